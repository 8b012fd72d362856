#!/bin/bash
# The C++ driver from FASTQ text with the host parser (run under gpurun), parser phase times included
TAG=${1:-r02ze}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2; do
SFB200_PARSE_TIMING=1 timeout 600 python scripts/cli_e2e.py --reads 16000000 $([ $i = 2 ] && echo --reuse) > $OUT/${TAG}_cli_e2e_host_parse.json 2> $OUT/${TAG}_cli_e2e_host_parse.log
echo "cli e2e (host parser, 16M reads) rc=$?"; cat $OUT/${TAG}_cli_e2e_host_parse.json; grep "parser phases" $OUT/${TAG}_cli_e2e_host_parse.log | tail -2
done
