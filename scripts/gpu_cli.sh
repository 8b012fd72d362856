#!/bin/bash
# The C++ driver from FASTQ text (run under gpurun): host parser and --deviceParse, 16 M reads; then BASELINE config 5 at full size.
# usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_cli.sh <tag> [cfg5]'
TAG=${1:-r02zb}
CFG5=${2:-}
OUT=gpurun_out
mkdir -p $OUT
t0=$(date +%s)
timeout 600 python scripts/cli_e2e.py --reads 16000000 > $OUT/${TAG}_cli_e2e_host_parse.json 2> $OUT/${TAG}_cli_e2e_host_parse.log
echo "cli e2e (host parser, 16M reads) rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e_host_parse.json
timeout 600 python scripts/cli_e2e.py --reads 16000000 --device-parse --reuse > $OUT/${TAG}_cli_e2e_device_parse.json 2> $OUT/${TAG}_cli_e2e_device_parse.log
echo "cli e2e (--deviceParse, 16M reads) rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e_device_parse.json
if [ -n "$CFG5" ]; then
    SFB200_VERBOSE=1 timeout 1200 python bench.py --config 5 --steps 1 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5.json 2> $OUT/${TAG}_bench_cfg5.log
    echo "cfg5 rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg5.json | cut -c1-600
fi
