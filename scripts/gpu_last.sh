#!/bin/bash
# Last GPU session: host-batch ramp-up (mapping parity through sfb200_map_batch) and the bench line.
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
t0=$(date +%s)
SFB200_VERBOSE=1 timeout 140 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_bench.json
timeout 100 python -m pytest tests/test_gpu_map.py tests/test_quant_cli.py -q --tb=short -p no:cacheprovider -x > $OUT/${TAG}_t_map.log 2>&1
echo "map tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -4 $OUT/${TAG}_t_map.log | cut -c1-300
