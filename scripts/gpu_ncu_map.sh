#!/bin/bash
# --set full of the mapping kernels on a batch in the middle of a step (class table populated): pack, scan, finalize, finalize (heavy pass)
TAG=${1:-r02zf}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads|k_pack_reads' --launch-skip 36 -c 4 -f -o $OUT/${TAG}_map \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_map.log 2>&1
echo "ncu map rc=$?"
