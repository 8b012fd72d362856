#!/bin/bash
# One-off probe (run under gpurun): --set full of k_part_bounds with source, then the bench line.
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_part_bounds' -c 1 -f -o $OUT/${TAG}_bounds \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_bounds.log 2>&1
echo "ncu bounds rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 python bench.py --no-cpu-baseline --no-realistic > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json
