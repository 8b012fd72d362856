#!/bin/bash
# Quick validation (run under gpurun): the whole GPU suite, the bench line, the launch list of the bench command.
# usage: /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_quick.sh <tag>'
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_all.log 2>&1
echo "gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t_all.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
grep -E "k_part_bounds|k_eq_fill" $OUT/${TAG}_launches.csv | tail -4 | cut -c1-200
