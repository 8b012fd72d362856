#!/bin/bash
# Mapping-kernel session (run under gpurun): the mapping parity tests first (stop early when they fail), the bench line, the ncu launch
# list of the bench command and a --set full capture of the scan / finalize kernels.  Everything lands in gpurun_out/.
# usage: /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_map_round.sh <tag> [full]'
TAG=${1:-r02b}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_map.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_map.log 2>&1
G=$?
echo "map tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_map.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t_map.log | head -20 | cut -c1-300; exit 1; fi
if [ -n "$FULL" ]; then
    timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_map.py > $OUT/${TAG}_t_all.log 2>&1
    echo "rest of the gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_all.log | cut -c1-300
fi
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads|k_pack_reads' --launch-skip 30 -c 3 -f -o $OUT/${TAG}_map \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_map.log 2>&1
echo "ncu map rc=$?  ($(( $(date +%s) - t0 )) s)"
