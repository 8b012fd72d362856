#!/bin/bash
# Mapping-kernel session (run under gpurun): the mapping parity tests first (stop early when they fail), the bench line, variant
# builds of the mapping kernels (sailfish_b200/variants/, `make variant`), the ncu launch list of the bench command and a --set full
# capture of the scan / finalize kernels.  Everything lands in gpurun_out/.
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
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.log | cut -c1-400
for lib in sailfish_b200/variants/libsfb200_*.so; do
    [ -f "$lib" ] || continue
    SFB200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-realistic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e6,1), 'Mreads/s map_ms', round(d['detail']['map_kernel_ms_per_step'],2), 'em_ms', round(d['detail']['em_loop_ms_per_step'],2))"
done 2>&1 | tee $OUT/${TAG}_variants.txt
echo "variants done ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads|k_pack_reads' --launch-skip 30 -c 3 -f -o $OUT/${TAG}_map \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_map.log 2>&1
echo "ncu map rc=$?  ($(( $(date +%s) - t0 )) s)"
