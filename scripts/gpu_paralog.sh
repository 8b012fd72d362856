#!/bin/bash
# Mapping on the paralog / repeat set (run under gpurun): the mapping parity tests, then bench.py --structure paralog with the heavy
# finalize pass on (default) and off, the iid bench line (the heavy pass must cost nothing there), and a --set full capture of the
# scan / finalize kernels on the paralog set.
# usage: /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_paralog.sh <tag>'
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_map.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_map.log 2>&1
G=$?
echo "map tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_map.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t_map.log | head -20 | cut -c1-300; exit 1; fi
run() {
    label=$1; shift
    env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-realistic --structure paralog --reads 4000000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('$label: map kernels %.2f ms per 4 M reads (%.0f M reads/s), EM %s %.2f ms, step %.2f ms, launches %d' % (x['map_kernel_ms_per_step'], x['map_kernel_reads_per_s']/1e6, x['em_kernel'], x['em_loop_ms_per_step'], d['ms_per_step'], d['gpu_launches']))"
}
{
run "heavy pass + extension words (default)" SFB200_X=0
run "one finalize pass" SFB200_NO_HEAVY_PASS=1
run "heavy pass, no extension words" SFB200_IVPOOL_WORDS=1
} 2>&1 | tee $OUT/${TAG}_paralog_ab.txt
echo "paralog A/B done ($(( $(date +%s) - t0 )) s)"
timeout 600 python bench.py --no-cpu-baseline --no-realistic > $OUT/${TAG}_bench_iid.json 2> $OUT/${TAG}_bench_iid.log
echo "iid bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_iid.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads' --launch-skip 12 -c 3 -f -o $OUT/${TAG}_paralog \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic --structure paralog --reads 4000000 > $OUT/${TAG}_ncu_paralog.log 2>&1
echo "ncu paralog rc=$?  ($(( $(date +%s) - t0 )) s)"
