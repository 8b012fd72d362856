#!/bin/bash
# EM session for the lagged stopping rule (run under gpurun): EM parity tests, then the converging / VBEM workloads with and without it
# usage: /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_lag.sh <tag>'
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_em.py tests/test_gpu_em_gather.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t.log 2>&1
G=$?
echo "em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error|assert" $OUT/${TAG}_t.log | head -30 | cut -c1-300; exit 1; fi
emline() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); x=d['detail']
print('$2: %s, %d iterations, %.2f us per iteration; step %.2f ms; host %s' % (x['em_kernel'], x['em_iters'], x['em_loop_ms_per_step']*1e3/max(x['em_iters'],1), d['ms_per_step'], x['host_wall_ms_per_step']))"; }
{
for lag in 0 1; do
    SFB200_EM_NO_LAG=$((1-lag)) timeout 600 python bench.py --config 4 --reads 5000000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_cfg4_lag$lag.json 2>/dev/null
    emline $OUT/${TAG}_cfg4_lag$lag.json "cfg4-shaped (5 M pairs, EM to convergence), lag=$lag"
    SFB200_EM_NO_LAG=$((1-lag)) timeout 600 python bench.py --config 3 --reads 5000000 --bootstraps 20 --gibbs 0 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_cfg3_lag$lag.json 2>/dev/null
    emline $OUT/${TAG}_cfg3_lag$lag.json "cfg3-shaped (5 M pairs, VBEM + 20 bootstraps), lag=$lag"
done
timeout 600 python bench.py --config 4 --reads 5000000 --em-iters 700 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_cfg4_fixed.json 2>/dev/null
emline $OUT/${TAG}_cfg4_fixed.json "cfg4-shaped, 700 fixed iterations (no stopping rule at all)"
} 2>&1 | tee $OUT/${TAG}_em_lag.txt
echo "EM lag A/B done ($(( $(date +%s) - t0 )) s)"
timeout 600 python -m pytest tests/test_gpu_map.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_map.log 2>&1
echo "map tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t_map.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-realistic --structure paralog --reads 4000000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('paralog set, heavy list in three bins: map kernels %.2f ms per 4 M reads (%.0f M reads/s), EM %s %.2f ms' % (x['map_kernel_ms_per_step'], x['map_kernel_reads_per_s']/1e6, x['em_kernel'], x['em_loop_ms_per_step']))" | tee $OUT/${TAG}_paralog.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_em_dense' -c 1 -f -o $OUT/${TAG}_em_lag \
    python bench.py --config 4 --reads 5000000 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
