#!/bin/bash
# EM session 4 (run under gpurun): EM parity tests, the cfg2 bench line (fixed-count EM: the headline), EM to convergence / VBEM with and
# without the lagged rule, BASELINE config 3 at full size.
# usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_em4.sh <tag> [cfg3]'
TAG=${1:-r02w}
CFG3=${2:-}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_em.py tests/test_gpu_em_gather.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t.log 2>&1
G=$?
echo "em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error|assert" $OUT/${TAG}_t.log | head -30 | cut -c1-300; exit 1; fi
timeout 600 python bench.py --no-cpu-baseline --no-realistic > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json
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
} 2>&1 | tee $OUT/${TAG}_em_lag.txt
echo "EM lag A/B done ($(( $(date +%s) - t0 )) s)"
if [ -n "$CFG3" ]; then
    SFB200_VERBOSE=1 timeout 1200 python bench.py --config 3 --steps 1 > $OUT/${TAG}_bench_cfg3.json 2> $OUT/${TAG}_bench_cfg3.log
    echo "cfg3 at full size rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg3.json | cut -c1-700
fi
