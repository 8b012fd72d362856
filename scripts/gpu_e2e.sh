#!/bin/bash
# End-to-end path session (run under gpurun): mapping + EM parity tests, then the bench line with different host piece sizes and with
# the offsets always copied (A/B of sfb200_map_batch's pipelining), and the cfg3-shaped run (VBEM + bootstraps) for the VBEM iteration time.
# usage: /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_e2e.sh <tag>'
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_em.py tests/test_gpu_em_gather.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t.log 2>&1
G=$?
echo "map + em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t.log | head -20 | cut -c1-300; exit 1; fi
run() {
    label=$1; shift
    env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-realistic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('$label: e2e %.1f M reads/s (%.2f ms per step, %.1f MB H2D), resident %.1f M reads/s (%.2f ms), map kernels %.2f ms, launches %d' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6, d['value']/1e6, d['ms_per_step'], x['map_kernel_ms_per_step'], d['gpu_launches']))"
}
{
run "default (1 M pieces, fixed-length entry point)" SFB200_X=0
run "512 k pieces" SFB200_HOST_PIECE=524288
run "whole batches (2.5 M)" SFB200_HOST_PIECE=4000000
} 2>&1 | tee $OUT/${TAG}_e2e_ab.txt
echo "e2e A/B done ($(( $(date +%s) - t0 )) s)"
SFB200_TIMING=1 timeout 600 python bench.py --config 3 --reads 5000000 --bootstraps 6 --gibbs 0 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg3_5M.json 2> $OUT/${TAG}_bench_cfg3_5M.log
echo "cfg3 (5 M pairs, 6 bootstraps) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg3_5M.json | cut -c1-600
grep "bootstrap [0-5]: em" $OUT/${TAG}_bench_cfg3_5M.log | tail -6
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
timeout 600 python bench.py --bootstraps 100 --steps 2 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_cfg2_boot100.json 2>/dev/null
emline $OUT/${TAG}_cfg2_boot100.json "cfg2 + 100 bootstraps (EM, 1000 fixed iterations each)"
} 2>&1 | tee $OUT/${TAG}_em_lag.txt
echo "EM lag A/B done ($(( $(date +%s) - t0 )) s)"
