#!/bin/bash
# EM session 3: hybrid with whole components per pool CTA (parity + paralog timing), bootstrap timing marks
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_map.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_map.log 2>&1
echo "map tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -4 $OUT/${TAG}_t_map.log | cut -c1-300; grep -E "^E " $OUT/${TAG}_t_map.log | head -8 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_em_gather.py -m gpu -x -q --tb=short -p no:cacheprovider -k "hybrid or streaming or dense_loop" > $OUT/${TAG}_t_em.log 2>&1
G=$?
echo "em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_em.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t_em.log | head -30 | cut -c1-300; fi
run() { local label=$1; shift
  env "$@" SFB200_VERBOSE=1 timeout 300 python bench.py --steps 2 --no-cpu-baseline --structure paralog --reads 4000000 --no-realistic 2> $OUT/${TAG}_probe.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label:', d['detail']['em_kernel'], round(d['detail']['em_loop_ms_per_step'],2), 'ms per 1000 iterations; step', round(d['ms_per_step'],2), 'ms; host', d['detail']['host_wall_ms_per_step'])"
  grep -E "EM pool|EM partition" $OUT/${TAG}_probe.log | tail -2 | cut -c1-250
}
run "hybrid" SFB200_EM_HYBRID=1
run "hybrid, 8 pool CTAs" SFB200_EM_HYBRID=1 SFB200_EM_POOL_CTAS=8
run "hybrid off" SFB200_EM_HYBRID=0
SFB200_TIMING=1 timeout 600 python bench.py --config 3 --reads 5000000 --steps 1 --bootstraps 6 --gibbs 0 --no-cpu-baseline > $OUT/${TAG}_bench_cfg3_t.json 2> $OUT/${TAG}_bench_cfg3_t.log
echo "cfg3 timing rc=$?  ($(( $(date +%s) - t0 )) s)"; grep -E "bootstrap [45]:|em: " $OUT/${TAG}_bench_cfg3_t.log | tail -24
