#!/bin/bash
# probes of the hybrid EM's pool loop on the paralog set + the class-table growth tests
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_em_gather.py -m gpu -x -q -k "hybrid or dense" --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
echo "map tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -4 $OUT/${TAG}_t_em.log | cut -c1-300; grep -E "^E " $OUT/${TAG}_t_em.log | head -10 | cut -c1-300
run() { # label, env...
  local label=$1; shift
  env "$@" SFB200_VERBOSE=1 timeout 300 python bench.py --steps 1 --no-cpu-baseline --structure paralog --reads 4000000 --no-realistic 2> $OUT/${TAG}_probe.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label:', d['detail']['em_kernel'], round(d['detail']['em_loop_ms_per_step'],2), 'ms per 1000 iterations')"
  grep -E "EM pool|EM partition" $OUT/${TAG}_probe.log | tail -2 | cut -c1-250
}
run "default" A=1
run "pool ctas 16" SFB200_EM_POOL_CTAS=16
run "pool ctas 32" SFB200_EM_POOL_CTAS=32
run "pool ctas 130" SFB200_EM_POOL_CTAS=130
run "beta in global" SFB200_EM_POOL_GLOBAL=1
run "hybrid off" SFB200_EM_HYBRID=0
echo "done ($(( $(date +%s) - t0 )) s)"
