#!/bin/bash
# Short GPU session: gather-loop parity (scaled and plain indices), host timing marks of em_run, CTAs-per-SM sweep of the EM loop.
TAG=${1:-probe}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_em_gather.py tests/test_gpu_em.py -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
echo "em tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_em.log
SFB200_EM_GATHER_UNSCALED=1 timeout 400 python -m pytest tests/test_gpu_em_gather.py -q --tb=short -p no:cacheprovider -k "converges or fixed" > $OUT/${TAG}_t_unscaled.log 2>&1
echo "unscaled tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t_unscaled.log
SFB200_TIMING=1 SFB200_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench2.json 2> $OUT/${TAG}_bench2.log
echo "bench (2 CTAs/SM) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench2.json; tail -22 $OUT/${TAG}_bench2.log
for n in 4 1; do
  SFB200_EM_CTAS_PER_SM=$n timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench$n.json 2> $OUT/${TAG}_bench$n.log
  echo "bench ($n CTAs/SM) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench$n.json
done
