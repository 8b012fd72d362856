#!/bin/bash
# GPU session: EM / partition parity after a change, then a sweep of the presence-filter size (SFB200_BLOOM_LOG2_WORDS).
TAG=${1:-sweep}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_em_gather.py tests/test_gpu_em.py -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
echo "em tests rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t_em.log
for lw in 23 22 21 24; do
  SFB200_BLOOM_LOG2_WORDS=$lw timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bloom$lw.json 2> $OUT/${TAG}_bloom$lw.log
  echo "bloom 2^$lw words rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bloom$lw.json
done
SFB200_BLOOM_LOG2_WORDS=22 SFB200_NO_L2_PERSIST=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bloom22np.json 2> $OUT/${TAG}_bloom22np.log
echo "bloom 2^22 words, no persisting window rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bloom22np.json
