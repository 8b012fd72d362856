#!/bin/bash
# GPU session: parity of the gather / dense EM loops, then the bench line with the dense loop on.
TAG=${1:-dense}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 500 python -m pytest tests/test_gpu_em_gather.py -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t.log 2>&1
G=$?
echo "gather/dense tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -25 $OUT/${TAG}_t.log | cut -c1-300
if [ $G -ne 0 ]; then
  SFB200_EM_DENSE=1 SFB200_EM_GATHER=1 timeout 200 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_em_gather.py -q --tb=line -p no:cacheprovider \
      -k "fixed_iterations and dense-2-0" > $OUT/${TAG}_memcheck.log 2>&1
  tail -30 $OUT/${TAG}_memcheck.log | cut -c1-300
fi
for grp in 4 1 2; do
  SFB200_EM_DENSE=1 SFB200_EM_DENSE_GROUP=$grp SFB200_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_g$grp.json 2> $OUT/${TAG}_bench_g$grp.log
  echo "bench (dense, $grp lanes per component) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_g$grp.json; grep -E "dense" $OUT/${TAG}_bench_g$grp.log | tail -1
done
