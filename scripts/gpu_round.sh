#!/bin/bash
# One GPU session (run under gpurun): parity tests, the bench line, the ncu launch list and --set full captures of the EM kernels.
# Everything lands in gpurun_out/.  usage: scripts/gpu_round.sh [tag]
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
t0=$(date +%s)
# 1. the gather-form EM loop on its own (all failures, not just the first)
timeout 400 python -m pytest tests/test_gpu_em_gather.py -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_gather.log 2>&1
G=$?
echo "gather tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_gather.log
if [ $G -ne 0 ]; then
  # where does it go wrong: memcheck on one small case
  SFB200_EM_GATHER=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_em_gather.py -q --tb=line -p no:cacheprovider \
      -k "fixed_iterations and 2-0" > $OUT/${TAG}_memcheck.log 2>&1
  tail -30 $OUT/${TAG}_memcheck.log
else
  export SFB200_EM_GATHER=1
fi
# 2. the whole GPU suite (with the gather loop on when it passed)
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_em_gather.py > $OUT/${TAG}_t_all.log 2>&1
echo "gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -8 $OUT/${TAG}_t_all.log
# 3. bench line
SFB200_VERBOSE=1 timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_bench.json
# 4. ncu: launch list of the bench command, then full captures of the EM kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_em_gather|k_gather_build|k_em_part' -c 3 -f -o $OUT/${TAG}_em \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
ls -la $OUT | tail -20
