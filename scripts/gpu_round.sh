#!/bin/bash
# One GPU session (run under gpurun, ~3 min on the box): parity of the EM loops, the whole GPU suite, the bench line, the C++ driver
# from FASTQ text, the ncu launch list of the bench command and a --set full capture of the EM kernels.  Everything lands in gpurun_out/.
# usage: /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r01d}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_em_gather.py -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
G=$?
echo "gather/dense tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -4 $OUT/${TAG}_t_em.log | cut -c1-300
if [ $G -eq 0 ]; then export SFB200_EM_DENSE=1; fi
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_em_gather.py > $OUT/${TAG}_t_all.log 2>&1
echo "gpu suite (SFB200_EM_DENSE=$SFB200_EM_DENSE) rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -4 $OUT/${TAG}_t_all.log | cut -c1-300
SFB200_VERBOSE=1 timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_bench.json; grep -E "dense layout" $OUT/${TAG}_bench.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_em_dense|k_dense_build|k_em_gather' -c 2 -f -o $OUT/${TAG}_em \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 python scripts/cli_e2e.py --reads 4000000 > $OUT/${TAG}_cli_e2e.json 2> $OUT/${TAG}_cli_e2e.log
echo "cli e2e rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e.json
SFB200_EM_DENSE=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_gather.json 2> $OUT/${TAG}_bench_gather.log
echo "bench (gather loop) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_gather.json
