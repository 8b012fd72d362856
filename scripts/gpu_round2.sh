#!/bin/bash
# GPU session: full parity suite, bench line, C++ driver from FASTQ, ncu launch list + EM kernel capture.  usage: scripts/gpu_round2.sh [tag]
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_all.log 2>&1
echo "gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -8 $OUT/${TAG}_t_all.log
SFB200_TIMING=1 timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_bench.json; grep timing $OUT/${TAG}_bench.log | tail -12
timeout 600 python scripts/cli_e2e.py --reads 4000000 > $OUT/${TAG}_cli_e2e.json 2> $OUT/${TAG}_cli_e2e.log
echo "cli e2e rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e.json; tail -5 $OUT/${TAG}_cli_e2e.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_em_gather|k_gather_build|k_part_bounds' -c 3 -f -o $OUT/${TAG}_em \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
