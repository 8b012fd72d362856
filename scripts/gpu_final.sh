#!/bin/bash
# Round-end evidence on one B200: the whole GPU suite, smoke(), the bench line of both arms, the ncu launch list of the bench command,
# --set full captures of the mapping and EM kernels.  Everything lands in gpurun_out/.
# usage: /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > $OUT/${TAG}_t_all.log 2>&1
echo "gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -6 $OUT/${TAG}_t_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -1 $OUT/${TAG}_smoke.log
SFB200_VERBOSE=1 timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
for k in ("parity","realistic","roofline","em_roofline","cpu_baseline","clocks","gpu_launches"):
    print(k+":", json.dumps(d.get(k))[:600])
PY
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.log
echo "reference arm rc=$?  ($(( $(date +%s) - t0 )) s)"; cut -c1-400 $OUT/${TAG}_bench_reference.json; tail -2 $OUT/${TAG}_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads|k_pack_reads' --launch-skip 32 -c 4 -f -o $OUT/${TAG}_map \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_map.log 2>&1
echo "ncu map rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_em_dense|k_em_part' -c 1 -f -o $OUT/${TAG}_em \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_reads|k_finalize_reads' --launch-skip 12 -c 3 -f -o $OUT/${TAG}_paralog \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic --structure paralog --reads 4000000 > $OUT/${TAG}_ncu_paralog.log 2>&1
echo "ncu paralog rc=$?  ($(( $(date +%s) - t0 )) s)"
