#!/bin/bash
# EM session (run under gpurun): the EM parity tests (gather / dense / hybrid pool loop), the whole suite, the bench line with the
# paralog pass, the L2 fetch-granularity switch, an ncu capture of the EM kernels on the paralog set.
# usage: /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_em_round.sh <tag>'
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_em_gather.py tests/test_gpu_em.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
G=$?
echo "em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_em.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t_em.log | head -30 | cut -c1-300; fi
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_em_gather.py --deselect tests/test_gpu_em.py > $OUT/${TAG}_t_all.log 2>&1
echo "rest of the gpu suite rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_all.log | cut -c1-300
SFB200_VERBOSE=1 timeout 900 python bench.py --steps 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench.json; grep -E "EM partition|dense layout" $OUT/${TAG}_bench.log | tail -4 | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
print("realistic:", json.dumps(d.get("realistic"))[:700]); print("parity:", d.get("parity")); print("clocks:", d.get("clocks"))
PY
for g in 32 64; do
  SFB200_L2_FETCH=$g timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-realistic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('L2 fetch $g:', round(d['value']/1e6,1), 'Mreads/s map_ms', round(d['detail']['map_kernel_ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1))"
done 2>&1 | tee $OUT/${TAG}_l2fetch.txt
SFB200_EM_HYBRID=0 timeout 300 python bench.py --steps 2 --no-cpu-baseline --structure paralog --reads 4000000 --no-realistic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('paralog, hybrid off:', d['detail']['em_kernel'], round(d['detail']['em_loop_ms_per_step'],2), 'ms')" | tee -a $OUT/${TAG}_l2fetch.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_em_dense' -c 2 -f -o $OUT/${TAG}_em \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-realistic --structure paralog --reads 4000000 > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu em rc=$?  ($(( $(date +%s) - t0 )) s)"
