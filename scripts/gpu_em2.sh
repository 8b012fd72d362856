#!/bin/bash
# EM session 2: EM parity tests (incl. the streaming dense variant), cfg3 at 5M pairs (VBEM bootstraps with the cheaper digamma), cfg5 at
# full size (streaming dense EM at 1 M transcripts)
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_em_gather.py tests/test_gpu_em.py tests/test_sampler_pins.py tests/test_gpu_bias.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_em.log 2>&1
G=$?
echo "em tests rc=$G  ($(( $(date +%s) - t0 )) s)"; tail -5 $OUT/${TAG}_t_em.log | cut -c1-300
if [ $G -ne 0 ]; then grep -E "^E |Error|error" $OUT/${TAG}_t_em.log | head -30 | cut -c1-300; fi
SFB200_VERBOSE=1 timeout 900 python bench.py --config 3 --reads 5000000 --steps 1 --no-cpu-baseline > $OUT/${TAG}_bench_cfg3_5M.json 2> $OUT/${TAG}_bench_cfg3_5M.log
echo "cfg3 (5M pairs) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg3_5M.json
SFB200_VERBOSE=1 timeout 1500 python bench.py --config 5 --steps 1 --no-cpu-baseline > $OUT/${TAG}_bench_cfg5.json 2> $OUT/${TAG}_bench_cfg5.log
echo "cfg5 rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg5.json; grep -E "EM partition|dense layout|Error|error" $OUT/${TAG}_bench_cfg5.log | tail -4 | cut -c1-300
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_cfg3_5M.json", "$OUT/${TAG}_bench_cfg5.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, {k:d["detail"].get(k) for k in ("em_kernel","em_iters","em_loop_ms_per_step","boot_loop_ms")}, d["em_roofline"].get("us_per_iter"), d["em_roofline"].get("frac"))
    except Exception as e: print(f, e)
PY
