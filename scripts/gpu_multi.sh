#!/bin/bash
# Multi-GPU session on ONE box with N GPUs: the 2-rank parity test, then the bench line at every N in the list, launched exactly as
# the driver does (torchrun for N > 1), optionally the reference arm under torchrun and config 4 at the largest N.
# usage: /usr/local/graft/bin/gpurun --gpus 8 --timeout 1800 -- '[SFB200_MULTI_AB=1] bash scripts/gpu_multi.sh <tag> "1 2 4 8" [cfg4]'
# (SFB200_MULTI_AB=1 adds the reference arm under torchrun and the largest N without NUMA binding)
TAG=${1:-r02n}
NS=${2:-"1 2"}
CFG4=${3:-}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_BENCH_CACHE=/dev/shm/sfb200_cache
t0=$(date +%s)
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -p no:cacheprovider -rs > $OUT/${TAG}_t_multi.log 2>&1
echo "multirank test rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_t_multi.log | cut -c1-400
last=1
for N in $NS; do
    last=$N
    if [ "$N" = "1" ]; then
        timeout 900 python bench.py --gpus 1 --no-realistic > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.log
    else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
            bench.py --gpus $N --no-realistic > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.log
    fi
    echo "bench N=$N rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_n$N.json 2>&1 | tail -4
done
if [ "$last" != "1" ] && [ -n "$SFB200_MULTI_AB" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port 29611 \
        bench.py --impl reference --gpus $last --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_n$last.json 2> $OUT/${TAG}_bench_reference_n$last.log
    echo "reference arm under torchrun N=$last rc=$?  ($(( $(date +%s) - t0 )) s)"; grep '^{' $OUT/${TAG}_bench_reference_n$last.json | cut -c1-300
fi
if [ "$last" != "1" ] && [ -n "$SFB200_MULTI_AB" ]; then
    # the same largest N with the page-locked buffers wherever the kernel puts them (no NUMA binding): what the binding buys
    SFB200_NO_BIND=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port 29622 \
        bench.py --gpus $last --no-realistic > $OUT/${TAG}_bench_n${last}_nobind.json 2> $OUT/${TAG}_bench_n${last}_nobind.log
    echo "bench N=$last, SFB200_NO_BIND=1 rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_n${last}_nobind.json 2>&1 | tail -4
fi
free -g | head -2; lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8; nvidia-smi topo -m 2>/dev/null | head -14
avail=$(free -g | awk '/^Mem:/ {print $7}')
if [ -n "$CFG4" ] && [ "$avail" -lt 400 ]; then echo "config 4 skipped: only $avail GB of host memory available"; CFG4=; fi
if [ -n "$CFG4" ]; then
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port 29633 \
        bench.py --config 4 --gpus $last --steps 3 --warmup 3 --no-realistic > $OUT/${TAG}_bench_cfg4_n$last.json 2> $OUT/${TAG}_bench_cfg4_n$last.log
    echo "config 4 N=$last rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg4_n$last.json 2>&1 | tail -6
    tail -5 $OUT/${TAG}_bench_cfg4_n$last.log | cut -c1-300
fi
