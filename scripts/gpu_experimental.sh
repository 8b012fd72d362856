#!/bin/bash
# First GPU session for the code written without a GPU (bias / GC correction end to end, balanced dense EM layout): the gated parity
# tests, then timings.  usage: /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_experimental.sh <tag>'
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
export SFB200_EXPERIMENTAL=1
t0=$(date +%s)
for f in tests/test_gpu_bias.py tests/test_gpu_map.py tests/test_gpu_em_gather.py tests/test_host_quant_cli.py; do
    n=$(basename $f .py)
    timeout 600 python -m pytest $f -m gpu -q --tb=short -p no:cacheprovider > $OUT/${TAG}_exp_$n.log 2>&1
    echo "$n rc=$?  ($(( $(date +%s) - t0 )) s)"; tail -3 $OUT/${TAG}_exp_$n.log | cut -c1-300
done
timeout 600 python scripts/bench_bias.py > $OUT/${TAG}_bench_bias.json 2> $OUT/${TAG}_bench_bias.log
echo "bench_bias rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_bench_bias.json
timeout 600 python scripts/cli_e2e.py --reads 4000000 > $OUT/${TAG}_cli_e2e_host_parse.json 2> $OUT/${TAG}_cli_e2e_host_parse.log
echo "cli e2e (host parser) rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e_host_parse.json
timeout 600 python scripts/cli_e2e.py --reads 4000000 --device-parse > $OUT/${TAG}_cli_e2e_device_parse.json 2> $OUT/${TAG}_cli_e2e_device_parse.log
echo "cli e2e (--deviceParse) rc=$?  ($(( $(date +%s) - t0 )) s)"; cat $OUT/${TAG}_cli_e2e_device_parse.json
SFB200_EM_DENSE_GROUP=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_dense0.json 2> $OUT/${TAG}_bench_dense0.log
echo "bench (balanced dense) rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_dense0.json
# ncu of the new kernels (one GPU, few launches): the bias / GC passes in both forms, and the device-side FASTQ extraction
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bias_expected|k_bias_efflen' -c 6 -f -o $OUT/${TAG}_bias \
    python scripts/bench_bias.py --genes 8000 > $OUT/${TAG}_ncu_bias.log 2>&1
echo "ncu bias rc=$?  ($(( $(date +%s) - t0 )) s)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fq_' -c 8 -f -o $OUT/${TAG}_fq \
    sailfish_b200/bin/sfb200-quant quant -t /dev/shm/sfb200_cli/t.fa -l U -r /dev/shm/sfb200_cli/r.fq -o /dev/shm/sfb200_cli/out_ncu --deviceParse > $OUT/${TAG}_ncu_fq.log 2>&1
echo "ncu fastq rc=$?  ($(( $(date +%s) - t0 )) s)"
