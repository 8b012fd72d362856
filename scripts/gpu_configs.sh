#!/bin/bash
# The other BASELINE configurations at their stated sizes on one B200 (bench.py --config N); everything lands in gpurun_out/.
# usage: /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_configs.sh <tag> [3] [5]'
TAG=${1:-r02i}; shift
OUT=gpurun_out
mkdir -p $OUT
t0=$(date +%s)
free -g | head -2; nvidia-smi --query-gpu=memory.total,memory.used --format=csv,noheader
for cfg in "$@"; do
  extra=""
  [ "$cfg" = "5" ] && extra="--no-cpu-baseline"      # the CPU arm would download a 34 GB k-mer table
  SFB200_VERBOSE=1 timeout 1500 python bench.py --config $cfg --steps 1 $extra > $OUT/${TAG}_bench_cfg$cfg.json 2> $OUT/${TAG}_bench_cfg$cfg.log
  echo "cfg$cfg rc=$?  ($(( $(date +%s) - t0 )) s)"; python scripts/show_bench.py $OUT/${TAG}_bench_cfg$cfg.json; grep -E "^\[bench\]|Error|error" $OUT/${TAG}_bench_cfg$cfg.log | tail -8 | cut -c1-400
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
