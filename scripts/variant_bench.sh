#!/bin/bash
# usage: scripts/variant_bench.sh lib1.so lib2.so ...   (map-kernel time of bench.py cfg2 for each library build)
for lib in "$@"; do
  SFB200_LIB=$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e6,1), 'Mreads/s map_ms', round(d['detail']['map_kernel_ms_per_step'],2), 'em_ms', round(d['detail']['em_loop_ms_per_step'],2))"
done
