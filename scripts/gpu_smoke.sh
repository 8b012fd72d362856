#!/bin/bash
# smoke() + the EM / sampler GPU tests (a last look at a rebuilt library)
TAG=${1:-r02zh}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_em.py tests/test_sampler_pins.py -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t.log 2>&1
echo "em + sampler tests rc=$?"; tail -2 $OUT/${TAG}_t.log | cut -c1-200
