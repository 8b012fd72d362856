#!/bin/bash
# the GPU tests of the two drivers (C++ sfb200-quant, python -m sailfish_b200.quant) only
TAG=${1:-r02zg}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_host_quant_cli.py tests/test_quant_cli.py tests/test_host_adaptors.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/${TAG}_t_host.log 2>&1
echo "driver gpu tests rc=$?"; tail -3 $OUT/${TAG}_t_host.log | cut -c1-300
