"""sailfish_b200 -- B200 (sm_100a) implementation of Sailfish's quantification hot path.

csrc/   CUDA kernels + the C ABI (libsfb200.so, declared in include/sfb200.h)
host/   C++ adaptors with the reference's class names on top of the C ABI
capi    ctypes binding used by tests/, bench.py and __graft_entry__
synth   deterministic synthetic transcriptomes / reads (BASELINE.json configs)
"""
from . import capi  # noqa: F401
