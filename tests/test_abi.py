"""CPU tests: the C-ABI library builds, loads and exports every symbol include/sfb200.h declares; calls fail loudly
without a CUDA device (no CPU fallback); the product never touches oracle/."""
import os
import re

import pytest

from sailfish_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb200_[a-z0-9_]+)\s*\(", src)) - {"sfb200_f64_row_cb", "sfb200_i32_row_cb"})


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libsfb200.so does not export %s" % s
        assert s in capi.SIGNATURES, "capi.py has no signature for %s" % s
    assert L.sfb200_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.Sfb200Error) as e:
        capi.Context(0)
    assert e.value.code == -1          # SFB200_ENODEV


def test_product_does_not_reference_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "sailfish_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|pyoracle|liboracle|libsfref|#\s*include\s*[\"<][^\">]*oracle|orc_[a-z_]+\s*\(")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or fn == "Makefile":
                for line in open(os.path.join(dp, fn)).read().splitlines():
                    assert not bad.search(line), (fn, line)


def test_struct_layouts():
    import ctypes as C
    assert C.sizeof(capi.MapOpts) == 40
    assert C.sizeof(capi.EMOpts) == 56
