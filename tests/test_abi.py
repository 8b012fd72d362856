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


def test_host_binding_is_a_no_op_without_a_device():
    """sfb200_bind_host_near_device: no such device (or no NUMA information) leaves the affinity alone and says so with 0 -- saying
    "no usable device" loudly is sfb200_ctx_create's job"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    before = os.sched_getaffinity(0)
    assert capi.bind_host_near_device(0) == 0 and capi.bind_host_near_device(7) == 0
    assert os.sched_getaffinity(0) == before


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


def test_new_wrappers_marshal_their_arguments():
    """capi's wrappers for the entry points that have not run on a GPU yet (map_fastq, map_set_bias / map_get_bias, em_run_bias,
    bias_eff_lens) called against ctypes callbacks with the SAME argtypes as the C functions: every argument must convert, arrive
    with the right value and the outputs must come back -- a marshalling slip would otherwise surface only on the GPU box."""
    import ctypes as C

    import numpy as np

    from sailfish_b200 import capi

    seen = {}

    class FakeLib:
        pass

    def proto(name, fn):
        restype, argtypes = capi.SIGNATURES[name]
        cb = C.CFUNCTYPE(restype, *argtypes)(fn)
        setattr(FakeLib, name, cb)

    def map_fastq(h, t1, n1, t2, n2, mx, n_rec, c1, c2):
        seen["map_fastq"] = (t1[:4] if t1 else None, n1, t2[:4] if t2 else None, n2, mx, bool(c2))
        n_rec[0] = 7; c1[0] = 11
        if c2:
            c2[0] = 13
        return 0

    def map_set_bias(h, s, g, n):
        seen["map_set_bias"] = (s, g, n)
        return 0

    def map_get_bias(h, rb, og):
        for i in range(4096):
            rb[i] = i + 1
        for i in range(101):
            og[i] = 2 * i
        return 0

    def em_run_bias(h, eff, n, nm, opts, model, alphas, eff_out, iters, mrd):
        m = model.contents
        seen["em_run_bias"] = (n, nm, opts.contents.use_vb, opts.contents.fixed_iters, m.mode, m.gc_samp, m.num_fwd, m.num_rc, m.read_bias[4095], m.observed_gc[100],
                               m.n_cdf, round(m.fld_cdf[2], 3), m.fld_max, eff[1])
        for i in range(n):
            alphas[i] = 10.0 * i; eff_out[i] = eff[i] + 1.0
        iters[0] = 123; mrd[0] = 0.25
        return 0

    def bias_eff_lens(h, model, eff_model, eff_in, alphas, n, out):
        seen["bias_eff_lens"] = (model.contents.mode, n, eff_model[0], eff_in[1], alphas[2])
        for i in range(n):
            out[i] = eff_in[i] * 2
        return 0

    for name, fn in (("sfb200_map_fastq", map_fastq), ("sfb200_map_set_bias", map_set_bias), ("sfb200_map_get_bias", map_get_bias),
                     ("sfb200_em_run_bias", em_run_bias), ("sfb200_bias_eff_lens", bias_eff_lens)):
        proto(name, fn)
    ctx = object.__new__(capi.Context)
    ctx.L = FakeLib; ctx.h = None                       # h None: close() / __del__ do nothing
    assert ctx.map_fastq(b"@r0\nAC\n+\nII\n", b"@r0\nGT\n+\nII\n", max_records=5) == (7, 11, 13)
    assert seen["map_fastq"] == (b"@r0\n", 12, b"@r0\n", 12, 5, True)
    assert ctx.map_fastq(b"@r0\nAC\n+\nII\n") == (7, 11, 0) and seen["map_fastq"][2:] == (None, 0, 0, False)
    ctx.map_set_bias(True, False, 1234)
    assert seen["map_set_bias"] == (1, 0, 1234)
    rb, og = ctx.map_get_bias()
    assert rb.dtype == np.uint32 and rb[0] == 1 and rb[4095] == 4096 and og[100] == 200
    eff = np.array([100.0, 200.0, 300.0])
    cdf = np.array([0.0, 0.25, 0.5, 1.0], np.float32)
    a, eo, it, mrd = ctx.em_run_bias(2, eff, 999, 60, 40, rb, og, cdf, 999, gc_samp=3, opts=capi.EMOpts.default(use_vb=1, fixed_iters=55))
    assert seen["em_run_bias"] == (3, 999, 1, 55, 2, 3, 60, 40, 4096, 200, 4, 0.5, 999, 200.0)
    assert a.tolist() == [0.0, 10.0, 20.0] and eo.tolist() == [101.0, 201.0, 301.0] and it == 123 and mrd == 0.25
    out = ctx.bias_eff_lens(1, eff, eff + 5, np.array([1.0, 2.0, 3.0]), 60, 40, rb, og, cdf, 999)
    assert seen["bias_eff_lens"] == (1, 3, 100.0, 205.0, 3.0) and out.tolist() == [210.0, 410.0, 610.0]
