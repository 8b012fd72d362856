"""Times the device-side bias / GC effective-length correction (sfb200_bias_eff_lens; default kernels and the sliding GC form) at a
BASELINE-config-2-like size.  usage (GPU box): python scripts/bench_bias.py [--genes 40000] [--gc-samp 1]
Prints one JSON line.  Not part of bench.py: bias correction is off by default in the reference.  Product code
only (nothing under oracle/ is used): parity is the tests' business (tests/test_gpu_bias.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200 import capi, efflen, synth          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=40000)
    ap.add_argument("--gc-samp", type=int, default=1)
    a = ap.parse_args()
    seq, off, ln = synth.make_transcriptome(a.genes, seed=1)
    T = len(ln)
    ctx = capi.Context(0)
    t0 = time.time()
    ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    t_index = time.time() - t0
    rng = np.random.default_rng(3)
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    cdf, mx = efflen.empirical_cdf(fld)
    eff = np.where(ln - 189.0 >= 1, ln - 189.0, ln).astype(np.float64)
    alphas = rng.lognormal(3, 2, size=T); alphas[rng.random(T) < 0.2] = 0.0
    rb = rng.integers(1, 3000, size=4096).astype(np.uint32); og = rng.integers(1, 8000, size=101).astype(np.uint32)
    out = {"transcripts": int(T), "bases": int(ln.sum()), "index_s": round(t_index, 3)}
    for mode, name in ((1, "seq"), (2, "gc"), (2, "gc_slide")):
        if name == "gc_slide":
            os.environ["SFB200_BIAS_GC_SLIDE"] = "1"          # read by the library at every call
        ctx.bias_eff_lens(mode, eff, eff, alphas, 600000, 590000, rb, og, cdf, mx, gc_samp=a.gc_samp)      # warm-up
        t0 = time.time()
        got = ctx.bias_eff_lens(mode, eff, eff, alphas, 600000, 590000, rb, og, cdf, mx, gc_samp=a.gc_samp)
        dt = time.time() - t0
        out[name + "_gpu_s"] = round(dt, 4)
        out[name + "_gpu_Mpos_per_s"] = round(float(ln.sum()) / dt / 1e6, 1)
        out[name + "_changed"] = int((got != eff).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
