"""BASELINE config 3 shape at a reduced read count: 200k-transcript index, paired-end 2x100 reads (-l IU), VBEM to
convergence, then bootstraps and Gibbs samples.  Prints timings and the invariants that do not need the oracle at size."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200 import capi, synth, efflen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=40000)
ap.add_argument("--pairs", type=int, default=5_000_000)
ap.add_argument("--boot", type=int, default=20)
ap.add_argument("--gibbs", type=int, default=20)
args = ap.parse_args()

t0 = time.time()
seq, off, ln = synth.make_transcriptome(args.genes, seed=42)
ctx = capi.Context(0)
st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
print("index: %d transcripts, %.2f GB (%.1fs)" % (len(ln), st["hbm_bytes"] / 1e9, time.time() - t0))
IU = 1 | (2 << 1) | (4 << 3)
chunk = 1_000_000
t_map = 0.0
ctx.map_begin(capi.MapOpts.default(IU))
done = 0
c = 0
while done < args.pairs:
    n = min(chunk, args.pairs - done)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, 100, seed=1235, paired=True, expr_seed=1235, stream=c)
    t = time.time()
    ctx.map_batch(b1, o1, b2, o2)
    ctx.sync()
    t_map += time.time() - t
    done += n
    c += 1
t = time.time()
g = ctx.map_finish()
t_map += time.time() - t
cnt = g["counters"]
print("mapped %d of %d pairs (%.2f%%), %d classes, nnz %d; mapping %.3fs = %.1f M pairs/s (kernel %.1f ms)" % (
    cnt[1], cnt[0], 100.0 * cnt[1] / cnt[0], g["n_classes"], g["nnz"], t_map, args.pairs / t_map / 1e6, ctx.last_map_kernel_ms()))
assert cnt[0] == args.pairs and g["fld"].sum() == 10000
eff = efflen.effective_lengths(ln, g["fld"])
nm = int(cnt[1])
for vb in (0, 1):
    t = time.time()
    a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
    print("%s: %d iterations, loop %.2f ms (%.1f us/iteration), call %.1f ms, sum %.3f, max rel diff %.4g" % (
        "VBEM" if vb else "EM", it, ctx.last_em_loop_ms(), 1e3 * ctx.last_em_loop_ms() / max(it, 1), 1e3 * (time.time() - t), a.sum(), mrd))
    if not vb:
        assert abs(a.sum() - nm) < 1e-6 * nm
        alphas = a
t = time.time()
rows = ctx.bootstrap_run(eff, args.boot, seed=5)
print("%d bootstraps: %.2f s (%.1f ms each); row sums ok: %s" % (args.boot, time.time() - t, 1e3 * (time.time() - t) / max(args.boot, 1),
                                                             bool(np.allclose(rows.sum(axis=1), nm, rtol=1e-9))))
t = time.time()
rows = ctx.gibbs_run(eff, alphas / alphas.sum(), nm, args.gibbs, seed=5)
print("%d Gibbs samples: %.2f s (%.1f ms each); row sums ok: %s" % (args.gibbs, time.time() - t, 1e3 * (time.time() - t) / max(args.gibbs, 1),
                                                                  bool((rows.sum(axis=1) == nm).all())))
big = alphas > 1000
print("Gibbs mean / EM estimate on %d large transcripts: median ratio %.4f" % (big.sum(), np.median(rows.mean(axis=0)[big] / alphas[big])))
