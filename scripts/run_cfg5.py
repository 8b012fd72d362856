"""BASELINE config 5 shape at reduced size: a metatranscriptome-scale index (default 100k genes x 5 = 500k transcripts, ~0.86 Gnt;
the full config is 1M transcripts) and paired-end 2x150 reads.  Reports the index footprint in HBM, the class-table load and the
throughput, plus the invariants that do not need the oracle at size."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200 import capi, synth, efflen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=100000)
ap.add_argument("--pairs", type=int, default=4_000_000)
args = ap.parse_args()

t0 = time.time()
seq, off, ln = synth.make_transcriptome(args.genes, seed=43)
print("transcriptome: %d transcripts, %.2f Gnt (%.1fs)" % (len(ln), seq.size / 1e9, time.time() - t0), flush=True)
ctx = capi.Context(0)
t0 = time.time()
st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
print("index build %.1fs: %d positions, %d k-mers, %d table slots (load %.2f), %.1f GB in HBM, max bucket %d" % (
    time.time() - t0, st["n_sa"], st["n_kmers"], st["table_slots"], st["n_kmers"] / st["table_slots"], st["hbm_bytes"] / 1e9, st["max_bucket"]), flush=True)
IU = 1 | (2 << 1) | (4 << 3)
ctx.map_begin(capi.MapOpts.default(IU))
done = 0; c = 0; t_map = 0.0
while done < args.pairs:
    n = min(1_000_000, args.pairs - done)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, 150, seed=1237, paired=True, frag_mean=300.0, frag_sd=40.0, expr_seed=1237, stream=c)
    t = time.time(); ctx.map_batch(b1, o1, b2, o2); ctx.sync(); t_map += time.time() - t
    done += n; c += 1
t = time.time(); g = ctx.map_finish(); t_map += time.time() - t
cnt = g["counters"]
print("mapped %d of %d pairs (%.2f%%); %d classes, nnz %d; class table load %.3f of %d slots; mapping kernels %.1f ms = %.1f M pairs/s" % (
    cnt[1], cnt[0], 100.0 * cnt[1] / cnt[0], g["n_classes"], g["nnz"], g["n_classes"] / float((1 << 21) * 4), (1 << 21) * 4,
    ctx.last_map_kernel_ms(), args.pairs / ctx.last_map_kernel_ms() / 1e3), flush=True)
eff = efflen.effective_lengths(ln, g["fld"])
nm = int(cnt[1])
os.environ["SFB200_VERBOSE"] = "1"
t = time.time()
a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(fixed_iters=1000))
print("EM 1000 iterations: loop %.2f ms (%.1f us/iteration), call %.1f ms, sum ok %s" % (
    ctx.last_em_loop_ms(), ctx.last_em_loop_ms(), 1e3 * (time.time() - t), bool(abs(a.sum() - nm) < 1e-6 * nm)))
