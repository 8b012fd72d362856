import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200 import capi, synth, efflen
seq, off, ln = synth.make_transcriptome(4000, seed=42)
b1, o1, _, _, _ = synth.make_reads(seq, off, ln, 1000000, 76, seed=1234)
ctx = capi.Context(0)
ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
ctx.map_begin(capi.MapOpts.default((3 << 1) | (4 << 3)))
ctx.map_batch(b1, o1)
g = ctx.map_finish()
print(g["n_classes"], g["nnz"])
eff = efflen.effective_lengths(ln, None, single_end=True)
for it in (1000,):
    a, iters, _ = ctx.em_run(eff, int(g["counters"][1]), capi.EMOpts.default(fixed_iters=it))
    print("em loop ms", ctx.last_em_loop_ms(), iters)
a, iters, _ = ctx.em_run(eff, int(g["counters"][1]), capi.EMOpts.default())
print("em converge ms", ctx.last_em_loop_ms(), iters)
