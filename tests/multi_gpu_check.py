"""Run under torchrun on N GPUs: every rank maps its shard of one read set; the merged-classes EM (default) and the
per-iteration all-reduce EM (SFB200_MULTI_EM_ALLREDUCE=1) must both reproduce the single-process oracle on all reads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200 import capi, synth, efflen  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = capi.Context(local)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(capi.Context.comm_unique_id()))
dist.broadcast(uid, 0)
ctx.comm_init(world, rank, uid.cpu().numpy())

seq, off, ln = synth.make_transcriptome(300, seed=3)
n = 200000
b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, 100, seed=8, paired=True, sub_rate=0.01)
ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
IU = 1 | (2 << 1) | (4 << 3)
lo, hi = rank * n // world, (rank + 1) * n // world
res = {}
for mode in ("merged", "allreduce"):
    if mode == "allreduce":
        os.environ["SFB200_MULTI_EM_ALLREDUCE"] = "1"
    else:
        os.environ.pop("SFB200_MULTI_EM_ALLREDUCE", None)
    ctx.map_begin(capi.MapOpts.default(IU))
    ctx.map_batch(b1, o1[lo:hi + 1], b2, o2[lo:hi + 1])
    g = ctx.map_finish()
    eff = efflen.effective_lengths(ln, g["fld"])
    for vb in (0, 1):
        a, it, mrd = ctx.em_run(eff, int(g["counters"][1]), capi.EMOpts.default(use_vb=vb))
        res[(mode, vb)] = (a, it, g["counters"].copy(), g["fld"].copy(), g["n_classes"])
    t = torch.from_numpy(res[(mode, 0)][0]).cuda()
    t0 = t.clone(); dist.broadcast(t0, 0)
    assert torch.allclose(t, t0, rtol=1e-9, atol=1e-9), "ranks disagree"
if rank == 0:
    from oracle import pyoracle as O
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    run = O.Run(O.Index(seqs, k=31), O.MapOpts.default(IU))
    run.map_batch(b1.tobytes(), o1, b2.tobytes(), o2, n_threads=8)
    w = run.finish()
    effo = efflen.effective_lengths(ln, w["fld"])
    for mode in ("merged", "allreduce"):
        a, it, counters, fld, ncls = res[(mode, 0)]
        assert counters.tolist() == w["counters"].tolist(), (mode, counters, w["counters"])
        assert fld.tolist() == w["fld"].tolist(), "fragment-length sample differs from the single-process one"
        if mode == "merged":
            assert ncls == len(w["counts"]), (ncls, len(w["counts"]))
        for vb in (0, 1):
            a, it, _, _, _ = res[(mode, vb)]
            rc, want, it_o, _ = O.em_run(len(ln), w["row_ptr"], w["labels"], w["counts"], effo, int(w["counters"][1]), O.EMOpts.default(use_vb=vb))
            assert it == it_o, (mode, vb, it, it_o)
            np.testing.assert_allclose(a, want, rtol=1e-4, atol=1e-6)
    print("multi-GPU check ok on %d ranks: merged and all-reduce EM both match the single-process oracle" % world)
dist.destroy_process_group()
