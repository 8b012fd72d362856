"""N > 1 host logic on CPU (gloo, world_size 2): reads are sharded by contiguous global read index over ranks, every rank
builds rank-LOCAL classes, and one all-reduce(sum) of the per-transcript vector per EM iteration gives exactly the merged
result, because class weights depend only on the label and the E-step is linear in the class count.
No GPU here, so the per-rank compute is played by the oracle -- this tests the sharding / reduction protocol
bench.py and libsfb200 (sfb200_comm_init, em_run with a communicator) implement, not the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as O
from sailfish_b200 import synth


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def em_step(T, rp, lab, cnt, eff, alpha):
    """one EMUpdate_ on local classes: returns the local alphaOut (CollapsedEMOptimizer.cpp:224-281)"""
    out = np.zeros(T)
    for e in range(len(cnt)):
        ids = lab[int(rp[e]):int(rp[e + 1])]
        if len(ids) == 1:
            out[ids[0]] += cnt[e]
            continue
        w = 1.0 / np.maximum(eff[ids], 1.0)
        w = w / w.sum()
        v = alpha[ids] * w
        den = v.sum()
        if den > 0:
            np.add.at(out, ids, v * (cnt[e] / den))
    return out


def worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq, off, ln = synth.make_transcriptome(30, seed=9)
    n = 6000
    b1, o1, _, _, _ = synth.make_reads(seq, off, ln, n, 76, seed=4)
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    ix = O.Index(seqs, k=31)
    fmt = O.parse_libtype("U")
    lo, hi = rank * n // world, (rank + 1) * n // world                     # contiguous global read ranges
    run = O.Run(ix, O.MapOpts.default(fmt))
    run.map_batch(b1.tobytes()[lo * 76:hi * 76], o1[:hi - lo + 1])
    loc = run.finish()
    T = len(ln)
    # once after mapping: counters and active flags (map_finish / em_run with a communicator)
    counters = torch.from_numpy(loc["counters"].astype(np.int64)); dist.all_reduce(counters)
    active = torch.zeros(T, dtype=torch.int64); active[torch.from_numpy(loc["labels"].astype(np.int64))] = 1
    dist.all_reduce(active)
    n_active = int((active > 0).sum()); num_mapped = int(counters[1])
    eff = np.full(T, 500.0)
    alpha = np.where(active.numpy() > 0, num_mapped / n_active, 0.0)
    for _ in range(25):                                                      # per iteration: ONE all-reduce of the T-vector
        out = torch.from_numpy(em_step(T, loc["row_ptr"], loc["labels"], loc["counts"], eff, alpha))
        dist.all_reduce(out)
        alpha = out.numpy().copy()
    if rank == 0:
        # merged single-process run over all reads
        run = O.Run(ix, O.MapOpts.default(fmt)); run.map_batch(b1.tobytes(), o1); g = run.finish()
        rc, want, it, _ = O.em_run(T, g["row_ptr"], g["labels"], g["counts"], eff, int(g["counters"][1]), O.EMOpts.default(fixed_iters=25))
        ret["counters_ok"] = counters.numpy().tolist() == g["counters"].astype(np.int64).tolist()
        alpha = np.where(alpha <= 1e-8, 0.0, alpha)                         # truncateCountVector (CollapsedEMOptimizer.cpp:37-44)
        ret["max_rel"] = float(np.max(np.abs(alpha - want) / np.maximum(want, 1e-3)))
        ret["sum_ok"] = abs(alpha.sum() - num_mapped) < 1e-6 * num_mapped
    dist.destroy_process_group()


def test_sharded_reads_allreduce_matches_merged():
    mgr = mp.Manager(); ret = mgr.dict()
    port = free_port()
    mp.spawn(worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["counters_ok"]
    assert ret["sum_ok"]
    assert ret["max_rel"] < 1e-9
