#!/usr/bin/env python
"""bench.py -- quantification throughput on the BASELINE.json configurations.

Default (what the driver runs) is BASELINE config 2: synthetic GENCODE-like transcriptome, 40 000 genes x 5 isoforms = 200 000
transcripts (seed 42), 10 M single-end 76 nt reads per GPU (seed 1234 + rank, 0.5% substitutions), library type U, k = 31, EM with
exactly 1000 iterations.  One step = one whole quantification of the read set:
    map_begin -> map_batch x B -> map_finish (equivalence classes) -> effective lengths -> em_run [-> bootstraps -> Gibbs samples].
`value` is reads/s with the reads already resident in HBM; `e2e` is the same metric through the C ABI with the reads in pinned HOST
memory (H2D inside the timed region, estimates read back).  `--config 3|4|5` select the other BASELINE configurations (paired-end;
reads drawn on the GPU with torch because numpy would take minutes; `--reads` scales them down, the workload string says so):
    3: 200 k transcripts, 100 M pairs 2x100, VBEM to convergence + 100 bootstraps + 100 Gibbs samples
    4: 200 k transcripts, 50 M pairs 2x100 per GPU (x 8 GPUs = 400 M), EM to convergence
    5: 1 M transcripts, 50 M pairs 2x150, EM to convergence
N > 1: reads are sharded over ranks (weak scaling); counters and the fragment-length sample are combined once and the ranks' class
tables are merged with one all-gather, after which every rank runs the EM locally (SFB200_MULTI_EM_ALLREDUCE=1
keeps the classes rank-local and all-reduces the per-transcript vector every iteration instead).

The line also carries: `parity` -- the first S reads of the SAME read set mapped by the CPU oracle and by the GPU, classes compared as
multisets of (label, count), counters compared, and the estimates after the same number of EM iterations compared; `realistic` -- the
same measurement on a transcriptome with paralog families and repeats (equivalence classes that cross genes, large connected
components), naming the EM kernel that ran; `roofline` / `em_roofline` / `cpu_baseline` as the contract asks.

`--impl reference` times the CPU oracle (oracle/, the restatement of the reference's algorithm; the reference binary itself cannot be
built offline) on all host threads on a bounded proportional sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sailfish_b200 import synth, efflen  # noqa: E402

METRIC = "quant_reads_per_sec"
UNIT = "reads/s"
LIB_U = (0 & 1) | (3 << 1) | (4 << 3)      # LibraryFormat(SINGLE_END, NONE, U).formatID()
LIB_IU = (1 & 1) | (2 << 1) | (4 << 3)     # LibraryFormat(PAIRED_END, TOWARD, U).formatID()

CONFIGS = {
    2: dict(genes=40_000, reads=10_000_000, read_len=76, paired=False, infer="em", fixed_iters=1000, n_boot=0, n_gibbs=0, batch=2_500_000),
    3: dict(genes=40_000, reads=100_000_000, read_len=100, paired=True, infer="vbem", fixed_iters=0, n_boot=100, n_gibbs=100, batch=4_000_000),
    4: dict(genes=40_000, reads=50_000_000, read_len=100, paired=True, infer="em", fixed_iters=0, n_boot=0, n_gibbs=0, batch=4_000_000),
    5: dict(genes=200_000, reads=50_000_000, read_len=150, paired=True, infer="em", fixed_iters=0, n_boot=0, n_gibbs=0, batch=4_000_000),
}
EM_KERNELS = ["k_em_persistent", "k_em_part", "k_em_gather", "k_em_transcript_pass+k_em_sweep", "k_em_dense", "k_em_dense (components + pool loop)"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Workload:
    """everything both arms must agree on: the transcriptome, the read model, the inference settings"""

    def __init__(self, cfg_id, reads=None, genes=None, em_iters=None, batch=None, structure="iid", n_boot=None, n_gibbs=None):
        c = dict(CONFIGS[cfg_id])
        self.cfg_id = cfg_id
        self.full_reads = c["reads"]
        if reads:
            c["reads"] = reads
        if genes:
            c["genes"] = genes
        if em_iters is not None:                     # a fixed iteration count (default for config 2; a diagnostic elsewhere)
            c["fixed_iters"] = em_iters
        if batch:
            c["batch"] = batch
        if n_boot is not None:
            c["n_boot"] = n_boot
        if n_gibbs is not None:
            c["n_gibbs"] = n_gibbs
        self.__dict__.update(c)
        self.structure = structure
        self.lib = LIB_IU if self.paired else LIB_U
        self.use_vb = 1 if self.infer == "vbem" else 0

    def describe(self):
        what = "%dk-transcript synthetic index (seed 42%s), %.1fM %s %dnt reads per GPU (seed 1234+rank)%s, -l %s, k=31, %s" % (
            self.genes * 5 // 1000, ", paralog families + repeats" if self.structure == "paralog" else "",
            self.reads / 1e6, "pairs of 2x" if self.paired else "single-end", self.read_len,
            "" if self.reads == self.full_reads else " [REDUCED from %.0fM]" % (self.full_reads / 1e6),
            "IU" if self.paired else "U",
            ("%s fixed %d iters" % (self.infer.upper(), self.fixed_iters)) if self.fixed_iters else ("%s to convergence" % self.infer.upper()))
        if self.n_boot or self.n_gibbs:
            what += " + %d bootstraps + %d Gibbs samples" % (self.n_boot, self.n_gibbs)
        return "cfg%d: %s" % (self.cfg_id, what)

    def config(self):
        bytes_in = self.reads * self.read_len * (2 if self.paired else 1)
        return {"workload": self.describe(), "baseline_config": self.cfg_id, "n_transcripts": self.genes * 5, "reads_per_gpu": self.reads,
                "read_len": self.read_len, "paired": self.paired, "inference": self.infer, "em_iters": self.fixed_iters,
                "bootstraps": self.n_boot, "gibbs_samples": self.n_gibbs, "batch_reads": self.batch, "structure": self.structure,
                "l2": "inputs larger than L2 (reads %d MB, index several GB)" % (bytes_in >> 20)}

    def transcriptome(self):
        cache = os.environ.get("SFB200_BENCH_CACHE")
        f = os.path.join(cache, "txome_%d_%s.npz" % (self.genes, self.structure)) if cache else None
        if f and os.path.exists(f):
            z = np.load(f)
            return z["seq"], z["off"], z["ln"]
        kw = dict(family_frac=0.05, repeat_frac=0.01) if self.structure == "paralog" else {}
        seq, off, ln = synth.make_transcriptome(self.genes, seed=42, **kw)
        if f:
            os.makedirs(cache, exist_ok=True)
            np.savez(f, seq=seq, off=off, ln=ln)
        return seq, off, ln

    def reads_numpy(self, seq, off, ln, n, seed, out1, out2=None, chunk=1_000_000):
        """fills uint8 arrays chunk by chunk; expression profile fixed by seed 1234"""
        L = self.read_len
        done = c = 0
        while done < n:
            m = min(chunk, n - done)
            b1, _, b2, _, _ = synth.make_reads(seq, off, ln, m, L, seed=seed, expr_seed=1234, stream=c, paired=self.paired)
            out1[done * L:(done + m) * L] = b1
            if self.paired:
                out2[done * L:(done + m) * L] = b2
            done += m
            c += 1

    def eff_lens(self, ln, fld):
        if not self.paired:
            return efflen.effective_lengths(ln, None, single_end=True)     # SailfishQuantify.cpp:1039-1042 (Gaussian prior)
        return efflen.effective_lengths(ln, fld)                           # :648-838 from the observed fragment lengths

    def b_map_bytes(self, work, n_reads, label_words):
        """per-run algorithmic bytes of the mapper = sum over fragments of
        bases + 32*P + sizeof(IndexT)*S + 64*S + X + 4*|label| + 16   (IndexT = u32; ASCII bases)"""
        P, S, X = (float(w) for w in work)
        bases = self.read_len * (2 if self.paired else 1)
        return bases * n_reads + 32.0 * P + 4.0 * S + 64.0 * S + X + 4.0 * label_words + 16.0 * n_reads


def oracle_sample(wl, oidx, b1, b2, n_total, threads, target_s, keep_labels=0):
    """One bounded, proportional sample of the workload on the CPU oracle: map S fragments + finish + the EM for fixed_iters * S /
    n_total iterations (fixed-iteration workloads) or to convergence, on the sample's own classes.  -> (run(S), S)"""
    from oracle import pyoracle as O
    opts = O.MapOpts.default(wl.lib)
    L = wl.read_len

    def run(S, labels=0):
        off = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
        t0 = time.perf_counter()
        r = O.Run(oidx, opts)
        if labels:
            r.keep_labels(True)
        if wl.paired:
            r.map_batch(b1[:S * L], off, b2[:S * L], off, n_threads=threads)
        else:
            r.map_batch(b1[:S * L], off, n_threads=threads)
        res = r.finish()
        t1 = time.perf_counter()
        eff = wl.eff_lens(oidx.txp_len, res["fld"])
        if wl.fixed_iters:
            iters = max(1, int(round(wl.fixed_iters * S / float(n_total))))
            eo = O.EMOpts.default(fixed_iters=iters, use_vb=wl.use_vb)
        else:
            eo = O.EMOpts.default(use_vb=wl.use_vb)
        nm = int(res["counters"][1])
        rc, alphas, it, _ = O.em_run(len(oidx.txp_len), res["row_ptr"], res["labels"], res["counts"], eff, nm, eo, n_threads=threads)
        t2 = time.perf_counter()
        return dict(S=S, t_map=t1 - t0, t_em=t2 - t1, iters=it, work=r.work(), res=res, alphas=alphas, eff=eff, run=r, em_opts=eo)

    probe = run(min(50_000, n_total))
    rate = probe["S"] / (probe["t_map"] + probe["t_em"])
    S = int(min(n_total, max(100_000, rate * target_s)))
    return run, S


def truth_stats(run_obj, truth, n):
    """sensitivity / precision of the mapping against the synthetic read origins (oracle labels; the GPU's classes equal the oracle's)"""
    n = min(n, len(truth))
    mapped = hit = 0
    for i in range(n):
        lab = run_obj.last_label(i)
        if lab is None or len(lab) == 0:
            continue
        mapped += 1
        if int(truth[i]) in lab:
            hit += 1
    return {"reads": n, "sensitivity": hit / float(n), "precision": hit / float(max(mapped, 1)), "mapped_frac": mapped / float(n)}


def class_multiset(rp, lab, cnt):
    rp = np.asarray(rp, np.int64)
    return sorted((tuple(lab[rp[i]:rp[i + 1]].tolist()), int(cnt[i])) for i in range(len(cnt)))


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries (NCCL prints its version banner) write to fd 1 behind Python's
    back: point fd 1 at stderr for the duration of the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def reference_arm(args, wl):
    """the CPU arm: oracle port on all host threads, bounded proportional sample.  No GPU, no libsfb200."""
    from oracle import pyoracle as O
    threads = host_threads()
    seq, off, ln = wl.transcriptome()
    log("[bench] transcriptome: %d transcripts, %.1f Mnt" % (len(ln), seq.size / 1e6))
    n_total = wl.reads
    n_gen = min(n_total, 2_000_000)
    L = wl.read_len
    b1 = np.empty(n_gen * L, np.uint8)
    b2 = np.empty(n_gen * L, np.uint8) if wl.paired else None
    wl.reads_numpy(seq, off, ln, n_gen, 1234, b1, b2)
    t0 = time.time()
    oidx = O.Index.from_text(seq, off, ln, k=31, n_threads=threads)
    log("[bench] oracle index built on %d host threads (%.1fs)" % (threads, time.time() - t0))
    run, S = oracle_sample(wl, oidx, b1, b2, n_total, threads, target_s=4.0)
    S = min(S, n_gen)
    for _ in range(args.warmup):
        run(S)
    ts = []
    for _ in range(args.steps):
        r = run(S)
        ts.append(r["t_map"] + r["t_em"])
    t_step = float(np.mean(ts))
    val = S / t_step
    sample = "%d of %d fragments mapped + %d EM iterations per step (%s), oracle port, %d threads" % (
        S, n_total, r["iters"], "proportional sample of %d" % wl.fixed_iters if wl.fixed_iters else "to convergence on the sample's classes", threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64+f64", "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "detail": {"map_reads_per_s": S / r["t_map"], "em_iters_per_s": r["iters"] / max(r["t_em"], 1e-9)}}
    emit(line)
    return 0


class GpuRun:
    """the product arm on one rank: context, index, reads (device + pinned host), the step"""

    def __init__(self, wl, rank, local_rank, world, dist):
        import torch
        from sailfish_b200 import capi
        self.torch, self.capi = torch, capi
        self.wl, self.rank, self.world, self.dist = wl, rank, world, dist
        self.ctx = capi.Context(local_rank)
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.from_numpy(capi.Context.comm_unique_id()))
            dist.broadcast(uid, 0)
            self.ctx.comm_init(world, rank, uid.cpu().numpy())
        t0 = time.time()
        self.seq, self.off, self.ln = wl.transcriptome()
        log("[bench] transcriptome (%s): %d transcripts, %.1f Mnt (%.1fs)" % (wl.structure, len(self.ln), self.seq.size / 1e6, time.time() - t0))
        t0 = time.time()
        self.st = self.ctx.index_build(seq=self.seq, txp_off=self.off, txp_len=self.ln, k=31)
        log("[bench] index: %d positions, %d k-mers, %.2f GB in HBM, max bucket %d (%.1fs)" % (
            self.st["n_sa"], self.st["n_kmers"], self.st["hbm_bytes"] / 1e9, self.st["max_bucket"], time.time() - t0))
        # the page-locked read buffers are allocated on the GPU's NUMA node (sfb200_bind_host_near_device: what the host program does
        # for its parser threads); the CPU set is restored afterwards so that the CPU leg keeps every host core
        cpus = os.sched_getaffinity(0)
        self.numa_cpus = capi.bind_host_near_device(local_rank)
        try:
            self.make_reads()
        finally:
            os.sched_setaffinity(0, cpus)
        self.map_opts = capi.MapOpts.default(wl.lib)
        if wl.fixed_iters:
            self.em_opts = capi.EMOpts.default(fixed_iters=wl.fixed_iters, use_vb=wl.use_vb)
        else:
            self.em_opts = capi.EMOpts.default(use_vb=wl.use_vb)
        n = wl.reads
        self.cuts = list(range(0, n, wl.batch)) + [n]
        self.phase = {"begin": 0.0, "batches": 0.0, "finish": 0.0, "efflen": 0.0, "em": 0.0, "boot": 0.0, "gibbs": 0.0}
        self.eff_se = wl.eff_lens(self.ln, None) if not wl.paired else None

    def make_reads(self):
        torch, wl = self.torch, self.wl
        n, L = wl.reads, wl.read_len
        t0 = time.time()
        self.truth = None
        nm = 2 if wl.paired else 1
        self.h_b = [torch.empty(n * L + 8, dtype=torch.uint8).pin_memory() for _ in range(nm)]
        if wl.cfg_id == 2 and wl.structure == "iid":
            # the round-1 read set, drawn with numpy (both arms can produce it without a GPU)
            cache = os.environ.get("SFB200_BENCH_CACHE")
            f = os.path.join(cache, "reads_%d_%d_%d.npy" % (wl.genes, n, 1234 + self.rank)) if cache else None
            if f and os.path.exists(f):
                self.h_b[0].numpy()[:] = np.load(f)
            else:
                wl.reads_numpy(self.seq, self.off, self.ln, n, 1234 + self.rank, self.h_b[0].numpy())
                if f:
                    np.save(f, self.h_b[0].numpy())
            self.d_b = [self.h_b[0].cuda()]
        else:
            seq_d = torch.from_numpy(self.seq).cuda()
            d1 = torch.empty(n * L + 8, dtype=torch.uint8, device="cuda")
            d2 = torch.empty(n * L + 8, dtype=torch.uint8, device="cuda") if wl.paired else None
            _, _, self.truth = synth.make_reads_device(seq_d, self.off, self.ln, n, L, seed=1234 + self.rank, paired=wl.paired,
                                                       out1=d1[:n * L], out2=d2[:n * L] if wl.paired else None)
            self.d_b = [d1] + ([d2] if wl.paired else [])
            del seq_d
            for h, d in zip(self.h_b, self.d_b):
                h.copy_(d)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
        self.h_off = (torch.arange(n + 1, dtype=torch.int64) * L).pin_memory()
        self.d_off = self.h_off.cuda()
        log("[bench] reads: %d x %d x %d nt (%.1fs)" % (n, nm, L, time.time() - t0))

    def step(self, host, cuts=None):
        ctx, wl, ph = self.ctx, self.wl, self.phase
        cuts = cuts or self.cuts
        t_a = time.perf_counter()
        ctx.map_begin(self.map_opts)
        t_b = time.perf_counter()
        B = self.h_b if host else self.d_b
        O_ = self.h_off if host else self.d_off
        L = wl.read_len
        for a, b in zip(cuts[:-1], cuts[1:]):
            if host:
                # the benchmark's reads all have one length: the fixed-length entry point (what sfb200-quant calls for such a batch)
                ctx.map_batch_fixed_ptr(B[0].data_ptr() + a * L, L, B[1].data_ptr() + a * L if wl.paired else 0, L if wl.paired else 0, b - a)
            else:
                p2 = B[1].data_ptr() if wl.paired else 0
                o2 = O_.data_ptr() + 8 * a if wl.paired else 0
                ctx.map_batch_ptr(B[0].data_ptr(), O_.data_ptr() + 8 * a, p2, o2, b - a, device=True)
        t_c = time.perf_counter()
        g = ctx.map_finish()
        t_d = time.perf_counter()
        eff = self.eff_se if not wl.paired else wl.eff_lens(self.ln, g["fld"])
        t_e = time.perf_counter()
        nm = int(g["counters"][1])               # summed over ranks by map_finish when a communicator is set
        alphas, iters, _ = ctx.em_run(eff, nm, self.em_opts)
        self.last_em_ms = ctx.last_em_loop_ms()          # of the optimizer run itself (bootstrap_run adds its replicates' loops)
        t_f = time.perf_counter()
        extra = {}
        if wl.n_boot:
            rows = ctx.bootstrap_run(eff, wl.n_boot, seed=7, opts=self.em_opts)
            extra["boot_mean_total"] = float(rows.sum(axis=1).mean())
            extra["boot_loop_ms"] = ctx.last_em_loop_ms()
        t_g = time.perf_counter()
        if wl.n_gibbs:
            rows = ctx.gibbs_run(eff, alphas / alphas.sum(), nm, wl.n_gibbs, seed=7)
            extra["gibbs_total_ok"] = bool((rows.sum(axis=1) == nm).all())
        t_h = time.perf_counter()
        ph["begin"] += t_b - t_a; ph["batches"] += t_c - t_b; ph["finish"] += t_d - t_c; ph["efflen"] += t_e - t_d
        ph["em"] += t_f - t_e; ph["boot"] += t_g - t_f; ph["gibbs"] += t_h - t_g
        return g, alphas, iters, eff, extra

    def timed(self, host, steps):
        torch, dist, ctx = self.torch, self.dist, self.ctx
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        map_ms = em_ms = 0.0
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(steps):
                out = self.step(host)
                map_ms += ctx.last_map_kernel_ms(); em_ms += self.last_em_ms
            e1.record(self.stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, map_ms, em_ms, out

    def measure(self, steps, warmup, with_e2e=True, sample_clocks=False, local_rank=0):
        """-> dict of the measured numbers of this workload"""
        wl = self.wl
        for _ in range(warmup):
            self.step(False)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        for kk in self.phase:
            self.phase[kk] = 0.0
        ms, launches, map_ms, em_ms, out = self.timed(False, steps)
        g, alphas, iters, eff, extra = out
        host_phase_ms = {kk: round(v * 1e3 / steps, 3) for kk, v in self.phase.items()}
        clocks = sampler.stop() if sampler else None
        r = {"ms": ms, "launches": launches, "map_ms": map_ms / steps, "em_ms": em_ms / steps, "g": g, "alphas": alphas, "iters": iters,
             "eff": eff, "extra": extra, "host_phase_ms": host_phase_ms, "clocks": clocks, "em_kernel": EM_KERNELS[self.ctx.last_em_kernel()], "em_variant": self.ctx.last_em_variant()}
        total_reads = wl.reads * self.world
        r["value"] = total_reads * steps / (ms / 1e3)
        if with_e2e:
            self.step(True)
            ms_e2e, _, _, _, _ = self.timed(True, steps)
            r["ms_e2e"] = ms_e2e
            r["h2d_map_bytes"] = self.ctx.map_h2d_bytes()          # of the last step (the counter restarts at map_begin)
            r["e2e_value"] = total_reads * steps / (ms_e2e / 1e3)
        return r

    def parity(self, cpu, S):
        """the CPU sample's reads through the GPU: classes as multisets, counters, estimates at equal iteration count"""
        ctx, wl, capi = self.ctx, self.wl, self.capi
        cuts = [0, S // 2, S]
        ctx.map_begin(self.map_opts)
        for a, b in zip(cuts[:-1], cuts[1:]):
            p2 = self.d_b[1].data_ptr() if wl.paired else 0
            o2 = self.d_off.data_ptr() + 8 * a if wl.paired else 0
            ctx.map_batch_ptr(self.d_b[0].data_ptr(), self.d_off.data_ptr() + 8 * a, p2, o2, b - a, device=True)
        g = ctx.map_finish()
        rp, lab, cnt = ctx.eq_export()
        res = cpu["res"]
        counters_equal = g["counters"].tolist() == res["counters"].tolist()
        fld_equal = g["fld"].tolist() == res["fld"].tolist()
        classes_equal = (len(cnt) == len(res["counts"]) and class_multiset(rp, lab, cnt) == class_multiset(res["row_ptr"], res["labels"], res["counts"]))
        eo = cpu["em_opts"]
        go = capi.EMOpts.default(use_vb=wl.use_vb, fixed_iters=cpu["iters"])      # the same number of iterations as the CPU ran
        a, it, _ = ctx.em_run(cpu["eff"], int(g["counters"][1]), go)
        want = cpu["alphas"]
        big = want > 1e-3
        rel = float(np.max(np.abs(a[big] - want[big]) / want[big])) if big.any() else 0.0
        zeros_equal = bool(((a == 0) == (want == 0)).all())
        ok = counters_equal and fld_equal and classes_equal and rel <= 1e-4 and zeros_equal and it == cpu["iters"]
        del eo
        return {"ok": bool(ok), "reads": int(S), "n_classes": int(len(cnt)), "classes_equal": bool(classes_equal), "counters_equal": bool(counters_equal),
                "fld_equal": bool(fld_equal), "em_iters": int(it), "em_max_rel_err": rel, "em_zero_pattern_equal": zeros_equal, "tolerance": 1e-4}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="reads (fragments) per GPU; default: the configuration's")
    ap.add_argument("--genes", type=int, default=0)
    ap.add_argument("--em-iters", type=int, default=None)
    ap.add_argument("--batch", type=int, default=0, help="fragments per map_batch call")
    ap.add_argument("--bootstraps", type=int, default=None)
    ap.add_argument("--gibbs", type=int, default=None)
    ap.add_argument("--structure", default="iid", choices=["iid", "paralog"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-realistic", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = Workload(args.config, reads=args.reads, genes=args.genes, em_iters=args.em_iters, batch=args.batch, structure=args.structure,
                  n_boot=args.bootstraps, n_gibbs=args.gibbs)
    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, wl)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    run = GpuRun(wl, rank, local_rank, world, dist)
    m = run.measure(args.steps, args.warmup, with_e2e=True, sample_clocks=(rank == 0), local_rank=local_rank)
    if wl.fixed_iters:
        assert m["iters"] == wl.fixed_iters
    n, L, T = wl.reads, wl.read_len, len(run.ln)
    h2d = m["h2d_map_bytes"] + T * 8               # what sfb200_map_batch sent (bases; offsets unless the reads have one length) + eff. lengths
    d2h = T * 8 + 6 * 8 + 4000 + wl.n_boot * T * 8 + wl.n_gibbs * T * 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    g = m["g"]
    E, nnz = g["n_classes"], g["nnz"]
    steps = args.steps
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": m["ms"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
            "data": "synthetic", "config": wl.config(), "clocks": m["clocks"],
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": m["ms_e2e"] / steps},
            "gpu_launches": m["launches"],
            "detail": {"map_kernel_ms_per_step": m["map_ms"], "em_loop_ms_per_step": m["em_ms"], "em_iters": m["iters"], "em_kernel": m["em_kernel"] + (" (streaming variant)" if m["em_variant"] & 1 else "") + (" (lagged stopping rule)" if m["em_variant"] & 2 else ""),
                       "map_kernel_reads_per_s": n / (m["map_ms"] / 1e3), "em_iters_per_s": m["iters"] / (m["em_ms"] / 1e3),
                       "host_wall_ms_per_step": m["host_phase_ms"], "n_classes": E, "nnz": nnz, "mapped": int(g["counters"][1]),
                       "observed": int(g["counters"][0]), "index_hbm_gb": round(run.st["hbm_bytes"] / 1e9, 2),
                       "host_cpus_on_gpu_node": run.numa_cpus, **m["extra"]}}
    # EM roofline: B_em = 12 nnz + 12 E + 32 T algorithmic bytes per iteration.  The dense / gather / part loops keep their
    # working set in shared memory (no HBM traffic after the first iteration): their bound is on-chip, the HBM figure is an equivalent
    b_em = 12.0 * nnz + 12.0 * E + 32.0 * T
    em_gbs = b_em * m["iters"] / (m["em_ms"] / 1e3) / 1e9
    streamed = bool(m["em_variant"] & 1)             # k_em_dense reading counts / base / 1/effLen from a global block every iteration
    on_chip = m["em_kernel"].startswith(("k_em_dense", "k_em_gather", "k_em_part")) and not streamed
    em_traffic = None
    try:
        em_traffic = None if streamed else json.load(open(os.path.join(ROOT, "profiles", "em_kernel_traffic.json"))).get(m["em_kernel"])
    except Exception:
        pass
    line["em_roofline"] = {"bound": "on-chip (shared memory / issue)" if on_chip else "hbm", "achieved": em_gbs, "peak": peak, "unit": "GB/s",
                           "frac": em_gbs / peak, "bytes_per_iter": b_em, "kernel": m["em_kernel"], "us_per_iter": m["em_ms"] * 1e3 / max(m["iters"], 1),
                           "traffic": em_traffic,
                           "note": ("algorithmic bytes (SURVEY 8d) over time, an HBM-EQUIVALENT rate: the %.0f MB working set is staged in shared memory once, "
                                    "so frac may exceed 1 and is not an HBM utilisation" % (b_em / 1e6)) if on_chip else "algorithmic bytes (SURVEY 8d) over time"}

    cpu = None
    per_read = traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "map_kernel_traffic.json")))
        traffic = tj.get("dram_bytes_per_read") if wl.cfg_id == 2 and wl.structure == "iid" else None
        per_read = tj.get("algorithmic_bytes_per_read") if wl.cfg_id == 2 and wl.structure == "iid" else None
    except Exception:
        pass
    if not args.no_cpu_baseline and world == 1:
        from oracle import pyoracle as O
        threads = host_threads()
        t0 = time.time()
        words, sa_pos, sa_tid = run.ctx.index_export()
        oidx = O.Index.from_table(words, run.st["text_len"], run.ln, 31, sa_pos, sa_tid, run.ctx.index_export_table())
        del words, sa_pos, sa_tid
        log("[bench] oracle index from the device arrays (bit-identical to the oracle's own build: tests) (%.1fs)" % (time.time() - t0))
        S_max = min(n, 4_000_000)
        hb1 = run.h_b[0].numpy()[:S_max * L]
        hb2 = run.h_b[1].numpy()[:S_max * L] if wl.paired else None
        orun, S = oracle_sample(wl, oidx, hb1, hb2, n, threads, target_s=10.0)
        S = min(S, S_max)
        orun(min(S, 200_000))
        r = orun(S, labels=1 if run.truth is not None else 0)
        cpu_val = S / (r["t_map"] + r["t_em"])
        cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d of %d fragments mapped + %d EM iterations (%s), oracle port, %d threads, %.1fs" % (
                   S, n, r["iters"], "proportional share of %d" % wl.fixed_iters if wl.fixed_iters else "to convergence on the sample's classes",
                   threads, r["t_map"] + r["t_em"]),
               "map_reads_per_s": S / r["t_map"], "em_iters_per_s": r["iters"] / max(r["t_em"], 1e-9)}
        # roofline of the mapping kernels: algorithmic bytes per fragment from the oracle's work counters on the sample
        lab_words = float((np.diff(r["res"]["row_ptr"]).astype(np.float64) * r["res"]["counts"].astype(np.float64)).sum())
        per_read = wl.b_map_bytes(r["work"], S, lab_words) / S
        # parity at benchmark scale: the same S fragments through the GPU
        line["parity"] = run.parity(r, S)
        if run.truth is not None:
            line["parity"]["truth"] = truth_stats(r["run"], run.truth[:100_000].cpu().numpy(), 100_000)
        log("[bench] parity: %s" % json.dumps(line["parity"]))
    if per_read:
        gbs = per_read * n / (m["map_ms"] / 1e3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                            "traffic": traffic * n if traffic else None, "kernel": "k_pack_reads+k_scan_reads+k_finalize_reads",
                            "bytes_per_read": per_read, "peak_source": peak_src,
                            "note": "random 32-byte-sector access: the honest bound is sectors/s, reported as bytes (SURVEY 8d); traffic = "
                                    "dram bytes of one ncu --set full capture (profiles/map_kernel_traffic.json) scaled to the run"}
    else:
        line["roofline"] = dict(line["em_roofline"], peak_source=peak_src)
    line["cpu_baseline"] = cpu

    # the same measurement on a transcriptome with paralog families and repeats (classes that cross genes, large components)
    if not args.no_realistic and world == 1 and wl.structure == "iid" and wl.cfg_id == 2:
        try:
            h_keep = None
            run.ctx.close()
            del run
            torch.cuda.empty_cache()
            wl2 = Workload(args.config, reads=min(wl.reads, 4_000_000), genes=args.genes, em_iters=args.em_iters, batch=min(wl.batch, 2_000_000),
                           structure="paralog")
            run2 = GpuRun(wl2, rank, local_rank, world, None)
            m2 = run2.measure(max(1, min(steps, 2)), 3, with_e2e=False)
            g2 = m2["g"]
            b2 = 12.0 * g2["nnz"] + 12.0 * g2["n_classes"] + 32.0 * T
            line["realistic"] = {"workload": wl2.describe(), "value": m2["value"], "unit": UNIT, "ms_per_step": m2["ms"] / max(1, min(steps, 2)),
                                 "map_kernel_ms_per_step": m2["map_ms"], "map_kernel_reads_per_s": wl2.reads / (m2["map_ms"] / 1e3),
                                 "em_kernel": m2["em_kernel"], "em_loop_ms_per_step": m2["em_ms"], "em_iters": m2["iters"],
                                 "em_us_per_iter": m2["em_ms"] * 1e3 / max(m2["iters"], 1), "em_bytes_per_iter": b2,
                                 "em_hbm_equiv_gbs": b2 * m2["iters"] / (m2["em_ms"] / 1e3) / 1e9,
                                 "n_classes": g2["n_classes"], "nnz": g2["nnz"], "mapped": int(g2["counters"][1]), "max_bucket": run2.st["max_bucket"]}
            del h_keep
        except Exception as e:       # the headline must survive a failure of the secondary measurement
            line["realistic"] = {"error": str(e)[:300]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
