#!/usr/bin/env python
"""bench.py -- quantification throughput on BASELINE.json config 2.

Workload (config.workload): synthetic GENCODE-like transcriptome, 40 000 genes x 5 isoforms = 200 000 transcripts
(seed 42), 10 M single-end 76 nt reads per GPU (seed 1234 + rank, 0.5% substitutions), library type U, k = 31,
EM with exactly 1000 iterations.  One step = one whole quantification of the read set:
    map_begin -> map_batch x B -> map_finish (equivalence classes) -> em_run(fixed 1000 iterations).
`value` is reads/s with the reads already resident in HBM; `e2e` is the same metric through the C ABI with the reads in
pinned HOST memory (H2D inside the timed region, estimates read back).  N > 1: reads are sharded over ranks (weak scaling,
10 M per rank); counters and the fragment-length sample are combined once and the ranks' class tables are merged with one
all-gather, after which every rank runs the EM locally (SFB200_MULTI_EM_ALLREDUCE=1 keeps the classes
rank-local and all-reduces the per-transcript vector every iteration instead).

`--impl reference` times the CPU oracle (oracle/, the restatement of the reference's algorithm; the reference binary itself
cannot be built offline) on all host threads on a bounded proportional sample of the same workload.
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
READ_LEN = 76
LIB_U = (0 & 1) | (3 << 1) | (4 << 3)      # LibraryFormat(SINGLE_END, NONE, U).formatID()


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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def gen_reads(seq, off, ln, n_reads, seed, out_bases, chunk=1_000_000):
    """fills out_bases (uint8[n_reads*READ_LEN]) chunk by chunk; expression profile fixed by seed 1234"""
    done = 0
    c = 0
    while done < n_reads:
        n = min(chunk, n_reads - done)
        b1, _, _, _, _ = synth.make_reads(seq, off, ln, n, READ_LEN, seed=seed, expr_seed=1234, stream=c)
        out_bases[done * READ_LEN:(done + n) * READ_LEN] = b1
        done += n
        c += 1


def b_map_bytes(work, n_reads, label_words):
    """per-run algorithmic bytes of the mapper = sum over reads of
    bases + 32*P + sizeof(IndexT)*S + 64*S + X + 4*|label| + 16   (IndexT = u32; ASCII bases)"""
    P, S, X = (float(w) for w in work)
    return READ_LEN * n_reads + 32.0 * P + 4.0 * S + 64.0 * S + X + 4.0 * label_words + 16.0 * n_reads


def cpu_sample(oidx, bases, n_total, em_iters_full, threads, target_s, want_work=False):
    """One bounded, proportional sample of the workload on the CPU oracle: map S reads + finish + EM for
    em_iters_full * S / n_total iterations on the sample's own classes.  Returns a closure running it and S."""
    from oracle import pyoracle as O
    opts = O.MapOpts.default(LIB_U)

    def run(S):
        off = np.arange(S + 1, dtype=np.uint64) * np.uint64(READ_LEN)
        t0 = time.perf_counter()
        r = O.Run(oidx, opts)
        r.map_batch(bases[:S * READ_LEN], off, n_threads=threads)
        res = r.finish()
        t1 = time.perf_counter()
        eff = efflen.effective_lengths(oidx.txp_len, None, single_end=True)
        iters = max(1, int(round(em_iters_full * S / float(n_total))))
        rc, alphas, it, _ = O.em_run(len(oidx.txp_len), res["row_ptr"], res["labels"], res["counts"], eff, int(res["counters"][1]),
                                     O.EMOpts.default(fixed_iters=iters), n_threads=threads)
        t2 = time.perf_counter()
        return dict(S=S, t_map=t1 - t0, t_em=t2 - t1, iters=iters, work=r.work(), res=res)

    probe = run(min(50_000, n_total))
    rate = probe["S"] / (probe["t_map"] + probe["t_em"])
    S = int(min(n_total, max(100_000, rate * target_s)))
    return run, S


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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--genes", type=int, default=40_000)
    ap.add_argument("--em-iters", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=2_500_000, help="reads per map_batch call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    have_gpu = torch.cuda.is_available()
    workload = "cfg2: %dk-transcript synthetic index (seed 42), %.1fM single-end %dnt reads per GPU (seed 1234+rank), -l U, k=31, EM fixed %d iters" % (
        args.genes * 5 // 1000, args.reads / 1e6, READ_LEN, args.em_iters)
    config = {"workload": workload, "n_transcripts": args.genes * 5, "reads_per_gpu": args.reads, "read_len": READ_LEN,
              "em_iters": args.em_iters, "batch_reads": args.batch, "l2": "inputs larger than L2 (reads %d MB, index several GB)" % (args.reads * READ_LEN >> 20)}

    t0 = time.time()
    # SFB200_BENCH_CACHE=<dir>: keep the generated (deterministic) inputs between invocations of one session
    cache = os.environ.get("SFB200_BENCH_CACHE")
    tx_file = os.path.join(cache, "txome_%d.npz" % args.genes) if cache else None
    if tx_file and os.path.exists(tx_file):
        z = np.load(tx_file); seq, off, ln = z["seq"], z["off"], z["ln"]
    else:
        seq, off, ln = synth.make_transcriptome(args.genes, seed=42)
        if tx_file:
            os.makedirs(cache, exist_ok=True)
            np.savez(tx_file, seq=seq, off=off, ln=ln)
    log("[bench] transcriptome: %d transcripts, %.1f Mnt (%.1fs)" % (len(ln), seq.size / 1e6, time.time() - t0))
    eff = efflen.effective_lengths(ln, None, single_end=True)     # SailfishQuantify.cpp:1039-1042 (single-end: Gaussian prior)

    # ---------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        from oracle import pyoracle as O
        threads = host_threads()
        n_total = args.reads
        bases = np.empty(min(n_total, 4_000_000) * READ_LEN, np.uint8)
        gen_reads(seq, off, ln, bases.size // READ_LEN, 1234, bases)
        t0 = time.time()
        if have_gpu:
            # index construction is outside the measured path of both arms: reuse the device-built index (bit-identical to
            # the oracle's own build, tests/test_gpu_map.py::test_index_matches_oracle); the timed region is oracle-only
            from sailfish_b200 import capi
            ctx = capi.Context(local_rank)
            ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
            words, sa_pos, sa_tid = ctx.index_export()
            oidx = O.Index.from_table(words, ctx.index_stats()["text_len"], ln, 31, sa_pos, sa_tid, ctx.index_export_table())
            ctx.close()
        else:
            seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
            oidx = O.Index(seqs, k=31)
        log("[bench] oracle index ready (%.1fs)" % (time.time() - t0))
        run, S = cpu_sample(oidx, bases, n_total, args.em_iters, threads, target_s=4.0)
        S = min(S, bases.size // READ_LEN)
        for _ in range(args.warmup):
            run(S)
        ts = []
        for _ in range(args.steps):
            r = run(S)
            ts.append(r["t_map"] + r["t_em"])
        t_step = float(np.mean(ts))
        val = S / t_step
        sample = "%d of %d reads mapped + %d of %d EM iterations per step (proportional sample), oracle port, %d threads" % (
            S, n_total, r["iters"], args.em_iters, threads)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int64+f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "detail": {"map_reads_per_s": S / r["t_map"], "em_iters_per_s": r["iters"] / r["t_em"]}}
        emit(line)
        return 0

    # ---------------------------------------------------------------------------------------------------------------
    if not have_gpu:
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    from sailfish_b200 import capi
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(capi.Context.comm_unique_id()))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, uid.cpu().numpy())

    t0 = time.time()
    st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    log("[bench] index: %d positions, %d k-mers, %.2f GB in HBM, max bucket %d (%.1fs)" % (
        st["n_sa"], st["n_kmers"], st["hbm_bytes"] / 1e9, st["max_bucket"], time.time() - t0))
    config["index_hbm_gb"] = round(st["hbm_bytes"] / 1e9, 2)

    n = args.reads
    t0 = time.time()
    h_bases = torch.empty(n * READ_LEN + 8, dtype=torch.uint8).pin_memory()
    rd_file = os.path.join(cache, "reads_%d_%d_%d.npy" % (args.genes, n, 1234 + rank)) if cache else None
    if rd_file and os.path.exists(rd_file):
        h_bases.numpy()[:] = np.load(rd_file)
    else:
        gen_reads(seq, off, ln, n, 1234 + rank, h_bases.numpy())
        if rd_file:
            np.save(rd_file, h_bases.numpy())
    h_off = (torch.arange(n + 1, dtype=torch.int64) * READ_LEN).pin_memory()
    log("[bench] reads: %d x %d nt (%.1fs)" % (n, READ_LEN, time.time() - t0))
    d_bases = h_bases.cuda()
    d_off = h_off.cuda()
    h_eff = np.ascontiguousarray(eff)
    em_opts = capi.EMOpts.default(fixed_iters=args.em_iters)
    map_opts = capi.MapOpts.default(LIB_U)
    cuts = list(range(0, n, args.batch)) + [n]
    num_mapped_global = [0]

    phase = {"begin": 0.0, "batches": 0.0, "finish": 0.0, "em": 0.0}

    def step(host):
        t_a = time.perf_counter()
        ctx.map_begin(map_opts)
        t_b = time.perf_counter()
        for a, b in zip(cuts[:-1], cuts[1:]):
            if host:
                ctx.map_batch_ptr(h_bases.data_ptr(), h_off.data_ptr() + 8 * a, 0, 0, b - a, device=False)
            else:
                ctx.map_batch_ptr(d_bases.data_ptr(), d_off.data_ptr() + 8 * a, 0, 0, b - a, device=True)
        t_c = time.perf_counter()
        g = ctx.map_finish()
        t_d = time.perf_counter()
        nm = int(g["counters"][1])               # summed over ranks by map_finish when a communicator is set
        alphas, iters, _ = ctx.em_run(h_eff, nm, em_opts)
        t_e = time.perf_counter()
        phase["begin"] += t_b - t_a; phase["batches"] += t_c - t_b; phase["finish"] += t_d - t_c; phase["em"] += t_e - t_d
        return g, alphas, iters

    def timed(host, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        map_ms = em_ms = 0.0
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                g, alphas, iters = step(host)
                map_ms += ctx.last_map_kernel_ms(); em_ms += ctx.last_em_loop_ms()
            e1.record(stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, map_ms, em_ms, g, alphas, iters

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for kk in phase:
        phase[kk] = 0.0
    ms, launches, map_ms, em_ms, g, alphas, iters = timed(False, args.steps)
    host_phase_ms = {kk: round(v * 1e3 / args.steps, 3) for kk, v in phase.items()}
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(1):
        step(True)
    ms_e2e, _, _, _, g2, alphas2, _ = timed(True, args.steps)
    assert iters == args.em_iters
    total_reads = n * world
    value = total_reads * args.steps / (ms / 1e3)
    e2e_value = total_reads * args.steps / (ms_e2e / 1e3)
    h2d = n * READ_LEN + (n + 1) * 8 + len(ln) * 8
    d2h = len(ln) * 8 + 6 * 8 + 4000

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

    E, nnz, T = g["n_classes"], g["nnz"], len(ln)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "detail": {"map_kernel_ms_per_step": map_ms / args.steps, "em_loop_ms_per_step": em_ms / args.steps,
                       "map_kernel_reads_per_s": n / (map_ms / args.steps / 1e3), "em_iters_per_s": args.em_iters / (em_ms / args.steps / 1e3),
                       "host_wall_ms_per_step": host_phase_ms, "n_classes": E, "nnz": nnz, "mapped": int(g["counters"][1]), "observed": int(g["counters"][0])}}
    # EM roofline: B_em = 12 nnz + 12 E + 32 T bytes per iteration
    b_em = 12.0 * nnz + 12.0 * E + 32.0 * T
    em_gbs = b_em * args.em_iters / (em_ms / args.steps / 1e3) / 1e9
    em_kernel = ["k_em_persistent", "k_em_part", "k_em_gather", "k_em_transcript_pass+k_em_sweep", "k_em_dense"][ctx.last_em_kernel()]
    line["em_roofline"] = {"bound": "hbm", "achieved": em_gbs, "peak": peak, "unit": "GB/s", "frac": em_gbs / peak,
                           "bytes_per_iter": b_em, "kernel": em_kernel,
                           "note": "algorithmic bytes (SURVEY 8d) over time; the %.0f MB working set is staged in shared memory once, so no HBM traffic after the first iteration" % (b_em / 1e6)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import pyoracle as O
        threads = host_threads()
        t0 = time.time()
        words, sa_pos, sa_tid = ctx.index_export()
        oidx = O.Index.from_table(words, st["text_len"], ln, 31, sa_pos, sa_tid, ctx.index_export_table())
        log("[bench] oracle index from device arrays (%.1fs)" % (time.time() - t0))
        run, S = cpu_sample(oidx, h_bases.numpy(), n, args.em_iters, threads, target_s=10.0)
        run(min(S, 200_000))
        r = run(S)
        cpu_val = S / (r["t_map"] + r["t_em"])
        cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d of %d reads mapped + %d of %d EM iterations (proportional sample), oracle port, %d threads, %.1fs" % (
                   S, n, r["iters"], args.em_iters, threads, r["t_map"] + r["t_em"]),
               "map_reads_per_s": S / r["t_map"], "em_iters_per_s": r["iters"] / r["t_em"]}
        # roofline of the dominant kernel (k_map_reads): algorithmic bytes per read from the oracle's work counters
        lab_words = float((np.diff(r["res"]["row_ptr"]).astype(np.float64) * r["res"]["counts"].astype(np.float64)).sum())
        per_read = b_map_bytes(r["work"], S, lab_words) / S
        gbs = per_read * n / (map_ms / args.steps / 1e3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "map_kernel_traffic.json"))).get("dram_bytes_per_read")
            traffic = traffic * n if traffic else None
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
                            "kernel": "k_pack_reads+k_scan_reads+k_finalize_reads", "bytes_per_read": per_read, "peak_source": peak_src,
                            "note": "random 32-byte-sector access: the honest bound is sectors/s, reported as bytes (SURVEY 8d)"}
    else:
        # no CPU sample in this run (N > 1 or --no-cpu-baseline): algorithmic bytes per read from the committed measurement
        per_read = traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "map_kernel_traffic.json")))
            per_read = tj.get("algorithmic_bytes_per_read"); traffic = tj.get("dram_bytes_per_read")
        except Exception:
            pass
        if per_read:
            gbs = per_read * n / (map_ms / args.steps / 1e3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "traffic": traffic * n if traffic else None, "kernel": "k_scan_reads+k_finalize_reads (per GPU)",
                                "bytes_per_read": per_read, "peak_source": peak_src,
                                "note": "bytes per read from profiles/map_kernel_traffic.json (oracle work counters on this workload)"}
        else:
            line["roofline"] = dict(line["em_roofline"], traffic=None, peak_source=peak_src)
    line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
