"""Generates the committed golden fixtures under tests/golden/ from the REAL reference pieces compiled into oracle/_ref
(run in the build container, where /root/reference exists):

  xxh64_kat.json      XXH64 known answers from the reference's own src/xxhash.c (ref_xxh64) and
                      TranscriptGroup hashes (src/TranscriptGroup.cpp:9-12)
  sample_data.npz     the bundled sample_data.tgz (15 transcripts, 10 000 read pairs x 50 nt) as arrays, the oracle's
                      equivalence classes for it (mapping spec v1, -l IU), and the estimates the reference's OWN
                      CollapsedEMOptimizer::optimize produces on those classes (EM and VBEM)
  synth_em.npz        a synthetic class set + the reference optimizer's EM / VBEM estimates
  eqbuilder.json      EquivalenceClassBuilder behaviour on a small add sequence (counts, finish() totals)
  ref_samplers.npz    a class set + per-transcript moments of the replicates the reference's OWN gatherBootstraps
                      (src/CollapsedEMOptimizer.cpp:557-709, EM and VBEM) and CollapsedGibbsSampler::sample
                      (src/CollapsedGibbsSampler.cpp:199-291) produce on it (both compiled unmodified into libsfref_em.so)
  bias_efflens.npz    inputs and outputs of the reference's OWN updateEffectiveLengths (src/SailfishUtils.cpp:611-926,
                      compiled unmodified into oracle/_ref/libsfref_em.so) for --biasCorrect and --gcBiasCorrect

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O          # noqa: E402
from sailfish_b200 import synth           # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def read_fasta(path):
    names, seqs, cur = [], [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur)); cur = []
            names.append(line[1:].split()[0])
        elif line:
            cur.append(line)
    if cur:
        seqs.append("".join(cur))
    return names, seqs


def read_fastq(path):
    out = []
    with open(path) as f:
        while True:
            h = f.readline()
            if not h:
                break
            out.append(f.readline().strip())
            f.readline(); f.readline()
    return out


def make_bias_fixture():
    """bias / GC effective-length correction: a small random transcriptome, read-start 6-mer counts,
    observed fragment GC histogram, fragment length counts, strand tallies, abundances -> the reference's corrected lengths"""
    rng = np.random.default_rng(20260117)
    T = 48
    lens = rng.integers(120, 2200, size=T)
    lens[:3] = [5, 7, 150]                                                      # shorter than the 6-mer window / than the FLD
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(l), p=[0.3, 0.2, 0.2, 0.3]).astype(np.uint8)) for l in lens]
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    eff_model = np.where(lens - 190.0 + 1 >= 1, lens - 190.0 + 1, lens).astype(np.float64)
    eff_in = eff_model * rng.uniform(0.9, 1.1, size=T)                          # as after an earlier correction round
    alphas = rng.lognormal(3, 2, size=T); alphas[rng.random(T) < 0.2] = 0.0; alphas[5] = 5e-9
    read_bias = rng.integers(1, 3000, size=4096).astype(np.uint32)
    observed_gc = rng.integers(1, 8000, size=101).astype(np.uint32)
    out = dict(seq=np.frombuffer(b"".join(seqs), np.uint8), txp_len=lens.astype(np.uint32), fld=fld, eff_model=eff_model, eff_in=eff_in,
               alphas=alphas, read_bias=read_bias, observed_gc=observed_gc, num_fwd=np.int64(61234), num_rc=np.int64(58766))
    for mode, tag in ((1, "seq"), (2, "gc")):
        for samp in ((1,) if mode == 1 else (1, 3)):
            Rb = O.RefBias(mode, seqs, eff_model, read_bias, observed_gc, fld, 61234, 58766, gc_samp=samp)
            rc, ref = Rb.update(alphas, eff_in)
            assert rc == 0
            out["ref_%s_samp%d" % (tag, samp)] = ref
            # a second round on the corrected lengths, as the optimizer does at iterations 500 and 1000
            rc, ref2 = Rb.update(alphas * 1.5, ref)
            out["ref_%s_samp%d_round2" % (tag, samp)] = ref2
    cdf, mx = Rb.fld(1000)
    out["ref_fld_cdf"] = cdf; out["ref_fld_max"] = np.int64(mx)
    # the whole optimizer with the correction switched on (lengths recomputed at iterations 50 / 500 / 1000): the reference's own
    # CollapsedEMOptimizer::optimize on classes over these transcripts, EM and VBEM
    rp, lab, cnt = synth.make_classes(T, 150, seed=11, gene_size=4)
    nm = int(cnt.sum())
    out.update(row_ptr=rp, labels=lab, counts=cnt, num_mapped=np.int64(nm))
    for mode, tag in ((1, "seq"), (2, "gc")):
        for vb in (0, 1):
            Ro = O.RefBias(mode, seqs, eff_model, read_bias, observed_gc, fld, 61234, 58766, classes=(rp, lab, cnt), num_mapped=nm, use_vb=bool(vb))
            rc, est, eff_after = Ro.optimize(tol=1e-5)          # tighter than the default 0.01: the run must pass iteration 50
            assert rc == 0
            out["opt_%s_vb%d_est" % (tag, vb)] = est; out["opt_%s_vb%d_eff" % (tag, vb)] = eff_after
    np.savez_compressed(os.path.join(OUT, "bias_efflens.npz"), **out)
    print("bias_efflens.npz: %d transcripts, %d corrected (seq), %d corrected (gc)" % (
        T, int((out["ref_seq_samp1"] != eff_in).sum()), int((out["ref_gc_samp1"] != eff_in).sum())))


def make_sampler_fixture():
    """bootstrap / Gibbs: the reference seeds its generators from std::random_device, so what can be
    pinned is the DISTRIBUTION of its replicates -- per-transcript mean and variance over many replicates, plus the exact
    invariants (every replicate conserves the fragment total)."""
    T = 300
    rp, lab, cnt = synth.make_classes(T, 600, seed=41, max_len=4)
    eff = np.random.default_rng(5).uniform(200, 3000, size=T)
    txp_len = np.maximum(eff, 1).astype(np.uint32)
    nm = int(cnt.sum())
    out = dict(txp_len=txp_len, eff=eff, row_ptr=rp, labels=lab, counts=cnt, num_mapped=np.int64(nm))
    NB, NG, BURN, NCHAIN = 400, 900, 200, 64
    for vb in (0, 1):
        ref = O.RefEM(txp_len, eff, rp, lab, cnt, nm, use_vb=bool(vb), n_boot=NB)
        rc, rows = ref.bootstraps()
        assert rc == 0 and rows.shape == (NB, T)
        if not vb:
            assert np.allclose(rows.sum(axis=1), nm, rtol=1e-9)          # EM hands out exactly the resampled total
        out["boot_vb%d_mean" % vb] = rows.mean(axis=0); out["boot_vb%d_var" % vb] = rows.var(axis=0)
        out["boot_vb%d_sum" % vb] = rows.sum(axis=1)
    out["boot_n"] = np.int64(NB)
    # Gibbs: chains started at the EM estimate mix slowly for some transcripts (integrated autocorrelation times of 100 and more), so
    # the pinned quantity is the mean / variance over a FIXED window of the chain (samples BURN .. NG), and its spread over
    # NCHAIN independent runs of the reference's sampler
    wm, wv = [], []
    for ch in range(NCHAIN):
        ref = O.RefEM(txp_len, eff, rp, lab, cnt, nm)
        rc, est, mass = ref.optimize()
        assert rc == 0
        rc, g = ref.gibbs(NG)
        assert rc == 0 and (g.sum(axis=1) == nm).all()
        g = g[BURN:].astype(np.float64)
        wm.append(g.mean(axis=0)); wv.append(g.var(axis=0))
    wm, wv = np.array(wm), np.array(wv)
    out["em_est"] = est
    out["gibbs_n"] = np.int64(NG); out["gibbs_burn"] = np.int64(BURN); out["gibbs_chains"] = np.int64(NCHAIN)
    out["gibbs_wmean_mean"] = wm.mean(axis=0); out["gibbs_wmean_sd"] = wm.std(axis=0, ddof=1)
    out["gibbs_wvar_mean"] = wv.mean(axis=0); out["gibbs_wvar_sd"] = wv.std(axis=0, ddof=1)
    np.savez_compressed(os.path.join(OUT, "ref_samplers.npz"), **out)
    print("ref_samplers.npz: %d bootstraps x 2 modes, %d Gibbs chains of %d samples (burn-in %d) of the reference's own samplers" % (NB, NCHAIN, NG, BURN))


def main():
    if "--bias-only" in sys.argv:
        make_bias_fixture()
        return
    if "--samplers-only" in sys.argv:
        make_sampler_fixture()
        return
    make_sampler_fixture()
    R = O.ref()
    assert R is not None, "oracle/_ref/libsfref.so missing: run make -C oracle"
    rng = np.random.default_rng(20261017)
    kat = []
    msgs = [b"", bytes(4), np.array([0, 1], np.uint32).tobytes(), np.array([3, 7, 11], np.uint32).tobytes(),
            np.arange(8, dtype=np.uint32).tobytes(), np.arange(9, dtype=np.uint32).tobytes(),
            np.array([0x0123456789ABCDEF], np.uint64).tobytes()]
    for n in list(range(0, 70)) + [127, 128, 129, 255, 256, 799, 800]:
        msgs.append(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
    for m in msgs:
        for seed in (0, 1, 0x9E3779B97F4A7C15):
            kat.append({"hex": m.hex(), "seed": seed, "xxh64": "%016x" % R.ref_xxh64(m, len(m), seed)})
    tg = []
    for n in (1, 2, 3, 5, 8, 9, 17, 64, 200):
        ids = np.sort(rng.choice(200000, size=n, replace=False)).astype(np.uint32)
        tg.append({"ids": ids.tolist(), "hash": "%016x" % R.ref_tgroup_hash(O._ptr(ids, O.u32p), n)})
    json.dump({"source": "reference src/xxhash.c via oracle/_ref/libsfref.so", "xxh64": kat, "transcript_group": tg},
              open(os.path.join(OUT, "xxh64_kat.json"), "w"))

    # EquivalenceClassBuilder behaviour (include/EquivalenceClassBuilder.hpp:64-108)
    adds = [[1, 2], [7], [1, 2], [1, 2, 3], [7], [1, 2], [2, 1], [5, 5], [1, 2]]
    h = R.ref_eqb_create()
    for a in adds:
        arr = np.array(a, np.uint32)
        R.ref_eqb_add(h, O._ptr(arr, O.u32p), len(a))
    import ctypes as C
    nnz = C.c_uint64()
    E = R.ref_eqb_finish(h, C.byref(nnz))
    rp = np.zeros(E + 1, np.uint64); lab = np.zeros(nnz.value, np.uint32); cnt = np.zeros(E, np.uint64); w = np.zeros(nnz.value, np.float64)
    R.ref_eqb_export(h, O._ptr(rp, O.u64p), O._ptr(lab, O.u32p), O._ptr(cnt, O.u64p), O._ptr(w, O.f64p))
    R.ref_eqb_free(h)
    classes = sorted((lab[int(rp[i]):int(rp[i + 1])].tolist(), int(cnt[i])) for i in range(E))
    json.dump({"adds": adds, "classes": classes, "total": int(cnt.sum())}, open(os.path.join(OUT, "eqbuilder.json"), "w"))

    # sample_data
    tmp = tempfile.mkdtemp()
    tarfile.open(os.path.join(REF, "sample_data.tgz")).extractall(tmp)
    names, seqs = read_fasta(os.path.join(tmp, "sample_data", "transcripts.fasta"))
    r1 = read_fastq(os.path.join(tmp, "sample_data", "reads_1.fastq"))
    r2 = read_fastq(os.path.join(tmp, "sample_data", "reads_2.fastq"))
    ix = O.Index(seqs, k=31)
    b1, o1 = O.pack_reads(r1); b2, o2 = O.pack_reads(r2)
    run = O.Run(ix, O.MapOpts.default(O.parse_libtype("IU")))
    run.map_batch(b1, o1, b2, o2)
    res = run.finish()
    txp_len = np.array([len(s) for s in seqs], np.uint32)
    eff = O.eff_lens(txp_len, res["fld"])
    num_mapped = int(res["counters"][1])
    out = dict(txp_seq=np.frombuffer("".join(seqs).encode(), np.uint8), txp_len=txp_len,
               names=np.array(names), reads1=np.frombuffer(b1, np.uint8), off1=o1, reads2=np.frombuffer(b2, np.uint8), off2=o2,
               counters=res["counters"], fld=res["fld"], row_ptr=res["row_ptr"], labels=res["labels"], counts=res["counts"],
               eff=eff, num_mapped=np.uint64(num_mapped))
    for vb in (0, 1):
        ref = O.RefEM(txp_len, eff, res["row_ptr"], res["labels"], res["counts"], num_mapped, use_vb=bool(vb))
        rc, est, mass = ref.optimize()
        assert rc == 0
        out["ref_est_vb%d" % vb] = est
        out["ref_mass_vb%d" % vb] = mass
    np.savez_compressed(os.path.join(OUT, "sample_data.npz"), **out)

    # synthetic classes
    T = 2000
    rp, lab, cnt = synth.make_classes(T, 3000, seed=11, long_frac=0.01)
    txp_len = rng.integers(200, 5000, size=T).astype(np.uint32)
    eff = O.eff_lens(txp_len, None, single_end=True)
    nm = int(cnt.sum())
    out = dict(txp_len=txp_len, eff=eff, row_ptr=rp, labels=lab, counts=cnt, num_mapped=np.uint64(nm))
    for vb in (0, 1):
        ref = O.RefEM(txp_len, eff, rp, lab, cnt, nm, use_vb=bool(vb))
        rc, est, mass = ref.optimize()
        assert rc == 0
        out["ref_est_vb%d" % vb] = est
    np.savez_compressed(os.path.join(OUT, "synth_em.npz"), **out)
    print("golden fixtures written to", OUT)
    subprocess.call(["ls", "-la", OUT])


if __name__ == "__main__" and "--bias-only" not in sys.argv and "--samplers-only" not in sys.argv:
    make_bias_fixture()

if __name__ == "__main__":
    main()
