"""CPU pins for the bias / GC effective-length correction ("next" row N3): the oracle restatement
(oracle/orc_bias.cpp) against the committed outputs of the reference's OWN updateEffectiveLengths (src/SailfishUtils.cpp:611-926,
compiled unmodified into oracle/_ref/libsfref_em.so; fixture written by tests/golden/make_golden.py) and, where oracle/_ref is
present, against that function live on fresh random inputs.  No CUDA kernel consumes this yet: the device side of row A18 is
not built; this is the checker it will be held to."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bias_efflens.npz")


def split(seq, lens):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return [seq[off[i]:off[i + 1]].tobytes() for i in range(len(lens))]


@pytest.mark.parametrize("mode,tag,samp", [(1, "seq", 1), (2, "gc", 1), (2, "gc", 3)])
def test_oracle_matches_reference_golden(mode, tag, samp):
    d = dict(np.load(GOLDEN))
    seqs = split(d["seq"], d["txp_len"])
    rc, out = O.update_eff_lens(mode, seqs, d["eff_model"], d["eff_in"], d["alphas"], int(d["num_fwd"]), int(d["num_rc"]),
                                d["read_bias"], d["observed_gc"], d["fld"], gc_samp=samp)
    assert rc == 0
    ref = d["ref_%s_samp%d" % (tag, samp)]
    np.testing.assert_allclose(out, ref, rtol=1e-12)                      # Eigen's .sum() adds in another order
    assert ((out != d["eff_in"]) == (ref != d["eff_in"])).all()           # the same transcripts were corrected
    assert (ref != d["eff_in"]).sum() > 10
    # transcripts shorter than the 6-mer window, with alpha < 1e-8 or alpha == 0 keep their length
    assert out[0] == d["eff_in"][0] and out[1] == d["eff_in"][1] and out[5] == d["eff_in"][5]
    assert (out[d["alphas"] == 0] == d["eff_in"][d["alphas"] == 0]).all()
    rc, out2 = O.update_eff_lens(mode, seqs, d["eff_model"], ref, d["alphas"] * 1.5, int(d["num_fwd"]), int(d["num_rc"]),
                                 d["read_bias"], d["observed_gc"], d["fld"], gc_samp=samp)
    np.testing.assert_allclose(out2, d["ref_%s_samp%d_round2" % (tag, samp)], rtol=1e-12)


@pytest.mark.parametrize("mode,tag", [(1, "seq"), (2, "gc")])
@pytest.mark.parametrize("vb", [0, 1])
def test_optimizer_with_correction_matches_reference_golden(mode, tag, vb):
    """optimize() with --biasCorrect / --gcBiasCorrect: lengths recomputed at iterations 50 / 500 / 1000 and the class weights
    rebuilt (CollapsedEMOptimizer.cpp:816-840); estimates and final effective lengths of the reference's own optimizer"""
    d = dict(np.load(GOLDEN))
    seqs = split(d["seq"], d["txp_len"])
    rc, a, eff, it, _ = O.em_run_bias(mode, seqs, d["row_ptr"], d["labels"], d["counts"], d["eff_model"], int(d["num_mapped"]),
                                      int(d["num_fwd"]), int(d["num_rc"]), d["read_bias"], d["observed_gc"], d["fld"],
                                      opts=O.EMOpts.default(use_vb=vb, tol=1e-5))
    assert rc == 0 and it > 50
    np.testing.assert_allclose(a, d["opt_%s_vb%d_est" % (tag, vb)], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(eff, d["opt_%s_vb%d_eff" % (tag, vb)], rtol=1e-12)
    assert (eff != np.maximum(d["eff_model"], 1.0)).sum() > 5               # the correction did change lengths
    # and it matters: without it the estimates differ
    rc0, a0, _, _ = O.em_run(len(seqs), d["row_ptr"], d["labels"], d["counts"], d["eff_model"], int(d["num_mapped"]), O.EMOpts.default(use_vb=vb, tol=1e-5))
    assert rc0 == 0 and np.max(np.abs(a0 - a)) > 1e-3


def test_degenerate_inputs():
    d = dict(np.load(GOLDEN))
    seqs = split(d["seq"], d["txp_len"])
    # no strand tallies: "Had no fragments from which to estimate fwd vs. rev-comp mapping rate" -> lengths unchanged (:627-632)
    rc, out = O.update_eff_lens(1, seqs, d["eff_model"], d["eff_in"], d["alphas"], 0, 0, d["read_bias"], d["observed_gc"], d["fld"])
    assert rc == 0 and (out == d["eff_in"]).all()
    # a base other than A/C/G/T is refused (the reference would index its tables with UINT32_MAX)
    bad = list(seqs); bad[7] = bad[7][:20] + b"N" + bad[7][21:]
    rc, _ = O.update_eff_lens(1, bad, d["eff_model"], d["eff_in"], d["alphas"], 5, 5, d["read_bias"], d["observed_gc"], d["fld"])
    assert rc == -2


def test_oracle_matches_reference_live():
    R = O.ref_em()
    if R is None or not hasattr(R, "ref_bias_session"):
        pytest.skip("oracle/_ref/libsfref_em.so with the bias correction is not available here")
    rng = np.random.default_rng(99)
    T = 30
    lens = rng.integers(100, 1500, size=T)
    seqs = [bytes(rng.choice(list(b"ACGTacgt"), size=int(l)).astype(np.uint8)) for l in lens]       # lower case as well
    x = np.arange(800)
    fld = np.round(5000 * np.exp(-0.5 * ((x - 250) / 40.0) ** 2)).astype(np.uint32)
    eff_model = np.maximum(lens - 249.0, 1.0)
    alphas = rng.lognormal(2, 2, size=T)
    rb = rng.integers(1, 500, size=4096).astype(np.uint32); og = rng.integers(1, 900, size=101).astype(np.uint32)
    for mode in (1, 2):
        Rb = O.RefBias(mode, seqs, eff_model, rb, og, fld, 1000, 3000, gc_samp=2)
        rc_r, ref = Rb.update(alphas, eff_model)
        rc, out = O.update_eff_lens(mode, seqs, eff_model, eff_model, alphas, 1000, 3000, rb, og, fld, gc_samp=2)
        assert rc == 0 and rc_r == 0
        np.testing.assert_allclose(out, ref, rtol=1e-12)


def test_fld_cdf_matches_reference_table():
    d = dict(np.load(GOLDEN))
    cdf, mx = O.fld_cdf(d["fld"])
    assert mx == int(d["ref_fld_max"])
    full = np.ones(1000, np.float32); full[:len(cdf)] = cdf
    assert (full == d["ref_fld_cdf"]).all()                                # float table of the reference's EmpiricalDistribution, bit for bit


def _rc(b):
    return bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(b))


def test_sample_collection_matches_reference_helpers():
    """read-start 6-mer contexts and the observed fragment GC histogram collected while mapping (SailfishQuantify.cpp:255-287,
    372-389, 555-583): the oracle mapper against the reference's own ReadKmerDist<6>::update and Transcript::gcFrac applied to the
    known origin of every read (single-isoform genes: every read has exactly one hit)"""
    R = O.ref_em()
    if R is None or not hasattr(R, "ref_readbias_update"):
        pytest.skip("oracle/_ref/libsfref_em.so with the bias correction is not available here")
    rng = np.random.default_rng(7)
    T, L = 12, 60
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(n)).astype(np.uint8)) for n in rng.integers(300, 900, size=T)]
    ix = O.Index(seqs, k=31)
    # ---- single-end: forward and reverse-complement reads, including starts at the very ends of a transcript
    reads, want = [], np.ones(4096, np.int64)
    n_samples = 0
    budget = 150
    for i in range(400):
        t = int(rng.integers(0, T)); n = len(seqs[t])
        pos = int(rng.choice([0, 1, 2, 3, n - L, n - L - 1, n - L - 2, int(rng.integers(0, n - L + 1))]))
        fwd = bool(rng.integers(0, 2))
        frag = seqs[t][pos:pos + L]
        reads.append(frag if fwd else _rc(frag))
        start = pos if fwd else pos + L
        idx = -1
        if 0 < start < n:
            idx = R.ref_readbias_update(seqs[t], n, start, int(fwd))
        if idx >= 0 and n_samples < budget:
            want[idx] += 1; n_samples += 1
    b = np.frombuffer(b"".join(reads), np.uint8); off = np.arange(len(reads) + 1, dtype=np.uint64) * np.uint64(L)
    run = O.Run(ix, O.MapOpts.default(O.parse_libtype("U")))
    run.set_bias(True, False, budget)
    run.map_batch(b.tobytes(), off, n_threads=3)
    res = run.finish()
    assert int(res["counters"][1]) == len(reads)
    rb, og = run.finish_bias()
    assert rb.tolist() == want.tolist() and n_samples == budget
    assert (og == 1).all()                                                # no fragment GC from single-end reads
    # ---- paired-end: inward pairs; the GC observation uses gcFrac(start, start + fragLen) when it lies inside the transcript
    r1, r2, want_gc = [], [], np.ones(101, np.int64)
    want_b = np.ones(4096, np.int64)
    for i in range(300):
        t = int(rng.integers(0, T)); n = len(seqs[t])
        fl = int(rng.integers(L + 20, 260))
        fs = int(rng.choice([0, 1, n - fl, n - fl - 1, int(rng.integers(0, n - fl + 1))]))
        left, right = seqs[t][fs:fs + L], _rc(seqs[t][fs + fl - L:fs + fl])
        flip = bool(rng.integers(0, 2))
        r1.append(right if flip else left); r2.append(left if flip else right)
        # mate 1 decides the bias sample: forward mate 1 starts at its position, a reverse-complemented one at position + length
        pos1, fwd1 = (fs + fl - L, False) if flip else (fs, True)
        start1 = pos1 if fwd1 else pos1 + L
        if 0 < start1 < n:
            idx = R.ref_readbias_update(seqs[t], n, start1, int(fwd1))
            if idx >= 0:
                want_b[idx] += 1
        if fs > 0 and fs + fl < n:
            want_gc[R.ref_gc_frac(seqs[t], n, fs, fs + fl)] += 1
    b1 = np.frombuffer(b"".join(r1), np.uint8); b2 = np.frombuffer(b"".join(r2), np.uint8)
    off = np.arange(len(r1) + 1, dtype=np.uint64) * np.uint64(L)
    run = O.Run(ix, O.MapOpts.default(O.parse_libtype("IU")))
    run.set_bias(True, True, 10 ** 6)
    run.map_batch(b1.tobytes(), off, b2.tobytes(), off, n_threads=2)
    res = run.finish()
    assert int(res["counters"][1]) == len(r1)
    rb, og = run.finish_bias()
    assert og.tolist() == want_gc.tolist()
    assert rb.tolist() == want_b.tolist()
