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
