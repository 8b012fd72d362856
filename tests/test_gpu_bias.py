"""GPU parity test of the device-side bias / GC effective-length correction (sailfish_b200/csrc/bias.cu, sfb200_bias_eff_lens;
SURVEY 8a row A18) against the CPU oracle, which is pinned to the reference's own updateEffectiveLengths
(tests/test_oracle_bias.py).  First green run on a B200: profiles/r01f_experimental_gpu.txt.  The entry point is not called by
the quantification pipeline yet."""
import numpy as np
import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,gc_samp", [(1, 1), (2, 1), (2, 3)])
def test_bias_eff_lens_matches_oracle(ctx, mode, gc_samp):
    rng = np.random.default_rng(17 + mode)
    T = 300
    lens = rng.integers(100, 3000, size=T)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(n), p=[0.3, 0.2, 0.2, 0.3]).astype(np.uint8)) for n in lens]
    ctx.index_build(seqs=seqs, k=31)
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    cdf, mx = O.fld_cdf(fld)
    eff_model = np.where(lens - 189.0 >= 1, lens - 189.0, lens).astype(np.float64)
    eff_in = eff_model * rng.uniform(0.9, 1.1, size=T)
    alphas = rng.lognormal(3, 2, size=T); alphas[rng.random(T) < 0.2] = 0.0; alphas[3] = 5e-9
    rb = rng.integers(1, 3000, size=4096).astype(np.uint32); og = rng.integers(1, 8000, size=101).astype(np.uint32)
    got = ctx.bias_eff_lens(mode, eff_model, eff_in, alphas, 61234, 58766, rb, og, cdf, mx, gc_samp=gc_samp)
    rc, want = O.update_eff_lens(mode, seqs, eff_model, eff_in, alphas, 61234, 58766, rb, og, fld, gc_samp=gc_samp)
    assert rc == 0
    np.testing.assert_allclose(got, want, rtol=1e-9)
    assert ((got != eff_in) == (want != eff_in)).all() and (want != eff_in).sum() > 50
    # no strand tallies: the correction is skipped
    assert (ctx.bias_eff_lens(mode, eff_model, eff_in, alphas, 0, 0, rb, og, cdf, mx) == eff_in).all()


# ---- the optimizer with the correction inside (sfb200_em_run_bias): effective lengths recomputed at iterations 50 / 500 / 1000 ------------
# The launch scheduling is also checked on CPU (tests/em_segments_test.cpp).  First B200 run: profiles/r02a_experimental_gpu.txt.
import os

from sailfish_b200 import capi, synth



@pytest.mark.parametrize("gc_samp", [1, 3, 7])
def test_bias_eff_lens_sliding_gc_passes(ctx, monkeypatch, gc_samp):
    """SFB200_BIAS_GC_SLIDE=1: the fragment GC passes with one thread per fragment length (k_bias_*_gc_slide; CPU check of the same text:
    tests/bias_core_test.cpp) against the oracle and against the default kernels"""
    rng = np.random.default_rng(29)
    T = 300
    lens = rng.integers(100, 3000, size=T)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(n), p=[0.3, 0.2, 0.2, 0.3]).astype(np.uint8)) for n in lens]
    ctx.index_build(seqs=seqs, k=31)
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    cdf, mx = O.fld_cdf(fld)
    eff_model = np.where(lens - 189.0 >= 1, lens - 189.0, lens).astype(np.float64)
    eff_in = eff_model * rng.uniform(0.9, 1.1, size=T)
    alphas = rng.lognormal(3, 2, size=T); alphas[rng.random(T) < 0.2] = 0.0
    rb = rng.integers(1, 3000, size=4096).astype(np.uint32); og = rng.integers(1, 8000, size=101).astype(np.uint32)
    plain = ctx.bias_eff_lens(2, eff_model, eff_in, alphas, 61234, 58766, rb, og, cdf, mx, gc_samp=gc_samp)
    monkeypatch.setenv("SFB200_BIAS_GC_SLIDE", "1")
    got = ctx.bias_eff_lens(2, eff_model, eff_in, alphas, 61234, 58766, rb, og, cdf, mx, gc_samp=gc_samp)
    rc, want = O.update_eff_lens(2, seqs, eff_model, eff_in, alphas, 61234, 58766, rb, og, fld, gc_samp=gc_samp)
    assert rc == 0
    np.testing.assert_allclose(got, want, rtol=1e-9)
    np.testing.assert_allclose(got, plain, rtol=1e-9)
    assert ((got != eff_in) == (want != eff_in)).all() and (want != eff_in).sum() > 50


@pytest.mark.parametrize("loops", ["default", "scatter", "steps"])
@pytest.mark.parametrize("mode,vb,kw", [
    (1, 0, {}), (2, 0, {}), (1, 1, {}), (2, 1, {}),                              # default limits: one recomputation, at iteration 50
    (1, 0, {"fixed_iters": 49}), (1, 0, {"fixed_iters": 50}), (2, 0, {"fixed_iters": 51}),
    (1, 0, {"fixed_iters": 520}), (2, 1, {"fixed_iters": 1003}),                # two and three recomputations
    (2, 0, {"min_iter": 10, "max_iter": 30}), (1, 0, {"min_iter": 600, "max_iter": 40}),
])
def test_em_run_bias_matches_oracle(ctx, monkeypatch, loops, mode, vb, kw):
    if loops == "scatter":
        monkeypatch.setenv("SFB200_EM_GATHER", "0"); monkeypatch.setenv("SFB200_EM_DENSE", "0")
    elif loops == "steps":
        monkeypatch.setenv("SFB200_EM_MODE", "steps")
    rng = np.random.default_rng(100 * mode + vb)
    T = 400
    lens = rng.integers(250, 3000, size=T)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(n), p=[0.3, 0.2, 0.2, 0.3]).astype(np.uint8)) for n in lens]
    ctx.index_build(seqs=seqs, k=31)
    rp, lab, cnt = synth.make_classes(T, 1000, seed=5 + mode)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    cdf, mx = O.fld_cdf(fld)
    eff = np.where(lens - 189.0 >= 1, lens - 189.0, lens).astype(np.float64)
    rb = rng.integers(1, 3000, size=4096).astype(np.uint32); og = rng.integers(1, 8000, size=101).astype(np.uint32)
    nf, nr = 61234, 58766
    a, eff_got, it, mrd = ctx.em_run_bias(mode, eff, nm, nf, nr, rb, og, cdf, mx, opts=capi.EMOpts.default(use_vb=vb, **kw))
    rc, want, eff_want, it_o, mrd_o = O.em_run_bias(mode, seqs, rp, lab, cnt, eff, nm, nf, nr, rb, og, fld, opts=O.EMOpts.default(use_vb=vb, **kw))
    assert rc == 0 and it == it_o
    np.testing.assert_allclose(eff_got, eff_want, rtol=1e-6)
    np.testing.assert_allclose(a, want, rtol=1e-4, atol=1e-6)
    assert (a == 0).tolist() == (want == 0).tolist()
    if it > 50:                                                                 # the recomputation sits at the TOP of iteration 50 (:822)
        assert (eff_got != np.maximum(eff, 1.0)).sum() > 50                     # the correction did happen
        assert abs(mrd - mrd_o) <= 1e-5 * max(abs(mrd_o), 1e-12)
    else:
        assert (eff_got == np.maximum(eff, 1.0)).all()
