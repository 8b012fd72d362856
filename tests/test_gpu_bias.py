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
