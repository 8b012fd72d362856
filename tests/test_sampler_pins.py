"""Bootstrap (A16) and Gibbs (A17) pinned to the reference's OWN samplers.

The reference seeds mt19937 from std::random_device (CollapsedEMOptimizer.cpp:657, CollapsedGibbsSampler.cpp:241), so replicate-for-
replicate equality does not exist even between two runs of the reference.  What is pinned:
  * exact invariants -- every replicate conserves the fragment total (EM bootstrap: sum == numMappedFragments; Gibbs: integer sum);
  * the distribution -- per-transcript mean and variance over many replicates, against the moments of the reference's own
    gatherBootstraps (src/CollapsedEMOptimizer.cpp:557-709) / CollapsedGibbsSampler::sample (src/CollapsedGibbsSampler.cpp:199-291,
    including `bool numInternalRounds = 10` at :248, i.e. ONE internal round per sample) committed in tests/golden/ref_samplers.npz
    (tests/golden/make_golden.py --samplers-only), and -- where oracle/_ref is built -- against a live run of the same code.
CPU tests pin the oracle (oracle/orc_em.cpp); the gpu tests pin the device path through the C ABI.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_samplers.npz")


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(GOLD))


def z_mean(mean_a, var_a, n_a, mean_b, var_b, n_b, ess_a=1.0, ess_b=1.0):
    """z-scores of the difference of two sample means; ess_* < 1 shrinks the effective sample count of autocorrelated chains"""
    se = np.sqrt(var_a / (n_a * ess_a) + var_b / (n_b * ess_b))
    return np.abs(mean_a - mean_b) / np.maximum(se, 1e-9)


def check_moments(rows, ref_mean, ref_var, ref_n, sel, ess=1.0, ref_ess=1.0, var_lo=0.6, var_hi=1.7, chain=False):
    rows = np.asarray(rows, np.float64)
    n = rows.shape[0]
    # a short autocorrelated chain underestimates its own variance: size both standard errors with the long reference chain's
    z = z_mean(rows.mean(axis=0), ref_var if chain else rows.var(axis=0), n, ref_mean, ref_var, ref_n, ess, ref_ess)[sel]
    assert np.mean(z < 3.0) >= 0.95 and z.max() < 6.0, (np.mean(z < 3.0), z.max())
    # spreads: the ratio of variances of the selected transcripts concentrates around one
    vr = (rows.var(axis=0)[sel] + 1.0) / (ref_var[sel] + 1.0)
    assert var_lo < np.median(vr) < var_hi, np.median(vr)
    assert np.mean((vr > var_lo / 2) & (vr < var_hi * 2)) > 0.95, vr


def check_gibbs_window(rows, fx):
    """Chains started at the EM estimate mix slowly for some transcripts, so a chain is compared over the SAME window (samples
    burn .. n) with the spread of that window's mean over 64 independent runs of the reference's own sampler."""
    n, burn, K = int(fx["gibbs_n"]), int(fx["gibbs_burn"]), int(fx["gibbs_chains"])
    assert rows.shape[0] == n
    w = rows[burn:].astype(np.float64)
    sel = fx["gibbs_wmean_mean"] > 100
    assert sel.sum() > 40
    sd = np.maximum(fx["gibbs_wmean_sd"], 0.5) * np.sqrt(1.0 + 1.0 / K)
    z = (np.abs(w.mean(axis=0) - fx["gibbs_wmean_mean"]) / sd)[sel]
    assert np.mean(z < 3.0) >= 0.90 and z.max() < 7.0, (np.mean(z < 3.0), z.max())
    # window variances: the same comparison, on the log scale (they are heavy-tailed for the slow mixers)
    vsd = np.maximum(fx["gibbs_wvar_sd"], 0.05 * fx["gibbs_wvar_mean"] + 1.0) * np.sqrt(1.0 + 1.0 / K)
    zv = (np.abs(w.var(axis=0) - fx["gibbs_wvar_mean"]) / vsd)[sel]
    assert np.mean(zv < 3.5) >= 0.88, np.mean(zv < 3.5)
    vr = (w.var(axis=0)[sel] + 1.0) / (fx["gibbs_wvar_mean"][sel] + 1.0)
    assert 0.7 < np.median(vr) < 1.4, np.median(vr)


# --------------------------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("vb", [0, 1])
def test_oracle_bootstrap_matches_reference_moments(fx, vb):
    T = len(fx["txp_len"])
    nm = int(fx["num_mapped"])
    n = 200
    rc, rows = O.bootstrap(T, fx["row_ptr"], fx["labels"], fx["counts"], fx["eff"], n, seed=7, opts=O.EMOpts.default(use_vb=vb))
    assert rc == 0 and rows.shape == (n, T)
    if not vb:
        np.testing.assert_allclose(rows.sum(axis=1), nm, rtol=1e-9)                    # exact total conservation
    else:
        # VBEM replicates carry the prior's mass as the reference's do
        assert abs(rows.sum(axis=1).mean() - fx["boot_vb1_sum"].mean()) < 0.05
    sel = fx["boot_vb%d_mean" % vb] > 50
    assert sel.sum() > 50
    check_moments(rows, fx["boot_vb%d_mean" % vb], fx["boot_vb%d_var" % vb], int(fx["boot_n"]), sel)
    # transcripts the reference never gives mass to stay empty (VBEM: a rare second mode may appear in a replicate)
    dead = fx["boot_vb%d_mean" % vb] == 0
    assert np.mean(rows[:, dead] > 0) <= (0.001 if vb else 0.0)


def test_oracle_gibbs_matches_reference_moments(fx):
    T = len(fx["txp_len"])
    nm = int(fx["num_mapped"])
    est = fx["em_est"]
    rc, rows = O.gibbs(T, fx["row_ptr"], fx["labels"], fx["counts"], fx["eff"], est / est.sum(), nm, int(fx["gibbs_n"]), seed=3)
    assert rc == 0 and rows.dtype == np.int32
    assert (rows.sum(axis=1) == nm).all() and (rows >= 0).all()
    check_gibbs_window(rows, fx)
    inactive = np.ones(T, bool); inactive[fx["labels"]] = False
    assert (rows[:, inactive] == 0).all()


def test_fixture_matches_live_reference(fx):
    """the committed moments are what the reference TU compiled here produces today (skipped where oracle/_ref is absent)"""
    if O.ref_em() is None:
        pytest.skip("oracle/_ref not built")
    T = len(fx["txp_len"])
    nm = int(fx["num_mapped"])
    ref = O.RefEM(fx["txp_len"], fx["eff"], fx["row_ptr"], fx["labels"], fx["counts"], nm, n_boot=150)
    rc, rows = ref.bootstraps()
    assert rc == 0
    np.testing.assert_allclose(rows.sum(axis=1), nm, rtol=1e-9)
    check_moments(rows, fx["boot_vb0_mean"], fx["boot_vb0_var"], int(fx["boot_n"]), fx["boot_vb0_mean"] > 50)
    ref = O.RefEM(fx["txp_len"], fx["eff"], fx["row_ptr"], fx["labels"], fx["counts"], nm)
    rc, est, _ = ref.optimize()
    assert rc == 0
    np.testing.assert_allclose(est, fx["em_est"], rtol=1e-12)
    rc, g = ref.gibbs(int(fx["gibbs_n"]))
    assert rc == 0 and (g.sum(axis=1) == nm).all()
    check_gibbs_window(g, fx)


# --------------------------------------------------------------------------------------------------------------- GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("vb", [0, 1])
def test_gpu_bootstrap_matches_reference_moments(ctx, fx, vb):
    from sailfish_b200 import capi
    T = len(fx["txp_len"])
    nm = int(fx["num_mapped"])
    ctx.eq_import(T, fx["row_ptr"], fx["labels"], fx["counts"])
    n = 240
    rows = ctx.bootstrap_run(fx["eff"], n, seed=11, opts=capi.EMOpts.default(use_vb=vb))
    assert rows.shape == (n, T)
    if not vb:
        np.testing.assert_allclose(rows.sum(axis=1), nm, rtol=1e-9)
    else:
        assert abs(rows.sum(axis=1).mean() - fx["boot_vb1_sum"].mean()) < 0.05
    sel = fx["boot_vb%d_mean" % vb] > 50
    check_moments(rows, fx["boot_vb%d_mean" % vb], fx["boot_vb%d_var" % vb], int(fx["boot_n"]), sel)
    dead = fx["boot_vb%d_mean" % vb] == 0
    assert np.mean(rows[:, dead] > 0) <= (0.001 if vb else 0.0)


@pytest.mark.gpu
def test_gpu_gibbs_matches_reference_moments(ctx, fx):
    T = len(fx["txp_len"])
    nm = int(fx["num_mapped"])
    ctx.eq_import(T, fx["row_ptr"], fx["labels"], fx["counts"])
    alphas, _, _ = ctx.em_run(fx["eff"], nm)
    np.testing.assert_allclose(alphas, fx["em_est"], rtol=1e-6, atol=1e-6)
    rows = ctx.gibbs_run(fx["eff"], alphas / alphas.sum(), nm, int(fx["gibbs_n"]), seed=5)
    assert rows.dtype == np.int32 and (rows >= 0).all()
    assert (rows.sum(axis=1) == nm).all()
    check_gibbs_window(rows, fx)
    inactive = np.ones(T, bool); inactive[fx["labels"]] = False
    assert (rows[:, inactive] == 0).all()
