"""GPU parity tests (-m gpu) for the inference kernels, called through the C ABI (libsfb200.so).

Bar (BASELINE.json north_star): TPM / NumReads within 1e-4 relative of the reference CPU CollapsedEMOptimizer on identical
inputs, compared at equal iteration count (SURVEY section 7 "parity at a convergence threshold").
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from sailfish_b200 import capi, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4          # north_star tolerance
ATOL = 1e-6          # counts below this are noise next to the 1e-8 truncation threshold


def assert_close(got, want, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def test_device_xxh64_known_answers(ctx, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "xxh64_kat.json")))
    kat = [e for e in g["xxh64"] if len(e["hex"]) % 8 == 0]            # labels are whole 32-bit words
    for seed in (0, 1, 0x9E3779B97F4A7C15):
        es = [e for e in kat if e["seed"] == seed]
        out = ctx.xxh64([bytes.fromhex(e["hex"]) for e in es], seed)
        assert ["%016x" % int(x) for x in out] == [e["xxh64"] for e in es]
    tg = g["transcript_group"]
    out = ctx.xxh64([np.array(e["ids"], np.uint32).tobytes() for e in tg], 0)
    assert ["%016x" % int(x) for x in out] == [e["hash"] for e in tg]


def test_device_digamma(ctx):
    from scipy.special import digamma
    xs = np.concatenate([10.0 ** np.linspace(-300, 9, 400), np.linspace(0.001, 30, 500)])
    got = ctx.digamma(xs)
    want = digamma(xs)
    ok = (np.abs(got - want) <= 1e-12 * np.abs(want)) | (np.abs(got - want) < 1e-13)
    assert ok.all()


def test_device_exp_digamma(ctx):
    """exp(digamma(x)) as the VBEM kernels evaluate it (csrc/vb_math.hpp: series without the logarithm) against scipy; the host build
    of the same header is checked against mpmath at 50 digits in tests/test_vb_math.py"""
    from scipy.special import digamma
    xs = np.concatenate([10.0 ** np.linspace(-2.5, 9, 600), np.linspace(0.002, 40, 800), [15.999999, 16.0, 16.000001]])
    got = ctx.exp_digamma(xs)
    want = np.exp(digamma(xs))
    assert (np.abs(got - want) <= 2e-12 * want).all()
    assert (ctx.exp_digamma(np.array([1e-3, 5e-4, 1e-9])) == 0.0).all()         # a transcript without reads: exactly 0, as exp(-1000 - logNorm)


@pytest.mark.parametrize("name", ["sample_data", "synth_em"])
@pytest.mark.parametrize("vb", [0, 1])
def test_em_matches_reference_optimizer_golden(ctx, name, vb, sample_data, synth_em):
    """GPU optimize() vs the estimates of the reference's own CollapsedEMOptimizer (committed fixture)."""
    d = sample_data if name == "sample_data" else synth_em
    T = len(d["txp_len"])
    ctx.eq_import(T, d["row_ptr"], d["labels"], d["counts"])
    alphas, iters, mrd = ctx.em_run(d["eff"], int(d["num_mapped"]), capi.EMOpts.default(use_vb=vb))
    rc, want, it_o, mrd_o = O.em_run(T, d["row_ptr"], d["labels"], d["counts"], d["eff"], int(d["num_mapped"]),
                                     O.EMOpts.default(use_vb=vb))
    assert rc == 0
    assert iters == it_o, "GPU stopped at a different iteration than the oracle"
    assert_close(alphas, want)
    assert_close(alphas, d["ref_est_vb%d" % vb])
    assert (alphas == 0).tolist() == (want == 0).tolist()
    assert abs(mrd - mrd_o) <= 1e-6 * max(abs(mrd_o), 1e-12)
    nm = int(d["num_mapped"])
    assert_close(capi.tpm(alphas, d["eff"], nm), O.tpm(want, d["eff"], nm), atol=1e-4)


@pytest.mark.parametrize("vb", [0, 1])
@pytest.mark.parametrize("fixed", [1, 10, 50, 333])
def test_em_fixed_iterations(ctx, vb, fixed):
    T = 5000
    rp, lab, cnt = synth.make_classes(T, 12000, seed=21 + fixed, long_frac=0.01)
    rng = np.random.default_rng(fixed)
    eff = rng.uniform(0.5, 4000, size=T)           # includes lengths below 1 (clamped, CollapsedEMOptimizer.cpp:736-738)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    alphas, iters, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb, fixed_iters=fixed))
    rc, want, it_o, _ = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, fixed_iters=fixed))
    assert rc == 0 and iters == fixed == it_o
    assert_close(alphas, want)
    assert abs(alphas.sum() - want.sum()) <= 1e-9 * want.sum()


def test_em_steps_mode_equals_persistent(ctx, monkeypatch):
    """one launch per phase (the multi-rank path) == the persistent cooperative kernel"""
    T = 3000
    rp, lab, cnt = synth.make_classes(T, 7000, seed=5)
    eff = np.random.default_rng(1).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        a1, it1, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))          # CTA-partitioned loop (em_part.cuh)
        monkeypatch.setenv("SFB200_EM_MODE", "steps")
        a2, it2, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        monkeypatch.delenv("SFB200_EM_MODE")
        assert it1 == it2
        assert_close(a1, a2, rtol=1e-7)      # same iterates up to rounding (the gather-form loop sums in a different order)
    # the binned-layout persistent kernel (what runs when a CTA's slice does not fit in shared memory)
    monkeypatch.setenv("SFB200_NO_PARTITION", "1")
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        a3, it3, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        rc, want, it_o, _ = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb))
        assert it3 == it_o
        assert_close(a3, want)
    monkeypatch.delenv("SFB200_NO_PARTITION")
    ctx.eq_import(T, rp, lab, cnt)


def test_em_partition_with_pool(ctx):
    """classes that cross CTA ranges (and everything they touch) go through the global-memory pool; the rest stays in
    shared memory: both halves together must reproduce the oracle"""
    T = 20000
    rp, lab, cnt = synth.make_classes(T, 30000, seed=77, max_len=5, long_frac=0.002)
    eff = np.random.default_rng(8).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        for kw in (dict(), dict(fixed_iters=77)):
            a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb, **kw))
            rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, **kw), n_threads=4)
            assert rc == 0 and it == it_o
            assert_close(a, want)
            assert abs(mrd - mrd_o) <= 1e-6 * max(abs(mrd_o), 1e-12)


def test_em_edge_cases(ctx):
    # a single class with a single transcript; transcripts that appear in no class stay 0
    rp = np.array([0, 1], np.uint64); lab = np.array([3], np.uint32); cnt = np.array([17], np.uint64)
    ctx.eq_import(6, rp, lab, cnt)
    a, it, _ = ctx.em_run(np.full(6, 100.0), 17)
    assert a.tolist() == [0, 0, 0, 17.0, 0, 0]
    # duplicate transcript ids inside a label (orphans mapping both mates to one transcript)
    rp = np.array([0, 2, 5], np.uint64); lab = np.array([1, 1, 0, 1, 2], np.uint32); cnt = np.array([10, 30], np.uint64)
    ctx.eq_import(3, rp, lab, cnt)
    eff = np.array([100.0, 200.0, 300.0])
    a, it, _ = ctx.em_run(eff, 40)
    rc, want, it_o, _ = O.em_run(3, rp, lab, cnt, eff, 40)
    assert it == it_o
    assert_close(a, want)
    # no classes at all: "no transcripts are expressed" (CollapsedEMOptimizer.cpp:794-798)
    ctx.eq_import(4, np.array([0], np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint64))
    with pytest.raises(capi.Sfb200Error) as e:
        ctx.em_run(np.full(4, 10.0), 0)
    assert e.value.code == -4
    # label with an out-of-range transcript id is rejected
    with pytest.raises(capi.Sfb200Error):
        ctx.eq_import(2, np.array([0, 1], np.uint64), np.array([7], np.uint32), np.array([1], np.uint64))


def test_eq_import_export_roundtrip(ctx):
    rp, lab, cnt = synth.make_classes(1000, 2500, seed=9, long_frac=0.02)
    ctx.eq_import(1000, rp, lab, cnt)
    rp2, lab2, cnt2 = ctx.eq_export()
    assert rp2.tolist() == rp.tolist() and lab2.tolist() == lab.tolist() and cnt2.tolist() == cnt.tolist()


@pytest.mark.parametrize("vb", [0, 1])
def test_bootstrap_em_same_resampled_counts(ctx, vb):
    """doBootstrap's loop (no minimum iteration count, gate on the OLD alpha) on identical resampled counts"""
    T = 2000
    rp, lab, cnt = synth.make_classes(T, 5000, seed=31)
    eff = np.random.default_rng(2).uniform(100, 3000, size=T)
    rng = np.random.default_rng(99)
    total = int(cnt.sum())
    samp = rng.multinomial(total, cnt / cnt.sum()).astype(np.uint64)
    ctx.eq_import(T, rp, lab, cnt)
    a, it = ctx.bootstrap_em(eff, samp, capi.EMOpts.default(use_vb=vb))
    rc, want, it_o = O.bootstrap_em(T, rp, lab, samp, eff, O.EMOpts.default(use_vb=vb))
    assert rc == 0 and it == it_o
    assert_close(a, want)


def test_bootstrap_run_distribution(ctx):
    """Bootstrap replicates: RNG differs from the reference (std::random_device there), so parity is distributional:
    per-transcript mean over replicates matches the oracle's replicates within sampling error, totals are exact."""
    T = 300
    rp, lab, cnt = synth.make_classes(T, 600, seed=41, max_len=4)
    eff = np.full(T, 500.0)
    total = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    rows = ctx.bootstrap_run(eff, 40, seed=7)
    assert rows.shape == (40, T)
    np.testing.assert_allclose(rows.sum(axis=1), total, rtol=1e-9)
    rc, orows = O.bootstrap(T, rp, lab, cnt, eff, 40, seed=7)
    assert rc == 0
    big = orows.mean(axis=0) > 200
    assert big.sum() > 10
    se = np.sqrt(rows.var(axis=0) / 40 + orows.var(axis=0) / 40)
    zscore = np.abs(rows.mean(axis=0) - orows.mean(axis=0))[big] / np.maximum(se[big], 1e-9)
    assert np.mean(zscore < 4) > 0.97
    # reproducible for a fixed seed
    rows2 = ctx.bootstrap_run(eff, 3, seed=7)
    np.testing.assert_allclose(rows2, rows[:3], rtol=1e-9)


def test_em_full_size_properties(ctx):
    """BASELINE config 2 scale (200k transcripts): size-independent properties -- mass conservation and fixed point."""
    T = 200000
    rp, lab, cnt = synth.make_classes(T, 400000, seed=77, max_len=5)
    eff = np.random.default_rng(3).uniform(100, 5000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    a, it, _ = ctx.em_run(eff, nm, capi.EMOpts.default(fixed_iters=200))
    assert it == 200
    assert abs(a.sum() - nm) <= 1e-8 * nm                      # every class hands out exactly its count
    active = np.zeros(T, bool); active[lab] = True
    assert (a[~active] == 0).all()
    rc, want, _, _ = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(fixed_iters=200), n_threads=8)
    assert_close(a, want)


def test_gibbs_invariants_and_distribution(ctx):
    """Collapsed Gibbs (CollapsedGibbsSampler.cpp): the RNG and the class visiting order differ from the reference's
    (random_device-seeded mt19937, libcuckoo bucket order), so parity is distributional: every sample conserves the
    fragment total exactly, and posterior means / spreads agree with the oracle's sequential sampler within sampling error."""
    T = 400
    rp, lab, cnt = synth.make_classes(T, 900, seed=51, max_len=4)
    eff = np.random.default_rng(5).uniform(200, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    alphas, _, _ = ctx.em_run(eff, nm)
    masses = alphas / alphas.sum()
    n = 300
    rows = ctx.gibbs_run(eff, masses, nm, n, seed=3)
    assert rows.shape == (n, T) and rows.dtype == np.int32
    assert (rows >= 0).all()
    assert (rows.sum(axis=1) == nm).all()                       # every round only moves fragments between class members
    inactive = np.ones(T, bool); inactive[lab] = False
    assert (rows[:, inactive] == 0).all()
    rc, orows = O.gibbs(T, rp, lab, cnt, eff, masses, nm, n, seed=3)
    assert rc == 0 and (orows.sum(axis=1) == nm).all()
    burn = 50
    g, o = rows[burn:].astype(np.float64), orows[burn:].astype(np.float64)
    big = o.mean(axis=0) > 100
    assert big.sum() > 20
    # chains are autocorrelated: allow a generous effective-sample-size factor
    se = np.sqrt((g.var(axis=0) + o.var(axis=0)) / (n - burn) * 10.0) + 1.0
    zs = np.abs(g.mean(axis=0) - o.mean(axis=0))[big] / se[big]
    assert np.mean(zs < 4) > 0.95, zs
    # the posterior mean stays close to the EM estimate it was started from
    rel = np.abs(g.mean(axis=0)[big] - alphas[big]) / alphas[big]
    assert np.median(rel) < 0.1
    # reproducible for a fixed seed
    rows2 = ctx.gibbs_run(eff, masses, nm, 5, seed=3)
    assert (rows2 == rows[:5]).all()


def test_binomial_sampler_moments(ctx):
    """the conditional-binomial multinomial used by Gibbs: a 2-member class resampled many times has the right mean/variance"""
    T = 2
    rp = np.array([0, 2], np.uint64); lab = np.array([0, 1], np.uint32); cnt = np.array([100000], np.uint64)
    eff = np.array([1000.0, 1000.0])
    ctx.eq_import(T, rp, lab, cnt)
    masses = np.array([0.3, 0.7])
    rows = ctx.gibbs_run(eff, masses, 100000, 400, seed=11)
    assert (rows.sum(axis=1) == 100000).all()
    # symmetric weights: the chain wanders, but each round redistributes a binomial share; sanity: values stay in range
    assert rows.min() >= 0 and rows.max() <= 100000
