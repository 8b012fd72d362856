"""GPU parity tests (-m gpu) of the atomic-free EM loops through the C ABI: k_em_gather (sailfish_b200/csrc/em_gather.cuh, one
thread per class / per transcript over a transposed layout) and k_em_dense (em_dense.cuh, one thread per connected component).

They run when every multi-member class is local to one CTA's transcript range (gene-local classes: the BASELINE workloads),
k_em_dense additionally needs components of at most 8 transcripts.  Every test runs once per loop, asserts that it is the
kernel that ran (sfb200_last_em_kernel == 2 / 4) and compares it with the oracle (= the reference's CollapsedEMOptimizer
restated) within the north_star tolerance, and with the scatter-form kernels."""
import numpy as np
import pytest

from oracle import pyoracle as O
from sailfish_b200 import capi, synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6
GATHER, DENSE = 2, 4


import os

# "dense-balanced" (lanes per component chosen from its class count, SFB200_EM_DENSE_GROUP=0) is opt-in.  Its layout and iteration
# are checked on CPU (tests/em_dense_layout_test.cpp) and every case below is green with it on a B200 (profiles/r02a_experimental_gpu.txt)
_KINDS = ["gather", "dense", "dense-balanced"]


@pytest.fixture(autouse=True, params=_KINDS)
def loop_kind(request, monkeypatch):
    monkeypatch.setenv("SFB200_EM_GATHER", "1")
    monkeypatch.setenv("SFB200_EM_DENSE", "0" if request.param == "gather" else "1")
    if request.param == "dense-balanced":
        monkeypatch.setenv("SFB200_EM_DENSE_GROUP", "0")
    else:
        monkeypatch.delenv("SFB200_EM_DENSE_GROUP", raising=False)
    return GATHER if request.param == "gather" else DENSE


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("vb", [0, 1])
@pytest.mark.parametrize("T,E", [(5000, 12000), (60000, 150000), (300, 700)])
def test_gather_loop_converges_like_the_oracle(ctx, vb, T, E, loop_kind):
    rp, lab, cnt = synth.make_classes(T, E, seed=T + vb)                    # members stay inside a 5-transcript gene
    eff = np.random.default_rng(T).uniform(0.5, 4000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
    assert ctx.last_em_kernel() == loop_kind
    rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb), n_threads=4)
    assert rc == 0 and it == it_o
    close(a, want)
    assert (a == 0).tolist() == (want == 0).tolist()
    assert abs(mrd - mrd_o) <= 1e-6 * max(abs(mrd_o), 1e-12)


@pytest.mark.parametrize("vb", [0, 1])
@pytest.mark.parametrize("fixed", [1, 2, 49, 50, 51, 400])
def test_gather_loop_fixed_iterations(ctx, vb, fixed, loop_kind):
    T = 8000
    rp, lab, cnt = synth.make_classes(T, 20000, seed=3 + fixed)
    eff = np.random.default_rng(fixed).uniform(50, 4000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb, fixed_iters=fixed))
    assert ctx.last_em_kernel() == loop_kind
    rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, fixed_iters=fixed))
    assert rc == 0 and it == fixed == it_o
    close(a, want)
    assert abs(a.sum() - want.sum()) <= 1e-9 * want.sum()
    assert abs(mrd - mrd_o) <= 1e-6 * max(abs(mrd_o), 1e-12)


def test_gather_loop_iteration_limits(ctx, loop_kind):
    """the loop rule (CollapsedEMOptimizer.cpp:820): itNum < minIter || (itNum < maxIter && !converged)"""
    T = 2000
    rp, lab, cnt = synth.make_classes(T, 5000, seed=12)
    eff = np.full(T, 800.0)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for kw in (dict(max_iter=7, min_iter=3), dict(max_iter=5, min_iter=20), dict(min_iter=0, max_iter=10000), dict(tol=1e-5)):
        a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(**kw))
        assert ctx.last_em_kernel() == loop_kind
        rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(**kw))
        assert rc == 0 and it == it_o, kw
        close(a, want)


def test_gather_loop_equals_scatter_kernels(ctx, monkeypatch, loop_kind):
    T = 30000
    rp, lab, cnt = synth.make_classes(T, 70000, seed=5)
    eff = np.random.default_rng(1).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        a1, it1, m1 = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        assert ctx.last_em_kernel() == loop_kind
        monkeypatch.setenv("SFB200_EM_GATHER", "0"); monkeypatch.setenv("SFB200_EM_DENSE", "0")
        a2, it2, m2 = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        assert ctx.last_em_kernel() == 1                                  # k_em_part
        monkeypatch.setenv("SFB200_EM_GATHER", "1"); monkeypatch.setenv("SFB200_EM_DENSE", "1" if loop_kind == DENSE else "0")
        assert it1 == it2
        close(a1, a2, rtol=1e-7)
        assert abs(m1 - m2) <= 1e-6 * abs(m2)


def test_gather_loop_duplicate_ids_long_classes_and_idle_transcripts(ctx, loop_kind):
    """labels with a repeated transcript id (orphan pairs) and transcripts that belong to no class or only to
    single-member classes; a slot mask has no multiplicity, so k_em_dense hands such class sets to k_em_gather"""
    rng = np.random.default_rng(4)
    T = 4000
    labels = {}
    for g in range(0, T - 10, 10):
        if g % 50 == 40:
            continue                                                       # an idle gene: its transcripts stay at 0
        for _ in range(6):
            n = int(rng.integers(1, 9))
            ids = np.sort(rng.integers(g, g + 10, size=n))                 # with replacement: duplicate ids
            labels[tuple(int(x) for x in ids)] = int(rng.integers(1, 500))
    labs = sorted(labels)
    rp = np.zeros(len(labs) + 1, np.uint64); rp[1:] = np.cumsum([len(l) for l in labs])
    lab = np.array([t for l in labs for t in l], np.uint32)
    cnt = np.array([labels[l] for l in labs], np.uint64)
    eff = rng.uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        a, it, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        assert ctx.last_em_kernel() == GATHER
        rc, want, it_o, _ = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb))
        assert rc == 0 and it == it_o
        close(a, want)
        idle = np.ones(T, bool); idle[lab] = False
        assert (a[idle] == 0).all()


@pytest.mark.parametrize("vb", [0, 1])
def test_gather_loop_bootstrap_counts(ctx, vb, loop_kind):
    """doBootstrap's loop (gate on the OLD alpha, no minimum) on resampled counts, including classes resampled to 0"""
    T = 6000
    rp, lab, cnt = synth.make_classes(T, 15000, seed=31)
    eff = np.random.default_rng(2).uniform(100, 3000, size=T)
    total = int(cnt.sum())
    samp = np.random.default_rng(99).multinomial(total, cnt / cnt.sum()).astype(np.uint64)
    assert (samp == 0).any()
    ctx.eq_import(T, rp, lab, cnt)
    a, it = ctx.bootstrap_em(eff, samp, capi.EMOpts.default(use_vb=vb))
    assert ctx.last_em_kernel() == loop_kind
    rc, want, it_o = O.bootstrap_em(T, rp, lab, samp, eff, O.EMOpts.default(use_vb=vb))
    assert rc == 0 and it == it_o
    close(a, want)


def test_gather_loop_after_device_side_finish(ctx, loop_kind):
    """mapping -> device-side class flatten -> partition -> gather layout -> EM, against the oracle on the same reads"""
    seq, off, ln = synth.make_transcriptome(400, seed=15)
    b1, o1, _, _, _ = synth.make_reads(seq, off, ln, 60000, 76, seed=16)
    fmt = O.parse_libtype("U")
    ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    ctx.map_begin(capi.MapOpts.default(fmt))
    ctx.map_batch(b1, o1, None, None)
    g = ctx.map_finish()
    rp, lab, cnt = ctx.eq_export()
    eff = np.maximum(ln.astype(np.float64) - 75.0, 1.0)
    nm = int(g["counters"][1])
    for vb in (0, 1):
        a, it, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
        assert ctx.last_em_kernel() == loop_kind
        rc, want, it_o, _ = O.em_run(len(ln), rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb))
        assert rc == 0 and it == it_o
        close(a, want)
    rows = ctx.bootstrap_run(eff, 3, seed=5)
    assert ctx.last_em_kernel() == loop_kind
    np.testing.assert_allclose(rows.sum(axis=1), int(cnt.sum()), rtol=1e-9)


def test_dense_loop_component_sizes(ctx, loop_kind):
    """components of exactly 8 transcripts run on one thread; 9 do not (the gather loop takes over); genes of 2 and 3 too"""
    for gene, want in ((8, loop_kind), (9, GATHER), (2, loop_kind), (3, loop_kind)):
        T = 40 * gene * 10
        n_cls = min(25 * T // gene, (T // gene) * (2 ** gene - 1) // 2)          # at most half of the subsets of every gene
        rp, lab, cnt = synth.make_classes(T, n_cls, seed=gene, gene_size=gene, max_len=gene)
        # one class spanning the whole gene, so that the component really has `gene` transcripts
        labs = {tuple(int(x) for x in lab[int(rp[i]):int(rp[i + 1])]): int(cnt[i]) for i in range(len(cnt))}
        for g0 in range(0, T, gene):
            labs[tuple(range(g0, g0 + gene))] = 7
        keys = sorted(labs)
        rp = np.zeros(len(keys) + 1, np.uint64); rp[1:] = np.cumsum([len(k) for k in keys])
        lab = np.array([t for k in keys for t in k], np.uint32); cnt = np.array([labs[k] for k in keys], np.uint64)
        eff = np.random.default_rng(gene).uniform(100, 3000, size=T)
        nm = int(cnt.sum())
        ctx.eq_import(T, rp, lab, cnt)
        for vb in (0, 1):
            a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
            assert ctx.last_em_kernel() == want, (gene, ctx.last_em_kernel())
            rc, ref, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb), n_threads=4)
            assert rc == 0 and it == it_o
            close(a, ref)
            assert abs(mrd - mrd_o) <= 1e-6 * max(abs(mrd_o), 1e-12)


def paralog_classes(T, seed, n_fam=6, fam_genes=(8, 60), gene=5):
    """gene-local classes plus families whose members are scattered over the transcript order: classes that cross genes (and CTA
    ranges), connected components of hundreds of transcripts, labels of up to ~100 members"""
    rng = np.random.default_rng(seed)
    labs = {}
    n_genes = T // gene
    for g in range(n_genes):
        for _ in range(4):
            n = int(rng.integers(1, gene + 1))
            ids = tuple(sorted(int(x) for x in rng.choice(np.arange(g * gene, (g + 1) * gene), size=n, replace=False)))
            labs[ids] = int(max(1, rng.lognormal(2.0, 2.0)))
    for f in range(n_fam):
        genes = rng.choice(n_genes, size=int(rng.integers(*fam_genes)), replace=False)
        members = np.concatenate([np.arange(g * gene, (g + 1) * gene) for g in genes])
        for _ in range(12 * len(genes)):
            n = int(rng.integers(2, min(100, len(members)) + 1)) if rng.random() < 0.2 else int(rng.integers(2, 7))
            ids = tuple(sorted(int(x) for x in rng.choice(members, size=n, replace=False)))
            labs[ids] = int(max(1, rng.lognormal(2.0, 2.0)))
    keys = sorted(labs)
    rp = np.zeros(len(keys) + 1, np.uint64); rp[1:] = np.cumsum([len(k) for k in keys])
    return rp, np.array([t for k in keys for t in k], np.uint32), np.array([labs[k] for k in keys], np.uint64)


@pytest.mark.parametrize("T,seed", [(6000, 1), (40000, 2)])
def test_hybrid_components_and_pool_loop(ctx, monkeypatch, loop_kind, T, seed):
    """class sets with paralog families: the small components run on component threads, everything else -- large components,
    classes that cross CTA ranges -- in the pool loop on CTAs of its own (sfb200_last_em_kernel == 5); EM and VBEM, converging and
    fixed-iteration runs, and a bootstrap-style run against the oracle = the reference's optimizer"""
    if loop_kind != DENSE:
        pytest.skip("the pool loop belongs to the dense kernel")
    monkeypatch.setenv("SFB200_EM_HYBRID", "1")                         # opt-in: measured slower than k_em_part so far
    rp, lab, cnt = paralog_classes(T, seed)
    eff = np.random.default_rng(seed).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    for vb in (0, 1):
        for kw in ({}, {"fixed_iters": 37}, {"min_iter": 5, "max_iter": 60}):
            a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb, **kw))
            assert ctx.last_em_kernel() == 5, ctx.last_em_kernel()
            rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, **kw), n_threads=4)
            assert rc == 0 and it == it_o, (vb, kw, it, it_o)
            close(a, want)
            assert (a == 0).tolist() == (want == 0).tolist()
            if not kw.get("fixed_iters"):
                assert abs(mrd - mrd_o) <= 1e-5 * max(abs(mrd_o), 1e-12)
    # the same classes with the hybrid switched off: the partitioned scatter loop; both agree
    a1, it1, _ = ctx.em_run(eff, nm, capi.EMOpts.default(fixed_iters=50))
    monkeypatch.setenv("SFB200_EM_HYBRID", "0")
    ctx.eq_import(T, rp, lab, cnt)
    a2, it2, _ = ctx.em_run(eff, nm, capi.EMOpts.default(fixed_iters=50))
    assert ctx.last_em_kernel() in (0, 1)
    close(a1, a2, rtol=1e-7)
    monkeypatch.setenv("SFB200_EM_HYBRID", "1")
    ctx.eq_import(T, rp, lab, cnt)
    # resampled counts (bootstrap): doBootstrap's loop rule on the same layout
    samp = np.random.default_rng(9).multinomial(nm, cnt / cnt.sum()).astype(np.uint64)
    a, it = ctx.bootstrap_em(eff, samp, capi.EMOpts.default())
    rc, want, it_o = O.bootstrap_em(T, rp, lab, samp, eff, O.EMOpts.default())
    assert rc == 0 and it == it_o
    close(a, want)


@pytest.mark.parametrize("vb", [0, 1])
def test_dense_loop_streaming_variant(ctx, monkeypatch, loop_kind, vb):
    """k_em_dense with the class counts, base and 1/effLen streamed from global memory (what runs when a CTA's slice does not fit in
    shared memory: 1 M transcripts) reproduces the oracle and the shared-memory variant"""
    if loop_kind != DENSE or os.environ.get("SFB200_EM_DENSE_GROUP") == "0":
        pytest.skip("the streaming variant belongs to the dense kernel with two lanes per component")
    T = 30000
    rp, lab, cnt = synth.make_classes(T, 70000, seed=15)
    eff = np.random.default_rng(2).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    ctx.eq_import(T, rp, lab, cnt)
    a0, it0, _ = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb))
    assert ctx.last_em_kernel() == DENSE
    monkeypatch.setenv("SFB200_EM_FORCE_STREAM", "1")
    ctx.eq_import(T, rp, lab, cnt)                                        # the layout is rebuilt for the new setting
    for kw in ({}, {"fixed_iters": 23}):
        a, it, mrd = ctx.em_run(eff, nm, capi.EMOpts.default(use_vb=vb, **kw))
        assert ctx.last_em_kernel() == DENSE
        rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, **kw), n_threads=4)
        assert rc == 0 and it == it_o
        close(a, want)
        if not kw:
            close(a, a0, rtol=1e-9)


@pytest.mark.parametrize("vb", [0, 1])
def test_dense_loop_lagged_stopping_rule(ctx, monkeypatch, loop_kind, vb):
    """k_em_dense reads the global quantities of an iteration (max relative change, VBEM's alpha sum) three iterations late instead of
    behind a grid barrier per iteration, and returns the alphas of the iteration that met the rule from a shared-memory ring
    (em_dense.cuh): same iteration count and estimates as the oracle and as the synchronous loop (SFB200_EM_NO_LAG=1) -- to
    convergence, at the iteration cap (cap below / at / just above the lag), with minIter above the converging iteration, with a
    loose tolerance that is met inside the first lag window, and with a fixed count (VBEM).  (The bootstrap's form of the rule -- gate on
    the old alphas, no minIter -- runs through the same code in the bootstrap tests.)"""
    if loop_kind != DENSE or os.environ.get("SFB200_EM_DENSE_GROUP", "2") != "2":
        pytest.skip("the lagged rule belongs to the dense kernel with two lanes per component")
    T = 30000
    rp, lab, cnt = synth.make_classes(T, 70000, seed=15)
    eff = np.random.default_rng(2).uniform(100, 3000, size=T)
    nm = int(cnt.sum())
    cases = [{}, {"max_iter": 1, "min_iter": 1}, {"max_iter": 2, "min_iter": 1}, {"max_iter": 3, "min_iter": 2}, {"max_iter": 4, "min_iter": 1},
             {"max_iter": 5, "min_iter": 1}, {"max_iter": 37}, {"max_iter": 60}, {"min_iter": 1}, {"min_iter": 400},
             {"min_iter": 1, "tol": 0.5}]                               # stops inside the first lag window
    if vb:
        cases += [{"fixed_iters": 1}, {"fixed_iters": 2}, {"fixed_iters": 23}]
    for kw in cases:
        o = capi.EMOpts.default(use_vb=vb, **kw)
        monkeypatch.delenv("SFB200_EM_NO_LAG", raising=False)
        ctx.eq_import(T, rp, lab, cnt)
        a, it, mrd = ctx.em_run(eff, nm, o)
        assert ctx.last_em_kernel() == DENSE
        monkeypatch.setenv("SFB200_EM_NO_LAG", "1")
        a_s, it_s, mrd_s = ctx.em_run(eff, nm, o)
        rc, want, it_o, mrd_o = O.em_run(T, rp, lab, cnt, eff, nm, O.EMOpts.default(use_vb=vb, **kw), n_threads=4)
        assert rc == 0 and it == it_o == it_s, (kw, it, it_s, it_o)
        close(a, want)
        if vb:
            close(a, a_s, rtol=1e-9)
        else:
            assert (a == a_s).all() and mrd == mrd_s                     # EM: the very same arithmetic, CTA by CTA
