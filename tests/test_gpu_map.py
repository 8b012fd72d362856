"""GPU parity tests (-m gpu) for index construction and the read -> equivalence-class kernel, through the C ABI.

Integer results (index arrays, class labels and counts, the six counters, the fragment-length histogram) must be
bit-exact against the CPU oracle on identical inputs (BASELINE.json north_star).
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from sailfish_b200 import capi, synth
from conftest import split_seqs

pytestmark = pytest.mark.gpu


def small_txome(n_genes=60, seed=42):
    seq, off, ln = synth.make_transcriptome(n_genes, seed=seed)
    return seq, off, ln


def run_both(ctx, seq, off, ln, b1, o1, b2, o2, libtype, k=31, batches=1, **kw):
    st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=k)
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    oix = O.Index(seqs, k=k)
    fmt = O.parse_libtype(libtype)
    ctx.map_begin(capi.MapOpts.default(fmt, **kw))
    run = O.Run(oix, O.MapOpts.default(fmt, **kw))
    n = len(o1) - 1
    cuts = np.linspace(0, n, batches + 1).astype(int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b2 is None:
            ctx.map_batch(b1, o1[a:b + 1])
            run.map_batch(b1.tobytes(), o1[a:b + 1] if a == 0 else o1[a:b + 1])
        else:
            ctx.map_batch(b1, o1[a:b + 1], b2, o2[a:b + 1])
            run.map_batch(b1.tobytes(), o1[a:b + 1], b2.tobytes(), o2[a:b + 1])
    g = ctx.map_finish()
    w = run.finish()
    return st, oix, g, w


def assert_same_classes(ctx, g, w):
    assert g["counters"].tolist() == w["counters"].tolist()
    assert g["fld"].tolist() == w["fld"].tolist()
    rp, lab, cnt = ctx.eq_export()
    assert g["n_classes"] == len(w["counts"])
    assert rp.tolist() == w["row_ptr"].tolist()
    assert lab.tolist() == w["labels"].tolist()
    assert cnt.tolist() == w["counts"].tolist()
    assert int(cnt.sum()) == int(g["counters"][1])           # sum of class counts == numMappedFragments


def test_index_matches_oracle(ctx):
    seq, off, ln = small_txome(40)
    # sprinkle non-ACGT characters and lower case into the transcript text
    seq = seq.copy()
    rng = np.random.default_rng(0)
    seq[rng.integers(0, len(seq), 200)] = ord("N")
    low = rng.integers(0, len(seq), 500)
    seq[low] = np.char.lower(seq[low].view("S1")).view(np.uint8)
    for k in (31, 21, 15):
        st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=k)
        seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
        oix = O.Index(seqs, k=k)
        e = oix.export()
        words, sa_pos, sa_tid = ctx.index_export()
        ow, n = oix.text_words()
        assert st["text_len"] == n and st["n_sa"] == len(e["sa_pos"]) and st["n_kmers"] == len(e["kmers"])
        assert words.tolist() == ow.tolist()
        assert sa_pos.tolist() == e["sa_pos"].tolist()
        assert sa_tid.tolist() == e["sa_tid"].tolist()
        assert st["max_bucket"] == int(e["cnt"].max())


def test_index_short_transcripts_and_bad_k(ctx):
    seqs = [b"ACGTACGTAC", b"A" * 40 + b"C" * 40, b"ACG", b"GATTACA" * 20]
    st = ctx.index_build(seqs=seqs, k=31)
    oix = O.Index(seqs, k=31)
    assert st["n_sa"] == O.lib().orc_index_n_sa(oix.h)
    with pytest.raises(capi.Sfb200Error):
        ctx.index_build(seqs=seqs, k=30)         # even k refused (SailfishIndexer.cpp:199-205)


@pytest.mark.parametrize("libtype", ["U", "SF", "SR"])
def test_single_end_matches_oracle(ctx, libtype):
    seq, off, ln = small_txome()
    b1, o1, _, _, _ = synth.make_reads(seq, off, ln, 30000, 76, seed=11, sub_rate=0.01, n_rate=0.002)
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, None, None, libtype, batches=3)
    assert_same_classes(ctx, g, w)
    assert g["counters"][1] > 25000


@pytest.mark.parametrize("libtype,kw", [("IU", {}), ("ISF", {}), ("ISR", {"enforce_compat": 1}), ("OU", {}), ("MU", {}),
                                        ("IU", {"allow_orphans": 0}), ("IU", {"strict_intersect": 1}),
                                        ("ISF", {"ignore_compat": 1}), ("IU", {"allow_dovetail": 1}),
                                        ("IU", {"max_read_occs": 3}), ("IU", {"num_frag_samples": 500}),
                                        ("IU", {"max_interval": 2})])
def test_paired_end_matches_oracle(ctx, libtype, kw):
    seq, off, ln = small_txome()
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 20000, 100, seed=12, paired=True, sub_rate=0.01, n_rate=0.002)
    # break some mates so that orphans occur: overwrite mate 2 of every 7th pair with random sequence
    b2 = b2.copy().reshape(-1, 100)
    rng = np.random.default_rng(1)
    b2[::7] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=b2[::7].shape)]
    b2 = b2.reshape(-1)
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, b2, o2, libtype, batches=2, **kw)
    assert_same_classes(ctx, g, w)


def test_sample_data_matches_golden(ctx, sample_data):
    """BASELINE config 1: the bundled sample (15 transcripts, 10 000 x 2 x 50 nt, -l IU) against the committed fixture."""
    d = sample_data
    seqs = split_seqs(d["txp_seq"], d["txp_len"])
    ctx.index_build(seqs=seqs, k=31)
    ctx.map_begin(capi.MapOpts.default(O.parse_libtype("IU")))
    ctx.map_batch(d["reads1"], d["off1"], d["reads2"], d["off2"])
    g = ctx.map_finish()
    assert g["counters"].tolist() == d["counters"].tolist()
    assert g["fld"].tolist() == d["fld"].tolist()
    rp, lab, cnt = ctx.eq_export()
    assert rp.tolist() == d["row_ptr"].tolist() and lab.tolist() == d["labels"].tolist() and cnt.tolist() == d["counts"].tolist()
    # straight into inference on the classes that are already on the device
    alphas, iters, _ = ctx.em_run(d["eff"], int(d["num_mapped"]))
    np.testing.assert_allclose(alphas, d["ref_est_vb0"], rtol=1e-4, atol=1e-6)


def test_ragged_and_degenerate_reads(ctx):
    seq, off, ln = small_txome(20)
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    t0 = seqs[0]
    reads = [t0[10:86], b"", b"ACGT", b"N" * 76, t0[0:300], t0[5:36], b"A" * 80, t0[100:130] + b"N" + t0[131:200],
             t0[50:81].lower(), bytes(reversed(t0[20:96])), t0[-76:], t0[:31]]
    b1, o1 = capi.pack_reads(reads)
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, None, None, "U")
    assert_same_classes(ctx, g, w)
    assert g["counters"][0] == len(reads)


def test_small_k_and_many_hits(ctx):
    """k = 15 on a repetitive transcriptome: big buckets, reads with more hits than max_read_occs"""
    rng = np.random.default_rng(3)
    unit = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=300)].tobytes()
    seqs = [unit + np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=100)].tobytes() for _ in range(300)]
    reads = [unit[i:i + 60] for i in range(0, 200, 7)] + [s[290:350] for s in seqs[:50]]
    b1, o1 = capi.pack_reads(reads)
    seq = np.frombuffer(b"".join(seqs), np.uint8)
    ln = np.array([len(s) for s in seqs], np.uint32)
    off = np.zeros(len(seqs), np.uint64); off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
    for kw in ({}, {"max_read_occs": 400}, {"max_interval": 100}):
        st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, None, None, "U", k=15, **kw)
        assert_same_classes(ctx, g, w)


@pytest.mark.parametrize("paired", [False, True])
def test_repeats_and_families_match_oracle(ctx, paired, monkeypatch):
    """Paralog families and repeat elements (synth.make_transcriptome): seed buckets of hundreds of positions next to ordinary reads.
    The finalize kernel sets the reads with a bucket of more than 32 positions aside and runs them in a second pass, with the scan's
    per-round extension words instead of its 32-bit mask; the classes must not change -- with the pool (default), with a pool too
    small for most buckets (SFB200_IVPOOL_WORDS: the finalize kernel extends the entries itself) and in one pass (SFB200_NO_HEAVY_PASS)."""
    seq, off, ln = synth.make_transcriptome(1500, seed=5, family_frac=0.2, family_genes=(10, 120), repeat_frac=0.5, repeat_len=200)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 40000, 76, seed=9, paired=paired, sub_rate=0.01, n_rate=0.002)
    lib = "IU" if paired else "U"
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, b2, o2, lib, batches=2)
    assert st["max_bucket"] > 200                              # the case is what it says
    assert_same_classes(ctx, g, w)
    for env in ({"SFB200_IVPOOL_WORDS": "64"}, {"SFB200_NO_HEAVY_PASS": "1"}):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        ctx.map_begin(capi.MapOpts.default(O.parse_libtype(lib)))
        if paired:
            ctx.map_batch(b1, o1, b2, o2)
        else:
            ctx.map_batch(b1, o1)
        g2 = ctx.map_finish()
        assert_same_classes(ctx, g2, w)
        for k_ in env:
            monkeypatch.delenv(k_)


@pytest.mark.parametrize("paired", [False, True])
def test_host_batches_in_small_pieces(ctx, paired, monkeypatch):
    """sfb200_map_batch copies a host batch piece by piece through three staging sets (copy of piece j+1 | enqueue of j | kernels of
    j-1).  With 1024-read pieces a 30 000-read batch rotates through the sets ten times; ragged read lengths make every piece's byte
    range different."""
    seq, off, ln = small_txome()
    reads1, reads2 = [], []
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 30000, 100, seed=21, paired=paired, sub_rate=0.01, n_rate=0.002)
    rng = np.random.default_rng(4)
    cut = rng.integers(40, 101, size=30000)
    for i in range(30000):
        reads1.append(b1[int(o1[i]):int(o1[i]) + int(cut[i])].tobytes())
        if paired:
            reads2.append(b2[int(o2[i]):int(o2[i]) + int(cut[(i * 7) % 30000])].tobytes())
    b1, o1 = capi.pack_reads(reads1)
    if paired:
        b2, o2 = capi.pack_reads(reads2)
    lib = "IU" if paired else "U"
    monkeypatch.setenv("SFB200_HOST_PIECE", "1024")
    for ramp in ("0", "128"):
        monkeypatch.setenv("SFB200_MAP_RAMP", ramp)
        st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, b2 if paired else None, o2 if paired else None, lib, batches=2)
        assert_same_classes(ctx, g, w)


@pytest.mark.parametrize("paired", [False, True])
def test_fixed_length_host_batches(ctx, paired, monkeypatch):
    """sfb200_map_batch_fixed: reads of one length per mate stored back to back, no offsets array (they are written on the device);
    mates of different lengths, several calls, small pieces"""
    seq, off, ln = small_txome()
    n, L1, L2 = 20000, 100, 83
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, L1, seed=23, paired=paired, sub_rate=0.01, n_rate=0.002)
    if paired:
        b2 = np.ascontiguousarray(b2[:n * L1].reshape(n, L1)[:, :L2]).reshape(-1)
        o2 = np.arange(n + 1, dtype=np.uint64) * L2
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    fmt = O.parse_libtype("IU" if paired else "U")
    run = O.Run(O.Index(seqs, k=31), O.MapOpts.default(fmt))
    if paired:
        run.map_batch(b1.tobytes(), o1, b2.tobytes(), o2)
    else:
        run.map_batch(b1.tobytes(), o1)
    w = run.finish()
    monkeypatch.setenv("SFB200_HOST_PIECE", "2048")
    ctx.map_begin(capi.MapOpts.default(fmt))
    for a, b in ((0, 7000), (7000, 7001), (7001, n)):
        if paired:
            ctx.map_batch_fixed(b1[a * L1:b * L1], L1, b2[a * L2:b * L2], L2)
        else:
            ctx.map_batch_fixed(b1[a * L1:b * L1], L1)
    g = ctx.map_finish()
    assert_same_classes(ctx, g, w)
    assert ctx.map_h2d_bytes() == n * L1 + (n * L2 if paired else 0)      # bases only


def test_full_size_properties(ctx):
    """Larger run (2 000 genes = 10 000 transcripts, 400k reads): invariants that do not need the oracle at size, plus
    the oracle on the same input with 8 host threads."""
    seq, off, ln = synth.make_transcriptome(2000, seed=42)
    b1, o1, _, _, tid = synth.make_reads(seq, off, ln, 400000, 76, seed=1234)
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, None, None, "U", batches=4)
    assert_same_classes(ctx, g, w)
    c = g["counters"]
    assert c[0] == 400000 and c[1] <= c[3] <= c[0] and c[4] + c[5] == c[2]
    # mapping twice gives the same table (idempotence of begin/finish)
    ctx.map_begin(capi.MapOpts.default(O.parse_libtype("U")))
    ctx.map_batch(b1, o1)
    g2 = ctx.map_finish()
    assert g2["counters"].tolist() == c.tolist() and g2["n_classes"] == g["n_classes"]


def test_oracle_on_device_built_index(ctx):
    """bench.py's CPU arm runs the oracle on the index the GPU built (arrays + k-mer table): same answers as the oracle's
    own index, so the CPU baseline at full size measures the same algorithm on the same data."""
    seq, off, ln = small_txome(50, seed=7)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 5000, 100, seed=3, paired=True, sub_rate=0.01)
    st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    words, sa_pos, sa_tid = ctx.index_export()
    table = ctx.index_export_table()
    assert int((table[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)).sum()) == st["n_kmers"]
    o_dev = O.Index.from_table(words, st["text_len"], ln, 31, sa_pos, sa_tid, table)
    seqs = [seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes() for i in range(len(ln))]
    o_own = O.Index(seqs, k=31)
    fmt = O.parse_libtype("IU")
    res = []
    for ix in (o_dev, o_own):
        run = O.Run(ix, O.MapOpts.default(fmt))
        run.map_batch(b1.tobytes(), o1, b2.tobytes(), o2, n_threads=4)
        res.append(run.finish())
    for key in ("counters", "fld", "row_ptr", "labels", "counts"):
        assert res[0][key].tolist() == res[1][key].tolist()


def test_large_batches_are_chunked(ctx, monkeypatch):
    """a batch larger than the per-launch hand-over buffers is walked in chunks: same classes, counters, FLD sample"""
    seq, off, ln = small_txome(30, seed=5)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 20000, 100, seed=21, paired=True, sub_rate=0.01)
    monkeypatch.setenv("SFB200_MAX_CHUNK", "3000")
    st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, b2, o2, "IU", batches=2)
    assert_same_classes(ctx, g, w)


# ---- bias / GC sample collection while mapping (sfb200_map_set_bias / sfb200_map_get_bias) ----------------------------------------
# The per-hit arithmetic is also checked on CPU (tests/bias_core_test.cpp).  First B200 run: profiles/r02a_experimental_gpu.txt.
import os



@pytest.mark.parametrize("paired,libtype,seq_bias,gc_bias,n_samples,kw", [
    (True, "IU", 1, 1, 1000000, {}), (True, "IU", 1, 0, 3000, {}), (True, "ISF", 0, 1, 0, {}), (True, "IU", 1, 1, 1000000, {"allow_orphans": 0}),
    (True, "OU", 1, 1, 1000000, {"max_read_occs": 3}), (False, "U", 1, 0, 1000000, {}), (False, "SR", 1, 1, 2500, {}),
])
def test_bias_samples_match_oracle(ctx, paired, libtype, seq_bias, gc_bias, n_samples, kw):
    seq, off, ln = small_txome()
    n, L = 20000, 100 if paired else 76
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, L, seed=21, paired=paired, sub_rate=0.01, n_rate=0.002)
    if paired:
        b2 = b2.copy().reshape(-1, L)
        rng = np.random.default_rng(2)
        b2[::7] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=b2[::7].shape)]       # orphans
        b2 = b2.reshape(-1)
    else:
        b2 = o2 = None
    ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    oix = O.Index(split_seqs(seq, ln), k=31)
    fmt = O.parse_libtype(libtype)
    ctx.map_begin(capi.MapOpts.default(fmt, **kw))
    ctx.map_set_bias(seq_bias, gc_bias, n_samples)
    run = O.Run(oix, O.MapOpts.default(fmt, **kw))
    run.set_bias(seq_bias, gc_bias, n_samples)
    cuts = np.linspace(0, n, 4).astype(int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        if paired:
            ctx.map_batch(b1, o1[a:b + 1], b2, o2[a:b + 1]); run.map_batch(b1.tobytes(), o1[a:b + 1], b2.tobytes(), o2[a:b + 1])
        else:
            ctx.map_batch(b1, o1[a:b + 1]); run.map_batch(b1.tobytes(), o1[a:b + 1])
    g = ctx.map_finish(); w = run.finish()
    assert_same_classes(ctx, g, w)                                   # the classes are what they are without the collection
    rb, og = ctx.map_get_bias()
    rb_o, og_o = run.finish_bias()
    assert rb.tolist() == rb_o.tolist()
    assert og.tolist() == og_o.tolist()
    if seq_bias:
        assert int(rb.sum()) - 4096 == min(n_samples, int(rb.sum()) - 4096) and (n_samples < 5000 or int(rb.sum()) - 4096 > 10000)
    else:
        assert int(rb.sum()) == 4096
    if gc_bias and paired:
        assert int(og.sum()) - 101 > 5000
    else:
        assert int(og.sum()) == 101


@pytest.mark.parametrize("paired,eol,block,fasta", [(True, "\n", 0, False), (True, "\r\n", 300000, False), (False, "\n", 70000, False), (True, "\n", 9000, False),
                                                    (True, "\n", 50000, True)])
def test_map_fastq_equals_map_batch(ctx, paired, eol, block, fasta):
    """sfb200_map_fastq (FASTQ text parsed on the device, fastq.cu) gives the classes, counters and FLD of sfb200_map_batch on the same
    reads -- whole text at once and block-wise with the unconsumed tail carried over (mate 2 has longer headers, so its blocks hold
    fewer records)"""
    seq, off, ln = small_txome()
    n, L = 12000, 100 if paired else 76
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, n, L, seed=31, paired=paired, sub_rate=0.01, n_rate=0.002)

    def text(bases, offs, long_names):
        recs = []
        for i in range(n):
            s_ = bases[int(offs[i]):int(offs[i + 1])].tobytes().decode()
            if fasta:
                recs.append(">r%d%s%s%s%s" % (i, " description of the read" * 2 if long_names else "", eol, s_, eol))
            else:
                recs.append("@r%d%s%s%s%s+%s%s%s" % (i, " description of the read" * 2 if long_names else "", eol, s_, eol, eol, "@" * len(s_), eol))
        return "".join(recs).encode()

    t1 = text(b1, o1, False)
    t2 = text(b2, o2, True) if paired else None
    ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    oix = O.Index(split_seqs(seq, ln), k=31)
    fmt = O.parse_libtype("IU" if paired else "U")
    run = O.Run(oix, O.MapOpts.default(fmt))
    if paired:
        run.map_batch(b1.tobytes(), o1, b2.tobytes(), o2, n_threads=4)
    else:
        run.map_batch(b1.tobytes(), o1, n_threads=4)
    w = run.finish()
    ctx.map_begin(capi.MapOpts.default(fmt))
    if block == 0:
        got, c1, c2 = ctx.map_fastq(t1, t2)
        assert got == n and c1 == len(t1) and (not paired or c2 == len(t2))
    else:
        p1 = p2 = 0
        buf1 = buf2 = b""
        total = calls = 0
        while True:
            take1 = max(0, block - len(buf1)); buf1 += t1[p1:p1 + take1]; p1 += take1
            if paired:
                take2 = max(0, block - len(buf2)); buf2 += t2[p2:p2 + take2]; p2 += take2
            if not buf1 and not buf2:
                break
            got, c1, c2 = ctx.map_fastq(buf1, buf2 if paired else None)
            assert got > 0
            total += got; calls += 1
            buf1 = buf1[c1:]
            if paired:
                buf2 = buf2[c2:]
        assert total == n and calls > 3
    g = ctx.map_finish()
    assert_same_classes(ctx, g, w)
    # malformed text is refused
    ctx.map_begin(capi.MapOpts.default(fmt))
    with pytest.raises(capi.Sfb200Error):
        ctx.map_fastq(b">a\nACGT\nAC\n>b\nGGCC\nGG\n", b">a\nACGT\nAC\n>b\nGGCC\nGG\n" if paired else None)     # wrapped FASTA
    ctx.map_finish()


@pytest.mark.parametrize("paired", [False, True])
def test_class_table_and_arena_grow(ctx, monkeypatch, paired):
    """a class table of 16 buckets and a label arena of 1024 words: both fill up many times over; the reads whose upsert failed are
    replayed after every doubling (libcuckoo grows too, include/cuckoohash_map.hh) -- classes, counters and FLD equal the oracle's"""
    monkeypatch.setenv("SFB200_EQ_LOG2_BUCKETS", "4")
    monkeypatch.setenv("SFB200_EQ_ARENA_LOG2", "10")
    monkeypatch.setenv("SFB200_MAX_CHUNK", "7000")
    seq, off, ln = synth.make_transcriptome(600, seed=11)
    if paired:
        b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 60000, 100, seed=5, paired=True, sub_rate=0.01)
        st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, b2, o2, "IU", batches=3)
    else:
        b1, o1, _, _, _ = synth.make_reads(seq, off, ln, 60000, 76, seed=5)
        st, oix, g, w = run_both(ctx, seq, off, ln, b1, o1, None, None, "U", batches=3)
    assert g["n_classes"] > 1500                                         # far more classes than the 64 + 1024 slots it started with
    assert_same_classes(ctx, g, w)


def test_index_file_round_trip(ctx, tmp_path):
    """sfb200_index_save / sfb200_index_load (what SailfishIndex::load reads from the index directory, SailfishIndex.hpp:80-144): a
    fresh context that loads the file maps reads to the same classes as the context that built the index; damaged files are refused"""
    seq, off, ln = small_txome(40, seed=9)
    b1, o1, b2, o2, _ = synth.make_reads(seq, off, ln, 8000, 100, seed=4, paired=True, sub_rate=0.01)
    st = ctx.index_build(seq=seq, txp_off=off, txp_len=ln, k=31)
    path = str(tmp_path / "index.bin")
    ctx.index_save(path)
    fmt = O.parse_libtype("IU")
    ctx.map_begin(capi.MapOpts.default(fmt)); ctx.map_batch(b1, o1, b2, o2); g = ctx.map_finish()
    want = (g["counters"].tolist(), g["fld"].tolist(), [a.tolist() for a in ctx.eq_export()])
    other = capi.Context(0)
    try:
        st2 = other.index_load(path)
        assert {k: st2[k] for k in ("text_len", "n_sa", "n_kmers", "max_bucket")} == {k: st[k] for k in ("text_len", "n_sa", "n_kmers", "max_bucket")}
        w1, p1, t1 = ctx.index_export(); w2, p2, t2 = other.index_export()
        assert w1.tolist() == w2.tolist() and p1.tolist() == p2.tolist() and t1.tolist() == t2.tolist()
        other.map_begin(capi.MapOpts.default(fmt)); other.map_batch(b1, o1, b2, o2); g2 = other.map_finish()
        assert (g2["counters"].tolist(), g2["fld"].tolist(), [a.tolist() for a in other.eq_export()]) == want
        blob = open(path, "rb").read()
        open(path, "wb").write(blob[:len(blob) // 2])
        with pytest.raises(capi.Sfb200Error):
            other.index_load(path)                                    # truncated
        open(path, "wb").write(b"XXXXXXXX" + blob[8:])
        with pytest.raises(capi.Sfb200Error):
            other.index_load(path)                                    # not an index file
    finally:
        other.close()
