"""Deterministic synthetic transcriptomes and reads for the BASELINE.json configurations.

A transcriptome is `n_genes` genes x `n_iso` isoforms; each gene has `n_exons` exons with log-normal lengths, and an
isoform is an ordered random subset of its gene's exons, so isoforms of a gene share sequence and reads multi-map within a
gene (that is what makes equivalence classes with more than one transcript).  Reads are drawn from transcripts with
log-normal expression (a fraction of transcripts unexpressed), uniform start, random strand and a substitution rate.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def make_transcriptome(n_genes, n_iso=5, n_exons=12, seed=42, gc=0.45, mu=5.3, sigma=0.6, p_keep=0.6,
                       family_frac=0.0, family_genes=(10, 400), family_share=0.5, repeat_frac=0.0, repeat_len=300):
    """-> (seq uint8[total], txp_off uint64[T], txp_len uint32[T])

    family_frac > 0 gives the class structure of a real annotation (paralog families, repeats) instead of genes that share nothing:
    that fraction of the genes is grouped into families of family_genes[0]..family_genes[1] genes (log-uniform sizes, members
    scattered over the gene order as paralogs are over a genome); every member takes each exon slot from the family's founder
    with probability family_share, so reads from those exons map to 50..2000 transcripts of many genes, equivalence classes cross
    gene boundaries and the (class x transcript) graph gets large connected components.  repeat_frac > 0 additionally replaces one
    exon of that fraction of ALL genes by one of 8 repeat elements of repeat_len bases (k-mer buckets of thousands of positions)."""
    rng = np.random.default_rng(seed)
    ex_len = np.clip(rng.lognormal(mu, sigma, size=(n_genes, n_exons)), 60, 2000).astype(np.int64)
    ex_src = np.arange(n_genes * n_exons, dtype=np.int64)        # exon slot -> the slot whose sequence it carries
    if family_frac > 0 or repeat_frac > 0:
        frng = np.random.default_rng([seed, 77])
        if family_frac > 0:
            pool = frng.permutation(n_genes)[:int(n_genes * family_frac)]
            at = 0
            while at < len(pool):
                lo, hi = family_genes
                size = int(np.exp(frng.uniform(np.log(lo), np.log(hi))))
                mem = pool[at:at + size]
                at += size
                if len(mem) < 2:
                    break
                founder = mem[0]
                share = frng.random((len(mem) - 1, n_exons)) < family_share
                for e in range(n_exons):
                    g = mem[1:][share[:, e]]
                    ex_src[g * n_exons + e] = founder * n_exons + e
                    ex_len[g, e] = ex_len[founder, e]
        if repeat_frac > 0:
            hosts = frng.permutation(n_genes)[:int(n_genes * repeat_frac)]
            donors = hosts[:8]
            ex_len[donors, 0] = repeat_len
            slot = frng.integers(0, n_exons, size=len(hosts))
            which = donors[frng.integers(0, len(donors), size=len(hosts))]
            ex_src[hosts[8:] * n_exons + slot[8:]] = which[8:] * n_exons
            ex_len[hosts[8:], slot[8:]] = repeat_len
    ex_off = np.zeros(n_genes * n_exons + 1, np.int64)
    ex_off[1:] = np.cumsum(ex_len.ravel())
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    exon_seq = _ACGT[rng.choice(4, size=int(ex_off[-1]), p=p)]
    keep = rng.random((n_genes, n_iso, n_exons)) < p_keep
    # at least two exons per isoform
    few = keep.sum(axis=2) < 2
    keep[few, 0] = True
    keep[few, 1] = True
    T = n_genes * n_iso
    lens = (keep * ex_len[:, None, :]).sum(axis=2).reshape(T)
    txp_len = lens.astype(np.uint32)
    txp_off = np.zeros(T, np.uint64)
    txp_off[1:] = np.cumsum(lens)[:-1]
    seq = np.empty(int(lens.sum()), np.uint8)
    # copy exon by exon, vectorised over (gene, isoform) pairs per exon slot
    cur = txp_off.astype(np.int64).copy()
    keep2 = keep.reshape(T, n_exons)
    gene_of = np.repeat(np.arange(n_genes), n_iso)
    for e in range(n_exons):
        sel = np.nonzero(keep2[:, e])[0]
        if sel.size == 0:
            continue
        el = ex_len[gene_of[sel], e]
        src0 = ex_off[ex_src[gene_of[sel] * n_exons + e]]
        tot = int(el.sum())
        seg_start = np.zeros(sel.size, np.int64)
        seg_start[1:] = np.cumsum(el)[:-1]
        within = np.arange(tot, dtype=np.int64) - np.repeat(seg_start, el)
        seq[np.repeat(cur[sel], el) + within] = exon_seq[np.repeat(src0, el) + within]
        cur[sel] += el
    return seq, txp_off, txp_len


def make_reads(seq, txp_off, txp_len, n_reads, read_len, seed=1234, paired=False, frag_mean=200.0, frag_sd=25.0,
               sub_rate=0.005, zero_frac=0.3, n_rate=0.0, expr_seed=None, stream=0):
    """-> (bases1, off1, bases2|None, off2|None, truth_tid).  Fixed-length reads, so off = arange * read_len.
    Expression comes from `expr_seed` (default: seed); the reads themselves from (seed, stream), so a large read set can be
    produced chunk by chunk (stream = chunk number) or shard by shard over one expression profile."""
    erng = np.random.default_rng(seed if expr_seed is None else expr_seed)
    rng = np.random.default_rng([seed, stream])
    T = len(txp_len)
    expr = erng.lognormal(0.0, 2.0, size=T)
    expr[erng.random(T) < zero_frac] = 0.0
    min_len = read_len if not paired else max(read_len, 100)
    expr[txp_len < min_len] = 0.0
    w = expr * np.maximum(txp_len.astype(np.float64) - min_len + 1, 0)
    w /= w.sum()
    tid = rng.choice(T, size=n_reads, p=w)
    tl = txp_len[tid].astype(np.int64)
    if paired:
        fl = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=n_reads)), max(read_len, 100), 999).astype(np.int64)
        fl = np.minimum(fl, tl)
    else:
        fl = np.full(n_reads, read_len, np.int64)
    start = (rng.random(n_reads) * (tl - fl + 1)).astype(np.int64)
    base0 = txp_off[tid].astype(np.int64) + start
    ar = np.arange(read_len, dtype=np.int64)
    flip = rng.random(n_reads) < 0.5

    def mutate(b):
        if sub_rate > 0:
            m = rng.random(b.shape) < sub_rate
            b[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
        if n_rate > 0:
            m = rng.random(b.shape) < n_rate
            b[m] = ord("N")
        return b

    fwd = seq[base0[:, None] + ar[None, :]]                       # left end of the fragment, forward strand
    if not paired:
        rc = _COMP[fwd[:, ::-1]]
        b1 = np.where(flip[:, None], rc, fwd)
        b1 = mutate(np.ascontiguousarray(b1))
        off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
        return b1.reshape(-1), off, None, None, tid
    right = seq[(base0 + fl - read_len)[:, None] + ar[None, :]]   # right end, forward strand
    right_rc = _COMP[right[:, ::-1]]
    fwd_rc = _COMP[fwd[:, ::-1]]
    # inward pair: mate1 = left end fwd, mate2 = right end rc; flipped fragment swaps the roles
    b1 = np.where(flip[:, None], right_rc, fwd)
    b2 = np.where(flip[:, None], fwd, right_rc)
    del fwd_rc
    b1 = mutate(np.ascontiguousarray(b1)); b2 = mutate(np.ascontiguousarray(b2))
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return b1.reshape(-1), off, b2.reshape(-1), off.copy(), tid


def make_classes(n_txp, n_classes, seed=7, max_len=8, gene_size=5, long_frac=0.0):
    """Random equivalence classes (CSR, label-lexicographic, unique labels) for inference-only tests/benches."""
    rng = np.random.default_rng(seed)
    labels = set()
    n_genes = max(1, n_txp // gene_size)
    while len(labels) < n_classes:
        g = int(rng.integers(0, n_genes))
        lo = g * gene_size
        hi = min(n_txp, lo + gene_size)
        if rng.random() < long_frac:
            n = int(rng.integers(33, 80))
            ids = np.sort(rng.choice(n_txp, size=min(n, n_txp), replace=False))
        else:
            n = int(rng.integers(1, min(max_len, hi - lo) + 1))
            ids = np.sort(rng.choice(np.arange(lo, hi), size=n, replace=False))
        labels.add(tuple(int(x) for x in ids))
    labs = sorted(labels)
    row_ptr = np.zeros(len(labs) + 1, np.uint64)
    row_ptr[1:] = np.cumsum([len(l) for l in labs])
    flat = np.array([t for l in labs for t in l], dtype=np.uint32)
    counts = np.maximum(1, rng.lognormal(2.0, 2.0, size=len(labs))).astype(np.uint64)
    return row_ptr, flat, counts


def expression_weights(txp_len, read_len, paired, expr_seed, zero_frac=0.3):
    """the per-transcript sampling weights make_reads / make_reads_device draw fragments with (expression x effective length)"""
    erng = np.random.default_rng(expr_seed)
    T = len(txp_len)
    expr = erng.lognormal(0.0, 2.0, size=T)
    expr[erng.random(T) < zero_frac] = 0.0
    min_len = read_len if not paired else max(read_len, 100)
    expr[txp_len < min_len] = 0.0
    w = expr * np.maximum(txp_len.astype(np.float64) - min_len + 1, 0)
    return w / w.sum()


def make_reads_device(seq_d, txp_off, txp_len, n_reads, read_len, seed=1234, paired=False, frag_mean=200.0, frag_sd=25.0,
                      sub_rate=0.005, expr_seed=1234, chunk=4_000_000, out1=None, out2=None):
    """make_reads on the GPU with torch (plumbing for the 100 M-pair configurations, which numpy would take minutes to draw).
    seq_d: the transcriptome as a uint8 CUDA tensor.  Same model as make_reads (same expression profile for the same expr_seed),
    different random stream.  Fills / returns uint8 CUDA tensors of n_reads*read_len ASCII bases per mate and the true transcript
    of every fragment (int32)."""
    import torch
    dev = seq_d.device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    w = expression_weights(txp_len, read_len, paired, expr_seed)
    cdf = torch.from_numpy(np.cumsum(w)).to(dev)
    off_d = torch.from_numpy(txp_off.astype(np.int64)).to(dev)
    len_d = torch.from_numpy(txp_len.astype(np.int64)).to(dev)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    b1 = out1 if out1 is not None else torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    b2 = None
    if paired:
        b2 = out2 if out2 is not None else torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    truth = torch.empty(n_reads, dtype=torch.int32, device=dev)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    min_fl = max(read_len, 100)
    for a0 in range(0, n_reads, chunk):
        n = min(chunk, n_reads - a0)
        u = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        tid = torch.searchsorted(cdf, u).clamp_(max=len(txp_len) - 1)
        truth[a0:a0 + n] = tid.to(torch.int32)
        tl = len_d[tid]
        if paired:
            fl = torch.round(torch.randn(n, generator=g, device=dev) * frag_sd + frag_mean).to(torch.int64).clamp_(min_fl, 999)
            fl = torch.minimum(fl, tl)
        else:
            fl = torch.full((n,), read_len, dtype=torch.int64, device=dev)
        start = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (tl - fl + 1).to(torch.float64)).to(torch.int64)
        base0 = off_d[tid] + start
        flip = torch.rand(n, generator=g, device=dev) < 0.5

        def mutate(x):
            if sub_rate > 0:
                m = torch.rand(x.shape, generator=g, device=dev) < sub_rate
                r = acgt[torch.randint(0, 4, x.shape, generator=g, device=dev)]
                x = torch.where(m, r, x)
            return x

        fwd = seq_d[base0[:, None] + ar[None, :]]
        if not paired:
            rc = comp[fwd.flip(1).to(torch.int64)]
            b1[a0 * read_len:(a0 + n) * read_len] = mutate(torch.where(flip[:, None], rc, fwd)).reshape(-1)
            continue
        right = seq_d[(base0 + fl - read_len)[:, None] + ar[None, :]]
        right_rc = comp[right.flip(1).to(torch.int64)]
        b1[a0 * read_len:(a0 + n) * read_len] = mutate(torch.where(flip[:, None], right_rc, fwd)).reshape(-1)
        b2[a0 * read_len:(a0 + n) * read_len] = mutate(torch.where(flip[:, None], fwd, right_rc)).reshape(-1)
    return b1, b2, truth
