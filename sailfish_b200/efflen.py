"""Effective transcript lengths (host side, O(T + maxFragLen)) -- the tail of quasiMapReads.

Follows the reference's src/SailfishQuantify.cpp: getNormalFragLengthDist (:648-673), correctionFactorsFromCounts
(:769-807), computeSmoothedEffectiveLengths (:809-838), setEffectiveLengthsDirect (:707-715) and the mode selection at
:937-992 (paired) / :1034-1043 (single-end).  fp64 throughout, same operation order (running sums).
"""
import numpy as np


def normal_correction_factors(max_len, mean, sd):
    i = np.arange(max_len, dtype=np.float64)
    inv = 1.0 / sd
    x = inv * (i - mean)
    dens = np.exp(-0.5 * x * x) * inv
    cum_mass = np.cumsum(i * dens)
    cum_dens = np.cumsum(dens)
    cf = np.zeros(max_len, np.float64)
    ok = cum_dens > 0
    cf[ok] = cum_mass[ok] / cum_dens[ok]
    return cf


def correction_factors_from_counts(hist):
    hist = np.asarray(hist, dtype=np.uint32)
    n = len(hist)
    vals = np.zeros(n, np.float64)
    mult = np.zeros(n, np.uint32)
    cf = np.zeros(n, np.float64)
    mult[0] = hist[0]
    acc = 0.0
    m = np.uint32(hist[0])
    for i in range(1, n):                      # :789-801 (vals[0] stays 0: the loop starts at 1)
        acc = float(int(hist[i]) * i) + acc
        m = np.uint32((int(m) + int(hist[i])) & 0xFFFFFFFF)
        vals[i] = acc
        mult[i] = m
        if m > 0:
            cf[i] = acc / float(m)
    return cf


def empirical_effective_lengths(txp_len, fld_hist):
    """computeEmpiricalEffectiveLengths (--unsmoothedFLD, :717-767): sum_l pdf(l) * (RefLength - l + 1) with EmpiricalDistribution's
    float pdf over jointMap (every length 0 .. maxFragLen-1, :944-947); transcripts not longer than the median keep RefLength"""
    txp_len = np.asarray(txp_len, dtype=np.uint32)
    counts = np.asarray(fld_hist, dtype=np.uint32)
    n = len(counts)
    total = 0.0
    for c in counts:
        total += float(c)
    if not total > 0:                                               # nothing observed: no distribution to correct with
        return txp_len.astype(np.float64)
    cum, last, maxval = 0.0, 0, 1
    while last < n:
        cum += float(counts[last]) / total
        maxval = last
        if cum > 1.0 - 1e-6:
            break
        last += 1
    kept = 0.0
    for c in counts[:min(last, n)]:
        kept += float(c)
    pdf = np.zeros(n, np.float32)                                   # pdf(l) = 0 beyond the table
    pdf[:maxval] = (counts[:maxval].astype(np.float64) / kept).astype(np.float32)
    i, j = 0, n - 1                                                 # the median by walking in from both ends (EmpiricalDistribution.cpp:83-93)
    u, v = int(counts[0]), int(counts[n - 1])
    while i < j:
        if u <= v:
            v -= u; i += 1; u = int(counts[i])
        else:
            u -= v; j -= 1; v = int(counts[j])
    median = float(np.float32(i))
    max_val = n - 1
    eff = np.zeros(len(txp_len), np.float64)
    pd = pdf.astype(np.float64)
    for t, L in enumerate(txp_len.tolist()):
        if L <= median or not max_val > 0:
            eff[t] = L
            continue
        e = 0.0
        for l in range(0, min(L, max_val) + 1):                     # the reference's running sum
            e += pd[l] * (L - l + 1.0)
        eff[t] = e
    return eff


def effective_lengths(txp_len, fld_hist=None, max_frag_len=1000, num_frag_samples=10000, single_end=False,
                      no_correction=False, prior_mean=200.0, prior_sd=80.0, unsmoothed=False):
    """-> float64[T]: Transcript::EffectiveLength for the default (smoothed) mode, --unsmoothedFLD and --noEffectiveLengthCorrection."""
    txp_len = np.asarray(txp_len, dtype=np.uint32)
    if no_correction:
        return txp_len.astype(np.float64)
    enough = (not single_end) and fld_hist is not None and int(np.asarray(fld_hist, np.uint64).sum()) >= num_frag_samples
    if enough and unsmoothed:
        return empirical_effective_lengths(txp_len, fld_hist)
    cf = correction_factors_from_counts(fld_hist) if enough else normal_correction_factors(max_frag_len, prior_mean, prior_sd)
    idx = np.minimum(txp_len.astype(np.int64), max_frag_len - 1)
    eff = txp_len.astype(np.float64) - cf[idx] + 1.0
    return np.where(eff < 1.0, txp_len.astype(np.float64), eff)


def normal_frag_length_counts(max_len=1000, total_count=10000, mean=200.0, sd=80.0):
    """getNormalFragLengthCounts (:675-704): the prior normal as rounded counts -- what setFragLengthDist receives when fewer than
    numFragSamples fragment lengths were observed"""
    i = np.arange(max_len, dtype=np.float64)
    inv = 1.0 / sd
    x = inv * (i - mean)
    dens = np.exp(-0.5 * x * x) * inv
    total = 0.0
    for d in dens:                                   # the reference's running sum
        total += float(d)
    if not total > 0:
        return np.zeros(max_len, np.uint32)
    v = dens * total_count / total
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64).astype(np.uint32)      # std::round


def empirical_cdf(counts):
    """EmpiricalDistribution(pos = 0..n-1, counts) (src/EmpiricalDistribution.cpp:29-90) -> (float32 cdf table, maxValue()).
    The table stops where the cumulative mass passes 1 - 1e-6; pdf and cdf are float, the cdf a running float sum."""
    counts = np.asarray(counts, dtype=np.uint32)
    n = len(counts)
    total = 0.0
    for c in counts:
        total += float(c)
    if not total > 0:                                # nothing observed (the reference would divide by zero): an empty table
        return np.zeros(0, np.float32), (n - 1 if n else 0)
    cum, last, maxval = 0.0, 0, 1
    while last < n:
        cum += float(counts[last]) / total
        maxval = last
        if cum > 1.0 - 1e-6:
            break
        last += 1
    kept = 0.0
    for c in counts[:min(last, n)]:
        kept += float(c)
    cdf = np.zeros(maxval if n else 0, np.float32)
    run = np.float32(0.0)
    for v in range(len(cdf)):
        pdf = np.float32(float(counts[v]) / kept)
        run = pdf if v == 0 else np.float32(run + pdf)
        cdf[v] = run
    return cdf, (n - 1 if n else 0)
