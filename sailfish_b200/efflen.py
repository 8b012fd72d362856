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


def effective_lengths(txp_len, fld_hist=None, max_frag_len=1000, num_frag_samples=10000, single_end=False,
                      no_correction=False, prior_mean=200.0, prior_sd=80.0):
    """-> float64[T]: Transcript::EffectiveLength for the default (smoothed) mode and --noEffectiveLengthCorrection."""
    txp_len = np.asarray(txp_len, dtype=np.uint32)
    if no_correction:
        return txp_len.astype(np.float64)
    enough = (not single_end) and fld_hist is not None and int(np.asarray(fld_hist, np.uint64).sum()) >= num_frag_samples
    cf = correction_factors_from_counts(fld_hist) if enough else normal_correction_factors(max_frag_len, prior_mean, prior_sd)
    idx = np.minimum(txp_len.astype(np.int64), max_frag_len - 1)
    eff = txp_len.astype(np.float64) - cf[idx] + 1.0
    return np.where(eff < 1.0, txp_len.astype(np.float64), eff)
