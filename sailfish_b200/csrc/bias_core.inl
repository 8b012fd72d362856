// bias_core.inl -- the per-position arithmetic of the bias / GC effective-length correction (bias.cu), written against four
// macros so that the same text is CUDA device code and plain host code for the CPU check (tests/bias_core_test.cpp):
//   SFB_BD            function qualifiers          SFB_LDG(p)      read-only load of *p
//   SFB_POPC64(x)     population count of a u64    SFB_D2I_RN(x)   double -> int, round to nearest even
// Text layout: the device index's 2-bit text (base p at bits 2*(p%32) of word p/32, codes A 0 C 1 G 2 T 3), transcripts
// concatenated without separators, txp_start[t] = first position of transcript t.

constexpr int BK = 6;                       // ReadKmerDist<6, ...> (include/ReadExperiment.hpp:211)
constexpr uint32_t BNK = 4096;

struct BiasView {
    const uint64_t* words; const uint64_t* txp_start; const uint32_t* txp_len; const uint32_t* gcw;   // gcw[w] = G/C bases in words [0, w)
    const float* cdf; uint32_t n_cdf;
    const double* eff_model; const double* eff_in; const double* alphas;
    uint32_t T;
    double probFwd, probRC;
    int32_t fldLow, fldHigh, gcSamp;
};

SFB_BD double b_cdf(const BiasView& v, int32_t x) { return (uint32_t)x < v.n_cdf ? (double)SFB_LDG(v.cdf + x) : 1.0; }   // EmpiricalDistribution::cdf (float)

// the six bases starting at text position p, base p in the two lowest bits
SFB_BD uint32_t b_win6(const uint64_t* __restrict__ w, uint64_t p) {
    const uint64_t idx = p >> 5; const uint32_t sh = 2 * (uint32_t)(p & 31);
    uint64_t x = SFB_LDG(w + idx) >> sh;
    if (sh > 52) x |= SFB_LDG(w + idx + 1) << (64 - sh);
    return (uint32_t)x & 0xFFFu;
}
// indexForKmer(s, 6, FORWARD) (include/UtilityFunctions.hpp:96-119): first base most significant
SFB_BD uint32_t b_idx_fwd(uint32_t win) {
    return ((win & 0x003u) << 10) | ((win & 0x00Cu) << 6) | ((win & 0x030u) << 2) | ((win & 0x0C0u) >> 2) | ((win & 0x300u) >> 6) | ((win & 0xC00u) >> 10);
}
// indexForKmer(s, 6, REVERSE_COMPLEMENT) (:120-140): complement of the last base most significant = bitwise not of the window
SFB_BD uint32_t b_idx_rc(uint32_t win) { return (~win) & 0xFFFu; }

// G/C bases of the text in [0, p)
SFB_BD uint32_t b_gc_upto(const BiasView& v, uint64_t p) {
    const uint64_t idx = p >> 5; const uint32_t r = (uint32_t)(p & 31);
    uint32_t n = SFB_LDG(v.gcw + idx);
    if (r) { const uint64_t w = SFB_LDG(v.words + idx); const uint64_t m = (w ^ (w >> 1)) & 0x5555555555555555ULL; n += SFB_POPC64(m & ((1ULL << (2 * r)) - 1)); }
    return n;
}
// Transcript::gcFrac(s, e) (include/Transcript.hpp:85-96): G/C bases in (s, e] over e - s + 1, rounded to nearest even
SFB_BD int32_t b_gc_frac(const BiasView& v, uint64_t t0, int32_t s, int32_t e) {
    const uint32_t n = b_gc_upto(v, t0 + e + 1) - b_gc_upto(v, t0 + s + 1);
    return SFB_D2I_RN((100.0 * n) / (double)(e - s + 1));
}

SFB_BD bool b_eligible(const BiasView& v, uint32_t t, int32_t& refLen, int32_t& unproc) {    // :712-722
    refLen = (int32_t)SFB_LDG(v.txp_len + t);
    const int32_t elen = (int32_t)SFB_LDG(v.eff_model + t);
    unproc = refLen - elen > 0 ? refLen - elen : 0;
    return !(SFB_LDG(v.alphas + t) < 1e-8 || unproc <= 0);
}

// ---- per-position bodies shared by the kernels and the CPU check ---------------------------------------------------------------
// pass 1, sequence bias (:728-741, :763-781): both strands' contributions of position i; add(bin, value)
template <typename Add>
SFB_BD void b_expected_seq(const BiasView& v, uint64_t t0, int32_t refLen, int32_t i, double contribution, Add add) {
    const uint32_t win = b_win6(v.words, t0 + i);
    add(b_idx_rc(win), v.probFwd * contribution * b_cdf(v, refLen - i - 1));        // forward: fragment starts at i + 2, at most refLen - i - 1 long
    if (i + 5 < refLen) add(b_idx_fwd(win), v.probRC * contribution * b_cdf(v, i + 5));   // reverse complement: "starts" at i + 4
}
// pass 1, fragment GC bias (:746-758)
template <typename Add>
SFB_BD void b_expected_gc(const BiasView& v, uint64_t t0, int32_t refLen, int32_t i, double contribution, Add add) {
    double prev = b_cdf(v, 0);
    for (int32_t fl = v.fldLow; fl <= v.fldHigh; fl += v.gcSamp) {
        const int32_t fragEnd = i + fl - 1;
        if (fragEnd >= refLen) break;
        const double cur = b_cdf(v, fl);
        add((uint32_t)b_gc_frac(v, t0, i, fragEnd), contribution * (cur - prev));
        prev = cur;
    }
}
// pass 2 (:828-838, :875-893 / :840-860): position i's share of the transcript's corrected length (before the normaliser)
SFB_BD double b_eff_seq(const BiasView& v, const double* ratio, uint64_t t0, int32_t refLen, int32_t i) {
    const uint32_t win = b_win6(v.words, t0 + i);
    double s = 0.0;
    if (i + 2 < refLen) s += v.probFwd * SFB_LDG(ratio + b_idx_rc(win)) * b_cdf(v, refLen - i - 1);
    if (i + 4 < refLen) s += v.probRC * SFB_LDG(ratio + b_idx_fwd(win)) * b_cdf(v, i + 5);
    return s;
}
SFB_BD double b_eff_gc(const BiasView& v, const double* ratio, uint64_t t0, int32_t refLen, int32_t i) {
    double prev = b_cdf(v, 0), s = 0.0;
    for (int32_t fl = v.fldLow; fl <= v.fldHigh; fl += v.gcSamp) {
        const int32_t fragEnd = i + fl - 1;
        if (fragEnd >= refLen) break;
        const double cur = b_cdf(v, fl);
        const double sampleProb = SFB_LDG(ratio + b_gc_frac(v, t0, i, fragEnd)) * (cur - prev);
        prev = cur;
        s += sampleProb * v.probFwd; s += sampleProb * v.probRC;                   // gcFactors[fragStart] and gcFactors[fragEnd]
    }
    return s;
}

// ---- the mapper's side: what one hit contributes to the observed distributions (SailfishQuantify.cpp:255-287, :372-389, :555-583) ---------
// G/C bases of the text in [a, b), straight from the 2-bit words (a fragment spans at most a few dozen words)
SFB_BD uint32_t b_gc_range(const uint64_t* __restrict__ words, uint64_t a, uint64_t b) {
    if (b <= a) return 0;
    const uint64_t w0 = a >> 5, w1 = (b - 1) >> 5;
    uint32_t n = 0;
    for (uint64_t w = w0; w <= w1; ++w) {
        const uint64_t x = SFB_LDG(words + w);
        uint64_t m = (x ^ (x >> 1)) & 0x5555555555555555ULL;
        if (w == w0) m &= ~0ULL << (2 * (uint32_t)(a & 31));
        if (w == w1) { const uint32_t hi = 2 * (uint32_t)((b - 1) & 31) + 2; if (hi < 64) m &= (1ULL << hi) - 1; }
        n += SFB_POPC64(m);
    }
    return n;
}
// Transcript::gcFrac(s, e) without the prefix array
SFB_BD int32_t b_gc_frac_range(const uint64_t* __restrict__ words, uint64_t t0, int32_t s, int32_t e) {
    return SFB_D2I_RN((100.0 * b_gc_range(words, t0 + s + 1, t0 + e + 1)) / (double)(e - s + 1));
}
// ReadKmerDist<6>::update (include/ReadKmerDist.hpp:36-72) for a hit at `pos` of a read of `readLen` bases: the bin of the 6-mer context
// around the read's start on the transcript (2 bases before it for a forward hit, reverse-complemented; 4 before for a
// reverse-complement hit), or -1 when the start lies outside (0, refLen) or the window does not fit
SFB_BD int32_t b_read_start_index(const uint64_t* __restrict__ words, uint64_t t0, int32_t refLen, int32_t pos, bool fwd, uint32_t readLen) {
    const int32_t startPos = fwd ? pos : pos + (int32_t)readLen;
    if (!(startPos > 0 && startPos < refLen)) return -1;
    if (fwd) {
        const int32_t p = startPos - 2;
        if (!(startPos >= 2 && p + BK < refLen)) return -1;
        return (int32_t)b_idx_rc(b_win6(words, t0 + p));
    }
    const int32_t p = startPos - 4;
    if (!(startPos >= 4 && p + BK < refLen)) return -1;
    return (int32_t)b_idx_fwd(b_win6(words, t0 + p));
}

// ---- fragment GC passes in sliding form ------------------------------------------------------------------------------------------
// For a fixed fragment length fl the bin of the fragment starting at i is lrint(100 n / fl), n = G/C among positions i+1 .. i+fl-1
// (gcFrac's divisor e - s + 1 IS fl), and moving the start by one base changes n by the base entering minus the base leaving.
// The weight of a fragment length, cdf(fl) - cdf(previous sampled fl), does not depend on the position either.  So one thread
// can own one fragment length and walk the transcript: no prefix loads, no fp64 division, no atomics (bias.cu, k_bias_*_slide).
//
// lrint(100 n / fl) in integers, ties to even.  Exact: 100 n / fl is a multiple of 1 / fl, so it is either a tie or at least
// 1 / (2 fl) >= 5e-4 away from one -- the correctly rounded double quotient rounds the same way
SFB_BD int32_t b_gc_bin(uint32_t n, uint32_t fl) {
    const uint32_t x = 200u * n + fl, d = 2u * fl;
    uint32_t q = x / d;
    if (q * d == x && (q & 1u)) --q;
    return (int32_t)q;
}
SFB_BD uint32_t b_gc_bit(const uint64_t* __restrict__ words, uint64_t p) {
    const uint64_t w = SFB_LDG(words + (p >> 5));
    const uint32_t sh = 2 * (uint32_t)(p & 31);
    return (uint32_t)(((w >> sh) ^ (w >> (sh + 1))) & 1ULL);
}
// last start of a fragment of fl bases the two passes visit on a transcript of refLen bases (i <= refLen - BK - 1 and
// fragEnd = i + fl - 1 < refLen); negative = none
SFB_BD int32_t b_gc_last_start(int32_t refLen, int32_t fl) { const int32_t a = refLen - BK - 1, b = refLen - fl; return a < b ? a : b; }
// visit(bin) for every start 0 .. b_gc_last_start; fl >= 1
template <typename Visit>
SFB_BD void b_gc_slide(const uint64_t* __restrict__ words, uint64_t t0, int32_t refLen, int32_t fl, Visit visit) {
    const int32_t hi = b_gc_last_start(refLen, fl);
    if (hi < 0) return;
    uint32_t n = b_gc_range(words, t0 + 1, t0 + (uint64_t)fl);
    for (int32_t i = 0; i <= hi; ++i) {
        visit(b_gc_bin(n, (uint32_t)fl));
        n += b_gc_bit(words, t0 + (uint64_t)(i + fl)) - b_gc_bit(words, t0 + (uint64_t)(i + 1));
    }
}
