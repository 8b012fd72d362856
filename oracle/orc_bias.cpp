// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's bias / GC effective-length correction:
// sailfish::utils::updateEffectiveLengths (reference src/SailfishUtils.cpp:611-926) with the helpers it uses:
// indexForKmer / nextKmerIndex (include/UtilityFunctions.hpp:40-140), Transcript::gcFrac over the inclusive GC prefix counts
// (include/Transcript.hpp:85-96,183-197, gcSampFactor 1), ReadKmerDist<6>::totalCount (include/ReadKmerDist.hpp:27-31) and
// EmpiricalDistribution's float cdf (orc_empdist.hpp).  Pinned against the reference's own function body, compiled unmodified
// into oracle/_ref/libsfref_em.so (oracle/Makefile cuts it out of the reference file at build time), by tests/test_oracle_bias.py
// and the committed fixture tests/golden/bias_efflens.npz.  Nothing under sailfish_b200/ may use this file.
//
// Sequences must be A/C/G/T (either case): for any other character the reference indexes its 4096-bin tables with
// UINT32_MAX (indexForKmer's error value) -- undefined behaviour this restatement refuses to imitate (returns -2).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_empdist.hpp"

namespace {

constexpr int K = 6;                     // ReadKmerDist<6, ...> (include/ReadExperiment.hpp:211)
constexpr uint32_t NK = 4096;

inline int code_fwd(char c) {            // UtilityFunctions.hpp:98-116
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return -1; }
}
// indexForKmer(s, K, FORWARD) (:96-119)
inline uint32_t index_fwd(const char* s) {
    uint32_t idx = 0;
    for (int i = 0; i < K; ++i) { idx += (uint32_t)code_fwd(s[i]); if (i < K - 1) idx <<= 2; }
    return idx;
}
// indexForKmer(s, K, REVERSE_COMPLEMENT) (:120-140): complement codes, last base first
inline uint32_t index_rc(const char* s) {
    uint32_t idx = 0;
    for (int i = K - 1; i >= 0; --i) { idx += (uint32_t)(3 - code_fwd(s[i])); if (i > 0) idx <<= 2; }
    return idx;
}
// nextKmerIndex (:40-86): shift in the next base (complemented for the reverse-complement direction), keep 2K bits
inline uint32_t next_index(uint32_t idx, char n, bool rc) {
    const int c = code_fwd(n);
    idx = (idx << 2) + (uint32_t)(rc ? 3 - c : c);
    return idx & (0xFFFFFFFFu >> (32 - 2 * K));
}

struct GC {                              // Transcript::computeGCContent_ / gcFrac with gcSampFactor 1
    std::vector<uint32_t> cnt;           // cnt[i] = G/C bases in [0, i]
    void build(const char* s, uint32_t n) {
        cnt.resize(n);
        uint32_t tot = 0;
        for (uint32_t i = 0; i < n; ++i) { const char c = (char)std::toupper((unsigned char)s[i]); if (c == 'G' || c == 'C') ++tot; cnt[i] = tot; }
    }
    int32_t frac(int32_t s, int32_t e) const { return (int32_t)std::lrint((100.0 * (cnt[e] - cnt[s])) / (e - s + 1)); }   // note: excludes base s
};

}  // namespace

// mode 1 = --biasCorrect (sequence-specific), 2 = --gcBiasCorrect.  eff_model = Transcript::EffectiveLength (the FLD model's
// value, truncated to int32 by the reference), eff_in = the optimizer's current vector.  Returns 0, or -2 for a non-ACGT base.
extern "C" int orc_update_eff_lens(int mode, uint32_t T, const char* seq, const uint64_t* off, const uint32_t* len, const double* eff_model,
                                   const double* eff_in, const double* alphas, int64_t num_fwd, int64_t num_rc, const uint32_t* read_bias,
                                   const uint32_t* observed_gc, const uint32_t* fld_counts, uint32_t n_fld, uint32_t gc_samp, double* eff_out) {
    const double minAlpha = 1e-8;
    const bool gcBias = mode == 2, seqBias = mode == 1;
    const int64_t numMappings = num_fwd + num_rc;
    for (uint32_t t = 0; t < T; ++t) eff_out[t] = eff_in[t];
    if (numMappings == 0 || (!gcBias && !seqBias)) return 0;                       // :627-632
    for (uint32_t t = 0; t < T; ++t)
        for (uint32_t i = 0; i < len[t]; ++i) if (code_fwd(seq[off[t] + i]) < 0) return -2;
    const double probFwd = static_cast<double>(num_fwd) / numMappings, probRC = static_cast<double>(num_rc) / numMappings;
    uint32_t tc = 0;                                                              // totalCount(): CountT (uint32) accumulator
    for (uint32_t i = 0; i < NK; ++i) tc += read_bias[i];
    const double readNormFactor = static_cast<double>(tc);
    std::vector<double> kdist(NK, 1.0), gdist(101, 1.0);
    EmpDist fld;
    { std::vector<uint32_t> pos(n_fld), cnt(fld_counts, fld_counts + n_fld); for (uint32_t i = 0; i < n_fld; ++i) pos[i] = i; fld.build(pos, cnt); }
    double readGCNormFactor = 0.0;
    int32_t fldLow = 0, fldHigh = 1;
    if (gcBias) {                                                                 // :668-687
        bool first = false, second = false;
        for (size_t i = 0; i <= fld.maxVal; ++i) {
            const float density = fld.cdf((unsigned)i);
            if (!first && density >= 0.005) { first = true; fldLow = (int32_t)i; }
            if (!second && density >= 0.995) { second = true; fldHigh = (int32_t)i; }
        }
        for (int i = 0; i < 101; ++i) readGCNormFactor += observed_gc[i];
    }
    const int32_t trunc = K;
    std::vector<GC> gc(gcBias ? T : 0);
    if (gcBias) for (uint32_t t = 0; t < T; ++t) gc[t].build(seq + off[t], len[t]);
    auto eligible = [&](uint32_t t, int32_t& refLen, int32_t& unprocessedLen) {
        refLen = static_cast<int32_t>(len[t]);
        const int32_t elen = static_cast<int32_t>(eff_model[t]);
        unprocessedLen = std::max(0, refLen - elen);
        return !(alphas[t] < minAlpha || unprocessedLen <= 0);
    };
    // ---- pass 1: expected distributions (:696-786)
    for (uint32_t t = 0; t < T; ++t) {
        int32_t refLen, unprocessedLen;
        if (!eligible(t, refLen, unprocessedLen)) continue;
        const double contribution = alphas[t] / eff_in[t];
        const char* tseq = seq + off[t];
        bool firstKmer = true; uint32_t idx = 0;
        for (int32_t i = refLen - trunc - 1; i >= 0; --i) {
            if (seqBias) {
                const int32_t fragStartPos = i + 2;
                if (firstKmer) { idx = index_rc(tseq + i); firstKmer = false; } else idx = next_index(idx, tseq[i], true);
                const int32_t maxFragLen = refLen - fragStartPos + 1;
                if (maxFragLen >= 0 && maxFragLen < refLen) kdist[idx] += probFwd * contribution * fld.cdf((unsigned)maxFragLen);
            }
            if (gcBias) {
                double prevFLMass = fld.cdf(0);
                for (int32_t fl = fldLow; fl <= fldHigh; fl += (int32_t)gc_samp) {
                    const int32_t fragEnd = i + fl - 1;
                    if (fragEnd < refLen) {
                        const int32_t g = gc[t].frac(i, fragEnd);
                        gdist[g] += contribution * (fld.cdf((unsigned)fl) - prevFLMass);
                        prevFLMass = fld.cdf((unsigned)fl);
                    } else break;
                }
            }
        }
        firstKmer = true; idx = 0;
        if (seqBias) {
            for (int32_t i = 0; i <= refLen - trunc - 1; ++i) {
                const int32_t kmerEndPos = i + K - 1, fragStartPos = i + 4;
                if (firstKmer) { idx = index_fwd(tseq); firstKmer = false; } else idx = next_index(idx, tseq[kmerEndPos], false);
                const int32_t maxFragLen = fragStartPos + 1;
                if (maxFragLen >= 0 && maxFragLen < refLen) kdist[idx] += probRC * contribution * fld.cdf((unsigned)maxFragLen);
            }
        }
    }
    // ---- priors and normalisers (:789-804)
    double txomeGCNormFactor = 0.0, gcPrior = 0.0;
    if (gcBias) { for (double m : gdist) txomeGCNormFactor += m; const double pmass = 101.0; gcPrior = ((pmass / (readGCNormFactor - pmass)) * txomeGCNormFactor) / 101.0; }
    double txomeNormFactor = 0.0, seqPrior = 0.0;
    if (seqBias) { for (double m : kdist) txomeNormFactor += m; const double pmass = static_cast<double>(NK); seqPrior = ((pmass / (readNormFactor - pmass)) * txomeNormFactor) / pmass; }
    // ---- pass 2: effective lengths (:811-924)
    std::vector<double> seqFactors, gcFactors;
    for (uint32_t t = 0; t < T; ++t) {
        double effLength = 0.0;
        int32_t refLen, unprocessedLen;
        const bool go = eligible(t, refLen, unprocessedLen);
        if (go) {
            seqFactors.assign(refLen, 0.0); gcFactors.assign(refLen, 0.0);
            const char* tseq = seq + off[t];
            bool firstKmer = true; uint32_t idx = 0;
            for (int32_t i = refLen - trunc - 1; i >= 0; --i) {
                if (seqBias) {
                    const int32_t fragStartPos = i + 2;
                    if (firstKmer) { idx = index_rc(tseq + i); firstKmer = false; } else idx = next_index(idx, tseq[i], true);
                    const int32_t maxFragLen = refLen - fragStartPos + 1;
                    if (fragStartPos >= 0 && fragStartPos < refLen)
                        seqFactors[fragStartPos] += probFwd * (read_bias[idx] / (kdist[idx] + seqPrior)) * fld.cdf((unsigned)maxFragLen);
                }
                if (gcBias) {
                    double prevFLMass = fld.cdf(0);
                    for (int32_t fl = fldLow; fl <= fldHigh; fl += (int32_t)gc_samp) {
                        const int32_t fragEnd = i + fl - 1;
                        if (fragEnd < refLen) {
                            const int32_t g = gc[t].frac(i, fragEnd);
                            const double sampleProb = (observed_gc[g] / (gcPrior + gdist[g])) * (fld.cdf((unsigned)fl) - prevFLMass);
                            prevFLMass = fld.cdf((unsigned)fl);
                            gcFactors[i] += sampleProb * probFwd;
                            gcFactors[fragEnd] += sampleProb * probRC;
                        } else break;
                    }
                }
            }
            firstKmer = true; idx = 0;
            if (seqBias) {
                for (int32_t i = 0; i <= refLen - trunc - 1; ++i) {
                    const int32_t kmerEndPos = i + K - 1, fragStartPos = i + 4;
                    if (firstKmer) { idx = index_fwd(tseq); firstKmer = false; } else idx = next_index(idx, tseq[kmerEndPos], false);
                    const int32_t maxFragLen = fragStartPos + 1;
                    if (fragStartPos >= 0 && fragStartPos < refLen)
                        seqFactors[fragStartPos] += probRC * (read_bias[idx] / (kdist[idx] + seqPrior)) * fld.cdf((unsigned)maxFragLen);
                }
            }
            if (seqBias) { double s = 0.0; for (double v : seqFactors) s += v; effLength = s; effLength *= (txomeNormFactor / readNormFactor); }
            else { double s = 0.0; for (double v : gcFactors) s += v; effLength = s; effLength *= (txomeGCNormFactor / readGCNormFactor); }
        }
        eff_out[t] = (unprocessedLen > 0.0 && effLength > unprocessedLen) ? effLength : eff_in[t];   // :916-922
    }
    return 0;
}

// EmpiricalDistribution over fragment-length counts (positions 0..n-1, as ReadExperiment::setFragLengthDist builds it): the float
// cdf table (cdf(x) = 1 beyond it) and maxValue().  Returns the table length.
extern "C" uint32_t orc_fld_cdf(const uint32_t* fld_counts, uint32_t n_fld, float* cdf_out, uint32_t cap, uint32_t* max_value) {
    EmpDist fld;
    std::vector<uint32_t> pos(n_fld), cnt(fld_counts, fld_counts + n_fld);
    for (uint32_t i = 0; i < n_fld; ++i) pos[i] = i;
    fld.build(pos, cnt);
    if (max_value) *max_value = fld.maxVal;
    for (uint32_t i = 0; i < cap && i < fld.cdfvals.size(); ++i) cdf_out[i] = fld.cdfvals[i];
    return static_cast<uint32_t>(fld.cdfvals.size());
}
