// TEST INFRASTRUCTURE ONLY (see oracle.h).
// CPU restatement of the inference stage of kingsfordgroup/sailfish v0.10.0:
//   CollapsedEMOptimizer::optimize / EMUpdate_ / VBEMUpdate_ / gatherBootstraps / doBootstrap
//   CollapsedGibbsSampler::sample / initCountMap_ / sampleRound_, MultinomialSampler,
//   effective-length helpers of SailfishQuantify.cpp and the TPM formula of GZipWriter.cpp.
// fp64 throughout, operation order preserved, built -O3 WITHOUT -ffast-math.
#include "oracle.h"
#include "orc_empdist.hpp"
#include "orc_threads.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <random>
#include <vector>

namespace {

// CollapsedEMOptimizer.cpp:33-34
constexpr double minEQClassWeight = std::numeric_limits<double>::denorm_min();
constexpr double minWeight = std::numeric_limits<double>::denorm_min();

// ---- digamma: replaces boost::math::digamma (Boost is not in /root/reference) -------------------
// psi(x) for x > 0: recurrence psi(x) = psi(x+1) - 1/x up to x >= 12, then the asymptotic series
// ln x - 1/(2x) - sum_n B_2n / (2n x^2n).  Checked against scipy.special.digamma in tests/.
double digamma_pos(double x) {
    double acc = 0.0;
    while (x < 12.0) { acc -= 1.0 / x; x += 1.0; }
    const double inv = 1.0 / x, inv2 = inv * inv;
    // B2/2=1/12, B4/4=-1/120, B6/6=1/252, B8/8=-1/240, B10/10=1/132, B12/12=-691/32760, B14/14=1/12
    double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0 -
                    inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
    return acc + std::log(x) - 0.5 * inv - series;
}

struct Classes {
    uint32_t T;
    uint64_t E;
    const uint64_t* row_ptr;
    const uint32_t* labels;
    const uint64_t* counts;
};

// CollapsedEMOptimizer.cpp:745-772 (and :625-650, :527-555): w_i = count/effLen[t_i]; w_i *= 1/sum(w)
void compute_weights(const Classes& c, const std::vector<double>& effLens, std::vector<double>& w, orc::Pool& pool) {
    w.assign(c.row_ptr[c.E], 0.0);
    pool.parallel_for(c.E, [&](size_t b, size_t e) {
        for (size_t eq = b; eq < e; ++eq) {
            double wsum = 0.0;
            const double cnt = static_cast<double>(c.counts[eq]);
            for (uint64_t j = c.row_ptr[eq]; j < c.row_ptr[eq + 1]; ++j) {
                w[j] = cnt / effLens[c.labels[j]];
                wsum += w[j];
            }
            const double wnorm = 1.0 / wsum;
            for (uint64_t j = c.row_ptr[eq]; j < c.row_ptr[eq + 1]; ++j) w[j] *= wnorm;
        }
    });
}

// CollapsedEMOptimizer.cpp:70-79 (CAS loop) / :85-87 (plain)
inline void incLoop(std::atomic<double>& val, double inc) {
    double oldv = val.load(std::memory_order_relaxed);
    while (!val.compare_exchange_weak(oldv, oldv + inc, std::memory_order_relaxed)) {}
}
inline void incLoop(double& val, double inc) { val += inc; }

// One class of EMUpdate_ (CollapsedEMOptimizer.cpp:105-140 serial / :235-277 parallel)
template <typename VecIn, typename VecOut>
inline void em_class(const Classes& c, const std::vector<double>& w, const uint8_t* valid, size_t eq,
                     const VecIn& alphaIn, VecOut& alphaOut, const uint64_t* counts) {
    if (valid && !valid[eq]) return;
    const uint64_t count = counts[eq];
    const uint64_t b = c.row_ptr[eq], e = c.row_ptr[eq + 1];
    const size_t groupSize = e - b;
    if (groupSize > 1) {
        double denom = 0.0;
        for (uint64_t j = b; j < e; ++j) denom += alphaIn[c.labels[j]] * w[j];
        if (denom <= minEQClassWeight) {
            // class skipped
        } else {
            const double invDenom = count / denom;
            for (uint64_t j = b; j < e; ++j) {
                const double v = alphaIn[c.labels[j]] * w[j];
                if (!std::isnan(v)) incLoop(alphaOut[c.labels[j]], v * invDenom);
            }
        }
    } else {
        incLoop(alphaOut[c.labels[b]], static_cast<double>(count));
    }
}

// One class of VBEMUpdate_ (:180-216 serial / :325-366 parallel)
template <typename VecOut>
inline void vbem_class(const Classes& c, const std::vector<double>& w, const uint8_t* valid, size_t eq,
                       const std::vector<double>& expTheta, VecOut& alphaOut, const uint64_t* counts) {
    if (valid && !valid[eq]) return;
    const uint64_t count = counts[eq];
    const uint64_t b = c.row_ptr[eq], e = c.row_ptr[eq + 1];
    if (e - b > 1) {
        double denom = 0.0;
        for (uint64_t j = b; j < e; ++j) {
            const double th = expTheta[c.labels[j]];
            if (th > 0.0) denom += th * w[j];
        }
        if (denom <= minEQClassWeight) {
        } else {
            const double invDenom = count / denom;
            for (uint64_t j = b; j < e; ++j) {
                const double th = expTheta[c.labels[j]];
                if (th > 0.0) incLoop(alphaOut[c.labels[j]], (th * w[j]) * invDenom);
            }
        }
    } else {
        incLoop(alphaOut[c.labels[b]], static_cast<double>(count));
    }
}

// CollapsedEMOptimizer.cpp:37-44
double truncateCountVector(std::vector<double>& alphas, double cutoff) {
    double alphaSum = 0.0;
    for (size_t i = 0; i < alphas.size(); ++i) {
        if (alphas[i] <= cutoff) alphas[i] = 0.0;
        alphaSum += alphas[i];
    }
    return alphaSum;
}

// The iteration loop shared by optimize() (parallel, gate on NEW alpha, minIter) and doBootstrap()
// (serial, gate on OLD alpha, no minIter).  `gate_old` selects the convergence gate (:499 vs :852).
struct LoopCfg {
    bool use_vb; double prior; double tol; uint32_t min_iter, max_iter, fixed_iters; double check_cutoff; bool gate_old;
};

// `at_top`, when given, runs at the top of every iteration with the iteration number (optimize()'s effective-length
// recomputation at iterations 50 / 500 / 1000, :824-840; it may rewrite `w` in place).
void em_loop(const Classes& c, const std::vector<double>& w, const uint8_t* valid, const uint64_t* counts,
             const LoopCfg& cfg, orc::Pool& pool, std::vector<double>& alphas, uint32_t* iters_out, double* mrd_out,
             const std::function<void(uint32_t)>* at_top = nullptr) {
    const size_t T = c.T;
    std::vector<double> expTheta(T, 0.0);
    const bool par = pool.size() > 1;
    std::vector<std::atomic<double>> aPrimeAt(par ? T : 0);
    std::vector<double> aPrime(par ? 0 : T, 0.0);
    if (par) for (auto& a : aPrimeAt) a.store(0.0, std::memory_order_relaxed);

    bool converged = false;
    double maxRelDiff = -std::numeric_limits<double>::max();
    uint32_t itNum = 0;
    auto keep_going = [&]() {
        if (cfg.fixed_iters > 0) return itNum < cfg.fixed_iters;
        return itNum < cfg.min_iter || (itNum < cfg.max_iter && !converged);      // :820 / :486
    };
    while (keep_going()) {
        if (at_top) (*at_top)(itNum);
        if (cfg.use_vb) {
            double alphaSum = 0.0;                                                 // :300-303 / :162-165
            for (size_t i = 0; i < T; ++i) alphaSum += alphas[i];
            const double logNorm = digamma_pos(alphaSum);
            pool.parallel_for(T, [&](size_t b, size_t e) {                         // :305-320 / :171-178
                for (size_t i = b; i < e; ++i) {
                    expTheta[i] = (alphas[i] > minWeight) ? std::exp(digamma_pos(alphas[i]) - logNorm) : 0.0;
                    if (par) aPrimeAt[i].store(cfg.prior, std::memory_order_relaxed); else aPrime[i] = cfg.prior;
                }
            });
            pool.parallel_for(c.E, [&](size_t b, size_t e) {
                for (size_t eq = b; eq < e; ++eq) {
                    if (par) vbem_class(c, w, valid, eq, expTheta, aPrimeAt, counts);
                    else vbem_class(c, w, valid, eq, expTheta, aPrime, counts);
                }
            });
        } else {
            pool.parallel_for(c.E, [&](size_t b, size_t e) {
                for (size_t eq = b; eq < e; ++eq) {
                    if (par) em_class(c, w, valid, eq, alphas, aPrimeAt, counts);
                    else em_class(c, w, valid, eq, alphas, aPrime, counts);
                }
            });
        }
        converged = true;                                                          // :849-861 / :496-508
        maxRelDiff = -std::numeric_limits<double>::max();
        for (size_t i = 0; i < T; ++i) {
            const double ap = par ? aPrimeAt[i].load(std::memory_order_relaxed) : aPrime[i];
            const double gate = cfg.gate_old ? alphas[i] : ap;
            if (gate > cfg.check_cutoff) {
                const double relDiff = std::fabs(alphas[i] - ap) / ap;
                maxRelDiff = (relDiff > maxRelDiff) ? relDiff : maxRelDiff;
                if (relDiff > cfg.tol) converged = false;
            }
            alphas[i] = ap;
            if (par) aPrimeAt[i].store(0.0, std::memory_order_relaxed); else aPrime[i] = 0.0;
        }
        ++itNum;
    }
    if (iters_out) *iters_out = itNum;
    if (mrd_out) *mrd_out = maxRelDiff;
}

void clamp_eff(const double* eff_in, uint32_t T, std::vector<double>& effLens) {   // :733-740
    effLens.resize(T);
    for (uint32_t i = 0; i < T; ++i) effLens[i] = (eff_in[i] <= 1.0) ? 1.0 : eff_in[i];
}

size_t mark_active(const Classes& c, std::vector<uint8_t>& active) {                // :774-782
    active.assign(c.T, 0);
    size_t n = 0;
    for (uint64_t j = 0; j < c.row_ptr[c.E]; ++j) {
        if (!active[c.labels[j]]) { active[c.labels[j]] = 1; ++n; }
    }
    return n;
}

// MultinomialSampler::operator() (include/MultinomialSampler.hpp:13-64); n and k are uint32_t there.
// z[i] = sum_{j<i} p[j] accumulated left to right: identical values to the reference's O(k^2) table.
struct Multinomial {
    std::mt19937 gen;
    std::uniform_real_distribution<> u01{0.0, 1.0};
    std::vector<double> z;
    explicit Multinomial(uint64_t seed) : gen(static_cast<uint32_t>(seed ^ (seed >> 32))) {}
    void operator()(uint64_t* sample, uint32_t n, uint32_t k, const double* probs) {
        z.assign(static_cast<size_t>(k) + 1, 0.0);
        for (uint32_t i = 0; i < k; ++i) sample[i] = 0;
        double sum = 0.0;
        for (uint32_t i = 1; i <= k; ++i) { sum += probs[i - 1]; z[i] = sum; }
        if (k <= 100) {                                                              // :37-47
            for (uint32_t j = 0; j < n; ++j) {
                const double u = u01(gen);
                for (uint32_t i = 0; i < k; ++i) {
                    if ((z[i] < u) && (u <= z[i + 1])) { sample[i]++; break; }
                }
            }
        } else {                                                                     // :48-63
            for (uint32_t j = 0; j < n; ++j) {
                const double u = u01(gen);
                auto it = std::lower_bound(z.begin(), z.end() - 1, u);
                size_t offset = static_cast<size_t>(std::distance(z.begin(), it));
                if (*it > u && offset > 0) offset -= 1;
                sample[offset]++;
            }
        }
    }
};

}  // namespace

extern "C" double orc_digamma(double x) { return digamma_pos(x); }

extern "C" void orc_em_default_opts(orc_em_opts* o) {
    o->use_vb = 0; o->prior_alpha = 0.01; o->tol = 0.01; o->min_iter = 50; o->max_iter = 10000;
    o->fixed_iters = 0; o->check_cutoff = 1e-2; o->min_alpha = 1e-8;
}

// CollapsedEMOptimizer::optimize (CollapsedEMOptimizer.cpp:711-893)
extern "C" int orc_em_run(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                          const uint64_t* counts, const double* eff_lens, uint64_t num_mapped, const orc_em_opts* o,
                          int n_threads, double* alphas_out, uint32_t* iters_out, double* mrd_out) {
    Classes c{n_txp, n_classes, row_ptr, labels, counts};
    orc::Pool pool(n_threads);
    std::vector<double> effLens; clamp_eff(eff_lens, n_txp, effLens);
    std::vector<double> w; compute_weights(c, effLens, w, pool);
    std::vector<uint8_t> active;
    const size_t nActive = mark_active(c, active);
    if (nActive == 0) return -1;                                                     // :794-798
    const double totalNumFrags = static_cast<double>(num_mapped);                    // :792
    const double scale = 1.0 / nActive;                                              // :800
    std::vector<double> alphas(n_txp);
    for (uint32_t i = 0; i < n_txp; ++i) alphas[i] = active[i] ? scale * totalNumFrags : 0.0;

    LoopCfg cfg{o->use_vb != 0, o->prior_alpha, o->tol, o->min_iter, o->max_iter, o->fixed_iters, o->check_cutoff, false};
    em_loop(c, w, nullptr, counts, cfg, pool, alphas, iters_out, mrd_out);

    const double cutoff = cfg.use_vb ? (o->prior_alpha + o->min_alpha) : o->min_alpha;   // :812
    const double alphaSum = truncateCountVector(alphas, cutoff);                     // :875
    std::memcpy(alphas_out, alphas.data(), sizeof(double) * n_txp);
    if (alphaSum < minWeight) return -2;                                             // :877-881
    return 0;
}

// optimize() with --biasCorrect / --gcBiasCorrect (CollapsedEMOptimizer.cpp:717,816,824-840,888): at iterations 50, 500 and 1000
// the effective lengths are recomputed from the current alphas (orc_update_eff_lens) and the class weights rebuilt from them;
// eff_out receives the lengths the reference stores back into Transcript::EffectiveLength.
extern "C" int orc_update_eff_lens(int mode, uint32_t T, const char* seq, const uint64_t* off, const uint32_t* len, const double* eff_model,
                                   const double* eff_in, const double* alphas, int64_t num_fwd, int64_t num_rc, const uint32_t* read_bias,
                                   const uint32_t* observed_gc, const uint32_t* fld_counts, uint32_t n_fld, uint32_t gc_samp, double* eff_out);
extern "C" int orc_em_run_bias(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels, const uint64_t* counts,
                               const double* eff_lens, uint64_t num_mapped, const orc_em_opts* o, int n_threads, int mode, const char* seq,
                               const uint64_t* off, const uint32_t* len, int64_t num_fwd, int64_t num_rc, const uint32_t* read_bias,
                               const uint32_t* observed_gc, const uint32_t* fld_counts, uint32_t n_fld, uint32_t gc_samp,
                               double* alphas_out, double* eff_out, uint32_t* iters_out, double* mrd_out) {
    Classes c{n_txp, n_classes, row_ptr, labels, counts};
    orc::Pool pool(n_threads);
    std::vector<double> effLens; clamp_eff(eff_lens, n_txp, effLens);
    std::vector<double> w; compute_weights(c, effLens, w, pool);
    std::vector<uint8_t> active;
    const size_t nActive = mark_active(c, active);
    if (nActive == 0) return -1;
    const double totalNumFrags = static_cast<double>(num_mapped);
    const double scale = 1.0 / nActive;
    std::vector<double> alphas(n_txp);
    for (uint32_t i = 0; i < n_txp; ++i) alphas[i] = active[i] ? scale * totalNumFrags : 0.0;
    int bias_rc = 0;
    const std::function<void(uint32_t)> recompute = [&](uint32_t itNum) {               // :824-840
        if (itNum != 50 && itNum != 500 && itNum != 1000) return;
        std::vector<double> next(n_txp);
        const int rc = orc_update_eff_lens(mode, n_txp, seq, off, len, eff_lens /* Transcript::EffectiveLength */, effLens.data(), alphas.data(),
                                           num_fwd, num_rc, read_bias, observed_gc, fld_counts, n_fld, gc_samp, next.data());
        if (rc) { bias_rc = rc; return; }
        effLens.swap(next);
        compute_weights(c, effLens, w, pool);                                          // updateEqClassWeights :527-556
    };
    LoopCfg cfg{o->use_vb != 0, o->prior_alpha, o->tol, o->min_iter, o->max_iter, o->fixed_iters, o->check_cutoff, false};
    em_loop(c, w, nullptr, counts, cfg, pool, alphas, iters_out, mrd_out, &recompute);
    if (bias_rc) return bias_rc;
    const double cutoff = cfg.use_vb ? (o->prior_alpha + o->min_alpha) : o->min_alpha;
    const double alphaSum = truncateCountVector(alphas, cutoff);
    std::memcpy(alphas_out, alphas.data(), sizeof(double) * n_txp);
    if (eff_out) std::memcpy(eff_out, effLens.data(), sizeof(double) * n_txp);          // :888
    if (alphaSum < minWeight) return -2;
    return 0;
}

// GZipWriter::writeAbundances numerics (GZipWriter.cpp:216-241)
extern "C" void orc_tpm(uint32_t n_txp, const double* alphas, const double* eff_lens, uint64_t num_mapped, double* tpm) {
    const double numMappedFrags = static_cast<double>(num_mapped);
    double tfracDenom = 0.0;
    for (uint32_t i = 0; i < n_txp; ++i) tfracDenom += (alphas[i] / numMappedFrags) / eff_lens[i];
    for (uint32_t i = 0; i < n_txp; ++i) {
        const double npm = alphas[i] / numMappedFrags;
        const double tfrac = (npm / eff_lens[i]) / tfracDenom;
        tpm[i] = tfrac * 1000000.0;
    }
}

// gatherBootstraps + doBootstrap (CollapsedEMOptimizer.cpp:557-709, :438-525).  One worker, bootstraps in order.
// The reference seeds mt19937 from std::random_device; here sample b uses seed+b so runs are reproducible.
extern "C" int orc_bootstrap(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                             const uint64_t* counts, const double* eff_lens, const orc_em_opts* o, uint32_t n_boot,
                             uint64_t seed, orc_f64_row_cb cb, void* user) {
    Classes c{n_txp, n_classes, row_ptr, labels, counts};
    orc::Pool pool(1);
    std::vector<double> effLens; clamp_eff(eff_lens, n_txp, effLens);
    std::vector<uint8_t> active;
    const size_t nActive = mark_active(c, active);
    if (nActive == 0) return -1;                                                     // :606-610
    const double scale = 1.0 / nActive;
    // markDegenerateClasses (:372-433) runs on the weights left behind by optimize() and uniform alphas:
    // every member of a class is active, so denom > 0 and no class is ever dropped; kept for fidelity.
    std::vector<double> w; compute_weights(c, effLens, w, pool);
    std::vector<uint8_t> valid(n_classes, 1);
    for (uint64_t eq = 0; eq < n_classes; ++eq) {
        double denom = 0.0;
        for (uint64_t j = row_ptr[eq]; j < row_ptr[eq + 1]; ++j) {
            const double v = scale * w[j];   // alpha is the same positive constant for all active members
            if (!std::isnan(v)) denom += v;
        }
        if (denom <= minEQClassWeight) valid[eq] = 0;
    }
    uint64_t totalCount = 0;                                                         // :662-674
    for (uint64_t eq = 0; eq < n_classes; ++eq) if (valid[eq]) totalCount += counts[eq];
    const double floatCount = static_cast<double>(totalCount);
    std::vector<double> samplingWeights(n_classes, 0.0);                             // :676-680
    for (uint64_t eq = 0; eq < n_classes; ++eq) samplingWeights[eq] = valid[eq] ? counts[eq] / floatCount : 0.0;

    std::vector<uint64_t> sampCounts(n_classes, 0);
    std::vector<double> alphas(n_txp);
    LoopCfg cfg{o->use_vb != 0, o->prior_alpha, o->tol, 0, o->max_iter, o->fixed_iters, o->check_cutoff, true};
    const double cutoff = cfg.use_vb ? (o->prior_alpha + o->min_alpha) : o->min_alpha;   // :484
    for (uint32_t b = 0; b < n_boot; ++b) {
        Multinomial msamp(seed + b);
        msamp(sampCounts.data(), static_cast<uint32_t>(totalCount), static_cast<uint32_t>(n_classes),
              samplingWeights.data());                                               // :468
        for (uint32_t i = 0; i < n_txp; ++i) alphas[i] = active[i] ? scale * totalCount : 0.0;   // :471-474
        em_loop(c, w, valid.data(), sampCounts.data(), cfg, pool, alphas, nullptr, nullptr);
        const double alphaSum = truncateCountVector(alphas, cutoff);                 // :514
        if (alphaSum < minWeight) return -2;
        if (cb && cb(user, alphas.data(), n_txp) != 0) return -3;
    }
    return 0;
}

// Serial EM on caller-supplied (e.g. resampled) counts with the bootstrap loop rule (doBootstrap :476-514).
extern "C" int orc_bootstrap_em(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                                const uint64_t* samp_counts, const double* eff_lens, const orc_em_opts* o,
                                double* alphas_out, uint32_t* iters_out) {
    Classes c{n_txp, n_classes, row_ptr, labels, samp_counts};
    orc::Pool pool(1);
    std::vector<double> effLens; clamp_eff(eff_lens, n_txp, effLens);
    std::vector<uint8_t> active;
    const size_t nActive = mark_active(c, active);
    if (nActive == 0) return -1;
    // weights are computed from the ORIGINAL counts in the reference; count cancels in the normalisation,
    // except for classes whose count is 0 (0/eff -> 0, 1/0 -> inf, 0*inf -> nan).  Use count 1 to mirror "original".
    std::vector<uint64_t> ones(n_classes, 1);
    Classes cw{n_txp, n_classes, row_ptr, labels, ones.data()};
    std::vector<double> w; compute_weights(cw, effLens, w, pool);
    uint64_t totalCount = 0;
    for (uint64_t eq = 0; eq < n_classes; ++eq) totalCount += samp_counts[eq];
    const double scale = 1.0 / nActive;
    std::vector<double> alphas(n_txp);
    for (uint32_t i = 0; i < n_txp; ++i) alphas[i] = active[i] ? scale * totalCount : 0.0;
    LoopCfg cfg{o->use_vb != 0, o->prior_alpha, o->tol, 0, o->max_iter, o->fixed_iters, o->check_cutoff, true};
    em_loop(c, w, nullptr, samp_counts, cfg, pool, alphas, iters_out, nullptr);
    const double cutoff = cfg.use_vb ? (o->prior_alpha + o->min_alpha) : o->min_alpha;
    const double alphaSum = truncateCountVector(alphas, cutoff);
    std::memcpy(alphas_out, alphas.data(), sizeof(double) * n_txp);
    return alphaSum < minWeight ? -2 : 0;
}

// CollapsedGibbsSampler::sample (CollapsedGibbsSampler.cpp:199-270) as ONE chain (== one TBB range chunk).
extern "C" int orc_gibbs(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                         const uint64_t* counts, const double* eff_lens, const double* masses, uint64_t num_mapped,
                         uint32_t n_samples, uint64_t seed, orc_i32_row_cb cb, void* user) {
    Classes c{n_txp, n_classes, row_ptr, labels, counts};
    orc::Pool pool(1);
    std::vector<double> effLens; clamp_eff(eff_lens, n_txp, effLens);
    std::vector<double> w; compute_weights(c, effLens, w, pool);     // weights left in eqVec by optimize()
    const double priorAlpha = 1e-8;                                   // :215
    std::vector<double> mass(n_txp);
    for (uint32_t i = 0; i < n_txp; ++i) mass[i] = priorAlpha + masses[i] * static_cast<double>(num_mapped);   // :219-221

    Multinomial ms(seed);
    std::mt19937 fracGen(static_cast<uint32_t>((seed * 0x9E3779B97F4A7C15ULL) >> 32));
    std::uniform_real_distribution<> dis(0.25, 0.75);                 // :106
    const uint64_t nnz = row_ptr[n_classes];
    std::vector<uint64_t> countMap(nnz, 0);
    std::vector<double> probMap(nnz, 0.0);
    std::vector<int> txpCount(n_txp, 0);

    // initCountMap_ (:35-94)
    for (uint64_t eq = 0; eq < n_classes; ++eq) {
        const uint64_t b = row_ptr[eq], e = row_ptr[eq + 1];
        const size_t groupSize = e - b;
        const uint64_t classCount = counts[eq];
        double denom = 0.0;
        if (groupSize > 1) {
            for (uint64_t j = b; j < e; ++j) { denom += (priorAlpha + mass[labels[j]]) * w[j]; countMap[j] = 0; }
            if (denom > minEQClassWeight) {
                const double norm = 1.0 / denom;
                for (uint64_t j = b; j < e; ++j) probMap[j] = norm * ((priorAlpha + mass[labels[j]]) * w[j]);
                ms(&countMap[b], static_cast<uint32_t>(classCount), static_cast<uint32_t>(groupSize), &probMap[b]);
            }
        } else {
            countMap[b] = classCount;
        }
        for (uint64_t j = b; j < e; ++j) txpCount[labels[j]] += static_cast<int>(countMap[j]);
    }

    std::vector<uint64_t> txpResamp;
    std::vector<int32_t> row(n_txp);
    for (uint32_t s = 0; s < n_samples; ++s) {
        // `bool numInternalRounds = 10;` => exactly one round per sample (:248,257)
        for (uint64_t eq = 0; eq < n_classes; ++eq) {               // sampleRound_ (:96-186)
            const double sampleFrac = dis(fracGen);                 // drawn for every class (:115)
            const uint64_t b = row_ptr[eq], e = row_ptr[eq + 1];
            const size_t groupSize = e - b;
            if (groupSize > 1) {
                double denom = 0.0;
                uint64_t numResampled = 0;
                if (groupSize > txpResamp.size()) txpResamp.resize(groupSize, 0);
                for (uint64_t j = b; j < e; ++j) {
                    const uint32_t tid = labels[j];
                    const uint64_t currCount = countMap[j];
                    const uint64_t currResamp = static_cast<uint64_t>(std::round(sampleFrac * currCount));
                    numResampled += currResamp;
                    txpResamp[j - b] = currResamp;
                    txpCount[tid] -= static_cast<int>(currResamp);
                    countMap[j] -= currResamp;
                    denom += (priorAlpha + txpCount[tid]) * w[j];
                }
                if (denom > minEQClassWeight) {
                    const double norm = 1.0 / denom;
                    for (uint64_t j = b; j < e; ++j) probMap[j] = norm * ((priorAlpha + txpCount[labels[j]]) * w[j]);
                    ms(txpResamp.data(), static_cast<uint32_t>(numResampled), static_cast<uint32_t>(groupSize), &probMap[b]);
                }
                for (uint64_t j = b; j < e; ++j) {                  // :162-176 (both branches add txpResamp back)
                    countMap[j] += txpResamp[j - b];
                    txpCount[labels[j]] += static_cast<int>(txpResamp[j - b]);
                }
            }
        }
        for (uint32_t i = 0; i < n_txp; ++i) row[i] = txpCount[i];
        if (cb && cb(user, row.data(), n_txp) != 0) return -3;
    }
    return 0;
}

// ---- effective lengths (SailfishQuantify.cpp:648-838, 937-992, 1034-1043) ------------------------
namespace {
std::vector<double> normalFragLengthDist(uint32_t maxLen, double mean, double sd) {            // :648-673
    std::vector<double> cf(maxLen, 0.0);
    auto kernel = [mean, sd](double p) { double invStd = 1.0 / sd; double x = invStd * (p - mean); return std::exp(-0.5 * x * x) * invStd; };
    double cumulativeMass = 0.0, cumulativeDensity = 0.0;
    for (size_t i = 0; i < maxLen; ++i) {
        const double d = kernel(static_cast<double>(i));
        cumulativeMass += i * d;
        cumulativeDensity += d;
        if (cumulativeDensity > 0) cf[i] = cumulativeMass / cumulativeDensity;
    }
    return cf;
}
std::vector<double> correctionFactorsFromCounts(const uint32_t* hist, uint32_t maxLen) {       // :769-807
    std::vector<double> cf(maxLen, 0.0), vals(maxLen, 0.0);
    std::vector<uint32_t> mult(maxLen, 0);
    mult[0] = hist[0];
    for (size_t i = 1; i < maxLen; ++i) {
        const uint32_t v = hist[i];
        vals[i] = static_cast<double>(v * i) + vals[i - 1];   // uint32 * size_t -> 64-bit product (:797)
        mult[i] = v + mult[i - 1];
        if (mult[i] > 0) cf[i] = vals[i] / static_cast<double>(mult[i]);
    }
    return cf;
}
void smoothedEffLens(const uint32_t* len, uint32_t T, const std::vector<double>& cf, uint32_t maxLen, double* out) {   // :809-838
    for (uint32_t t = 0; t < T; ++t) {
        const uint32_t origLen = len[t];
        const double c = (origLen >= maxLen) ? cf[maxLen - 1] : cf[origLen];
        double effLen = static_cast<double>(origLen) - c + 1.0;
        if (effLen < 1.0) effLen = static_cast<double>(origLen);
        out[t] = effLen;
    }
}
}  // namespace

extern "C" int orc_eff_lens(const uint32_t* txp_len, uint32_t T, const uint32_t* fld_hist, uint32_t maxLen,
                            int32_t num_frag_samples, int single_end, int mode, double prior_mean, double prior_sd,
                            double* eff_out) {
    if (mode == 1) {                                                                  // setEffectiveLengthsDirect :707-715
        for (uint32_t t = 0; t < T; ++t) eff_out[t] = txp_len[t];
        return 0;
    }
    uint64_t seen = 0;
    if (!single_end && fld_hist) for (uint32_t i = 0; i < maxLen; ++i) seen += fld_hist[i];
    const bool enough = !single_end && fld_hist && static_cast<int64_t>(seen) >= num_frag_samples;   // remainingFLOps <= 0 (:966)
    if (!enough) {                                                                    // :966-976, :1039-1042
        smoothedEffLens(txp_len, T, normalFragLengthDist(maxLen, prior_mean, prior_sd), maxLen, eff_out);
        return 0;
    }
    if (mode == 2) {                                                                  // computeEmpiricalEffectiveLengths :717-767
        std::vector<uint32_t> vals, mults;
        for (uint32_t i = 0; i < maxLen; ++i) { vals.push_back(i); mults.push_back(fld_hist[i]); }   // jointMap holds every i (:944-947)
        EmpDist d; d.build(vals, mults);
        for (uint32_t t = 0; t < T; ++t) {
            const bool validSupport = d.maxVal > d.minVal;
            const double refLen = txp_len[t];
            if (refLen <= d.med || !validSupport) { eff_out[t] = refLen; continue; }
            double eff = 0.0;
            for (size_t l = d.minVal; l <= std::min(txp_len[t], d.maxVal); ++l) eff += d.pdf(static_cast<unsigned>(l)) * (txp_len[t] - l + 1.0);
            eff_out[t] = eff;
        }
        return 0;
    }
    smoothedEffLens(txp_len, T, correctionFactorsFromCounts(fld_hist, maxLen), maxLen, eff_out);   // :988-989
    return 0;
}
