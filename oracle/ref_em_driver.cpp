// TEST INFRASTRUCTURE ONLY.  Drives the reference's OWN inference code -- src/CollapsedEMOptimizer.cpp and
// src/CollapsedGibbsSampler.cpp compiled unmodified from /root/reference against the stub headers in
// oracle/shim_em/ (TBB, Boost, RapMap are not in the tree) -- so tests can pin the oracle's EM/VBEM restatement
// against the real thing.  Built by oracle/Makefile into oracle/_ref/libsfref_em.so.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <string>
#include <unistd.h>
#include <vector>

#include "CollapsedEMOptimizer.hpp"
#include "CollapsedGibbsSampler.hpp"
#include "ReadExperiment.hpp"
#include "SailfishOpts.hpp"
#include "spdlog/sinks/null_sink.h"

// boost::math::digamma stand-in (declared in shim_em/boost/math/special_functions/digamma.hpp): recurrence to x >= 12,
// then the asymptotic expansion; validated against scipy.special.digamma in tests/test_oracle_pins.py.
namespace boost { namespace math {
double digamma(double x) {
    double acc = 0.0;
    while (x < 12.0) { acc -= 1.0 / x; x += 1.0; }
    const double inv = 1.0 / x, inv2 = inv * inv;
    const double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0 -
                          inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
    return acc + std::log(x) - 0.5 * inv - series;
}
}}

// optimize() references this template for --biasCorrect / --gcBiasCorrect.  With SFREF_HAVE_BIAS the body is the reference's
// own text (src/SailfishUtils.cpp:611-926, cut out at build time into oracle/_ref/upd_efflens.inc by oracle/Makefile);
// otherwise a stub that must never be reached.
#ifdef SFREF_HAVE_BIAS
#include "ReadKmerDist.hpp"
#include "UtilityFunctions.hpp"
namespace sailfish { namespace utils {
#include "upd_efflens.inc"
template Eigen::VectorXd updateEffectiveLengths<std::vector<tbb::atomic<double>>>(
    SailfishOpts&, ReadExperiment&, Eigen::VectorXd&, std::vector<tbb::atomic<double>>&);
template Eigen::VectorXd updateEffectiveLengths<std::vector<double>>(
    SailfishOpts&, ReadExperiment&, Eigen::VectorXd&, std::vector<double>&);
}}
#else
namespace sailfish { namespace utils {
template <typename AbundanceVecT>
Eigen::VectorXd updateEffectiveLengths(SailfishOpts&, ReadExperiment&, Eigen::VectorXd& effLensIn, AbundanceVecT&) {
    std::fprintf(stderr, "ref_em_driver: bias correction is not part of this build\n");
    std::abort();
    return effLensIn;
}
template Eigen::VectorXd updateEffectiveLengths<std::vector<tbb::atomic<double>>>(
    SailfishOpts&, ReadExperiment&, Eigen::VectorXd&, std::vector<tbb::atomic<double>>&);
}}
#endif

int rapMapSAIndex(int, char*[]) { return 1; }

namespace {
struct Session {
    std::string dir;
    SailfishOpts sopt;
    std::vector<ReadLibrary> libs;
    std::unique_ptr<ReadExperiment> exp;
};
}

extern "C" {

// Build a ReadExperiment through the reference's own constructor (ReadExperiment.hpp:39-63) from (lengths, eff lengths),
// fill its EquivalenceClassBuilder through addGroup/finish, and set the mapped-fragment counter.
void* ref_em_session(uint32_t n_txp, const uint32_t* txp_len, const double* eff_len, uint64_t n_classes,
                     const uint64_t* row_ptr, const uint32_t* labels, const uint64_t* counts, uint64_t num_mapped,
                     int use_vb, int n_boot) {
    auto* s = new Session();
    char tmpl[] = "/tmp/sfref_XXXXXX";
    if (!mkdtemp(tmpl)) { delete s; return nullptr; }
    s->dir = tmpl;
    { std::ofstream f(s->dir + "/versionInfo.json"); f << "{\n \"indexVersion\": 2,\n \"kmerLength\": 31\n}\n"; }
    { std::ofstream f(s->dir + "/header.json"); cereal::JSONOutputArchive ar(f); IndexHeader h; ar(h); }
    { std::ofstream f(s->dir + "/txpinfo.txt"); for (uint32_t t = 0; t < n_txp; ++t) f << "t" << t << " " << txp_len[t] << "\n"; }
    auto sink = std::make_shared<spdlog::sinks::null_sink_st>();
    s->sopt.jointLog = std::make_shared<spdlog::logger>("refem", sink);
    s->sopt.numThreads = 1;
    s->sopt.useVBOpt = use_vb != 0;
    s->sopt.noEffectiveLengthCorrection = false;
    s->sopt.biasCorrect = false; s->sopt.gcBiasCorrect = false; s->sopt.gcSampFactor = 1;
    s->sopt.numBootstraps = n_boot; s->sopt.numGibbsSamples = 0;
    s->sopt.fragLenDistMax = 1000; s->sopt.fragLenDistPriorMean = 200; s->sopt.fragLenDistPriorSD = 80;
    boost::filesystem::path p(s->dir);
    s->exp.reset(new ReadExperiment(s->libs, p, s->sopt));
    auto& txps = s->exp->transcripts();
    for (uint32_t t = 0; t < n_txp; ++t) txps[t].EffectiveLength = eff_len[t];
    auto& eqb = s->exp->equivalenceClassBuilder();
    eqb.start();
    for (uint64_t e = 0; e < n_classes; ++e) {
        std::vector<uint32_t> ids(labels + row_ptr[e], labels + row_ptr[e + 1]);
        std::vector<double> aux(ids.size(), 1.0);
        TranscriptGroup tg(ids);
        eqb.addGroup(std::move(tg), aux);
    }
    eqb.finish();
    // one addGroup per class gave count 1; store the real counts on the flattened vector optimize() reads
    for (auto& kv : eqb.eqVec()) {
        const auto& ids = kv.first.txps;
        // find the class (labels are unique): linear probe through a hash of the first call order is overkill; use a map
        (void)ids;
    }
    {
        std::unordered_map<std::string, uint64_t> cnt;
        for (uint64_t e = 0; e < n_classes; ++e)
            cnt[std::string(reinterpret_cast<const char*>(labels + row_ptr[e]), 4 * (row_ptr[e + 1] - row_ptr[e]))] = counts[e];
        for (auto& kv : eqb.eqVec()) {
            std::string key(reinterpret_cast<const char*>(kv.first.txps.data()), 4 * kv.first.txps.size());
            kv.second.count.store(cnt[key]);
        }
    }
    s->exp->numMappedFragmentsAtomic().store(num_mapped);
    return s;
}

// Bias / GC correction.  A session as above plus what updateEffectiveLengths reads: the transcript
// sequences (through the reference's own Transcript::setSequence / GC tables), the read-start 6-mer counts
// (ReadExperiment::readBias()), the observed fragment GC histogram, the fragment length distribution (setFragLengthDist ->
// EmpiricalDistribution) and the strand tallies.  mode: 1 = --biasCorrect, 2 = --gcBiasCorrect.
void* ref_bias_session(uint32_t n_txp, const uint32_t* txp_len, const double* eff_len, const char* seq_concat, int mode,
                       const uint32_t* read_bias_counts /* 4096 */, const uint32_t* observed_gc /* 101 */,
                       const int32_t* fld_counts, uint32_t n_fld, int64_t num_fwd, int64_t num_rc, uint32_t gc_samp,
                       uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels, const uint64_t* counts, uint64_t num_mapped,
                       int use_vb) {
#ifndef SFREF_HAVE_BIAS
    return nullptr;
#else
    auto* s = new Session();
    char tmpl[] = "/tmp/sfref_XXXXXX";
    if (!mkdtemp(tmpl)) { delete s; return nullptr; }
    s->dir = tmpl;
    { std::ofstream f(s->dir + "/versionInfo.json"); f << "{\n \"indexVersion\": 2,\n \"kmerLength\": 31\n}\n"; }
    { std::ofstream f(s->dir + "/header.json"); cereal::JSONOutputArchive ar(f); IndexHeader h; ar(h); }
    { std::ofstream f(s->dir + "/txpinfo.txt"); for (uint32_t t = 0; t < n_txp; ++t) f << "t" << t << " " << txp_len[t] << "\n"; }
    { std::ofstream f(s->dir + "/seq.txt"); size_t o = 0; for (uint32_t t = 0; t < n_txp; ++t) { f.write(seq_concat + o, txp_len[t]); f << "\n"; o += txp_len[t]; } }
    auto sink = std::make_shared<spdlog::sinks::null_sink_st>();
    s->sopt.jointLog = std::make_shared<spdlog::logger>("refbias", sink);
    s->sopt.numThreads = 1;
    s->sopt.useVBOpt = use_vb != 0;
    s->sopt.noEffectiveLengthCorrection = false;
    s->sopt.biasCorrect = mode == 1; s->sopt.gcBiasCorrect = mode == 2; s->sopt.gcSampFactor = 1; s->sopt.pdfSampFactor = gc_samp;
    s->sopt.numBootstraps = 0; s->sopt.numGibbsSamples = 0;
    s->sopt.fragLenDistMax = 1000; s->sopt.fragLenDistPriorMean = 200; s->sopt.fragLenDistPriorSD = 80;
    boost::filesystem::path p(s->dir);
    s->exp.reset(new ReadExperiment(s->libs, p, s->sopt));
    auto& txps = s->exp->transcripts();
    for (uint32_t t = 0; t < n_txp; ++t) txps[t].EffectiveLength = eff_len[t];
    for (size_t i = 0; i < 4096; ++i) s->exp->readBias().counts[i].store(read_bias_counts[i]);
    for (size_t i = 0; i < 101; ++i) s->exp->observedGC()[i].store(observed_gc[i]);
    s->exp->setFragLengthDist(std::vector<int32_t>(fld_counts, fld_counts + n_fld));
    s->exp->addNumFwd(static_cast<int32_t>(num_fwd)); s->exp->addNumRC(static_cast<int32_t>(num_rc));
    if (n_classes) {
        auto& eqb = s->exp->equivalenceClassBuilder();
        eqb.start();
        for (uint64_t e = 0; e < n_classes; ++e) {
            std::vector<uint32_t> ids(labels + row_ptr[e], labels + row_ptr[e + 1]);
            std::vector<double> aux(ids.size(), 1.0);
            TranscriptGroup tg(ids);
            eqb.addGroup(std::move(tg), aux);
        }
        eqb.finish();
        std::unordered_map<std::string, uint64_t> cnt;
        for (uint64_t e = 0; e < n_classes; ++e)
            cnt[std::string(reinterpret_cast<const char*>(labels + row_ptr[e]), 4 * (row_ptr[e + 1] - row_ptr[e]))] = counts[e];
        for (auto& kv : eqb.eqVec()) {
            std::string key(reinterpret_cast<const char*>(kv.first.txps.data()), 4 * kv.first.txps.size());
            kv.second.count.store(cnt[key]);
        }
    }
    s->exp->numMappedFragmentsAtomic().store(num_mapped);
    return s;
#endif
}

// sailfish::utils::updateEffectiveLengths (src/SailfishUtils.cpp:611-926) on (alphas, effLensIn) -> effLensOut
int ref_bias_update(void* h, const double* alphas, const double* eff_in, double* eff_out) {
#ifndef SFREF_HAVE_BIAS
    return -1;
#else
    auto* s = static_cast<Session*>(h);
    const size_t T = s->exp->transcripts().size();
    std::vector<double> a(alphas, alphas + T);
    Eigen::VectorXd in(T);
    for (size_t i = 0; i < T; ++i) in(i) = eff_in[i];
    Eigen::VectorXd out = sailfish::utils::updateEffectiveLengths(s->sopt, *s->exp, in, a);
    for (size_t i = 0; i < T; ++i) eff_out[i] = out(i);
    return 0;
#endif
}

// the fragment length distribution as the reference's EmpiricalDistribution sees it (float cdf, maxValue)
uint32_t ref_bias_fld(void* h, float* cdf_out, uint32_t n) {
    auto* s = static_cast<Session*>(h);
    EmpiricalDistribution* d = s->exp->fragLengthDist();
    for (uint32_t i = 0; i < n; ++i) cdf_out[i] = d->cdf(i);
    return d->maxValue();
}

void ref_em_free(void* h) {
    auto* s = static_cast<Session*>(h);
    if (!s) return;
    std::string cmd = "rm -rf '" + s->dir + "'";
    if (std::system(cmd.c_str()) != 0) {}
    delete s;
}

// the order in which optimize() will walk the classes (libcuckoo bucket-major, cuckoohash_map.hh:1963-1977)
uint64_t ref_em_eq_order(void* h, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts) {
    auto* s = static_cast<Session*>(h);
    auto& v = s->exp->equivalenceClassBuilder().eqVec();
    if (row_ptr) {
        uint64_t e = 0, z = 0;
        row_ptr[0] = 0;
        for (auto& kv : v) {
            for (auto t : kv.first.txps) labels[z++] = t;
            counts[e] = kv.second.count.load();
            row_ptr[++e] = z;
        }
    }
    return v.size();
}

// CollapsedEMOptimizer::optimize (CollapsedEMOptimizer.cpp:711-893); outputs Transcript::estCount / mass
int ref_em_optimize(void* h, double tol, uint32_t max_iter, double* est_count, double* mass) {
    auto* s = static_cast<Session*>(h);
    CollapsedEMOptimizer opt;
    const bool ok = opt.optimize(*s->exp, s->sopt, tol, max_iter);
    auto& txps = s->exp->transcripts();
    for (size_t t = 0; t < txps.size(); ++t) { est_count[t] = txps[t].estCount(); mass[t] = txps[t].mass(); }
    return ok ? 0 : -1;
}

// ReadKmerDist<6>::update (include/ReadKmerDist.hpp:36-72) for a read that starts at txp[start_pos]: the bin it incremented, -1 if
// the context window did not fit
int ref_readbias_update(const char* txp, int len, int start_pos, int fwd) {
#ifndef SFREF_HAVE_BIAS
    return -2;
#else
    ReadKmerDist<6, uint32_t> d;
    const bool ok = d.update(txp, txp + start_pos, txp + len, fwd ? sailfish::utils::Direction::FORWARD : sailfish::utils::Direction::REVERSE_COMPLEMENT);
    if (!ok) return -1;
    for (size_t i = 0; i < d.counts.size(); ++i) if (d.counts[i] == 2) return static_cast<int>(i);
    return -1;
#endif
}
// Transcript::gcFrac(s, e) (include/Transcript.hpp:85-96) through the reference's own GC tables (gcSampFactor 1)
int ref_gc_frac(const char* seq, int len, int s, int e) {
    Transcript t(0, "t", static_cast<uint32_t>(len));
    t.setSequence(seq, true, 1);
    return t.gcFrac(s, e);
}

// Transcript::EffectiveLength of every transcript (optimize() stores the bias-corrected lengths there, :888)
void ref_txp_eff_lens(void* h, double* out) {
    auto* s = static_cast<Session*>(h);
    auto& txps = s->exp->transcripts();
    for (size_t t = 0; t < txps.size(); ++t) out[t] = txps[t].EffectiveLength;
}

// CollapsedEMOptimizer::gatherBootstraps (:557-709); rows appended to out[n_boot][n_txp]
int ref_em_bootstraps(void* h, double tol, uint32_t max_iter, double* out) {
    auto* s = static_cast<Session*>(h);
    CollapsedEMOptimizer opt;
    size_t row = 0;
    const size_t T = s->exp->transcripts().size();
    std::function<bool(const std::vector<double>&)> cb = [&](const std::vector<double>& a) -> bool {
        std::copy(a.begin(), a.end(), out + (row++) * T);
        return true;
    };
    return opt.gatherBootstraps(*s->exp, s->sopt, cb, tol, max_iter) ? 0 : -1;
}

// CollapsedGibbsSampler::sample (CollapsedGibbsSampler.cpp:199-291); rows appended to out[n_samples][n_txp]
int ref_em_gibbs(void* h, uint32_t n_samples, int32_t* out) {
    auto* s = static_cast<Session*>(h);
    CollapsedGibbsSampler g;
    size_t row = 0;
    const size_t T = s->exp->transcripts().size();
    std::function<bool(const std::vector<int>&)> cb = [&](const std::vector<int>& a) -> bool {
        std::copy(a.begin(), a.end(), out + (row++) * T);
        return true;
    };
    return g.sample(*s->exp, s->sopt, cb, n_samples) ? 0 : -1;
}

}  // extern "C"
