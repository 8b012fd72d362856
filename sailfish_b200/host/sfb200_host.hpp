// sfb200_host.hpp -- C++ adaptors with the reference's own class / method names on top of the C ABI (include/sfb200.h).
//
// Sailfish (kingsfordgroup/sailfish v0.10.0) has no plugin layer: the quantification hot path is five ordinary C++ call
// sites inside src/SailfishQuantify.cpp.  The classes below keep those signatures, so mainQuantify's body can
// call them unchanged (INTEGRATION.md shows the patch):
//
//   sfb200::GpuQuasiMapper::processReads        <- processReadsQuasi<IndexT>            SailfishQuantify.cpp:105-113, 458-464
//   sfb200::EquivalenceClassBuilder             <- EquivalenceClassBuilder              include/EquivalenceClassBuilder.hpp:53-117
//   sfb200::CollapsedEMOptimizer::optimize      <- CollapsedEMOptimizer::optimize       include/CollapsedEMOptimizer.hpp:25-28
//   sfb200::CollapsedEMOptimizer::gatherBootstraps <- ...::gatherBootstraps             include/CollapsedEMOptimizer.hpp:30-36
//   sfb200::CollapsedGibbsSampler::sample       <- CollapsedGibbsSampler::sample        include/CollapsedGibbsSampler.hpp:27-31
//
// They are templates over the experiment / options types, so they compile both against the reference's ReadExperiment /
// Transcript / SailfishOpts headers and against any type with the same members (tests/host_adaptor_test.cpp uses a mock).
// Header-only, plain C++11, no CUDA headers: link with -lsfb200.  There is no CPU fallback: without a CUDA device the
// constructor of Device throws.
#ifndef SFB200_HOST_HPP
#define SFB200_HOST_HPP

#include <cstdint>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfb200.h"

namespace sfb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// one sfb200_ctx (a CUDA device with its index, class table and inference scratch)
class Device {
public:
    // bindHost: keep the calling thread (and the parser / reader threads it starts later) on the GPU's NUMA node -- page-locked
    // batches are then node-local, which is what lets eight ranks of one host copy at the PCIe rate each (sfb200.h)
    explicit Device(int device = 0, bool bindHost = false) {
        if (bindHost) sfb200_bind_host_near_device(device);
        const int rc = sfb200_ctx_create(device, &ctx_);
        if (rc != SFB200_OK) throw Error(rc, "sfb200: no usable CUDA device (there is no CPU fallback)");
    }
    ~Device() { sfb200_ctx_destroy(ctx_); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    sfb200_ctx* get() const { return ctx_; }
    void check(int rc) const { if (rc != SFB200_OK) throw Error(rc, sfb200_last_error(ctx_)); }

    // what ReadExperiment takes from the quasi index (include/ReadExperiment.hpp:103-116): concatenated transcript
    // sequences, their offsets and lengths
    void buildIndex(const std::string& seq, const std::vector<uint64_t>& txpOffsets, const std::vector<uint32_t>& txpLens, int k) {
        check(sfb200_index_build(ctx_, seq.data(), txpOffsets.data(), txpLens.data(), static_cast<uint32_t>(txpLens.size()), k));
        nTxp_ = static_cast<uint32_t>(txpLens.size());
    }
    // the built index as one file and back (what SailfishIndex::load takes from the index directory, include/SailfishIndex.hpp:80-144)
    void saveIndex(const std::string& path) const { check(sfb200_index_save(ctx_, path.c_str())); }
    void loadIndex(const std::string& path, uint32_t nTxp) { check(sfb200_index_load(ctx_, path.c_str())); nTxp_ = nTxp; }
    uint32_t numTranscripts() const { return nTxp_; }
    void setNumTranscripts(uint32_t n) { nTxp_ = n; }

private:
    sfb200_ctx* ctx_ = nullptr;
    uint32_t nTxp_ = 0;
};

// ---- equivalence classes -----------------------------------------------------------------------------------------------------
// TranscriptGroup / TGValue as the rest of Sailfish sees them through eqVec() (EquivalenceClassBuilder.hpp:18-51,110)
struct TranscriptGroup { std::vector<uint32_t> txps; };
struct TGValue { std::vector<double> weights; uint64_t count; };

class EquivalenceClassBuilder {
public:
    explicit EquivalenceClassBuilder(Device& d) : dev_(d) {}
    // == start() (EquivalenceClassBuilder.hpp:62): resets the class table, the counters and the FLD sampler
    void start(const sfb200_map_opts& opts) { opts_ = opts; dev_.check(sfb200_map_begin(dev_.get(), &opts)); active_ = true; bias_ = false; }
    // sopt.biasCorrect / sopt.gcBiasCorrect: collect readExp.readBias() and readExp.observedGC() while mapping
    // (SailfishQuantify.cpp:255-287, :372-389, :555-583); call between start() and the first batch
    void collectBias(bool seqBias, bool gcBias, int32_t numBiasSamples) {
        dev_.check(sfb200_map_set_bias(dev_.get(), seqBias ? 1 : 0, gcBias ? 1 : 0, numBiasSamples));
        bias_ = true;
    }
    // == finish() (:64-80): flattens the table; the classes stay on the device for optimize()
    bool finish() {
        fld_.assign(opts_.max_frag_len, 0);
        dev_.check(sfb200_map_finish(dev_.get(), counters_, fld_.data(), &nClasses_, &nnz_));
        if (bias_) { readBias_.assign(4096, 0); observedGC_.assign(101, 0); dev_.check(sfb200_map_get_bias(dev_.get(), readBias_.data(), observedGC_.data())); }
        active_ = false;
        haveVec_ = false;
        return true;
    }
    // == eqVec() (:110): materialised on demand (aux/eq_classes.txt, --dumpEq); weights are 1/n as normalizeAux leaves them
    std::vector<std::pair<const TranscriptGroup, TGValue>>& eqVec() {
        if (!haveVec_) {
            std::vector<uint64_t> rowPtr(nClasses_ + 1), counts(nClasses_ ? nClasses_ : 1);
            std::vector<uint32_t> labels(nnz_ ? nnz_ : 1);
            dev_.check(sfb200_eq_export(dev_.get(), rowPtr.data(), labels.data(), counts.data()));
            countVec_.clear();
            countVec_.reserve(nClasses_);
            for (uint64_t e = 0; e < nClasses_; ++e) {
                TranscriptGroup tg;
                tg.txps.assign(labels.begin() + rowPtr[e], labels.begin() + rowPtr[e + 1]);
                TGValue v;
                v.count = counts[e];
                v.weights.assign(tg.txps.size(), 1.0 / static_cast<double>(tg.txps.size()));
                countVec_.emplace_back(std::move(tg), std::move(v));
            }
            haveVec_ = true;
        }
        return countVec_;
    }
    // the inverse (the commented-out loadEquivClasses, SailfishQuantify.cpp:1444-1495): classes from aux/eq_classes.txt
    void import(uint32_t nTxp, const std::vector<uint64_t>& rowPtr, const std::vector<uint32_t>& labels, const std::vector<uint64_t>& counts) {
        dev_.check(sfb200_eq_import(dev_.get(), nTxp, counts.size(), rowPtr.data(), labels.data(), counts.data()));
        dev_.setNumTranscripts(nTxp);
        nClasses_ = counts.size(); nnz_ = labels.size(); haveVec_ = false;
    }
    // ReadExperiment's atomics after the mapping threads joined (ReadExperiment.hpp:74-97)
    uint64_t numObservedFragments() const { return counters_[0]; }
    uint64_t numMappedFragments() const { return counters_[1]; }
    uint64_t numFragHits() const { return counters_[2]; }
    uint64_t upperBoundHits() const { return counters_[3]; }
    uint64_t numFwd() const { return counters_[4]; }
    uint64_t numRC() const { return counters_[5]; }
    const std::vector<uint32_t>& fragLengthCounts() const { return fld_; }      // flMap (SailfishQuantify.cpp:867)
    const std::vector<uint32_t>& readBiasCounts() const { return readBias_; }   // readExp.readBias().counts, after collectBias()
    const std::vector<uint32_t>& observedGC() const { return observedGC_; }     // readExp.observedGC()
    uint64_t numClasses() const { return nClasses_; }

private:
    Device& dev_;
    sfb200_map_opts opts_{};
    bool active_ = false, haveVec_ = false, bias_ = false;
    uint64_t counters_[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> fld_, readBias_, observedGC_;
    uint64_t nClasses_ = 0, nnz_ = 0;
    std::vector<std::pair<const TranscriptGroup, TGValue>> countVec_;
};

// ---- mapping -------------------------------------------------------------------------------------------------------------------
// processReadsQuasi (SailfishQuantify.cpp:105-452 paired, :458-646 single) for one parser job.  A job is whatever the parser
// hands a worker thread (paired_parser::job / single_parser::job, :175-180,510-515): `n` reads, read i = seq(i) or
// (seq1(i), seq2(i)).  Host worker threads may call this concurrently: batches are serialised on the device context.
class GpuQuasiMapper {
public:
    // Parser jobs are small (1000 reads, SailfishQuantify.cpp:73); a kernel launch wants far more.  Jobs are appended to a
    // staging batch under a mutex and the batch goes to the device when it holds `flushAt` fragments; call flush() after the
    // worker threads joined (before EquivalenceClassBuilder::finish()).
    explicit GpuQuasiMapper(Device& d, size_t flushAt = 1u << 20) : dev_(d), flushAt_(flushAt) { o1_.push_back(0); o2_.push_back(0); }
    template <typename SeqOf>
    void processReads(size_t n, SeqOf seq) {                                   // single-end
        std::lock_guard<std::mutex> lk(mu_);
        paired_ = false;
        for (size_t i = 0; i < n; ++i) { b1_.append(seq(i)); o1_.push_back(b1_.size()); }
        if (o1_.size() - 1 >= flushAt_) flushLocked();
    }
    template <typename SeqOf1, typename SeqOf2>
    void processReads(size_t n, SeqOf1 seq1, SeqOf2 seq2) {                     // paired-end
        std::lock_guard<std::mutex> lk(mu_);
        paired_ = true;
        for (size_t i = 0; i < n; ++i) {
            b1_.append(seq1(i)); o1_.push_back(b1_.size());
            b2_.append(seq2(i)); o2_.push_back(b2_.size());
        }
        if (o1_.size() - 1 >= flushAt_) flushLocked();
    }
    void flush() { std::lock_guard<std::mutex> lk(mu_); flushLocked(); }

private:
    void flushLocked() {
        const size_t n = o1_.size() - 1;
        if (n == 0) return;
        b1_.push_back('\0');
        if (paired_) {
            b2_.push_back('\0');
            dev_.check(sfb200_map_batch(dev_.get(), b1_.data(), o1_.data(), b2_.data(), o2_.data(), n));
        } else {
            dev_.check(sfb200_map_batch(dev_.get(), b1_.data(), o1_.data(), nullptr, nullptr, n));
        }
        b1_.clear(); b2_.clear(); o1_.assign(1, 0); o2_.assign(1, 0);
    }
    Device& dev_;
    size_t flushAt_;
    bool paired_ = false;
    std::mutex mu_;
    std::string b1_, b2_;
    std::vector<uint64_t> o1_, o2_;
};

// ---- inference -----------------------------------------------------------------------------------------------------------------
namespace detail {
template <typename ExpT, typename OptsT>
std::vector<double> effLens(ExpT& readExp, OptsT& sopt) {                        // CollapsedEMOptimizer.cpp:733-740
    auto& transcripts = readExp.transcripts();
    std::vector<double> e(transcripts.size());
    for (size_t i = 0; i < transcripts.size(); ++i)
        e[i] = sopt.noEffectiveLengthCorrection ? static_cast<double>(transcripts[i].RefLength) : transcripts[i].EffectiveLength;
    return e;
}
inline sfb200_em_opts emOpts(bool useVB, double tol, uint32_t maxIter) {
    sfb200_em_opts o;
    sfb200_em_default_opts(&o);
    o.use_vb = useVB ? 1 : 0; o.tol = tol; o.max_iter = maxIter;
    return o;
}
}  // namespace detail

// What updateEffectiveLengths reads from the experiment (src/SailfishUtils.cpp:611-690): the two observed distributions, the strand
// tallies and readExp.fragLengthDist().  fragLengthCounts = what setFragLengthDist received (ReadExperiment.hpp:160-167: the
// observed histogram, or getNormalFragLengthCounts when too few fragments were sampled); its cdf table is built as
// EmpiricalDistribution does (src/EmpiricalDistribution.cpp:29-90: float pdf / cdf, truncated where the cumulative mass passes
// 1 - 1e-6; maxValue() is the largest position, whatever its count).
class BiasModel {
public:
    BiasModel(bool gcBias, const std::vector<uint32_t>& readBias, const std::vector<uint32_t>& observedGC, uint64_t numFwd, uint64_t numRC,
              const std::vector<uint32_t>& fragLengthCounts, uint32_t pdfSampFactor = 1)
        : readBias_(readBias), observedGC_(observedGC) {
        const size_t n = fragLengthCounts.size();
        double total = 0.0;
        for (uint32_t c : fragLengthCounts) total += c;
        size_t last = 0, maxval = 1;
        double cum = 0.0;
        for (; last < n; ++last) { cum += fragLengthCounts[last] / total; maxval = last; if (cum > 1.0 - 1e-6) break; }
        double kept = 0.0;
        for (size_t i = 0; i < last && i < n; ++i) kept += fragLengthCounts[i];
        cdf_.resize(n ? maxval : 0);
        float run = 0.0f;
        for (size_t v = 0; v < cdf_.size(); ++v) {
            const float pdf = static_cast<float>(fragLengthCounts[v] / kept);
            run = v ? run + pdf : pdf;
            cdf_[v] = run;
        }
        m_.mode = gcBias ? 2 : 1; m_.gc_samp = pdfSampFactor ? pdfSampFactor : 1;
        m_.num_fwd = static_cast<int64_t>(numFwd); m_.num_rc = static_cast<int64_t>(numRC);
        m_.read_bias = readBias_.data(); m_.observed_gc = observedGC_.data();
        m_.fld_cdf = cdf_.data(); m_.n_cdf = static_cast<uint32_t>(cdf_.size()); m_.fld_max = n ? static_cast<uint32_t>(n - 1) : 0;
    }
    BiasModel(const BiasModel&) = delete;
    BiasModel& operator=(const BiasModel&) = delete;
    const sfb200_bias_model* get() const { return &m_; }
    const std::vector<float>& cdf() const { return cdf_; }

private:
    std::vector<uint32_t> readBias_, observedGC_;
    std::vector<float> cdf_;
    sfb200_bias_model m_;
};

class CollapsedEMOptimizer {
public:
    explicit CollapsedEMOptimizer(Device& d) : dev_(d) {}

    // bool CollapsedEMOptimizer::optimize(ReadExperiment&, SailfishOpts&, double relDiffTolerance, uint32_t maxIter)
    template <typename ExpT, typename OptsT>
    bool optimize(ExpT& readExp, OptsT& sopt, double relDiffTolerance = 0.01, uint32_t maxIter = 10000) {
        auto& transcripts = readExp.transcripts();
        const std::vector<double> eff = detail::effLens(readExp, sopt);
        std::vector<double> alphas(transcripts.size());
        const sfb200_em_opts o = detail::emOpts(sopt.useVBOpt, relDiffTolerance, maxIter);
        uint32_t iters = 0; double mrd = 0.0;
        const int rc = sfb200_em_run(dev_.get(), eff.data(), static_cast<uint32_t>(eff.size()), readExp.numMappedFragments(), &o,
                                     alphas.data(), &iters, &mrd);
        lastIters_ = iters;
        if (rc == SFB200_ENOACTIVE || rc == SFB200_ESMALLSUM) { lastError_ = sfb200_last_error(dev_.get()); return false; }   // :794-798, :877-881
        dev_.check(rc);
        double alphaSum = 0.0;
        for (double a : alphas) alphaSum += a;
        for (size_t i = 0; i < transcripts.size(); ++i) {                       // :883-890
            transcripts[i].setEstCount(alphas[i]);
            transcripts[i].setMass(alphas[i] / alphaSum);
        }
        return true;
    }

    // optimize() when sopt.biasCorrect or sopt.gcBiasCorrect is set (CollapsedEMOptimizer.cpp:820-840, :888): the effective lengths are
    // recomputed on the device at iterations 50 / 500 / 1000 and the transcripts leave with the corrected EffectiveLength
    template <typename ExpT, typename OptsT>
    bool optimizeWithBias(ExpT& readExp, OptsT& sopt, const BiasModel& model, double relDiffTolerance = 0.01, uint32_t maxIter = 10000) {
        auto& transcripts = readExp.transcripts();
        const std::vector<double> eff = detail::effLens(readExp, sopt);
        std::vector<double> alphas(transcripts.size()), effOut(transcripts.size());
        const sfb200_em_opts o = detail::emOpts(sopt.useVBOpt, relDiffTolerance, maxIter);
        uint32_t iters = 0; double mrd = 0.0;
        const int rc = sfb200_em_run_bias(dev_.get(), eff.data(), static_cast<uint32_t>(eff.size()), readExp.numMappedFragments(), &o, model.get(),
                                          alphas.data(), effOut.data(), &iters, &mrd);
        lastIters_ = iters;
        if (rc == SFB200_ENOACTIVE || rc == SFB200_ESMALLSUM) { lastError_ = sfb200_last_error(dev_.get()); return false; }
        dev_.check(rc);
        double alphaSum = 0.0;
        for (double a : alphas) alphaSum += a;
        for (size_t i = 0; i < transcripts.size(); ++i) {                       // :883-890
            transcripts[i].EffectiveLength = effOut[i];
            transcripts[i].setEstCount(alphas[i]);
            transcripts[i].setMass(alphas[i] / alphaSum);
        }
        return true;
    }

    // bool gatherBootstraps(ReadExperiment&, SailfishOpts&, std::function<bool(const std::vector<double>&)>&, double, uint32_t)
    template <typename ExpT, typename OptsT>
    bool gatherBootstraps(ExpT& readExp, OptsT& sopt, std::function<bool(const std::vector<double>&)>& writeBootstrap,
                          double relDiffTolerance = 0.01, uint32_t maxIter = 10000, uint64_t seed = 0x5f3759dfULL) {
        const std::vector<double> eff = detail::effLens(readExp, sopt);
        const sfb200_em_opts o = detail::emOpts(sopt.useVBOpt, relDiffTolerance, maxIter);
        struct Ctx { std::function<bool(const std::vector<double>&)>* f; std::vector<double> row; } cx{&writeBootstrap, {}};
        auto tramp = [](void* u, const double* row, size_t n) -> int {
            Ctx* c = static_cast<Ctx*>(u);
            c->row.assign(row, row + n);
            return (*c->f)(c->row) ? 0 : 1;
        };
        const int rc = sfb200_bootstrap_run(dev_.get(), eff.data(), static_cast<uint32_t>(eff.size()), &o, sopt.numBootstraps, seed,
                                            tramp, &cx);
        if (rc == SFB200_ENOACTIVE || rc == SFB200_ESMALLSUM || rc == SFB200_ECALLBACK) { lastError_ = sfb200_last_error(dev_.get()); return false; }
        dev_.check(rc);
        return true;
    }
    uint32_t lastIterations() const { return lastIters_; }
    const std::string& lastError() const { return lastError_; }

private:
    Device& dev_;
    uint32_t lastIters_ = 0;
    std::string lastError_;
};

class CollapsedGibbsSampler {
public:
    explicit CollapsedGibbsSampler(Device& d) : dev_(d) {}
    // template <typename ExpT> bool sample(ExpT&, SailfishOpts&, std::function<bool(const std::vector<int>&)>&, uint32_t)
    template <typename ExpT, typename OptsT>
    bool sample(ExpT& readExp, OptsT& sopt, std::function<bool(const std::vector<int>&)>& writeSample, uint32_t numSamples = 500,
                uint64_t seed = 0x2545F491ULL) {
        auto& transcripts = readExp.transcripts();
        const std::vector<double> eff = detail::effLens(readExp, sopt);
        std::vector<double> masses(transcripts.size());
        const double numMapped = static_cast<double>(readExp.numMappedFragments());
        for (size_t i = 0; i < transcripts.size(); ++i) masses[i] = transcripts[i].mass();
        struct Ctx { std::function<bool(const std::vector<int>&)>* f; std::vector<int> row; } cx{&writeSample, {}};
        auto tramp = [](void* u, const int32_t* row, size_t n) -> int {
            Ctx* c = static_cast<Ctx*>(u);
            c->row.assign(row, row + n);
            return (*c->f)(c->row) ? 0 : 1;
        };
        dev_.check(sfb200_gibbs_run(dev_.get(), eff.data(), masses.data(), static_cast<uint32_t>(eff.size()),
                                    readExp.numMappedFragments(), numSamples, seed, tramp, &cx));
        // the reference overwrites Transcript::mass_ with prior + mass * numMapped and never restores it
        // (CollapsedGibbsSampler.cpp:219-221): keep that observable side effect
        for (size_t i = 0; i < transcripts.size(); ++i) transcripts[i].setMass(1e-8 + masses[i] * numMapped);
        return true;
    }

private:
    Device& dev_;
};

}  // namespace sfb200
#endif
