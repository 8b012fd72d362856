// TEST INFRASTRUCTURE ONLY.  C API over pieces of the REAL reference, compiled from the sources where
// they lie under /root/reference (never copied into this repo).  Built by oracle/Makefile into
// oracle/_ref/libsfref.so and used by tests/ to pin the oracle restatement.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "xxhash.h"                      // /root/reference/include/xxhash.h
#include "TranscriptGroup.hpp"           // /root/reference/include/TranscriptGroup.hpp
#include "EquivalenceClassBuilder.hpp"   // /root/reference/include/EquivalenceClassBuilder.hpp
#include "LibraryFormat.hpp"             // /root/reference/include/LibraryFormat.hpp
#include "EmpiricalDistribution.hpp"     // /root/reference/include/EmpiricalDistribution.hpp
#include "spdlog/sinks/null_sink.h"

extern "C" {

uint64_t ref_xxh64(const void* p, size_t len, uint64_t seed) { return XXH64(p, len, seed); }

uint64_t ref_tgroup_hash(const uint32_t* ids, uint32_t n) {
    TranscriptGroup tg(std::vector<uint32_t>(ids, ids + n));
    return tg.hash;
}

struct ref_eqb {
    std::shared_ptr<spdlog::logger> log;
    std::unique_ptr<EquivalenceClassBuilder> b;
};

ref_eqb* ref_eqb_create() {
    auto* r = new ref_eqb();
    auto sink = std::make_shared<spdlog::sinks::null_sink_st>();
    r->log = std::make_shared<spdlog::logger>("ref", sink);
    r->b.reset(new EquivalenceClassBuilder(r->log));
    r->b->start();
    return r;
}
void ref_eqb_free(ref_eqb* r) { delete r; }
// the call processReadsQuasi makes: TranscriptGroup tg(ids); eqBuilder.addGroup(std::move(tg), auxProbs)  (SailfishQuantify.cpp:402-403)
void ref_eqb_add(ref_eqb* r, const uint32_t* ids, uint32_t n) {
    std::vector<uint32_t> v(ids, ids + n);
    std::vector<double> aux(n, 1.0);
    TranscriptGroup tg(v);
    r->b->addGroup(std::move(tg), aux);
}
uint64_t ref_eqb_finish(ref_eqb* r, uint64_t* nnz) {
    r->b->finish();
    uint64_t z = 0;
    for (auto& kv : r->b->eqVec()) z += kv.first.txps.size();
    if (nnz) *nnz = z;
    return r->b->eqVec().size();
}
// eqVec() in the reference's own (bucket-major) order
void ref_eqb_export(ref_eqb* r, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts, double* weights) {
    uint64_t e = 0, z = 0;
    row_ptr[0] = 0;
    for (auto& kv : r->b->eqVec()) {
        for (size_t i = 0; i < kv.first.txps.size(); ++i) {
            labels[z] = kv.first.txps[i];
            if (weights) weights[z] = kv.second.weights[i];
            ++z;
        }
        counts[e] = kv.second.count.load();
        row_ptr[++e] = z;
    }
}

int ref_format_id(int type, int orientation, int strandedness) {
    LibraryFormat f(static_cast<ReadType>(type), static_cast<ReadOrientation>(orientation), static_cast<ReadStrandedness>(strandedness));
    return f.formatID();
}
int ref_format_roundtrip(int id) { return LibraryFormat::formatFromID(static_cast<uint8_t>(id)).formatID(); }
int ref_format_check(int id) { return LibraryFormat::formatFromID(static_cast<uint8_t>(id)).check() ? 1 : 0; }

// EmpiricalDistribution over (vals, lens): writes pdf[0..n_pdf), returns median; min/max by pointer
float ref_empdist(const uint32_t* vals, const uint32_t* lens, uint32_t n, float* pdf, uint32_t n_pdf, uint32_t* mn, uint32_t* mx) {
    std::vector<uint32_t> v(vals, vals + n), l(lens, lens + n);
    EmpiricalDistribution d(v, l);
    for (uint32_t i = 0; i < n_pdf; ++i) pdf[i] = d.pdf(i);
    *mn = d.minValue(); *mx = d.maxValue();
    return d.median();
}

}  // extern "C"
