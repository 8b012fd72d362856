// common.cuh -- shared host/device plumbing of libsfb200 (context, error handling, device helpers).
// Product code: nothing here may reference oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/sfb200.h"

#define SFB_CUDA(ctx, call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                 \
            return SFB200_ECUDA;                                                                              \
        }                                                                                                     \
    } while (0)

#define SFB_FAIL(ctx, code, msg)                                                                              \
    do {                                                                                                      \
        (ctx)->err = (msg);                                                                                   \
        return (code);                                                                                        \
    } while (0)

// ---- device buffer that remembers its capacity (grow-only) ----------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        // 16 elements of slack: bulk (TMA) copies of 16-byte-aligned supersets may read past the logical end
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), (n + 16) * sizeof(T));
        if (e == cudaSuccess) cap = n ? n : 1;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    size_t bytes() const { return cap * sizeof(T); }
};

// A function's own scratch buffers: released when the function returns, on the error paths of SFB_CUDA / SFB_FAIL as well.
// (DevBuf itself has no destructor: the long-lived ones are members that are copied and swapped when tables grow.)
template <typename... Bufs>
struct DevBufScope {
    std::tuple<Bufs&...> bufs;
    explicit DevBufScope(Bufs&... b) : bufs(b...) {}
    ~DevBufScope() { std::apply([](auto&... b) { (b.release(), ...); }, bufs); }
    DevBufScope(const DevBufScope&) = delete;
    DevBufScope& operator=(const DevBufScope&) = delete;
};

// ---- the device index (mapping spec v1) ---------------------------------------------------
struct DevIndex {
    int k = 0;
    uint32_t n_txp = 0;
    uint64_t text_len = 0;       // packed coordinate space (transcripts concatenated, no separators)
    uint64_t n_sa = 0;           // valid k-mer start positions
    uint64_t n_kmers = 0;        // distinct k-mers
    uint64_t table_slots = 0;    // power of two
    uint32_t max_bucket = 0;
    DevBuf<uint64_t> words;      // 2-bit text, 32 bases per word, base p at bits 2*(p%32)
    DevBuf<uint64_t> txp_start;  // n_txp+1
    DevBuf<uint32_t> txp_len;    // n_txp
    DevBuf<uint64_t> txp_end;    // n_txp: txp_start[t] + txp_len[t]
    // n_sa entries sorted by (k-mer value, position), 16 bytes each: {transcript, position inside it, text position, bases left to the
    // transcript's end}.  Positions inside / bases left are redundant with txp_start / txp_end, but having them in the entry removes a
    // dependent random load from every match extension and every projected hit, and each kernel reads only its 8-byte half
    DevBuf<uint4> sa;
    DevBuf<uint4> table;         // {key lo, key hi, lb, cnt}; empty = cnt 0
    // presence filter: bitmap over the m-mers of the text (kmer_filter.hpp); one bit decides k-m+1 read positions
    DevBuf<uint32_t> mfilter;
    int mf_m = 0;
    bool ready = false;
    size_t hbm_bytes() const {
        return words.bytes() + txp_start.bytes() + txp_len.bytes() + txp_end.bytes() + sa.bytes() + table.bytes() + mfilter.bytes();
    }
};

// ---- equivalence classes prepared for the inference kernels ----------------------------------------------------
// Classes with >= 2 members are stored "binned": bin b holds the classes whose member count n satisfies
// g/2 < n <= g for g = 2,4,8,16,32 (b = 0..4) and n > 32 (b = 5), so a sub-warp group of g lanes owns one class.
// Classes with one member never enter the sweep: their counts form the per-transcript vector `single`.
constexpr int SFB_NBINS = 6;
// CTA-partitioned copy of the multi-member classes (em_part.cuh): local classes grouped by (CTA, bin), then the pool by bin
struct DevPartition {
    bool valid = false;      // built for the current classes
    bool usable = false;     // every CTA's slice fits in shared memory
    uint32_t n_cta = 0;
    uint64_t pool_cls[SFB_NBINS + 1] = {0, 0, 0, 0, 0, 0, 0};
    uint64_t n_pool = 0, pool_nnz = 0;
    uint32_t n_pool_cta = 0;            // hybrid runs (em_dense.cuh): CTAs that run the pool loop, launched after the n_cta component CTAs
    uint32_t n_dirty = 0;               // pool transcripts
    DevBuf<uint32_t> dlist;             // hybrid runs: the pool's transcripts
    uint64_t pool_nz = 0;               // label entries of the pool classes
    uint64_t max_cta_bytes = 0, max_cta_bytes_vb = 0, smem_limit = 0;
    int per_sm = 1;
    DevBuf<uint32_t> start, len, lab, src, bounds, owner, load;
    DevBuf<double> cnt, w, cnt_s;
    DevBuf<unsigned long long> tbl, grp, pre;
    DevBuf<uint8_t> dirty;
    // gather layout of the atomic-free loop (em_gather.cuh): one region per CTA, geometry kept as opaque words
    bool gather_ok = false, gather_tried = false;
    uint32_t gth_geom[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t gather_smem = 0;           // dynamic shared memory the largest CTA needs
    DevBuf<uint32_t> gth;
    // dense-component layout (em_dense.cuh)
    bool dense_ok = false, dense_tried = false;
    uint32_t dns_geom[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t dense_smem = 0;
    uint64_t dense_smem_lag = 0;        // with the alpha ring of the lagged stopping rule (0: does not fit)
    uint32_t dense_ns = 0;
    bool dense_stream = false;          // counts / base / 1/effLen streamed from a per-CTA global block (em_dense.cuh)
    uint32_t stream_ent = 0, stream_state = 0;
    DevBuf<double> dns_f64;
    DevBuf<uint32_t> dns;
    void release() { start.release(); len.release(); lab.release(); src.release(); bounds.release(); owner.release(); load.release();
                     cnt.release(); w.release(); cnt_s.release(); tbl.release(); grp.release(); pre.release(); dirty.release(); gth.release(); dns.release(); dlist.release(); dns_f64.release(); }
};
struct DevClasses {
    DevPartition part;
    uint32_t n_txp = 0;
    uint64_t E = 0, nnz = 0;            // as imported (all classes)
    uint64_t Em = 0, nnzm = 0;          // multi-member classes only
    uint64_t bin_cls[SFB_NBINS + 1] = {0, 0, 0, 0, 0, 0, 0};   // class index range of each bin
    uint64_t n_active = 0;
    uint64_t total_count = 0;
    DevBuf<uint32_t> start;             // Em: first entry of the class in lab/w
    DevBuf<uint32_t> len;               // Em: member count
    DevBuf<uint32_t> lab;               // nnzm transcript ids
    DevBuf<double>   w;                 // nnzm weights (computed per run from eff_lens)
    DevBuf<double>   cnt;               // Em counts as f64 (exact below 2^53)
    // "canonical index" of a class = its position in the order bootstrap count vectors are indexed by:
    // import path: the caller's order; device-finish path: multi-member classes in binned order, then the singles.
    DevBuf<uint32_t> perm;              // Em: binned position -> canonical class index
    DevBuf<uint32_t> sgl_cls, sgl_tid;  // single-member classes: canonical index and transcript
    uint64_t n_sgl = 0;
    DevBuf<unsigned long long> cnt_all; // E counts in canonical order (device-finish path)
    DevBuf<double>   single;            // n_txp: count of the class {t}
    DevBuf<uint8_t>  active;            // n_txp
    // host copy in EXPORT order (label-lexicographic for the device-finish path, caller's order for eq_import);
    // materialised lazily by sfb_classes_host() because only eq_export / bootstrap need it
    bool host_valid = false;
    bool from_device = false;
    bool merged = false;                // multi-rank: the classes of all ranks were merged (every rank holds the global set)
    std::vector<uint64_t> h_row_ptr;
    std::vector<uint32_t> h_labels;
    std::vector<uint64_t> h_counts;
    std::vector<uint64_t> export_to_canon;   // empty = identity
    bool ready = false;
};

struct MapState;   // map.cu

struct sfb200_ctx {
    int device = 0;
    int num_sms = 0;
    int coop = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    double last_em_ms = 0.0;
    int last_em_kernel = 0;             // see sfb200_last_em_kernel
    int last_em_variant = 0;            // see sfb200_last_em_variant
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevIndex index;
    DevClasses cls;
    MapState* map = nullptr;
    void* em_extra = nullptr;           // em.cu: per-sample buffers
    // communicator (NCCL loaded at run time, see comm.cu)
    void* comm = nullptr;
    int n_ranks = 1, rank = 0;
    // scratch
    DevBuf<double> em_alpha;            // 3 * n_txp (rotating alpha buffers)
    DevBuf<double> em_theta;            // n_txp (VBEM expTheta)
    DevBuf<double> em_base;             // n_txp (initial value of every output buffer: single counts (+ prior))
    DevBuf<unsigned long long> em_ctl;  // control block of the persistent kernel
    DevBuf<double> eff;                 // n_txp clamped effective lengths
    bool eff_resident = false;          // set by a bootstrap run after its first replicate: c->eff already holds this run's lengths
    void* fastq = nullptr;              // fastq.cu: staging buffers of the device-side FASTQ extraction
};

int sfb_comm_allreduce_f64(sfb200_ctx* ctx, double* d_buf, size_t n);
int sfb_comm_allreduce_u64(sfb200_ctx* ctx, unsigned long long* d_buf, size_t n);
int sfb_comm_allgather(sfb200_ctx* ctx, const void* send, void* recv, size_t bytes);
void sfb_map_state_free(sfb200_ctx* ctx);
void sfb_em_extra_free(sfb200_ctx* ctx);
void sfb_fastq_free(sfb200_ctx* ctx);
int sfb_classes_from_host(sfb200_ctx* ctx, uint32_t n_txp, uint64_t E, const uint64_t* row_ptr, const uint32_t* labels,
                          const uint64_t* counts);
int sfb_classes_host(sfb200_ctx* ctx);   // make cls.h_* valid (downloads + sorts after a device-side finish)

// ---- device helpers ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// XXH64 (reference: src/xxhash.c:346-455) -- written from the published algorithm description:
// 4 accumulator lanes over 32-byte stripes, merge, then 8/4/1-byte tail, then avalanche.
constexpr uint64_t XXP1 = 11400714785074694791ULL;
constexpr uint64_t XXP2 = 14029467366897019727ULL;
constexpr uint64_t XXP3 = 1609587929392839161ULL;
constexpr uint64_t XXP4 = 9650029242287828579ULL;
constexpr uint64_t XXP5 = 2870177450012600261ULL;

__host__ __device__ __forceinline__ uint64_t xx_round(uint64_t acc, uint64_t in) {
    acc += in * XXP2;
    acc = (acc << 31) | (acc >> 33);
    return acc * XXP1;
}
__host__ __device__ __forceinline__ uint64_t xx_merge(uint64_t h, uint64_t v) {
    h ^= xx_round(0, v);
    return h * XXP1 + XXP4;
}
__host__ __device__ __forceinline__ uint64_t xx_avalanche(uint64_t h) {
    h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
    return h;
}
// XXH64 of one 8-byte word (the k-mer key hasher)
__host__ __device__ __forceinline__ uint64_t xxh64_u64(uint64_t w, uint64_t seed) {
    uint64_t h = seed + XXP5 + 8;
    h ^= xx_round(0, w);
    h = ((h << 27) | (h >> 37)) * XXP1 + XXP4;
    return xx_avalanche(h);
}
// XXH64 of n 32-bit words (a label); `get(i)` returns word i
template <typename F>
__host__ __device__ __forceinline__ uint64_t xxh64_words(F get, uint32_t n, uint64_t seed) {
    const uint64_t len = 4ull * n;
    uint64_t h;
    uint32_t i = 0;
    if (n >= 8) {
        uint64_t v1 = seed + XXP1 + XXP2, v2 = seed + XXP2, v3 = seed, v4 = seed - XXP1;
        for (; i + 8 <= n; i += 8) {
            v1 = xx_round(v1, (uint64_t)get(i) | ((uint64_t)get(i + 1) << 32));
            v2 = xx_round(v2, (uint64_t)get(i + 2) | ((uint64_t)get(i + 3) << 32));
            v3 = xx_round(v3, (uint64_t)get(i + 4) | ((uint64_t)get(i + 5) << 32));
            v4 = xx_round(v4, (uint64_t)get(i + 6) | ((uint64_t)get(i + 7) << 32));
        }
        h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
        h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
    } else {
        h = seed + XXP5;
    }
    h += len;
    for (; i + 2 <= n; i += 2) {
        const uint64_t k1 = xx_round(0, (uint64_t)get(i) | ((uint64_t)get(i + 1) << 32));
        h ^= k1;
        h = ((h << 27) | (h >> 37)) * XXP1 + XXP4;
    }
    if (i < n) {
        h ^= (uint64_t)get(i) * XXP1;
        h = ((h << 23) | (h >> 41)) * XXP2 + XXP3;
    }
    return xx_avalanche(h);
}

__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }

#include "kmer_filter.hpp"     // k-mer table hash + the presence filter's addressing (also compiled by the CPU tests)

#include "vb_math.hpp"         // sfb_digamma, sfb_exp_digamma (also compiled by the CPU tests)
#endif
