// infer.cu -- posterior sampling on top of the EM machinery: bootstrap (CollapsedEMOptimizer::gatherBootstraps,
// reference src/CollapsedEMOptimizer.cpp:557-709 + doBootstrap :438-525 + include/MultinomialSampler.hpp:13-64) and the
// collapsed Gibbs sampler (src/CollapsedGibbsSampler.cpp:35-186,199-291).
//
// Random numbers: the reference seeds std::mt19937 from std::random_device, so its draws are not reproducible and
// parity is distributional.  Here every draw is a counter-based Philox4x32-10 value keyed by (seed, sample index),
// which makes runs reproducible and lets every draw be made by an independent thread.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

int sfb_bootstrap_em_device(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, const unsigned long long* d_samp,
                            uint64_t total, const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out);

namespace {

constexpr double DENORM_MIN = 4.9406564584124654e-324;

// ---- Philox4x32-10 (Salmon et al., SC'11), written from the published round function -------------------------------------
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ void operator()(uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) const {
        uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
__host__ __device__ inline double u01_from(uint32_t hi, uint32_t lo) {       // [0,1) with 53 random bits
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

// MultinomialSampler::operator() for k > 100 (:48-63): lower_bound over z[0..k), step back if z[it] > u.
// z has k+1 entries, z[0] = 0, z[i] = p_0 + ... + p_{i-1} accumulated left to right as the reference's table is.
__device__ __forceinline__ uint32_t pick_class(const double* __restrict__ z, uint32_t k, double u) {
    uint32_t lo = 0, hi = k;                           // search [0, k): first i with z[i] >= u, k if none
    while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (__ldg(z + mid) < u) lo = mid + 1; else hi = mid; }
    uint32_t off = lo;
    if (__ldg(z + off) > u && off > 0) off -= 1;
    return off < k ? off : k - 1;
}

// one draw per thread (two per Philox call); N = total fragments, k = number of classes
__global__ void k_multinomial_draws(const double* __restrict__ z, uint32_t k, uint64_t n_draws, uint64_t seed,
                                    unsigned long long* __restrict__ samp) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t d0 = 2 * i;
    if (d0 >= n_draws) return;
    uint32_t r[4];
    Philox ph(seed);
    ph(i, 0x6d756c74ULL /* "mult" */, r);
    atomicAdd(samp + pick_class(z, k, u01_from(r[0], r[1])), 1ULL);
    if (d0 + 1 < n_draws) atomicAdd(samp + pick_class(z, k, u01_from(r[2], r[3])), 1ULL);
}

inline unsigned gridn(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int sfb200_bootstrap_run(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, const sfb200_em_opts* opts,
                                    uint32_t n_boot, uint64_t seed, sfb200_f64_row_cb cb, void* user) {
    if (!c || !eff_lens || !opts) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "bootstrap_run: no classes");
    if (n_txp != c->cls.n_txp) SFB_FAIL(c, SFB200_EINVAL, "bootstrap_run: n_txp differs from the class table's");
    cudaSetDevice(c->device);
    const DevClasses& k = c->cls;
    const uint64_t E = k.E;
    if (E == 0 || k.n_active == 0) SFB_FAIL(c, SFB200_ENOACTIVE, "The optimizer has no active transcripts: no transcripts are expressed");
    if (E >= 0xFFFFFFFFull) SFB_FAIL(c, SFB200_EINVAL, "too many classes");
    // markDegenerateClasses (:372-433) never drops a class here: with uniform positive alphas over the active set every
    // denominator is positive (see oracle note); all classes are valid.
    const uint64_t totalCount = k.total_count;                                     // :662-674
    const double floatCount = static_cast<double>(totalCount);
    // class counts in the order the per-sample count vectors are indexed by (canonical order)
    std::vector<uint64_t> canon_counts(E);
    if (k.from_device) {
        SFB_CUDA(c, cudaMemcpy(canon_counts.data(), k.cnt_all.p, E * 8, cudaMemcpyDeviceToHost));
    } else {
        canon_counts = k.h_counts;
    }
    std::vector<double> z(E + 1);
    double sum = 0.0;
    z[0] = 0.0;
    for (uint64_t e = 0; e < E; ++e) { sum += static_cast<double>(canon_counts[e]) / floatCount; z[e + 1] = sum; }   // :676-680, MultinomialSampler.hpp:30-34
    cudaStream_t s = c->stream;
    DevBuf<double> d_z; DevBuf<unsigned long long> d_samp;
    SFB_CUDA(c, d_z.reserve(E + 1)); SFB_CUDA(c, d_samp.reserve(E));
    SFB_CUDA(c, cudaMemcpyAsync(d_z.p, z.data(), (E + 1) * 8, cudaMemcpyHostToDevice, s));
    std::vector<double> alphas(n_txp);
    // MultinomialSampler takes n as uint32_t (:15): the reference wraps above 2^32 fragments; we keep 64 bits
    int rc = SFB200_OK;
    double loop_ms = 0.0;
    for (uint32_t b = 0; b < n_boot && rc == SFB200_OK; ++b) {
        cudaMemsetAsync(d_samp.p, 0, E * 8, s);
        if (totalCount) {
            k_multinomial_draws<<<gridn((totalCount + 1) / 2, 256), 256, 0, s>>>(d_z.p, (uint32_t)E, totalCount, seed + 0x9E3779B97F4A7C15ULL * (b + 1), d_samp.p);
            c->launches++;
        }
        uint32_t iters = 0;
        rc = sfb_bootstrap_em_device(c, eff_lens, n_txp, d_samp.p, totalCount, opts, alphas.data(), &iters);
        loop_ms += c->last_em_ms;
        if (rc == SFB200_OK && cb && cb(user, alphas.data(), n_txp) != 0) { c->err = "bootstrap row callback failed"; rc = SFB200_ECALLBACK; }
    }
    c->last_em_ms = loop_ms;
    d_z.release(); d_samp.release();
    return rc;
}

extern "C" int sfb200_gibbs_run(sfb200_ctx* c, const double* eff_lens, const double* masses, uint32_t n_txp, uint64_t num_mapped,
                                uint32_t n_samples, uint64_t seed, sfb200_i32_row_cb cb, void* user) {
    if (!c) return SFB200_EINVAL;
    SFB_FAIL(c, SFB200_EINVAL, "gibbs_run: not built yet");
}
