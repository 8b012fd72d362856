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
#include <ctime>
#include <cstdlib>

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

inline unsigned gridn(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// ======================================================================================================================
// Collapsed Gibbs sampler (reference src/CollapsedGibbsSampler.cpp)
// ======================================================================================================================
// sampleRound_ (:96-186) walks the classes one after another because consecutive classes may share transcripts (the
// conditional of a class depends on txpCount of its members).  Classes that share NO transcript are conditionally
// independent, so the walk is re-ordered by a greedy colouring (classes of one colour are pairwise disjoint): colours are
// visited in sequence, the classes of a colour in parallel -- the same Markov kernel as a sequential sweep in colour order,
// which is a valid scan order of the same sampler (the reference's own order is libcuckoo's arbitrary bucket order).
namespace {

struct Rng {                       // counter-based stream: (seed; sample, purpose, class) -> as many uniforms as asked for
    Philox ph; uint64_t lo, hi; uint32_t buf[4]; int left;
    __device__ Rng(uint64_t seed, uint64_t sample, uint64_t purpose, uint64_t item) : ph(seed), lo(item << 16), hi((sample << 8) | purpose), left(0) {}
    __device__ double u01() {
        if (left < 2) { ph(lo++, hi, buf); left = 4; }
        left -= 2;
        return u01_from(buf[left], buf[left + 1]);
    }
};

__device__ __forceinline__ double stirling_tail(double k) {
    const double t[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509, 0.0166446911898211,
                          0.0138761288230707, 0.0118967099458917, 0.0104112652619720, 0.00925546218271273, 0.00833056343336287};
    if (k <= 9.0) return t[(int)k];
    const double kp1sq = (k + 1) * (k + 1);
    return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / (k + 1);
}

// exact Binomial(n, p) draw: sequential inversion for small n*p, Hormann's BTRS transformed rejection otherwise
__device__ uint64_t binomial_draw(Rng& g, uint64_t n, double p) {
    if (n == 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    if (flip) p = 1.0 - p;
    const double dn = (double)n, q = 1.0 - p;
    uint64_t x;
    if (dn * p < 10.0) {
        // waiting-time inversion: sum of geometric gaps
        double qn = exp(dn * log1p(-p));
        const double bound = fmin(dn, dn * p + 10.0 * sqrt(dn * p * q + 1.0));
        double px = qn, U = g.u01();
        double X = 0.0;
        while (U > px) {
            X += 1.0;
            if (X > bound) { X = 0.0; px = qn; U = g.u01(); }
            else { U -= px; px = ((dn - X + 1.0) * p * px) / (X * q); }
        }
        x = (uint64_t)X;
    } else {
        const double stddev = sqrt(dn * p * q);
        const double b = 1.15 + 2.53 * stddev, a = -0.0873 + 0.0248 * b + 0.01 * p, cc = dn * p + 0.5;
        const double v_r = 0.92 - 4.2 / b, r = p / q, alpha = (2.83 + 5.1 / b) * stddev, m = floor((dn + 1.0) * p);
        for (;;) {
            const double u = g.u01() - 0.5;
            double v = g.u01();
            const double us = 0.5 - fabs(u);
            const double k = floor((2.0 * a / us + b) * u + cc);
            if (us >= 0.07 && v <= v_r) { x = (uint64_t)k; break; }
            if (k < 0.0 || k > dn) continue;
            v = log(v * alpha / (a / (us * us) + b));
            const double ub = (m + 0.5) * log((m + 1.0) / (r * (dn - m + 1.0))) + (dn + 1.0) * log((dn - m + 1.0) / (dn - k + 1.0)) +
                              (k + 0.5) * log(r * (dn - k + 1.0) / (k + 1.0)) + stirling_tail(m) + stirling_tail(dn - m) -
                              stirling_tail(k) - stirling_tail(dn - k);
            if (v <= ub) { x = (uint64_t)k; break; }
        }
    }
    return flip ? n - x : x;
}

// Multinomial(n; probs[0..k)) by conditional binomials, written into out[0..k) (MultinomialSampler.hpp:13-64 draws the same
// distribution one uniform at a time)
__device__ void multinomial_draw(Rng& g, uint64_t n, uint32_t k, const double* probs, unsigned long long* out) {
    double rem_p = 0.0;
    for (uint32_t i = 0; i < k; ++i) rem_p += probs[i];
    uint64_t rem = n;
    for (uint32_t i = 0; i < k; ++i) {
        uint64_t x = 0;
        if (rem > 0) {
            if (i + 1 == k || !(rem_p > probs[i])) x = rem;
            else x = binomial_draw(g, rem, fmin(1.0, fmax(0.0, probs[i] / rem_p)));
        }
        out[i] = x;
        rem -= x;
        rem_p -= probs[i];
    }
}

struct GibbsParams {
    const unsigned long long* row_ptr; const uint32_t* labels; const unsigned long long* counts; const double* w;
    unsigned long long* countMap; double* probMap; int* txpCount;
    const double* mass;            // prior + mass * numMapped (:219-221)
    double prior;
    uint64_t seed;
};

// initCountMap_ (:35-94): every class splits its count over its members, independently of the other classes
__global__ void k_gibbs_init(const GibbsParams p, uint64_t E) {
    const uint64_t eq = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (eq >= E) return;
    const uint64_t b = p.row_ptr[eq], e = p.row_ptr[eq + 1];
    const uint32_t k = (uint32_t)(e - b);
    const uint64_t classCount = p.counts[eq];
    if (k > 1) {
        double denom = 0.0;
        for (uint64_t j = b; j < e; ++j) { denom += (p.prior + p.mass[p.labels[j]]) * p.w[j]; p.countMap[j] = 0; }
        if (denom > DENORM_MIN) {
            const double norm = 1.0 / denom;
            for (uint64_t j = b; j < e; ++j) p.probMap[j] = norm * ((p.prior + p.mass[p.labels[j]]) * p.w[j]);
            Rng g(p.seed, 0, 1, eq);
            multinomial_draw(g, classCount, k, p.probMap + b, p.countMap + b);
        }
    } else if (k == 1) {
        p.countMap[b] = classCount;
    }
    for (uint64_t j = b; j < e; ++j) atomicAdd(p.txpCount + p.labels[j], (int)p.countMap[j]);
}

// sampleRound_ (:96-186) for the classes of one colour
__global__ void k_gibbs_round(const GibbsParams p, const uint32_t* __restrict__ order, uint64_t lo, uint64_t hi, uint64_t sample) {
    const uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint64_t eq = order[i];
    const uint64_t b = p.row_ptr[eq], e = p.row_ptr[eq + 1];
    const uint32_t k = (uint32_t)(e - b);
    Rng g(p.seed, sample + 1, 2, eq);
    const double sampleFrac = 0.25 + 0.5 * g.u01();                      // uniform_real_distribution(0.25, 0.75) (:106,115)
    double denom = 0.0;
    unsigned long long numResampled = 0;
    for (uint64_t j = b; j < e; ++j) {                                      // :124-136
        const uint32_t tid = p.labels[j];
        const unsigned long long curr = p.countMap[j];
        const unsigned long long res = (unsigned long long)round(sampleFrac * (double)curr);
        numResampled += res;
        p.txpCount[tid] -= (int)res;
        p.countMap[j] = curr - res;
        p.probMap[j] = (double)res;                                       // parked: restored below if the class is skipped
        denom += (p.prior + (double)p.txpCount[tid]) * p.w[j];
    }
    if (denom > DENORM_MIN) {                                               // :138-160
        // the resample sizes are needed again only if the class is skipped; it is not, so probMap can take the probabilities
        const double norm = 1.0 / denom;
        for (uint64_t j = b; j < e; ++j) p.probMap[j] = norm * ((p.prior + (double)p.txpCount[p.labels[j]]) * p.w[j]);
        // draw into a scratch row that aliases nothing: reuse countMap increments through a second pass
        Rng g2(p.seed, sample + 1, 3, eq);
        double rem_p = 0.0;
        for (uint64_t j = b; j < e; ++j) rem_p += p.probMap[j];
        unsigned long long rem = numResampled;
        for (uint64_t j = b; j < e; ++j) {
            unsigned long long x = 0;
            if (rem > 0) {
                if (j + 1 == e || !(rem_p > p.probMap[j])) x = rem;
                else x = binomial_draw(g2, rem, fmin(1.0, fmax(0.0, p.probMap[j] / rem_p)));
            }
            rem -= x; rem_p -= p.probMap[j];
            p.countMap[j] += x;                                            // :162-176
            p.txpCount[p.labels[j]] += (int)x;
        }
    } else {
        for (uint64_t j = b; j < e; ++j) {                                  // class skipped: put the removed counts back
            const unsigned long long res = (unsigned long long)p.probMap[j];
            p.countMap[j] += res;
            p.txpCount[p.labels[j]] += (int)res;
        }
    }
    (void)k;
}

__global__ void k_entry_weights(const unsigned long long* __restrict__ row_ptr, const uint32_t* __restrict__ labels,
                                const unsigned long long* __restrict__ counts, const double* __restrict__ eff_in, uint64_t E,
                                double* __restrict__ w) {
    const uint64_t eq = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (eq >= E) return;
    const double count = (double)counts[eq];
    double wsum = 0.0;
    for (uint64_t j = row_ptr[eq]; j < row_ptr[eq + 1]; ++j) {              // CollapsedEMOptimizer.cpp:745-772
        const double e = eff_in[labels[j]];
        const double v = count / (e <= 1.0 ? 1.0 : e);
        w[j] = v; wsum += v;
    }
    const double wnorm = 1.0 / wsum;
    for (uint64_t j = row_ptr[eq]; j < row_ptr[eq + 1]; ++j) w[j] *= wnorm;
}

}  // namespace

extern "C" int sfb200_gibbs_run(sfb200_ctx* c, const double* eff_lens, const double* masses, uint32_t n_txp, uint64_t num_mapped,
                                uint32_t n_samples, uint64_t seed, sfb200_i32_row_cb cb, void* user) {
    if (!c || !eff_lens || !masses) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "gibbs_run: no classes");
    if (n_txp != c->cls.n_txp) SFB_FAIL(c, SFB200_EINVAL, "gibbs_run: n_txp differs from the class table's");
    cudaSetDevice(c->device);
    { const int rc = sfb_classes_host(c); if (rc) return rc; }
    const DevClasses& k = c->cls;
    const uint64_t E = k.E, nnz = k.nnz;
    cudaStream_t s = c->stream;
    // greedy colouring: colour(class) = max over its members of the next free colour of that transcript
    std::vector<uint32_t> next_free(n_txp, 0), colour(E, 0);
    uint32_t n_colours = 0;
    for (uint64_t e = 0; e < E; ++e) {
        const uint64_t b = k.h_row_ptr[e], en = k.h_row_ptr[e + 1];
        if (en - b <= 1) { colour[e] = 0xFFFFFFFFu; continue; }            // single-member classes are never resampled (:120)
        uint32_t col = 0;
        for (uint64_t j = b; j < en; ++j) col = std::max(col, next_free[k.h_labels[j]]);
        for (uint64_t j = b; j < en; ++j) next_free[k.h_labels[j]] = col + 1;
        colour[e] = col;
        n_colours = std::max(n_colours, col + 1);
    }
    std::vector<uint64_t> col_start(n_colours + 1, 0);
    for (uint64_t e = 0; e < E; ++e) if (colour[e] != 0xFFFFFFFFu) col_start[colour[e] + 1]++;
    for (uint32_t q = 0; q < n_colours; ++q) col_start[q + 1] += col_start[q];
    std::vector<uint32_t> order(col_start[n_colours] ? col_start[n_colours] : 1);
    { std::vector<uint64_t> cur(col_start.begin(), col_start.end() - 1);
      for (uint64_t e = 0; e < E; ++e) if (colour[e] != 0xFFFFFFFFu) order[cur[colour[e]]++] = (uint32_t)e; }

    DevBuf<unsigned long long> d_rp, d_cnt, d_cmap; DevBuf<uint32_t> d_lab, d_order; DevBuf<double> d_w, d_prob, d_mass, d_eff; DevBuf<int> d_txp;
    auto cleanup = [&]() { d_rp.release(); d_cnt.release(); d_cmap.release(); d_lab.release(); d_order.release(); d_w.release();
                           d_prob.release(); d_mass.release(); d_eff.release(); d_txp.release(); };
#define GB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); c->err = std::string(#call) + ": " + cudaGetErrorString(e__); return SFB200_ECUDA; } } while (0)
    GB_CUDA(d_rp.reserve(E + 1)); GB_CUDA(d_cnt.reserve(E)); GB_CUDA(d_cmap.reserve(nnz)); GB_CUDA(d_lab.reserve(nnz));
    GB_CUDA(d_order.reserve(order.size())); GB_CUDA(d_w.reserve(nnz)); GB_CUDA(d_prob.reserve(nnz)); GB_CUDA(d_mass.reserve(n_txp));
    GB_CUDA(d_eff.reserve(n_txp)); GB_CUDA(d_txp.reserve(n_txp));
    std::vector<double> mass(n_txp);
    const double prior = 1e-8;                                              // :215
    for (uint32_t i = 0; i < n_txp; ++i) mass[i] = prior + masses[i] * static_cast<double>(num_mapped);   // :219-221
    GB_CUDA(cudaMemcpyAsync(d_rp.p, k.h_row_ptr.data(), (E + 1) * 8, cudaMemcpyHostToDevice, s));
    if (E) GB_CUDA(cudaMemcpyAsync(d_cnt.p, k.h_counts.data(), E * 8, cudaMemcpyHostToDevice, s));
    if (nnz) GB_CUDA(cudaMemcpyAsync(d_lab.p, k.h_labels.data(), nnz * 4, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(d_order.p, order.data(), order.size() * 4, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(d_mass.p, mass.data(), n_txp * 8ull, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(d_eff.p, eff_lens, n_txp * 8ull, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemsetAsync(d_txp.p, 0, n_txp * 4ull, s));
    if (E) { k_entry_weights<<<gridn(E, 128), 128, 0, s>>>(d_rp.p, d_lab.p, d_cnt.p, d_eff.p, E, d_w.p); c->launches++; }
    GibbsParams gp;
    gp.row_ptr = d_rp.p; gp.labels = d_lab.p; gp.counts = d_cnt.p; gp.w = d_w.p; gp.countMap = d_cmap.p; gp.probMap = d_prob.p;
    gp.txpCount = d_txp.p; gp.mass = d_mass.p; gp.prior = prior; gp.seed = seed;
    if (E) { k_gibbs_init<<<gridn(E, 128), 128, 0, s>>>(gp, E); c->launches++; }
    GB_CUDA(cudaGetLastError());
    std::vector<int32_t> row(n_txp);
    int rc = SFB200_OK;
    for (uint32_t smp = 0; smp < n_samples && rc == SFB200_OK; ++smp) {
        // `bool numInternalRounds = 10;` in the reference => exactly one round per sample (:248,257)
        for (uint32_t q = 0; q < n_colours; ++q) {
            const uint64_t lo = col_start[q], hi = col_start[q + 1];
            if (hi > lo) { k_gibbs_round<<<gridn(hi - lo, 128), 128, 0, s>>>(gp, d_order.p, lo, hi, smp); c->launches++; }
        }
        GB_CUDA(cudaGetLastError());
        GB_CUDA(cudaMemcpyAsync(row.data(), d_txp.p, n_txp * 4ull, cudaMemcpyDeviceToHost, s));
        GB_CUDA(cudaStreamSynchronize(s));
        if (cb && cb(user, row.data(), n_txp) != 0) { c->err = "gibbs row callback failed"; rc = SFB200_ECALLBACK; }
    }
#undef GB_CUDA
    cleanup();
    return rc;
}

// ---- bootstrap resampling -----------------------------------------------------------------------------------------------------
// gatherBootstraps draws, per replicate, numMappedFragments class indices from the multinomial over the class counts
// (MultinomialSampler.hpp:13-64: one uniform and one binary search over the cumulative table per fragment).  The count vector it
// ends up with is Multinomial(N; counts / N), and that is drawn here directly, by conditional binomials down a tree of fan-out
// SPLIT_FAN over the classes: a node's share is split over its children by one thread, all nodes of a level in parallel.  E binomial
// draws instead of N uniform draws with N binary searches and N atomics (E = 4e5, N = 1e7 at cfg2) -- the same distribution, a
// different stream of random numbers (the reference's own is seeded from std::random_device: parity is distributional either way).
namespace {
constexpr uint32_t SPLIT_FAN = 8;        // 7 levels of 8 sequential draws per thread at 4e5 classes (48: 4 levels of 48, 3x the latency)

template <typename W>
__global__ void k_level_sums(const W* __restrict__ lower, uint64_t n_lower, uint64_t n_upper, double* __restrict__ upper) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n_upper) return;
    const uint64_t b = q * SPLIT_FAN, e = b + SPLIT_FAN < n_lower ? b + SPLIT_FAN : n_lower;
    double s = 0.0;
    for (uint64_t i = b; i < e; ++i) s += (double)lower[i];             // exact: integers below 2^53
    upper[q] = s;
}
// one thread per node of the upper level: its share over its children, weights = the children's totals
template <typename W>
__global__ void k_level_split(const W* __restrict__ lower_w, uint64_t n_lower, const unsigned long long* __restrict__ upper_share,
                              uint64_t n_upper, uint64_t seed, uint32_t level, unsigned long long* __restrict__ lower_share) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n_upper) return;
    const uint64_t b = q * SPLIT_FAN, e = b + SPLIT_FAN < n_lower ? b + SPLIT_FAN : n_lower;
    Rng g(seed, level, 7, q);
    double rem_p = 0.0;
    for (uint64_t i = b; i < e; ++i) rem_p += (double)lower_w[i];
    unsigned long long rem = upper_share[q];
    for (uint64_t i = b; i < e; ++i) {
        const double pi = (double)lower_w[i];
        unsigned long long x = 0;
        if (rem > 0) {
            if (i + 1 == e || !(rem_p > pi)) x = rem;
            else x = binomial_draw(g, rem, fmin(1.0, fmax(0.0, pi / rem_p)));
        }
        lower_share[i] = x;
        rem -= x; rem_p -= pi;
    }
}

struct SplitTree {
    std::vector<uint64_t> n;                       // nodes per level, n[0] = classes, n.back() = 1
    std::vector<double*> sums;                     // level >= 1
    std::vector<unsigned long long*> share;        // level >= 1 (level 0 is the caller's sample vector)
    DevBuf<double> d_sums; DevBuf<unsigned long long> d_share;
};
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
    // class counts in the order the per-sample count vectors are indexed by (canonical order)
    std::vector<uint64_t> canon_counts(E);
    if (k.from_device) {
        SFB_CUDA(c, cudaMemcpy(canon_counts.data(), k.cnt_all.p, E * 8, cudaMemcpyDeviceToHost));
    } else {
        canon_counts = k.h_counts;
    }
    cudaStream_t s = c->stream;
    DevBuf<unsigned long long> d_samp, d_cnt;
    SplitTree tr;
    DevBufScope<DevBuf<unsigned long long>, DevBuf<unsigned long long>, DevBuf<double>, DevBuf<unsigned long long>> scope(d_samp, d_cnt, tr.d_sums, tr.d_share);
    tr.n.push_back(E);
    while (tr.n.back() > 1) tr.n.push_back((tr.n.back() + SPLIT_FAN - 1) / SPLIT_FAN);
    if (tr.n.size() == 1) tr.n.push_back(1);                                          // a single class still has a root above it
    uint64_t upper_total = 0;
    for (size_t l = 1; l < tr.n.size(); ++l) upper_total += tr.n[l];
    SFB_CUDA(c, d_samp.reserve(E)); SFB_CUDA(c, d_cnt.reserve(E)); SFB_CUDA(c, tr.d_sums.reserve(upper_total)); SFB_CUDA(c, tr.d_share.reserve(upper_total));
    tr.sums.assign(tr.n.size(), nullptr); tr.share.assign(tr.n.size(), nullptr);
    { uint64_t at = 0; for (size_t l = 1; l < tr.n.size(); ++l) { tr.sums[l] = tr.d_sums.p + at; tr.share[l] = tr.d_share.p + at; at += tr.n[l]; } }
    SFB_CUDA(c, cudaMemcpyAsync(d_cnt.p, canon_counts.data(), E * 8, cudaMemcpyHostToDevice, s));
    k_level_sums<unsigned long long><<<gridn(tr.n[1], 128), 128, 0, s>>>(d_cnt.p, E, tr.n[1], tr.sums[1]);
    for (size_t l = 2; l < tr.n.size(); ++l) k_level_sums<double><<<gridn(tr.n[l], 128), 128, 0, s>>>(tr.sums[l - 1], tr.n[l - 1], tr.n[l], tr.sums[l]);
    c->launches += tr.n.size() - 1;
    const unsigned long long h_total = totalCount;
    SFB_CUDA(c, cudaMemcpyAsync(tr.share.back(), &h_total, 8, cudaMemcpyHostToDevice, s));    // the root's share is N, for every replicate
    std::vector<double> alphas(n_txp);
    // MultinomialSampler takes n as uint32_t (:15): the reference wraps above 2^32 fragments; we keep 64 bits
    int rc = SFB200_OK;
    double loop_ms = 0.0;
    const bool timing = getenv("SFB200_TIMING") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t_mark = now_ms();
    for (uint32_t b = 0; b < n_boot && rc == SFB200_OK; ++b) {
        const uint64_t sd = seed + 0x9E3779B97F4A7C15ULL * (b + 1);
        for (size_t l = tr.n.size() - 1; l >= 2; --l)
            k_level_split<double><<<gridn(tr.n[l], 64), 64, 0, s>>>(tr.sums[l - 1], tr.n[l - 1], tr.share[l], tr.n[l], sd, (uint32_t)l, tr.share[l - 1]);
        k_level_split<unsigned long long><<<gridn(tr.n[1], 64), 64, 0, s>>>(d_cnt.p, E, tr.share[1], tr.n[1], sd, 1u, d_samp.p);
        c->launches += tr.n.size() - 1;
        if (timing) { cudaStreamSynchronize(s); const double t1 = now_ms(); fprintf(stderr, "[sfb200-timing] bootstrap %u: resampling %.3f ms\n", b, t1 - t_mark); t_mark = t1; }
        uint32_t iters = 0;
        rc = sfb_bootstrap_em_device(c, eff_lens, n_txp, d_samp.p, totalCount, opts, alphas.data(), &iters);
        if (timing) { const double t1 = now_ms(); fprintf(stderr, "[sfb200-timing] bootstrap %u: em %.3f ms (loop %.3f ms, %u iterations)\n", b, t1 - t_mark, c->last_em_ms, iters); t_mark = t1; }
        c->eff_resident = true;                             // the same lengths for every replicate of this run
        loop_ms += c->last_em_ms;
        if (rc == SFB200_OK && cb && cb(user, alphas.data(), n_txp) != 0) { c->err = "bootstrap row callback failed"; rc = SFB200_ECALLBACK; }
        if (timing) { const double t1 = now_ms(); fprintf(stderr, "[sfb200-timing] bootstrap %u: callback %.3f ms\n", b, t1 - t_mark); t_mark = t1; }
    }
    c->eff_resident = false;
    c->last_em_ms = loop_ms;
    return rc;                                              // `scope` releases the replicate buffers
}

