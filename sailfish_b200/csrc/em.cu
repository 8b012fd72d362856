// em.cu -- inference over the ragged (equivalence class x transcript) structure on sm_100a.
//
// Replaces CollapsedEMOptimizer::optimize / EMUpdate_ / VBEMUpdate_ (reference src/CollapsedEMOptimizer.cpp:224-369,
// 711-893) and doBootstrap's inner loop (:476-514).  DESIGN.md section 4 describes the layout and the kernels:
//
//   * classes with >= 2 members are stored binned by member count so that a sub-warp group of g = 2,4,8,16,32 lanes
//     owns one class: one coalesced load of (label, weight) per lane, a shuffle reduction for the denominator
//     (E-step) and one red.global.add.f64 per lane (M-step scatter);
//   * single-member classes contribute a constant per-transcript vector which is the initial value of every output
//     buffer, so they never enter the sweep;
//   * three alpha buffers rotate (in / out / spare): while iteration n sweeps in -> out, the same pass compares
//     spare (alpha_{n-1}) with in (alpha_n) for the convergence rule and re-initialises spare for iteration n+1,
//     so an EM iteration costs ONE grid-wide barrier and no host round trip: the whole loop is one persistent
//     cooperative kernel (one CTA per SM).  Tile -> CTA assignment is static, so a CTA re-reads the same slice of
//     labels / weights every iteration and small problems are served from that SM's L1.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>

#include "common.cuh"
#include "em_segments.hpp"

namespace {

constexpr double DENORM_MIN = 4.9406564584124654e-324;   // std::numeric_limits<double>::denorm_min() (CollapsedEMOptimizer.cpp:33-34)
constexpr int EM_THREADS = 1024;

// control block (unsigned long long words)
enum { CTL_BAR_COUNT = 0, CTL_BAR_GEN = 1, CTL_MAXREL = 2 /*4 slots*/, CTL_CSUM = 6 /*4 slots*/, CTL_ITERS = 10,
       CTL_RESULT_BUF = 11, CTL_MRD = 12, CTL_TSUM = 13 /* sum of the truncated result */,
       CTL_PBAR_COUNT = 14, CTL_PBAR_GEN = 15 /* barrier among the pool CTAs of a hybrid run (em_dense.cuh) */,
       // k_em_dense's lagged stopping rule (em_dense.cuh): 16 slots each of max relative change, alpha sum and arrivals
       CTL_LAG_MAX = 16, CTL_LAG_SUM = 32, CTL_LAG_ARR = 48, CTL_WORDS = 64 };

struct EmParams {
    const uint32_t* start; const uint32_t* len; const uint32_t* lab; const double* w; const double* cnt;
    const double* base;     // T: initial value of an output buffer
    double* X;              // 3*T rotating alpha buffers
    double* theta;          // T (VBEM)
    unsigned long long* ctl;
    uint32_t T;
    uint64_t tile_start[SFB_NBINS + 1];
    uint64_t cls_start[SFB_NBINS + 1];
    int use_vb, gate_old;
    double tol, cutoff;
    uint32_t min_iter, max_iter, fixed_iters;
    double sum0;            // sum of alpha_0 (VBEM logNorm of the first iteration)
    double base_sum;        // sum of base
    uint32_t smem_bytes;    // dynamic shared memory given to the persistent kernel for its class slice
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) { return __ldcg(p); }

// all CTAs of the (cooperatively launched, hence co-resident) grid meet here
__device__ __forceinline__ void grid_barrier(unsigned long long* ctl, unsigned int nblocks, unsigned long long& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += 1;
        __threadfence();
        const unsigned long long prev = atomicAdd(&ctl[CTL_BAR_COUNT], 1ULL);
        if (prev + 1 == gen * nblocks) {
            st_release_u64(&ctl[CTL_BAR_GEN], gen);
        } else {
            while (ld_acquire_u64(&ctl[CTL_BAR_GEN]) < gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ---- the CTA's slice of the class structure --------------------------------------------------------------------------
// Tile -> CTA assignment is static, so a CTA reads the same classes every iteration.  When the slice fits in shared
// memory (it does for BASELINE configs 1-4: ~120 KB of the 227 KB per SM) it is staged there ONCE, at kernel start, with
// TMA bulk copies (cp.async.bulk + mbarrier), and a 1000-iteration run then touches global memory only for the alpha
// gathers and the red.add scatters.  Otherwise the same code reads the slice through the read-only path.
struct Slice {
    const uint32_t* start; const uint32_t* len; const double* cnt; const uint32_t* lab; const double* w;
    uint64_t c0;      // class index that start/len/cnt[0] correspond to
    uint64_t e0;      // entry index that lab/w[0] correspond to
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// first class of a tile (tiles of bin b hold 32 >> (b+1) classes, the long bin one)
template <typename P>
__device__ __forceinline__ uint64_t tile_first_class(const P& p, uint64_t tile) {
    int b = 0;
    while (b < SFB_NBINS - 1 && tile >= p.tile_start[b + 1]) ++b;
    if (tile >= p.tile_start[SFB_NBINS]) return p.cls_start[SFB_NBINS];
    return p.cls_start[b] + ((tile - p.tile_start[b]) << (b < SFB_NBINS - 1 ? 4 - b : 0));
}

constexpr int EM_ILP = 4;     // warp tiles in flight per warp: the alpha gathers of EM_ILP tiles are issued back to back

// ---- E-step + M-step scatter for the tiles of this CTA ---------------------------------------------------------------------
// EMUpdate_ (CollapsedEMOptimizer.cpp:235-277) / VBEMUpdate_ (:325-366).  `in` is alpha (EM) or expTheta (VBEM); members with a
// non-positive VBEM theta are skipped as in :342,357.  Returns this thread's sum of contributions (VBEM's alpha sum).
struct Bins { uint64_t cls_start[SFB_NBINS + 1], tile_start[SFB_NBINS + 1]; };
__device__ __forceinline__ void bins_set_tiles(Bins& b) {
    uint64_t t = 0;
    for (int i = 0; i < SFB_NBINS; ++i) {
        b.tile_start[i] = t;
        const uint64_t n = b.cls_start[i + 1] - b.cls_start[i];
        const uint64_t per = i < SFB_NBINS - 1 ? (32u >> (i + 1)) : 1u;
        t += (n + per - 1) / per;
    }
    b.tile_start[SFB_NBINS] = t;
}

__device__ __forceinline__ Bins em_bins(const EmParams& p) {
    Bins b;
    for (int i = 0; i <= SFB_NBINS; ++i) { b.cls_start[i] = p.cls_start[i]; b.tile_start[i] = p.tile_start[i]; }
    return b;
}

// SMA: alpha vectors live in shared memory, indexed by (transcript - toff) (the CTA-local partition, see em_part.cuh)
// count / denom: through the correctly rounded reciprocal (<= 1 ulp from the quotient; the tolerance is 1e-4) unless the
// denominator is so small that its reciprocal would overflow (a resampled count of 0 must still give 0, not 0 * inf)
__device__ __forceinline__ double sfb_div_count(double cnt, double denom) {
    return denom > 1e-290 ? cnt * __drcp_rn(denom) : cnt / denom;
}

// one bin: SH = log2(lanes per class); classes [cb, ce) and tiles [tb, te) are relative to the bin's first class / tile
template <bool VB, bool SMA, int SH>
__device__ __forceinline__ double sweep_bin(const Slice& sl, uint32_t bin_c0 /* bin's first class, slice-relative */, uint32_t bin_nc,
                                            uint32_t tb, uint32_t te, const double* __restrict__ in, double* __restrict__ out,
                                            uint32_t toff) {
    constexpr uint32_t G = 1u << SH, PER = 32u >> SH;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t e0 = (uint32_t)sl.e0;
    const uint32_t j = lane & (G - 1), sub = lane >> SH;
    double contrib = 0.0;
    for (uint32_t t0 = tb + warp; t0 < te; t0 += EM_ILP * W) {
        uint32_t tid[EM_ILP], ci[EM_ILP];
        double wv[EM_ILP], a[EM_ILP];
        bool ev[EM_ILP];
#pragma unroll
        for (int u = 0; u < EM_ILP; ++u) {
            const uint32_t tile = t0 + u * W;
            const uint32_t cl = tile * PER + sub;                       // class index inside the bin
            ev[u] = false; a[u] = 0.0; wv[u] = 0.0; tid[u] = 0; ci[u] = bin_c0 + cl;
            if (tile < te && cl < bin_nc) {
                const uint32_t o0 = sl.start[ci[u]] - e0, n = sl.len[ci[u]];
                if (j < n) {
                    tid[u] = sl.lab[o0 + j] - toff;
                    wv[u] = sl.w[o0 + j];
                    a[u] = SMA ? in[tid[u]] : ld_cg_f64(in + tid[u]);
                    ev[u] = true;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EM_ILP; ++u) {
            double v = 0.0;
            bool e = ev[u];
            if (e) { if (VB && !(a[u] > 0.0)) e = false; else v = a[u] * wv[u]; }
            double denom = v;
#pragma unroll
            for (uint32_t m = G >> 1; m >= 1; m >>= 1) denom += __shfl_xor_sync(0xffffffffu, denom, m);
            if (e && denom > DENORM_MIN && !isnan(v)) {
                // count / denom through the correctly rounded reciprocal (<= 1 ulp from the quotient; tolerance is 1e-4)
                const double add = v * sfb_div_count(sl.cnt[ci[u]], denom);
                atomicAdd(out + tid[u], add);
                contrib += add;
            }
        }
    }
    return contrib;
}

template <bool VB, bool SMA>
__device__ __forceinline__ double sweep_block(const Bins& p, const Slice& sl, uint64_t tile_lo, uint64_t tile_hi,
                                              const double* __restrict__ in, double* __restrict__ out, uint32_t toff) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    double contrib = 0.0;
    // bins 0..4: intersect the CTA's tile range with the bin's tiles
#define SFB_SWEEP_BIN(B)                                                                                                        \
    {                                                                                                                           \
        const uint64_t lo = tile_lo > p.tile_start[B] ? tile_lo : p.tile_start[B];                                              \
        const uint64_t hi = tile_hi < p.tile_start[B + 1] ? tile_hi : p.tile_start[B + 1];                                      \
        if (lo < hi)                                                                                                            \
            contrib += sweep_bin<VB, SMA, B + 1>(sl, (uint32_t)(p.cls_start[B] - sl.c0), (uint32_t)(p.cls_start[B + 1] - p.cls_start[B]), \
                                                 (uint32_t)(lo - p.tile_start[B]), (uint32_t)(hi - p.tile_start[B]), in, out, toff); \
    }
    SFB_SWEEP_BIN(0) SFB_SWEEP_BIN(1) SFB_SWEEP_BIN(2) SFB_SWEEP_BIN(3) SFB_SWEEP_BIN(4)
#undef SFB_SWEEP_BIN
    // long classes (more than 32 members): the whole warp walks the class twice
    const uint64_t long_lo = tile_lo > p.tile_start[SFB_NBINS - 1] ? tile_lo : p.tile_start[SFB_NBINS - 1];
    for (uint64_t tile = long_lo + warp; tile < tile_hi; tile += W) {
        const uint64_t c = p.cls_start[SFB_NBINS - 1] + (tile - p.tile_start[SFB_NBINS - 1]);
        if (c >= p.cls_start[SFB_NBINS]) continue;
        const uint32_t o0 = sl.start[c - sl.c0] - (uint32_t)sl.e0, n = sl.len[c - sl.c0];
        double denom = 0.0;
        for (uint32_t j = lane; j < n; j += 32) {
            const double al = SMA ? in[sl.lab[o0 + j] - toff] : ld_cg_f64(in + sl.lab[o0 + j]);
            if (!VB || al > 0.0) denom += al * sl.w[o0 + j];
        }
        denom = warp_sum(denom);
        if (denom > DENORM_MIN) {
            const double inv = sfb_div_count(sl.cnt[c - sl.c0], denom);
            for (uint32_t j = lane; j < n; j += 32) {
                const uint32_t t = sl.lab[o0 + j] - toff;
                const double al = SMA ? in[t] : ld_cg_f64(in + t);
                if (VB && !(al > 0.0)) continue;
                const double v = al * sl.w[o0 + j];
                if (!isnan(v)) { const double add = v * inv; atomicAdd(out + t, add); contrib += add; }
            }
        }
    }
    return contrib;
}

// Stage the CTA's slice in shared memory (dynamic smem, 16-byte aligned) with TMA bulk copies.  Source ranges are widened
// to 16-byte boundaries (the arrays are allocated with slack), which shifts c0 / e0 down accordingly.
// Returns false (and leaves `sl` pointing at global memory) if the slice does not fit in `smem_bytes`.
__device__ bool stage_slice(const EmParams& p, uint64_t tile_lo, uint64_t tile_hi, unsigned char* smem, uint32_t smem_bytes,
                            uint64_t* bar, Slice& sl) {
    sl.start = p.start; sl.len = p.len; sl.cnt = p.cnt; sl.lab = p.lab; sl.w = p.w; sl.c0 = 0; sl.e0 = 0;
    const uint64_t Em = p.cls_start[SFB_NBINS];
    uint64_t c_lo = tile_first_class(p, tile_lo), c_hi = tile_first_class(p, tile_hi);
    if (c_hi > Em) c_hi = Em;
    if (c_lo >= c_hi) return true;                                  // nothing to do for this CTA
    uint64_t e_lo = __ldg(p.start + c_lo);
    uint64_t e_hi = (uint64_t)__ldg(p.start + c_hi - 1) + __ldg(p.len + c_hi - 1);
    c_lo &= ~3ULL; e_lo &= ~3ULL;                                    // 16-byte aligned u32 ranges (and 32-byte f64 ranges)
    const uint64_t nc = ((c_hi - c_lo) + 3) & ~3ULL, ne = ((e_hi - e_lo) + 3) & ~3ULL;
    const uint64_t need = nc * 4 * 2 + nc * 8 + ne * 4 + ne * 8;
    if (need > smem_bytes) return false;
    double* s_cnt = reinterpret_cast<double*>(smem);
    double* s_w = s_cnt + nc;
    uint32_t* s_start = reinterpret_cast<uint32_t*>(s_w + ne);
    uint32_t* s_len = s_start + nc;
    uint32_t* s_lab = s_len + nc;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"((uint32_t)need) : "memory");
        tma_load_1d(s_cnt, p.cnt + c_lo, (uint32_t)(nc * 8), bar);
        tma_load_1d(s_w, p.w + e_lo, (uint32_t)(ne * 8), bar);
        tma_load_1d(s_start, p.start + c_lo, (uint32_t)(nc * 4), bar);
        tma_load_1d(s_len, p.len + c_lo, (uint32_t)(nc * 4), bar);
        tma_load_1d(s_lab, p.lab + e_lo, (uint32_t)(ne * 4), bar);
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    }
    sl.start = s_start; sl.len = s_len; sl.cnt = s_cnt; sl.lab = s_lab; sl.w = s_w; sl.c0 = c_lo; sl.e0 = e_lo;
    return true;
}

// ---- per-transcript pass: convergence rule + re-initialise the spare buffer (+ VBEM expTheta) -------------------------
// CollapsedEMOptimizer.cpp:849-861 (gate on the NEW alpha) / :496-508 (doBootstrap: gate on the OLD alpha);
// VBEM: :300-320.  Returns this thread's max relative difference encoded as bits+1 (0 = no transcript passed the gate).
template <bool VB>
__device__ __forceinline__ unsigned long long transcript_pass(const EmParams& p, const double* prev, const double* cur,
                                                              double* spare_out, bool compare, bool reset, double logNorm,
                                                              uint64_t tid0, uint64_t stride) {
    unsigned long long best = 0ULL;
    const double thetaScale = (VB && reset) ? exp(-logNorm) : 0.0;
    for (uint64_t t = tid0; t < p.T; t += stride) {
        const double c = ld_cg_f64(cur + t);
        if (compare) {
            const double pv = ld_cg_f64(prev + t);
            const double gate = p.gate_old ? pv : c;
            if (gate > p.cutoff) {
                const double rel = fabs(pv - c) / c;
                // rel >= 0 (or +inf / nan): non-negative doubles order like their bit patterns
                const unsigned long long bits = (unsigned long long)__double_as_longlong(rel) + 1ULL;
                best = bits > best ? bits : best;
            }
        }
        if (reset) spare_out[t] = __ldg(p.base + t);
        if (VB && reset) p.theta[t] = (c > DENORM_MIN) ? sfb_exp_theta(c, logNorm, thetaScale) : 0.0;
    }
    return best;
}

__device__ __forceinline__ void block_max_to_slot(unsigned long long v, unsigned long long* slot, unsigned long long* sm) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
        v = o > v ? o : v;
    }
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : 0ULL;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
            v = o > v ? o : v;
        }
        if (lane == 0 && v) atomicMax(slot, v);
    }
}

__device__ __forceinline__ void block_sum_to_slot(double v, double* slot, double* sm) {
    v = warp_sum(v);
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(slot, v);
    }
}

__device__ __forceinline__ double decode_mrd(unsigned long long bits1) {
    return bits1 ? __longlong_as_double((long long)(bits1 - 1ULL)) : -1.7976931348623157e308;
}

// ---- the persistent EM loop --------------------------------------------------------------------------------------------
template <bool VB>
__global__ void __launch_bounds__(EM_THREADS, 1) k_em_persistent(const EmParams p) {
    __shared__ unsigned long long sm_u[32];
    __shared__ double sm_d[32];
    const unsigned nblocks = gridDim.x;
    unsigned long long gen = 0;
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gstride = (uint64_t)nblocks * blockDim.x;
    const uint64_t n_tiles = p.tile_start[SFB_NBINS];
    const uint64_t tile_lo = n_tiles * blockIdx.x / nblocks, tile_hi = n_tiles * (blockIdx.x + 1ULL) / nblocks;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ uint64_t tma_bar;
    Slice sl;
    stage_slice(p, tile_lo, tile_hi, dyn_smem, p.smem_bytes, &tma_bar, sl);

    unsigned bi = 0, bo = 1, bs = 2;                       // buffer indices: in / out / spare
    const bool fixed = p.fixed_iters > 0;
    uint32_t n = 0;
    for (;;) {
        double* in = p.X + (size_t)bi * p.T;
        double* out = p.X + (size_t)bo * p.T;
        double* spare = p.X + (size_t)bs * p.T;
        unsigned long long* slot = p.ctl + CTL_MAXREL + (n & 3u);
        const bool last = fixed ? (n >= p.fixed_iters) : (n >= p.max_iter && n >= p.min_iter);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.ctl[CTL_MAXREL + ((n + 2u) & 3u)] = 0ULL;
            p.ctl[CTL_CSUM + ((n + 2u) & 3u)] = 0ULL;
        }
        double logNorm = 0.0;
        if (VB && !last) {
            const double asum = (n == 0) ? p.sum0
                : p.base_sum + __longlong_as_double((long long)ld_cg_u64(p.ctl + CTL_CSUM + (n & 3u)));
            logNorm = sfb_digamma(asum);
        }
        const unsigned long long best = transcript_pass<VB>(p, spare, in, spare, n > 0, !last, logNorm, gtid, gstride);
        if (n > 0) block_max_to_slot(best, slot, sm_u);
        if (last) {
            grid_barrier(p.ctl, nblocks, gen);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = bi; p.ctl[CTL_MRD] = ld_cg_u64(slot);
            }
            return;
        }
        if (VB) grid_barrier(p.ctl, nblocks, gen);         // expTheta complete before anyone gathers it
        const double* src = VB ? p.theta : in;
        double contrib = 0.0;
        contrib = sweep_block<VB, false>(em_bins(p), sl, tile_lo, tile_hi, src, out, 0u);
        if (VB) block_sum_to_slot(contrib, reinterpret_cast<double*>(p.ctl + CTL_CSUM + ((n + 1u) & 3u)), sm_d);
        grid_barrier(p.ctl, nblocks, gen);
        // check(alpha_{n-1}, alpha_n) is complete now: would the reference loop (:820 / :486) have stopped at itNum == n?
        if (!fixed && n > 0 && n >= p.min_iter) {
            const unsigned long long mr = ld_cg_u64(slot);
            const bool converged = !(decode_mrd(mr) > p.tol);
            if (converged) {
                if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = bi; p.ctl[CTL_MRD] = mr; }
                return;
            }
        }
        const unsigned tmp = bs; bs = bi; bi = bo; bo = tmp;
        ++n;
    }
}

// ---- the same iteration as separate launches (multi-rank runs with an all-reduce in between, and profiling) ------------
template <bool VB>
__global__ void __launch_bounds__(EM_THREADS, 1) k_em_transcript_pass(const EmParams p, unsigned bi, unsigned bs, uint32_t n,
                                                                      int compare, int reset, double asum) {
    __shared__ unsigned long long sm_u[32];
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
    double* in = p.X + (size_t)bi * p.T;
    double* spare = p.X + (size_t)bs * p.T;
    const double logNorm = (VB && reset) ? sfb_digamma(asum) : 0.0;
    const unsigned long long best = transcript_pass<VB>(p, spare, in, spare, compare != 0, reset != 0, logNorm, gtid, gstride);
    if (compare) block_max_to_slot(best, p.ctl + CTL_MAXREL + (n & 3u), sm_u);
}

template <bool VB>
__global__ void __launch_bounds__(EM_THREADS, 1) k_em_sweep(const EmParams p, unsigned bi, unsigned bo, uint32_t n) {
    __shared__ double sm_d[32];
    const unsigned nblocks = gridDim.x;
    const uint64_t n_tiles = p.tile_start[SFB_NBINS];
    const uint64_t tile_lo = n_tiles * blockIdx.x / nblocks, tile_hi = n_tiles * (blockIdx.x + 1ULL) / nblocks;
    const double* src = VB ? p.theta : p.X + (size_t)bi * p.T;
    double* out = p.X + (size_t)bo * p.T;
    Slice sl;
    sl.start = p.start; sl.len = p.len; sl.cnt = p.cnt; sl.lab = p.lab; sl.w = p.w; sl.c0 = 0; sl.e0 = 0;
    const double contrib = sweep_block<VB, false>(em_bins(p), sl, tile_lo, tile_hi, src, out, 0u);
    if (VB) block_sum_to_slot(contrib, reinterpret_cast<double*>(p.ctl + CTL_CSUM + ((n + 1u) & 3u)), sm_d);
}

#include "em_part.cuh"
#include "em_gather.cuh"
#include "em_dense.cuh"

// truncateCountVector (CollapsedEMOptimizer.cpp:37-44): alpha <= cutoff -> 0, and the sum of what is left
__global__ void k_truncate(double* __restrict__ x, uint32_t n, double cutoff, double* __restrict__ sum_out) {
    __shared__ double sm_d[32];
    double s = 0.0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v <= cutoff) { v = 0.0; x[i] = 0.0; }
        s += v;
    }
    block_sum_to_slot(s, sum_out, sm_d);
}

__global__ void k_sum_f64(const double* __restrict__ x, uint32_t n, double* __restrict__ out) {
    __shared__ double sm_d[32];
    double s = 0.0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += x[i];
    block_sum_to_slot(s, out, sm_d);
}

// ---- set-up kernels -------------------------------------------------------------------------------------------------------
// CollapsedEMOptimizer.cpp:733-740 (clamp) and :745-772: w_i = count / effLen[t_i]; w_i *= 1 / sum_i w_i
__global__ void k_clamp_eff(const double* __restrict__ eff_in, uint32_t T, double* __restrict__ eff) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) { const double e = eff_in[i]; eff[i] = (e <= 1.0) ? 1.0 : e; }
}
__global__ void k_class_weights(const uint32_t* __restrict__ start, const uint32_t* __restrict__ len, const uint32_t* __restrict__ lab,
                                const double* __restrict__ cnt, const double* __restrict__ eff, uint64_t Em, double* __restrict__ w) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= Em) return;
    const uint32_t b = start[c], e = b + len[c];
    const double count = cnt[c];
    double wsum = 0.0;
    for (uint32_t j = b; j < e; ++j) { const double v = count / eff[lab[j]]; w[j] = v; wsum += v; }
    const double wnorm = 1.0 / wsum;
    for (uint32_t j = b; j < e; ++j) w[j] *= wnorm;
}
// alpha_0 (:800-803), base = single (+ prior), and the first output buffer
__global__ void k_em_init(const uint8_t* __restrict__ active, const double* __restrict__ single, uint32_t T, double alpha0,
                          double prior_term, double* __restrict__ X, double* __restrict__ base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const double b = single[i] + prior_term;
    base[i] = b;
    X[i] = active[i] ? alpha0 : 0.0;
    X[(size_t)T + i] = b;
    X[2 * (size_t)T + i] = b;
}
// per-sample counts (bootstrap): binned counts and the single-member vector from canonical-order counts
__global__ void k_permute_counts(const unsigned long long* __restrict__ samp, const uint32_t* __restrict__ perm, uint64_t Em,
                                 double* __restrict__ cnt) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < Em) cnt[c] = (double)samp[perm ? perm[c] : c];
}
__global__ void k_scatter_single(const unsigned long long* __restrict__ samp, const uint32_t* __restrict__ sgl_cls,
                                 const uint32_t* __restrict__ sgl_tid, uint64_t n_sgl, double* __restrict__ single) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n_sgl) atomicAdd(single + sgl_tid[i], (double)samp[sgl_cls[i]]);
}

inline unsigned grid_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// SFB200_TIMING=1: host wall-clock between marks of an em_run, on stderr (where does the set-up time go)
struct HostMarks {
    bool on; double last;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    HostMarks() : on(getenv("SFB200_TIMING") != nullptr), last(now()) {}
    void mark(const char* what) { if (!on) return; const double t = now(); fprintf(stderr, "[sfb200-timing] %-28s %8.3f ms\n", what, t - last); last = t; }
};

}  // namespace

// ======================================================================================================================
// host side
// ======================================================================================================================

struct EmExtra {   // per-sample device arrays of the bootstrap path
    DevBuf<double> cnt_s;                 // per-sample binned counts
    DevBuf<double> single_s;              // per-sample single-member vector
    DevBuf<unsigned long long> samp;      // canonical-order per-sample counts on the device
};
static EmExtra* g_extra_for(sfb200_ctx* c) {
    if (!c->em_extra) c->em_extra = new EmExtra();
    return static_cast<EmExtra*>(c->em_extra);
}
void sfb_em_extra_free(sfb200_ctx* c) {
    EmExtra* e = static_cast<EmExtra*>(c->em_extra);
    if (!e) return;
    e->cnt_s.release(); e->single_s.release(); e->samp.release();
    delete e;
    c->em_extra = nullptr;
}

// Build the binned device layout from canonical CSR on the host.
int sfb_classes_from_host(sfb200_ctx* c, uint32_t n_txp, uint64_t E, const uint64_t* row_ptr, const uint32_t* labels,
                          const uint64_t* counts) {
    DevClasses& k = c->cls;
    k.ready = false;
    k.part.valid = false;
    const uint64_t nnz = E ? row_ptr[E] : 0;
    for (uint64_t i = 0; i < nnz; ++i) if (labels[i] >= n_txp) SFB_FAIL(c, SFB200_EINVAL, "label holds a transcript id >= n_txp");
    k.n_txp = n_txp; k.E = E; k.nnz = nnz;
    k.h_row_ptr.assign(row_ptr, row_ptr + E + 1);
    if (E == 0) k.h_row_ptr.assign(1, 0);
    k.h_labels.assign(labels, labels + nnz);
    k.h_counts.assign(counts, counts + E);
    k.export_to_canon.clear();
    k.host_valid = true; k.from_device = false; k.merged = false;

    auto bin_of = [](uint64_t n) { return n <= 2 ? 0 : n <= 4 ? 1 : n <= 8 ? 2 : n <= 16 ? 3 : n <= 32 ? 4 : 5; };
    uint64_t bin_n[SFB_NBINS] = {0, 0, 0, 0, 0, 0};
    uint64_t nnzm = 0, n_sgl = 0, total = 0;
    for (uint64_t e = 0; e < E; ++e) {
        const uint64_t n = row_ptr[e + 1] - row_ptr[e];
        total += counts[e];
        if (n == 1) ++n_sgl;
        else if (n >= 2) { bin_n[bin_of(n)]++; nnzm += n; }
    }
    if (nnzm >= 0xFFFFFFFFull) SFB_FAIL(c, SFB200_EINVAL, "more than 2^32 label entries");
    k.total_count = total;
    k.bin_cls[0] = 0;
    for (int b = 0; b < SFB_NBINS; ++b) k.bin_cls[b + 1] = k.bin_cls[b] + bin_n[b];
    k.Em = k.bin_cls[SFB_NBINS]; k.nnzm = nnzm;

    std::vector<uint32_t> perm(k.Em), start(k.Em), len(k.Em), lab(nnzm), sgl_cls(n_sgl), sgl_tid(n_sgl);
    std::vector<double> cnt(k.Em), single(n_txp, 0.0);
    std::vector<uint8_t> active(n_txp, 0);
    uint64_t cur[SFB_NBINS];
    for (int b = 0; b < SFB_NBINS; ++b) cur[b] = k.bin_cls[b];
    uint64_t si = 0;
    for (uint64_t e = 0; e < E; ++e) {
        const uint64_t n = row_ptr[e + 1] - row_ptr[e];
        if (n == 1) {
            const uint32_t t = labels[row_ptr[e]];
            single[t] += static_cast<double>(counts[e]);
            sgl_cls[si] = static_cast<uint32_t>(e); sgl_tid[si] = t; ++si;
        } else if (n >= 2) {
            perm[cur[bin_of(n)]++] = static_cast<uint32_t>(e);
        }
    }
    uint64_t o = 0;
    for (uint64_t i = 0; i < k.Em; ++i) {
        const uint64_t e = perm[i];
        start[i] = static_cast<uint32_t>(o);
        len[i] = static_cast<uint32_t>(row_ptr[e + 1] - row_ptr[e]);
        cnt[i] = static_cast<double>(counts[e]);
        for (uint64_t j = row_ptr[e]; j < row_ptr[e + 1]; ++j) lab[o++] = labels[j];
    }
    uint64_t n_active = 0;
    for (uint64_t i = 0; i < nnz; ++i) if (!active[labels[i]]) { active[labels[i]] = 1; ++n_active; }   // :774-782
    k.n_active = n_active;

    cudaSetDevice(c->device);
    k.n_sgl = n_sgl;
    SFB_CUDA(c, k.start.reserve(k.Em)); SFB_CUDA(c, k.len.reserve(k.Em)); SFB_CUDA(c, k.lab.reserve(nnzm)); SFB_CUDA(c, k.w.reserve(nnzm));
    SFB_CUDA(c, k.cnt.reserve(k.Em)); SFB_CUDA(c, k.perm.reserve(k.Em)); SFB_CUDA(c, k.single.reserve(n_txp));
    SFB_CUDA(c, k.active.reserve(n_txp)); SFB_CUDA(c, k.sgl_cls.reserve(n_sgl)); SFB_CUDA(c, k.sgl_tid.reserve(n_sgl));
    cudaStream_t s = c->stream;
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(k.start.p, start.data(), k.Em * 4, cudaMemcpyHostToDevice, s));
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(k.len.p, len.data(), k.Em * 4, cudaMemcpyHostToDevice, s));
    if (nnzm) SFB_CUDA(c, cudaMemcpyAsync(k.lab.p, lab.data(), nnzm * 4, cudaMemcpyHostToDevice, s));
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(k.cnt.p, cnt.data(), k.Em * 8, cudaMemcpyHostToDevice, s));
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(k.perm.p, perm.data(), k.Em * 4, cudaMemcpyHostToDevice, s));
    if (n_txp) SFB_CUDA(c, cudaMemcpyAsync(k.single.p, single.data(), n_txp * 8ull, cudaMemcpyHostToDevice, s));
    if (n_txp) SFB_CUDA(c, cudaMemcpyAsync(k.active.p, active.data(), n_txp, cudaMemcpyHostToDevice, s));
    if (n_sgl) SFB_CUDA(c, cudaMemcpyAsync(k.sgl_cls.p, sgl_cls.data(), n_sgl * 4, cudaMemcpyHostToDevice, s));
    if (n_sgl) SFB_CUDA(c, cudaMemcpyAsync(k.sgl_tid.p, sgl_tid.data(), n_sgl * 4, cudaMemcpyHostToDevice, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));   // the host vectors die here
    k.ready = true;
    return SFB200_OK;
}

// Host copy of the classes in export order.  After a device-side finish the classes exist only on the device, in binned
// order; eq_export (aux/eq_classes.txt, tests) wants them label-lexicographic, so download and sort here, on demand.
int sfb_classes_host(sfb200_ctx* c) {
    DevClasses& k = c->cls;
    if (k.host_valid) return SFB200_OK;
    if (!k.from_device) SFB_FAIL(c, SFB200_EINVAL, "classes have no host copy");
    cudaSetDevice(c->device);
    std::vector<uint32_t> start(k.Em), len(k.Em), lab(k.nnzm ? k.nnzm : 1), sgl_tid(k.n_sgl);
    std::vector<uint64_t> cnt_all(k.E);
    cudaStream_t s = c->stream;
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(start.data(), k.start.p, k.Em * 4, cudaMemcpyDeviceToHost, s));
    if (k.Em) SFB_CUDA(c, cudaMemcpyAsync(len.data(), k.len.p, k.Em * 4, cudaMemcpyDeviceToHost, s));
    if (k.nnzm) SFB_CUDA(c, cudaMemcpyAsync(lab.data(), k.lab.p, k.nnzm * 4, cudaMemcpyDeviceToHost, s));
    if (k.n_sgl) SFB_CUDA(c, cudaMemcpyAsync(sgl_tid.data(), k.sgl_tid.p, k.n_sgl * 4, cudaMemcpyDeviceToHost, s));
    if (k.E) SFB_CUDA(c, cudaMemcpyAsync(cnt_all.data(), k.cnt_all.p, k.E * 8, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    auto lab_of = [&](uint64_t ci, uint32_t& n) -> const uint32_t* {
        if (ci < k.Em) { n = len[ci]; return lab.data() + start[ci]; }
        n = 1; return sgl_tid.data() + (ci - k.Em);
    };
    std::vector<uint64_t> order(k.E);
    for (uint64_t i = 0; i < k.E; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        uint32_t na, nb; const uint32_t* pa = lab_of(a, na); const uint32_t* pb = lab_of(b, nb);
        return std::lexicographical_compare(pa, pa + na, pb, pb + nb);
    });
    k.h_row_ptr.assign(k.E + 1, 0); k.h_counts.resize(k.E); k.h_labels.resize(k.nnz);
    uint64_t z = 0;
    for (uint64_t e = 0; e < k.E; ++e) {
        uint32_t n; const uint32_t* pl = lab_of(order[e], n);
        std::memcpy(k.h_labels.data() + z, pl, n * 4ull);
        z += n; k.h_row_ptr[e + 1] = z; k.h_counts[e] = cnt_all[order[e]];
    }
    k.export_to_canon = order;
    k.host_valid = true;
    return SFB200_OK;
}

extern "C" int sfb200_eq_import(sfb200_ctx* c, uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr,
                                const uint32_t* labels, const uint64_t* counts) {
    if (!c) return SFB200_EINVAL;
    if (n_classes && (!row_ptr || !labels || !counts)) SFB_FAIL(c, SFB200_EINVAL, "eq_import: null array");
    static const uint64_t zero = 0;
    return sfb_classes_from_host(c, n_txp, n_classes, n_classes ? row_ptr : &zero, labels, counts);
}

extern "C" int sfb200_eq_export(sfb200_ctx* c, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts) {
    if (!c) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "eq_export: no classes (call map_finish or eq_import first)");
    { const int rc = sfb_classes_host(c); if (rc) return rc; }
    const DevClasses& k = c->cls;
    if (row_ptr) std::memcpy(row_ptr, k.h_row_ptr.data(), (k.E + 1) * 8);
    if (labels && k.nnz) std::memcpy(labels, k.h_labels.data(), k.nnz * 4);
    if (counts && k.E) std::memcpy(counts, k.h_counts.data(), k.E * 8);
    return SFB200_OK;
}

extern "C" void sfb200_em_default_opts(sfb200_em_opts* o) {
    o->use_vb = 0; o->prior_alpha = 0.01; o->tol = 0.01; o->min_iter = 50; o->max_iter = 10000; o->fixed_iters = 0;
    o->check_cutoff = 1e-2; o->min_alpha = 1e-8;
}

extern "C" double sfb200_last_em_loop_ms(const sfb200_ctx* c) { return c ? c->last_em_ms : 0.0; }
extern "C" int sfb200_last_em_kernel(const sfb200_ctx* c) { return c ? c->last_em_kernel : 0; }
extern "C" int sfb200_last_em_variant(const sfb200_ctx* c) { return c ? c->last_em_variant : 0; }

namespace {


// Build the CTA partition of the current classes (em_part.cuh).  Device kernels do the per-class work; the host only scans
// a T-long load histogram and the (n_cta+1) x 6 group table.
int build_gather(sfb200_ctx* c, const std::vector<unsigned long long>& tbl);
// the atomic-free loop is chosen when the classes allow it; SFB200_EM_GATHER=0 / 1 overrides the default
constexpr bool SFB_GATHER_DEFAULT = true;
bool gather_enabled() { const char* e = getenv("SFB200_EM_GATHER"); return e ? atoi(e) != 0 : SFB_GATHER_DEFAULT; }
// one thread per connected component (em_dense.cuh) when every component is small; SFB200_EM_DENSE=0 / 1 overrides the default
constexpr bool SFB_DENSE_DEFAULT = true;
bool dense_enabled() { const char* e = getenv("SFB200_EM_DENSE"); return e ? atoi(e) != 0 : SFB_DENSE_DEFAULT; }
int build_dense(sfb200_ctx* c, const std::vector<unsigned long long>& tbl, bool mark_large, bool* marked);

// hybrid runs (em_dense.cuh): small components on the component CTAs, everything else in the pool loop; SFB200_EM_HYBRID=0 switches it off
bool hybrid_enabled() { const char* e = getenv("SFB200_EM_HYBRID"); return e ? atoi(e) != 0 : false; }   // opt-in: 16.9 vs 17.5 us per iteration on the paralog set, not worth a default

// the partition for n_cta ranges; *marked_again is set when the dense builder sent large components to the pool and wants another pass
int build_partition_n(sfb200_ctx* c, uint32_t n_cta, int per_sm) {
    DevClasses& k = c->cls;
    DevPartition& P = k.part;
    const uint32_t T = k.n_txp;
    const uint64_t Em = k.Em, nnzm = k.nnzm;
    cudaStream_t s = c->stream;
    HostMarks hm;
    P.n_cta = n_cta;
    SFB_CUDA(c, P.load.reserve(T)); SFB_CUDA(c, P.bounds.reserve(n_cta + 1)); SFB_CUDA(c, P.owner.reserve(Em)); SFB_CUDA(c, P.dirty.reserve(T));
    SFB_CUDA(c, P.grp.reserve(3 * (size_t)(n_cta + 1) * SFB_NBINS + 4));
    SFB_CUDA(c, P.start.reserve(Em)); SFB_CUDA(c, P.len.reserve(Em)); SFB_CUDA(c, P.lab.reserve(nnzm)); SFB_CUDA(c, P.src.reserve(Em));
    SFB_CUDA(c, P.cnt.reserve(Em)); SFB_CUDA(c, P.w.reserve(nnzm)); SFB_CUDA(c, P.tbl.reserve((size_t)n_cta * PT_WORDS));
    // 1. ranges balanced by sweep cost (lanes occupied by the classes whose smallest member falls in the range)
    SFB_CUDA(c, cudaMemsetAsync(P.load.p, 0, T * 4ull, s));
    // crossing profile (reuses the owner buffer as int[T+1] scratch)
    SFB_CUDA(c, P.owner.reserve(std::max<uint64_t>(Em, (uint64_t)T + 1)));
    int* d_diff = reinterpret_cast<int*>(P.owner.p);
    SFB_CUDA(c, cudaMemsetAsync(d_diff, 0, (T + 1) * 4ull, s));
    k_part_span<<<grid_for(Em, 256), 256, 0, s>>>(k.start.p, k.len.p, k.lab.p, Em, d_diff, P.load.p);
    c->launches++;
    // boundaries: balanced by sweep cost, each moved to the nearest position no class crosses (fewer pool classes); on the device,
    // the host reads them back together with the group sizes below
    std::vector<uint32_t> bounds(n_cta + 1);
    SFB_CUDA(c, P.pre.reserve(T));
    k_part_bounds<<<1, 1024, (n_cta + 1) * sizeof(uint32_t), s>>>(P.load.p, d_diff, T, n_cta, std::max<uint32_t>(8, T / n_cta / 2), P.pre.p, P.bounds.p);
    c->launches++;
    hm.mark("part: span + bounds launch");
    SFB_CUDA(c, cudaMemsetAsync(P.dirty.p, 0, T, s));
    unsigned int* d_changed = reinterpret_cast<unsigned int*>(P.grp.p + 3 * (size_t)(n_cta + 1) * SFB_NBINS);
    std::vector<unsigned long long> tbl((size_t)n_cta * PT_WORDS, 0);
    int max_optin = 0;
    SFB_CUDA(c, cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    P.smem_limit = (uint64_t)max_optin;
    for (int pass = 0; pass < 3; ++pass) {
        // 2. closure of "crosses a range or touches a dirty transcript" (pass > 0: the dense builder marked large components dirty)
        for (int round = 0; round < 256; ++round) {
            SFB_CUDA(c, cudaMemsetAsync(d_changed, 0, 4, s));
            k_part_owner<<<grid_for(Em, 256), 256, 0, s>>>(k.start.p, k.len.p, k.lab.p, Em, P.bounds.p, n_cta, P.dirty.p, P.owner.p, d_changed);
            c->launches++;
            unsigned int ch = 0;
            SFB_CUDA(c, cudaMemcpyAsync(&ch, d_changed, 4, cudaMemcpyDeviceToHost, s));
            SFB_CUDA(c, cudaStreamSynchronize(s));
            if (!ch) break;
        }
        hm.mark("part: closure rounds");
        // 3. group sizes -> offsets
        const size_t G = (size_t)(n_cta + 1) * SFB_NBINS;
        unsigned long long* d_grp = P.grp.p; unsigned long long* d_cls_off = P.grp.p + G; unsigned long long* d_nnz_off = P.grp.p + 2 * G;
        SFB_CUDA(c, cudaMemsetAsync(d_grp, 0, G * 8, s));
        k_part_count<<<grid_for(Em, 256), 256, 0, s>>>(P.owner.p, k.len.p, Em, n_cta, d_grp);
        c->launches++;
        std::vector<unsigned long long> grp(G), cls_off(G + 1), nnz_off(G + 1);
        SFB_CUDA(c, cudaMemcpyAsync(grp.data(), d_grp, G * 8, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaMemcpyAsync(bounds.data(), P.bounds.p, (n_cta + 1) * 4ull, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        cls_off[0] = 0; nnz_off[0] = 0;
        for (size_t g = 0; g < G; ++g) { cls_off[g + 1] = cls_off[g] + (grp[g] >> 32); nnz_off[g + 1] = nnz_off[g] + (grp[g] & 0xFFFFFFFFULL); }
        SFB_CUDA(c, cudaMemcpyAsync(d_cls_off, cls_off.data(), G * 8, cudaMemcpyHostToDevice, s));
        SFB_CUDA(c, cudaMemcpyAsync(d_nnz_off, nnz_off.data(), G * 8, cudaMemcpyHostToDevice, s));
        SFB_CUDA(c, cudaMemsetAsync(d_grp, 0, G * 8, s));                  // reused as the fill cursors
        k_part_fill<<<grid_for(Em, 256), 256, 0, s>>>(P.owner.p, k.start.p, k.len.p, k.lab.p, k.cnt.p, Em, n_cta, d_cls_off, d_nnz_off, d_grp,
                                                       P.start.p, P.len.p, P.lab.p, P.cnt.p, P.src.p);
        c->launches++;
        hm.mark("part: counts + fill launch");
        // 4. per-CTA table and the shared-memory budget
        uint64_t max_bytes = 0, max_bytes_vb = 0;
        for (uint32_t i = 0; i < n_cta; ++i) {
            unsigned long long* row = tbl.data() + (size_t)i * PT_WORDS;
            for (int b = 0; b <= SFB_NBINS; ++b) row[PT_CLS + b] = cls_off[(size_t)i * SFB_NBINS + b];
            row[PT_ENT0] = nnz_off[(size_t)i * SFB_NBINS]; row[PT_ENT1] = nnz_off[(size_t)(i + 1) * SFB_NBINS];
            row[PT_TXP0] = bounds[i]; row[PT_TXP1] = bounds[i + 1];
            const uint64_t nc = row[PT_CLS + SFB_NBINS] - (row[PT_CLS] & ~3ULL) + 4, ne = row[PT_ENT1] - (row[PT_ENT0] & ~3ULL) + 4;
            const uint64_t nt = bounds[i + 1] - bounds[i] + 4;
            max_bytes = std::max<uint64_t>(max_bytes, nc * 16 + ne * 12 + nt * (8 * 2 + 1) + 256);
            max_bytes_vb = std::max<uint64_t>(max_bytes_vb, nc * 16 + ne * 12 + nt * (8 * 3 + 1) + 256);
        }
        SFB_CUDA(c, cudaMemcpyAsync(P.tbl.p, tbl.data(), tbl.size() * 8, cudaMemcpyHostToDevice, s));
        for (int b = 0; b <= SFB_NBINS; ++b) P.pool_cls[b] = cls_off[(size_t)n_cta * SFB_NBINS + std::min(b, SFB_NBINS)];
        P.pool_cls[SFB_NBINS] = Em;
        P.n_pool = Em - P.pool_cls[0];
        P.pool_nnz = nnz_off[G] - nnz_off[(size_t)n_cta * SFB_NBINS];
        P.max_cta_bytes = max_bytes; P.max_cta_bytes_vb = max_bytes_vb; P.per_sm = per_sm;
        P.usable = (max_bytes + 2048) * per_sm <= (uint64_t)max_optin + 1024 * (uint64_t)(per_sm - 1);
        SFB_CUDA(c, cudaStreamSynchronize(s));
        hm.mark("part: table + sync");
        { const int rc = build_gather(c, tbl); if (rc) return rc; }
        hm.mark("part: gather layout");
        bool marked = false;
        // components too large for a thread go to the pool only when there is a pool anyway (classes that cross ranges): a class
        // set that is local everywhere stays with the on-chip gather loop
        { const int rc = build_dense(c, tbl, pass < 2 && P.n_pool > 0, &marked); if (rc) return rc; }
        hm.mark("part: dense layout");
        if (!marked) break;
    }
    if (getenv("SFB200_VERBOSE"))
        fprintf(stderr, "[sfb200] EM partition: %u CTAs, %llu classes (%llu in the pool, %u pool transcripts), largest CTA slice %llu bytes (limit %d) -> %s\n",
                n_cta, (unsigned long long)Em, (unsigned long long)P.n_pool, P.n_dirty, (unsigned long long)P.max_cta_bytes, max_optin,
                P.dense_ok ? (P.n_pool ? "component threads + pool loop" : "component threads") : P.usable ? "shared-memory loop" : "binned global loop");
    return SFB200_OK;
}

__global__ void k_dirty_list(const uint8_t* __restrict__ dirty, uint32_t T, uint32_t* __restrict__ list, unsigned int* __restrict__ n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T && dirty[t]) list[atomicAdd(n, 1u)] = t;
}

int build_partition(sfb200_ctx* c) {
    DevClasses& k = c->cls;
    DevPartition& P = k.part;
    P.valid = true; P.usable = false; P.gather_ok = false; P.gather_tried = gather_enabled(); P.dense_ok = false; P.dense_tried = dense_enabled();
    P.n_pool_cta = 0;
    if (getenv("SFB200_NO_PARTITION")) return SFB200_OK;
    // CTAs per SM for the partitioned loop: two half-size CTAs fill each other's __syncthreads bubbles
    int per_sm = 2;
    if (const char* e = getenv("SFB200_EM_CTAS_PER_SM")) per_sm = std::max(1, std::min(4, atoi(e)));
    const uint32_t n_full = (uint32_t)(c->num_sms * per_sm);
    if (k.Em == 0 || k.n_txp == 0) return SFB200_OK;
    { const int rc = build_partition_n(c, n_full, per_sm); if (rc) return rc; }
    P.n_dirty = 0;
    if (P.dense_ok && P.n_pool > 0) {
        // hybrid: the pool's transcripts as a list (every CTA looks after a share of them, em_dense.cuh)
        cudaStream_t s = c->stream;
        unsigned int* d_n = reinterpret_cast<unsigned int*>(P.grp.p);
        SFB_CUDA(c, P.dlist.reserve(k.n_txp));
        SFB_CUDA(c, cudaMemsetAsync(d_n, 0, 4, s));
        k_dirty_list<<<grid_for(k.n_txp, 256), 256, 0, s>>>(P.dirty.p, k.n_txp, P.dlist.p, d_n);
        c->launches++;
        unsigned int nd = 0;
        SFB_CUDA(c, cudaMemcpyAsync(&nd, d_n, 4, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        P.n_dirty = nd;
        if (getenv("SFB200_VERBOSE")) fprintf(stderr, "[sfb200] EM hybrid: %llu pool classes over %u transcripts swept by all CTAs\n", (unsigned long long)P.n_pool, nd);
    }
    return SFB200_OK;
}

// Build the gather layout (em_gather.cuh) of the current partition: one k_gather_build CTA per partition range.
// Leaves P.gather_ok false (and the atomic kernels in charge) when the pool is not empty, an index would not fit in
// 16 bits, a region overflows or the largest CTA does not fit in shared memory.
int build_gather(sfb200_ctx* c, const std::vector<unsigned long long>& tbl) {
    DevPartition& P = c->cls.part;
    P.gather_ok = false;
    P.gather_tried = gather_enabled();
    if (!gather_enabled() || P.n_pool != 0 || P.n_cta == 0) return SFB200_OK;
    static_assert(sizeof(GatherGeom) == sizeof(P.gth_geom), "GatherGeom is stored as 16 opaque words");
    uint64_t max_nc = 0, max_ne = 0, max_nt = 0;
    for (uint32_t i = 0; i < P.n_cta; ++i) {
        const unsigned long long* row = tbl.data() + (size_t)i * PT_WORDS;
        max_nc = std::max<uint64_t>(max_nc, row[PT_CLS + SFB_NBINS] - row[PT_CLS]);
        max_ne = std::max<uint64_t>(max_ne, row[PT_ENT1] - row[PT_ENT0]);
        max_nt = std::max<uint64_t>(max_nt, row[PT_TXP1] - row[PT_TXP0]);
    }
    auto up = [](uint64_t x, uint64_t m) { return (x + m - 1) / m * m; };
    if (up(max_nc, 32) > 65535 || up(max_nt, 32) > 65535 || max_ne > (1u << 24)) return SFB200_OK;
    const GatherGeom g = gather_make_geom(max_nc, max_ne, max_nt, !getenv("SFB200_EM_GATHER_UNSCALED"));
    cudaStream_t s = c->stream;
    SFB_CUDA(c, P.gth.reserve((size_t)P.n_cta * g.region_words));
    const size_t scratch = 4 * gather_scratch_words(max_nc, max_nt, g);
    if (scratch + 1024 > P.smem_limit) return SFB200_OK;
    SFB_CUDA(c, cudaFuncSetAttribute(reinterpret_cast<const void*>(&k_gather_build), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
    k_gather_build<<<P.n_cta, 256, scratch, s>>>(P.start.p, P.len.p, P.lab.p, P.tbl.p, g, P.gth.p);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    std::vector<uint32_t> hdr((size_t)P.n_cta * GH_WORDS);
    SFB_CUDA(c, cudaMemcpy2DAsync(hdr.data(), GH_WORDS * 4, P.gth.p, (size_t)g.region_words * 4, GH_WORDS * 4, P.n_cta,
                                  cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    bool ok = true;
    uint64_t need = 0;
    uint32_t max_len = 0, max_deg = 0;
    for (uint32_t i = 0; i < P.n_cta; ++i) {
        const uint32_t* h = hdr.data() + (size_t)i * GH_WORDS;
        ok = ok && h[GH_OK] == 1u;
        need = std::max<uint64_t>(need, gather_smem_need(h[GH_TILES_E], h[GH_TILES_T], h[GH_ENT_E], h[GH_ENT_T]));
        max_len = std::max(max_len, h[GH_MAXLEN]); max_deg = std::max(max_deg, h[GH_MAXDEG]);
    }
    need += 256;
    const bool fits = (need + 2048) * P.per_sm <= P.smem_limit + 1024 * (uint64_t)(P.per_sm - 1);
    std::memcpy(P.gth_geom, &g, sizeof(g));
    P.gather_smem = need;
    P.gather_ok = ok && fits;
    if (getenv("SFB200_VERBOSE"))
        fprintf(stderr, "[sfb200] EM gather layout: largest CTA %llu bytes of shared memory, largest class %u, largest degree %u -> %s\n",
                (unsigned long long)need, max_len, max_deg, P.gather_ok ? "atomic-free loop" : (ok ? "does not fit" : "region overflow"));
    return SFB200_OK;
}

// Build the dense-component layout (em_dense.cuh): usable when every CTA reports components of at most DN_MAX_SLOTS transcripts.
int build_dense(sfb200_ctx* c, const std::vector<unsigned long long>& tbl, bool mark_large, bool* marked) {
    DevPartition& P = c->cls.part;
    P.dense_ok = false;
    P.dense_tried = dense_enabled();
    *marked = false;
    const bool hybrid = hybrid_enabled() && c->coop;
    if (!dense_enabled() || (P.n_pool != 0 && !hybrid) || P.n_cta == 0) return SFB200_OK;
    static_assert(sizeof(DenseGeom) == sizeof(P.dns_geom), "DenseGeom is stored as 16 opaque words");
    uint64_t max_nc = 0, max_nt = 0;
    for (uint32_t i = 0; i < P.n_cta; ++i) {
        const unsigned long long* row = tbl.data() + (size_t)i * PT_WORDS;
        max_nc = std::max<uint64_t>(max_nc, row[PT_CLS + SFB_NBINS] - row[PT_CLS]);
        max_nt = std::max<uint64_t>(max_nt, row[PT_TXP1] - row[PT_TXP0]);
    }
    uint32_t group = 2;                                                // lanes per component (SFB200_EM_DENSE_GROUP = 1, 2, 4; swept on B200);
                                                                       // 0 = balanced by class count (parity green on B200, untimed: opt-in)
    if (const char* e = getenv("SFB200_EM_DENSE_GROUP")) group = (uint32_t)atoi(e);
    const DenseGeom g = dense_make_geom(max_nc, max_nt, group);
    cudaStream_t s = c->stream;
    SFB_CUDA(c, P.dns.reserve((size_t)P.n_cta * g.region_words));
    const size_t scratch = 4 * dense_scratch_words(max_nt, g);
    if (scratch + 1024 > P.smem_limit) return SFB200_OK;
    SFB_CUDA(c, cudaFuncSetAttribute(reinterpret_cast<const void*>(&k_dense_build), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
    k_dense_build<<<P.n_cta, 256, scratch, s>>>(P.start.p, P.len.p, P.lab.p, P.tbl.p, g, P.dns.p, P.dirty.p, (hybrid && mark_large) ? 1 : 0);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    std::vector<uint32_t> hdr((size_t)P.n_cta * DH_WORDS);
    SFB_CUDA(c, cudaMemcpy2DAsync(hdr.data(), DH_WORDS * 4, P.dns.p, (size_t)g.region_words * 4, DH_WORDS * 4, P.n_cta, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    bool ok = true;
    uint32_t ns = 2, rounds = 0;
    for (uint32_t i = 0; i < P.n_cta; ++i) if (hdr[(size_t)i * DH_WORDS + DH_KIND] == 2u) *marked = true;
    if (*marked) return SFB200_OK;                                     // large components were sent to the pool: the caller builds again
    for (uint32_t i = 0; i < P.n_cta && ok; ++i) {
        const uint32_t* h = hdr.data() + (size_t)i * DH_WORDS;
        ok = h[DH_KIND] == 1u;
        if (ok) { ns = std::max(ns, h[DH_NS]); rounds = std::max(rounds, h[DH_ROUNDS]); }
    }
    uint64_t need = 0;
    if (ok) for (uint32_t i = 0; i < P.n_cta; ++i) {
        const uint32_t* h = hdr.data() + (size_t)i * DH_WORDS;
        need = std::max<uint64_t>(need, dense_smem_need(h[DH_TILES], h[DH_ENT], ns, g.group, h[DH_NCOMP]));
    }
    need += 256;
    bool fits = (need + 2048) * P.per_sm <= P.smem_limit + 1024 * (uint64_t)(P.per_sm - 1);
    std::memcpy(P.dns_geom, &g, sizeof(g));
    P.dense_stream = false;
    if (getenv("SFB200_EM_FORCE_STREAM")) fits = false;                 // tests: the streaming variant on class sets that would fit
    if (ok && !fits && g.group == 2 && ns <= DN_MAX_SLOTS && !getenv("SFB200_EM_NO_STREAM")) {
        // the slice of a CTA does not fit: keep only beta and alpha in shared memory, stream counts / base / 1/effLen from a global block
        uint64_t need_s = 0, max_ent = 0, max_state = 0;
        for (uint32_t i = 0; i < P.n_cta; ++i) {
            const uint32_t* h = hdr.data() + (size_t)i * DH_WORDS;
            need_s = std::max<uint64_t>(need_s, dense_smem_need_stream(h[DH_TILES], ns, g.group));
            max_ent = std::max<uint64_t>(max_ent, (h[DH_ENT] + 1u) & ~1u);
            max_state = std::max<uint64_t>(max_state, (uint64_t)ns * (((uint64_t)h[DH_TILES] << 5) / g.group));
        }
        need_s += 256;
        if ((need_s + 2048) * P.per_sm <= P.smem_limit + 1024 * (uint64_t)(P.per_sm - 1)) {
            P.stream_ent = (uint32_t)max_ent; P.stream_state = (uint32_t)max_state;
            SFB_CUDA(c, P.dns_f64.reserve((size_t)P.n_cta * (max_ent + 2 * max_state)));
            P.dense_stream = true; fits = true; need = need_s;
        }
    }
    P.dense_smem = need; P.dense_ns = ns;
    // the alpha ring of the lagged stopping rule (em_dense.cuh), when it fits next to the working set
    P.dense_smem_lag = 0;
    if (ok && fits && !P.dense_stream) {
        uint64_t ring = 0;
        for (uint32_t i = 0; i < P.n_cta; ++i) {
            const uint32_t* h = hdr.data() + (size_t)i * DH_WORDS;
            ring = std::max<uint64_t>(ring, dense_smem_ring(h[DH_TILES], ns, g.group, h[DH_NCOMP]));
        }
        if ((need + ring + 2048) * P.per_sm <= P.smem_limit + 1024 * (uint64_t)(P.per_sm - 1)) P.dense_smem_lag = need + ring;
    }
    P.dense_ok = ok && fits && ns <= DN_MAX_SLOTS && P.per_sm <= 2;
    if (getenv("SFB200_VERBOSE"))
        fprintf(stderr, "[sfb200] EM dense layout: %s (largest component %u transcripts, %u propagation rounds, %llu bytes of shared memory)\n",
                P.dense_ok ? (P.dense_stream ? "one thread per component, counts streamed from global memory" : "one thread per component") : (ok ? "does not fit" : "components too large / not separable"), ns, rounds,
                (unsigned long long)need);
    return SFB200_OK;
}

struct LoopSpec { bool gate_old; uint32_t min_iter; };
enum LoopKind { LOOP_BINNED = 0, LOOP_PART = 1, LOOP_GATHER = 2, LOOP_DENSE = 4 };

// Runs the iteration loop on prepared device state (weights, base, X[0] = alpha_0, X[1] = X[2] = base).
// On return *buf_out says which third of X holds the result.
int run_loop(sfb200_ctx* c, EmParams& p, const sfb200_em_opts* o, LoopKind kind, uint32_t* iters_out, double* mrd_out, unsigned* buf_out) {
    cudaStream_t s = c->stream;
    SFB_CUDA(c, c->em_ctl.reserve(CTL_WORDS));
    SFB_CUDA(c, cudaMemsetAsync(c->em_ctl.p, 0, CTL_WORDS * 8, s));
    p.ctl = c->em_ctl.p;
    const bool vb = o->use_vb != 0;
    const char* mode = getenv("SFB200_EM_MODE");
    const bool sharded = c->n_ranks > 1 && !c->cls.merged;
    const bool steps = sharded || !c->coop || (mode && std::strcmp(mode, "steps") == 0);
    c->last_em_variant = 0;
    unsigned long long h_ctl[CTL_WORDS];
    SFB_CUDA(c, cudaEventRecord(c->ev0, s));
    if (!steps && kind == LOOP_DENSE) {
        const DevPartition& P = c->cls.part;
        DenseParams q;
        q.regions = P.dns.p; std::memcpy(&q.g, P.dns_geom, sizeof(q.g)); q.eff = c->eff.p;
        q.stream_buf = P.dns_f64.p; q.stream_ent = P.stream_ent; q.stream_state = P.stream_state; q.stream_stride = P.stream_ent + 2 * P.stream_state;
        q.dlist = P.dlist.p; q.n_dirty = P.n_pool ? P.n_dirty : 0;
        // the stopping rule (and VBEM's alpha sum) consumed DN_LAG iterations late instead of behind a grid barrier per iteration:
        // whenever an iteration needs a global quantity at all (not EM with a fixed count) and the alpha ring fits
        const char* e_lag = getenv("SFB200_EM_NO_LAG");
        const bool no_lag = e_lag && atoi(e_lag) != 0;
        q.lag = (!no_lag && P.dense_smem_lag && !P.dense_stream && q.g.group == 2 && q.n_dirty == 0 && (vb || o->fixed_iters == 0)) ? 1u : 0u;
        const size_t smem = (size_t)(q.lag ? P.dense_smem_lag : P.dense_smem);
        c->last_em_variant = (P.dense_stream ? 1 : 0) | (q.lag ? 2 : 0);
        void* args[] = {&p, &q};
        const void* fn = nullptr;
#define SFB_DENSE_FN(N, GG) (vb ? reinterpret_cast<const void*>(&k_em_dense<true, N, GG>) : reinterpret_cast<const void*>(&k_em_dense<false, N, GG>))
#define SFB_DENSE_FN_L(N) (vb ? reinterpret_cast<const void*>(&k_em_dense<true, N, 2, false, true>) : reinterpret_cast<const void*>(&k_em_dense<false, N, 2, false, true>))
#define SFB_DENSE_FN_S(N) (vb ? reinterpret_cast<const void*>(&k_em_dense<true, N, 2, true>) : reinterpret_cast<const void*>(&k_em_dense<false, N, 2, true>))
#define SFB_DENSE_CASE(N) case N: fn = P.dense_stream ? SFB_DENSE_FN_S(N) : q.lag ? SFB_DENSE_FN_L(N) : q.g.group == 4 ? SFB_DENSE_FN(N, 4) : q.g.group == 2 ? SFB_DENSE_FN(N, 2) : q.g.group == 0 ? SFB_DENSE_FN(N, 0) : SFB_DENSE_FN(N, 1); break;
        switch (P.dense_ns) { SFB_DENSE_CASE(2) SFB_DENSE_CASE(3) SFB_DENSE_CASE(4) SFB_DENSE_CASE(5) SFB_DENSE_CASE(6) SFB_DENSE_CASE(7) SFB_DENSE_CASE(8)
                              default: SFB_FAIL(c, SFB200_EINVAL, "dense EM: unexpected component size"); }
#undef SFB_DENSE_CASE
#undef SFB_DENSE_FN
#undef SFB_DENSE_FN_S
#undef SFB_DENSE_FN_L
        SFB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SFB_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(P.n_cta), dim3(DENSE_THREADS), args, smem, s));
        c->launches++;
        SFB_CUDA(c, cudaEventRecord(c->ev1, s));
        SFB_CUDA(c, cudaMemcpyAsync(h_ctl, c->em_ctl.p, sizeof(h_ctl), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        *iters_out = static_cast<uint32_t>(h_ctl[CTL_ITERS]);
        *buf_out = static_cast<unsigned>(h_ctl[CTL_RESULT_BUF]);
        const unsigned long long mr = h_ctl[CTL_MRD];
        double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
        *mrd_out = mr ? d : -std::numeric_limits<double>::max();
    } else if (!steps && kind == LOOP_GATHER) {
        const DevPartition& P = c->cls.part;
        GatherParams q;
        q.regions = P.gth.p; std::memcpy(&q.g, P.gth_geom, sizeof(q.g)); q.eff = c->eff.p;
        const size_t smem = (size_t)P.gather_smem;
        void* args[] = {&p, &q};
        const void* fn = q.g.shift == 3
            ? (vb ? reinterpret_cast<const void*>(&k_em_gather<true, 3>) : reinterpret_cast<const void*>(&k_em_gather<false, 3>))
            : (vb ? reinterpret_cast<const void*>(&k_em_gather<true, 0>) : reinterpret_cast<const void*>(&k_em_gather<false, 0>));
        SFB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned per_sm_ctas = P.n_cta / (unsigned)c->num_sms;
        SFB_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(P.n_cta), dim3(EM_THREADS / per_sm_ctas), args, smem, s));
        c->launches++;
        SFB_CUDA(c, cudaEventRecord(c->ev1, s));
        SFB_CUDA(c, cudaMemcpyAsync(h_ctl, c->em_ctl.p, sizeof(h_ctl), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        *iters_out = static_cast<uint32_t>(h_ctl[CTL_ITERS]);
        *buf_out = static_cast<unsigned>(h_ctl[CTL_RESULT_BUF]);
        const unsigned long long mr = h_ctl[CTL_MRD];
        double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
        *mrd_out = mr ? d : -std::numeric_limits<double>::max();
    } else if (!steps && kind == LOOP_PART) {
        const DevPartition& P = c->cls.part;
        PartParams q;
        q.tbl = P.tbl.p; q.dirty = P.dirty.p; q.has_pool = P.n_pool > 0 ? 1 : 0;
        const size_t smem = (size_t)(vb ? P.max_cta_bytes_vb : P.max_cta_bytes) + 1024;
        q.smem_bytes = (uint32_t)smem;
        void* args[] = {&p, &q};
        const void* fn = vb ? reinterpret_cast<const void*>(&k_em_part<true>) : reinterpret_cast<const void*>(&k_em_part<false>);
        SFB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned per_sm_ctas = P.n_cta / (unsigned)c->num_sms;
        SFB_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(P.n_cta), dim3(EM_THREADS / per_sm_ctas), args, smem, s));
        c->launches++;
        SFB_CUDA(c, cudaEventRecord(c->ev1, s));
        SFB_CUDA(c, cudaMemcpyAsync(h_ctl, c->em_ctl.p, sizeof(h_ctl), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        *iters_out = static_cast<uint32_t>(h_ctl[CTL_ITERS]);
        *buf_out = static_cast<unsigned>(h_ctl[CTL_RESULT_BUF]);
        const unsigned long long mr = h_ctl[CTL_MRD];
        double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
        *mrd_out = mr ? d : -std::numeric_limits<double>::max();
    } else if (!steps) {
        void* args[] = {&p};
        const void* fn = vb ? reinterpret_cast<const void*>(&k_em_persistent<true>) : reinterpret_cast<const void*>(&k_em_persistent<false>);
        // shared memory for the CTA's class slice: what the largest slice needs (+ alignment slack), capped by the opt-in maximum
        int max_optin = 0;
        SFB_CUDA(c, cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        // a warp tile of the densest bin holds 16 classes (16 B each) and up to 32 entries (12 B each) = 640 B
        const uint64_t want = (p.tile_start[SFB_NBINS] / c->num_sms + 2) * 640 + 4096;
        const size_t smem = (size_t)std::min<uint64_t>(want, (uint64_t)max_optin - 1024);
        SFB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        p.smem_bytes = (uint32_t)smem;
        int per_sm = 0;
        SFB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, EM_THREADS, smem));
        if (per_sm < 1) SFB_FAIL(c, SFB200_ECUDA, "EM kernel does not fit on an SM");
        SFB_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(c->num_sms), dim3(EM_THREADS), args, smem, s));
        c->launches++;
        SFB_CUDA(c, cudaEventRecord(c->ev1, s));
        SFB_CUDA(c, cudaMemcpyAsync(h_ctl, c->em_ctl.p, sizeof(h_ctl), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        *iters_out = static_cast<uint32_t>(h_ctl[CTL_ITERS]);
        *buf_out = static_cast<unsigned>(h_ctl[CTL_RESULT_BUF]);
        const unsigned long long mr = h_ctl[CTL_MRD];
        double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
        *mrd_out = mr ? d : -std::numeric_limits<double>::max();
    } else {
        // one launch per phase; with a communicator the output buffer is summed over ranks before it is looked at
        const unsigned grid = c->num_sms;
        unsigned bi = 0, bo = 1, bs = 2;
        const bool fixed = o->fixed_iters > 0;
        uint32_t n = 0;
        double asum = p.sum0;
        unsigned long long mr = 0;
        for (;;) {
            const bool last = fixed ? (n >= o->fixed_iters) : (n >= o->max_iter && n >= o->min_iter);
            // slots (n+2)&3 are recycled; clear them (the persistent kernel does this on the device)
            SFB_CUDA(c, cudaMemsetAsync(c->em_ctl.p + CTL_MAXREL + ((n + 2u) & 3u), 0, 8, s));
            SFB_CUDA(c, cudaMemsetAsync(c->em_ctl.p + CTL_CSUM + ((n + 2u) & 3u), 0, 8, s));
            if (vb) k_em_transcript_pass<true><<<grid, EM_THREADS, 0, s>>>(p, bi, bs, n, n > 0, !last, asum);
            else k_em_transcript_pass<false><<<grid, EM_THREADS, 0, s>>>(p, bi, bs, n, n > 0, !last, asum);
            c->launches++;
            if (!last) {
                if (vb) k_em_sweep<true><<<grid, EM_THREADS, 0, s>>>(p, bi, bo, n);
                else k_em_sweep<false><<<grid, EM_THREADS, 0, s>>>(p, bi, bo, n);
                c->launches++;
                if (sharded) {
                    const int rc = sfb_comm_allreduce_f64(c, p.X + (size_t)bo * p.T, p.T);
                    if (rc) return rc;
                    if (vb) {   // the contribution sums are rank-local: recompute the total from the reduced vector
                        double* slot = reinterpret_cast<double*>(c->em_ctl.p + CTL_CSUM + ((n + 1u) & 3u));
                        SFB_CUDA(c, cudaMemsetAsync(slot, 0, 8, s));
                        k_sum_f64<<<grid, EM_THREADS, 0, s>>>(p.X + (size_t)bo * p.T, p.T, slot);
                        c->launches++;
                    }
                }
            }
            const bool need_host = last || (!fixed && n > 0 && n >= o->min_iter) || vb;
            if (need_host) {
                SFB_CUDA(c, cudaMemcpyAsync(h_ctl, c->em_ctl.p, sizeof(h_ctl), cudaMemcpyDeviceToHost, s));
                SFB_CUDA(c, cudaStreamSynchronize(s));
                mr = h_ctl[CTL_MAXREL + (n & 3u)];
                if (vb) {
                    double cs; std::memcpy(&cs, &h_ctl[CTL_CSUM + ((n + 1u) & 3u)], 8);
                    asum = sharded ? cs : p.base_sum + cs;
                }
            }
            if (last) break;
            if (!fixed && n > 0 && n >= o->min_iter) {
                double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
                const double mrd = mr ? d : -std::numeric_limits<double>::max();
                if (!(mrd > o->tol)) break;
            }
            const unsigned tmp = bs; bs = bi; bi = bo; bo = tmp;
            ++n;
        }
        SFB_CUDA(c, cudaGetLastError());
        SFB_CUDA(c, cudaEventRecord(c->ev1, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        *iters_out = n; *buf_out = bi;
        double d; const unsigned long long b = mr ? mr - 1 : 0; std::memcpy(&d, &b, 8);
        *mrd_out = mr ? d : -std::numeric_limits<double>::max();
    }
    float ms = 0.f;
    SFB_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_em_ms = ms;
    c->last_em_kernel = steps ? 3 : (kind == LOOP_DENSE && c->cls.part.n_pool) ? 5 : (int)kind;
    return SFB200_OK;
}

// bias / GC correction inside the optimizer (optimize() :820-840): the model, and where the final effective lengths go
struct EmBias { const sfb200_bias_model* model; double* eff_out; };

__global__ void k_em_restart(const double* __restrict__ base, uint32_t T, double* __restrict__ X) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) { const double b = base[i]; X[(size_t)T + i] = b; X[2 * (size_t)T + i] = b; }
}

// shared by em_run and bootstrap_em: everything from effective lengths to truncated alphas
int em_common(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, double total_frags, const sfb200_em_opts* o,
              const double* d_cnt, const double* d_single, LoopSpec spec, double* alphas_out, uint32_t* iters_out,
              double* mrd_out, const EmBias* eb = nullptr) {
    DevClasses& k = c->cls;
    cudaStream_t s = c->stream;
    const uint32_t T = n_txp;
    HostMarks hm;
    SFB_CUDA(c, c->eff.reserve(2ull * T));
    SFB_CUDA(c, c->em_alpha.reserve(3ull * T));
    SFB_CUDA(c, c->em_theta.reserve(T));
    SFB_CUDA(c, c->em_base.reserve(T));
    double* d_eff_in = c->eff.p + T;
    if (!c->eff_resident) {                                 // a bootstrap run uploads the lengths once, not once per replicate
        SFB_CUDA(c, cudaMemcpyAsync(d_eff_in, eff_lens, T * 8ull, cudaMemcpyHostToDevice, s));
        k_clamp_eff<<<grid_for(T, 256), 256, 0, s>>>(d_eff_in, T, c->eff.p);
        c->launches++;
    }
    // which layout runs: the CTA-partitioned one (em_part.cuh) when a single rank drives a cooperative launch and every
    // CTA's slice fits in shared memory; otherwise the binned layout
    const char* mode_env = getenv("SFB200_EM_MODE");
    const bool sharded = c->n_ranks > 1 && !k.merged;      // rank-local classes: one all-reduce per iteration
    const bool steps_mode = sharded || !c->coop || (mode_env && std::strcmp(mode_env, "steps") == 0);
    bool use_part = false, use_gather = false, use_dense = false;
    hm.mark("em: eff H2D + clamp");
    if (!steps_mode && k.Em) {
        if (!k.part.valid || (gather_enabled() && !k.part.gather_tried) || (dense_enabled() && !k.part.dense_tried)) { const int rc = build_partition(c); if (rc) return rc; }
        use_dense = k.part.dense_ok && dense_enabled();                      // one thread per component (em_dense.cuh)
        use_gather = use_dense || (k.part.gather_ok && gather_enabled());    // atomic-free loop (em_gather.cuh)
        use_part = use_gather || k.part.usable;                              // all read the partition-ordered arrays
        if (use_part && !use_gather && o->use_vb)                   // VBEM keeps expTheta in shared memory as well
            use_part = (k.part.max_cta_bytes_vb + 2048) * k.part.per_sm <= k.part.smem_limit + 1024 * (uint64_t)(k.part.per_sm - 1);
    }
    hm.mark("em: partition (total)");
    DevPartition& P = k.part;
    const uint32_t* a_start = use_part ? P.start.p : k.start.p;
    const uint32_t* a_len = use_part ? P.len.p : k.len.p;
    const uint32_t* a_lab = use_part ? P.lab.p : k.lab.p;
    double* a_w = use_part ? P.w.p : k.w.p;
    const bool hybrid_pool = use_dense && P.n_pool > 0;                 // the pool of a hybrid run is swept in the weighted scatter form
    if (k.Em && (!use_gather || hybrid_pool)) {
        // weights always come from the ORIGINAL counts (the reference computes them once in optimize(), :745-772)
        k_class_weights<<<grid_for(k.Em, 128), 128, 0, s>>>(a_start, a_len, a_lab, use_part ? P.cnt.p : k.cnt.p, c->eff.p, k.Em, a_w);
        c->launches++;
    }
    const double* a_cnt = d_cnt;
    if (use_part) {
        if (d_cnt == k.cnt.p) a_cnt = P.cnt.p;
        else {                                                   // per-sample counts (bootstrap) into partition order
            SFB_CUDA(c, P.cnt_s.reserve(k.Em));
            k_part_gather_counts<<<grid_for(k.Em, 256), 256, 0, s>>>(d_cnt, P.src.p, k.Em, P.cnt_s.p);
            c->launches++;
            a_cnt = P.cnt_s.p;
        }
    }
    uint64_t n_active = k.n_active;
    if (c->n_ranks > 1) {
        // the active set is the union over ranks: sum the 0/1 flags as f64 through the alpha scratch
        // (done by the caller of a multi-rank run through sfb200_em_run below)
    }
    if (n_active == 0) SFB_FAIL(c, SFB200_ENOACTIVE, "The optimizer has no active transcripts: no transcripts are expressed");
    const bool vb = o->use_vb != 0;
    const double alpha0 = (1.0 / static_cast<double>(n_active)) * total_frags;           // :800-803
    const double prior_term = vb ? ((sharded && c->rank != 0) ? 0.0 : o->prior_alpha) : 0.0;
    k_em_init<<<grid_for(T, 256), 256, 0, s>>>(k.active.p, d_single, T, alpha0, prior_term, c->em_alpha.p, c->em_base.p);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());

    EmParams p;
    std::memset(&p, 0, sizeof(p));
    p.start = a_start; p.len = a_len; p.lab = a_lab; p.w = a_w; p.cnt = a_cnt; p.base = c->em_base.p; p.X = c->em_alpha.p;
    p.theta = c->em_theta.p; p.T = T;
    const uint64_t* bin_cls = use_part ? P.pool_cls : k.bin_cls;        // partitioned: the global-memory path sweeps only the pool
    uint64_t tiles = 0;
    for (int b = 0; b < SFB_NBINS; ++b) {
        p.cls_start[b] = bin_cls[b]; p.tile_start[b] = tiles;
        const uint64_t ncls = bin_cls[b + 1] - bin_cls[b];
        const uint64_t per = (b < SFB_NBINS - 1) ? (32u >> (b + 1)) : 1;
        tiles += (ncls + per - 1) / per;
    }
    p.cls_start[SFB_NBINS] = bin_cls[SFB_NBINS]; p.tile_start[SFB_NBINS] = tiles;
    p.use_vb = vb; p.gate_old = spec.gate_old; p.tol = o->tol; p.cutoff = o->check_cutoff;
    p.min_iter = spec.min_iter; p.max_iter = o->max_iter; p.fixed_iters = o->fixed_iters;
    // sums the reference forms by a serial pass over the vector (VBEMUpdate_ :300-303); alpha_0 is n_active equal terms
    double sum0 = 0.0;
    if (vb) for (uint64_t i = 0; i < n_active; ++i) sum0 += alpha0;      // only VBEM's first logNorm reads it
    p.sum0 = sum0;
    double single_sum = 0.0;
    if (vb) {                                               // counts of the single-member classes: integers, any summation order is exact
        SFB_CUDA(c, c->em_ctl.reserve(CTL_WORDS));
        double* d_sum = reinterpret_cast<double*>(c->em_ctl.p + CTL_TSUM);
        SFB_CUDA(c, cudaMemsetAsync(d_sum, 0, 8, s));
        k_sum_f64<<<std::min(grid_for(T, 256), 1024u), 256, 0, s>>>(d_single, T, d_sum);
        c->launches++;
        SFB_CUDA(c, cudaMemcpyAsync(&single_sum, d_sum, 8, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
    }
    p.base_sum = single_sum + static_cast<double>(T) * o->prior_alpha;

    hm.mark("em: init + sums");
    unsigned buf = 0;
    sfb200_em_opts oo = *o;
    oo.min_iter = spec.min_iter;
    const LoopKind kind = use_dense ? LOOP_DENSE : use_gather ? LOOP_GATHER : use_part ? LOOP_PART : LOOP_BINNED;
    if (!eb) {
        const int rc = run_loop(c, p, &oo, kind, iters_out, mrd_out, &buf);
        if (rc) return rc;
    } else {
        // one launch per stretch between the iterations at which the reference recomputes the effective lengths (em_segments.hpp);
        // between two launches: alphas to the host, sfb200_bias_eff_lens, new lengths into c->eff (the gather / dense loops read it
        // at launch, the scatter loops through the class weights, recomputed as updateEqClassWeights does, :527-556)
        static const uint32_t pauses[3] = {50, 500, 1000};
        std::vector<double> h_alpha(T), h_eff(T), h_next(T);
        for (uint32_t i = 0; i < T; ++i) h_eff[i] = eff_lens[i] <= 1.0 ? 1.0 : eff_lens[i];      // what k_clamp_eff left in c->eff
        float loop_ms = 0.f;
        auto update = [&](uint32_t) -> int {
            SFB_CUDA(c, cudaMemcpyAsync(h_alpha.data(), c->em_alpha.p + (size_t)buf * T, T * 8ull, cudaMemcpyDeviceToHost, s));
            SFB_CUDA(c, cudaStreamSynchronize(s));
            const int rc = sfb200_bias_eff_lens(c, eb->model, eff_lens, h_eff.data(), h_alpha.data(), T, h_next.data());
            if (rc) return rc;
            h_eff.swap(h_next);
            SFB_CUDA(c, cudaMemcpyAsync(c->eff.p, h_eff.data(), T * 8ull, cudaMemcpyHostToDevice, s));
            if (k.Em && (!use_gather || hybrid_pool)) {
                k_class_weights<<<grid_for(k.Em, 128), 128, 0, s>>>(a_start, a_len, a_lab, use_part ? P.cnt.p : k.cnt.p, c->eff.p, k.Em, a_w);
                c->launches++;
            }
            // the next launch starts from the current alphas: X[0] = alphas, X[1] = X[2] = base
            if (buf != 0) SFB_CUDA(c, cudaMemcpyAsync(c->em_alpha.p, c->em_alpha.p + (size_t)buf * T, T * 8ull, cudaMemcpyDeviceToDevice, s));
            k_em_restart<<<grid_for(T, 256), 256, 0, s>>>(c->em_base.p, T, c->em_alpha.p);
            c->launches++;
            buf = 0;
            double sum = 0.0;
            if (vb) for (uint32_t i = 0; i < T; ++i) sum += h_alpha[i];                          // VBEMUpdate_'s serial alpha sum (:300-303)
            p.sum0 = sum;
            return SFB200_OK;
        };
        auto run = [&](const sfb::SegLimits& lim, uint32_t, uint32_t* it, double* mrd) -> int {
            sfb200_em_opts so = oo;
            so.min_iter = lim.min_iter; so.max_iter = lim.max_iter; so.fixed_iters = lim.fixed_iters;
            p.min_iter = lim.min_iter; p.max_iter = lim.max_iter; p.fixed_iters = lim.fixed_iters;
            const int rc = run_loop(c, p, &so, kind, it, mrd, &buf);
            loop_ms += c->last_em_ms;
            return rc;
        };
        const int rc = sfb::run_segments(spec.min_iter, o->max_iter, o->fixed_iters, o->tol, pauses, 3, run, update, iters_out, mrd_out);
        if (rc) return rc;
        c->last_em_ms = loop_ms;
        if (eb->eff_out) std::memcpy(eb->eff_out, h_eff.data(), T * 8ull);                        // :888 the lengths quant.sf reports
    }
    hm.mark("em: loop (launch .. sync)");

    const double cutoff = vb ? (o->prior_alpha + o->min_alpha) : o->min_alpha;           // :812
    double alphaSum = 0.0;                                                               // truncateCountVector :37-44, on the device
    k_truncate<<<std::min(grid_for(T, 256), 1024u), 256, 0, s>>>(c->em_alpha.p + (size_t)buf * T, T, cutoff,
                                                                reinterpret_cast<double*>(c->em_ctl.p + CTL_TSUM));
    c->launches++;
    SFB_CUDA(c, cudaMemcpyAsync(alphas_out, c->em_alpha.p + (size_t)buf * T, T * 8ull, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaMemcpyAsync(&alphaSum, c->em_ctl.p + CTL_TSUM, 8, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    hm.mark("em: truncate + alphas D2H");
    if (alphaSum < DENORM_MIN) SFB_FAIL(c, SFB200_ESMALLSUM, "Total alpha weight was too small! Make sure you ran sailfish correctly.");
    return SFB200_OK;
}

}  // namespace

extern "C" int sfb200_em_run(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, uint64_t num_mapped,
                             const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out, double* max_rel_diff_out) {
    if (!c || !eff_lens || !opts || !alphas_out) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "em_run: no classes (call map_finish or eq_import first)");
    if (n_txp != c->cls.n_txp) SFB_FAIL(c, SFB200_EINVAL, "em_run: n_txp differs from the class table's");
    cudaSetDevice(c->device);
    if (c->n_ranks > 1 && !c->cls.merged) {
        // union of the active sets over ranks (classes are rank-local)
        DevClasses& k = c->cls;
        std::vector<uint8_t> act(n_txp);
        std::vector<unsigned long long> cntv(n_txp);
        SFB_CUDA(c, cudaMemcpy(act.data(), k.active.p, n_txp, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < n_txp; ++i) cntv[i] = act[i];
        DevBuf<unsigned long long> d; SFB_CUDA(c, d.reserve(n_txp));
        SFB_CUDA(c, cudaMemcpyAsync(d.p, cntv.data(), n_txp * 8ull, cudaMemcpyHostToDevice, c->stream));
        const int rc = sfb_comm_allreduce_u64(c, d.p, n_txp);
        if (rc) { d.release(); return rc; }
        SFB_CUDA(c, cudaMemcpyAsync(cntv.data(), d.p, n_txp * 8ull, cudaMemcpyDeviceToHost, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
        d.release();
        uint64_t na = 0;
        for (uint32_t i = 0; i < n_txp; ++i) { act[i] = cntv[i] ? 1 : 0; na += act[i]; }
        SFB_CUDA(c, cudaMemcpy(k.active.p, act.data(), n_txp, cudaMemcpyHostToDevice));
        k.n_active = na;
    }
    uint32_t iters = 0; double mrd = 0.0;
    LoopSpec spec{false, opts->min_iter};
    const int rc = em_common(c, eff_lens, n_txp, static_cast<double>(num_mapped), opts, c->cls.cnt.p, c->cls.single.p, spec,
                             alphas_out, &iters, &mrd);
    if (iters_out) *iters_out = iters;
    if (max_rel_diff_out) *max_rel_diff_out = mrd;
    return rc;
}

/* optimize() with --biasCorrect / --gcBiasCorrect: the effective lengths are recomputed from the current abundances at the top of
 * iterations 50, 500 and 1000 (:820-840) and the final ones are returned (:888) */
extern "C" int sfb200_em_run_bias(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, uint64_t num_mapped, const sfb200_em_opts* opts,
                                  const sfb200_bias_model* model, double* alphas_out, double* eff_out, uint32_t* iters_out,
                                  double* max_rel_diff_out) {
    if (!c || !eff_lens || !opts || !alphas_out || !model) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "em_run_bias: no classes (call map_finish or eq_import first)");
    if (n_txp != c->cls.n_txp) SFB_FAIL(c, SFB200_EINVAL, "em_run_bias: n_txp differs from the class table's");
    if (c->n_ranks > 1 && !c->cls.merged) SFB_FAIL(c, SFB200_EINVAL, "em_run_bias: classes must be merged over ranks first (map_finish with the communicator attached)");
    if (!c->index.ready || c->index.n_txp != n_txp) SFB_FAIL(c, SFB200_EINVAL, "em_run_bias: the correction reads the transcript sequences of the index; build or load it first");
    cudaSetDevice(c->device);
    uint32_t iters = 0; double mrd = 0.0;
    LoopSpec spec{false, opts->min_iter};
    EmBias eb{model, eff_out};
    const int rc = em_common(c, eff_lens, n_txp, static_cast<double>(num_mapped), opts, c->cls.cnt.p, c->cls.single.p, spec,
                             alphas_out, &iters, &mrd, &eb);
    if (iters_out) *iters_out = iters;
    if (max_rel_diff_out) *max_rel_diff_out = mrd;
    return rc;
}

// device-resident per-sample counts -> binned counts + single vector, then doBootstrap's loop (:476-514)
int sfb_bootstrap_em_device(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, const unsigned long long* d_samp,
                            uint64_t total, const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out) {
    DevClasses& k = c->cls;
    EmExtra* x = g_extra_for(c);
    cudaStream_t s = c->stream;
    SFB_CUDA(c, x->cnt_s.reserve(k.Em));
    SFB_CUDA(c, x->single_s.reserve(n_txp));
    SFB_CUDA(c, cudaMemsetAsync(x->single_s.p, 0, n_txp * 8ull, s));
    if (k.Em) { k_permute_counts<<<grid_for(k.Em, 256), 256, 0, s>>>(d_samp, k.from_device ? nullptr : k.perm.p, k.Em, x->cnt_s.p); c->launches++; }
    if (k.n_sgl) { k_scatter_single<<<grid_for(k.n_sgl, 256), 256, 0, s>>>(d_samp, k.sgl_cls.p, k.sgl_tid.p, k.n_sgl, x->single_s.p); c->launches++; }
    uint32_t iters = 0; double mrd = 0.0;
    LoopSpec spec{true, 0};
    const int rc = em_common(c, eff_lens, n_txp, static_cast<double>(total), opts, x->cnt_s.p, x->single_s.p, spec, alphas_out,
                             &iters, &mrd);
    if (iters_out) *iters_out = iters;
    return rc;
}

extern "C" int sfb200_bootstrap_em(sfb200_ctx* c, const double* eff_lens, uint32_t n_txp, const uint64_t* samp_counts,
                                   const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out) {
    if (!c || !eff_lens || !opts || !alphas_out || !samp_counts) return SFB200_EINVAL;
    if (!c->cls.ready) SFB_FAIL(c, SFB200_EINVAL, "bootstrap_em: no classes");
    if (n_txp != c->cls.n_txp) SFB_FAIL(c, SFB200_EINVAL, "bootstrap_em: n_txp differs from the class table's");
    cudaSetDevice(c->device);
    EmExtra* x = g_extra_for(c);
    const uint64_t E = c->cls.E;
    SFB_CUDA(c, x->samp.reserve(E));
    uint64_t total = 0;
    for (uint64_t e = 0; e < E; ++e) total += samp_counts[e];
    { const int rc = sfb_classes_host(c); if (rc) return rc; }      // samp_counts come in eq_export order
    std::vector<uint64_t> canon;
    const uint64_t* src = samp_counts;
    if (!c->cls.export_to_canon.empty()) {
        canon.resize(E);
        for (uint64_t e = 0; e < E; ++e) canon[c->cls.export_to_canon[e]] = samp_counts[e];
        src = canon.data();
    }
    SFB_CUDA(c, cudaMemcpyAsync(x->samp.p, src, E * 8, cudaMemcpyHostToDevice, c->stream));
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return sfb_bootstrap_em_device(c, eff_lens, n_txp, x->samp.p, total, opts, alphas_out, iters_out);
}
