// em_gather.cuh -- the atomic-free, weight-free EM loop over the CTA partition (included by em.cu after em_part.cuh).
//
// k_em_part keeps the reference's shape (weighted scatter with f64 atomics into alphaOut) and was measured instruction-bound:
// ~190 thread instructions per label entry (shared-memory f64 atomics are CAS loops, one reciprocal per lane, shuffle
// reductions over half-empty groups).  The update is rewritten so that each pass is a plain segmented sum:
//
//   reference (CollapsedEMOptimizer.cpp:235-277, :760-769):   w_i = (1/eff_i) / sum_j (1/eff_j)      (count cancels)
//       denom_c = sum_i alpha_i w_i,     alphaOut_i += alpha_i w_i count_c / denom_c
//   with beta_i = alpha_i / eff_i the class normaliser cancels:
//       S_c = sum_{i in c} beta_i,  r_c = count_c / S_c            (E-step: one thread per CLASS, sums a column of beta)
//       alphaOut_i = single_i + beta_i * sum_{c containing i} r_c  (M-step: one thread per TRANSCRIPT, sums a column of r)
//   VBEM (:288-369): beta_i = expTheta_i / eff_i and alphaOut_i additionally starts from the prior.
//
// Same fixed point and same iterates up to fp64 rounding (the tolerance of the path is 1e-4 relative); no weight array,
// no atomics, no shuffles; the new alpha is produced by the thread that holds the old one, so the reference's convergence
// test (:849-861 / :496-508) is evaluated in the same pass and the verdict needs no speculative extra sweep.
// The layout (em_gather_build.inl) lives in global memory, one region per CTA, built once per class set by k_gather_build
// and staged into shared memory with TMA bulk copies at kernel start.  Requires an empty pool (every class local to a CTA).

#define SFB_GB_FN __device__ __forceinline__
#define SFB_GB_TID threadIdx.x
#define SFB_GB_NT blockDim.x
#define SFB_GB_SYNC() __syncthreads()
#define SFB_GB_ADD(p, v) atomicAdd((p), (v))
#define SFB_GB_MAX(p, v) atomicMax((p), (v))
#include "em_gather_build.inl"
#undef SFB_GB_FN
#undef SFB_GB_TID
#undef SFB_GB_NT
#undef SFB_GB_SYNC
#undef SFB_GB_ADD
#undef SFB_GB_MAX

__global__ void __launch_bounds__(256) k_gather_build(const uint32_t* __restrict__ start, const uint32_t* __restrict__ len,
                                                      const uint32_t* __restrict__ lab, const unsigned long long* __restrict__ tbl,
                                                      const GatherGeom g, uint32_t* __restrict__ regions) {
    extern __shared__ __align__(16) uint32_t gb_scratch[];
    const unsigned long long* row = tbl + (size_t)blockIdx.x * PT_WORDS;
    const uint32_t c_lo = (uint32_t)row[PT_CLS], nc = (uint32_t)(row[PT_CLS + SFB_NBINS] - row[PT_CLS]);
    const uint32_t t0 = (uint32_t)row[PT_TXP0], nt = (uint32_t)(row[PT_TXP1] - row[PT_TXP0]);
    gather_build_cta(start, len, lab, c_lo, nc, t0, nt, g, regions + (size_t)blockIdx.x * g.region_words, gb_scratch);
}

struct GatherParams {
    const uint32_t* regions;
    GatherGeom g;
    const double* eff;        // T clamped effective lengths
};

__device__ __forceinline__ uint32_t up4(uint32_t x) { return (x + 3u) & ~3u; }
__device__ __forceinline__ uint32_t up8(uint32_t x) { return (x + 7u) & ~7u; }

// shared memory a CTA of k_em_gather needs for (nc, nt, tiles, entries): mirrored on the host (gather_smem_bytes)
__host__ __device__ inline uint64_t gather_smem_need(uint32_t tiles_e, uint32_t tiles_t, uint32_t ent_e, uint32_t ent_t) {
    const uint64_t nc_pad = (uint64_t)tiles_e << 5, nt_pad = (uint64_t)tiles_t << 5;
    const uint64_t f64s = (nc_pad + 2) * 2 + (nt_pad + 2) * 4;                       // r, cnt | beta, alpha, base, inveff (even counts)
    const uint64_t u32s = 2 * (uint64_t)((tiles_e + 3u) & ~3u) + 2 * (uint64_t)((tiles_t + 3u) & ~3u);
    const uint64_t u16s = (uint64_t)((ent_e + 7u) & ~7u) + (uint64_t)((ent_t + 7u) & ~7u);
    return f64s * 8 + u32s * 4 + u16s * 2;
}

// sum of v[col[32 j]] for j < L.  L is uniform over the warp and small (a class has ~2-5 members, a transcript ~5-15 classes),
// and the per-element loop overhead was measured at ~45% of the kernel's instructions: dispatch once on L to a fully unrolled
// body (independent loads issued back to back, two accumulators), generic loop beyond 16
// SHIFT = 3: the stored value is already the byte offset of the f64 (no index scaling per element)
template <int SHIFT>
__device__ __forceinline__ double col_at(const double* __restrict__ v, uint32_t stored) {
    return SHIFT == 3 ? *reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(v) + stored) : v[stored];
}
template <int L, int SHIFT>
__device__ __forceinline__ double col_sum_fixed(const uint16_t* __restrict__ col, const double* __restrict__ v) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < L; ++j) { const double x = col_at<SHIFT>(v, col[j << 5]); if (j & 1) s1 += x; else s0 += x; }
    return s0 + s1;
}
template <int SHIFT>
__device__ __forceinline__ double col_sum(const uint16_t* __restrict__ col, const double* __restrict__ v, uint32_t L) {
    switch (L) {
        case 0: return 0.0;
        case 1: return col_sum_fixed<1, SHIFT>(col, v);
        case 2: return col_sum_fixed<2, SHIFT>(col, v);
        case 3: return col_sum_fixed<3, SHIFT>(col, v);
        case 4: return col_sum_fixed<4, SHIFT>(col, v);
        case 5: return col_sum_fixed<5, SHIFT>(col, v);
        case 6: return col_sum_fixed<6, SHIFT>(col, v);
        case 7: return col_sum_fixed<7, SHIFT>(col, v);
        case 8: return col_sum_fixed<8, SHIFT>(col, v);
        case 9: return col_sum_fixed<9, SHIFT>(col, v);
        case 10: return col_sum_fixed<10, SHIFT>(col, v);
        case 11: return col_sum_fixed<11, SHIFT>(col, v);
        case 12: return col_sum_fixed<12, SHIFT>(col, v);
        case 13: return col_sum_fixed<13, SHIFT>(col, v);
        case 14: return col_sum_fixed<14, SHIFT>(col, v);
        case 15: return col_sum_fixed<15, SHIFT>(col, v);
        case 16: return col_sum_fixed<16, SHIFT>(col, v);
        default: break;
    }
    double s0 = 0.0, s1 = 0.0;
    uint32_t j = 0;
    for (; j + 4 <= L; j += 4) {
        const uint32_t i0 = col[j << 5], i1 = col[(j + 1) << 5], i2 = col[(j + 2) << 5], i3 = col[(j + 3) << 5];
        const double x0 = col_at<SHIFT>(v, i0), x1 = col_at<SHIFT>(v, i1), x2 = col_at<SHIFT>(v, i2), x3 = col_at<SHIFT>(v, i3);
        s0 += x0; s1 += x1; s0 += x2; s1 += x3;
    }
    for (; j < L; ++j) s0 += col_at<SHIFT>(v, col[j << 5]);
    return s0 + s1;
}

// count / S for the E-step: hardware reciprocal seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps (<= 2 ulp; the path's
// tolerance is 1e-4) while S is comfortably normal.  Branch-free for the two cases that occur -- a normal S and S == 0 (padding
// entries, classes whose members all vanished: ratio 0, :260) -- and an out-of-line exact quotient for a denormal or huge S,
// so that a warp with some zero denominators does not walk through a division.
__device__ __noinline__ double em_ratio_rare(double cnt, double S) { return (S > DENORM_MIN) ? cnt / S : 0.0; }
__device__ __forceinline__ double em_ratio(double cnt, double S) {
    const bool normal = S > 1e-280 && S < 1e280;
    const double Ss = normal ? S : 1.0;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(Ss));
    double e = fma(-Ss, y, 1.0);
    y = fma(y, e, y);
    e = fma(-Ss, y, 1.0);
    y = fma(y, e, y);
    double r = normal ? cnt * y : 0.0;
    if (!normal && S != 0.0) r = em_ratio_rare(cnt, S);
    return r;
}

template <bool VB, int SHIFT>
__global__ void __launch_bounds__(EM_THREADS, 1) k_em_gather(const EmParams p, const GatherParams q) {
    __shared__ unsigned long long sm_u[32];
    __shared__ double sm_d[32];
    __shared__ uint64_t tma_bar;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const unsigned nblocks = gridDim.x;
    unsigned long long gen = 0;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, W = blockDim.x >> 5;

    // ---- this CTA's region
    const uint32_t* region = q.regions + (size_t)blockIdx.x * q.g.region_words;
    const uint32_t nc = region[GH_NC], nt = region[GH_NT], tiles_e = region[GH_TILES_E], tiles_t = region[GH_TILES_T];
    const uint32_t ent_e = region[GH_ENT_E], ent_t = region[GH_ENT_T];
    const uint32_t nc_pad = tiles_e << 5, nt_pad = tiles_t << 5;
    double* s_r = reinterpret_cast<double*>(dyn_smem);          // nc_pad + 1 (sentinel) (+1 to stay even)
    double* s_cnt = s_r + nc_pad + 2;                            // nc_pad (+2)
    double* s_beta = s_cnt + nc_pad + 2;                         // nt_pad + 1 (sentinel)
    double* s_alpha = s_beta + nt_pad + 2;
    double* s_base = s_alpha + nt_pad + 2;
    double* s_inveff = s_base + nt_pad + 2;
    uint32_t* s_eoff = reinterpret_cast<uint32_t*>(s_inveff + nt_pad + 2);
    uint32_t* s_elen = s_eoff + up4(tiles_e);
    uint32_t* s_toff = s_elen + up4(tiles_e);
    uint32_t* s_tlen = s_toff + up4(tiles_t);
    uint16_t* s_lab_e = reinterpret_cast<uint16_t*>(s_tlen + up4(tiles_t));
    uint16_t* s_cls_t = s_lab_e + up8(ent_e);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&tma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t b_te = up4(tiles_e) * 4u, b_tt = up4(tiles_t) * 4u, b_ee = up8(ent_e) * 2u, b_et = up8(ent_t) * 2u;
    const uint32_t tx_bytes = 2u * b_te + 2u * b_tt + b_ee + b_et;
    if (tx_bytes && threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&tma_bar)), "r"(tx_bytes) : "memory");
        if (b_te) { tma_load_1d(s_eoff, region + q.g.o_tile_e_off, b_te, &tma_bar); tma_load_1d(s_elen, region + q.g.o_tile_e_len, b_te, &tma_bar); }
        if (b_tt) { tma_load_1d(s_toff, region + q.g.o_tile_t_off, b_tt, &tma_bar); tma_load_1d(s_tlen, region + q.g.o_tile_t_len, b_tt, &tma_bar); }
        if (b_ee) tma_load_1d(s_lab_e, region + q.g.o_lab_e, b_ee, &tma_bar);
        if (b_et) tma_load_1d(s_cls_t, region + q.g.o_cls_t, b_et, &tma_bar);
    }
    // per-run vectors (counts, alpha_0, base, 1/eff) are gathered through the index maps while the bulk copies fly
    const uint32_t* cperm = region + q.g.o_cperm;
    const uint32_t* tmap = region + q.g.o_tmap;
    for (uint32_t i = threadIdx.x; i < nc_pad; i += blockDim.x) { s_cnt[i] = i < nc ? p.cnt[cperm[i]] : 0.0; s_r[i] = 0.0; }
    for (uint32_t i = threadIdx.x; i < nt_pad; i += blockDim.x) {
        double a = 0.0, b = 0.0, ie = 0.0;
        if (i < nt) { const uint32_t t = tmap[i]; a = p.X[t]; b = p.base[t]; ie = 1.0 / q.eff[t]; }
        s_alpha[i] = a; s_base[i] = b; s_inveff[i] = ie;
    }
    if (threadIdx.x == 0) { s_r[nc_pad] = 0.0; s_beta[nt_pad] = 0.0; }
    if (tx_bytes) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(done) : "r"(smem_u32(&tma_bar)) : "memory");
        }
    }
    __syncthreads();

    const bool fixed = p.fixed_iters > 0;
    // beta of alpha_0
    {
        const double logNorm = VB ? sfb_digamma(p.sum0) : 0.0, thetaScale = VB ? exp(-logNorm) : 0.0;
        for (uint32_t i = threadIdx.x; i < nt_pad; i += blockDim.x) {
            const double a = s_alpha[i];
            const double th = VB ? ((a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0) : a;
            s_beta[i] = th * s_inveff[i];
        }
    }
    __syncthreads();
    uint32_t n = 0;                                                 // completed iterations (the reference's itNum)
    unsigned long long mr_final = 0ULL;
    for (;;) {
        if (fixed ? (n >= p.fixed_iters) : (n >= p.max_iter && n >= p.min_iter)) break;
        // ---- E-step: r_c = count_c / sum of beta over the class's column
        for (uint32_t k0 = 0, rnd = 0; k0 < tiles_e; k0 += W, ++rnd) {
            const uint32_t k = k0 + ((rnd & 1u) ? W - 1u - warp : warp);      // tiles are sorted by size: serpentine deal
            if (k >= tiles_e) continue;
            const double S = col_sum<SHIFT>(s_lab_e + s_eoff[k] + lane, s_beta, s_elen[k]);
            const uint32_t c = (k << 5) + lane;
            s_r[c] = em_ratio(s_cnt[c], S);
        }
        __syncthreads();
        // ---- M-step + the convergence test of this iteration
        const uint32_t m = n + 1;
        const bool do_cmp = fixed ? (m >= p.fixed_iters) : (m >= p.min_iter);
        unsigned long long best = 0ULL;
        double asum = 0.0;
        for (uint32_t k0 = 0, rnd = 0; k0 < tiles_t; k0 += W, ++rnd) {
            const uint32_t k = k0 + ((rnd & 1u) ? W - 1u - warp : warp);
            if (k >= tiles_t) continue;
            const double acc = col_sum<SHIFT>(s_cls_t + s_toff[k] + lane, s_r, s_tlen[k]);
            const uint32_t i = (k << 5) + lane;
            const double a_old = s_alpha[i];
            const double a_new = s_beta[i] * acc + s_base[i];
            if (do_cmp) {
                const double gate = p.gate_old ? a_old : a_new;
                if (gate > p.cutoff) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a_old - a_new) / a_new) + 1ULL;
                    best = bits > best ? bits : best;
                }
            }
            s_alpha[i] = a_new;
            if (VB) asum += a_new; else s_beta[i] = a_new * s_inveff[i];
        }
        n = m;
        if (VB || do_cmp) {
            unsigned long long* slot = p.ctl + CTL_MAXREL + (m & 3u);
            double* csum = reinterpret_cast<double*>(p.ctl + CTL_CSUM + (m & 3u));
            if (do_cmp) block_max_to_slot(best, slot, sm_u);
            if (VB) block_sum_to_slot(asum, csum, sm_d);
            grid_barrier(p.ctl, nblocks, gen);
            // slots of iteration m-1 were last read before this barrier; they serve iteration m+3 next
            if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_MAXREL + ((m + 3u) & 3u)] = 0ULL; p.ctl[CTL_CSUM + ((m + 3u) & 3u)] = 0ULL; }
            if (do_cmp) {
                mr_final = ld_cg_u64(slot);
                if (fixed) break;                                    // m == fixed_iters
                if (!(decode_mrd(mr_final) > p.tol) || m >= p.max_iter) break;
            }
            if (VB) {
                const double logNorm = sfb_digamma(__longlong_as_double((long long)ld_cg_u64(p.ctl + CTL_CSUM + (m & 3u))));
                const double thetaScale = exp(-logNorm);
                for (uint32_t i = threadIdx.x; i < nt_pad; i += blockDim.x) {
                    const double a = s_alpha[i];
                    s_beta[i] = ((a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0) * s_inveff[i];
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = 0ULL; p.ctl[CTL_MRD] = mr_final; }
    for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) p.X[tmap[i]] = s_alpha[i];
}
