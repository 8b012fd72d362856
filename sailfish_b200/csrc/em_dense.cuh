// em_dense.cuh -- the EM loop for class structures that fall apart into small connected components (included by em.cu after
// em_gather.cuh).
//
// k_em_gather was measured bound by shared-memory wavefronts: ~4000 per SM and iteration at cfg2, two thirds of them bank
// conflicts of the f64 gathers (32 lanes read 32 unrelated transcripts / classes).  But the structure is far more local than a
// CTA range: the isoforms of a gene and the classes over them form a connected component of a handful of transcripts.  When no
// component has more than 8 transcripts (em_dense_build.inl checks that on the device), ONE THREAD owns a component:
//   * its transcripts are slots 0..NS-1 and live in the thread's registers for the whole iteration (beta_j, accumulators);
//   * a class is (count, slot mask): S_c = sum of beta over the mask, r_c = count_c / S_c, accumulate r_c into the masked slots
//     (E-step and M-step of the component back to back, same arithmetic as k_em_gather);
//   * everything a thread reads per class is one f64 and one byte, column-major over the 32 components of a warp tile:
//     conflict-free, no gathers, no atomics, no shuffles and -- components being independent -- NO barrier inside an iteration.
// CTA barriers and the grid barrier appear only where the reference needs a global quantity (the stopping rule, VBEM's
// digamma(sum alpha)), exactly as in k_em_gather.  Transcripts outside every multi-member class ("idle") are constant after the
// first iteration (alpha = count of their single-member class [+ prior]) and are only touched for that comparison and the result.

#define SFB_GB_FN __device__ __forceinline__
#define SFB_GB_TID threadIdx.x
#define SFB_GB_NT blockDim.x
#define SFB_GB_SYNC() __syncthreads()
#define SFB_GB_ADD(p, v) atomicAdd((p), (v))
#define SFB_GB_MAX(p, v) atomicMax((p), (v))
#define SFB_GB_MIN(p, v) atomicMin((p), (v))
#include "em_dense_build.inl"
#undef SFB_GB_FN
#undef SFB_GB_TID
#undef SFB_GB_NT
#undef SFB_GB_SYNC
#undef SFB_GB_ADD
#undef SFB_GB_MAX
#undef SFB_GB_MIN

__global__ void __launch_bounds__(256) k_dense_build(const uint32_t* __restrict__ start, const uint32_t* __restrict__ len,
                                                     const uint32_t* __restrict__ lab, const unsigned long long* __restrict__ tbl,
                                                     const DenseGeom g, uint32_t* __restrict__ regions, uint8_t* dirty, int mark_large) {
    extern __shared__ __align__(16) uint32_t db_scratch[];
    const unsigned long long* row = tbl + (size_t)blockIdx.x * PT_WORDS;
    const uint32_t c_lo = (uint32_t)row[PT_CLS], nc = (uint32_t)(row[PT_CLS + SFB_NBINS] - row[PT_CLS]);
    const uint32_t t0 = (uint32_t)row[PT_TXP0], nt = (uint32_t)(row[PT_TXP1] - row[PT_TXP0]);
    dense_build_cta(start, len, lab, c_lo, nc, t0, nt, g, regions + (size_t)blockIdx.x * g.region_words, db_scratch, dirty, mark_large ? dirty : nullptr);
}

struct DenseParams {
    const uint32_t* regions;
    DenseGeom g;
    const double* eff;        // T clamped effective lengths
    // streaming runs (the per-CTA working set does not fit in shared memory: metatranscriptome scale): class counts, base and 1/effLen
    // live in a per-CTA global block (stream_buf + blockIdx.x * stream_stride: [cnt: stream_ent][base: stream_state][1/eff: stream_state])
    // and the masks are read from the region; only beta and alpha stay in shared memory
    double* stream_buf; uint32_t stream_stride, stream_ent, stream_state;
    // hybrid runs: the classes the components do not cover -- components too large for a thread, classes that cross CTA ranges: the
    // "pool", p.start/len/lab/w/cnt binned as k_em_part's pool -- are swept by ALL CTAs in the scatter form every iteration
    const uint32_t* dlist;    // the pool's transcripts
    uint32_t n_dirty;         // 0: no pool
    // lagged stopping rule (see k_em_dense): 0 = the global quantities of an iteration are known before the next one starts (a grid
    // barrier per iteration), 1 = they are consumed DN_LAG iterations later
    uint32_t lag;
};

// Lagged stopping rule.  The components of a CTA do not depend on other CTAs; what ties the grid together is the reference's stopping
// rule (max relative change over ALL transcripts, every iteration from minIter on: CollapsedEMOptimizer.cpp:849-861) and VBEM's
// digamma(sum alpha).  A grid barrier per iteration for them costs more than the iteration (measured: 10.2 us per iteration to
// convergence against 2.6 us with a fixed count, VBEM 13 us): every CTA waits for the slowest one, then for two L2 round trips.
// Instead a CTA publishes its part of iteration m (atomicMax / atomicAdd into slot m % 16, then an arrival count) and goes on; the
// decision for iteration m is read DN_LAG iterations later, when every CTA has long arrived.  The alphas of the last DN_LAG + 1
// iterations stay in a shared-memory ring, so when iteration x turns out to have met the rule, alpha_x is what the run returns:
// same iteration count, same numbers as the synchronous loop.  VBEM's expTheta takes digamma(sum alpha) of DN_LAG iterations ago:
// the sum is the same number up to rounding (every iteration redistributes the same counts) and scales ALL expThetas alike, which
// cancels in every class's shares.
// The arrival for iteration m is sent one iteration later (its atomics are performed by then, so the fence before it is free), and
// what is due at iteration m is loaded before m's sweep and looked at after it: neither costs a round trip on the critical path.
// Slot reuse: CTAs are at most DN_LAG iterations apart (nobody passes m + DN_LAG before everybody has arrived at m), so the slots of
// iterations m - 2 DN_LAG .. m + DN_LAG may be live when CTA 0 is at m; it clears the slot of m + DN_LAG + 1 (last used 16 iterations
// earlier) before its own arrival at m.  Arrival counters only grow.
constexpr uint32_t DN_LAG = 3, DN_RING = 4, DN_LAG_SLOTS = 16;
static_assert(DN_RING == DN_LAG + 1 && (DN_RING & (DN_RING - 1)) == 0 && DN_LAG_SLOTS > 3 * DN_LAG + 1, "lagged stopping rule geometry");

constexpr int DENSE_THREADS = 256;
constexpr int DENSE_ILP = 4;          // classes of one component in flight per lane

template <int N>
__device__ __forceinline__ double dense_tree_sum(const double* v) {
    if constexpr (N == 1) return v[0];
    else return dense_tree_sum<N / 2>(v) + dense_tree_sum<N - N / 2>(v + N / 2);
}

// shared memory a CTA of k_em_dense needs (mirrored on the host)
__host__ __device__ inline uint64_t dense_smem_need_stream(uint32_t tiles, uint32_t ns, uint32_t group) {
    const uint64_t ncomp_pad = ((uint64_t)tiles << 5) / (group ? group : 1);
    return 2 * (uint64_t)ns * ncomp_pad * 8 + 2 * (uint64_t)((tiles + 3u) & ~3u) * 4;
}
__host__ __device__ inline uint64_t dense_smem_need(uint32_t tiles, uint32_t ent, uint32_t ns, uint32_t group, uint32_t ncomp) {
    const uint64_t ncomp_pad = group ? ((uint64_t)tiles << 5) / group : (((uint64_t)ncomp + 31u) & ~31ull);
    return (uint64_t)((ent + 1u) & ~1u) * 8 + 4 * (uint64_t)ns * ncomp_pad * 8 + 2 * (uint64_t)((tiles + 3u) & ~3u) * 4 + (uint64_t)((ent + 15u) & ~15u) +
           (group ? 0 : (uint64_t)tiles * 32 * 4);
}

// the alpha ring of the lagged stopping rule: DN_RING - 1 more copies of the alpha array
__host__ __device__ inline uint64_t dense_smem_ring(uint32_t tiles, uint32_t ns, uint32_t group, uint32_t ncomp) {
    const uint64_t ncomp_pad = group ? ((uint64_t)tiles << 5) / group : (((uint64_t)ncomp + 31u) & ~31ull);
    return (uint64_t)(DN_RING - 1) * ns * ncomp_pad * 8 + 16;
}

// G lanes share a component: each takes every G-th class of it (all G hold the component's beta), the accumulators are summed
// over the group with shuffles, lane 0 of the group writes the component's new state.  The longest class list of a tile sets
// the pace of its warp, and nothing else runs on that warp: G = 4 shortens that list fourfold for 10 shuffles per slot.
// STREAM: the class counts, base and 1/effLen of the CTA are read from global memory every iteration (coalesced 256-byte rows; the whole
// run's stream is a few tens of MB and stays in L2) instead of shared memory -- for class sets whose per-CTA slice does not fit.
template <bool VB, int NS, int G, bool STREAM = false, bool LAGGED = false>
__global__ void __launch_bounds__(DENSE_THREADS, 2) k_em_dense(const EmParams p, const DenseParams q) {
    __shared__ unsigned long long sm_u[32];
    __shared__ double sm_d[32];
    __shared__ uint64_t tma_bar;
    __shared__ unsigned long long s_pmax[2], s_res[4];          // lagged rule: the CTA's max of iteration m (by parity), what is due
    __shared__ double s_psum[2];                                 //              the CTA's alpha sum of iteration m (by parity)
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const unsigned nblocks = gridDim.x;
    unsigned long long gen = 0;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    if (threadIdx.x < 2) { s_pmax[threadIdx.x] = 0ULL; s_psum[threadIdx.x] = 0.0; }

    const uint32_t* region = q.regions + (size_t)blockIdx.x * q.g.region_words;
    const uint32_t tiles = region[DH_TILES], ent = region[DH_ENT], nidle = region[DH_NIDLE];
    const uint32_t ncomp_pad = G ? (tiles << 5) / G : ((region[DH_NCOMP] + 31u) & ~31u);      // G == 0: balanced layout, lanes per component vary
    double* gbuf = STREAM ? q.stream_buf + (size_t)blockIdx.x * q.stream_stride : nullptr;
    double* s_cnt = STREAM ? gbuf : reinterpret_cast<double*>(dyn_smem);         // ent (even)
    double* s_beta = STREAM ? reinterpret_cast<double*>(dyn_smem) : s_cnt + ((ent + 1u) & ~1u);   // [NS][ncomp_pad] each
    double* s_alpha = s_beta + (size_t)NS * ncomp_pad;
    double* s_base = STREAM ? gbuf + q.stream_ent : s_alpha + (size_t)NS * ncomp_pad;
    double* s_inveff = STREAM ? s_base + q.stream_state : s_base + (size_t)NS * ncomp_pad;
    uint32_t* s_toff = reinterpret_cast<uint32_t*>(STREAM ? s_alpha + (size_t)NS * ncomp_pad : s_inveff + (size_t)NS * ncomp_pad);
    uint32_t* s_tlen = s_toff + ((tiles + 3u) & ~3u);
    uint8_t* s_mask = STREAM ? const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(region + q.g.o_mask))
                             : reinterpret_cast<uint8_t*>(s_tlen + ((tiles + 3u) & ~3u));          // ent (padded to 16)
    uint32_t* s_lane = reinterpret_cast<uint32_t*>(s_mask + ((ent + 15u) & ~15u));   // G == 0: 32 * tiles lane descriptors (never with STREAM)
    constexpr bool lagged = LAGGED;                                                   // host (q.lag): never with STREAM or a pool
    double* s_ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_lane + (G ? 0u : tiles * 32u)) + 15u) & ~(uintptr_t)15u);
    const size_t ring_stride = (size_t)NS * ncomp_pad;
    // alphas after iteration `it` (lagged runs): slot it % DN_RING of the ring, slot 0 being s_alpha itself
    auto ring = [&](uint32_t it) -> double* { const uint32_t h = it & (DN_RING - 1u); return h == 0u ? s_alpha : s_ring + (size_t)(h - 1u) * ring_stride; };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&tma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t b_t = ((tiles + 3u) & ~3u) * 4u, b_m = STREAM ? 0u : ((ent + 15u) & ~15u), b_l = G ? 0u : tiles * 128u;
    const uint32_t tx_bytes = 2u * b_t + b_m + b_l;
    if (tx_bytes && threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&tma_bar)), "r"(tx_bytes) : "memory");
        if (b_t) { tma_load_1d(s_toff, region + q.g.o_tile_off, b_t, &tma_bar); tma_load_1d(s_tlen, region + q.g.o_tile_len, b_t, &tma_bar); }
        if (b_m) tma_load_1d(s_mask, region + q.g.o_mask, b_m, &tma_bar);
        if (b_l) tma_load_1d(s_lane, region + q.g.o_lane, b_l, &tma_bar);
    }
    // per-run vectors through the index maps while the bulk copies fly
    const uint32_t* cperm = region + q.g.o_cperm;
    const uint32_t* tmap = region + q.g.o_tmap;
    const uint32_t* idle = region + q.g.o_idle;
    for (uint32_t i = threadIdx.x; i < ent; i += blockDim.x) { const uint32_t c = cperm[i]; s_cnt[i] = c != DN_NONE ? p.cnt[c] : 0.0; }
    for (uint32_t i = threadIdx.x; i < (uint32_t)NS * ncomp_pad; i += blockDim.x) {
        // the region stores [slot][component] for DN_MAX_SLOTS slots; a run with NS < DN_MAX_SLOTS reads the first NS rows
        const uint32_t t = tmap[i];
        double a = 0.0, b = 0.0, ie = 0.0;
        if (t != DN_NONE) { a = p.X[t]; b = p.base[t]; ie = 1.0 / q.eff[t]; }
        s_alpha[i] = a; s_base[i] = b; s_inveff[i] = ie;
    }
    if (lagged) for (uint32_t i = threadIdx.x; i < (DN_RING - 1u) * (uint32_t)ring_stride; i += blockDim.x) s_ring[i] = 0.0;
    // idle transcripts: constant from the first iteration on; their sum feeds VBEM's alpha sum
    double idle_sum = 0.0;
    if (VB) for (uint32_t i = threadIdx.x; i < nidle; i += blockDim.x) idle_sum += p.base[idle[i]];
    if (tx_bytes) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(done) : "r"(smem_u32(&tma_bar)) : "memory");
        }
    }
    __syncthreads();

    const bool fixed = p.fixed_iters > 0;
    // ---- the pool (hybrid runs): its transcripts keep their alphas in the three rotating global buffers of k_em_persistent (in / out /
    // spare), every CTA sweeps its share of the pool's tiles with gathers and red.add through L2 and looks after its share of the pool's
    // transcripts; the pool costs one grid barrier per iteration (sweep complete) -- the price of classes that tie CTAs together
    const bool has_pool = q.n_dirty > 0;
    const Bins pb = em_bins(p);
    const uint64_t pool_tiles = p.tile_start[SFB_NBINS];
    const uint64_t ptile_lo = pool_tiles * blockIdx.x / nblocks, ptile_hi = pool_tiles * (blockIdx.x + 1ULL) / nblocks;
    const uint32_t d_lo = (uint32_t)((uint64_t)q.n_dirty * blockIdx.x / nblocks), d_hi = (uint32_t)((uint64_t)q.n_dirty * (blockIdx.x + 1ULL) / nblocks);
    Slice sl_pool; sl_pool.start = p.start; sl_pool.len = p.len; sl_pool.cnt = p.cnt; sl_pool.lab = p.lab; sl_pool.w = p.w; sl_pool.c0 = 0; sl_pool.e0 = 0;
    unsigned bi = 0, bo = 1, bs = 2;
    if (VB && has_pool) {
        const double logNorm = sfb_digamma(p.sum0), thetaScale = exp(-logNorm);
        for (uint32_t i = d_lo + threadIdx.x; i < d_hi; i += blockDim.x) {
            const uint32_t t = q.dlist[i];
            const double a = p.X[t];
            p.theta[t] = (a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0;
        }
        grid_barrier(p.ctl, nblocks, gen);
    }
    {
        const double logNorm = VB ? sfb_digamma(p.sum0) : 0.0, thetaScale = VB ? exp(-logNorm) : 0.0;
        for (uint32_t i = threadIdx.x; i < (uint32_t)NS * ncomp_pad; i += blockDim.x) {
            const double a = s_alpha[i];
            const double th = VB ? ((a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0) : a;
            s_beta[i] = th * s_inveff[i];
        }
    }
    __syncthreads();
    uint32_t n = 0;
    unsigned long long mr_final = 0ULL;
    // the idle transcripts are compared once (m == 1); afterwards their alpha no longer changes and all that is left of them in the
    // stopping rule is "a transcript past the gate with relative change 0"
    unsigned long long idle_later = 0ULL;
    for (uint32_t i = threadIdx.x; i < nidle; i += blockDim.x) if (p.base[idle[i]] > p.cutoff) idle_later = 1ULL;
    for (;;) {
        if (fixed ? (n >= p.fixed_iters) : (n >= p.max_iter && n >= p.min_iter)) break;
        const uint32_t m = n + 1;
        const bool do_cmp = fixed ? (m >= p.fixed_iters) : (m >= p.min_iter);
        unsigned long long best = 0ULL;
        double asum = 0.0;
        double bnum = -1.0, bden = 1.0;
        // lagged runs: what the grid found DN_LAG iterations ago is loaded now and looked at after the sweep
        unsigned long long pf_arr = 0ULL, pf_mr = 0ULL, pf_sum = 0ULL;
        if (lagged && threadIdx.x == 0 && m > DN_LAG) {
            const uint32_t xs = (m - DN_LAG) & (DN_LAG_SLOTS - 1u);
            pf_arr = ld_acquire_u64(p.ctl + CTL_LAG_ARR + xs);
            pf_mr = ld_cg_u64(p.ctl + CTL_LAG_MAX + xs);
            pf_sum = ld_cg_u64(p.ctl + CTL_LAG_SUM + xs);
        }
        const double* a_rd = lagged ? ring(n) : s_alpha;
        double* a_wr = lagged ? ring(m) : s_alpha;
        // ---- one EM iteration of every component of this warp's tiles (tile -> warp is fixed, so a tile's state is only ever
        //      touched by its own warp: no barrier)
        for (uint32_t k = warp; k < tiles; k += W) {
            uint32_t qi, glanes = G, grank = G ? lane % (G ? G : 1) : 0;
            bool idle_lane = false;
            if (G) qi = k * (32u / (G ? G : 1)) + lane / (G ? G : 1);
            else {
                const uint32_t info = s_lane[(k << 5) + lane];
                idle_lane = info == DN_NONE;
                qi = idle_lane ? 0u : (info & 0xFFFFu);
                glanes = 1u << ((info >> 16) & 0xFu); grank = info >> 20;
            }
            double b[NS], acc[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) { b[j] = idle_lane ? 0.0 : s_beta[(size_t)j * ncomp_pad + qi]; acc[j] = 0.0; }
            const uint32_t L = s_tlen[k];
            const double* cn = s_cnt + s_toff[k] + lane;
            const uint8_t* mk = s_mask + s_toff[k] + lane;
            // DENSE_ILP classes at a time: this warp has little else to run, so the S sums (as trees), the reciprocal chains and the
            // accumulator updates of several classes are kept independent of each other.  Denominators that are neither normal
            // nor zero are left to a rare fix-up outside the straight-line code (a call inside it would fence the scheduler's
            // reordering).  (Measured: no faster than one class at a time -- the loop waits for its heaviest warp, profiles/README.md.)
            for (uint32_t e = 0; e < L; e += DENSE_ILP) {
                double cntv[DENSE_ILP], S[DENSE_ILP], r[DENSE_ILP];
                uint32_t msk[DENSE_ILP];
                bool rare = false;
#pragma unroll
                for (int u = 0; u < DENSE_ILP; ++u) {
                    const bool in = e + u < L;
                    cntv[u] = in ? cn[(e + u) << 5] : 0.0;
                    msk[u] = in ? (uint32_t)mk[(e + u) << 5] : 0u;
                }
#pragma unroll
                for (int u = 0; u < DENSE_ILP; ++u) {
                    double v[NS];
#pragma unroll
                    for (int j = 0; j < NS; ++j) v[j] = ((msk[u] >> j) & 1u) ? b[j] : 0.0;
                    S[u] = dense_tree_sum<NS>(v);
                }
#pragma unroll
                for (int u = 0; u < DENSE_ILP; ++u) {
                    const bool normal = S[u] > 1e-280 && S[u] < 1e280;
                    const double Ss = normal ? S[u] : 1.0;
                    double y;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(Ss));
                    double t = fma(-Ss, y, 1.0);
                    y = fma(y, t, y);
                    t = fma(-Ss, y, 1.0);
                    y = fma(y, t, y);
                    r[u] = normal ? cntv[u] * y : 0.0;
                    rare = rare || (!normal && S[u] != 0.0);
                }
                if (__any_sync(0xffffffffu, rare)) {
#pragma unroll
                    for (int u = 0; u < DENSE_ILP; ++u) r[u] = em_ratio(cntv[u], S[u]);
                }
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    double w[DENSE_ILP];
#pragma unroll
                    for (int u = 0; u < DENSE_ILP; ++u) w[u] = ((msk[u] >> j) & 1u) ? r[u] : 0.0;
                    acc[j] += dense_tree_sum<DENSE_ILP>(w);
                }
            }
            if (G > 1) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
#pragma unroll
                    for (int o = 1; o < G; o <<= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
                }
                if (grank) continue;                                   // the group's first lane owns the component's state
            } else if (G == 0) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
#pragma unroll
                    for (int o = 1; o < (int)DN_MAX_GROUP; o <<= 1) {  // groups are aligned powers of two: lane ^ o stays inside for o < glanes
                        const double other = __shfl_xor_sync(0xffffffffu, acc[j], o);
                        if ((uint32_t)o < glanes) acc[j] += other;
                    }
                }
                if (idle_lane || grank) continue;
            }
            if (do_cmp) {
                // max of |old - new| / new without a division per slot: the largest quotient is kept as (numerator, denominator) and
                // compared by cross-multiplication; it is divided once per thread and iteration (rounding is monotonic, so that IS
                // the max of the rounded quotients, up to the rounding of the products between near-equal candidates).  A loop of
                // its own, so that iterations without the rule (all but the last of a fixed-count run) skip it with one branch.
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const size_t i = (size_t)j * ncomp_pad + qi;
                    double a_old;
                    if constexpr (lagged) a_old = a_rd[i]; else a_old = s_alpha[i];
                    const double a_new = b[j] * acc[j] + s_base[i];
                    const double gate = p.gate_old ? a_old : a_new;
                    const double num = fabs(a_old - a_new);
                    if (gate > p.cutoff && num * bden > bnum * a_new) { bnum = num; bden = a_new; }
                }
            }
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const size_t i = (size_t)j * ncomp_pad + qi;
                const double a_new = b[j] * acc[j] + s_base[i];
                if constexpr (lagged) a_wr[i] = a_new; else s_alpha[i] = a_new;       // (the same array; a second base register costs an IMAD per slot)
                if (VB) asum += a_new; else s_beta[i] = a_new * s_inveff[i];
            }
        }
        if (do_cmp && bnum >= 0.0) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(bnum / bden) + 1ULL;
            best = bits > best ? bits : best;
        }
        if (has_pool) {
            const double* in = p.X + (size_t)bi * p.T;
            double* out = p.X + (size_t)bo * p.T;
            double* spare = p.X + (size_t)bs * p.T;
            // the spare buffer (the input of the previous iteration) becomes an output buffer again: back to its initial value for this
            // CTA's share of the pool's transcripts (every CTA keeps its share, and nobody gathers from that buffer any more)
            if (n > 0) for (uint32_t i = d_lo + threadIdx.x; i < d_hi; i += blockDim.x) { const uint32_t t = q.dlist[i]; spare[t] = __ldg(p.base + t); }
            sweep_block<VB, false>(pb, sl_pool, ptile_lo, ptile_hi, VB ? p.theta : in, out, 0u);
            grid_barrier(p.ctl, nblocks, gen);
            if (VB || do_cmp) {
                for (uint32_t i = d_lo + threadIdx.x; i < d_hi; i += blockDim.x) {
                    const uint32_t t = q.dlist[i];
                    const double a_new = ld_cg_f64(out + t);
                    if (do_cmp) {
                        const double a_old = ld_cg_f64(in + t);
                        const double gate = p.gate_old ? a_old : a_new;
                        if (gate > p.cutoff) {
                            const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a_old - a_new) / a_new) + 1ULL;
                            best = bits > best ? bits : best;
                        }
                    }
                    asum += a_new;
                }
            }
            { const unsigned tmp = bs; bs = bi; bi = bo; bo = tmp; }   // bi now names the pool's newest alphas
        }
        n = m;
        if (do_cmp && m == 1u) {                                       // idle transcripts: alpha_0 -> base at m == 1, base -> base after (idle_later)
            for (uint32_t i = threadIdx.x; i < nidle; i += blockDim.x) {
                const uint32_t t = idle[i];
                const double a_new = p.base[t];
                const double a_old = (m == 1) ? p.X[t] : a_new;
                const double gate = p.gate_old ? a_old : a_new;
                if (gate > p.cutoff) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a_old - a_new) / a_new) + 1ULL;
                    best = bits > best ? bits : best;
                }
            }
        }
        if (do_cmp && m > 1u) best = idle_later > best ? idle_later : best;
        if (lagged) {
            // ---- ONE CTA barrier per iteration.  Before it: the warps fold their part of iteration m into shared memory (atomics, two
            // slots by the parity of m) and thread 0 puts what the grid found in iteration m - DN_LAG (loaded before the sweep) where
            // everybody can read it.  After it: thread 0 sends the arrival for m - 1 (its atomics went out an iteration ago, so the
            // fence in front of it is free), then the CTA's part of m to the global slots; everybody else is already in the next sweep.
            const uint32_t par = m & 1u;
            if (do_cmp) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o > best ? o : best; }
                if (lane == 0 && best) atomicMax(&s_pmax[par], best);
            }
            if (VB) {
                const double v = warp_sum(asum + idle_sum);
                if (lane == 0) atomicAdd(&s_psum[par], v);
            }
            const bool final_it = fixed ? (m >= p.fixed_iters) : (m >= p.max_iter && m >= p.min_iter);
            uint32_t x = m > DN_LAG ? m - DN_LAG : 0u;
            if (threadIdx.x == 0 && x >= 1u) {
                const unsigned long long want = (unsigned long long)nblocks * ((x - 1u) / DN_LAG_SLOTS + 1u);
                if (pf_arr < want) {                                   // not there yet when it was loaded before the sweep: rare
                    while (ld_acquire_u64(p.ctl + CTL_LAG_ARR + (x & (DN_LAG_SLOTS - 1u))) < want) { }
                    pf_mr = ld_cg_u64(p.ctl + CTL_LAG_MAX + (x & (DN_LAG_SLOTS - 1u)));
                    pf_sum = ld_cg_u64(p.ctl + CTL_LAG_SUM + (x & (DN_LAG_SLOTS - 1u)));
                }
                s_res[par * 2u] = pf_mr; s_res[par * 2u + 1u] = pf_sum;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (m > 1u) { __threadfence(); atomicAdd(p.ctl + CTL_LAG_ARR + ((m - 1u) & (DN_LAG_SLOTS - 1u)), 1ULL); }
                const uint32_t slot = m & (DN_LAG_SLOTS - 1u);
                if (do_cmp) { const unsigned long long v = s_pmax[par]; s_pmax[par] = 0ULL; if (v) atomicMax(p.ctl + CTL_LAG_MAX + slot, v); }
                if (VB) { const double v = s_psum[par]; s_psum[par] = 0.0; atomicAdd(reinterpret_cast<double*>(p.ctl + CTL_LAG_SUM + slot), v); }
                if (blockIdx.x == 0) {
                    const uint32_t z = (m + DN_LAG + 1u) & (DN_LAG_SLOTS - 1u);
                    p.ctl[CTL_LAG_MAX + z] = 0ULL; p.ctl[CTL_LAG_SUM + z] = 0ULL;
                }
                if (final_it) { __threadfence(); atomicAdd(p.ctl + CTL_LAG_ARR + slot, 1ULL); }        // no later iteration
            }
            bool stop = false;
            unsigned long long sum_bits = 0ULL;
            bool have_sum = false;
            if (x >= 1u) {
                const unsigned long long mr = s_res[par * 2u];
                sum_bits = s_res[par * 2u + 1u]; have_sum = true;
                const bool checked = fixed ? (x >= p.fixed_iters) : (x >= p.min_iter);
                if (checked && (fixed || x >= p.max_iter || !(decode_mrd(mr) > p.tol))) { stop = true; mr_final = mr; n = x; }   // alpha_x is still in the ring
            }
            if (!stop && final_it) {
                // the last iteration: what is left (m - DN_LAG + 1 .. m) is waited for and looked at in order
                for (x = x + 1u; x <= m; ++x) {
                    __syncthreads();                                   // everybody has read sm_u
                    if (threadIdx.x == 0) {
                        const unsigned long long want = (unsigned long long)nblocks * ((x - 1u) / DN_LAG_SLOTS + 1u);
                        while (ld_acquire_u64(p.ctl + CTL_LAG_ARR + (x & (DN_LAG_SLOTS - 1u))) < want) { }
                        sm_u[0] = ld_cg_u64(p.ctl + CTL_LAG_MAX + (x & (DN_LAG_SLOTS - 1u)));
                        sm_u[1] = ld_cg_u64(p.ctl + CTL_LAG_SUM + (x & (DN_LAG_SLOTS - 1u)));
                    }
                    __syncthreads();
                    const unsigned long long mr = sm_u[0];
                    const bool checked = fixed ? (x >= p.fixed_iters) : (x >= p.min_iter);
                    if (checked && (fixed || x >= p.max_iter || !(decode_mrd(mr) > p.tol))) { stop = true; mr_final = mr; n = x; break; }
                }
            }
            if (stop) break;
            if (VB) {
                const double logNorm = sfb_digamma(have_sum ? __longlong_as_double((long long)sum_bits) : p.sum0);
                const double thetaScale = exp(-logNorm);
                for (uint32_t i = threadIdx.x; i < (uint32_t)ring_stride; i += blockDim.x) {
                    const double a = a_wr[i];
                    s_beta[i] = ((a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0) * s_inveff[i];
                }
                __syncthreads();
            }
            continue;
        }
        if (VB || do_cmp) {
            unsigned long long* slot = p.ctl + CTL_MAXREL + (m & 3u);
            double* csum = reinterpret_cast<double*>(p.ctl + CTL_CSUM + (m & 3u));
            if (do_cmp) block_max_to_slot(best, slot, sm_u);
            if (VB) block_sum_to_slot(asum + idle_sum, csum, sm_d);
            grid_barrier(p.ctl, nblocks, gen);
            if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_MAXREL + ((m + 3u) & 3u)] = 0ULL; p.ctl[CTL_CSUM + ((m + 3u) & 3u)] = 0ULL; }
            if (do_cmp) {
                mr_final = ld_cg_u64(slot);
                if (fixed) break;
                if (!(decode_mrd(mr_final) > p.tol) || m >= p.max_iter) break;
            }
            if (VB) {
                const double logNorm = sfb_digamma(__longlong_as_double((long long)ld_cg_u64(p.ctl + CTL_CSUM + (m & 3u))));
                const double thetaScale = exp(-logNorm);
                for (uint32_t i = threadIdx.x; i < (uint32_t)NS * ncomp_pad; i += blockDim.x) {
                    const double a = s_alpha[i];
                    s_beta[i] = ((a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0) * s_inveff[i];
                }
                __syncthreads();
                if (has_pool) {                                        // the pool's expTheta, complete before anyone's next sweep gathers it
                    const double* cur = p.X + (size_t)bi * p.T;
                    for (uint32_t i = d_lo + threadIdx.x; i < d_hi; i += blockDim.x) {
                        const uint32_t t = q.dlist[i];
                        const double a = ld_cg_f64(cur + t);
                        p.theta[t] = (a > DENORM_MIN) ? sfb_exp_theta(a, logNorm, thetaScale) : 0.0;
                    }
                    grid_barrier(p.ctl, nblocks, gen);
                }
            }
        }
    }
    __syncthreads();
    // the components leave their result in the first third of X: so does the pool
    if (has_pool && bi != 0) {
        const double* cur = p.X + (size_t)bi * p.T;
        for (uint32_t i = d_lo + threadIdx.x; i < d_hi; i += blockDim.x) { const uint32_t t = q.dlist[i]; p.X[t] = ld_cg_f64(cur + t); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = 0ULL; p.ctl[CTL_MRD] = mr_final; }
    const double* a_res = lagged ? ring(n) : s_alpha;
    for (uint32_t i = threadIdx.x; i < (uint32_t)NS * ncomp_pad; i += blockDim.x) { const uint32_t t = tmap[i]; if (t != DN_NONE) p.X[t] = a_res[i]; }
    if (n > 0) for (uint32_t i = threadIdx.x; i < nidle; i += blockDim.x) { const uint32_t t = idle[i]; p.X[t] = p.base[t]; }
}
