// em_part.cuh -- the CTA-partitioned EM loop (included by em.cu inside its anonymous namespace).
//
// Measured on the binned-layout kernel: an EM sweep is bound by the rate at which an SM can issue SCATTERED global accesses
// (the alpha gathers and the red.add scatters), not by bytes.  But the class x transcript structure is local: the isoforms of
// a gene sit next to each other in the transcript order and a class rarely reaches outside its gene family.  So:
//
//   * the transcript range is cut into one contiguous range per CTA, balanced by work;
//   * a class is LOCAL to a CTA if all its members are "clean" transcripts of that CTA's range; a class that crosses ranges
//     goes to the POOL and marks its members dirty, which in turn sends every other class touching them to the pool
//     (closure, a few rounds of a marking kernel);
//   * a CTA keeps its range of alpha (in / out / base [/ expTheta]) and its local classes in SHARED memory: gathers, the
//     denominator reduction and the scatter (shared-memory atomics) never leave the SM;
//   * the pool (possibly empty) is swept by all CTAs through global memory exactly as before.
//
// With an empty pool (the synthetic BASELINE workloads) an iteration needs no grid barrier at all unless the stopping rule
// has to be evaluated (after min_iter) or VBEM needs the global alpha sum; fixed-iteration runs are barrier-free.
// If any CTA's slice does not fit in shared memory the whole run falls back to k_em_persistent.

constexpr uint32_t PART_POOL = 0xFFFFFFFFu;
enum { PT_CLS = 0 /* 7 boundaries */, PT_ENT0 = 7, PT_ENT1 = 8, PT_TXP0 = 9, PT_TXP1 = 10, PT_WORDS = 12 };

__device__ __forceinline__ int part_bin(uint32_t n) { return n <= 2 ? 0 : n <= 4 ? 1 : n <= 8 ? 2 : n <= 16 ? 3 : n <= 32 ? 4 : 5; }

__device__ __forceinline__ uint32_t part_range_of(const uint32_t* __restrict__ bounds, uint32_t n_cta, uint32_t t) {
    uint32_t lo = 0, hi = n_cta;                     // bounds[lo] <= t < bounds[hi]
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(bounds + mid) <= t) lo = mid; else hi = mid; }
    return lo;
}

// crossing profile: diff[t] accumulates +1 at (min member)+1 and -1 at (max member)+1 of every class, so that the prefix sum
// cross[t] is the number of classes with min < t <= max, i.e. the classes a range boundary placed at t would cut
// The same pass charges the class's sweep cost (the lanes its sub-warp group occupies) to its smallest member: ranges are
// balanced by that, since the sweep is instruction-bound and a warp tile costs the same whatever its classes hold.
__global__ void k_part_span(const uint32_t* __restrict__ start, const uint32_t* __restrict__ len, const uint32_t* __restrict__ lab,
                            uint64_t Em, int* __restrict__ diff, uint32_t* __restrict__ load) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= Em) return;
    const uint32_t b = start[c], n = len[c];
    uint32_t mn = lab[b], mx = mn;
    for (uint32_t j = 1; j < n; ++j) { const uint32_t t = lab[b + j]; mn = t < mn ? t : mn; mx = t > mx ? t : mx; }
    if (mn < mx) { atomicAdd(diff + mn + 1, 1); atomicSub(diff + mx + 1, 1); }
    const uint32_t g = n <= 2 ? 2u : n <= 4 ? 4u : n <= 8 ? 8u : n <= 16 ? 16u : n <= 32 ? 32u : 2u * n;
    atomicAdd(load + mn, g);
}

// Range boundaries on the device (one CTA): prefix sums of the per-transcript sweep cost (1 + load) and of the crossing profile,
// then boundary i = the first position where the cost prefix reaches i/n_cta of the total, moved to the nearest position no
// class crosses (within `win`), then made monotone.  pre: T scratch words; diff is turned into its prefix sum in place.
__global__ void __launch_bounds__(1024) k_part_bounds(const uint32_t* __restrict__ load, int* __restrict__ diff, uint32_t T, uint32_t n_cta,
                                                      uint32_t win, unsigned long long* __restrict__ pre, uint32_t* __restrict__ bounds) {
    __shared__ unsigned long long s_wa[32];
    __shared__ long long s_wd[32];
    __shared__ unsigned long long s_total;
    extern __shared__ uint32_t s_b[];                    // n_cta + 1
    const uint32_t nth = blockDim.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nw = nth >> 5;
    // block-wide inclusive scans over tiles of PB_ITEMS * blockDim elements (the profile has T + 1 entries).  A tile is PB_ITEMS
    // sub-tiles of blockDim consecutive elements, thread tid holding element tid of each (coalesced loads and stores: with a thread's
    // elements next to each other every load touched 32 sectors and this one-CTA kernel spent 0.3 ms waiting for them); each warp scans
    // its 32 elements of every sub-tile, the PB_ITEMS x 32 warp totals are scanned in shared memory, three barriers per tile.
    constexpr int PB_ITEMS = 8;
    __shared__ unsigned long long s_ta[PB_ITEMS * 32];
    __shared__ long long s_td[PB_ITEMS * 32];
    unsigned long long carry_a = 0; long long carry_d = 0;
    for (uint32_t base = 0; base <= T; base += nth * PB_ITEMS) {
        unsigned long long av[PB_ITEMS]; long long dv[PB_ITEMS];
#pragma unroll
        for (int q = 0; q < PB_ITEMS; ++q) {
            const uint32_t t = base + q * nth + tid;
            av[q] = t < T ? 1ull + load[t] : 0ull;
            dv[q] = t <= T ? (long long)diff[t] : 0ll;
        }
#pragma unroll
        for (int q = 0; q < PB_ITEMS; ++q) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ua = __shfl_up_sync(0xffffffffu, av[q], o);
                const long long ud = __shfl_up_sync(0xffffffffu, dv[q], o);
                if ((int)lane >= o) { av[q] += ua; dv[q] += ud; }
            }
            if (lane == 31) { s_ta[q * 32 + warp] = av[q]; s_td[q * 32 + warp] = dv[q]; }
        }
        if (nw < 32 && lane == 31) for (int q = 0; q < PB_ITEMS; ++q) for (uint32_t w = nw; w < 32; ++w) { s_ta[q * 32 + w] = 0; s_td[q * 32 + w] = 0; }
        __syncthreads();
        // warp q scans the 32 warp totals of sub-tile q (inclusive); s_wa / s_wd take the sub-tile totals
        if (warp < (uint32_t)PB_ITEMS) {
            unsigned long long wa = s_ta[warp * 32 + lane];
            long long wd = s_td[warp * 32 + lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ua = __shfl_up_sync(0xffffffffu, wa, o);
                const long long ud = __shfl_up_sync(0xffffffffu, wd, o);
                if ((int)lane >= o) { wa += ua; wd += ud; }
            }
            s_ta[warp * 32 + lane] = wa; s_td[warp * 32 + lane] = wd;
            if (lane == 31) { s_wa[warp] = wa; s_wd[warp] = wd; }
        }
        __syncthreads();
        unsigned long long run_a = carry_a; long long run_d = carry_d;              // everything in front of sub-tile q
#pragma unroll
        for (int q = 0; q < PB_ITEMS; ++q) {
            const uint32_t t = base + q * nth + tid;
            const unsigned long long ex_a = run_a + (warp ? s_ta[q * 32 + warp - 1] : 0ull);
            const long long ex_d = run_d + (warp ? s_td[q * 32 + warp - 1] : 0ll);
            if (t < T) pre[t] = ex_a + av[q];
            if (t <= T) diff[t] = (int)(ex_d + dv[q]);
            run_a += s_wa[q]; run_d += s_wd[q];
        }
        carry_a = run_a; carry_d = run_d;
        __syncthreads();
    }
    if (tid == 0) s_total = carry_a;
    __syncthreads();
    const unsigned long long total = s_total;
    // a WARP per boundary: a 32-ary search for the cost target (4 rounds of one load per lane instead of 18 dependent loads) and the
    // lanes of the warp looking at 32 window offsets at a time (a thread by itself walked the window with two dependent loads per
    // offset: one boundary far from any uncrossed position cost win x 2 round trips, ~0.3 ms, whatever the other threads did)
    for (uint32_t i = warp; i <= n_cta; i += nw) {
        uint32_t b;
        if (i == 0) b = 0;
        else if (i == n_cta) b = T;
        else {
            const unsigned long long target = total * i / n_cta;
            uint32_t l = 0, h = T;                       // first t with pre[t] >= target (T if none): the answer lies in [l, h]
            while (l < h) {
                const uint32_t step = (h - l + 31u) / 32u;
                const uint32_t c_lo = l + lane * step;                          // this lane's chunk [c_lo, c_hi)
                const uint32_t c_hi = c_lo + step < h ? c_lo + step : h;
                const bool ge = c_lo < c_hi && pre[c_hi - 1] >= target;         // the chunk's last element has reached the target
                const unsigned m = __ballot_sync(0xffffffffu, ge);
                if (m == 0u) { l = h; break; }                                  // nothing below h has
                const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
                const uint32_t k_lo = l + k * step, k_hi = (k_lo + step < h ? k_lo + step : h);
                l = k_lo; h = k_hi - 1u;                                        // pre[k_hi - 1] >= target: the answer is at most k_hi - 1
            }
            b = l < T ? l + 1 : T;
            uint32_t best = b;
            for (uint32_t d0 = 0; d0 <= win; d0 += 32u) {
                const uint32_t dd = d0 + lane;
                const bool okm = dd <= win && b >= dd && diff[b - dd] == 0;
                const bool okp = dd <= win && b + dd < T && diff[b + dd] == 0;
                const unsigned mm = __ballot_sync(0xffffffffu, okm), mp = __ballot_sync(0xffffffffu, okp);
                if (mm | mp) {                                                  // the smallest offset, the position below before the one above
                    const uint32_t k = (uint32_t)__ffs((int)(mm | mp)) - 1u;
                    best = ((mm >> k) & 1u) ? b - (d0 + k) : b + (d0 + k);
                    break;
                }
            }
            b = best;
        }
        if (lane == 0) s_b[i] = b;
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t i = 1; i <= n_cta; ++i) if (s_b[i] < s_b[i - 1]) s_b[i] = s_b[i - 1];
    }
    __syncthreads();
    for (uint32_t i = tid; i <= n_cta; i += nth) bounds[i] = s_b[i];
}

// one closure round: classes that span ranges or touch a dirty transcript go to the pool and dirty all their members
__global__ void k_part_owner(const uint32_t* __restrict__ start, const uint32_t* __restrict__ len, const uint32_t* __restrict__ lab,
                             uint64_t Em, const uint32_t* __restrict__ bounds, uint32_t n_cta, uint8_t* __restrict__ dirty,
                             uint32_t* __restrict__ owner, unsigned int* __restrict__ changed) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= Em) return;
    const uint32_t b = start[c], n = len[c];
    const uint32_t o = part_range_of(bounds, n_cta, lab[b]);
    bool cut = false;
    for (uint32_t j = 0; j < n && !cut; ++j) {
        const uint32_t t = lab[b + j];
        cut = dirty[t] != 0 || part_range_of(bounds, n_cta, t) != o;
    }
    if (cut) {
        owner[c] = PART_POOL;
        for (uint32_t j = 0; j < n; ++j) { const uint32_t t = lab[b + j]; if (!dirty[t]) { dirty[t] = 1; *changed = 1u; } }
    } else {
        owner[c] = o;
    }
}

// group = owner * 6 + bin (pool = n_cta); packed counters: classes in the high half, entries in the low half
__global__ void k_part_count(const uint32_t* __restrict__ owner, const uint32_t* __restrict__ len, uint64_t Em, uint32_t n_cta,
                             unsigned long long* __restrict__ grp) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= Em) return;
    const uint32_t o = owner[c] == PART_POOL ? n_cta : owner[c];
    atomicAdd(grp + (size_t)o * SFB_NBINS + part_bin(len[c]), (1ULL << 32) | len[c]);
}

__global__ void k_part_fill(const uint32_t* __restrict__ owner, const uint32_t* __restrict__ start, const uint32_t* __restrict__ len,
                            const uint32_t* __restrict__ lab, const double* __restrict__ cnt, uint64_t Em, uint32_t n_cta,
                            const unsigned long long* __restrict__ cls_off, const unsigned long long* __restrict__ nnz_off,
                            unsigned long long* __restrict__ cursor, uint32_t* __restrict__ start2, uint32_t* __restrict__ len2,
                            uint32_t* __restrict__ lab2, double* __restrict__ cnt2, uint32_t* __restrict__ src) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= Em) return;
    const uint32_t o = owner[c] == PART_POOL ? n_cta : owner[c];
    const uint32_t n = len[c];
    const size_t g = (size_t)o * SFB_NBINS + part_bin(n);
    const unsigned long long tk = atomicAdd(cursor + g, (1ULL << 32) | n);
    const uint64_t pos = cls_off[g] + (tk >> 32), e = nnz_off[g] + (tk & 0xFFFFFFFFULL);
    start2[pos] = (uint32_t)e; len2[pos] = n; cnt2[pos] = cnt[c]; src[pos] = (uint32_t)c;
    const uint32_t b = start[c];
    for (uint32_t j = 0; j < n; ++j) lab2[e + j] = lab[b + j];
}

__global__ void k_part_gather_counts(const double* __restrict__ cnt, const uint32_t* __restrict__ src, uint64_t Em, double* __restrict__ cnt2) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < Em) cnt2[c] = cnt[src[c]];
}

struct PartParams {
    const unsigned long long* tbl;     // n_cta rows of PT_WORDS
    const uint8_t* dirty;              // T
    uint32_t smem_bytes;
    int has_pool;
};

// ---- the partitioned persistent loop ------------------------------------------------------------------------------------------
template <bool VB>
__global__ void __launch_bounds__(EM_THREADS, 1) k_em_part(const EmParams p, const PartParams q) {
    __shared__ unsigned long long sm_u[32];
    __shared__ double sm_d[32];
    __shared__ uint64_t tma_bar;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const unsigned nblocks = gridDim.x;
    unsigned long long gen = 0;
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gstride = (uint64_t)nblocks * blockDim.x;

    // ---- this CTA's slice
    const unsigned long long* row = q.tbl + (size_t)blockIdx.x * PT_WORDS;
    Bins lb;
    for (int i = 0; i <= SFB_NBINS; ++i) lb.cls_start[i] = row[PT_CLS + i];
    bins_set_tiles(lb);
    const uint64_t e0 = row[PT_ENT0], e1 = row[PT_ENT1];
    const uint32_t t0 = (uint32_t)row[PT_TXP0], t1 = (uint32_t)row[PT_TXP1];
    const uint32_t nt = t1 - t0;
    const uint64_t c_lo = lb.cls_start[0] & ~3ULL, e_lo = e0 & ~3ULL;
    const uint64_t nc = ((lb.cls_start[SFB_NBINS] - c_lo) + 3) & ~3ULL, ne = ((e1 - e_lo) + 3) & ~3ULL;
    const uint32_t ntp = (nt + 3u) & ~3u;
    double* s_cnt = reinterpret_cast<double*>(dyn_smem);
    double* s_w = s_cnt + nc;
    double* s_a = s_w + ne;
    double* s_b = s_a + ntp;
    double* s_theta = s_b + ntp;                                      // VBEM only (space reserved only then)
    uint32_t* s_start = reinterpret_cast<uint32_t*>(s_theta + (VB ? ntp : 0));
    uint32_t* s_len = s_start + nc;
    uint32_t* s_lab = s_len + nc;
    uint8_t* s_dirty = reinterpret_cast<uint8_t*>(s_lab + ne);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&tma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool have_cls = lb.cls_start[SFB_NBINS] > lb.cls_start[0];
    if (have_cls && threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(nc * 16 + ne * 12);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&tma_bar)), "r"(bytes) : "memory");
        tma_load_1d(s_cnt, p.cnt + c_lo, (uint32_t)(nc * 8), &tma_bar);
        tma_load_1d(s_w, p.w + e_lo, (uint32_t)(ne * 8), &tma_bar);
        tma_load_1d(s_start, p.start + c_lo, (uint32_t)(nc * 4), &tma_bar);
        tma_load_1d(s_len, p.len + c_lo, (uint32_t)(nc * 4), &tma_bar);
        tma_load_1d(s_lab, p.lab + e_lo, (uint32_t)(ne * 4), &tma_bar);
    }
    for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) {
        const uint8_t d = q.dirty[t0 + i];
        s_dirty[i] = d;
        s_a[i] = d ? 0.0 : p.X[t0 + i];                               // alpha_0 (buffer 0)
        s_b[i] = d ? 0.0 : p.base[t0 + i];
    }
    if (have_cls) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(done) : "r"(smem_u32(&tma_bar)) : "memory");
        }
    }
    __syncthreads();
    Slice sl_loc; sl_loc.start = s_start; sl_loc.len = s_len; sl_loc.cnt = s_cnt; sl_loc.lab = s_lab; sl_loc.w = s_w; sl_loc.c0 = c_lo; sl_loc.e0 = e_lo;
    const uint64_t loc_tiles = lb.tile_start[SFB_NBINS];
    // ---- the pool (global memory), split evenly over the CTAs
    Slice sl_pool; sl_pool.start = p.start; sl_pool.len = p.len; sl_pool.cnt = p.cnt; sl_pool.lab = p.lab; sl_pool.w = p.w; sl_pool.c0 = 0; sl_pool.e0 = 0;
    const Bins pb = em_bins(p);
    const uint64_t pool_tiles = p.tile_start[SFB_NBINS];
    const uint64_t ptile_lo = pool_tiles * blockIdx.x / nblocks, ptile_hi = pool_tiles * (blockIdx.x + 1ULL) / nblocks;
    const bool has_pool = q.has_pool != 0;

    double* s_in = s_a; double* s_prev = s_b;                          // s_prev doubles as the output buffer of the sweep
    unsigned bi = 0, bo = 1, bs = 2;
    const bool fixed = p.fixed_iters > 0;
    uint32_t n = 0;
    for (;;) {
        double* in = p.X + (size_t)bi * p.T;
        double* out = p.X + (size_t)bo * p.T;
        double* spare = p.X + (size_t)bs * p.T;
        unsigned long long* slot = p.ctl + CTL_MAXREL + (n & 3u);
        const bool last = fixed ? (n >= p.fixed_iters) : (n >= p.max_iter && n >= p.min_iter);
        const bool check_now = !fixed && n > 0 && n >= p.min_iter;
        const bool bar_this = has_pool || VB || check_now;
        const bool do_cmp = n > 0 && (last || check_now);
        if (bar_this && blockIdx.x == 0 && threadIdx.x == 0) {
            p.ctl[CTL_MAXREL + ((n + 2u) & 3u)] = 0ULL;
            p.ctl[CTL_CSUM + ((n + 2u) & 3u)] = 0ULL;
        }
        double logNorm = 0.0;
        if (VB && !last) {
            const double asum = (n == 0) ? p.sum0
                : p.base_sum + __longlong_as_double((long long)ld_cg_u64(p.ctl + CTL_CSUM + (n & 3u)));
            logNorm = sfb_digamma(asum);
        }
        const double thetaScale = (VB && !last) ? exp(-logNorm) : 0.0;
        unsigned long long best = 0ULL;
        if (has_pool) best = transcript_pass<VB>(p, spare, in, spare, do_cmp, !last, logNorm, gtid, gstride);
        for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) {       // local transcript pass
            const double cur = s_in[i];
            if (do_cmp) {
                const double pv = s_prev[i];
                const double gate = p.gate_old ? pv : cur;
                if (gate > p.cutoff) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(pv - cur) / cur) + 1ULL;
                    best = bits > best ? bits : best;
                }
            }
            if (!last) {
                s_prev[i] = s_dirty[i] ? 0.0 : __ldg(p.base + t0 + i);     // base stays in global memory (coalesced, L2-resident)
                if (VB) s_theta[i] = (cur > DENORM_MIN) ? sfb_exp_theta(cur, logNorm, thetaScale) : 0.0;
            }
        }
        if (do_cmp) block_max_to_slot(best, slot, sm_u);
        if (last) {
            grid_barrier(p.ctl, nblocks, gen);
            if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = bi; p.ctl[CTL_MRD] = ld_cg_u64(slot); }
            for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) if (!s_dirty[i]) in[t0 + i] = s_in[i];
            return;
        }
        __syncthreads();
        if (VB && has_pool) grid_barrier(p.ctl, nblocks, gen);          // the pool's expTheta is complete
        double contrib = sweep_block<VB, true>(lb, sl_loc, 0, loc_tiles, VB ? s_theta : s_in, s_prev, t0);
        if (has_pool) contrib += sweep_block<VB, false>(pb, sl_pool, ptile_lo, ptile_hi, VB ? p.theta : in, out, 0u);
        if (VB) block_sum_to_slot(contrib, reinterpret_cast<double*>(p.ctl + CTL_CSUM + ((n + 1u) & 3u)), sm_d);
        __syncthreads();
        if (bar_this) grid_barrier(p.ctl, nblocks, gen);
        if (check_now) {
            const unsigned long long mr = ld_cg_u64(slot);
            if (!(decode_mrd(mr) > p.tol)) {
                if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctl[CTL_ITERS] = n; p.ctl[CTL_RESULT_BUF] = bi; p.ctl[CTL_MRD] = mr; }
                for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) if (!s_dirty[i]) in[t0 + i] = s_in[i];
                return;
            }
        }
        double* tmpd = s_in; s_in = s_prev; s_prev = tmpd;
        const unsigned tmp = bs; bs = bi; bi = bo; bo = tmp;
        ++n;
    }
}
