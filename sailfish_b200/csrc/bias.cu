// bias.cu -- effective lengths corrected for sequence-specific or fragment-GC bias on the device.
//
// Replaces sailfish::utils::updateEffectiveLengths (reference src/SailfishUtils.cpp:611-926), which optimize() calls at
// iterations 50 / 500 / 1000 when --biasCorrect or --gcBiasCorrect is given (src/CollapsedEMOptimizer.cpp:816-840).  Two passes
// over every position of every expressed transcript:
//   pass 1  expected distributions: a 4096-bin histogram over the 6-mer context of every possible fragment start (both strands,
//           weighted by alpha/effLen and the fragment-length cdf), or a 101-bin histogram over the GC percentage of every
//           (start, fragment length) pair;
//   pass 2  per transcript, the sum over positions of observed/expected ratios = its corrected effective length.
// O(sum of transcript lengths x [1 | FLD window]) -- the most expensive optional stage of the reference, embarrassingly
// parallel.  Here: persistent CTAs deal transcripts round-robin, a CTA keeps its histogram in shared memory and adds it to the
// global one once; the 6-mer context of a position is 12 bits of the index's 2-bit text (reverse complement = bitwise not, as
// the codes are A 0, C 1, G 2, T 3); GC counts of any interval come from a per-word prefix of G/C counts plus one popcount.
// The text is the device index's: a base that was not A/C/G/T in the FASTA is the deterministic substitute the index stores
// (DESIGN.md section 3), where the reference would run off its tables (indexForKmer returns UINT32_MAX for such a window).
//
// STATUS: parity with the CPU oracle (oracle/orc_bias.cpp, which is pinned to the reference's own function body) is green on a
// B200 (tests/test_gpu_bias.py; profiles/r01f_experimental_gpu.txt); not timed or profiled yet, and the quantification drivers
// do not call it yet.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int BIAS_THREADS = 256;

#define SFB_BD __device__ __forceinline__
#define SFB_LDG(p) __ldg(p)
#define SFB_POPC64(x) __popcll(x)
#define SFB_D2I_RN(x) __double2int_rn(x)
#include "bias_core.inl"
#undef SFB_BD
#undef SFB_LDG
#undef SFB_POPC64
#undef SFB_D2I_RN

__global__ void k_bias_gc_words(const uint64_t* __restrict__ words, uint64_t n_words, uint32_t* __restrict__ cnt) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint64_t w = words[i];
    cnt[i] = (uint32_t)__popcll((w ^ (w >> 1)) & 0x5555555555555555ULL);
}

// ---- pass 1 (:696-786): expected distributions.  MODE 1: hist has 4096 bins, MODE 2: 101
template <int MODE>
__global__ void __launch_bounds__(BIAS_THREADS) k_bias_expected(const BiasView v, double* __restrict__ hist) {
    constexpr uint32_t NB = MODE == 1 ? BNK : 101u;
    __shared__ double s_h[NB];
    for (uint32_t i = threadIdx.x; i < NB; i += blockDim.x) s_h[i] = 0.0;
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < v.T; t += gridDim.x) {
        int32_t refLen, unproc;
        if (!b_eligible(v, t, refLen, unproc)) continue;                       // uniform over the CTA
        const double contribution = __ldg(v.alphas + t) / __ldg(v.eff_in + t);
        const uint64_t t0 = __ldg(v.txp_start + t);
        for (int32_t i = (int32_t)threadIdx.x; i <= refLen - BK - 1; i += blockDim.x) {
            auto add = [&](uint32_t bin, double x) { atomicAdd(&s_h[bin], x); };
            if (MODE == 1) b_expected_seq(v, t0, refLen, i, contribution, add);
            else b_expected_gc(v, t0, refLen, i, contribution, add);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < NB; i += blockDim.x) if (s_h[i] != 0.0) atomicAdd(hist + i, s_h[i]);
}

// ---- pass 2 (:811-924): ratio[bin] = observed / (expected + prior); one transcript per CTA at a time, block sum
template <int MODE>
__global__ void __launch_bounds__(BIAS_THREADS) k_bias_efflen(const BiasView v, const double* __restrict__ ratio, double norm,
                                                              double* __restrict__ eff_out) {
    __shared__ double s_w[BIAS_THREADS / 32];
    for (uint32_t t = blockIdx.x; t < v.T; t += gridDim.x) {
        int32_t refLen, unproc;
        const bool go = b_eligible(v, t, refLen, unproc);
        double sum = 0.0;
        if (go) {
            const uint64_t t0 = __ldg(v.txp_start + t);
            for (int32_t i = (int32_t)threadIdx.x; i <= refLen - BK - 1; i += blockDim.x) {
                sum += MODE == 1 ? b_eff_seq(v, ratio, t0, refLen, i) : b_eff_gc(v, ratio, t0, refLen, i);
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
        __syncthreads();
        if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double eff = 0.0;
            for (int w = 0; w < BIAS_THREADS / 32; ++w) eff += s_w[w];
            eff *= norm;                                                          // txomeNormFactor / readNormFactor (:901-912)
            const double in = __ldg(v.eff_in + t);
            eff_out[t] = (go && unproc > 0 && eff > (double)unproc) ? eff : in;   // :916-922
        }
    }
}

// ---- the fragment GC passes in sliding form (bias_core.inl, b_gc_slide): a thread owns one sampled fragment length and walks the
// transcript, so its G/C count changes by one base per step.  SLIDE_THREADS fragment lengths per tile.
constexpr int SLIDE_THREADS = 128;
constexpr int SLIDE_STRIDE = SLIDE_THREADS + 1;             // counts laid out [bin][fragment length]: conflict-free both ways

// pass 1: per (fragment length, bin) integer counts in shared memory (a private column per thread: plain read-modify-write),
// folded into the CTA's 101 doubles once per transcript and tile by one thread per bin; w[k] = cdf(fl_k) - cdf(fl_{k-1})
__global__ void __launch_bounds__(SLIDE_THREADS) k_bias_expected_gc_slide(const BiasView v, const double* __restrict__ w, int n_fl,
                                                                          double* __restrict__ hist) {
    extern __shared__ uint32_t s_cnt[];                      // 101 x SLIDE_STRIDE
    __shared__ double s_h[101];
    for (uint32_t i = threadIdx.x; i < 101u * SLIDE_STRIDE; i += blockDim.x) s_cnt[i] = 0;
    if (threadIdx.x < 101) s_h[threadIdx.x] = 0.0;
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < v.T; t += gridDim.x) {
        int32_t refLen, unproc;
        if (!b_eligible(v, t, refLen, unproc)) continue;                           // uniform over the CTA
        const double contribution = __ldg(v.alphas + t) / __ldg(v.eff_in + t);
        const uint64_t t0 = __ldg(v.txp_start + t);
        for (int tile = 0; tile < n_fl; tile += SLIDE_THREADS) {
            const int k = tile + (int)threadIdx.x;
            if (k < n_fl) {
                int32_t last = -1; uint32_t run = 0;                               // consecutive starts mostly fall into the same bin
                b_gc_slide(v.words, t0, refLen, v.fldLow + k * v.gcSamp, [&](int32_t bin) {
                    if (bin == last) { ++run; return; }
                    if (run) s_cnt[last * SLIDE_STRIDE + threadIdx.x] += run;
                    last = bin; run = 1;
                });
                if (run) s_cnt[last * SLIDE_STRIDE + threadIdx.x] += run;
            }
            __syncthreads();
            if (threadIdx.x < 101) {
                const int nf = n_fl - tile < SLIDE_THREADS ? n_fl - tile : SLIDE_THREADS;
                uint32_t* row = s_cnt + threadIdx.x * SLIDE_STRIDE;
                double acc = 0.0;
                for (int f = 0; f < nf; ++f) { const uint32_t c = row[f]; if (c) { acc += __ldg(w + tile + f) * (double)c; row[f] = 0; } }
                s_h[threadIdx.x] += contribution * acc;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < 101 && s_h[threadIdx.x] != 0.0) atomicAdd(hist + threadIdx.x, s_h[threadIdx.x]);
}

// pass 2: wf[k] = w[k] probFwd + w[k] probRC; a transcript's length = norm * sum_k wf[k] * sum_starts ratio[bin(start, fl_k)]
__global__ void __launch_bounds__(SLIDE_THREADS) k_bias_efflen_gc_slide(const BiasView v, const double* __restrict__ ratio,
                                                                        const double* __restrict__ wf, int n_fl, double norm,
                                                                        double* __restrict__ eff_out) {
    __shared__ double s_ratio[101];
    __shared__ double s_w[SLIDE_THREADS / 32];
    if (threadIdx.x < 101) s_ratio[threadIdx.x] = ratio[threadIdx.x];
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < v.T; t += gridDim.x) {
        int32_t refLen, unproc;
        const bool go = b_eligible(v, t, refLen, unproc);
        double sum = 0.0;
        if (go) {
            const uint64_t t0 = __ldg(v.txp_start + t);
            for (int k = (int)threadIdx.x; k < n_fl; k += SLIDE_THREADS) {
                double acc = 0.0;
                b_gc_slide(v.words, t0, refLen, v.fldLow + k * v.gcSamp, [&](int32_t bin) { acc += s_ratio[bin]; });
                sum += __ldg(wf + k) * acc;
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
        __syncthreads();
        if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double eff = 0.0;
            for (int q = 0; q < SLIDE_THREADS / 32; ++q) eff += s_w[q];
            eff *= norm;
            const double in = __ldg(v.eff_in + t);
            eff_out[t] = (go && unproc > 0 && eff > (double)unproc) ? eff : in;
        }
    }
}

inline unsigned b_grid(uint64_t n, unsigned th) { return (unsigned)((n + th - 1) / th); }

}  // namespace

extern "C" int sfb200_bias_eff_lens(sfb200_ctx* c, const sfb200_bias_model* m, const double* eff_model, const double* eff_in,
                                    const double* alphas, uint32_t n_txp, double* eff_out) {
    if (!c || !m || !eff_model || !eff_in || !alphas || !eff_out) return SFB200_EINVAL;
    if (!c->index.ready) SFB_FAIL(c, SFB200_EINVAL, "bias_eff_lens: build the index first (the correction reads the transcript sequences)");
    if (n_txp != c->index.n_txp) SFB_FAIL(c, SFB200_EINVAL, "bias_eff_lens: n_txp differs from the index's");
    if (m->mode != 1 && m->mode != 2) SFB_FAIL(c, SFB200_EINVAL, "bias_eff_lens: mode must be 1 (sequence bias) or 2 (fragment GC bias)");
    if (!m->fld_cdf || (m->mode == 1 && !m->read_bias) || (m->mode == 2 && !m->observed_gc)) SFB_FAIL(c, SFB200_EINVAL, "bias_eff_lens: null model array");
    const int64_t numMappings = m->num_fwd + m->num_rc;
    if (numMappings == 0) { std::memcpy(eff_out, eff_in, n_txp * sizeof(double)); return SFB200_OK; }      // :627-632: correction skipped
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    const DevIndex& ix = c->index;
    const uint32_t T = n_txp;
    const bool seq = m->mode == 1;
    const uint32_t NB = seq ? BNK : 101u;
    const uint64_t n_words = ix.text_len / 32 + 2;
    DevBuf<uint32_t> d_gcw; DevBuf<unsigned char> d_tmp; DevBuf<float> d_cdf; DevBuf<double> d_vec, d_hist, d_w;
    auto cleanup = [&]() { d_gcw.release(); d_tmp.release(); d_cdf.release(); d_vec.release(); d_hist.release(); d_w.release(); };
#define BIAS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return SFB200_ECUDA; } } while (0)
    BIAS_CUDA(d_cdf.reserve(m->n_cdf ? m->n_cdf : 1));
    BIAS_CUDA(d_vec.reserve(3ull * T + T));                       // eff_model | eff_in | alphas | eff_out
    BIAS_CUDA(d_hist.reserve(2ull * NB));                         // expected | ratio
    if (m->n_cdf) BIAS_CUDA(cudaMemcpyAsync(d_cdf.p, m->fld_cdf, m->n_cdf * sizeof(float), cudaMemcpyHostToDevice, s));
    BIAS_CUDA(cudaMemcpyAsync(d_vec.p, eff_model, T * 8ull, cudaMemcpyHostToDevice, s));
    BIAS_CUDA(cudaMemcpyAsync(d_vec.p + T, eff_in, T * 8ull, cudaMemcpyHostToDevice, s));
    BIAS_CUDA(cudaMemcpyAsync(d_vec.p + 2ull * T, alphas, T * 8ull, cudaMemcpyHostToDevice, s));
    BiasView v;
    v.words = ix.words.p; v.txp_start = ix.txp_start.p; v.txp_len = ix.txp_len.p; v.gcw = nullptr;
    v.cdf = d_cdf.p; v.n_cdf = m->n_cdf; v.eff_model = d_vec.p; v.eff_in = d_vec.p + T; v.alphas = d_vec.p + 2ull * T; v.T = T;
    v.probFwd = static_cast<double>(m->num_fwd) / numMappings; v.probRC = static_cast<double>(m->num_rc) / numMappings;
    v.fldLow = 0; v.fldHigh = 1; v.gcSamp = (int32_t)std::max<uint32_t>(1, m->gc_samp);
    auto cdf = [&](uint32_t x) -> float { return x < m->n_cdf ? m->fld_cdf[x] : 1.0f; };
    if (!seq) {                                                   // :668-687
        bool first = false, second = false;
        for (uint32_t i = 0; i <= m->fld_max; ++i) {
            const float density = cdf(i);
            if (!first && density >= 0.005) { first = true; v.fldLow = (int32_t)i; }
            if (!second && density >= 0.995) { second = true; v.fldHigh = (int32_t)i; }
        }
    }
    // the GC passes in sliding form: opt-in (SFB200_BIAS_GC_SLIDE=1) until they have had their first GPU run
    const char* slide_env = getenv("SFB200_BIAS_GC_SLIDE");
    const bool slide = !seq && slide_env && atoi(slide_env) != 0 && v.fldLow >= 1 && v.fldHigh >= v.fldLow;
    int n_fl = 0;
    if (slide) {
        std::vector<double> w2;
        double prev = static_cast<double>(cdf(0));
        for (int32_t fl = v.fldLow; fl <= v.fldHigh; fl += v.gcSamp) {
            const double cur = static_cast<double>(cdf((uint32_t)fl));
            w2.push_back(cur - prev);
            prev = cur;
        }
        n_fl = (int)w2.size();
        for (int k = 0; k < n_fl; ++k) w2.push_back(w2[k] * v.probFwd + w2[k] * v.probRC);
        BIAS_CUDA(d_w.reserve(w2.size()));
        BIAS_CUDA(cudaMemcpyAsync(d_w.p, w2.data(), w2.size() * 8ull, cudaMemcpyHostToDevice, s));
        BIAS_CUDA(cudaStreamSynchronize(s));                  // w2 is a local
    } else if (!seq) {
        // per-word G/C counts -> exclusive prefix
        BIAS_CUDA(d_gcw.reserve(n_words + 1));
        k_bias_gc_words<<<b_grid(n_words, 256), 256, 0, s>>>(ix.words.p, n_words, d_gcw.p);
        c->launches++;
        size_t tmp = 0;
        BIAS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_gcw.p, d_gcw.p, (int)n_words, s));
        BIAS_CUDA(d_tmp.reserve(tmp));
        BIAS_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp, d_gcw.p, d_gcw.p, (int)n_words, s));
        c->launches++;
        v.gcw = d_gcw.p;
    }
    // ---- pass 1: expected distribution, starting from 1.0 in every bin (:649-651, :669-671)
    std::vector<double> h_hist(NB, 1.0);
    BIAS_CUDA(cudaMemcpyAsync(d_hist.p, h_hist.data(), NB * 8ull, cudaMemcpyHostToDevice, s));
    const unsigned grid = (unsigned)std::min<uint64_t>(T, (uint64_t)c->num_sms * 8);
    const size_t slide_smem = 101ull * SLIDE_STRIDE * sizeof(uint32_t);
    if (slide) {
        BIAS_CUDA(cudaFuncSetAttribute(k_bias_expected_gc_slide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slide_smem));
        k_bias_expected_gc_slide<<<grid, SLIDE_THREADS, slide_smem, s>>>(v, d_w.p, n_fl, d_hist.p);
    } else if (seq) k_bias_expected<1><<<grid, BIAS_THREADS, 0, s>>>(v, d_hist.p);
    else k_bias_expected<2><<<grid, BIAS_THREADS, 0, s>>>(v, d_hist.p);
    c->launches++;
    BIAS_CUDA(cudaGetLastError());
    BIAS_CUDA(cudaMemcpyAsync(h_hist.data(), d_hist.p, NB * 8ull, cudaMemcpyDeviceToHost, s));
    BIAS_CUDA(cudaStreamSynchronize(s));
    // ---- priors, normalisers, observed / expected (:789-804)
    double txomeNorm = 0.0, readNorm = 0.0;
    for (uint32_t i = 0; i < NB; ++i) txomeNorm += h_hist[i];
    std::vector<double> ratio(NB);
    if (seq) {
        uint32_t tc = 0;                                          // ReadKmerDist::totalCount() accumulates in CountT (uint32)
        for (uint32_t i = 0; i < NB; ++i) tc += m->read_bias[i];
        readNorm = static_cast<double>(tc);
        const double pmass = static_cast<double>(BNK);
        const double prior = ((pmass / (readNorm - pmass)) * txomeNorm) / pmass;
        for (uint32_t i = 0; i < NB; ++i) ratio[i] = m->read_bias[i] / (h_hist[i] + prior);
    } else {
        for (uint32_t i = 0; i < NB; ++i) readNorm += m->observed_gc[i];
        const double pmass = 101.0;
        const double prior = ((pmass / (readNorm - pmass)) * txomeNorm) / 101.0;
        for (uint32_t i = 0; i < NB; ++i) ratio[i] = m->observed_gc[i] / (prior + h_hist[i]);
    }
    BIAS_CUDA(cudaMemcpyAsync(d_hist.p + NB, ratio.data(), NB * 8ull, cudaMemcpyHostToDevice, s));
    // ---- pass 2
    double* d_out = d_vec.p + 3ull * T;
    const double norm = txomeNorm / readNorm;
    if (slide) k_bias_efflen_gc_slide<<<grid, SLIDE_THREADS, 0, s>>>(v, d_hist.p + NB, d_w.p + n_fl, n_fl, norm, d_out);
    else if (seq) k_bias_efflen<1><<<grid, BIAS_THREADS, 0, s>>>(v, d_hist.p + NB, norm, d_out);
    else k_bias_efflen<2><<<grid, BIAS_THREADS, 0, s>>>(v, d_hist.p + NB, norm, d_out);
    c->launches++;
    BIAS_CUDA(cudaGetLastError());
    BIAS_CUDA(cudaMemcpyAsync(eff_out, d_out, T * 8ull, cudaMemcpyDeviceToHost, s));
    BIAS_CUDA(cudaStreamSynchronize(s));
#undef BIAS_CUDA
    cleanup();
    return SFB200_OK;
}
