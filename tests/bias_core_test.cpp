// CPU check of the per-position arithmetic of the device-side bias / GC correction (sailfish_b200/csrc/bias_core.inl, the text
// bias.cu compiles for the GPU): 2-bit text windows, 6-mer context indices, GC prefix counts, the two passes.  The kernels' loop
// structure is replayed serially here (transcripts, positions, histogram, normalisers, per-transcript sums) and the result is
// compared with the CPU oracle (oracle/orc_bias.cpp, pinned to the reference's own updateEffectiveLengths).  Built and run by
// tests/test_bias_core.py; links oracle/liboracle.so -- test infrastructure on both sides.
#include <stdint.h>
#include <stddef.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#define SFB_BD static inline
#define SFB_LDG(p) (*(p))
#define SFB_POPC64(x) __builtin_popcountll(x)
#define SFB_D2I_RN(x) ((int32_t)std::lrint(x))
#include "../sailfish_b200/csrc/bias_core.inl"

extern "C" int orc_update_eff_lens(int mode, uint32_t T, const char* seq, const uint64_t* off, const uint32_t* len, const double* eff_model,
                                   const double* eff_in, const double* alphas, int64_t num_fwd, int64_t num_rc, const uint32_t* read_bias,
                                   const uint32_t* observed_gc, const uint32_t* fld_counts, uint32_t n_fld, uint32_t gc_samp, double* eff_out);
struct orc_index;
extern "C" orc_index* orc_index_build(const char* seq, const uint64_t* txp_off, const uint32_t* txp_len, uint32_t n_txp, int k, int n_threads);
extern "C" void orc_index_free(orc_index*);
extern "C" int32_t orc_bias_context_index(const orc_index*, uint32_t tid, int32_t pos, int fwd, uint32_t read_len);
extern "C" int32_t orc_gc_frac(const orc_index*, uint32_t tid, int32_t s, int32_t e);
extern "C" uint32_t orc_fld_cdf(const uint32_t* fld_counts, uint32_t n_fld, float* cdf_out, uint32_t cap, uint32_t* max_value);

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

static int run(int mode, uint32_t gc_samp, uint64_t seed, bool slide = false) {
    std::mt19937_64 rng(seed);
    const uint32_t T = 90;
    std::vector<uint32_t> len(T);
    std::vector<uint64_t> off(T), tstart(T + 1);
    std::string seq;
    for (uint32_t t = 0; t < T; ++t) {
        len[t] = t < 3 ? 4 + 2 * t : (t < 8 ? 60 + 20 * t : 200 + (uint32_t)(rng() % 1500));            // a few transcripts shorter than the 6-mer window
        off[t] = seq.size(); tstart[t] = seq.size();
        for (uint32_t i = 0; i < len[t]; ++i) seq.push_back("ACGT"[(rng() % 10) < 3 ? 0 : (rng() % 4)]);
    }
    tstart[T] = seq.size();
    // the device index's text: 2 bits per base, base p at bits 2*(p%32) of word p/32, two words of slack
    std::vector<uint64_t> words(seq.size() / 32 + 2, 0);
    for (size_t p = 0; p < seq.size(); ++p) { const uint64_t c = seq[p] == 'A' ? 0 : seq[p] == 'C' ? 1 : seq[p] == 'G' ? 2 : 3; words[p >> 5] |= c << (2 * (p & 31)); }
    std::vector<uint32_t> gcw(words.size() + 1, 0);
    for (size_t w = 0; w < words.size(); ++w) gcw[w + 1] = gcw[w] + (uint32_t)__builtin_popcountll((words[w] ^ (words[w] >> 1)) & 0x5555555555555555ULL);
    // helpers against the plain string
    for (int k = 0; k < 2000; ++k) {
        const uint32_t t = 3 + (uint32_t)(rng() % (T - 3));
        const uint32_t i = (uint32_t)(rng() % (len[t] - 6));
        const uint32_t win = b_win6(words.data(), tstart[t] + i);
        uint32_t f = 0, r = 0;
        for (int j = 0; j < 6; ++j) { const uint32_t c = (uint32_t)std::string("ACGT").find(seq[off[t] + i + j]); f = (f << 2) | c; }
        for (int j = 5; j >= 0; --j) { const uint32_t c = 3 - (uint32_t)std::string("ACGT").find(seq[off[t] + i + j]); r = (r << 2) | c; }
        CHECK(b_idx_fwd(win) == f && b_idx_rc(win) == r, "6-mer index at transcript %u position %u", t, i);
    }
    // the mapper-side helpers against the oracle's per-hit functions (positions on, off and around both transcript ends)
    {
        orc_index* oix = orc_index_build(seq.data(), off.data(), len.data(), T, 31, 1);
        CHECK(oix != nullptr, "oracle index");
        size_t n_valid = 0, n_gc = 0;
        for (int k = 0; k < 40000; ++k) {
            const uint32_t t = (uint32_t)(rng() % T);
            const uint32_t rl = 20 + (uint32_t)(rng() % 130);
            const int L = (int)len[t];
            const int32_t pos = (k & 3) == 0 ? (int32_t)(rng() % 12) - 6 : (k & 3) == 1 ? L - (int32_t)(rng() % (rl + 12)) : (int32_t)(rng() % (uint32_t)(L + 40)) - 20;
            const bool fwd = (rng() & 1) != 0;
            const int32_t got_i = b_read_start_index(words.data(), tstart[t], L, pos, fwd, rl);
            const int32_t want_i = orc_bias_context_index(oix, t, pos, fwd ? 1 : 0, rl);
            CHECK(got_i == want_i, "read-start context: transcript %u (len %d) pos %d fwd %d readLen %u: %d vs %d", t, L, pos, (int)fwd, rl, got_i, want_i);
            n_valid += got_i >= 0;
            if (L > 3) {
                const int32_t s0 = 1 + (int32_t)(rng() % (uint32_t)(L - 2)), e0 = s0 + (int32_t)(rng() % (uint32_t)(L - s0 - 1 + 1));
                if (s0 > 0 && e0 < L && e0 >= s0) {
                    CHECK(b_gc_frac_range(words.data(), tstart[t], s0, e0) == orc_gc_frac(oix, t, s0, e0), "gcFrac transcript %u (%d, %d]", t, s0, e0);
                    CHECK(b_gc_frac_range(words.data(), tstart[t], s0, e0) == b_gc_frac(BiasView{words.data(), tstart.data(), len.data(), gcw.data()}, tstart[t], s0, e0), "gcFrac (prefix form) transcript %u", t);
                    ++n_gc;
                }
            }
        }
        orc_index_free(oix);
        CHECK(n_valid > 5000 && n_gc > 5000, "too few helper cases: %zu contexts, %zu intervals", n_valid, n_gc);
    }
    std::vector<uint32_t> fld(1000);
    for (uint32_t x = 0; x < 1000; ++x) fld[x] = (uint32_t)std::lround(30000.0 * std::exp(-0.5 * std::pow((x - 190.0) / 30.0, 2)));
    std::vector<float> cdf(1001);
    uint32_t fld_max = 0;
    const uint32_t n_cdf = orc_fld_cdf(fld.data(), 1000, cdf.data(), 1001, &fld_max);
    std::vector<double> eff_model(T), eff_in(T), alphas(T), want(T), got(T);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (uint32_t t = 0; t < T; ++t) {
        eff_model[t] = len[t] > 189 ? len[t] - 189.0 : len[t];
        eff_in[t] = eff_model[t] * (0.9 + 0.2 * U(rng));
        alphas[t] = U(rng) < 0.2 ? 0.0 : std::exp(3 + 2 * (U(rng) - 0.5) * 3);
    }
    std::vector<uint32_t> rb(4096), og(101);
    for (auto& x : rb) x = 1 + (uint32_t)(rng() % 3000);
    for (auto& x : og) x = 1 + (uint32_t)(rng() % 8000);
    const int64_t nf = 61234, nr = 58766;
    CHECK(orc_update_eff_lens(mode, T, seq.data(), off.data(), len.data(), eff_model.data(), eff_in.data(), alphas.data(), nf, nr, rb.data(), og.data(),
                              fld.data(), 1000, gc_samp, want.data()) == 0, "oracle failed");
    // ---- what sfb200_bias_eff_lens does, serially
    BiasView v;
    v.words = words.data(); v.txp_start = tstart.data(); v.txp_len = len.data(); v.gcw = gcw.data();
    v.cdf = cdf.data(); v.n_cdf = n_cdf; v.eff_model = eff_model.data(); v.eff_in = eff_in.data(); v.alphas = alphas.data(); v.T = T;
    v.probFwd = (double)nf / (nf + nr); v.probRC = (double)nr / (nf + nr);
    v.fldLow = 0; v.fldHigh = 1; v.gcSamp = (int32_t)gc_samp;
    if (mode == 2) {
        bool first = false, second = false;
        for (uint32_t i = 0; i <= fld_max; ++i) {
            const float d = i < n_cdf ? cdf[i] : 1.0f;
            if (!first && d >= 0.005) { first = true; v.fldLow = (int32_t)i; }
            if (!second && d >= 0.995) { second = true; v.fldHigh = (int32_t)i; }
        }
    }
    const uint32_t NB = mode == 1 ? BNK : 101u;
    std::vector<double> hist(NB, 1.0), ratio(NB);
    auto add = [&](uint32_t bin, double x) { hist[bin] += x; };
    // sliding form (k_bias_expected_gc_slide / k_bias_efflen_gc_slide): the sampled fragment lengths and their weights
    std::vector<int32_t> fls;
    std::vector<double> w, wf;
    if (slide) {
        CHECK(mode == 2 && v.fldLow >= 1, "sliding form needs mode 2 and fldLow >= 1");
        double prev = b_cdf(v, 0);
        for (int32_t fl = v.fldLow; fl <= v.fldHigh; fl += v.gcSamp) {
            const double cur = b_cdf(v, fl);
            fls.push_back(fl); w.push_back(cur - prev); wf.push_back((cur - prev) * v.probFwd + (cur - prev) * v.probRC);
            prev = cur;
        }
        for (int k = 0; k < 3000; ++k) {                                        // the integer rounding against lrint of the double quotient
            const uint32_t fl = 1 + (uint32_t)(rng() % 1200), n = (uint32_t)(rng() % fl);
            CHECK(b_gc_bin(n, fl) == (int32_t)std::lrint((100.0 * n) / (double)fl), "gc bin n %u fl %u", n, fl);
        }
        for (uint32_t fl = 1; fl <= 1000; ++fl) for (uint32_t n = 0; n < fl; ++n) if ((200 * n) % (2 * fl) == fl) CHECK(b_gc_bin(n, fl) == (int32_t)std::lrint((100.0 * n) / (double)fl), "tie n %u fl %u", n, fl);
    }
    for (uint32_t t = 0; t < T && slide; ++t) {
        int32_t refLen, unproc;
        if (!b_eligible(v, t, refLen, unproc)) continue;
        const double contribution = alphas[t] / eff_in[t];
        std::vector<double> tsum(101, 0.0);
        for (size_t k = 0; k < fls.size(); ++k) {                                // one thread per fragment length: integer counts per bin
            uint32_t cnt[101] = {0};
            b_gc_slide(words.data(), tstart[t], refLen, fls[k], [&](int32_t bin) { cnt[bin]++; });
            for (int b = 0; b < 101; ++b) if (cnt[b]) tsum[b] += w[k] * cnt[b];
        }
        for (int b = 0; b < 101; ++b) hist[b] += contribution * tsum[b];
    }
    for (uint32_t t = 0; t < T && !slide; ++t) {
        int32_t refLen, unproc;
        if (!b_eligible(v, t, refLen, unproc)) continue;
        const double contribution = alphas[t] / eff_in[t];
        for (int32_t i = 0; i <= refLen - BK - 1; ++i) {
            if (mode == 1) b_expected_seq(v, tstart[t], refLen, i, contribution, add); else b_expected_gc(v, tstart[t], refLen, i, contribution, add);
        }
    }
    double txomeNorm = 0.0, readNorm = 0.0;
    for (double h : hist) txomeNorm += h;
    if (mode == 1) {
        uint32_t tc = 0; for (uint32_t x : rb) tc += x; readNorm = tc;
        const double prior = ((4096.0 / (readNorm - 4096.0)) * txomeNorm) / 4096.0;
        for (uint32_t i = 0; i < NB; ++i) ratio[i] = rb[i] / (hist[i] + prior);
    } else {
        for (uint32_t x : og) readNorm += x;
        const double prior = ((101.0 / (readNorm - 101.0)) * txomeNorm) / 101.0;
        for (uint32_t i = 0; i < NB; ++i) ratio[i] = og[i] / (prior + hist[i]);
    }
    size_t changed = 0;
    for (uint32_t t = 0; t < T; ++t) {
        int32_t refLen, unproc;
        const bool go = b_eligible(v, t, refLen, unproc);
        double sum = 0.0;
        if (go && slide) {
            for (size_t k = 0; k < fls.size(); ++k) {
                double acc = 0.0;
                b_gc_slide(words.data(), tstart[t], refLen, fls[k], [&](int32_t bin) { acc += ratio[bin]; });
                sum += wf[k] * acc;
            }
        } else if (go) for (int32_t i = 0; i <= refLen - BK - 1; ++i) sum += mode == 1 ? b_eff_seq(v, ratio.data(), tstart[t], refLen, i) : b_eff_gc(v, ratio.data(), tstart[t], refLen, i);
        const double eff = sum * (txomeNorm / readNorm);
        got[t] = (go && unproc > 0 && eff > (double)unproc) ? eff : eff_in[t];
        CHECK(std::fabs(got[t] - want[t]) <= 1e-10 * std::fabs(want[t]), "mode %d transcript %u: %.17g vs oracle %.17g", mode, t, got[t], want[t]);
        changed += got[t] != eff_in[t];
    }
    CHECK(changed >= 15, "only %zu transcripts were corrected", changed);
    return 0;
}

int main() {
    if (run(1, 1, 1) || run(2, 1, 2) || run(2, 3, 3) || run(1, 1, 4)) return 1;
    if (run(2, 1, 5, true) || run(2, 3, 6, true) || run(2, 7, 7, true)) return 1;                    // the sliding form of the GC passes
    printf("bias core ok\n");
    return 0;
}
