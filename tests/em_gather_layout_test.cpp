// CPU check of the gather layout of the EM loop (sailfish_b200/csrc/em_gather_build.inl) and of the beta / r formulation the
// GPU kernel k_em_gather iterates (sailfish_b200/csrc/em_gather.cuh).  The build body is compiled here as a single-thread host
// function (one "thread" per CTA, atomics are plain updates); the E-step / M-step loops below walk the layout exactly as the
// kernel's lanes do and are compared with the update written in the reference's shape (CollapsedEMOptimizer.cpp:235-277, :760-769).
// Built and run by tests/test_em_gather_layout.py.
#include <stdint.h>
#include <stddef.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <vector>

static inline uint32_t gb_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
static inline void gb_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
#define SFB_GB_FN static
#define SFB_GB_TID 0u
#define SFB_GB_NT 1u
#define SFB_GB_SYNC() do { } while (0)
#define SFB_GB_ADD(p, v) gb_add((p), (v))
#define SFB_GB_MAX(p, v) gb_max((p), (v))
#include "../sailfish_b200/csrc/em_gather_build.inl"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

struct Slice {
    uint32_t c_lo, nc, t0, nt;
    std::vector<uint32_t> start, len, lab;      // partition arrays (global indexing: class c_lo + c)
};

static Slice make_slice(std::mt19937_64& rng, uint32_t nc, uint32_t nt, uint32_t max_len, bool dups) {
    Slice s;
    s.c_lo = 37; s.nc = nc; s.t0 = 1000; s.nt = nt;
    s.start.assign(s.c_lo + nc, 0); s.len.assign(s.c_lo + nc, 0);
    s.lab.assign(123, 0xFFFFFFFFu);               // entries of other CTAs in front
    for (uint32_t c = 0; c < nc; ++c) {
        uint32_t n = 2 + (uint32_t)(rng() % 3);
        const uint32_t roll = (uint32_t)(rng() % 100);
        if (roll < 10) n = 2 + (uint32_t)(rng() % 14);
        if (roll < 2) n = 2 + (uint32_t)(rng() % (max_len - 1));
        n = std::min(n, max_len);
        if (!dups) n = std::min(n, nt);
        s.start[s.c_lo + c] = (uint32_t)s.lab.size(); s.len[s.c_lo + c] = n;
        // members from a window of the range (gene locality), the upper 20% of the range is never used (degree 0)
        const uint32_t usable = std::max<uint32_t>(1, nt - nt / 5);
        std::vector<uint32_t> m;
        while (m.size() < n) {
            const uint32_t t = (uint32_t)(rng() % usable);
            if (!dups && std::find(m.begin(), m.end(), t) != m.end()) { if (usable <= m.size()) break; continue; }
            m.push_back(t);
        }
        std::sort(m.begin(), m.end());
        s.len[s.c_lo + c] = (uint32_t)m.size();
        for (uint32_t t : m) s.lab.push_back(s.t0 + t);
    }
    return s;
}

static int run_case(uint64_t seed, uint32_t nc, uint32_t nt, uint32_t max_len, bool dups, bool vb, bool scaled) {
    std::mt19937_64 rng(seed);
    Slice s = make_slice(rng, nc, nt, max_len, dups);
    uint64_t ne = 0;
    for (uint32_t c = 0; c < nc; ++c) ne += s.len[s.c_lo + c];
    const GatherGeom g = gather_make_geom(nc, ne, nt, scaled);
    const uint32_t sh = g.shift;
    CHECK(sh == ((scaled && ((nc + 31) / 32 * 32) <= 8191 && ((nt + 31) / 32 * 32) <= 8191) ? 3u : 0u), "shift");
    CHECK(g.region_words % 4 == 0 && g.o_lab_e % 4 == 0 && g.o_cls_t % 4 == 0 && g.o_cperm % 4 == 0 && g.o_tmap % 4 == 0, "geometry alignment");
    std::vector<uint32_t> region(g.region_words + 8, 0xDEADBEEFu), scratch(gather_scratch_words(nc, nt, g) + 8, 0xABABABABu);
    const uint32_t guard_r = 0x13572468u;
    for (int i = 0; i < 8; ++i) region[g.region_words + i] = guard_r;
    gather_build_cta(s.start.data(), s.len.data(), s.lab.data(), s.c_lo, nc, s.t0, nt, g, region.data(), scratch.data());
    for (int i = 0; i < 8; ++i) CHECK(region[g.region_words + i] == guard_r, "region overrun");
    for (int i = 0; i < 8; ++i) CHECK(scratch[gather_scratch_words(nc, nt, g) + i] == 0xABABABABu, "scratch overrun");
    const uint32_t* h = region.data();
    CHECK(h[GH_OK] == 1, "layout flagged not ok (ent_e %u ent_t %u cap %u)", h[GH_ENT_E], h[GH_ENT_T], g.cap_ent);
    CHECK(h[GH_NC] == nc && h[GH_NT] == nt, "header counts");
    const uint32_t tiles_e = h[GH_TILES_E], tiles_t = h[GH_TILES_T], nc_pad = tiles_e * 32, nt_pad = tiles_t * 32;
    CHECK(tiles_e == (nc + 31) / 32 && tiles_t == (nt + 31) / 32, "tile counts");
    CHECK(h[GH_ENT_E] <= g.cap_ent && h[GH_ENT_T] <= g.cap_ent, "entries exceed the region");
    const uint32_t* eoff = h + g.o_tile_e_off; const uint32_t* elen = h + g.o_tile_e_len;
    const uint32_t* toff = h + g.o_tile_t_off; const uint32_t* tlen = h + g.o_tile_t_len;
    const uint32_t* cperm = h + g.o_cperm; const uint32_t* tmap = h + g.o_tmap;
    const uint16_t* lab_e = reinterpret_cast<const uint16_t*>(h + g.o_lab_e);
    const uint16_t* cls_t = reinterpret_cast<const uint16_t*>(h + g.o_cls_t);
    // permutations
    std::vector<int> seen_c(nc, 0), seen_t(nt, 0);
    std::vector<uint32_t> tnew_of(nt);
    for (uint32_t i = 0; i < nc; ++i) { CHECK(cperm[i] >= s.c_lo && cperm[i] < s.c_lo + nc, "cperm range"); seen_c[cperm[i] - s.c_lo]++; }
    for (uint32_t i = 0; i < nt; ++i) { CHECK(tmap[i] >= s.t0 && tmap[i] < s.t0 + nt, "tmap range"); seen_t[tmap[i] - s.t0]++; tnew_of[tmap[i] - s.t0] = i; }
    for (uint32_t i = 0; i < nc; ++i) CHECK(seen_c[i] == 1, "cperm is not a permutation");
    for (uint32_t i = 0; i < nt; ++i) CHECK(seen_t[i] == 1, "tmap is not a permutation");
    // tile offsets are a prefix sum of 32 * L
    { uint32_t acc = 0; for (uint32_t k = 0; k < tiles_e; ++k) { CHECK(eoff[k] == acc, "class tile offset"); acc += 32 * elen[k]; } CHECK(acc == h[GH_ENT_E], "class entries"); }
    { uint32_t acc = 0; for (uint32_t k = 0; k < tiles_t; ++k) { CHECK(toff[k] == acc, "transcript tile offset"); acc += 32 * tlen[k]; } CHECK(acc == h[GH_ENT_T], "transcript entries"); }
    // class columns hold the members in label order, then the sentinel
    std::vector<std::multiset<uint32_t>> want_t(nt);   // per new transcript index: new class indices
    static_assert(sizeof(uint16_t) == 2, "");
    for (uint32_t i = 0; i < nc_pad; ++i) {
        const uint32_t k = i >> 5, lane = i & 31;
        uint32_t n = 0, b = 0;
        if (i < nc) { n = s.len[cperm[i]]; b = s.start[cperm[i]]; CHECK(n <= elen[k], "tile L below a member count"); }
        for (uint32_t j = 0; j < elen[k]; ++j) {
            const uint32_t raw = lab_e[eoff[k] + 32 * j + lane], v = raw >> sh;
            CHECK((v << sh) == raw, "stored member index is not a multiple of the scale");
            if (j < n) { CHECK(v < nt && tmap[v] == s.lab[b + j], "class %u member %u", i, j); want_t[v].insert(i); }
            else CHECK(v == nt_pad, "class %u padding %u holds %u", i, j, v);
        }
    }
    // sizes are non-increasing from tile to tile (this is what bounds the padding)
    // (sizes above 255 share the last bucket and stay unsorted among themselves)
    if (h[GH_MAXLEN] < GB_BUCKETS) for (uint32_t k = 1; k < tiles_e; ++k) CHECK(elen[k] <= elen[k - 1], "class tiles not sorted");
    if (h[GH_MAXDEG] < GB_BUCKETS) for (uint32_t k = 1; k < tiles_t; ++k) CHECK(tlen[k] <= tlen[k - 1], "transcript tiles not sorted");
    for (uint32_t i = 0; i < nt_pad; ++i) {
        const uint32_t k = i >> 5, lane = i & 31;
        std::multiset<uint32_t> got;
        for (uint32_t j = 0; j < tlen[k]; ++j) {
            const uint32_t raw = cls_t[toff[k] + 32 * j + lane], v = raw >> sh;
            CHECK((v << sh) == raw, "stored class index is not a multiple of the scale");
            if (v == nc_pad) continue;
            CHECK(v < nc, "transcript %u holds class %u", i, v);
            got.insert(v);
        }
        if (i < nt) CHECK(got == want_t[i], "transposed list of transcript %u differs", i);
        else CHECK(got.empty(), "padding transcript %u has classes", i);
    }

    // ---- iterate: the kernel's loops on the layout against the update in the reference's shape
    const uint32_t T = s.t0 + nt + 50;
    std::vector<double> eff(T), single(T, 0.0), alpha(T, 0.0), cnt(s.c_lo + nc, 0.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (uint32_t t = 0; t < T; ++t) { eff[t] = 1.0 + 3000.0 * U(rng); if (U(rng) < 0.3) single[t] = std::floor(100.0 * U(rng)); }
    for (uint32_t c = 0; c < nc; ++c) cnt[s.c_lo + c] = std::floor(1.0 + 50.0 * U(rng) * U(rng));
    const double prior = vb ? 0.01 : 0.0;
    for (uint32_t t = s.t0; t < s.t0 + nt; ++t) alpha[t] = 7.5;
    // reference-shaped state
    std::vector<double> a_ref(alpha);
    // kernel-shaped state
    std::vector<double> s_r(nc_pad + 2, 0.0), s_cnt(nc_pad + 2, 0.0), s_beta(nt_pad + 2, 0.0), s_alpha(nt_pad + 2, 0.0), s_base(nt_pad + 2, 0.0),
        s_inveff(nt_pad + 2, 0.0);
    for (uint32_t i = 0; i < nc; ++i) s_cnt[i] = cnt[cperm[i]];
    for (uint32_t i = 0; i < nt; ++i) { const uint32_t t = tmap[i]; s_alpha[i] = alpha[t]; s_base[i] = single[t] + prior; s_inveff[i] = 1.0 / eff[t]; }
    auto digamma = [](double x) {
        double acc = 0.0;
        while (x < 12.0) { acc -= 1.0 / x; x += 1.0; }
        const double inv = 1.0 / x, inv2 = inv * inv;
        return acc + std::log(x) - 0.5 * inv - inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0)));
    };
    for (int it = 0; it < 25; ++it) {
        // reference shape: weights w_i = (cnt/eff_i) / sum, denom = sum theta_i w_i, out_i += theta_i w_i cnt / denom
        std::vector<double> theta(a_ref), out(T, 0.0);
        if (vb) {
            double sum = 0.0; for (uint32_t t = s.t0; t < s.t0 + nt; ++t) sum += a_ref[t];
            const double ln = digamma(sum);
            for (uint32_t t = s.t0; t < s.t0 + nt; ++t) theta[t] = a_ref[t] > 0 ? std::exp(digamma(a_ref[t]) - ln) : 0.0;
        }
        for (uint32_t t = s.t0; t < s.t0 + nt; ++t) out[t] = single[t] + prior;
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t b = s.start[s.c_lo + c], n = s.len[s.c_lo + c];
            std::vector<double> w(n); double ws = 0.0;
            for (uint32_t j = 0; j < n; ++j) { w[j] = cnt[s.c_lo + c] / eff[s.lab[b + j]]; ws += w[j]; }
            double denom = 0.0;
            for (uint32_t j = 0; j < n; ++j) { w[j] *= 1.0 / ws; denom += theta[s.lab[b + j]] * w[j]; }
            if (!(denom > 0.0)) continue;
            const double inv = cnt[s.c_lo + c] / denom;
            for (uint32_t j = 0; j < n; ++j) out[s.lab[b + j]] += theta[s.lab[b + j]] * w[j] * inv;
        }
        for (uint32_t t = s.t0; t < s.t0 + nt; ++t) a_ref[t] = out[t];
        // kernel shape
        {
            double ln = 0.0;
            if (vb) { double sum = 0.0; for (uint32_t i = 0; i < nt_pad; ++i) sum += s_alpha[i]; ln = digamma(sum); }
            for (uint32_t i = 0; i < nt_pad; ++i) {
                const double a = s_alpha[i];
                s_beta[i] = (vb ? (a > 0 ? std::exp(digamma(a) - ln) : 0.0) : a) * s_inveff[i];
            }
        }
        for (uint32_t k = 0; k < tiles_e; ++k)
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint16_t* col = lab_e + eoff[k] + lane;
                double S = 0.0;
                for (uint32_t j = 0; j < elen[k]; ++j) S += s_beta[col[j << 5] >> sh];
                s_r[(k << 5) + lane] = S > 0.0 ? s_cnt[(k << 5) + lane] / S : 0.0;
            }
        for (uint32_t k = 0; k < tiles_t; ++k)
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint16_t* col = cls_t + toff[k] + lane;
                double acc = 0.0;
                for (uint32_t j = 0; j < tlen[k]; ++j) acc += s_r[col[j << 5] >> sh];
                const uint32_t i = (k << 5) + lane;
                s_alpha[i] = s_beta[i] * acc + s_base[i];
            }
        for (uint32_t i = 0; i < nt; ++i) {
            const double want = a_ref[tmap[i]], got = s_alpha[i];
            CHECK(std::fabs(want - got) <= 1e-9 * std::max(1.0, std::fabs(want)), "iteration %d transcript %u: %.17g vs %.17g", it, tmap[i], got, want);
        }
        for (uint32_t i = nt; i < nt_pad; ++i) CHECK(s_alpha[i] == 0.0, "padding transcript became %g", s_alpha[i]);
    }
    return 0;
}

int main() {
    struct Case { uint64_t seed; uint32_t nc, nt, max_len; bool dups, vb; };
    const Case cases[] = {
        {1, 1381, 676, 40, false, false},   // a cfg2-sized CTA slice
        {2, 1381, 676, 40, true, true},
        {3, 1, 1, 2, true, false},          // one class {t, t}
        {4, 31, 32, 8, false, false},
        {5, 33, 31, 8, false, true},
        {6, 64, 64, 200, true, false},      // maxReadOccs-long classes
        {7, 5000, 3000, 300, true, false},  // sizes beyond the 255 bucket
        {8, 0, 17, 2, false, false},        // no classes at all
        {9, 200, 5, 6, true, false},        // degrees far above the bucket cap on few transcripts
        {10, 3000, 40, 30, true, true},
        {11, 8160, 8191, 12, false, false}, // the largest sizes the scaled (byte-offset) indices can address
        {12, 8161, 100, 12, false, false},  // one class more: falls back to plain indices
    };
    for (int scaled = 0; scaled < 2; ++scaled)
    for (const Case& c : cases) {
        if (run_case(c.seed, c.nc, c.nt, c.max_len, c.dups, c.vb, scaled != 0)) { fprintf(stderr, "case seed %llu failed\n", (unsigned long long)c.seed); return 1; }
    }
    printf("em_gather layout ok (%zu cases)\n", sizeof(cases) / sizeof(cases[0]));
    return 0;
}
