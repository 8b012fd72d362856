// CPU check of the dense-component layout of the EM loop (sailfish_b200/csrc/em_dense_build.inl) and of the per-component
// iteration k_em_dense runs on it (sailfish_b200/csrc/em_dense.cuh).  The build body is compiled here as a single-thread host
// function; the loops below walk the layout as the kernel's lanes do (one lane = one component) and are compared with the update
// written in the reference's shape (CollapsedEMOptimizer.cpp:235-277, :760-769).  Built and run by tests/test_em_gather_layout.py.
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
static inline void gb_min(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
#define SFB_GB_FN static
#define SFB_GB_TID 0u
#define SFB_GB_NT 1u
#define SFB_GB_SYNC() do { } while (0)
#define SFB_GB_ADD(p, v) gb_add((p), (v))
#define SFB_GB_MAX(p, v) gb_max((p), (v))
#define SFB_GB_MIN(p, v) gb_min((p), (v))
#include "../sailfish_b200/csrc/em_dense_build.inl"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

struct Slice {
    uint32_t c_lo, nc, t0, nt;
    std::vector<uint32_t> start, len, lab;
};

// genes of `gsize` consecutive transcripts (every `idle_every`-th gene unused); classes are subsets of one gene;
// chain = true links consecutive transcripts pairwise only, so that the label propagation needs several rounds
static Slice make_slice(std::mt19937_64& rng, uint32_t n_genes, uint32_t gsize, uint32_t classes_per_gene, bool chain, bool dup, uint32_t idle_every) {
    Slice s;
    s.c_lo = 11; s.t0 = 500; s.nt = n_genes * gsize + 3;            // three trailing transcripts in no class
    s.start.assign(s.c_lo, 0); s.len.assign(s.c_lo, 0);
    s.lab.assign(57, 0xFFFFFFFFu);
    std::set<std::vector<uint32_t>> seen;
    for (uint32_t g = 0; g < n_genes; ++g) {
        if (idle_every && g % idle_every == idle_every - 1) continue;
        const uint32_t base = g * gsize;
        if (chain) {
            for (uint32_t j = 0; j + 1 < gsize; ++j) { s.start.push_back((uint32_t)s.lab.size()); s.len.push_back(2); s.lab.push_back(s.t0 + base + j); s.lab.push_back(s.t0 + base + j + 1); }
            continue;
        }
        const uint32_t ncls = 1 + (uint32_t)(rng() % classes_per_gene);
        for (uint32_t c = 0; c < ncls; ++c) {
            std::vector<uint32_t> m;
            const uint32_t n = 2 + (uint32_t)(rng() % (gsize - 1));
            while (m.size() < n) { const uint32_t t = base + (uint32_t)(rng() % gsize); if (dup || std::find(m.begin(), m.end(), t) == m.end()) m.push_back(t); }
            std::sort(m.begin(), m.end());
            if (!seen.insert(m).second) continue;
            s.start.push_back((uint32_t)s.lab.size()); s.len.push_back((uint32_t)m.size());
            for (uint32_t t : m) s.lab.push_back(s.t0 + t);
        }
    }
    s.nc = (uint32_t)s.start.size() - s.c_lo;
    return s;
}

static int run_case(uint64_t seed, uint32_t n_genes, uint32_t gsize, uint32_t cpg, bool chain, bool dup, uint32_t idle_every, bool expect_dense, bool vb, uint32_t G) {
    std::mt19937_64 rng(seed);
    Slice s = make_slice(rng, n_genes, gsize, cpg, chain, dup, idle_every);
    const uint32_t nc = s.nc, nt = s.nt;
    const DenseGeom g = dense_make_geom(nc, nt, G);
    CHECK(g.group == G, "group");
    CHECK(g.region_words % 4 == 0 && g.o_mask % 4 == 0 && g.o_tmap % 4 == 0 && g.o_idle % 4 == 0 && g.o_cperm % 4 == 0, "geometry alignment");
    std::vector<uint32_t> region(g.region_words + 8, 0xDEADBEEFu), scratch(dense_scratch_words(nt, g) + 8, 0xABABABABu);
    for (int i = 0; i < 8; ++i) region[g.region_words + i] = 0x13572468u;
    dense_build_cta(s.start.data(), s.len.data(), s.lab.data(), s.c_lo, nc, s.t0, nt, g, region.data(), scratch.data());
    for (int i = 0; i < 8; ++i) CHECK(region[g.region_words + i] == 0x13572468u, "region overrun");
    for (int i = 0; i < 8; ++i) CHECK(scratch[dense_scratch_words(nt, g) + i] == 0xABABABABu, "scratch overrun");
    const uint32_t* h = region.data();
    CHECK((h[DH_KIND] == 1) == expect_dense, "kind %u, expected %d", h[DH_KIND], (int)expect_dense);
    if (!expect_dense) return 0;
    const uint32_t ncomp = h[DH_NCOMP], tiles = h[DH_TILES], ent = h[DH_ENT], NS = h[DH_NS], nidle = h[DH_NIDLE], ncomp_pad = G ? tiles * 32 / G : (ncomp + 31) / 32 * 32;
    const uint32_t* lane_tab = h + g.o_lane;                              // balanced layout (G == 0): component | log2(lanes) << 16 | rank << 20 per lane
    auto comp_of = [&](uint32_t k, uint32_t lane) -> uint32_t { if (G) return k * (32 / G) + lane / G; const uint32_t i = lane_tab[32 * k + lane]; return i == DN_NONE ? DN_NONE : (i & 0xFFFFu); };
    auto lanes_of = [&](uint32_t k, uint32_t lane) -> uint32_t { if (G) return G; const uint32_t i = lane_tab[32 * k + lane]; return i == DN_NONE ? 1u : 1u << ((i >> 16) & 0xFu); };
    auto rank_of = [&](uint32_t k, uint32_t lane) -> uint32_t { if (G) return lane % G; const uint32_t i = lane_tab[32 * k + lane]; return i == DN_NONE ? 1u : i >> 20; };
    CHECK(h[DH_GROUP] == G, "header group");
    CHECK(ncomp == 0 ? NS == 0 : (NS >= 2 && NS <= DN_MAX_SLOTS && NS <= gsize), "slots %u", NS);
    CHECK((G == 0 || tiles == (ncomp * G + 31) / 32) && ent <= g.cap_ent && tiles <= g.cap_tiles, "tiles / entries");
    if (G == 0) {
        // every component owns an aligned run of 1, 2, 4 or 8 lanes with ranks 0..lanes-1, enough for <= DN_ROWS classes per lane (up to 32 classes)
        std::vector<uint32_t> seen_lanes(ncomp, 0);
        for (uint32_t i = 0; i < tiles * 32; ++i) {
            const uint32_t info = lane_tab[i];
            if (info == DN_NONE) continue;
            const uint32_t q = info & 0xFFFFu, gl = 1u << ((info >> 16) & 0xFu), r = info >> 20;
            CHECK(q < ncomp && gl <= DN_MAX_GROUP && r < gl && (i - r) % gl == 0, "lane %u: component %u, %u lanes, rank %u", i, q, gl, r);
            CHECK((lane_tab[i - r] & 0xFFFFu) == q, "lane %u is not in its component's run", i);
            seen_lanes[q] += 1;
        }
        for (uint32_t q = 0; q < ncomp; ++q) CHECK(seen_lanes[q] >= 1, "component %u has no lane", q);
    }
    const uint32_t* toff = h + g.o_tile_off; const uint32_t* tlen = h + g.o_tile_len; const uint32_t* cperm = h + g.o_cperm;
    const uint8_t* mask = reinterpret_cast<const uint8_t*>(h + g.o_mask);
    const uint32_t* tmap = h + g.o_tmap; const uint32_t* idle = h + g.o_idle;
    { uint32_t acc = 0; for (uint32_t k = 0; k < tiles; ++k) { CHECK(toff[k] == acc, "tile offset"); acc += 32 * tlen[k]; if (k && G) CHECK(tlen[k] <= tlen[k - 1], "tiles not sorted"); if (!G) CHECK(tlen[k] <= std::max<uint32_t>(DN_ROWS, 300 / DN_MAX_GROUP + 1), "balanced tile with %u rows", tlen[k]); } CHECK(acc == ent, "entries"); }
    // every transcript is either idle or in exactly one (slot, component); slots of a component ascend with the transcript id
    std::vector<int> where(nt, 0);
    for (uint32_t i = 0; i < nidle; ++i) { CHECK(idle[i] >= s.t0 && idle[i] < s.t0 + nt, "idle range"); where[idle[i] - s.t0] += 1; }
    std::map<uint32_t, std::pair<uint32_t, uint32_t>> pos_of;   // global t -> (component, slot)
    for (uint32_t q = 0; q < ncomp_pad; ++q) {
        uint32_t prev = 0; bool ended = false;
        for (uint32_t j = 0; j < DN_MAX_SLOTS; ++j) {
            const uint32_t t = tmap[j * ncomp_pad + q];
            if (t == DN_NONE) { ended = true; continue; }
            CHECK(!ended && q < ncomp && j < NS, "hole in the slots of component %u", q);
            CHECK(j == 0 || t > prev, "slots not in transcript order");
            prev = t; where[t - s.t0] += 1; pos_of[t] = {q, j};
        }
    }
    for (uint32_t t = 0; t < nt; ++t) CHECK(where[t] == 1, "transcript %u appears %d times", t, where[t]);
    // every class appears once, in the column of its component, with the mask of its members
    std::vector<int> seen_c(nc, 0);
    for (uint32_t k = 0; k < tiles; ++k)
        for (uint32_t e = 0; e < tlen[k]; ++e)
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t pos = toff[k] + 32 * e + lane, c = cperm[pos], q = comp_of(k, lane);
                if (c == DN_NONE) { CHECK(mask[pos] == 0, "padding entry with a mask"); continue; }
                CHECK(q != DN_NONE, "class entry on an idle lane");
                CHECK(c >= s.c_lo && c < s.c_lo + nc, "cperm range"); seen_c[c - s.c_lo]++;
                uint32_t want = 0;
                for (uint32_t j = 0; j < s.len[c]; ++j) { auto it = pos_of.find(s.lab[s.start[c] + j]); CHECK(it != pos_of.end() && it->second.first == q, "class %u is in the wrong column", c); want |= 1u << it->second.second; }
                CHECK(mask[pos] == want, "mask of class %u", c);
            }
    for (uint32_t c = 0; c < nc; ++c) CHECK(seen_c[c] == 1, "class %u appears %d times", c, seen_c[c]);

    // ---- iterate
    const uint32_t T = s.t0 + nt + 9;
    std::vector<double> eff(T), single(T, 0.0), cnt(s.c_lo + nc, 0.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (uint32_t t = 0; t < T; ++t) { eff[t] = 1.0 + 3000.0 * U(rng); if (U(rng) < 0.3) single[t] = std::floor(100.0 * U(rng)); }
    for (uint32_t c = 0; c < nc; ++c) cnt[s.c_lo + c] = std::floor(1.0 + 50.0 * U(rng) * U(rng));
    const double prior = vb ? 0.01 : 0.0;
    std::vector<double> a_ref(T, 0.0);
    for (uint32_t t = s.t0; t < s.t0 + nt; ++t) a_ref[t] = 3.25;
    std::vector<double> s_cnt(ent, 0.0), s_beta(NS * ncomp_pad, 0.0), s_alpha(NS * ncomp_pad, 0.0), s_base(NS * ncomp_pad, 0.0), s_inveff(NS * ncomp_pad, 0.0);
    for (uint32_t i = 0; i < ent; ++i) s_cnt[i] = cperm[i] != DN_NONE ? cnt[cperm[i]] : 0.0;
    for (uint32_t i = 0; i < NS * ncomp_pad; ++i) { const uint32_t t = tmap[i]; if (t != DN_NONE) { s_alpha[i] = a_ref[t]; s_base[i] = single[t] + prior; s_inveff[i] = 1.0 / eff[t]; } }
    std::vector<double> idle_alpha(nidle);
    for (uint32_t i = 0; i < nidle; ++i) idle_alpha[i] = a_ref[idle[i]];
    auto digamma = [](double x) {
        double acc = 0.0;
        while (x < 12.0) { acc -= 1.0 / x; x += 1.0; }
        const double inv = 1.0 / x, inv2 = inv * inv;
        return acc + std::log(x) - 0.5 * inv - inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0)));
    };
    for (int it = 0; it < 20; ++it) {
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
            if (vb) { double sum = 0.0; for (double a : s_alpha) sum += a; for (double a : idle_alpha) sum += a; ln = digamma(sum); }
            for (uint32_t i = 0; i < NS * ncomp_pad; ++i) { const double a = s_alpha[i]; s_beta[i] = (vb ? (a > 0 ? std::exp(digamma(a) - ln) : 0.0) : a) * s_inveff[i]; }
        }
        for (uint32_t k = 0; k < tiles; ++k) {
            double accs[32][DN_MAX_SLOTS], bs[32][DN_MAX_SLOTS];
            for (uint32_t lane = 0; lane < 32; ++lane) {                // every lane: the classes of its column
                const uint32_t qi = comp_of(k, lane);
                for (uint32_t j = 0; j < NS; ++j) { bs[lane][j] = qi == DN_NONE ? 0.0 : s_beta[j * ncomp_pad + qi]; accs[lane][j] = 0.0; }
                for (uint32_t e = 0; e < tlen[k]; ++e) {
                    const double c = s_cnt[toff[k] + 32 * e + lane]; const uint32_t msk = mask[toff[k] + 32 * e + lane];
                    double S = 0.0;
                    for (uint32_t j = 0; j < NS; ++j) S += ((msk >> j) & 1u) ? bs[lane][j] : 0.0;
                    const double r = S > 0.0 ? c / S : 0.0;
                    for (uint32_t j = 0; j < NS; ++j) accs[lane][j] += ((msk >> j) & 1u) ? r : 0.0;
                }
            }
            for (uint32_t j = 0; j < NS; ++j)                           // butterfly over the lanes of a group (masked by the lane's group size)
                for (uint32_t o = 1; o < (G ? G : DN_MAX_GROUP); o <<= 1) {
                    double tmp[32];
                    for (uint32_t lane = 0; lane < 32; ++lane) tmp[lane] = accs[lane][j] + (o < lanes_of(k, lane) ? accs[lane ^ o][j] : 0.0);
                    for (uint32_t lane = 0; lane < 32; ++lane) accs[lane][j] = tmp[lane];
                }
            for (uint32_t lane = 0; lane < 32; ++lane) {                // the group's first lane owns the state
                const uint32_t qi = comp_of(k, lane);
                if (qi == DN_NONE || rank_of(k, lane) != 0) continue;
                for (uint32_t j = 0; j < NS; ++j) s_alpha[j * ncomp_pad + qi] = bs[lane][j] * accs[lane][j] + s_base[j * ncomp_pad + qi];
            }
        }
        for (uint32_t i = 0; i < nidle; ++i) idle_alpha[i] = single[idle[i]] + prior;
        for (uint32_t i = 0; i < NS * ncomp_pad; ++i) {
            const uint32_t t = tmap[i];
            if (t == DN_NONE) { CHECK(s_alpha[i] == 0.0, "empty slot became %g", s_alpha[i]); continue; }
            CHECK(std::fabs(a_ref[t] - s_alpha[i]) <= 1e-9 * std::max(1.0, std::fabs(a_ref[t])), "iteration %d transcript %u: %.17g vs %.17g", it, t, s_alpha[i], a_ref[t]);
        }
        for (uint32_t i = 0; i < nidle; ++i) CHECK(idle_alpha[i] == a_ref[idle[i]], "idle transcript %u", idle[i]);
    }
    return 0;
}

int main() {
    struct Case { uint64_t seed; uint32_t n_genes, gsize, cpg; bool chain, dup; uint32_t idle_every; bool dense, vb; };
    const Case cases[] = {
        {1, 135, 5, 20, false, false, 4, true, false},   // a cfg2-sized CTA range: 5-isoform genes, a quarter of them silent
        {2, 135, 5, 20, false, false, 4, true, true},
        {3, 1, 2, 1, false, false, 0, true, false},      // one component, one class
        {4, 40, 8, 60, false, false, 0, true, false},    // the largest component one thread takes
        {5, 40, 9, 60, true, false, 0, false, false},    // nine transcripts in one component: not dense
        {6, 33, 8, 1, true, false, 0, true, true},       // chains: the label has to travel seven links
        {7, 50, 4, 10, false, true, 0, false, false},    // a transcript twice in a label: a mask cannot hold that
        {8, 700, 3, 300, false, false, 3, true, false},  // many components, class counts beyond one bucket width
        {9, 5, 5, 3, false, false, 1, true, false},      // no class at all (every gene silent): zero components
    };
    for (uint32_t G : {1u, 2u, 4u, 0u})
    for (const Case& c : cases)
        if (run_case(c.seed, c.n_genes, c.gsize, c.cpg, c.chain, c.dup, c.idle_every, c.dense, c.vb, G)) { fprintf(stderr, "case seed %llu failed\n", (unsigned long long)c.seed); return 1; }
    printf("em_dense layout ok (%zu cases)\n", sizeof(cases) / sizeof(cases[0]));
    return 0;
}
