// em_dense_build.inl -- per-CTA construction of the "dense component" layout of the EM loop (em_dense.cuh).
//
// Written against the same five macros as em_gather_build.inl (SFB_GB_FN, _TID, _NT, _SYNC, _ADD, _MAX) plus SFB_GB_MIN
// (atomic min on a uint32_t), so that the text is both the CUDA device code and a single-thread host function for the CPU
// test (tests/em_dense_layout_test.cpp).
//
// The (class x transcript) structure of one CTA's range falls apart into connected components -- for an annotated transcriptome,
// the isoforms of a gene and the classes over them.  When every component of every CTA has at most DN_MAX_SLOTS transcripts, a
// component is small enough for ONE thread: its transcripts are "slots" 0..NS-1, a class is a bit mask over the slots, and an
// EM iteration of the component needs nothing outside the thread (em_dense.cuh).  This builder
//   * finds the components (min-label propagation over the classes, a few rounds for gene-sized components),
//   * numbers the components by DESCENDING class count (32 consecutive components = one warp tile; the tile's class loop runs
//     to the largest count in the tile, the rest is padding with count 0),
//   * writes per component the global ids of its transcripts ([slot][component], 0xFFFFFFFF = empty slot), per class entry
//     (column-major inside the tile) the position of its count in the partition arrays and its slot mask, and the list of
//     "idle" transcripts (members of no multi-member class: alpha = their single-class count after the first iteration).
// DH_KIND stays 0 -- and the caller keeps the gather / scatter kernels -- if a component is larger than DN_MAX_SLOTS, a label
// holds a transcript twice (a mask has no multiplicity), the propagation does not settle in DN_MAX_ROUNDS rounds, or the
// region is too small.

#ifndef SFB_DENSE_GEOM_DEFINED
#define SFB_DENSE_GEOM_DEFINED
constexpr uint32_t DN_MAX_SLOTS = 8, DN_MAX_ROUNDS = 48, DN_BUCKETS = 256, DN_NONE = 0xFFFFFFFFu, DN_ROWS = 4, DN_MAX_GROUP = 8;
// identical for every CTA region; offsets in 32-bit words from the region start, all multiples of 4 (16 bytes)
struct DenseGeom {
    uint32_t region_words;
    uint32_t o_tile_off, o_tile_len, o_cperm, o_mask, o_tmap, o_idle;
    uint32_t cap_tiles;      // component tiles a region has room for
    uint32_t cap_ent;        // class entries (padded) a region has room for
    uint32_t cap_nt;         // transcripts
    uint32_t group;          // lanes per component (1, 2 or 4): a component's classes are dealt round-robin to its lanes;
                             // 0 = as many lanes (1, 2, 4 or 8) as it takes to leave a lane at most DN_ROWS classes ("balanced")
    uint32_t o_lane;         // group 0: per lane of every tile  component | log2(lanes of the component) << 16 | lane's rank << 20
    uint32_t pad_[4];
};
enum { DH_KIND = 0, DH_NCOMP, DH_TILES, DH_ENT, DH_NS, DH_NIDLE, DH_NT, DH_NC, DH_ROUNDS, DH_GROUP, DH_WORDS = 16 };
inline DenseGeom dense_make_geom(uint64_t max_nc, uint64_t max_nt, uint32_t group = 1) {
    auto up = [](uint64_t x, uint64_t m) { return (uint32_t)((x + m - 1) / m * m); };
    DenseGeom g;
    g.group = (group == 0 || group == 2 || group == 4) ? group : 1;
    // a component has at least two transcripts; balanced: at most 2 (cc / DN_ROWS + 1) lanes per component
    g.cap_tiles = g.group ? up(max_nt / 2 * g.group / 32 + 2, 4) : up((max_nc / 2 + max_nt) / 32 + 2, 4);
    g.cap_ent = up(max_nc + 64 * DN_BUCKETS, 16);
    g.cap_nt = up(max_nt + 1, 4);
    uint32_t o = DH_WORDS;
    g.o_tile_off = o; o += g.cap_tiles;
    g.o_tile_len = o; o += g.cap_tiles;
    g.o_cperm = o; o += g.cap_ent;
    g.o_mask = o; o += g.cap_ent / 4;
    g.o_tmap = o; o += DN_MAX_SLOTS * 32 * g.cap_tiles;
    g.o_idle = o; o += g.cap_nt;
    g.o_lane = o; o += 32 * g.cap_tiles;
    g.region_words = o;
    for (int i = 0; i < 4; ++i) g.pad_[i] = 0;
    return g;
}
// scratch words dense_build_cta needs
inline size_t dense_scratch_words(uint64_t max_nt, const DenseGeom& g) { return 9 * (size_t)max_nt + DN_BUCKETS + 2 * (size_t)g.cap_tiles + 16; }
#endif

// dirty_in (may be null): transcripts whose classes were all sent to the pool -- they belong to the pool loop, not to the idle list.
// dirty_out (may be null): instead of just giving up on a component that is too large for one thread (or on a propagation that does
// not settle), mark its transcripts there and report DH_KIND = 2: the caller sends their classes to the pool and builds again.
SFB_GB_FN void dense_build_cta(const uint32_t* start, const uint32_t* len, const uint32_t* lab, uint32_t c_lo, uint32_t nc,
                               uint32_t t0, uint32_t nt, const DenseGeom g, uint32_t* region, uint32_t* scratch,
                               const uint8_t* dirty_in = nullptr, uint8_t* dirty_out = nullptr) {
    const uint32_t tid = SFB_GB_TID, nth = SFB_GB_NT;
    uint32_t* s_comp = scratch;              // nt: component label (smallest local index of the component)
    uint32_t* s_deg = s_comp + nt;           // nt: number of local classes the transcript belongs to
    uint32_t* s_size = s_deg + nt;           // nt: per root, transcripts in the component
    uint32_t* s_ccnt = s_size + nt;          // nt: per root, classes in the component
    uint32_t* s_q = s_ccnt + nt;             // nt: per root, new component index
    uint32_t* s_cur = s_q + nt;              // nt: per root, fill cursor of the class column
    uint32_t* s_slot = s_cur + nt;           // nt: slot of a transcript inside its component
    uint32_t* s_hist = s_slot + nt;          // DN_BUCKETS
    uint32_t* s_tlen = s_hist + DN_BUCKETS;  // cap_tiles
    uint32_t* s_toff = s_tlen + g.cap_tiles; // cap_tiles
    uint32_t* s_misc = s_toff + g.cap_tiles; // [0] ok [1] changed [2] max size [3] components [4] idle [5] rounds [6] tiles [7] entries
    uint32_t* s_gq = s_misc + 16;            // nt: balanced layout, lanes of component q
    uint32_t* s_lb = s_gq + nt;              // nt: balanced layout, first lane of component q
    uint32_t* hdr = region;
    uint32_t* cperm = region + g.o_cperm;
    uint8_t* mask = reinterpret_cast<uint8_t*>(region + g.o_mask);
    uint32_t* tmap = region + g.o_tmap;
    uint32_t* idle = region + g.o_idle;

    for (uint32_t i = tid; i < nt; i += nth) { s_comp[i] = i; s_deg[i] = 0; s_size[i] = 0; s_ccnt[i] = 0; s_cur[i] = 0; s_slot[i] = 0; s_q[i] = DN_NONE; }
    for (uint32_t i = tid; i < DN_BUCKETS; i += nth) s_hist[i] = 0;
    for (uint32_t i = tid; i < g.cap_tiles; i += nth) s_tlen[i] = 0;
    if (tid == 0) { s_misc[0] = 1; s_misc[1] = 0; s_misc[2] = 0; s_misc[3] = 0; s_misc[4] = 0; s_misc[5] = 0; s_misc[6] = 0; s_misc[7] = 0; hdr[DH_KIND] = 0; }
    SFB_GB_SYNC();
    for (uint32_t c = tid; c < nc; c += nth) {
        const uint32_t n = len[c_lo + c], b = start[c_lo + c];
        for (uint32_t j = 0; j < n; ++j) SFB_GB_ADD(s_deg + (lab[b + j] - t0), 1u);
    }
    SFB_GB_SYNC();
    // ---- connected components: every class pulls its members down to the smallest label among them, until nothing moves
    uint32_t rounds = 0;
    for (;; ++rounds) {
        if (tid == 0) s_misc[1] = 0;
        SFB_GB_SYNC();
        for (uint32_t c = tid; c < nc; c += nth) {
            const uint32_t n = len[c_lo + c], b = start[c_lo + c];
            uint32_t m = DN_NONE;
            for (uint32_t j = 0; j < n; ++j) { const uint32_t v = s_comp[lab[b + j] - t0]; m = v < m ? v : m; }
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t t = lab[b + j] - t0;
                if (s_comp[t] > m) { SFB_GB_MIN(s_comp + t, m); s_misc[1] = 1; }
            }
        }
        SFB_GB_SYNC();
        const uint32_t changed = s_misc[1];
        SFB_GB_SYNC();
        if (!changed) break;
        if (rounds + 1 >= DN_MAX_ROUNDS) { if (tid == 0) s_misc[0] = 0; break; }
    }
    SFB_GB_SYNC();
    if (!s_misc[0]) {
        if (dirty_out) {                                               // not settled: everything with a local class goes to the pool
            for (uint32_t t = tid; t < nt; t += nth) if (s_deg[t]) dirty_out[t0 + t] = 1;
            if (tid == 0) hdr[DH_KIND] = 2u;
        }
        return;
    }
    // labels are not yet roots everywhere if a label moved after a class last looked at it: they are, because the loop only
    // ends after a full round in which no class saw two different labels among its members, and a label is always the index of
    // a transcript that carries it (the minimum never leaves its own transcript)
    for (uint32_t t = tid; t < nt; t += nth) if (s_deg[t]) SFB_GB_ADD(s_size + s_comp[t], 1u);
    for (uint32_t c = tid; c < nc; c += nth) SFB_GB_ADD(s_ccnt + s_comp[lab[start[c_lo + c]] - t0], 1u);
    SFB_GB_SYNC();
    for (uint32_t t = tid; t < nt; t += nth) {
        if (!s_deg[t]) { if (!(dirty_in && dirty_in[t0 + t])) { const uint32_t p = SFB_GB_ADD(s_misc + 4, 1u); idle[p] = t0 + t; } }
        else if (s_comp[t] == t) {
            SFB_GB_MAX(s_misc + 2, s_size[t]);
            const uint32_t cc = s_ccnt[t];
            SFB_GB_ADD(s_hist + (cc < DN_BUCKETS ? cc : DN_BUCKETS - 1), 1u);
        }
    }
    SFB_GB_SYNC();
    if (s_misc[2] > DN_MAX_SLOTS) {                                    // a component too large for one thread
        if (dirty_out) {
            for (uint32_t t = tid; t < nt; t += nth) if (s_deg[t] && s_size[s_comp[t]] > DN_MAX_SLOTS) dirty_out[t0 + t] = 1;
            if (tid == 0) hdr[DH_KIND] = 2u;
        }
        return;
    }
    if (tid == 0) {
        uint32_t acc = 0;
        for (int b = (int)DN_BUCKETS - 1; b >= 0; --b) { const uint32_t h = s_hist[b]; s_hist[b] = acc; acc += h; }
        s_misc[3] = acc;                                                // components
    }
    SFB_GB_SYNC();
    const uint32_t ncomp = s_misc[3];
    const uint32_t G = g.group;                                         // 0: balanced (lanes per component from its class count)
    for (uint32_t t = tid; t < nt; t += nth) {
        if (s_deg[t] && s_comp[t] == t) {
            const uint32_t cc = s_ccnt[t];
            const uint32_t q = SFB_GB_ADD(s_hist + (cc < DN_BUCKETS ? cc : DN_BUCKETS - 1), 1u);
            s_q[t] = q;
            uint32_t lanes = G;                                          // balanced: 1, 2, 4 or 8 lanes so that a lane holds <= DN_ROWS classes
            if (G == 0) { const uint32_t need = (cc + DN_ROWS - 1) / DN_ROWS; lanes = 1; while (lanes < need && lanes < DN_MAX_GROUP) lanes <<= 1; }
            s_gq[q] = lanes;
        }
    }
    SFB_GB_SYNC();
    // first lane of every component: components are numbered by descending class count, so their lane counts (powers of two)
    // never increase and every group starts on a multiple of its size
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t q = 0; q < ncomp; ++q) { s_lb[q] = acc; acc += s_gq[q]; }
        s_misc[6] = (acc + 31u) >> 5;
        if (ncomp > 0xFFFFu) s_misc[0] = 0;                             // the lane table keeps the component in 16 bits
        if (G == 0) for (uint32_t q = 1; q < ncomp; ++q) if (s_gq[q] > s_gq[q - 1]) s_misc[0] = 0;   // counts >= 255 share a bucket unsorted
    }
    SFB_GB_SYNC();
    const uint32_t tiles = s_misc[6];
    const uint32_t ncomp_pad = G ? (tiles << 5) / G : ((ncomp + 31u) & ~31u);
    if (tiles > g.cap_tiles || !s_misc[0]) return;
    for (uint32_t t = tid; t < nt; t += nth) {
        if (s_deg[t] && s_comp[t] == t) {
            const uint32_t q = s_q[t], gq = s_gq[q];
            SFB_GB_MAX(s_tlen + (s_lb[q] >> 5), (s_ccnt[t] + gq - 1) / gq);   // rows of the tile: a lane takes every gq-th class
        }
    }
    SFB_GB_SYNC();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t k = 0; k < tiles; ++k) { s_toff[k] = acc; acc += s_tlen[k] << 5; }
        s_misc[7] = acc;
        if (acc > g.cap_ent) s_misc[0] = 0;
    }
    SFB_GB_SYNC();
    if (!s_misc[0]) return;
    const uint32_t ent = s_misc[7];
    for (uint32_t k = tid; k < tiles; k += nth) { region[g.o_tile_off + k] = s_toff[k]; region[g.o_tile_len + k] = s_tlen[k]; }
    for (uint32_t i = tid; i < DN_MAX_SLOTS * ncomp_pad; i += nth) tmap[i] = DN_NONE;
    for (uint32_t i = tid; i < ent; i += nth) { cperm[i] = DN_NONE; mask[i] = 0; }
    if (G == 0) {
        uint32_t* lane_tab = region + g.o_lane;
        for (uint32_t i = tid; i < (tiles << 5); i += nth) lane_tab[i] = DN_NONE;
        SFB_GB_SYNC();
        for (uint32_t q = tid; q < ncomp; q += nth) {
            const uint32_t gq = s_gq[q], lg = gq == 1 ? 0u : gq == 2 ? 1u : gq == 4 ? 2u : 3u;
            for (uint32_t r = 0; r < gq; ++r) lane_tab[s_lb[q] + r] = q | (lg << 16) | (r << 20);
        }
    }
    SFB_GB_SYNC();
    // ---- slots: rank of a transcript among the members of its component, in transcript order
    for (uint32_t t = tid; t < nt; t += nth) {
        if (!s_deg[t]) continue;
        const uint32_t r = s_comp[t];
        uint32_t s = 0;
        for (uint32_t u = r; u < t; ++u) if (s_deg[u] && s_comp[u] == r) ++s;
        s_slot[t] = s;
        tmap[s * ncomp_pad + s_q[r]] = t0 + t;
    }
    SFB_GB_SYNC();
    // ---- class entries: count position and slot mask, one column per component
    for (uint32_t c = tid; c < nc; c += nth) {
        const uint32_t n = len[c_lo + c], b = start[c_lo + c];
        const uint32_t r = s_comp[lab[b] - t0], q = s_q[r];
        const uint32_t e = SFB_GB_ADD(s_cur + r, 1u);
        uint32_t mk = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t bit = 1u << s_slot[lab[b + j] - t0];
            if (mk & bit) s_misc[0] = 0;                                // the same transcript twice in one label
            mk |= bit;
        }
        const uint32_t gq = s_gq[q], lb = s_lb[q];
        const uint32_t pos = s_toff[lb >> 5] + ((e / gq) << 5) + (lb & 31u) + (e % gq);
        cperm[pos] = c_lo + c;
        mask[pos] = (uint8_t)mk;
    }
    SFB_GB_SYNC();
    if (tid == 0) {
        hdr[DH_NCOMP] = ncomp; hdr[DH_TILES] = tiles; hdr[DH_ENT] = ent; hdr[DH_NS] = s_misc[2]; hdr[DH_NIDLE] = s_misc[4];
        hdr[DH_NT] = nt; hdr[DH_NC] = nc; hdr[DH_ROUNDS] = rounds + 1; hdr[DH_GROUP] = G;
        hdr[DH_KIND] = s_misc[0] ? 1u : 0u;
    }
}
