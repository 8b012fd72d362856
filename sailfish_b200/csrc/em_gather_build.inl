// em_gather_build.inl -- per-CTA construction of the "gather" layout of the EM loop (em_gather.cuh).
//
// The body is written against five macros so that the same text is the CUDA device code (em_gather.cuh) and a
// single-thread host function (tests/em_gather_layout_test.cpp: SFB_GB_NT == 1, atomics are plain updates), which lets
// the index arithmetic be checked where there is no GPU:
//   SFB_GB_FN            function qualifiers
//   SFB_GB_TID / _NT     this thread's index in the CTA / threads per CTA
//   SFB_GB_SYNC()        CTA barrier
//   SFB_GB_ADD(p, v)     atomic add on a uint32_t, returns the old value
//   SFB_GB_MAX(p, v)     atomic max on a uint32_t
//
// What it builds, for one CTA's transcript range [t0, t0+nt) and its local classes [c_lo, c_lo+nc) of the partition arrays:
//   * transcripts renumbered by DESCENDING degree (number of local classes they belong to), classes renumbered by DESCENDING
//     member count; both in warp tiles of 32 consecutive new indices;
//   * lab_e: for class tile k with L_k = its largest member count, entry (j, lane) at tile_e_off[k] + 32 j + lane holds the
//     new transcript index of member j of class 32 k + lane, or the sentinel nt_pad (whose beta is 0) -- the E-step is one
//     thread per class summing beta over a conflict-free column;
//   * cls_t: the transpose, same shape, for transcript tiles: the classes a transcript belongs to, sentinel nc_pad (r = 0) --
//     the M-step is one thread per transcript summing r; no atomics, no weights (w_i is recomputed as beta_i / sum beta);
//   * cperm[new class] = position in the partition arrays (for the counts), tmap[new transcript] = global transcript id.
// Sorting by size keeps the padding below 64 * 256 entries per list (sizes are bucketed up to 255; larger ones share the last
// bucket, and if the padded lists outgrow the region the header's GH_OK stays 0 and the caller keeps the atomic kernel).

#ifndef SFB_GATHER_GEOM_DEFINED
#define SFB_GATHER_GEOM_DEFINED
// identical for every CTA region; offsets in 32-bit words from the region start, all multiples of 4 (16 bytes)
struct GatherGeom {
    uint32_t region_words;
    uint32_t o_tile_e_off, o_tile_e_len, o_tile_t_off, o_tile_t_len, o_cperm, o_tmap, o_lab_e, o_cls_t;
    uint32_t cap_tiles_e, cap_tiles_t;   // tiles per list a region has room for
    uint32_t cap_ent;                    // 16-bit entries per list a region has room for
    uint32_t shift;                      // stored indices are (index << shift): 3 = byte offsets into an f64 array (needs
                                         // every padded index <= 8191), 0 = plain indices
    uint32_t pad_[3];
};
enum { GH_NC = 0, GH_NT, GH_TILES_E, GH_TILES_T, GH_ENT_E, GH_ENT_T, GH_OK, GH_MAXLEN, GH_MAXDEG, GH_WORDS = 16 };
constexpr uint32_t GB_BUCKETS = 256;
// geometry for the largest class count / entry count / transcript count of any CTA (host side)
inline GatherGeom gather_make_geom(uint64_t max_nc, uint64_t max_ne, uint64_t max_nt, bool allow_scaled = true) {
    auto up = [](uint64_t x, uint64_t m) { return (uint32_t)((x + m - 1) / m * m); };
    GatherGeom g;
    g.cap_tiles_e = up(max_nc / 32 + 2, 4);
    g.cap_tiles_t = up(max_nt / 32 + 2, 4);
    g.cap_ent = up(max_ne + 64 * GB_BUCKETS, 8);
    uint32_t o = GH_WORDS;
    g.o_tile_e_off = o; o += g.cap_tiles_e;
    g.o_tile_e_len = o; o += g.cap_tiles_e;
    g.o_tile_t_off = o; o += g.cap_tiles_t;
    g.o_tile_t_len = o; o += g.cap_tiles_t;
    g.o_cperm = o; o += up(max_nc + 1, 4);
    g.o_tmap = o; o += up(max_nt + 1, 4);
    g.o_lab_e = o; o += g.cap_ent / 2;
    g.o_cls_t = o; o += g.cap_ent / 2;
    g.region_words = o;
    g.shift = (allow_scaled && up(max_nc, 32) <= 8191 && up(max_nt, 32) <= 8191) ? 3u : 0u;
    g.pad_[0] = g.pad_[1] = g.pad_[2] = 0;
    return g;
}
// scratch words gather_build_cta needs
inline size_t gather_scratch_words(uint64_t max_nc, uint64_t max_nt, const GatherGeom& g) {
    return 3 * (size_t)max_nt + (size_t)max_nc + 2 * GB_BUCKETS + 2 * ((size_t)g.cap_tiles_e + g.cap_tiles_t) + 8;
}
#endif

// scratch: 3 nt + nc + 2 * GB_BUCKETS + 2 (tiles_e + tiles_t) + 4 words
SFB_GB_FN void gather_build_cta(const uint32_t* start, const uint32_t* len, const uint32_t* lab, uint32_t c_lo, uint32_t nc,
                                uint32_t t0, uint32_t nt, const GatherGeom g, uint32_t* region, uint32_t* scratch) {
    const uint32_t tid = SFB_GB_TID, nth = SFB_GB_NT;
    const uint32_t tiles_e = (nc + 31u) >> 5, tiles_t = (nt + 31u) >> 5;
    const uint32_t nc_pad = tiles_e << 5, nt_pad = tiles_t << 5;
    uint32_t* s_deg = scratch;                      // nt: degree of a transcript (old local index)
    uint32_t* s_tnew = s_deg + nt;                  // nt: old local index -> new index
    uint32_t* s_tcur = s_tnew + nt;                 // nt: fill cursor of a transcript's class list (old local index)
    uint32_t* s_cnew = s_tcur + nt;                 // nc: class (partition-relative) -> new index
    uint32_t* s_hist_e = s_cnew + nc;               // GB_BUCKETS
    uint32_t* s_hist_t = s_hist_e + GB_BUCKETS;     // GB_BUCKETS
    uint32_t* s_len_e = s_hist_t + GB_BUCKETS;      // tiles_e: L_k
    uint32_t* s_off_e = s_len_e + tiles_e;          // tiles_e
    uint32_t* s_len_t = s_off_e + tiles_e;          // tiles_t
    uint32_t* s_off_t = s_len_t + tiles_t;          // tiles_t
    uint32_t* s_misc = s_off_t + tiles_t;           // [0] ok  [1] max class size  [2] max degree
    uint32_t* hdr = region;
    uint32_t* cperm = region + g.o_cperm;
    uint32_t* tmap = region + g.o_tmap;
    uint16_t* lab_e = reinterpret_cast<uint16_t*>(region + g.o_lab_e);
    uint16_t* cls_t = reinterpret_cast<uint16_t*>(region + g.o_cls_t);

    // ---- P0: clear
    for (uint32_t i = tid; i < nt; i += nth) { s_deg[i] = 0; s_tcur[i] = 0; }
    for (uint32_t i = tid; i < 2 * GB_BUCKETS; i += nth) s_hist_e[i] = 0;
    for (uint32_t i = tid; i < tiles_e; i += nth) s_len_e[i] = 0;
    for (uint32_t i = tid; i < tiles_t; i += nth) s_len_t[i] = 0;
    if (tid == 0) { s_misc[0] = 1; s_misc[1] = 0; s_misc[2] = 0; }
    SFB_GB_SYNC();
    // ---- P1: class-size histogram and transcript degrees
    for (uint32_t c = tid; c < nc; c += nth) {
        const uint32_t n = len[c_lo + c], b = start[c_lo + c];
        SFB_GB_ADD(s_hist_e + (n < GB_BUCKETS ? n : GB_BUCKETS - 1), 1u);
        SFB_GB_MAX(s_misc + 1, n);
        for (uint32_t j = 0; j < n; ++j) SFB_GB_ADD(s_deg + (lab[b + j] - t0), 1u);
    }
    SFB_GB_SYNC();
    // ---- P2: degree histogram
    for (uint32_t t = tid; t < nt; t += nth) {
        const uint32_t d = s_deg[t];
        SFB_GB_ADD(s_hist_t + (d < GB_BUCKETS ? d : GB_BUCKETS - 1), 1u);
        SFB_GB_MAX(s_misc + 2, d);
    }
    SFB_GB_SYNC();
    // ---- P3: descending exclusive prefix -> bucket cursors
    if (tid == 0) {
        uint32_t acc = 0;
        for (int b = (int)GB_BUCKETS - 1; b >= 0; --b) { const uint32_t h = s_hist_e[b]; s_hist_e[b] = acc; acc += h; }
        acc = 0;
        for (int b = (int)GB_BUCKETS - 1; b >= 0; --b) { const uint32_t h = s_hist_t[b]; s_hist_t[b] = acc; acc += h; }
    }
    SFB_GB_SYNC();
    // ---- P4: new indices, the inverse maps, per-tile maxima
    for (uint32_t c = tid; c < nc; c += nth) {
        const uint32_t n = len[c_lo + c];
        const uint32_t pos = SFB_GB_ADD(s_hist_e + (n < GB_BUCKETS ? n : GB_BUCKETS - 1), 1u);
        s_cnew[c] = pos;
        cperm[pos] = c_lo + c;
        SFB_GB_MAX(s_len_e + (pos >> 5), n);
    }
    for (uint32_t t = tid; t < nt; t += nth) {
        const uint32_t d = s_deg[t];
        const uint32_t pos = SFB_GB_ADD(s_hist_t + (d < GB_BUCKETS ? d : GB_BUCKETS - 1), 1u);
        s_tnew[t] = pos;
        tmap[pos] = t0 + t;
        SFB_GB_MAX(s_len_t + (pos >> 5), d);
    }
    SFB_GB_SYNC();
    // ---- P5: tile offsets, capacity check, header
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t k = 0; k < tiles_e; ++k) { s_off_e[k] = acc; acc += s_len_e[k] << 5; }
        const uint32_t ent_e = acc;
        acc = 0;
        for (uint32_t k = 0; k < tiles_t; ++k) { s_off_t[k] = acc; acc += s_len_t[k] << 5; }
        const uint32_t ent_t = acc;
        const bool ok = ent_e <= g.cap_ent && ent_t <= g.cap_ent && tiles_e <= g.cap_tiles_e && tiles_t <= g.cap_tiles_t &&
                        (nc_pad << g.shift) <= 65535u && (nt_pad << g.shift) <= 65535u;
        s_misc[0] = ok ? 1u : 0u;
        hdr[GH_NC] = nc; hdr[GH_NT] = nt; hdr[GH_TILES_E] = tiles_e; hdr[GH_TILES_T] = tiles_t;
        hdr[GH_ENT_E] = ent_e; hdr[GH_ENT_T] = ent_t; hdr[GH_OK] = ok ? 1u : 0u;
        hdr[GH_MAXLEN] = s_misc[1]; hdr[GH_MAXDEG] = s_misc[2];
    }
    SFB_GB_SYNC();
    if (!s_misc[0]) return;
    for (uint32_t k = tid; k < tiles_e; k += nth) { region[g.o_tile_e_off + k] = s_off_e[k]; region[g.o_tile_e_len + k] = s_len_e[k]; }
    for (uint32_t k = tid; k < tiles_t; k += nth) { region[g.o_tile_t_off + k] = s_off_t[k]; region[g.o_tile_t_len + k] = s_len_t[k]; }
    // ---- P6: member columns of the classes (+ padding), and the transposed lists through per-transcript cursors
    for (uint32_t c = tid; c < nc_pad; c += nth) {
        if (c < nc) {
            const uint32_t n = len[c_lo + c], b = start[c_lo + c];
            const uint32_t pos = s_cnew[c], k = pos >> 5;
            const uint32_t base = s_off_e[k] + (pos & 31u), L = s_len_e[k];
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t told = lab[b + j] - t0, tn = s_tnew[told];
                lab_e[base + (j << 5)] = (uint16_t)(tn << g.shift);
                const uint32_t slot = SFB_GB_ADD(s_tcur + told, 1u);
                cls_t[s_off_t[tn >> 5] + (slot << 5) + (tn & 31u)] = (uint16_t)(pos << g.shift);
            }
            for (uint32_t j = n; j < L; ++j) lab_e[base + (j << 5)] = (uint16_t)(nt_pad << g.shift);
        } else {                                      // lanes of the last tile that hold no class
            const uint32_t k = c >> 5, base = s_off_e[k] + (c & 31u), L = s_len_e[k];
            for (uint32_t j = 0; j < L; ++j) lab_e[base + (j << 5)] = (uint16_t)(nt_pad << g.shift);
        }
    }
    SFB_GB_SYNC();
    // ---- P7: padding of the transposed lists
    for (uint32_t t = tid; t < nt_pad; t += nth) {
        uint32_t tn, d;
        if (t < nt) { tn = s_tnew[t]; d = s_deg[t]; } else { tn = t; d = 0; }   // new indices nt..nt_pad-1 hold no transcript
        const uint32_t k = tn >> 5, base = s_off_t[k] + (tn & 31u), L = s_len_t[k];
        for (uint32_t j = d; j < L; ++j) cls_t[base + (j << 5)] = (uint16_t)(nc_pad << g.shift);
    }
}
