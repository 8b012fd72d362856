// map.cu -- read batch -> quasi-mapping -> equivalence-class counts on sm_100a.
//
// Replaces processReadsQuasi<IndexT> (reference src/SailfishQuantify.cpp:105-452 paired, :458-646 single) together
// with the EquivalenceClassBuilder it feeds (include/EquivalenceClassBuilder.hpp:62-110, TranscriptGroup.cpp:9-19).
// The quasi-mapping itself (RapMap's SACollector / mergeLeftRightHits) is not in the reference tree; the algorithm
// is "mapping spec v1" (DESIGN.md section 3).  DESIGN.md section 5 describes the kernel and the class table.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.cuh"

namespace {

constexpr int MAX_IV = 16;                 // spec v1: at most 16 maximal-match intervals per orientation scan
constexpr uint32_t MAX_READ_LEN = 256;     // spec v1: reads are clipped to 256 bases
constexpr int RW = MAX_READ_LEN / 32 + 1;  // packed words per read (+1 pad so a 32-base window never runs off the end)
constexpr int MAP_THREADS = 256;
constexpr int POS_BIAS = 4096;

// library-format ids (LibraryFormat::formatID, include/LibraryFormat.hpp:89-98)
enum { O_SAME = 0, O_AWAY = 1, O_TOWARD = 2, O_NONE = 3 };
enum { S_SA = 0, S_AS = 1, S_S = 2, S_A = 3, S_U = 4 };
__device__ __forceinline__ int mkfmt(int type, int orient, int strand) { return (type & 1) | ((orient & 3) << 1) | ((strand & 7) << 3); }
__device__ __forceinline__ int f_orient(int id) { return (id >> 1) & 3; }
__device__ __forceinline__ int f_strand(int id) { return (id >> 3) & 7; }

// compatibleHit for single-end reads and orphans (src/SailfishUtils.cpp:157-211)
__device__ __forceinline__ bool compat_single(int expected, bool fwd, int ms) {
    const int es = f_strand(expected);
    if (ms == 0) return fwd ? (es == S_U || es == S_S) : (es == S_U || es == S_A);
    if (f_orient(expected) == O_SAME) return es == S_U || (es == S_S && fwd) || (es == S_A && !fwd);
    if (ms == 1) return fwd ? (es == S_U || es == S_S) : (es == S_U || es == S_A);
    return fwd ? (es == S_U || es == S_A) : (es == S_U || es == S_S);
}
// hitType (:243-289) followed by the paired compatibleHit (:215-239)
__device__ __forceinline__ bool compat_paired(int expected, int32_t e1, bool fwd1, uint32_t len1, int32_t e2, bool fwd2,
                                              uint32_t len2, bool dovetail) {
    int obs;
    if (fwd1 != fwd2) {
        if (fwd1) { const int32_t st = dovetail ? (int32_t)len2 : 0; obs = (e1 <= e2 + st) ? mkfmt(1, O_TOWARD, S_SA) : mkfmt(1, O_AWAY, S_SA); }
        else { const int32_t st = dovetail ? (int32_t)len1 : 0; obs = (e2 <= e1 + st) ? mkfmt(1, O_TOWARD, S_AS) : mkfmt(1, O_AWAY, S_AS); }
    } else {
        obs = fwd1 ? mkfmt(1, O_SAME, S_S) : mkfmt(1, O_SAME, S_A);
    }
    if (f_orient(expected) != f_orient(obs)) return false;
    return f_strand(expected) == S_U || f_strand(expected) == f_strand(obs);
}

struct IndexView {
    const uint64_t* words; const uint64_t* txp_start; const uint64_t* txp_end;
    const uint4* sa; const uint4* table; const uint32_t* mfilter;
    uint64_t mask; int k; uint64_t kmask;
    int mf_m;                     // m of the m-mer presence bitmap (kmer_filter.hpp)
};
// a suffix entry is {transcript, position inside it, text position, bases left to the transcript's end}: the finalize kernel reads the
// first half, the scan kernel's match extension the second
__device__ __forceinline__ uint2 sa_tid_rel(const IndexView& ix, uint64_t e) { return __ldg(reinterpret_cast<const uint2*>(ix.sa + e)); }
__device__ __forceinline__ uint2 sa_pos_rem(const IndexView& ix, uint64_t e) { return __ldg(reinterpret_cast<const uint2*>(ix.sa + e) + 1); }

// the per-hit arithmetic of the bias / GC sample collection, shared with bias.cu and the CPU check (tests/bias_core_test.cpp)
#define SFB_BD __device__ __forceinline__
#define SFB_LDG(p) __ldg(p)
#define SFB_POPC64(x) __popcll(x)
#define SFB_D2I_RN(x) __double2int_rn(x)
#include "bias_core.inl"
#undef SFB_BD
#undef SFB_LDG
#undef SFB_POPC64
#undef SFB_D2I_RN
__device__ __forceinline__ int32_t hit_read_start_index(const IndexView& ix, uint32_t tid, int32_t pos, bool fwd, uint32_t readLen) {
    const uint64_t t0 = ix.txp_start[tid];
    return b_read_start_index(ix.words, t0, (int32_t)(ix.txp_end[tid] - t0), pos, fwd, readLen);
}

// ---- the equivalence-class table ------------------------------------------------------------------------------------------
// libcuckoo's layout (4 slots per bucket, two candidate buckets per key, include/cuckoohash_config.hh:9,
// cuckoohash_map.hh:1012-1026) with the BFS displacement replaced by a linear-probed overflow region: the table is kept
// under 50% load, so both buckets being full is rare.  A slot is one 64-bit word
//     [ arena offset : 34 | label length : 10 | hash fingerprint : 20 ]      (0 = empty)
// pointing at the label's transcript ids in an append-only arena; counts live in a parallel u64 array.  A bucket is one
// 32-byte sector.  Key equality is the full label (TranscriptGroup.cpp:53-55); XXH64 only chooses the buckets.
struct EqTable {
    unsigned long long* slot;      // n_buckets*4 + overflow
    unsigned long long* count;     // same length
    uint32_t* arena;               // label storage
    unsigned long long* cursor;    // [0] arena words used  [1] distinct labels  [2] error flags  [3] overflow inserts  [4] reads listed for a retry
    uint64_t n_buckets;            // power of two
    uint64_t n_overflow;           // power of two
    uint64_t arena_words;
};
constexpr unsigned long long ERR_ARENA_FULL = 1, ERR_TABLE_FULL = 2, ERR_LABEL_LONG = 4;

__device__ __forceinline__ unsigned long long pack_slot(uint64_t off, uint32_t len, uint64_t h) {
    return (off << 30) | ((unsigned long long)len << 20) | (h >> 44);
}

template <typename GetLabel>
__device__ __forceinline__ bool slot_matches(const EqTable& tb, unsigned long long sv, uint32_t len, uint64_t h, GetLabel get) {
    if (((sv ^ pack_slot(0, len, h)) & ((1ULL << 30) - 1)) != 0) return false;      // length + fingerprint
    const uint32_t* a = tb.arena + (sv >> 30);
    for (uint32_t j = 0; j < len; ++j) if (__ldcg(a + j) != get(j)) return false;
    return true;
}

// upsert: returns false only when the table or the arena is exhausted (error flag raised)
template <typename GetLabel>
__device__ bool eq_upsert(const EqTable& tb, uint32_t len, GetLabel get, unsigned long long add, uint64_t h /* XXH64 of the label */) {
    if (len >= 1024) { atomicOr(tb.cursor + 2, ERR_LABEL_LONG); return false; }
    const uint64_t bmask = tb.n_buckets - 1;
    const uint64_t b1 = h & bmask;
    const uint64_t b2 = (b1 ^ (((h >> 48) + 1) * 0x5bd1e995ULL)) & bmask;
    unsigned long long mine = 0;                                       // our published-candidate slot word (arena copy made)
    const uint64_t n_main = tb.n_buckets * 4;
    const uint64_t total_probe = 8 + tb.n_overflow;
    const uint64_t ov0 = h >> 20;
    for (uint64_t step = 0; step < total_probe; ++step) {
        uint64_t idx;
        if (step < 4) idx = b1 * 4 + step;
        else if (step < 8) idx = b2 * 4 + (step - 4);
        else idx = n_main + ((ov0 + (step - 8)) & (tb.n_overflow - 1));
        unsigned long long sv = __ldcg(tb.slot + idx);
        for (;;) {
            if (sv == 0ULL) {
                if (!mine) {
                    const unsigned long long off = atomicAdd(tb.cursor, (unsigned long long)len);
                    if (off + len > tb.arena_words) { atomicOr(tb.cursor + 2, ERR_ARENA_FULL); return false; }
                    for (uint32_t j = 0; j < len; ++j) tb.arena[off + j] = get(j);
                    mine = pack_slot(off, len, h);
                    __threadfence();                                   // label visible before the slot word
                }
                const unsigned long long prev = atomicCAS(tb.slot + idx, 0ULL, mine);
                if (prev == 0ULL) {
                    atomicAdd(tb.count + idx, add);
                    atomicAdd(tb.cursor + 1, 1ULL);
                    if (step >= 8) atomicAdd(tb.cursor + 3, 1ULL);
                    return true;
                }
                sv = prev;                                             // somebody else took the slot: look at what they put
                continue;
            }
            // no fence on this side: the slot word and the label are read from L2 (ld.cg), the label's address depends on the slot
            // word, and the writer published the label with a fence before its CAS (a fence here costs an L1 flush per probe)
            if (slot_matches(tb, sv, len, h, get)) { atomicAdd(tb.count + idx, add); return true; }
            break;
        }
    }
    atomicOr(tb.cursor + 2, ERR_TABLE_FULL);
    return false;
}

// ---- per-read state -----------------------------------------------------------------------------------------------------
// One mate.  The packed bases live in SHARED memory, lane-interleaved (word w of orientation o of this lane at
// sb[(o*RW + w) * 32]): every k-mer window and every extension compare reads them with a dynamic index, which in local
// memory would go out to L2 (a CTA's reads do not fit L1).  2 bits per base, base i at bits 2*(i%32) of word i/32;
// orientation 0 = as sequenced, 1 = reverse complement.  The invalid-base masks (same layout, 0b01 where the base is not
// A/C/G/T) are rare and stay in local memory, touched only when has_n.
extern __shared__ uint64_t smem_reads[];       // scan kernel: [warp][mate][orientation][word][lane]
struct Read {                                  // scalars only: selected by mate with ?: they stay in registers
    uint32_t sb;                               // index of this lane's column of this mate in smem_reads
    uint32_t len;
    bool has_n;
    __device__ __forceinline__ uint64_t word(int o, uint32_t w) const { return smem_reads[sb + (o * RW + w) * 32]; }
};
struct ReadN { uint64_t nm[2][RW]; };          // invalid-base masks of one mate: local memory, touched only when has_n

__device__ __forceinline__ uint64_t win32(const Read& r, int o, uint32_t pos) {
    const uint32_t idx = pos >> 5, sh = 2 * (pos & 31);
    uint64_t v = r.word(o, idx) >> sh;
    if (sh) v |= r.word(o, idx + 1) << (64 - sh);
    return v;
}
__device__ __forceinline__ uint64_t win32n(const ReadN& r, int o, uint32_t pos) {
    const uint32_t idx = pos >> 5, sh = 2 * (pos & 31);
    uint64_t v = r.nm[o][idx] >> sh;
    if (sh) v |= r.nm[o][idx + 1] << (64 - sh);
    return v;
}
__device__ __forceinline__ uint64_t win32g(const uint64_t* __restrict__ w, uint64_t pos) {
    const uint64_t idx = pos >> 5; const uint32_t sh = 2 * (pos & 31);
    uint64_t v = __ldg(w + idx) >> sh;
    if (sh) v |= __ldg(w + idx + 1) << (64 - sh);
    return v;
}
// reverse complement of one 32-base word: complement every 2-bit code, reverse the order of the codes
__device__ __forceinline__ uint64_t revcomp32(uint64_t w) {
    uint64_t x = __brevll(~w);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

// ---- read packing ------------------------------------------------------------------------------------------------------------
// k_pack_reads turns the ASCII batch into 2-bit words once (coalesced, one thread per 32-base word); the mapping kernels then
// fetch a read as a few 8-byte words.  pk / pkn: RWP words per (fragment, mate); meta: length | has_invalid_base << 16.
__global__ void k_pack_reads(const char* __restrict__ bases1, const uint64_t* __restrict__ off1, const char* __restrict__ bases2,
                             const uint64_t* __restrict__ off2, uint64_t n_frags, int n_mates, uint32_t rwp,
                             uint64_t* __restrict__ pk, uint64_t* __restrict__ pkn, uint32_t* __restrict__ meta) {
    const uint64_t gid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t total = n_frags * n_mates * rwp;
    if (gid >= total) return;
    const uint32_t w = (uint32_t)(gid % rwp);
    const uint64_t fm = gid / rwp;
    const uint64_t frag = fm / n_mates; const int mate = (int)(fm % n_mates);
    const char* bases = mate ? bases2 : bases1; const uint64_t* off = mate ? off2 : off1;
    const uint64_t beg = off[frag];
    uint64_t len64 = off[frag + 1] - beg;
    const uint32_t L = len64 > MAX_READ_LEN ? MAX_READ_LEN : (uint32_t)len64;
    uint64_t word = 0, bad = 0;
    const uint32_t i0 = w * 32;
    if (i0 < L) {
        const uint32_t n = (L - i0) < 32 ? (L - i0) : 32;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(bases) + beg + i0;
        uint32_t j = 0;
        if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
            // four bases per 32-bit load, converted with byte-parallel arithmetic
            for (; j + 4 <= n; j += 4) {
                const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(src + j));
                const uint32_t up = x & 0xDFDFDFDFu;
                uint32_t okm = 0;                                    // 0x80 in every byte that is A, C, G or T
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t K = q == 0 ? 0x41414141u : q == 1 ? 0x43434343u : q == 2 ? 0x47474747u : 0x54545454u;
                    const uint32_t z = up ^ K;
                    okm |= ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);     // exact zero-byte detector
                }
                uint32_t code = ((x >> 1) ^ (x >> 2)) & 0x03030303u;                 // A 0, C 1, G 2, T 3 in every byte
                const uint32_t okb = okm >> 7;                                       // 1 per valid byte
                code &= okb * 3u;                                                    // invalid bases carry code 0
                const uint32_t c8 = (code | (code >> 6) | (code >> 12) | (code >> 18)) & 0xFFu;
                const uint32_t nb = (~okb) & 0x01010101u;
                const uint32_t b8 = (nb | (nb >> 6) | (nb >> 12) | (nb >> 18)) & 0x55u;
                word |= (uint64_t)c8 << (2 * j);
                bad |= (uint64_t)b8 << (2 * j);
            }
        }
        for (; j < n; ++j) {
            const unsigned char ch = src[j];
            const unsigned char up = ch & 0xDF;
            const bool ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
            word |= (ok ? (uint64_t)(((ch >> 1) ^ (ch >> 2)) & 3) : 0ULL) << (2 * j);
            bad |= (ok ? 0ULL : 1ULL) << (2 * j);
        }
    }
    pk[gid] = word; pkn[gid] = bad;
    if (w == 0) atomicOr(meta + fm, L);                 // meta is zeroed before the launch; other words may add the flag
    if (bad) atomicOr(meta + fm, 1u << 16);
}

__global__ void k_max_read_len(const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2, uint64_t n_frags,
                               unsigned int* __restrict__ out, unsigned long long* __restrict__ clipped) {
    unsigned int mx = 0, nclip = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_frags; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t l = off1[i + 1] - off1[i];
        nclip += l > MAX_READ_LEN;
        if (off2) { const uint64_t l2 = off2[i + 1] - off2[i]; nclip += l2 > MAX_READ_LEN; l = l2 > l ? l2 : l; }
        if (l > MAX_READ_LEN) l = MAX_READ_LEN;
        mx = (unsigned int)l > mx ? (unsigned int)l : mx;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const unsigned int o = __shfl_xor_sync(0xffffffffu, mx, m); mx = o > mx ? o : mx; }
    if ((threadIdx.x & 31u) == 0 && mx) atomicMax(out, mx);
    if (nclip) atomicAdd(clipped, (unsigned long long)nclip);          // rare: no reduction
}

// packed read -> this lane's shared-memory column (both orientations) + the invalid-base masks when there are any
__device__ void load_packed(const uint64_t* __restrict__ pk, const uint64_t* __restrict__ pkn, const uint32_t* __restrict__ meta,
                            uint64_t fm, uint32_t rwp, Read& r, ReadN& rn) {
    const uint32_t me = meta[fm];
    const uint32_t L = me & 0xFFFFu;
    r.len = L;
    r.has_n = (me >> 16) & 1u;
    const uint32_t nw = (L + 31) >> 5;
    const uint64_t* src = pk + fm * rwp;
    for (uint32_t w = 0; w < RW; ++w) smem_reads[r.sb + w * 32] = w < nw ? __ldg(src + w) : 0ULL;
    // reverse complement: reversed words in reverse order, then shifted down by the padding of the last word
    const uint32_t pad = nw * 32 - L;
    for (uint32_t w = 0; w < RW; ++w) {
        uint64_t v = 0;
        if (w < nw) {
            // rc base j = comp(fw base L-1-j); in "reversed padded" coordinates that is position j + pad
            const uint32_t pos = w * 32 + pad, idx = pos >> 5, sh = 2 * (pos & 31);
            const uint64_t lo = idx < nw ? revcomp32(smem_reads[r.sb + (nw - 1 - idx) * 32]) : 0ULL;
            const uint64_t hi = (idx + 1) < nw ? revcomp32(smem_reads[r.sb + (nw - 2 - idx) * 32]) : 0ULL;
            v = lo >> sh;
            if (sh) v |= hi << (64 - sh);
            const uint32_t rem = L - w * 32;                            // bases of this word that exist
            if (rem < 32) v &= (1ULL << (2 * rem)) - 1;
        }
        smem_reads[r.sb + (RW + w) * 32] = v;
    }
    if (r.has_n) {
        const uint64_t* srcn = pkn + fm * rwp;
        for (uint32_t w = 0; w < RW; ++w) { rn.nm[0][w] = w < nw ? __ldg(srcn + w) : 0ULL; rn.nm[1][w] = 0ULL; }
        // masks of the reverse orientation: plain reversal of the 2-bit groups (no complement)
        for (uint32_t w = 0; w < nw; ++w) {
            const uint32_t pos = w * 32 + pad, idx = pos >> 5, sh = 2 * (pos & 31);
            auto rev = [](uint64_t x) { x = __brevll(x); return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1); };
            const uint64_t lo = idx < nw ? rev(rn.nm[0][nw - 1 - idx]) : 0ULL;
            const uint64_t hi = (idx + 1) < nw ? rev(rn.nm[0][nw - 2 - idx]) : 0ULL;
            uint64_t v = lo >> sh;
            if (sh) v |= hi << (64 - sh);
            rn.nm[1][w] = v;
        }
    }
}

// longest common extension of read[qpos..) with text[p..p+rem), counted from the k-mer start (always >= k); rem = bases from p to the
// end of p's transcript (carried by the suffix entry)
__device__ __forceinline__ uint32_t lcp_at(const IndexView& ix, const Read& r, const ReadN& rn, int o, uint32_t qpos, uint64_t p, uint32_t rem) {
    const uint32_t lim_r = r.len - qpos;
    const uint32_t lim = rem < lim_r ? rem : lim_r;
    uint32_t m = ix.k;
    while (m < lim) {
        const uint64_t x = win32(r, o, qpos + m) ^ win32g(ix.words, p + m);
        uint64_t y = (x | (x >> 1)) & 0x5555555555555555ULL;
        if (r.has_n) y |= win32n(rn, o, qpos + m);
        if (y) { m += static_cast<uint32_t>(__ffsll(static_cast<long long>(y)) - 1) >> 1; break; }
        m += 32;
    }
    return m < lim ? m : lim;
}

// mask: buckets of at most 32 positions -- which entries reach the maximal match m; larger buckets -- offset of the bucket's words in
// the chunk's IvPool (one word per 32 entries, see extend_coop), or IV_NO_EXT
struct Interval { uint32_t lb, cnt, qpos, m, mask; };
constexpr uint32_t IV_NO_EXT = 0xFFFFFFFFu;
struct IvPool { unsigned long long* words; unsigned long long* cursor; unsigned long long cap; };

// continue a table lookup whose first slot (already loaded) held another k-mer: linear probing from the next slot
__device__ __forceinline__ bool table_find_from(const IndexView& ix, uint64_t km, uint64_t h, uint32_t& lb, uint32_t& cnt) {
    for (;;) {
        h = (h + 1) & ix.mask;
        const uint4 sl = __ldg(ix.table + h);
        if (sl.w == 0) return false;
        if ((((uint64_t)sl.y << 32) | sl.x) == km) { lb = sl.z; cnt = sl.w; return true; }
    }
}

// intervals travel between the scan and the finalize kernel as one u64: lb | cnt << 32 | qpos << 46 | m << 55
__device__ __forceinline__ unsigned long long pack_iv(uint32_t lb, uint32_t cnt, uint32_t qpos, uint32_t m) {
    return (unsigned long long)lb | ((unsigned long long)cnt << 32) | ((unsigned long long)qpos << 46) | ((unsigned long long)m << 55);
}
__device__ __forceinline__ Interval unpack_iv(unsigned long long v) {
    Interval r; r.mask = 0; r.lb = (uint32_t)v; r.cnt = (uint32_t)(v >> 32) & 0x3FFFu; r.qpos = (uint32_t)(v >> 46) & 0x1FFu; r.m = (uint32_t)(v >> 55);
    return r;
}

// Spec v1 seed scans of ALL orientations of ALL mates of one fragment form ONE sequence of steps: scan s = 2*mate +
// orientation, position i, n intervals found so far.  A fragment has one cheap scan per mate (the orientation that matches:
// a hit, an extension, done) and one expensive one (the other strand: every k-mer absent).  One step decides as many positions as
// one round of loads allows, with the m-mer bitmap (kmer_filter.hpp): the LAST m-mer of the k-mer at i lies inside the k-mers of
// i .. i+J-1 (J = k-m+1), so if it is absent all J positions are decided -- and then the same test at i+J, whose load is already in
// flight, may decide J more; the m-mer in the MIDDLE of the k-mer decides J/2+1 positions when the last one is present.  Only a
// position that passes all of them is looked up in the k-mer table.  Skipped positions are exactly positions whose k-mer is
// absent (or whose window holds an invalid base -- never a seed either), so the result equals the one-position-at-a-time scan's.
// Returns true when the fragment's last scan has ended.
#ifndef SFB_EXT_BATCH
#define SFB_EXT_BATCH 4      // pending seeds a warp collects before it extends them together (extend_coop)
#endif
#ifndef SFB_SCAN_BLOCKS
#define SFB_SCAN_BLOCKS 4
#endif
struct ScanState { int s, n; uint32_t i; uint32_t pend_lb, pend_cnt; };   // pend_cnt != 0: a seed waits for its extension
__device__ __forceinline__ bool mf_test(const IndexView& ix, uint64_t key) {
    return (__ldg(ix.mfilter + (key >> 5)) >> (key & 31)) & 1u;
}
__device__ __forceinline__ bool scan_step(const IndexView& ix, const Read* rds, const ReadN* rns, int ns, uint32_t max_interval, ScanState& st,
                                          uint8_t* __restrict__ niv_out /* [ns] of this fragment */) {
    const uint32_t k = ix.k;
    const bool m1 = (st.s >> 1) != 0;
    Read r; r.sb = m1 ? rds[1].sb : rds[0].sb; r.len = m1 ? rds[1].len : rds[0].len; r.has_n = m1 ? rds[1].has_n : rds[0].has_n;
    const int o = st.s & 1;
    const uint32_t L = r.len;
    const uint32_t i = st.i;
    if (!(i + k <= L && st.n < MAX_IV)) { niv_out[st.s] = (uint8_t)st.n; ++st.s; st.i = 0; st.n = 0; return st.s >= ns; }
    if (r.has_n) {
        const uint64_t nn = win32n(rns[st.s >> 1], o, i) & ix.kmask;
        if (nn) { st.i = i + ((63 - __clzll(static_cast<long long>(nn))) >> 1) + 1; return false; }   // jump past the last invalid base
    }
    const int m = ix.mf_m;
    const uint32_t J = k - m + 1;
    const uint64_t km = win32(r, o, i) & ix.kmask;
    // the look-ahead position i+J: only when its window lies inside the read and (reads with invalid bases) is clean
    bool ahead = i + J + k <= L;
    uint64_t km2 = 0;
    if (ahead) {
        km2 = win32(r, o, i + J) & ix.kmask;
        if (r.has_n && (win32n(rns[st.s >> 1], o, i + J) & ix.kmask) != 0) ahead = false;
    }
    const uint32_t half = (J - 1) >> 1;
    const bool b_last = mf_test(ix, sfb_mfilter_key(km, k - m, m));
    const bool b_next = ahead ? mf_test(ix, sfb_mfilter_key(km2, k - m, m)) : true;
    const bool b_mid = half ? mf_test(ix, sfb_mfilter_key(km, half, m)) : true;
    if (!b_last) { st.i = i + ((ahead && !b_next) ? 2 * J : J); return false; }
    if (!b_mid) { st.i = i + half + 1; return false; }                 // the m-mer at `half` lies inside the k-mers of i .. i+half
    // homopolymer k-mers are never seeds: a k-mer equals itself shifted by one base
    const bool homo = ((km ^ (km >> 2)) & (ix.kmask >> 2)) == 0;
    bool hit = false;
    uint32_t lb = 0, cnt = 0;
    if (!homo) {
        const uint64_t hh = sfb_kmer_mix(km);
        const uint64_t h0 = (hh & (ix.mask >> 1)) << 1;               // even slot: h0 and h0+1 share a 32-byte sector
        const uint4 sl = __ldg(ix.table + h0), sl2 = __ldg(ix.table + h0 + 1);
        if (sl.w != 0) {
            if ((((uint64_t)sl.y << 32) | sl.x) == km) { hit = true; lb = sl.z; cnt = sl.w; }
            else if (sl2.w != 0) {
                if ((((uint64_t)sl2.y << 32) | sl2.x) == km) { hit = true; lb = sl2.z; cnt = sl2.w; }
                else hit = table_find_from(ix, km, h0 + 1, lb, cnt);
            }
        }
    }
    if (hit && cnt <= max_interval) {
        // a seed: the extension over its bucket is done later, together with other lanes' (see k_scan_reads)
        st.pend_lb = lb; st.pend_cnt = cnt;
        return false;
    }
    st.i = i + 1;
    return false;
}

// match extension of a pending seed: m = longest match over the bucket, mask = which of its first 32 entries reach it
__device__ __forceinline__ void extend_seed(const IndexView& ix, const Read* rds, const ReadN* rns, ScanState& st,
                                            unsigned long long* __restrict__ iv_out, uint32_t* __restrict__ ivmask_out) {
    const bool m1 = (st.s >> 1) != 0;
    Read r; r.sb = m1 ? rds[1].sb : rds[0].sb; r.len = m1 ? rds[1].len : rds[0].len; r.has_n = m1 ? rds[1].has_n : rds[0].has_n;
    const ReadN& rn = rns[st.s >> 1];
    const int o = st.s & 1;
    const uint32_t q = st.i, lb = st.pend_lb, cnt = st.pend_cnt;
    uint32_t m = 0, mask = 0;
    for (uint32_t e = 0; e < cnt; ++e) {
        const uint2 en = sa_pos_rem(ix, lb + e);
        const uint32_t l = lcp_at(ix, r, rn, o, q, en.x, en.y);
        if (l > m) { m = l; mask = e < 32 ? (1u << e) : 0u; }
        else if (l == m && e < 32) mask |= 1u << e;
    }
    iv_out[st.s * MAX_IV + st.n] = pack_iv(lb, cnt, q, m);
    ivmask_out[st.s * MAX_IV + st.n] = cnt <= 32u ? mask : IV_NO_EXT;                  // big bucket: the finalize kernel extends again
    ++st.n;
    st.i = q + m - ix.k + 1;                                                           // next k-mer ends one base past the match
    st.pend_cnt = 0;
}

// Match extension of the warp's pending seeds, done by the whole warp.  A seed's extension is a chain of dependent random loads
// (suffix entry -> text words) per bucket entry; run by the lane that found the seed it keeps 1-3 lanes of the warp busy for
// ~cnt x 3 memory round trips.  Here the warp splits into 8 groups of 4 lanes, a group takes one pending seed (its owner's packed read
// is readable by every lane: it lives in shared memory), the group's lanes take the bucket entries round-robin, and a lane has the
// first two text windows of its entry in flight together.  Seeds of reads with invalid bases stay with their owner (the masks are
// in its local memory): extend_seed below.
#ifndef SFB_EXT_GROUP
#define SFB_EXT_GROUP 4
#endif
constexpr unsigned EXT_G = SFB_EXT_GROUP, EXT_NG = 32 / EXT_G;
__device__ __forceinline__ uint32_t lcp_clean(const IndexView& ix, uint32_t sb, uint32_t len, int o, uint32_t qpos, uint64_t p, uint32_t rem) {
    const uint32_t lim_r = len - qpos;
    const uint32_t lim = rem < lim_r ? rem : lim_r;
    uint32_t m = ix.k;
    if (m >= lim) return lim;
    // text windows [p+m, p+m+32) and [p+m+32, p+m+64): three words, loaded together
    const uint64_t P = p + m;
    const uint64_t idx = P >> 5; const uint32_t sh = 2 * (P & 31);
    const uint64_t w0 = __ldg(ix.words + idx), w1 = __ldg(ix.words + idx + 1), w2 = __ldg(ix.words + idx + 2);
    const uint64_t t0 = sh ? ((w0 >> sh) | (w1 << (64 - sh))) : w0;
    const uint64_t t1 = sh ? ((w1 >> sh) | (w2 << (64 - sh))) : w1;
    auto rwin = [&](uint32_t pos) {
        const uint32_t i = pos >> 5, s2 = 2 * (pos & 31);
        uint64_t v = smem_reads[sb + (o * RW + i) * 32] >> s2;
        if (s2) v |= smem_reads[sb + (o * RW + i + 1) * 32] << (64 - s2);
        return v;
    };
    uint64_t x = rwin(qpos + m) ^ t0;
    uint64_t y = (x | (x >> 1)) & 0x5555555555555555ULL;
    if (y) { m += static_cast<uint32_t>(__ffsll(static_cast<long long>(y)) - 1) >> 1; return m < lim ? m : lim; }
    m += 32;
    if (m >= lim) return lim;
    x = rwin(qpos + m) ^ t1;
    y = (x | (x >> 1)) & 0x5555555555555555ULL;
    if (y) { m += static_cast<uint32_t>(__ffsll(static_cast<long long>(y)) - 1) >> 1; return m < lim ? m : lim; }
    m += 32;
    while (m < lim) {
        x = rwin(qpos + m) ^ win32g(ix.words, p + m);
        y = (x | (x >> 1)) & 0x5555555555555555ULL;
        if (y) { m += static_cast<uint32_t>(__ffsll(static_cast<long long>(y)) - 1) >> 1; break; }
        m += 32;
    }
    return m < lim ? m : lim;
}

// all 32 lanes call this; todo = lanes whose pending seed belongs to a read without invalid bases
__device__ __forceinline__ void extend_coop(const IndexView& ix, const Read* rds, ScanState& st, unsigned todo, unsigned lane,
                                            unsigned long long* __restrict__ iv_all, uint32_t* __restrict__ ivmask_all, uint64_t iv_base,
                                            const IvPool pool) {
    const unsigned grp = lane / EXT_G, sub = lane % EXT_G;
    const uint32_t my_len = (st.s >> 1) ? rds[1].len : rds[0].len;     // only read from lanes that own a pending seed
    // seeds with large buckets first (paralog families, repeats: hundreds of positions): the whole warp takes one seed, a bucket entry per
    // lane per round -- in a 4-lane group such a seed would keep the other 28 lanes waiting for cnt / 4 rounds of dependent loads
    unsigned big = todo & __ballot_sync(0xffffffffu, st.pend_cnt > 4u * EXT_G);
    todo &= ~big;
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t lb = __shfl_sync(0xffffffffu, st.pend_lb, src), cnt = __shfl_sync(0xffffffffu, st.pend_cnt, src);
        const uint32_t q = __shfl_sync(0xffffffffu, st.i, src), len = __shfl_sync(0xffffffffu, my_len, src);
        const int ss = __shfl_sync(0xffffffffu, st.s, src);
        const uint32_t sb = ((ss >> 1) ? rds[1].sb : rds[0].sb) + (unsigned)src - lane;
        uint32_t best = 0, mask = 0;
        if (cnt <= 32u) {
            if (lane < cnt) {
                const uint2 en = sa_pos_rem(ix, lb + lane);
                best = lcp_clean(ix, sb, len, ss & 1, q, en.x, en.y);
                mask = 1u << lane;
            }
#pragma unroll
            for (unsigned d = 1; d < 32; d <<= 1) {
                const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), om = __shfl_xor_sync(0xffffffffu, mask, d);
                if (ob > best) { best = ob; mask = om; } else if (ob == best) mask |= om;
            }
        } else {
            // more than 32 positions: one word per round of 32 entries goes to the chunk's pool -- the round's longest match in the high
            // half, which of its entries reach it in the low half; an entry reaches the seed's maximal match iff its round's maximum
            // IS that match and its bit is set.  The interval's mask word carries the pool offset (IV_NO_EXT: pool exhausted, the
            // finalize kernel then extends the entries itself)
            const uint32_t rounds = (cnt + 31u) >> 5;
            unsigned long long po = 0;
            if (lane == 0) po = atomicAdd(pool.cursor, (unsigned long long)rounds);
            po = __shfl_sync(0xffffffffu, po, 0);
            mask = po + rounds <= pool.cap ? (uint32_t)po : IV_NO_EXT;
            for (uint32_t r = 0; r < rounds; ++r) {
                const uint32_t e = r * 32u + lane;
                uint32_t l = 0;
                if (e < cnt) { const uint2 en = sa_pos_rem(ix, lb + e); l = lcp_clean(ix, sb, len, ss & 1, q, en.x, en.y); }
                const uint32_t mr = __reduce_max_sync(0xffffffffu, l);
                const unsigned bits = __ballot_sync(0xffffffffu, e < cnt && l == mr);
                if (lane == 0 && mask != IV_NO_EXT) pool.words[mask + r] = ((unsigned long long)mr << 32) | bits;
                best = mr > best ? mr : best;
            }
        }
        if ((int)lane == src) {
            iv_all[iv_base + st.s * MAX_IV + st.n] = pack_iv(st.pend_lb, st.pend_cnt, st.i, best);
            ivmask_all[iv_base + st.s * MAX_IV + st.n] = mask;
            ++st.n;
            st.i = st.i + best - ix.k + 1;
            st.pend_cnt = 0;
        }
    }
    while (todo) {
        const unsigned src = __fns(todo, 0, (int)grp + 1);             // owner lane of this group's seed, 0xFFFFFFFF if there is none
        const bool act = src < 32u;
        const unsigned sl = act ? src : lane;
        const uint32_t lb = __shfl_sync(0xffffffffu, st.pend_lb, sl);
        uint32_t cnt = __shfl_sync(0xffffffffu, st.pend_cnt, sl);
        if (!act) cnt = 0;                                             // a group without a seed idles through this round
        const uint32_t q = __shfl_sync(0xffffffffu, st.i, sl);
        const int ss = __shfl_sync(0xffffffffu, st.s, sl);
        const uint32_t len = __shfl_sync(0xffffffffu, my_len, sl);
        // the owner's packed mate in shared memory: same warp region, its lane's column
        const uint32_t sb = ((ss >> 1) ? rds[1].sb : rds[0].sb) + sl - lane;
        uint32_t best = 0, mask = 0;
        for (uint32_t e = sub; e < cnt; e += EXT_G) {
            const uint2 en = sa_pos_rem(ix, lb + e);
            const uint32_t l = lcp_clean(ix, sb, len, ss & 1, q, en.x, en.y);
            if (l > best) { best = l; mask = e < 32 ? (1u << e) : 0u; }
            else if (l == best && e < 32) mask |= 1u << e;
        }
#pragma unroll
        for (unsigned d = 1; d < EXT_G; d <<= 1) {
            const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), om = __shfl_xor_sync(0xffffffffu, mask, d);
            if (ob > best) { best = ob; mask = om; } else if (ob == best) mask |= om;
        }
        // hand the result back to the owners served in this round: the first EXT_NG set bits of todo
        const unsigned rank = __popc(todo & ((1u << lane) - 1));
        const bool mine = ((todo >> lane) & 1u) && rank < EXT_NG;
        const unsigned from = mine ? rank * EXT_G : lane;
        const uint32_t m = __shfl_sync(0xffffffffu, best, from), mk = __shfl_sync(0xffffffffu, mask, from);
        if (mine) {
            iv_all[iv_base + st.s * MAX_IV + st.n] = pack_iv(st.pend_lb, st.pend_cnt, st.i, m);
            ivmask_all[iv_base + st.s * MAX_IV + st.n] = mk;
            ++st.n;
            st.i = st.i + m - ix.k + 1;
            st.pend_cnt = 0;
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

// Per-thread hit lists.  Five regions of max_read_occs+1 entries per thread: the two orientation projections of the mate being
// collected (R_A, R_B), the left and right mate's merged lists, and the label under construction.  The first FIN_S entries of every
// region live in SHARED memory (lane-interleaved, [region][entry][lane] per warp) -- a read rarely hits more than a handful of
// transcripts -- the rest in a thread-interleaved global scratch (element j of thread t at base[j * stride + t], coalesced like local
// memory).  Lists are written, merged, filtered, hashed and compared entry by entry, each step a dependent access: in shared memory
// that is ~30 cycles, through L2 it was ~300 (ncu, round 1: 18% issue utilisation).
constexpr uint32_t FIN_S = 6;
enum { R_A = 0, R_B = 1, R_LEFT = 2, R_RIGHT = 3, R_LABEL = 4, N_REGIONS = 5 };
extern __shared__ unsigned long long smem_hits[];      // finalize kernel: [warp][region][entry][lane]
struct Scratch {
    uint32_t so;                   // index of this lane's column of its warp's block in smem_hits
    unsigned long long* base;      // this thread's column of the global scratch
    uint64_t stride; uint32_t cap1;
    __device__ __forceinline__ unsigned long long get(uint32_t region, uint32_t j) const {
        if (j < FIN_S) return smem_hits[so + (region * FIN_S + j) * 32];
        return base[(uint64_t)(region * cap1 + j) * stride];
    }
    __device__ __forceinline__ void set(uint32_t region, uint32_t j, unsigned long long v) const {
        if (j < FIN_S) smem_hits[so + (region * FIN_S + j) * 32] = v;
        else base[(uint64_t)(region * cap1 + j) * stride] = v;
    }
    // the same entry of another lane of this warp (delta = that lane - this lane)
    __device__ __forceinline__ unsigned long long peer(int delta, uint32_t region, uint32_t j) const {
        if (j < FIN_S) return smem_hits[(int)(so + (region * FIN_S + j) * 32) + delta];
        return base[(int64_t)((uint64_t)(region * cap1 + j) * stride) + delta];
    }
};
__device__ __forceinline__ unsigned long long pack_hit(uint32_t tid, int32_t pos, bool fwd) {
    return ((unsigned long long)tid << 32) | (uint32_t)(((pos + POS_BIAS) << 1) | (fwd ? 1 : 0));
}
__device__ __forceinline__ uint32_t hit_tid(unsigned long long h) { return (uint32_t)(h >> 32); }
__device__ __forceinline__ int32_t hit_pos(unsigned long long h) { return (int32_t)(((uint32_t)h) >> 1) - POS_BIAS; }
__device__ __forceinline__ bool hit_fwd(unsigned long long h) { return (h & 1ULL) != 0; }

// The finalize kernel does not stage the packed reads: it needs read bases only for bucket entries beyond the 32 the scan's mask
// covers (buckets of more than 32 positions).  Those few extensions read the packed mate straight from global memory.
struct ReadG { const uint64_t* pk; const uint64_t* pkn; uint32_t len; bool has_n; };
__device__ __forceinline__ uint64_t gwin32(const uint64_t* __restrict__ w, uint32_t pos) {
    const uint32_t idx = pos >> 5, sh = 2 * (pos & 31);
    uint64_t v = w[idx] >> sh;
    if (sh) v |= w[idx + 1] << (64 - sh);
    return v;
}
// 32-base window at `pos` of orientation o (1 = reverse complement; plain reversal for the invalid-base masks)
__device__ __forceinline__ uint64_t readg_win(const uint64_t* __restrict__ w, uint32_t len, int o, uint32_t pos, bool complement) {
    if (o == 0) return gwin32(w, pos);
    // rc base j = comp(fw base len-1-j): the window covers fw bases [len-pos-32, len-pos), reversed
    const int32_t s = (int32_t)len - (int32_t)pos - 32;
    uint64_t f = gwin32(w, s >= 0 ? (uint32_t)s : 0u);
    uint64_t x = __brevll(complement ? ~f : f);
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    return s >= 0 ? x : (x >> (2 * (uint32_t)(-s)));
}
__device__ __noinline__ uint32_t lcp_global(const IndexView& ix, const ReadG& r, int o, uint32_t qpos, uint64_t p, uint32_t rem) {
    const uint32_t lim_r = r.len - qpos;
    const uint32_t lim = rem < lim_r ? rem : lim_r;
    uint32_t m = ix.k;
    while (m < lim) {
        const uint64_t x = readg_win(r.pk, r.len, o, qpos + m, true) ^ win32g(ix.words, p + m);
        uint64_t y = (x | (x >> 1)) & 0x5555555555555555ULL;
        if (r.has_n) y |= readg_win(r.pkn, r.len, o, qpos + m, false);
        if (y) { m += static_cast<uint32_t>(__ffsll(static_cast<long long>(y)) - 1) >> 1; break; }
        m += 32;
    }
    return m < lim ? m : lim;
}

// which of the bucket entries [base, base + 32) of interval a reach its maximal match
__device__ __forceinline__ uint32_t reach_bits(const IndexView& ix, const ReadG& r, int o, const Interval& a, uint32_t base, const IvPool& pool) {
    if (a.cnt <= 32u) return a.mask;
    if (a.mask != IV_NO_EXT) {
        const unsigned long long w = __ldg(pool.words + a.mask + (base >> 5));
        return (uint32_t)(w >> 32) == a.m ? (uint32_t)w : 0u;
    }
    uint32_t bits = 0;                                   // no pool words (pool exhausted, or a read with invalid bases): extend here
    const uint32_t end = a.cnt - base < 32u ? a.cnt - base : 32u;
    for (uint32_t j = 0; j < end; ++j) {
        const uint2 pr = sa_pos_rem(ix, a.lb + base + j);
        if (lcp_global(ix, r, o, a.qpos, pr.x, pr.y) == a.m) bits |= 1u << j;
    }
    return bits;
}
__device__ __forceinline__ bool reach_one(const IndexView& ix, const ReadG& r, int o, const Interval& b, uint32_t e_rel, const IvPool& pool) {
    if (b.cnt <= 32u) return (b.mask >> e_rel) & 1u;
    if (b.mask != IV_NO_EXT) {
        const unsigned long long w = __ldg(pool.words + b.mask + (e_rel >> 5));
        return (uint32_t)(w >> 32) == b.m && ((w >> (e_rel & 31u)) & 1ULL);
    }
    const uint2 pr = sa_pos_rem(ix, b.lb + e_rel);
    return lcp_global(ix, r, o, b.qpos, pr.x, pr.y) == b.m;
}

// transcripts present with the maximal match in every interval; output ascending by transcript id;
// stops after cap+1 hits (list overflow).  The first four suffix entries of the first interval's bucket are loaded together (most
// buckets are no larger), so the walk over the bucket is one memory round trip instead of one per entry; only entries that reach
// the maximal match are visited (their bits: the scan's mask, or its pool words for buckets of more than 32 positions).
__device__ uint32_t project(const IndexView& ix, const ReadG& r, int o, const Interval* ivs, int niv, uint32_t cap,
                            const Scratch& out, uint32_t region, const IvPool& pool) {
    if (niv == 0) return 0;
    uint32_t n = 0;
    const Interval a = ivs[0];
    uint32_t lastTid = 0xFFFFFFFFu;                 // transcript ids are < 2^32 - 1
    uint2 pre0 = make_uint2(0, 0), pre1 = pre0, pre2 = pre0, pre3 = pre0;
    pre0 = sa_tid_rel(ix, a.lb);
    if (a.cnt > 1) pre1 = sa_tid_rel(ix, a.lb + 1);
    if (a.cnt > 2) pre2 = sa_tid_rel(ix, a.lb + 2);
    if (a.cnt > 3) pre3 = sa_tid_rel(ix, a.lb + 3);
    for (uint32_t base = 0; base < a.cnt; base += 32u) {
        uint32_t bits = reach_bits(ix, r, o, a, base, pool);
        while (bits) {
            const uint32_t e_rel = base + (uint32_t)__ffs((int)bits) - 1u;
            bits &= bits - 1u;
            uint2 en;
            if (e_rel < 4) en = e_rel == 0 ? pre0 : e_rel == 1 ? pre1 : e_rel == 2 ? pre2 : pre3;
            else en = sa_tid_rel(ix, a.lb + e_rel);
            const uint32_t tid = en.x;
            if (tid == lastTid) continue;
            lastTid = tid;
            bool all = true;
            for (int j = 1; j < niv && all; ++j) {
                const Interval b = ivs[j];
                uint32_t lo = b.lb, hi = b.lb + b.cnt;
                while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (__ldg(&ix.sa[mid].x) < tid) lo = mid + 1; else hi = mid; }
                bool found = false;
                for (uint32_t e2 = lo; e2 < b.lb + b.cnt; ++e2) {
                    if (__ldg(&ix.sa[e2].x) != tid) break;
                    if (reach_one(ix, r, o, b, e2 - b.lb, pool)) { found = true; break; }
                }
                all = found;
            }
            if (!all) continue;
            const int32_t pos = (int32_t)en.y - (int32_t)a.qpos;
            out.set(region, n, pack_hit(tid, pos, o == 0));
            ++n;
            if (n > cap) return n;
        }
    }
    return n;
}

// one mate: both orientations, optional strand vote, merge by transcript id.  Returns false on list overflow.  The list ends up in
// region `where` (a projection region when only one orientation hit -- the usual case -- else `dst`)
__device__ bool collect(const IndexView& ix, const ReadG& r, bool strict, uint32_t cap, const Interval* ivF, int nivF,
                        uint64_t scF, const Interval* ivR, int nivR, uint64_t scR, const Scratch& scr,
                        uint32_t dst, uint32_t& n_out, uint32_t& where, const IvPool& pool) {
    n_out = 0; where = dst;
    uint32_t nF = project(ix, r, 0, ivF, nivF, cap, scr, R_A, pool);
    uint32_t nR = project(ix, r, 1, ivR, nivR, cap, scr, R_B, pool);
    if (nF > cap || nR > cap) return false;
    if (strict && nF && nR) { if (scF > scR) nR = 0; else if (scR > scF) nF = 0; }
    if (nF + nR > cap) return false;
    if (nR == 0) { n_out = nF; where = R_A; return true; }
    if (nF == 0) { n_out = nR; where = R_B; return true; }
    uint32_t i = 0, j = 0, n = 0;
    while (i < nF || j < nR) {                     // stable merge: forward before reverse on equal transcript id
        bool takeF;
        if (i >= nF) takeF = false; else if (j >= nR) takeF = true;
        else takeF = hit_tid(scr.get(R_A, i)) <= hit_tid(scr.get(R_B, j));
        const unsigned long long h = takeF ? scr.get(R_A, i++) : scr.get(R_B, j++);
        scr.set(dst, n++, h);
    }
    n_out = n;
    return true;
}

struct MapParams {
    IndexView ix;
    EqTable tb;
    const char* bases1; const uint64_t* off1; const char* bases2; const uint64_t* off2;
    uint64_t n_reads;
    uint32_t cap;              // max_read_occs
    uint32_t max_frag_len, max_interval;
    int lib_fmt, strict_intersect, allow_orphans, allow_dovetail, ignore_compat, enforce_compat;
    unsigned long long* scratch; uint64_t n_threads_total;
    unsigned long long* counters;      // 6
    unsigned long long* next_read;     // work counters: [0] scan kernel, [1] finalize kernel, [2..4] reads set aside for the heavy pass, by bin
    int16_t* fld_val;                  // per read of the batch: fragment length if FLD-eligible, else -1
    // packed batch and the scan -> finalize hand-over
    const uint64_t* pk; const uint64_t* pkn; const uint32_t* meta; uint32_t rwp; int n_mates;
    unsigned long long* iv; uint8_t* niv; uint32_t* ivmask; IvPool pool;
    // bias / GC sample collection (k_finalize_reads_bias only)
    // class-table growth: reads whose upsert failed are listed (retry_out) and finalized again after the table has grown (retry_in)
    uint32_t* retry_out; const uint32_t* retry_in; uint64_t n_retry;
    // reads with a big seed bucket are set aside by the main finalize pass (heavy_out: three bins of n_reads entries, counted in
    // next_read[2..4]) and finalized by a second launch (heavy_in): map_finalize_body.inl
    uint32_t* heavy_out; const uint32_t* heavy_in;
    int16_t* bias_val;                 // per read of the batch: bin of its read-start context, or -1
    unsigned int* gc_hist;             // observed fragment GC histogram (101 bins), accumulated over batches
    int bias_seq, bias_gc;
};

struct LabelAcc {                       // the txpIDsAll / txpIDsCompat pair of processReadsQuasi folded into one buffer
    const Scratch& scr; uint32_t n; bool haveCompat; int32_t fw, rc; bool enforce;
    __device__ LabelAcc(const Scratch& s, bool enf) : scr(s), n(0), haveCompat(false), fw(0), rc(0), enforce(enf) {}
    __device__ __forceinline__ void add(uint32_t tid, bool compat, bool fwdHit) {
        if (compat) {
            if (!haveCompat) { haveCompat = true; n = 0; fw = 0; rc = 0; }      // switch from "all" to "compatible only"
            scr.set(R_LABEL, n++, tid); if (fwdHit) ++fw; else ++rc;
        } else if (!haveCompat && !enforce) {
            scr.set(R_LABEL, n++, tid); if (fwdHit) ++fw; else ++rc;
        }
    }
};

// ---- scan kernel: lanes pull fragments independently ------------------------------------------------------------------------
// The work per fragment varies (a read with a sequencing error scans ~30 more positions on the matching strand), so a
// warp that processes 32 fragments in lock step idles most of its lanes while the slowest finishes (measured: 13 of 32
// lanes busy).  Here a lane that has finished its fragment takes the next one from its warp's reservation (64 fragments per
// global atomic) and the seed intervals go to global memory for the finalize kernel.
__global__ void __launch_bounds__(MAP_THREADS, SFB_SCAN_BLOCKS) k_scan_reads(const MapParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const int n_mates = p.n_mates, ns = 2 * n_mates;
    Read rds[2];
    ReadN rns[2];
    rds[0].len = rds[1].len = 0; rds[0].has_n = rds[1].has_n = false;
    {
        const uint32_t wbase = (threadIdx.x >> 5) * n_mates * 2 * RW * 32 + lane;
        rds[0].sb = wbase;
        rds[1].sb = wbase + (n_mates - 1) * 2 * RW * 32;
    }
    bool have = false, done = false;
    uint32_t frag = 0;                                     // a chunk holds at most 2^22 fragments (map_chunk_device): 32-bit indices
    const uint32_t n_reads = (uint32_t)p.n_reads;
    ScanState st; st.s = 0; st.n = 0; st.i = 0; st.pend_lb = 0; st.pend_cnt = 0;
    uint32_t res_next = 0, res_end = 0;                    // this warp's reservation [res_next, res_end), warp-uniform
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have && !done);
        const unsigned busy = __ballot_sync(0xffffffffu, have);
        if (need && (busy == 0 || __popc(need) >= 8)) {     // refill several lanes at once: the refill code runs divergent
            if (res_next == res_end) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(p.next_read, 64ULL);
                b = __shfl_sync(0xffffffffu, b, 0);
                res_next = b < n_reads ? (uint32_t)b : n_reads;
                res_end = b + 64 < n_reads ? (uint32_t)b + 64 : n_reads;
            }
            const uint32_t avail = res_end - res_next;                 // 0 only when the global queue is exhausted
            if (!have && !done) {
                const unsigned my = __popc(need & ((1u << lane) - 1));
                if (my < avail) {
                    frag = res_next + my;
                    load_packed(p.pk, p.pkn, p.meta, (uint64_t)frag * n_mates, p.rwp, rds[0], rns[0]);
                    if (n_mates == 2) load_packed(p.pk, p.pkn, p.meta, (uint64_t)frag * n_mates + 1, p.rwp, rds[1], rns[1]);
                    st.s = 0; st.n = 0; st.i = 0; st.pend_cnt = 0;
                    have = true;
                } else if (avail == 0) {
                    done = true;
                }
            }
            const unsigned want = __popc(need);
            res_next += want < avail ? want : avail;
        }
        if (__ballot_sync(0xffffffffu, have) == 0 && __ballot_sync(0xffffffffu, !done) == 0) break;
        // extensions are batched: lanes with a pending seed wait until SFB_EXT_BATCH of them do (or nobody else can advance), then the
        // whole warp extends them together (extend_coop)
        bool pend = have && st.pend_cnt != 0;
        const unsigned pend_m = __ballot_sync(0xffffffffu, pend);
        const unsigned adv_m = __ballot_sync(0xffffffffu, have && !pend);
        if (pend_m && (__popc(pend_m) >= SFB_EXT_BATCH || adv_m == 0)) {
            const uint64_t ivb = frag * (uint64_t)(ns * MAX_IV);
            const bool dirty = pend && ((st.s >> 1) ? rds[1].has_n : rds[0].has_n);
            const unsigned clean_m = pend_m & ~__ballot_sync(0xffffffffu, dirty);
            if (clean_m) extend_coop(p.ix, rds, st, clean_m, lane, p.iv, p.ivmask, ivb, p.pool);
            if (dirty) extend_seed(p.ix, rds, rns, st, p.iv + ivb, p.ivmask + ivb);
            pend = false;                                  // every pending seed of the warp has been extended
        }
        if (have && !pend) {
            if (scan_step(p.ix, rds, rns, ns, p.max_interval, st, p.niv + (uint64_t)frag * ns)) have = false;
        }
    }
}

// ---- finalize kernel: projection, mate merge, compatibility filter, label, class upsert -----------------------------------------
#ifndef SFB_FIN_BLOCKS
#define SFB_FIN_BLOCKS 2      // 106 registers without spills; 3 CTAs (80 registers, spills) measured 5% slower (profiles/r02c_variants.txt)
#endif
#ifndef SFB_HEAVY_CNT
#define SFB_HEAVY_CNT 32u      // a read with a seed bucket larger than this is finalized in the heavy pass
#endif
__global__ void __launch_bounds__(MAP_THREADS, SFB_FIN_BLOCKS) k_finalize_reads(const MapParams p) {
#define SFB_FIN_BIAS 0
#include "map_finalize_body.inl"
#undef SFB_FIN_BIAS
}

__global__ void __launch_bounds__(MAP_THREADS, SFB_FIN_BLOCKS) k_finalize_reads_bias(const MapParams p) {
    __shared__ unsigned int s_gc[101];
    for (unsigned i = threadIdx.x; i < 101; i += blockDim.x) s_gc[i] = 0;
    __syncthreads();
    {
#define SFB_FIN_BIAS 1
#include "map_finalize_body.inl"
#undef SFB_FIN_BIAS
    }
    __syncthreads();
    if (p.bias_gc) for (unsigned i = threadIdx.x; i < 101; i += blockDim.x) if (s_gc[i]) atomicAdd(p.gc_hist + i, s_gc[i]);
}

// read-start contexts in global read order: the first `remaining` reads that have one (sfOpts.numBiasSamples, :270-285 at -p 1)
__global__ void k_bias_select(const int16_t* __restrict__ bias_val, uint64_t n_reads, unsigned int* __restrict__ hist, int* __restrict__ remaining) {
    __shared__ int s_base, s_rem;
    __shared__ int s_warp[32];
    if (threadIdx.x == 0) { s_rem = *remaining; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint64_t c0 = 0; c0 < n_reads; c0 += blockDim.x) {
        if (s_rem <= 0) break;
        const uint64_t i = c0 + threadIdx.x;
        const int v = i < n_reads ? bias_val[i] : -1;
        const unsigned bal = __ballot_sync(0xffffffffu, v >= 0);
        const int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int a = 0; for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { const int t = s_warp[w]; s_warp[w] = a; a += t; } s_base = a; }
        __syncthreads();
        if (v >= 0 && s_warp[warp] + pre < s_rem) atomicAdd(hist + v, 1u);
        __syncthreads();
        if (threadIdx.x == 0) s_rem -= s_base;
        __syncthreads();
    }
    if (threadIdx.x == 0) *remaining = s_rem < 0 ? 0 : s_rem;
}

// FLD sampling in global read order: the first `remaining` eligible fragments (SailfishQuantify.cpp:426-430 at -p 1)
__global__ void k_fld_select(const int16_t* __restrict__ fld_val, uint64_t n_reads, unsigned int* __restrict__ hist,
                             int* __restrict__ remaining, int16_t* __restrict__ samples, int total) {
    __shared__ int s_base, s_rem;
    __shared__ int s_warp[32];
    if (threadIdx.x == 0) { s_rem = *remaining; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint64_t c0 = 0; c0 < n_reads; c0 += blockDim.x) {
        if (s_rem <= 0) break;
        const uint64_t i = c0 + threadIdx.x;
        const int v = i < n_reads ? fld_val[i] : -1;
        const unsigned bal = __ballot_sync(0xffffffffu, v >= 0);
        const int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int a = 0; for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { const int t = s_warp[w]; s_warp[w] = a; a += t; } s_base = a; }
        __syncthreads();
        const int rank = s_warp[warp] + pre;
        if (v >= 0 && rank < s_rem) { atomicAdd(hist + v, 1u); samples[total - s_rem + rank] = (int16_t)v; }   // kept in read order
        __syncthreads();
        if (threadIdx.x == 0) s_rem -= s_base;
        __syncthreads();
    }
    if (threadIdx.x == 0) *remaining = s_rem < 0 ? 0 : s_rem;
}

// ---- eqBuilder.finish() on the device ---------------------------------------------------------------------------------------
// bin of a class by member count: 0..5 as in DevClasses (g = 2,4,8,16,32 lanes, then "long"), 6 = single-member
enum { FIN_CLS = 0, FIN_NNZ = 8, FIN_CUR_CLS = 16, FIN_ACTIVE = 32, FIN_TOTAL = 33, FIN_WORDS = 40 };
__device__ __forceinline__ int fin_bin(uint32_t n) { return n == 1 ? SFB_NBINS : n <= 2 ? 0 : n <= 4 ? 1 : n <= 8 ? 2 : n <= 16 ? 3 : n <= 32 ? 4 : 5; }

__global__ void k_eq_count(const unsigned long long* __restrict__ slot, uint64_t n_slots, unsigned long long* __restrict__ fin) {
    __shared__ unsigned long long s_cls[SFB_NBINS + 1], s_nnz[SFB_NBINS + 1];
    if (threadIdx.x <= SFB_NBINS) { s_cls[threadIdx.x] = 0; s_nnz[threadIdx.x] = 0; }
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long sv = slot[i];
        if (sv) {
            const uint32_t n = (uint32_t)((sv >> 20) & 1023);
            const int b = fin_bin(n);
            atomicAdd(&s_cls[b], 1ULL); atomicAdd(&s_nnz[b], (unsigned long long)n);
        }
    }
    __syncthreads();
    if (threadIdx.x <= SFB_NBINS && s_cls[threadIdx.x]) {
        atomicAdd(fin + FIN_CLS + threadIdx.x, s_cls[threadIdx.x]);
        atomicAdd(fin + FIN_NNZ + threadIdx.x, s_nnz[threadIdx.x]);
    }
}

struct FinParams {
    uint64_t cls_start[SFB_NBINS + 1], nnz_start[SFB_NBINS + 1];
    uint64_t Em;
    const unsigned long long* slot; const unsigned long long* count; const uint32_t* arena; uint64_t n_slots;
    unsigned long long* fin;
    uint32_t* start; uint32_t* len; uint32_t* lab; double* cnt; unsigned long long* cnt_all; double* single; uint8_t* active;
    uint32_t* sgl_cls; uint32_t* sgl_tid;
};

// Every occupied slot claims a class position (and label space) inside its bin, then copies itself.  One 64-bit cursor per bin hands
// out the class position (high half) and the label space (low half) together, so class order and label order agree and a CTA's slice
// of classes owns one contiguous slice of labels.  A CTA walks a contiguous range of slots twice: first it counts what its range
// holds per bin and reserves that with ONE global atomic per bin, then its threads draw their positions from shared-memory cursors
// (one global atomic per class -- 4e5 returning atomics on seven addresses -- made this kernel 0.28 ms at cfg2).
// The order inside a bin is whatever the atomics produce; the E-step/M-step sums are order-free anyway.
__global__ void __launch_bounds__(256) k_eq_fill(const FinParams p) {
    constexpr int NB = SFB_NBINS + 1;                                  // the bins of multi-member classes + the single-member classes
    __shared__ unsigned long long s_tot[NB], s_base[NB], s_cur[NB];
    const uint64_t chunk = (p.n_slots + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = blockIdx.x * chunk, hi = lo + chunk < p.n_slots ? lo + chunk : p.n_slots;
    if (threadIdx.x < NB) { s_tot[threadIdx.x] = 0ULL; s_cur[threadIdx.x] = 0ULL; }
    __syncthreads();
    unsigned long long loc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) loc[b] = 0ULL;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned long long sv = p.slot[i];
        if (!sv) continue;
        const uint32_t n = (uint32_t)((sv >> 20) & 1023);
        const int bn = fin_bin(n);
#pragma unroll
        for (int b = 0; b < NB; ++b) if (b == bn) loc[b] += (1ULL << 32) | (unsigned long long)n;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        unsigned long long v = loc[b];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&s_tot[b], v);
    }
    __syncthreads();
    if (threadIdx.x < NB && s_tot[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(p.fin + FIN_CUR_CLS + threadIdx.x, s_tot[threadIdx.x]);
    __syncthreads();
    unsigned long long total = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned long long sv = p.slot[i];
        if (!sv) continue;
        const uint32_t n = (uint32_t)((sv >> 20) & 1023);
        const uint32_t* a = p.arena + (sv >> 30);
        const unsigned long long cn = p.count[i];
        total += cn;
        const int b = fin_bin(n);
        const unsigned long long tk = s_base[b] + atomicAdd(&s_cur[b], (1ULL << 32) | (unsigned long long)n);
        const uint64_t pos = tk >> 32;
        if (b == SFB_NBINS) {
            const uint32_t t = a[0];
            p.sgl_tid[pos] = t; p.sgl_cls[pos] = (uint32_t)(p.Em + pos);
            p.cnt_all[p.Em + pos] = cn;
            atomicAdd(p.single + t, (double)cn);
            p.active[t] = 1;
        } else {
            const uint64_t c = p.cls_start[b] + pos;
            const uint64_t o = p.nnz_start[b] + (tk & 0xFFFFFFFFULL);
            p.start[c] = (uint32_t)o; p.len[c] = n; p.cnt[c] = (double)cn; p.cnt_all[c] = cn;
            for (uint32_t j = 0; j < n; ++j) { const uint32_t t = a[j]; p.lab[o + j] = t; p.active[t] = 1; }
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
    if ((threadIdx.x & 31u) == 0 && total) atomicAdd(p.fin + FIN_TOTAL, total);
}

__global__ void k_count_active(const uint8_t* __restrict__ active, uint32_t T, unsigned long long* __restrict__ out) {
    unsigned long long n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < T; i += (uint64_t)gridDim.x * blockDim.x) n += active[i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) n += __shfl_xor_sync(0xffffffffu, n, m);
    if ((threadIdx.x & 31u) == 0 && n) atomicAdd(out, n);
}

// growth: every class of the old table moves into a larger one (same arena, same slot word: the label does not move)
__global__ void k_eq_rehash(const unsigned long long* __restrict__ old_slot, const unsigned long long* __restrict__ old_count,
                            uint64_t n_old, const EqTable nt) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_old; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long sv = old_slot[i];
        if (!sv) continue;
        const uint32_t len = (uint32_t)((sv >> 20) & 1023);
        const uint32_t* a = nt.arena + (sv >> 30);
        const uint64_t h = xxh64_words([&](uint32_t j) { return a[j]; }, len, 0);
        const uint64_t bmask = nt.n_buckets - 1;
        const uint64_t b1 = h & bmask, b2 = (b1 ^ (((h >> 48) + 1) * 0x5bd1e995ULL)) & bmask;
        const uint64_t n_main = nt.n_buckets * 4, ov0 = h >> 20;
        bool placed = false;
        for (uint64_t step = 0; step < 8 + nt.n_overflow && !placed; ++step) {
            const uint64_t idx = step < 4 ? b1 * 4 + step : step < 8 ? b2 * 4 + (step - 4) : n_main + ((ov0 + (step - 8)) & (nt.n_overflow - 1));
            if (nt.slot[idx] == 0ULL && atomicCAS(nt.slot + idx, 0ULL, sv) == 0ULL) { nt.count[idx] = old_count[i]; placed = true; }
        }
        if (!placed) atomicOr(nt.cursor + 2, ERR_TABLE_FULL);             // cannot happen: the new table is twice the old one
    }
}

}  // namespace

// ======================================================================================================================
constexpr size_t FIN_SMEM = (size_t)(MAP_THREADS / 32) * N_REGIONS * FIN_S * 32 * 8;      // finalize: hit lists in shared memory

struct MapState {
    sfb200_map_opts o;
    bool begun = false;
    DevBuf<unsigned long long> slot, count, cursor, counters, next_read, scratch, fin;
    DevBuf<uint32_t> heavy;            // reads set aside by the main finalize pass
    DevBuf<unsigned long long> ivpool; // extension words of big seed buckets (IvPool)
    bool defer_heavy = true;           // SFB200_NO_HEAVY_PASS=1: one pass (A/B)
    uint64_t h2d_bytes = 0;            // host batches since map_begin: bytes sent to the device
    DevBuf<uint32_t> arena;
    DevBuf<unsigned int> fld_hist;
    DevBuf<int> remaining;
    DevBuf<int16_t> fld_val, fld_samples;
    DevBuf<uint64_t> pk, pkn;          // packed batch
    DevBuf<uint32_t> meta;
    DevBuf<unsigned long long> iv;     // seed intervals, scan -> finalize
    DevBuf<uint8_t> niv;
    DevBuf<uint32_t> ivmask;
    DevBuf<unsigned int> maxlen;
    DevBuf<unsigned long long> clipped;   // mates cut to MAX_READ_LEN since map_begin
    DevBuf<uint32_t> retry[2];            // reads of the last chunk whose class upsert found the table / arena full
    MapParams last_p;                     // the last chunk's launch parameters (its hand-over buffers stay valid until the next pack)
    bool have_last = false;
    uint32_t n_grown = 0;                 // how often the class table / arena grew since map_begin
    int grid_scan = 0;
    // multi-rank class merge / FLD gather buffers (grow-only: cudaMalloc / cudaFree per step cost more than the exchange)
    DevBuf<unsigned long long> mg_sizes, mg_cnt, mg_cnt_g;
    DevBuf<uint32_t> mg_start, mg_len, mg_lab, mg_start_g, mg_len_g, mg_lab_g;
    DevBuf<unsigned char> fld_send, fld_recv;
    // two staging sets for host batches: the H2D copy of batch j (copy stream) overlaps the mapping kernel of batch j-1
    // staging sets for host batches (N_STAGE = 3: the piece being copied, the piece whose kernels are being enqueued, the piece
    // whose kernels run)
    DevBuf<char> bases1[3], bases2[3];
    DevBuf<uint64_t> off1[3], off2[3];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copied[3] = {nullptr, nullptr, nullptr}, consumed[3] = {nullptr, nullptr, nullptr};
    bool in_use[3] = {false, false, false};
    bool primed = false;               // a host batch has been mapped since map_begin (see sfb200_map_batch)
    // --biasCorrect / --gcBiasCorrect sample collection (sfb200_map_set_bias)
    bool bias_seq = false, bias_gc = false;
    DevBuf<int16_t> bias_val;
    DevBuf<unsigned int> bias_hist;    // [0, 4096) read-start contexts, [4096, 4197) fragment GC percentages; without pseudo-counts
    DevBuf<int> bias_remaining;
    unsigned parity = 0;
    uint64_t n_buckets = 0, n_overflow = 0, arena_words = 0;
    uint64_t n_threads_total = 0;
    int grid = 0;
    std::vector<cudaEvent_t> ev;       // start/stop pairs around the mapping kernels of every batch since map_begin
    size_t ev_used = 0;
    double kernel_ms = 0.0;
};

void sfb_map_state_free(sfb200_ctx* c) {
    MapState* m = c->map;
    if (!m) return;
    m->slot.release(); m->count.release(); m->cursor.release(); m->counters.release(); m->next_read.release();
    m->scratch.release(); m->fin.release(); m->heavy.release(); m->ivpool.release(); m->arena.release(); m->fld_hist.release(); m->remaining.release(); m->fld_val.release(); m->fld_samples.release();
    m->mg_sizes.release(); m->mg_cnt.release(); m->mg_cnt_g.release(); m->mg_start.release(); m->mg_len.release(); m->mg_lab.release();
    m->mg_start_g.release(); m->mg_len_g.release(); m->mg_lab_g.release(); m->fld_send.release(); m->fld_recv.release();
    m->pk.release(); m->pkn.release(); m->meta.release(); m->iv.release(); m->niv.release(); m->ivmask.release(); m->maxlen.release(); m->clipped.release(); m->retry[0].release(); m->retry[1].release();
    m->bias_val.release(); m->bias_hist.release(); m->bias_remaining.release();
    for (int i = 0; i < 3; ++i) {
        m->bases1[i].release(); m->bases2[i].release(); m->off1[i].release(); m->off2[i].release();
        if (m->copied[i]) cudaEventDestroy(m->copied[i]);
        if (m->consumed[i]) cudaEventDestroy(m->consumed[i]);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
    delete m;
    c->map = nullptr;
}

extern "C" int sfb200_map_begin(sfb200_ctx* c, const sfb200_map_opts* o) {
    if (!c || !o) return SFB200_EINVAL;
    if (!c->index.ready) SFB_FAIL(c, SFB200_EINVAL, "map_begin: build the index first");
    if (o->max_frag_len == 0 || o->max_frag_len > 32767) SFB_FAIL(c, SFB200_EINVAL, "max_frag_len must be in [1, 32767]");
    if (o->max_read_occs == 0 || o->max_read_occs > 1000) SFB_FAIL(c, SFB200_EINVAL, "max_read_occs must be in [1, 1000]");
    if (o->max_interval > 16383) SFB_FAIL(c, SFB200_EINVAL, "max_interval must be <= 16383");
    cudaSetDevice(c->device);
    if (!c->map) c->map = new MapState();
    MapState* m = c->map;
    m->o = *o;
    cudaStream_t s = c->stream;
    // table geometry: SFB200_EQ_LOG2_BUCKETS (default 21 -> 8M slots) and SFB200_EQ_ARENA_LOG2 words (default 26)
    { const char* e = getenv("SFB200_NO_HEAVY_PASS"); m->defer_heavy = !(e && atoi(e) != 0); }
    int lb = 21, la = 26;
    if (const char* e = getenv("SFB200_EQ_LOG2_BUCKETS")) lb = std::max(4, std::min(30, atoi(e)));
    if (const char* e = getenv("SFB200_EQ_ARENA_LOG2")) la = std::max(10, std::min(33, atoi(e)));
    m->n_buckets = 1ull << lb; m->n_overflow = std::max<uint64_t>(1024, m->n_buckets / 4); m->arena_words = 1ull << la;
    const uint64_t n_slots = m->n_buckets * 4 + m->n_overflow;
    SFB_CUDA(c, m->slot.reserve(n_slots)); SFB_CUDA(c, m->count.reserve(n_slots)); SFB_CUDA(c, m->arena.reserve(m->arena_words));
    SFB_CUDA(c, m->cursor.reserve(8)); SFB_CUDA(c, m->counters.reserve(6)); SFB_CUDA(c, m->next_read.reserve(8)); SFB_CUDA(c, m->maxlen.reserve(1)); SFB_CUDA(c, m->clipped.reserve(1));
    SFB_CUDA(c, m->fld_hist.reserve(o->max_frag_len)); SFB_CUDA(c, m->remaining.reserve(1));
    SFB_CUDA(c, m->fld_samples.reserve((size_t)std::max(1, o->num_frag_samples)));
    SFB_CUDA(c, cudaMemsetAsync(m->slot.p, 0, n_slots * 8, s));
    SFB_CUDA(c, cudaMemsetAsync(m->count.p, 0, n_slots * 8, s));
    SFB_CUDA(c, cudaMemsetAsync(m->cursor.p, 0, 8 * 8, s));
    SFB_CUDA(c, cudaMemsetAsync(m->counters.p, 0, 6 * 8, s));
    SFB_CUDA(c, cudaMemsetAsync(m->clipped.p, 0, 8, s));
    SFB_CUDA(c, cudaMemsetAsync(m->fld_hist.p, 0, o->max_frag_len * 4ull, s));
    const int rem = o->num_frag_samples;
    SFB_CUDA(c, cudaMemcpyAsync(m->remaining.p, &rem, 4, cudaMemcpyHostToDevice, s));
    // launch geometry: every SM filled with resident CTAs; scratch sized for exactly those threads
    int per_sm = 0;
    // scan: the packed mates of a lane's fragment live in shared memory (paired-end: two mates); the grid is sized for the
    // single-end occupancy, a paired-end launch simply leaves some of those CTAs waiting for a slot
    const size_t smem_pe = (size_t)MAP_THREADS * 2 * 2 * RW * 8;
    SFB_CUDA(c, cudaFuncSetAttribute(k_scan_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pe));
    SFB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_reads, MAP_THREADS, smem_pe / 2));
    m->grid_scan = c->num_sms * std::max(per_sm, 1);
    // finalize: the hit lists' first FIN_S entries per region in shared memory
    SFB_CUDA(c, cudaFuncSetAttribute(k_finalize_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN_SMEM));
    SFB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_finalize_reads, MAP_THREADS, FIN_SMEM));
    if (per_sm < 1) per_sm = 1;
    m->grid = c->num_sms * per_sm;
    m->n_threads_total = (uint64_t)m->grid * MAP_THREADS;
    SFB_CUDA(c, m->scratch.reserve(m->n_threads_total * (uint64_t)N_REGIONS * (o->max_read_occs + 1)));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    // keep (as much as the device allows of) the m-mer bitmap resident in L2 while the table / suffix entries / text stream through it
    if (!getenv("SFB200_NO_L2_PERSIST")) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
        if (max_persist > 0 && max_window > 0) {
            const size_t bytes = c->index.mfilter.bytes();
            const size_t want = std::min<size_t>(bytes, (size_t)max_persist);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue av;
            std::memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = c->index.mfilter.p;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)std::max<size_t>(1, av.accessPolicyWindow.num_bytes));
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
            cudaGetLastError();   // best effort
        }
    }
    c->cls.ready = false;
    m->begun = true;
    m->ev_used = 0; m->kernel_ms = 0.0;
    m->in_use[0] = m->in_use[1] = m->in_use[2] = false; m->parity = 0; m->primed = false; m->h2d_bytes = 0;
    m->bias_seq = m->bias_gc = false;
    m->have_last = false; m->n_grown = 0;
    return SFB200_OK;
}

constexpr uint32_t BIAS_HIST_WORDS = BNK + 101;

extern "C" int sfb200_map_set_bias(sfb200_ctx* c, int seq_bias, int gc_bias, int32_t num_bias_samples) {
    if (!c) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun) SFB_FAIL(c, SFB200_EINVAL, "map_set_bias: call map_begin first");
    if (m->ev_used) SFB_FAIL(c, SFB200_EINVAL, "map_set_bias: call it before the first batch");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    m->bias_seq = seq_bias != 0; m->bias_gc = gc_bias != 0;
    SFB_CUDA(c, m->bias_hist.reserve(BIAS_HIST_WORDS)); SFB_CUDA(c, m->bias_remaining.reserve(1));
    SFB_CUDA(c, cudaMemsetAsync(m->bias_hist.p, 0, BIAS_HIST_WORDS * 4ull, s));
    const int rem = num_bias_samples < 0 ? 0 : num_bias_samples;
    SFB_CUDA(c, cudaMemcpyAsync(m->bias_remaining.p, &rem, 4, cudaMemcpyHostToDevice, s));
    SFB_CUDA(c, cudaFuncSetAttribute(k_finalize_reads_bias, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN_SMEM));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    return SFB200_OK;
}

extern "C" int sfb200_map_get_bias(sfb200_ctx* c, uint32_t* read_bias, uint32_t* observed_gc) {
    if (!c || !read_bias || !observed_gc) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun || !m->bias_hist.p) SFB_FAIL(c, SFB200_EINVAL, "map_get_bias: map_set_bias was not called");
    cudaSetDevice(c->device);
    std::vector<unsigned int> h(BIAS_HIST_WORDS);
    if (c->n_ranks > 1) {
        // every rank collected from its own reads: the model is the sum (the sampling limit then applies per rank)
        std::vector<unsigned long long> w(BIAS_HIST_WORDS);
        SFB_CUDA(c, cudaMemcpyAsync(h.data(), m->bias_hist.p, BIAS_HIST_WORDS * 4ull, cudaMemcpyDeviceToHost, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (uint32_t i = 0; i < BIAS_HIST_WORDS; ++i) w[i] = h[i];
        DevBuf<unsigned long long> d; SFB_CUDA(c, d.reserve(BIAS_HIST_WORDS));
        SFB_CUDA(c, cudaMemcpyAsync(d.p, w.data(), BIAS_HIST_WORDS * 8ull, cudaMemcpyHostToDevice, c->stream));
        const int rc = sfb_comm_allreduce_u64(c, d.p, BIAS_HIST_WORDS);
        if (rc) { d.release(); return rc; }
        SFB_CUDA(c, cudaMemcpyAsync(w.data(), d.p, BIAS_HIST_WORDS * 8ull, cudaMemcpyDeviceToHost, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
        d.release();
        for (uint32_t i = 0; i < BIAS_HIST_WORDS; ++i) h[i] = (unsigned int)w[i];
    } else {
        SFB_CUDA(c, cudaMemcpyAsync(h.data(), m->bias_hist.p, BIAS_HIST_WORDS * 4ull, cudaMemcpyDeviceToHost, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    // both distributions start from a count of one per bin (ReadKmerDist.hpp:20-24, ReadExperiment.hpp:50)
    for (uint32_t i = 0; i < BNK; ++i) read_bias[i] = h[i] + 1u;
    for (uint32_t i = 0; i < 101; ++i) observed_gc[i] = h[BNK + i] + 1u;
    return SFB200_OK;
}

static int map_chunk_device(sfb200_ctx* c, const char* d_bases1, const uint64_t* d_off1, const char* d_bases2,
                            const uint64_t* d_off2, uint64_t n_reads);

// ---- a class table that grows (libcuckoo grows too: include/cuckoohash_map.hh, cuckoo_expand_simple) --------------------------------
// The table starts at the geometry of map_begin.  When an upsert finds both buckets and the overflow region full, or the label arena
// exhausted, the read is put on a retry list instead of being lost; before the next chunk is packed (and in map_finish) the host looks
// at the error flags -- it synchronises there anyway -- doubles what was full (classes re-inserted by k_eq_rehash; labels stay where
// they are, the arena is copied) and finalizes the listed reads again, upsert only.
static int eq_resize(sfb200_ctx* c, MapState* m, uint64_t new_buckets, uint64_t new_arena_words) {
    cudaStream_t s = c->stream;
    if (new_arena_words > m->arena_words) {
        if (new_arena_words > (1ull << 34)) SFB_FAIL(c, SFB200_EFULL, "equivalence-class label arena cannot grow beyond 2^34 words");
        unsigned long long used = 0;
        SFB_CUDA(c, cudaMemcpyAsync(&used, m->cursor.p, 8, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        used = std::min<unsigned long long>(used, m->arena_words);       // reservations beyond the end were refused
        DevBuf<uint32_t> na;
        SFB_CUDA(c, na.reserve(new_arena_words));
        SFB_CUDA(c, cudaMemcpyAsync(na.p, m->arena.p, used * 4, cudaMemcpyDeviceToDevice, s));
        SFB_CUDA(c, cudaMemcpyAsync(m->cursor.p, &used, 8, cudaMemcpyHostToDevice, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        m->arena.release(); m->arena = na; m->arena_words = new_arena_words;
    }
    if (new_buckets > m->n_buckets) {
        const uint64_t n_old = m->n_buckets * 4 + m->n_overflow;
        const uint64_t n_over = std::max<uint64_t>(1024, new_buckets / 4), n_new = new_buckets * 4 + n_over;
        DevBuf<unsigned long long> ns, nc;
        SFB_CUDA(c, ns.reserve(n_new)); SFB_CUDA(c, nc.reserve(n_new));
        SFB_CUDA(c, cudaMemsetAsync(ns.p, 0, n_new * 8, s)); SFB_CUDA(c, cudaMemsetAsync(nc.p, 0, n_new * 8, s));
        EqTable nt;
        nt.slot = ns.p; nt.count = nc.p; nt.arena = m->arena.p; nt.cursor = m->cursor.p;
        nt.n_buckets = new_buckets; nt.n_overflow = n_over; nt.arena_words = m->arena_words;
        k_eq_rehash<<<(unsigned)std::min<uint64_t>((n_old + 255) / 256, (uint64_t)c->num_sms * 16), 256, 0, s>>>(m->slot.p, m->count.p, n_old, nt);
        c->launches++;
        SFB_CUDA(c, cudaGetLastError());
        SFB_CUDA(c, cudaStreamSynchronize(s));
        m->slot.release(); m->count.release();
        m->slot = ns; m->count = nc; m->n_buckets = new_buckets; m->n_overflow = n_over;
    }
    m->n_grown++;
    return SFB200_OK;
}

// called with the stream idle: grows what the last chunk found full and finalizes its listed reads again
static int eq_grow_and_retry(sfb200_ctx* c, MapState* m) {
    cudaStream_t s = c->stream;
    for (int round = 0; round < 48; ++round) {
        unsigned long long cur[8];
        SFB_CUDA(c, cudaMemcpyAsync(cur, m->cursor.p, sizeof(cur), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        const unsigned long long flags = cur[2] & (ERR_ARENA_FULL | ERR_TABLE_FULL), n_retry = cur[4];
        if (!flags && !n_retry) return SFB200_OK;
        if (!m->have_last) SFB_FAIL(c, SFB200_EFULL, "equivalence-class table full with nothing to replay");
        { const int rc = eq_resize(c, m, (flags & ERR_TABLE_FULL) ? m->n_buckets * 2 : m->n_buckets,
                                   (flags & ERR_ARENA_FULL) ? m->arena_words * 2 : m->arena_words); if (rc) return rc; }
        const unsigned long long keep = cur[2] & ~(ERR_ARENA_FULL | ERR_TABLE_FULL), zero = 0;
        SFB_CUDA(c, cudaMemcpyAsync(m->cursor.p + 2, &keep, 8, cudaMemcpyHostToDevice, s));
        SFB_CUDA(c, cudaMemcpyAsync(m->cursor.p + 4, &zero, 8, cudaMemcpyHostToDevice, s));
        SFB_CUDA(c, cudaMemsetAsync(m->next_read.p, 0, 64, s));
        MapParams p = m->last_p;
        p.tb.slot = m->slot.p; p.tb.count = m->count.p; p.tb.arena = m->arena.p;
        p.tb.n_buckets = m->n_buckets; p.tb.n_overflow = m->n_overflow; p.tb.arena_words = m->arena_words;
        p.retry_in = m->retry[round & 1].p; p.retry_out = m->retry[(round & 1) ^ 1].p; p.n_retry = n_retry;
        p.heavy_out = nullptr; p.heavy_in = nullptr;
        p.fld_val = nullptr; p.bias_val = nullptr; p.bias_seq = 0; p.bias_gc = 0;
        if (n_retry) {
            k_finalize_reads<<<(unsigned)std::min<uint64_t>(m->grid, (n_retry + MAP_THREADS - 1) / MAP_THREADS), MAP_THREADS, FIN_SMEM, s>>>(p);
            c->launches++;
            SFB_CUDA(c, cudaGetLastError());
        }
    }
    SFB_FAIL(c, SFB200_EFULL, "equivalence-class table still full after 48 doublings");
}

// Batches of any size: the per-fragment hand-over buffers (packed reads, seed intervals) are sized for at most
// MAX_CHUNK fragments, larger batches are walked chunk by chunk (offsets are absolute, so a chunk is a pointer shift).
extern "C" int sfb200_map_batch_device(sfb200_ctx* c, const char* d_bases1, const uint64_t* d_off1, const char* d_bases2,
                                       const uint64_t* d_off2, uint64_t n_reads) {
    if (!c) return SFB200_EINVAL;
    uint64_t MAX_CHUNK = 4u << 20;
    if (const char* e = getenv("SFB200_MAX_CHUNK")) MAX_CHUNK = std::min<long long>(1ll << 30, std::max<long long>(32, atoll(e)));   // the scan kernel indexes a chunk with 32 bits
    for (uint64_t a = 0; a < n_reads || a == 0; a += MAX_CHUNK) {
        const uint64_t n = std::min<uint64_t>(MAX_CHUNK, n_reads - a);
        const int rc = map_chunk_device(c, d_bases1, d_off1 ? d_off1 + a : nullptr, d_bases2, d_off2 ? d_off2 + a : nullptr, n);
        if (rc || n_reads == 0) return rc;
    }
    return SFB200_OK;
}

static int map_chunk_device(sfb200_ctx* c, const char* d_bases1, const uint64_t* d_off1, const char* d_bases2,
                            const uint64_t* d_off2, uint64_t n_reads) {
    if (!c) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun) SFB_FAIL(c, SFB200_EINVAL, "map_batch: call map_begin first");
    if (n_reads == 0) return SFB200_OK;
    if (!d_bases1 || !d_off1 || ((d_bases2 == nullptr) != (d_off2 == nullptr))) SFB_FAIL(c, SFB200_EINVAL, "map_batch: null array");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    const DevIndex& ix = c->index;
    MapParams p;
    std::memset(&p, 0, sizeof(p));
    p.ix.words = ix.words.p; p.ix.txp_start = ix.txp_start.p; p.ix.txp_end = ix.txp_end.p; p.ix.sa = ix.sa.p;
    p.ix.table = ix.table.p; p.ix.mfilter = ix.mfilter.p; p.ix.mf_m = ix.mf_m; p.ix.mask = ix.table_slots - 1; p.ix.k = ix.k;
    p.ix.kmask = (1ULL << (2 * ix.k)) - 1;
    p.tb.slot = m->slot.p; p.tb.count = m->count.p; p.tb.arena = m->arena.p; p.tb.cursor = m->cursor.p;
    p.tb.n_buckets = m->n_buckets; p.tb.n_overflow = m->n_overflow; p.tb.arena_words = m->arena_words;
    p.bases1 = d_bases1; p.off1 = d_off1; p.bases2 = d_bases2; p.off2 = d_off2; p.n_reads = n_reads;
    p.cap = m->o.max_read_occs; p.max_frag_len = m->o.max_frag_len; p.max_interval = m->o.max_interval;
    p.lib_fmt = m->o.lib_format_id; p.strict_intersect = m->o.strict_intersect; p.allow_orphans = m->o.allow_orphans;
    p.allow_dovetail = m->o.allow_dovetail; p.ignore_compat = m->o.ignore_compat; p.enforce_compat = m->o.enforce_compat;
    p.scratch = m->scratch.p; p.n_threads_total = m->n_threads_total; p.counters = m->counters.p; p.next_read = m->next_read.p;
    const bool want_fld = d_bases2 != nullptr;
    if (want_fld) { SFB_CUDA(c, m->fld_val.reserve(n_reads)); p.fld_val = m->fld_val.p; }
    SFB_CUDA(c, cudaMemsetAsync(m->next_read.p, 0, 64, s));
    const int n_mates = d_bases2 ? 2 : 1;
    if (m->ev.size() < m->ev_used + 2) { cudaEvent_t a, b; SFB_CUDA(c, cudaEventCreate(&a)); SFB_CUDA(c, cudaEventCreate(&b)); m->ev.push_back(a); m->ev.push_back(b); }
    SFB_CUDA(c, cudaEventRecord(m->ev[m->ev_used], s));
    // 1. longest read of the batch -> words per packed read
    SFB_CUDA(c, cudaMemsetAsync(m->maxlen.p, 0, 4, s));
    k_max_read_len<<<(unsigned)std::min<uint64_t>((n_reads + 255) / 256, 1024), 256, 0, s>>>(d_off1, d_off2, n_reads, m->maxlen.p, m->clipped.p);
    c->launches++;
    unsigned int maxlen = 0;
    unsigned long long h_cur[8];
    SFB_CUDA(c, cudaMemcpyAsync(&maxlen, m->maxlen.p, 4, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaMemcpyAsync(h_cur, m->cursor.p, sizeof(h_cur), cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    // the previous chunk is finished: if its upserts found the table or the arena full, grow and replay them while its hand-over
    // buffers are still intact (the pack below overwrites them)
    if ((h_cur[2] & (ERR_ARENA_FULL | ERR_TABLE_FULL)) || h_cur[4]) {
        const int rc = eq_grow_and_retry(c, m);
        if (rc) return rc;
        SFB_CUDA(c, cudaMemsetAsync(m->next_read.p, 0, 64, s));
        p.tb.slot = m->slot.p; p.tb.count = m->count.p; p.tb.arena = m->arena.p;
        p.tb.n_buckets = m->n_buckets; p.tb.n_overflow = m->n_overflow; p.tb.arena_words = m->arena_words;
    }
    SFB_CUDA(c, m->retry[0].reserve(n_reads)); SFB_CUDA(c, m->retry[1].reserve(n_reads));
    p.retry_out = m->retry[0].p; p.retry_in = nullptr; p.n_retry = 0;
    const uint32_t rwp = std::max<uint32_t>(1, (maxlen + 31) / 32);
    // 2. pack
    const uint64_t n_fm = n_reads * n_mates;
    SFB_CUDA(c, m->pk.reserve(n_fm * rwp)); SFB_CUDA(c, m->pkn.reserve(n_fm * rwp)); SFB_CUDA(c, m->meta.reserve(n_fm));
    SFB_CUDA(c, m->iv.reserve(n_reads * (uint64_t)(2 * n_mates) * MAX_IV)); SFB_CUDA(c, m->niv.reserve(n_reads * (uint64_t)(2 * n_mates)));
    SFB_CUDA(c, m->ivmask.reserve(n_reads * (uint64_t)(2 * n_mates) * MAX_IV));
    SFB_CUDA(c, cudaMemsetAsync(m->meta.p, 0, n_fm * 4, s));
    k_pack_reads<<<(unsigned)((n_fm * rwp + 255) / 256), 256, 0, s>>>(d_bases1, d_off1, d_bases2, d_off2, n_reads, n_mates, rwp, m->pk.p, m->pkn.p, m->meta.p);
    c->launches++;
    p.pk = m->pk.p; p.pkn = m->pkn.p; p.meta = m->meta.p; p.rwp = rwp; p.n_mates = n_mates; p.iv = m->iv.p; p.niv = m->niv.p; p.ivmask = m->ivmask.p;
    // per chunk: the pool of per-round extension words of
    // big seed buckets (cursor[6]; 32 bytes per read -- a bucket of 1000 positions takes 32 words; when the pool runs out the
    // finalize kernel extends those entries itself)
    uint64_t pool_cap = std::max<uint64_t>(1u << 16, 4 * n_reads);
    if (const char* e = getenv("SFB200_IVPOOL_WORDS")) pool_cap = (uint64_t)std::max<long long>(1, atoll(e));     // tests: a pool that runs out
    SFB_CUDA(c, m->ivpool.reserve(pool_cap));
    SFB_CUDA(c, cudaMemsetAsync(m->cursor.p + 5, 0, 16, s));
    p.pool.words = m->ivpool.p; p.pool.cursor = m->cursor.p + 6; p.pool.cap = pool_cap;
    // 3. seed scan (lanes pull fragments), 4. finalize (projection .. class upsert)
    const size_t smem = (size_t)MAP_THREADS * n_mates * 2 * RW * 8;
    const uint64_t blocks_needed = (n_reads + MAP_THREADS - 1) / MAP_THREADS;
    k_scan_reads<<<(unsigned)std::min<uint64_t>(m->grid_scan, blocks_needed), MAP_THREADS, smem, s>>>(p);
    c->launches++;
    const bool bias = m->bias_seq || m->bias_gc;
    if (bias) {
        if (m->bias_seq) { SFB_CUDA(c, m->bias_val.reserve(n_reads)); p.bias_val = m->bias_val.p; }
        p.gc_hist = m->bias_hist.p + BNK; p.bias_seq = m->bias_seq ? 1 : 0; p.bias_gc = (m->bias_gc && n_mates == 2) ? 1 : 0;
    }
    // main pass, then the reads it set aside (big seed buckets); the second launch reads their number on the device
    const bool defer = m->defer_heavy;
    if (defer) {
        SFB_CUDA(c, m->heavy.reserve(3 * n_reads));
        p.heavy_out = m->heavy.p;
    }
    for (int pass = 0; pass < (defer ? 2 : 1); ++pass) {
        if (pass == 1) {
            SFB_CUDA(c, cudaMemsetAsync(m->next_read.p + 1, 0, 8, s));
            p.heavy_out = nullptr; p.heavy_in = m->heavy.p;
        }
        if (!bias) k_finalize_reads<<<(unsigned)std::min<uint64_t>(m->grid, blocks_needed), MAP_THREADS, FIN_SMEM, s>>>(p);
        else k_finalize_reads_bias<<<(unsigned)std::min<uint64_t>(m->grid, blocks_needed), MAP_THREADS, FIN_SMEM, s>>>(p);
        c->launches++;
        SFB_CUDA(c, cudaGetLastError());
    }
    p.heavy_out = nullptr; p.heavy_in = nullptr;                       // a replay after table growth (eq_grow_and_retry) lists its reads itself
    m->last_p = p; m->have_last = true;
    SFB_CUDA(c, cudaEventRecord(m->ev[m->ev_used + 1], s));
    m->ev_used += 2;
    if (want_fld) {
        k_fld_select<<<1, 1024, 0, s>>>(m->fld_val.p, n_reads, m->fld_hist.p, m->remaining.p, m->fld_samples.p, m->o.num_frag_samples);
        c->launches++;
        SFB_CUDA(c, cudaGetLastError());
    }
    if (m->bias_seq) {
        k_bias_select<<<1, 1024, 0, s>>>(m->bias_val.p, n_reads, m->bias_hist.p, m->bias_remaining.p);
        c->launches++;
        SFB_CUDA(c, cudaGetLastError());
    }
    return SFB200_OK;
}

// Host batches travel in pieces: piece j+1 is on the copy stream while the kernels of piece j are enqueued (a host round trip: the
// chunk's longest read comes back before the pack kernel is sized) and the kernels of piece j-1 run -- three staging sets.  The copy
// engine never waits for the host, so a batch costs max(copy, kernels) plus the kernels of its LAST piece; pieces are therefore small
// (SFB200_HOST_PIECE reads, default 512 k: 0.5 ms of kernels; measured with fixed-length reads 19.6 ms per 10 M reads against 20.1 with
// 1 M pieces and 21.0 with whole 2.5 M batches, profiles/r02r_e2e_ab.txt), and the first ones after map_begin smaller still (nothing to hide
// behind yet).
constexpr unsigned N_STAGE = 3;
struct HostPiece { uint64_t at, n; unsigned set; };
// a host batch: offsets arrays, or (o == nullptr) reads of one length L stored back to back
struct HostBatch {
    const char* b1; const uint64_t* o1; uint64_t L1;
    const char* b2; const uint64_t* o2; uint64_t L2;
    uint64_t n;
    uint64_t at1(uint64_t i) const { return o1 ? o1[i] : i * L1; }
    uint64_t at2(uint64_t i) const { return o2 ? o2[i] : i * L2; }
};

// Reads of one length (the usual sequencer output, sfb200_map_batch_fixed): the piece's offsets are an arithmetic progression, written
// on the device instead of copied -- 8 of every read's ~84 bytes on a path that is bound by the host link (with eight ranks on one
// host the link is shared: profiles/r02o_scaling.json)
__global__ void k_iota_offsets(uint64_t* __restrict__ out, uint64_t first, uint64_t step, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = first + i * step;
}

static int copy_offsets(sfb200_ctx* c, MapState* m, uint64_t* dst, const uint64_t* o, uint64_t L, uint64_t at, uint64_t n, cudaStream_t cs) {
    if (!o) {
        k_iota_offsets<<<(unsigned)std::min<uint64_t>((n + 256) / 256, 1024), 256, 0, cs>>>(dst, at * L, L, n + 1);
        c->launches++;
        SFB_CUDA(c, cudaGetLastError());
    } else {
        SFB_CUDA(c, cudaMemcpyAsync(dst, o + at, (n + 1) * 8, cudaMemcpyHostToDevice, cs));
        m->h2d_bytes += (n + 1) * 8;
    }
    return SFB200_OK;
}

static int host_piece_copy(sfb200_ctx* c, const HostBatch& hb, HostPiece& pc) {
    MapState* m = c->map;
    const unsigned b = pc.set = m->parity;
    m->parity = (m->parity + 1u) % N_STAGE;
    // this staging set was last read by the kernels of three pieces ago
    if (m->in_use[b]) SFB_CUDA(c, cudaEventSynchronize(m->consumed[b]));
    cudaStream_t cs = m->copy_stream;
    const uint64_t f1 = hb.at1(pc.at), nb1 = hb.at1(pc.at + pc.n) - f1;
    SFB_CUDA(c, m->bases1[b].reserve(nb1 + 8)); SFB_CUDA(c, m->off1[b].reserve(pc.n + 1));
    SFB_CUDA(c, cudaMemcpyAsync(m->bases1[b].p, hb.b1 + f1, nb1, cudaMemcpyHostToDevice, cs));
    { const int rc = copy_offsets(c, m, m->off1[b].p, hb.o1, hb.L1, pc.at, pc.n, cs); if (rc) return rc; }
    m->h2d_bytes += nb1;
    if (hb.b2) {
        const uint64_t f2 = hb.at2(pc.at), nb2 = hb.at2(pc.at + pc.n) - f2;
        SFB_CUDA(c, m->bases2[b].reserve(nb2 + 8)); SFB_CUDA(c, m->off2[b].reserve(pc.n + 1));
        SFB_CUDA(c, cudaMemcpyAsync(m->bases2[b].p, hb.b2 + f2, nb2, cudaMemcpyHostToDevice, cs));
        { const int rc = copy_offsets(c, m, m->off2[b].p, hb.o2, hb.L2, pc.at, pc.n, cs); if (rc) return rc; }
        m->h2d_bytes += nb2;
    }
    SFB_CUDA(c, cudaEventRecord(m->copied[b], cs));
    return SFB200_OK;
}

static int host_piece_map(sfb200_ctx* c, const HostBatch& hb, const HostPiece& pc) {
    MapState* m = c->map;
    const unsigned b = pc.set;
    SFB_CUDA(c, cudaStreamWaitEvent(c->stream, m->copied[b], 0));
    // offsets are absolute positions in the caller's arrays: shift the base pointers instead of the offsets
    const char* d_b2 = hb.b2 ? m->bases2[b].p - hb.at2(pc.at) : nullptr;
    const int rc = sfb200_map_batch_device(c, m->bases1[b].p - hb.at1(pc.at), m->off1[b].p, d_b2, hb.b2 ? m->off2[b].p : nullptr, pc.n);
    if (rc) return rc;
    SFB_CUDA(c, cudaEventRecord(m->consumed[b], c->stream));
    m->in_use[b] = true;
    return SFB200_OK;
}

static int map_host_batch(sfb200_ctx* c, const HostBatch& hb) {
    MapState* m = c->map;
    const uint64_t n_reads = hb.n;
    cudaSetDevice(c->device);
    if (!m->copy_stream) {
        SFB_CUDA(c, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (unsigned i = 0; i < N_STAGE; ++i) {
            SFB_CUDA(c, cudaEventCreateWithFlags(&m->copied[i], cudaEventDisableTiming));
            SFB_CUDA(c, cudaEventCreateWithFlags(&m->consumed[i], cudaEventDisableTiming));
        }
    }
    uint64_t piece = 512u << 10, ramp = 128u << 10;
    if (const char* e = getenv("SFB200_HOST_PIECE")) piece = (uint64_t)std::max<long long>(1024, atoll(e));
    if (const char* e = getenv("SFB200_MAP_RAMP")) ramp = (uint64_t)std::max<long long>(0, atoll(e));
    if (m->primed || ramp == 0 || ramp > piece) ramp = piece;        // the first pieces after map_begin: 128 k, 256 k, ... reads
    m->primed = true;
    uint64_t at = 0;
    auto next_piece = [&](HostPiece& pc) {
        pc.at = at;
        pc.n = std::min<uint64_t>(ramp, n_reads - at);
        if (n_reads - at - pc.n < pc.n / 2) pc.n = n_reads - at;        // no sliver at the end
        at += pc.n;
        ramp = std::min<uint64_t>(piece, ramp * 2);
    };
    HostPiece cur, nxt;
    next_piece(cur);
    { const int rc = host_piece_copy(c, hb, cur); if (rc) return rc; }
    unsigned last_set = cur.set;
    for (;;) {
        const bool more = at < n_reads;
        if (more) {
            next_piece(nxt);
            const int rc = host_piece_copy(c, hb, nxt);
            if (rc) return rc;
            last_set = nxt.set;
        }
        const int rc = host_piece_map(c, hb, cur);
        if (rc) return rc;
        if (!more) break;
        cur = nxt;
    }
    // the caller may reuse its buffers as soon as we return: wait for the last copy (not for the kernels)
    SFB_CUDA(c, cudaEventSynchronize(m->copied[last_set]));
    return SFB200_OK;
}

extern "C" int sfb200_map_batch(sfb200_ctx* c, const char* bases1, const uint64_t* off1, const char* bases2,
                                const uint64_t* off2, uint64_t n_reads) {
    if (!c) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun) SFB_FAIL(c, SFB200_EINVAL, "map_batch: call map_begin first");
    if (n_reads == 0) return SFB200_OK;
    if (!bases1 || !off1 || ((bases2 == nullptr) != (off2 == nullptr))) SFB_FAIL(c, SFB200_EINVAL, "map_batch: null array");
    const HostBatch hb{bases1, off1, 0, bases2, off2, 0, n_reads};
    return map_host_batch(c, hb);
}

extern "C" int sfb200_map_batch_fixed(sfb200_ctx* c, const char* bases1, uint32_t len1, const char* bases2, uint32_t len2, uint64_t n_reads) {
    if (!c) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun) SFB_FAIL(c, SFB200_EINVAL, "map_batch_fixed: call map_begin first");
    if (n_reads == 0) return SFB200_OK;
    if (!bases1) SFB_FAIL(c, SFB200_EINVAL, "map_batch_fixed: null array");
    const HostBatch hb{bases1, nullptr, len1, bases2, nullptr, bases2 ? len2 : 0, n_reads};
    return map_host_batch(c, hb);
}

extern "C" uint64_t sfb200_map_h2d_bytes(const sfb200_ctx* c) { return (c && c->map) ? c->map->h2d_bytes : 0; }

extern "C" double sfb200_last_map_kernel_ms(const sfb200_ctx* c) { return (c && c->map) ? c->map->kernel_ms : 0.0; }

extern "C" uint64_t sfb200_map_clipped(sfb200_ctx* c) {
    if (!c || !c->map || !c->map->clipped.p) return 0;
    cudaSetDevice(c->device);
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, c->map->clipped.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 0;
    cudaStreamSynchronize(c->stream);
    return v;
}

// eqBuilder.finish() (EquivalenceClassBuilder.hpp:64-80): flatten the table -- on the device, straight into the binned
// layout the inference kernels read; nothing but a few counters crosses PCIe.
static int flatten_classes(sfb200_ctx* c, MapState* m, uint64_t* n_classes, uint64_t* nnz) {
    cudaStream_t s = c->stream;
    const uint64_t n_slots = m->n_buckets * 4 + m->n_overflow;
    DevClasses& k = c->cls;
    k.ready = false; k.host_valid = false; k.from_device = true; k.export_to_canon.clear(); k.part.valid = false;
    k.h_row_ptr.clear(); k.h_labels.clear(); k.h_counts.clear();
    const uint32_t T = c->index.n_txp;
    SFB_CUDA(c, m->fin.reserve(FIN_WORDS));
    SFB_CUDA(c, cudaMemsetAsync(m->fin.p, 0, FIN_WORDS * 8, s));
    const unsigned fgrid = (unsigned)std::min<uint64_t>((n_slots + 255) / 256, (uint64_t)c->num_sms * 16);
    k_eq_count<<<fgrid, 256, 0, s>>>(m->slot.p, n_slots, m->fin.p);
    c->launches++;
    unsigned long long h_fin[FIN_WORDS];
    SFB_CUDA(c, cudaMemcpyAsync(h_fin, m->fin.p, sizeof(h_fin), cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    uint64_t cls_start[SFB_NBINS + 2], nnz_start[SFB_NBINS + 2];
    cls_start[0] = 0; nnz_start[0] = 0;
    for (int b = 0; b <= SFB_NBINS; ++b) { cls_start[b + 1] = cls_start[b] + h_fin[FIN_CLS + b]; nnz_start[b + 1] = nnz_start[b] + h_fin[FIN_NNZ + b]; }
    const uint64_t Em = cls_start[SFB_NBINS], nnzm = nnz_start[SFB_NBINS];
    const uint64_t n_sgl = h_fin[FIN_CLS + SFB_NBINS];
    const uint64_t E = Em + n_sgl, z = nnzm + n_sgl;
    if (nnzm >= 0xFFFFFFFFull) SFB_FAIL(c, SFB200_EINVAL, "more than 2^32 label entries");
    k.n_txp = T; k.E = E; k.nnz = z; k.Em = Em; k.nnzm = nnzm; k.n_sgl = n_sgl;
    for (int b = 0; b <= SFB_NBINS; ++b) k.bin_cls[b] = cls_start[b];
    SFB_CUDA(c, k.start.reserve(Em)); SFB_CUDA(c, k.len.reserve(Em)); SFB_CUDA(c, k.lab.reserve(nnzm)); SFB_CUDA(c, k.w.reserve(nnzm));
    SFB_CUDA(c, k.cnt.reserve(Em)); SFB_CUDA(c, k.cnt_all.reserve(E)); SFB_CUDA(c, k.single.reserve(T)); SFB_CUDA(c, k.active.reserve(T));
    SFB_CUDA(c, k.sgl_cls.reserve(n_sgl)); SFB_CUDA(c, k.sgl_tid.reserve(n_sgl));
    SFB_CUDA(c, cudaMemsetAsync(k.single.p, 0, T * 8ull, s));
    SFB_CUDA(c, cudaMemsetAsync(k.active.p, 0, T, s));
    FinParams fp;
    for (int b = 0; b <= SFB_NBINS; ++b) { fp.cls_start[b] = cls_start[b]; fp.nnz_start[b] = nnz_start[b]; }
    fp.Em = Em;
    fp.slot = m->slot.p; fp.count = m->count.p; fp.arena = m->arena.p; fp.n_slots = n_slots; fp.fin = m->fin.p;
    fp.start = k.start.p; fp.len = k.len.p; fp.lab = k.lab.p; fp.cnt = k.cnt.p; fp.cnt_all = k.cnt_all.p; fp.single = k.single.p;
    fp.active = k.active.p; fp.sgl_cls = k.sgl_cls.p; fp.sgl_tid = k.sgl_tid.p;
    k_eq_fill<<<fgrid, 256, 0, s>>>(fp);
    c->launches++;
    k_count_active<<<(unsigned)std::min<uint64_t>((T + 255) / 256, 1024), 256, 0, s>>>(k.active.p, T, m->fin.p + FIN_ACTIVE);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    SFB_CUDA(c, cudaMemcpyAsync(h_fin, m->fin.p, sizeof(h_fin), cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    k.n_active = h_fin[FIN_ACTIVE];
    k.total_count = h_fin[FIN_TOTAL];
    if (n_classes) *n_classes = E;
    if (nnz) *nnz = z;
    k.ready = true;
    return SFB200_OK;
}

namespace {
// singles appended behind the multi-member classes so that one (start, len, count, label) quadruple describes every class
__global__ void k_merge_pack_singles(const uint32_t* __restrict__ sgl_tid, uint64_t n_sgl, uint64_t Em, uint64_t nnzm,
                                     uint32_t* __restrict__ start_all, uint32_t* __restrict__ len_all, uint32_t* __restrict__ lab_all) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_sgl) return;
    start_all[Em + i] = (uint32_t)(nnzm + i); len_all[Em + i] = 1; lab_all[nnzm + i] = sgl_tid[i];
}
__global__ void k_merge_upsert(const EqTable tb, const uint32_t* __restrict__ start_g, const uint32_t* __restrict__ len_g,
                               const unsigned long long* __restrict__ cnt_g, const uint32_t* __restrict__ lab_g,
                               const unsigned long long* __restrict__ sizes /* per rank: E, nnz */, uint64_t maxE, uint64_t maxZ,
                               int my_rank) {
    const int r = blockIdx.y;
    if (r == my_rank) return;
    const uint64_t E = sizes[2 * r];
    const uint32_t* st = start_g + (size_t)r * maxE; const uint32_t* ln = len_g + (size_t)r * maxE;
    const unsigned long long* cn = cnt_g + (size_t)r * maxE; const uint32_t* lb = lab_g + (size_t)r * maxZ;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < E; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t* a = lb + st[i];
        auto get = [&](uint32_t j) { return a[j]; };
        eq_upsert(tb, ln[i], get, cn[i], xxh64_words(get, ln[i], 0));
    }
}
}  // namespace

static int merge_classes_over_ranks(sfb200_ctx* c, MapState* m) {
    cudaStream_t s = c->stream;
    DevClasses& k = c->cls;
    const int R = c->n_ranks;
    const uint64_t E = k.E, Z = k.nnz, Em = k.Em, nnzm = k.nnzm;
    DevBuf<unsigned long long>& d_sizes = m->mg_sizes; DevBuf<unsigned long long>& d_cnt_g = m->mg_cnt_g; DevBuf<unsigned long long>& d_cnt = m->mg_cnt;
    DevBuf<uint32_t>& d_start = m->mg_start; DevBuf<uint32_t>& d_len = m->mg_len; DevBuf<uint32_t>& d_lab = m->mg_lab;
    DevBuf<uint32_t>& d_start_g = m->mg_start_g; DevBuf<uint32_t>& d_len_g = m->mg_len_g; DevBuf<uint32_t>& d_lab_g = m->mg_lab_g;
    auto cleanup = [&]() {};
#define MG(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); c->err = std::string(#call) + ": " + cudaGetErrorString(e__); return SFB200_ECUDA; } } while (0)
#define MGRC(call) do { const int r__ = (call); if (r__) { cleanup(); return r__; } } while (0)
    MG(d_sizes.reserve(2 * (size_t)R + 2));
    unsigned long long mine[2] = {E, Z};
    MG(cudaMemcpyAsync(d_sizes.p + 2 * (size_t)R, mine, 16, cudaMemcpyHostToDevice, s));
    MGRC(sfb_comm_allgather(c, d_sizes.p + 2 * (size_t)R, d_sizes.p, 16));
    std::vector<unsigned long long> sizes(2 * (size_t)R);
    MG(cudaMemcpyAsync(sizes.data(), d_sizes.p, sizes.size() * 8, cudaMemcpyDeviceToHost, s));
    MG(cudaStreamSynchronize(s));
    uint64_t maxE = 1, maxZ = 1;
    for (int r = 0; r < R; ++r) { maxE = std::max<uint64_t>(maxE, sizes[2 * r]); maxZ = std::max<uint64_t>(maxZ, sizes[2 * r + 1]); }
    maxE = (maxE + 3) & ~3ull; maxZ = (maxZ + 3) & ~3ull;
    MG(d_start.reserve(maxE)); MG(d_len.reserve(maxE)); MG(d_lab.reserve(maxZ));
    MG(d_start_g.reserve(maxE * R)); MG(d_len_g.reserve(maxE * R)); MG(d_lab_g.reserve(maxZ * R)); MG(d_cnt_g.reserve(maxE * R));
    if (Em) { MG(cudaMemcpyAsync(d_start.p, k.start.p, Em * 4, cudaMemcpyDeviceToDevice, s)); MG(cudaMemcpyAsync(d_len.p, k.len.p, Em * 4, cudaMemcpyDeviceToDevice, s)); }
    if (nnzm) MG(cudaMemcpyAsync(d_lab.p, k.lab.p, nnzm * 4, cudaMemcpyDeviceToDevice, s));
    if (k.n_sgl) { k_merge_pack_singles<<<(unsigned)((k.n_sgl + 255) / 256), 256, 0, s>>>(k.sgl_tid.p, k.n_sgl, Em, nnzm, d_start.p, d_len.p, d_lab.p); c->launches++; }
    // cnt_all is exactly E long; gather through a padded copy
    MG(d_cnt.reserve(maxE));
    if (E) MG(cudaMemcpyAsync(d_cnt.p, k.cnt_all.p, E * 8, cudaMemcpyDeviceToDevice, s));
    int rc = sfb_comm_allgather(c, d_start.p, d_start_g.p, maxE * 4);
    if (!rc) rc = sfb_comm_allgather(c, d_len.p, d_len_g.p, maxE * 4);
    if (!rc) rc = sfb_comm_allgather(c, d_cnt.p, d_cnt_g.p, maxE * 8);
    if (!rc) rc = sfb_comm_allgather(c, d_lab.p, d_lab_g.p, maxZ * 4);
    if (rc) return rc;
    // room for every rank's classes before they are added (an upsert that fails here has no chunk to replay)
    {
        uint64_t all_E = 0, other_Z = 0;
        for (int r = 0; r < R; ++r) { all_E += sizes[2 * r]; if (r != c->rank) other_Z += sizes[2 * r + 1]; }
        uint64_t nb = m->n_buckets, na = m->arena_words;
        while (nb * 4 < 2 * all_E) nb *= 2;
        unsigned long long used = 0;
        MG(cudaMemcpyAsync(&used, m->cursor.p, 8, cudaMemcpyDeviceToHost, s));
        MG(cudaStreamSynchronize(s));
        while (na < used + other_Z + 64) na *= 2;
        if (nb != m->n_buckets || na != m->arena_words) { const int rg = eq_resize(c, m, nb, na); if (rg) return rg; }
    }
    EqTable tb;
    tb.slot = m->slot.p; tb.count = m->count.p; tb.arena = m->arena.p; tb.cursor = m->cursor.p;
    tb.n_buckets = m->n_buckets; tb.n_overflow = m->n_overflow; tb.arena_words = m->arena_words;
    k_merge_upsert<<<dim3((unsigned)std::min<uint64_t>((maxE + 255) / 256, 4096), R), 256, 0, s>>>(tb, d_start_g.p, d_len_g.p, d_cnt_g.p, d_lab_g.p,
                                                                                                 d_sizes.p, maxE, maxZ, c->rank);
    c->launches++;
    MG(cudaGetLastError());
    unsigned long long h_cursor[4];
    MG(cudaMemcpyAsync(h_cursor, m->cursor.p, sizeof(h_cursor), cudaMemcpyDeviceToHost, s));
    MG(cudaStreamSynchronize(s));
    cleanup();
#undef MG
#undef MGRC
    if (h_cursor[2] & ERR_ARENA_FULL) SFB_FAIL(c, SFB200_EFULL, "equivalence-class label arena exhausted while merging ranks");
    if (h_cursor[2] & ERR_TABLE_FULL) SFB_FAIL(c, SFB200_EFULL, "equivalence-class table exhausted while merging ranks");
    return SFB200_OK;
}

extern "C" int sfb200_map_finish(sfb200_ctx* c, uint64_t counters[6], uint32_t* fld_hist, uint64_t* n_classes, uint64_t* nnz) {
    if (!c) return SFB200_EINVAL;
    MapState* m = c->map;
    if (!m || !m->begun) SFB_FAIL(c, SFB200_EINVAL, "map_finish: call map_begin first");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    if (c->n_ranks > 1) {
        int rc = sfb_comm_allreduce_u64(c, m->counters.p, 6);
        if (rc) return rc;
        // fld histogram: widen to u64 through the host (1000 entries)
    }
    { const int rc = eq_grow_and_retry(c, m); if (rc) return rc; }      // the last chunk may have found the table / arena full
    unsigned long long h_cursor[4], h_counters[6];
    SFB_CUDA(c, cudaMemcpyAsync(h_cursor, m->cursor.p, sizeof(h_cursor), cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaMemcpyAsync(h_counters, m->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> h_fld(m->o.max_frag_len);
    SFB_CUDA(c, cudaMemcpyAsync(h_fld.data(), m->fld_hist.p, m->o.max_frag_len * 4ull, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    m->kernel_ms = 0.0;
    for (size_t i = 0; i + 1 < m->ev_used; i += 2) { float ms = 0.f; if (cudaEventElapsedTime(&ms, m->ev[i], m->ev[i + 1]) == cudaSuccess) m->kernel_ms += ms; }
    if (h_cursor[2] & ERR_ARENA_FULL) SFB_FAIL(c, SFB200_EFULL, "equivalence-class label arena exhausted (raise SFB200_EQ_ARENA_LOG2)");
    if (h_cursor[2] & ERR_TABLE_FULL) SFB_FAIL(c, SFB200_EFULL, "equivalence-class table exhausted (raise SFB200_EQ_LOG2_BUCKETS)");
    if (h_cursor[2] & ERR_LABEL_LONG) SFB_FAIL(c, SFB200_EFULL, "a label has 1024 or more transcripts");
    if (c->n_ranks > 1 && m->o.num_frag_samples > 0) {
        // "the first num_frag_samples eligible fragments in global read order" with reads sharded by contiguous ranges:
        // rank 0's samples come first, then rank 1's, ... -- gather the ordered samples and cut at the budget
        const int total = m->o.num_frag_samples;
        const size_t row = ((size_t)total * 2 + 8 + 15) & ~(size_t)15;          // samples + taken count, per rank
        DevBuf<unsigned char>& d_send = m->fld_send; DevBuf<unsigned char>& d_recv = m->fld_recv;
        SFB_CUDA(c, d_send.reserve(row)); SFB_CUDA(c, d_recv.reserve(row * c->n_ranks));
        int rem = 0;
        SFB_CUDA(c, cudaMemcpyAsync(&rem, m->remaining.p, 4, cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        const long long taken = total - rem;
        SFB_CUDA(c, cudaMemcpyAsync(d_send.p, m->fld_samples.p, (size_t)total * 2, cudaMemcpyDeviceToDevice, s));
        SFB_CUDA(c, cudaMemcpyAsync(d_send.p + (size_t)total * 2, &taken, 8, cudaMemcpyHostToDevice, s));
        const int rc = sfb_comm_allgather(c, d_send.p, d_recv.p, row);
        if (rc) return rc;
        std::vector<unsigned char> all(row * c->n_ranks);
        SFB_CUDA(c, cudaMemcpyAsync(all.data(), d_recv.p, all.size(), cudaMemcpyDeviceToHost, s));
        SFB_CUDA(c, cudaStreamSynchronize(s));
        std::fill(h_fld.begin(), h_fld.end(), 0u);
        long long budget = total;
        for (int r = 0; r < c->n_ranks && budget > 0; ++r) {
            const int16_t* smp = reinterpret_cast<const int16_t*>(all.data() + row * r);
            long long n_r; std::memcpy(&n_r, all.data() + row * r + (size_t)total * 2, 8);
            for (long long i = 0; i < n_r && budget > 0; ++i, --budget) h_fld[smp[i]]++;
        }
    }
    if (counters) for (int i = 0; i < 6; ++i) counters[i] = h_counters[i];
    if (fld_hist) std::memcpy(fld_hist, h_fld.data(), h_fld.size() * 4);

    { const int rc = flatten_classes(c, m, n_classes, nnz); if (rc) return rc; }
    c->cls.merged = false;
    if (c->n_ranks > 1 && !getenv("SFB200_MULTI_EM_ALLREDUCE")) {
        // One exchange instead of one per EM iteration: all-gather every rank's (label, count) list, add the other ranks'
        // classes to the local table (same upsert as the mapper's) and flatten again.  Every rank then holds the merged
        // class set and runs the whole EM locally -- an EM iteration (~8 us) is shorter than an all-reduce of the
        // per-transcript vector (~25-40 us), see DESIGN.md section 7.
        const int rc = merge_classes_over_ranks(c, m);
        if (rc) return rc;
        const int rc2 = flatten_classes(c, m, n_classes, nnz);
        if (rc2) return rc2;
        c->cls.merged = true;
    }
    return SFB200_OK;
}
