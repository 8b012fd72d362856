// TEST INFRASTRUCTURE ONLY (see oracle.h).
// CPU restatement of the read -> equivalence-class stage.
//
//  * quasi-mapping: RapMap tag sf-v0.10.1 is NOT under /root/reference
//    (scripts/fetchRapMap.sh:20).  PARITY UNPINNED.  The algorithm below is "mapping spec v1"
//    (DESIGN.md section 3), a restatement of the published RapMap quasi-mapping procedure:
//    k-mer hash -> suffix-array interval -> maximum mappable prefix (MMP) -> skip to the next
//    informative position -> intersect the transcript sets of all intervals -> left/right merge.
//    Call sites it serves: SailfishQuantify.cpp:192-213 (PE), :526-528 (SE).
//  * hit filter -> label -> count (rows A4-A8): SailfishQuantify.cpp:215-439 (PE), :530-631 (SE),
//    TranscriptGroup.cpp:9-19,53-55, EquivalenceClassBuilder.hpp:90-108.
#include "oracle.h"
#include "orc_threads.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

constexpr int MAX_IV = 16;          // spec v1: at most 16 MMP intervals per orientation scan
constexpr uint64_t EMPTY_KEY = ~0ULL;
constexpr uint64_t MAX_READ_LEN = 256;

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

inline int base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

}  // namespace

struct orc_index {
    int k = 31;
    uint32_t T = 0;
    uint64_t text_len = 0;
    std::vector<uint64_t> words;        // 2-bit text, 32 bases per word, base p at bits 2*(p%32)
    std::vector<uint64_t> txp_start;    // T+1 prefix sums in packed coordinates
    std::vector<uint32_t> txp_len;
    std::vector<uint32_t> sa_pos;       // valid positions sorted by (k-mer value [base i at bits 2i], position)
    std::vector<uint32_t> sa_tid;
    std::vector<uint64_t> kmers;        // distinct k-mers ascending
    std::vector<uint32_t> lb, cnt;      // bucket [lb, lb+cnt) in sa_pos
    // oracle's own open-addressing table over `kmers` (slot -> index into kmers), XXH64-hashed
    std::vector<uint32_t> slot;
    uint64_t mask = 0;

    inline int base(uint64_t p) const { return static_cast<int>((words[p >> 5] >> (2 * (p & 31))) & 3); }
    uint64_t kmer_at(uint64_t p) const {
        uint64_t v = 0;
        for (int i = 0; i < k; ++i) v |= static_cast<uint64_t>(base(p + i)) << (2 * i);   // base i of the k-mer at bits 2i
        return v;
    }
    // optional externally built table (mix(k-mer) & mask, linear probing): {k-mer, lb | cnt << 32},
    // empty = all-ones k-mer.  Used by bench.py's CPU arm at full size (orc_index_from_table).
    std::vector<uint64_t> ext;
    // looks the k-mer up, counts probes; fills the bucket [lb, lb+cnt)
    bool find(uint64_t km, uint64_t& probes, uint32_t& olb, uint32_t& ocnt) const {
        uint64_t h = orc_xxh64(&km, 8, 0) & mask;
        if (!ext.empty()) {
            // the externally built table is slotted with a two-multiply mixer (sailfish_b200/csrc/common.cuh: sfb_kmer_mix)
            uint64_t x = km;
            x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 29;
            h = (x & (mask >> 1)) << 1;                                  // probing starts on an even slot
            for (;;) {
                ++probes;
                const uint64_t key = ext[2 * h];
                if (key == ~0ULL) return false;
                if (key == km) { olb = static_cast<uint32_t>(ext[2 * h + 1]); ocnt = static_cast<uint32_t>(ext[2 * h + 1] >> 32); return true; }
                h = (h + 1) & mask;
            }
        }
        for (;;) {
            ++probes;
            const uint32_t s = slot[h];
            if (s == 0xFFFFFFFFu) return false;
            if (kmers[s] == km) { olb = lb[s]; ocnt = cnt[s]; return true; }
            h = (h + 1) & mask;
        }
    }
    // everything derived from the suffix order: transcript of every entry, the distinct k-mers with their buckets, the k-mer table.
    // n_threads > 1: entries are cut into per-thread ranges (bucket heads found independently, lists concatenated in range order --
    // the result does not depend on the thread count) and the table is filled with compare-and-swap (slot order may differ between
    // runs; look-ups do not depend on it).
    void finish_from_sa(int n_threads = 1) {
        const uint64_t n = sa_pos.size();
        sa_tid.resize(n);
        kmers.clear(); lb.clear(); cnt.clear();
        orc::Pool pool(n_threads);
        const int W = pool.size();
        std::vector<std::vector<uint64_t>> t_km(W);
        std::vector<std::vector<uint32_t>> t_lb(W);
        std::vector<uint64_t> t_b(W, 0), t_e(W, 0);
        pool.run_each([&](size_t w) {
            const uint64_t per = (n + W - 1) / W;
            const uint64_t b = std::min<uint64_t>(n, per * w), e = std::min<uint64_t>(n, b + per);
            t_b[w] = b; t_e[w] = e;
            uint64_t prev = b ? kmer_at(sa_pos[b - 1]) : EMPTY_KEY;
            for (uint64_t i = b; i < e; ++i) {
                const uint64_t p = sa_pos[i];
                sa_tid[i] = static_cast<uint32_t>(std::upper_bound(txp_start.begin(), txp_start.end(), p) - txp_start.begin() - 1);
                const uint64_t km = kmer_at(p);
                if (km != prev) { t_km[w].push_back(km); t_lb[w].push_back(static_cast<uint32_t>(i)); prev = km; }
            }
        });
        size_t total = 0;
        for (int w = 0; w < W; ++w) total += t_km[w].size();
        kmers.reserve(total); lb.reserve(total);
        for (int w = 0; w < W; ++w) { kmers.insert(kmers.end(), t_km[w].begin(), t_km[w].end()); lb.insert(lb.end(), t_lb[w].begin(), t_lb[w].end()); }
        cnt.resize(total);
        for (size_t i = 0; i < total; ++i) cnt[i] = static_cast<uint32_t>((i + 1 < total ? lb[i + 1] : n) - lb[i]);
        uint64_t cap = 16;
        while (cap < 2 * kmers.size()) cap <<= 1;
        mask = cap - 1;
        slot.assign(cap, 0xFFFFFFFFu);
        pool.parallel_for(kmers.size(), [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                uint64_t h = orc_xxh64(&kmers[i], 8, 0) & mask;
                for (;;) {
                    if (slot[h] == 0xFFFFFFFFu && __sync_bool_compare_and_swap(&slot[h], 0xFFFFFFFFu, static_cast<uint32_t>(i))) break;
                    h = (h + 1) & mask;
                }
            }
        });
    }
};

extern "C" orc_index* orc_index_build(const char* seq, const uint64_t* txp_off, const uint32_t* txp_len,
                                      uint32_t n_txp, int k, int n_threads) {
    if (k < 1 || k > 31) return nullptr;
    orc_index* ix = new orc_index();
    ix->k = k; ix->T = n_txp;
    ix->txp_start.resize(n_txp + 1);
    ix->txp_len.assign(txp_len, txp_len + n_txp);
    uint64_t tot = 0;
    for (uint32_t t = 0; t < n_txp; ++t) { ix->txp_start[t] = tot; tot += txp_len[t]; }
    ix->txp_start[n_txp] = tot;
    ix->text_len = tot;
    ix->words.assign(tot / 32 + 2, 0);
    for (uint32_t t = 0; t < n_txp; ++t) {
        for (uint32_t i = 0; i < txp_len[t]; ++i) {
            const uint64_t p = ix->txp_start[t] + i;
            int c = base_code(seq[txp_off[t] + i]);
            if (c > 3) c = static_cast<int>(splitmix64(p) >> 62);     // spec v1: non-ACGT -> deterministic pseudo-random base
            ix->words[p >> 5] |= static_cast<uint64_t>(c) << (2 * (p & 31));
        }
    }
    // (k-mer, position) pairs sorted by k-mer then position.  Large inputs (bench.py's CPU arm: 3.4e8 positions): the pairs are
    // binned by the k-mer's top 12 bits, every thread emits the pairs of a contiguous transcript range into per-(thread, bin) runs laid
    // out in thread order -- so a bin holds its pairs in position order before it is sorted -- and the bins are sorted in parallel.
    orc::Pool pool(n_threads);
    const int W = pool.size();
    const int NB = 4096, bshift = 2 * k > 12 ? 2 * k - 12 : 0;
    std::vector<uint64_t> tb(W + 1, 0);
    for (int w = 0; w <= W; ++w) tb[w] = static_cast<uint64_t>(n_txp) * w / W;
    std::vector<uint64_t> counts(static_cast<size_t>(W) * NB, 0);
    auto each_kmer = [&](uint32_t t0, uint32_t t1, const std::function<void(uint64_t, uint32_t)>& fn) {
        for (uint32_t t = t0; t < t1; ++t) {
            if (txp_len[t] < static_cast<uint32_t>(k)) continue;
            const uint64_t s = ix->txp_start[t];
            uint64_t v = 0;
            for (uint32_t i = 0; i < txp_len[t]; ++i) {
                v = (v >> 2) | (static_cast<uint64_t>(ix->base(s + i)) << (2 * (k - 1)));
                if (i + 1 >= static_cast<uint32_t>(k)) fn(v, static_cast<uint32_t>(s + i + 1 - k));
            }
        }
    };
    pool.run_each([&](size_t w) {
        uint64_t* c = counts.data() + w * NB;
        each_kmer(static_cast<uint32_t>(tb[w]), static_cast<uint32_t>(tb[w + 1]), [&](uint64_t v, uint32_t) { ++c[(v >> bshift) & (NB - 1)]; });
    });
    std::vector<uint64_t> start(static_cast<size_t>(W) * NB, 0), bin_start(NB + 1, 0);
    uint64_t acc = 0;
    for (int b = 0; b < NB; ++b) {
        bin_start[b] = acc;
        for (int w = 0; w < W; ++w) { start[static_cast<size_t>(w) * NB + b] = acc; acc += counts[static_cast<size_t>(w) * NB + b]; }
    }
    bin_start[NB] = acc;
    std::vector<std::pair<uint64_t, uint32_t>> kp(acc);
    pool.run_each([&](size_t w) {
        uint64_t* c = start.data() + w * NB;
        each_kmer(static_cast<uint32_t>(tb[w]), static_cast<uint32_t>(tb[w + 1]), [&](uint64_t v, uint32_t p) { kp[c[(v >> bshift) & (NB - 1)]++] = {v, p}; });
    });
    pool.parallel_for(NB, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) std::sort(kp.begin() + bin_start[i], kp.begin() + bin_start[i + 1]);
    });
    ix->sa_pos.resize(kp.size());
    pool.parallel_for(kp.size(), [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) ix->sa_pos[i] = kp[i].second; });
    std::vector<std::pair<uint64_t, uint32_t>>().swap(kp);
    ix->finish_from_sa(n_threads);
    return ix;
}

// Build the oracle's index around an externally supplied packed text + suffix order (used by bench.py's
// cpu_baseline at full size so the CPU arm does not spend minutes sorting; the order itself is checked
// against orc_index_build at test sizes).
extern "C" orc_index* orc_index_from_arrays(const uint64_t* words, uint64_t text_len, const uint32_t* txp_len,
                                            uint32_t n_txp, int k, const uint32_t* sa_pos, uint64_t n_sa) {
    orc_index* ix = new orc_index();
    ix->k = k; ix->T = n_txp; ix->text_len = text_len;
    ix->words.assign(words, words + text_len / 32 + 2);
    ix->txp_len.assign(txp_len, txp_len + n_txp);
    ix->txp_start.resize(n_txp + 1);
    uint64_t tot = 0;
    for (uint32_t t = 0; t < n_txp; ++t) { ix->txp_start[t] = tot; tot += txp_len[t]; }
    ix->txp_start[n_txp] = tot;
    ix->sa_pos.assign(sa_pos, sa_pos + n_sa);
    ix->finish_from_sa();
    return ix;
}

// Same, with the suffix order's transcript ids and the k-mer table supplied too (nothing is rebuilt): table16 holds
// n_slots entries of {k-mer u64, first entry u32, entry count u32}, n_slots a power of two.
extern "C" orc_index* orc_index_from_table(const uint64_t* words, uint64_t text_len, const uint32_t* txp_len, uint32_t n_txp,
                                           int k, const uint32_t* sa_pos, const uint32_t* sa_tid, uint64_t n_sa,
                                           const uint64_t* table16, uint64_t n_slots) {
    orc_index* ix = new orc_index();
    ix->k = k; ix->T = n_txp; ix->text_len = text_len;
    ix->words.assign(words, words + text_len / 32 + 2);
    ix->txp_len.assign(txp_len, txp_len + n_txp);
    ix->txp_start.resize(n_txp + 1);
    uint64_t tot = 0;
    for (uint32_t t = 0; t < n_txp; ++t) { ix->txp_start[t] = tot; tot += txp_len[t]; }
    ix->txp_start[n_txp] = tot;
    ix->sa_pos.assign(sa_pos, sa_pos + n_sa);
    ix->sa_tid.assign(sa_tid, sa_tid + n_sa);
    ix->ext.assign(table16, table16 + 2 * n_slots);
    ix->mask = n_slots - 1;
    return ix;
}

extern "C" void orc_index_free(orc_index* ix) { delete ix; }
extern "C" uint64_t orc_index_n_sa(const orc_index* ix) { return ix->sa_pos.size(); }
extern "C" uint64_t orc_index_n_kmers(const orc_index* ix) { return ix->kmers.size(); }
extern "C" uint64_t orc_index_text_len(const orc_index* ix) { return ix->text_len; }
extern "C" void orc_index_export(const orc_index* ix, uint32_t* sa_pos, uint32_t* sa_tid, uint64_t* kmers, uint32_t* lb, uint32_t* cnt) {
    if (sa_pos) std::memcpy(sa_pos, ix->sa_pos.data(), 4 * ix->sa_pos.size());
    if (sa_tid) std::memcpy(sa_tid, ix->sa_tid.data(), 4 * ix->sa_tid.size());
    if (kmers) std::memcpy(kmers, ix->kmers.data(), 8 * ix->kmers.size());
    if (lb) std::memcpy(lb, ix->lb.data(), 4 * ix->lb.size());
    if (cnt) std::memcpy(cnt, ix->cnt.data(), 4 * ix->cnt.size());
}
extern "C" void orc_index_export_text(const orc_index* ix, uint64_t* words) {
    std::memcpy(words, ix->words.data(), 8 * (ix->text_len / 32 + 2));
}

// ------------------------------------------------------------------------------------------------
namespace {

struct Hit {           // the QuasiAlignment fields Sailfish consumes (SailfishQuantify.cpp:221,261-265,344-349,423-428)
    uint32_t tid; int32_t pos; bool fwd; uint32_t readLen;
    int mateStatus;    // 0 SE, 1 LEFT, 2 RIGHT, 3 PAIRED
    int32_t matePos = 0; bool mateIsFwd = false; uint32_t mateLen = 0; uint32_t fragLen = 0;
};

struct Interval { uint32_t lb, cnt, qpos, m; };

struct Work { uint64_t P = 0, S = 0, X = 0; };

struct ThreadOut {
    std::unordered_map<std::string, uint64_t> classes;   // key = raw bytes of the uint32 label
    uint64_t observed = 0, mapped = 0, fragHits = 0, ubHits = 0;
    int64_t numFwd = 0, numRC = 0;
    std::vector<int32_t> fld;                            // per read: fragLen if FLD-eligible else -1 (chunk order)
    std::vector<int32_t> bias;                           // per read: 6-mer context index of its bias sample else -1 (only when collected)
    uint32_t gc[101] = {0};                              // observed fragment GC histogram of this chunk
    Work work;
};

// ---- bias / GC sample collection (SailfishQuantify.cpp:255-287,372-389 paired; :555-583 single) ---------------------------------
struct BiasCfg { bool seq = false, gc = false; };

// ReadKmerDist<6>::update (include/ReadKmerDist.hpp:36-72) for a hit: the 6-mer context around the read start on the transcript,
// reverse-complemented for forward hits; -1 if the window does not fit (or the start lies outside (0, RefLength))
inline int32_t bias_context_index(const orc_index& ix, const Hit& h) {
    constexpr int K = 6;
    const int32_t refLen = static_cast<int32_t>(ix.txp_len[h.tid]);
    const int32_t startPos = h.fwd ? h.pos : h.pos + static_cast<int32_t>(h.readLen);      // :276
    if (!(startPos > 0 && startPos < refLen)) return -1;                                    // :278
    const uint64_t t0 = ix.txp_start[h.tid];
    uint32_t idx = 0;
    if (h.fwd) {                                                                             // Direction::FORWARD: window starts 2 before
        const int32_t p = startPos - 2;
        if (!(startPos >= 2 && p + K < refLen)) return -1;
        for (int i = K - 1; i >= 0; --i) { idx += static_cast<uint32_t>(3 - ix.base(t0 + p + i)); if (i > 0) idx <<= 2; }   // indexForKmer(.., REVERSE_COMPLEMENT)
    } else {                                                                                 // REVERSE_COMPLEMENT: window starts 4 before
        const int32_t p = startPos - 4;
        if (!(startPos >= 4 && p + K < refLen)) return -1;
        for (int i = 0; i < K; ++i) { idx += static_cast<uint32_t>(ix.base(t0 + p + i)); if (i < K - 1) idx <<= 2; }        // indexForKmer(.., FORWARD)
    }
    return static_cast<int32_t>(idx);
}
// Transcript::gcFrac(s, e) (include/Transcript.hpp:85-96): G/C bases in (s, e] over the closed interval's length
inline int32_t gc_frac(const orc_index& ix, uint32_t tid, int32_t s, int32_t e) {
    const uint64_t t0 = ix.txp_start[tid];
    uint32_t n = 0;
    for (int32_t i = s + 1; i <= e; ++i) { const int b = ix.base(t0 + i); n += (b == 1 || b == 2) ? 1u : 0u; }
    return static_cast<int32_t>(std::lrint((100.0 * n) / (e - s + 1)));
}
// the per-read part of both samplers: the first hit (in jointHits order) whose context window fits gives the read's bias sample;
// every properly paired hit inside its transcript adds one observation to the fragment GC histogram
inline void collect_bias_samples(const orc_index& ix, const BiasCfg& bc, const std::vector<Hit>& joint, ThreadOut& t) {
    int32_t sample = -1;
    for (const Hit& h : joint) {
        if (bc.seq && sample < 0) sample = bias_context_index(ix, h);
        if (bc.gc && h.mateStatus == 3) {                                                    // :375-388
            const int32_t start = std::min(h.pos, h.matePos), stop = start + static_cast<int32_t>(h.fragLen);
            if (start > 0 && stop < static_cast<int32_t>(ix.txp_len[h.tid])) t.gc[gc_frac(ix, h.tid, start, stop)]++;
        }
    }
    if (bc.seq) t.bias.push_back(sample);
}

struct Mapper {
    const orc_index& ix;
    const orc_map_opts& o;
    Work& work;
    Mapper(const orc_index& i, const orc_map_opts& oo, Work& w) : ix(i), o(oo), work(w) {}

    // longest common extension of codes[q..L) with text[p..end), both after the k-mer
    uint32_t lcp_at(const std::vector<uint8_t>& s, uint32_t qpos, uint32_t sa_i) const {
        const int k = ix.k;
        const uint64_t p = ix.sa_pos[sa_i];
        const uint64_t tend = ix.txp_start[ix.sa_tid[sa_i]] + ix.txp_len[ix.sa_tid[sa_i]];
        uint32_t m = k;
        while (qpos + m < s.size() && p + m < tend) {
            ++work.X;
            if (s[qpos + m] > 3 || s[qpos + m] != ix.base(p + m)) break;
            ++m;
        }
        return m;
    }

    void scan(const std::vector<uint8_t>& s, std::vector<Interval>& ivs) const {
        ivs.clear();
        const int k = ix.k;
        const uint32_t L = static_cast<uint32_t>(s.size());
        uint32_t i = 0;
        while (i + k <= L && static_cast<int>(ivs.size()) < MAX_IV) {
            int lastN = -1;
            for (int j = 0; j < k; ++j) if (s[i + j] > 3) lastN = j;
            if (lastN >= 0) { i += lastN + 1; continue; }            // jump past the last invalid base in the window
            uint64_t km = 0;
            bool homo = true;
            for (int j = 0; j < k; ++j) { km |= static_cast<uint64_t>(s[i + j]) << (2 * j); if (s[i + j] != s[i]) homo = false; }
            if (homo) { i += 1; continue; }                          // homopolymer k-mers are never used as seeds
            uint32_t lb = 0, cnt = 0;
            if (!ix.find(km, work.P, lb, cnt) || cnt > o.max_interval) { i += 1; continue; }
            uint32_t m = 0;
            for (uint32_t e = lb; e < lb + cnt; ++e) { ++work.S; m = std::max(m, lcp_at(s, i, e)); }
            ivs.push_back({lb, cnt, i, m});
            i += m - k + 1;                                          // next k-mer ends one base past the MMP
        }
    }

    // transcripts present (with the maximal match) in EVERY interval; first (lowest) position per transcript
    // taken from interval 0.  Output ascending by tid.  Stops after max_read_occs+1 hits.
    void project(const std::vector<uint8_t>& s, const std::vector<Interval>& ivs, bool fwd, int mateStatus,
                 std::vector<Hit>& out) const {
        out.clear();
        if (ivs.empty()) return;
        const Interval& a = ivs[0];
        int64_t lastTid = -1;
        for (uint32_t e = a.lb; e < a.lb + a.cnt; ++e) {
            const uint32_t tid = ix.sa_tid[e];
            if (static_cast<int64_t>(tid) == lastTid) continue;
            ++work.S;
            if (lcp_at(s, a.qpos, e) != a.m) continue;
            lastTid = tid;                                           // first maximal entry of this transcript decides
            bool all = true;
            for (size_t j = 1; j < ivs.size() && all; ++j) {
                const Interval& b = ivs[j];
                // bucket entries are position-sorted, hence tid-sorted: binary search the tid run
                uint32_t lo = b.lb, hi = b.lb + b.cnt;
                while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (ix.sa_tid[mid] < tid) lo = mid + 1; else hi = mid; }
                bool found = false;
                for (uint32_t e2 = lo; e2 < b.lb + b.cnt && ix.sa_tid[e2] == tid; ++e2) {
                    ++work.S;
                    if (lcp_at(s, b.qpos, e2) == b.m) { found = true; break; }
                }
                all = found;
            }
            if (!all) continue;
            Hit h;
            h.tid = tid;
            h.pos = static_cast<int32_t>(static_cast<int64_t>(ix.sa_pos[e]) - static_cast<int64_t>(ix.txp_start[tid]) - static_cast<int64_t>(a.qpos));
            h.fwd = fwd; h.readLen = static_cast<uint32_t>(s.size()); h.mateStatus = mateStatus;
            out.push_back(h);
            if (out.size() > o.max_read_occs) return;                // list overflow: caller treats as "too many hits"
        }
    }

    // one mate: scan both orientations, choose, merge by tid.  Returns false if a list overflowed.
    bool collect(const std::vector<uint8_t>& fw, int mateStatus, bool strict, std::vector<Hit>& out) const {
        std::vector<uint8_t> rc(fw.size());
        for (size_t i = 0; i < fw.size(); ++i) { const uint8_t c = fw[fw.size() - 1 - i]; rc[i] = c > 3 ? c : static_cast<uint8_t>(3 - c); }
        std::vector<Interval> ivF, ivR;
        std::vector<Hit> hF, hR;
        scan(fw, ivF); scan(rc, ivR);
        project(fw, ivF, true, mateStatus, hF);
        project(rc, ivR, false, mateStatus, hR);
        if (hF.size() > o.max_read_occs || hR.size() > o.max_read_occs) { out.clear(); return false; }
        uint64_t scF = 0, scR = 0;
        for (auto& v : ivF) scF += v.m;
        for (auto& v : ivR) scR += v.m;
        if (strict && !hF.empty() && !hR.empty()) {                  // orientation vote by MMP coverage; tie keeps both
            if (scF > scR) hR.clear(); else if (scR > scF) hF.clear();
        }
        out.resize(hF.size() + hR.size());
        std::merge(hF.begin(), hF.end(), hR.begin(), hR.end(), out.begin(),
                   [](const Hit& x, const Hit& y) { return x.tid < y.tid; });   // stable: fwd before rc on equal tid
        if (out.size() > o.max_read_occs) { out.clear(); return false; }
        return true;
    }
};

void encode(const char* b, uint64_t n, std::vector<uint8_t>& s) {
    if (n > MAX_READ_LEN) n = MAX_READ_LEN;                          // spec v1: reads are clipped to 256 bases
    s.resize(n);
    for (uint64_t i = 0; i < n; ++i) s[i] = static_cast<uint8_t>(base_code(b[i]));
}

inline void add_class(ThreadOut& t, const std::vector<uint32_t>& label) {
    // EquivalenceClassBuilder::addGroup (EquivalenceClassBuilder.hpp:90-108): key equality is the full
    // vector (TranscriptGroup.cpp:53-55); the XXH64 of the label only selects the bucket.
    t.classes[std::string(reinterpret_cast<const char*>(label.data()), label.size() * 4)]++;
}

// SailfishQuantify.cpp:215-439 (paired) for one fragment.
void process_pair(const Mapper& mp, const orc_map_opts& o, const BiasCfg& bc, const std::vector<uint8_t>& r1, const std::vector<uint8_t>& r2,
                  ThreadOut& t, std::vector<uint32_t>* dbgLabel) {
    std::vector<Hit> left, right, joint;
    const bool okL = mp.collect(r1, 1, true, left);                   // strict check (:192-202)
    const bool okR = mp.collect(r2, 2, true, right);
    const bool overflow = !okL || !okR;                                // a mate list exceeded max_read_occs: "too many hits"
    if (!overflow) {
        // mergeLeftRightHits[Fuzzy] (RapMap; call sites :204-213): two-pointer merge by transcript id
        size_t i = 0, j = 0;
        while (i < left.size() && j < right.size()) {
            if (left[i].tid < right[j].tid) ++i;
            else if (right[j].tid < left[i].tid) ++j;
            else {
                const uint32_t tid = left[i].tid;
                // one joint hit per transcript: its first left hit paired with its first right hit
                Hit h = left[i];
                const Hit& r = right[j];
                h.mateStatus = 3; h.matePos = r.pos; h.mateIsFwd = r.fwd; h.mateLen = r.readLen;
                const int32_t fs = std::min(h.pos, r.pos);
                const int32_t fe = std::max(h.pos + static_cast<int32_t>(h.readLen), r.pos + static_cast<int32_t>(r.readLen));
                h.fragLen = static_cast<uint32_t>(fe - fs);
                joint.push_back(h);
                ++i;
                while (i < left.size() && left[i].tid == tid) ++i;    // one joint hit per transcript
                while (j < right.size() && right[j].tid == tid) ++j;
            }
        }
        if (joint.empty() && !o.strict_intersect) {                   // fuzzy: fall back to orphans, left block then right block
            joint.insert(joint.end(), left.begin(), left.end());
            joint.insert(joint.end(), right.begin(), right.end());
        }
    }
    t.ubHits += (overflow || joint.size() > 0) ? 1 : 0;                // :215 (an overflowed mate did have hits)
    if (joint.size() > o.max_read_occs) joint.clear();                 // :217
    bool mappedFrag = false;
    if (!joint.empty()) {
        const bool isPaired = joint.front().mateStatus == 3;
        if (!o.allow_orphans && !isPaired) joint.clear();              // :226
        if (!isPaired && !joint.empty()) {                             // :231-246 merge the two orphan blocks by tid
            auto mid = std::partition_point(joint.begin(), joint.end(), [](const Hit& q) { return q.mateStatus == 1; });
            std::inplace_merge(joint.begin(), mid, joint.end(), [](const Hit& a, const Hit& b) { return a.tid < b.tid; });
        }
        int32_t fwAll = 0, fwCompat = 0, rcAll = 0, rcCompat = 0;
        bool haveCompat = false;
        std::vector<uint32_t> idsAll, idsCompat;
        if (bc.seq || bc.gc) collect_bias_samples(mp.ix, bc, joint, t);   // inside the hit loop in the reference (:255-287,372-389)
        for (const Hit& h : joint) {
            if (!isPaired) {                                           // :289-340
                bool compat = o.ignore_compat != 0;
                if (!compat) compat = orc_compat_single(o.lib_format_id, h.pos, h.fwd, h.mateStatus) != 0;
                bool fwdHit = false;
                if (h.mateStatus == 1) { if (h.fwd) fwdHit = true; }
                else if (h.mateStatus == 2) { if (!h.fwd) fwdHit = true; }
                if (compat) { haveCompat = true; idsCompat.push_back(h.tid); if (fwdHit) fwCompat++; else rcCompat++; }
                if (!haveCompat && !o.enforce_compat) { idsAll.push_back(h.tid); if (fwdHit) fwAll++; else rcAll++; }
            } else {                                                   // :341-369
                bool compat = o.ignore_compat != 0;
                if (!compat) {
                    const uint32_t end1Pos = h.fwd ? static_cast<uint32_t>(h.pos) : static_cast<uint32_t>(h.pos) + h.readLen;
                    const uint32_t end2Pos = h.mateIsFwd ? static_cast<uint32_t>(h.matePos) : static_cast<uint32_t>(h.matePos) + h.mateLen;
                    const int obs = orc_hit_type(static_cast<int32_t>(end1Pos), h.fwd, h.readLen, static_cast<int32_t>(end2Pos),
                                                 h.mateIsFwd, h.mateLen, o.allow_dovetail);
                    compat = orc_compat_paired(o.lib_format_id, obs) != 0;
                }
                const bool fwdHit = h.fwd;
                if (compat) { haveCompat = true; idsCompat.push_back(h.tid); if (fwdHit) fwCompat++; else rcCompat++; }
                if (!haveCompat && !o.enforce_compat) { idsAll.push_back(h.tid); if (fwdHit) fwAll++; else rcAll++; }
            }
        }
        if (haveCompat) {                                              // :399-416
            if (!idsCompat.empty()) { mappedFrag = true; add_class(t, idsCompat); t.numFwd += fwCompat; t.numRC += rcCompat; if (dbgLabel) *dbgLabel = idsCompat; }
        } else if (!idsAll.empty()) {
            mappedFrag = true; add_class(t, idsAll); t.numFwd += fwAll; t.numRC += rcAll; if (dbgLabel) *dbgLabel = idsAll;
        }
    }
    if (bc.seq && t.bias.size() < t.fld.size() + 1) t.bias.push_back(-1);   // a fragment without hits gives no sample
    int32_t fl = -1;                                                   // :419-434 (sampling decided later, in global read order)
    if (joint.size() == 1 && joint.front().mateStatus == 3 && mappedFrag && joint.front().fragLen < o.max_frag_len)
        fl = static_cast<int32_t>(joint.front().fragLen);
    t.fld.push_back(fl);
    t.mapped += mappedFrag ? 1 : 0;                                    // :436-439
    t.fragHits += joint.size();
    t.observed += 1;
}

// SailfishQuantify.cpp:526-631 (single-end) for one read.
void process_single(const Mapper& mp, const orc_map_opts& o, const BiasCfg& bc, const std::vector<uint8_t>& r, ThreadOut& t,
                    std::vector<uint32_t>* dbgLabel) {
    std::vector<Hit> joint;
    const bool ok = mp.collect(r, 0, false, joint);                    // default (non-strict) check (:526-528)
    t.ubHits += (!ok || joint.size() > 0) ? 1 : 0;                     // :530
    if (joint.size() > o.max_read_occs) joint.clear();                 // :533
    bool mappedFrag = false;
    if (!joint.empty()) {
        int32_t fwAll = 0, fwCompat = 0, rcAll = 0, rcCompat = 0;
        bool haveCompat = false;
        std::vector<uint32_t> idsAll, idsCompat;
        if (bc.seq) collect_bias_samples(mp.ix, BiasCfg{true, false}, joint, t);   // :555-583 (no fragment GC for single-end reads)
        for (const Hit& h : joint) {                                   // :547-605
            bool compat = o.ignore_compat != 0;
            if (!compat) compat = orc_compat_single(o.lib_format_id, h.pos, h.fwd, h.mateStatus) != 0;
            if (compat) { haveCompat = true; idsCompat.push_back(h.tid); if (h.fwd) fwCompat++; else rcCompat++; }
            if (!haveCompat && !o.enforce_compat) { idsAll.push_back(h.tid); if (h.fwd) fwAll++; else rcAll++; }
        }
        if (haveCompat) {                                              // :608-625
            if (!idsCompat.empty()) { mappedFrag = true; add_class(t, idsCompat); t.numFwd += fwCompat; t.numRC += rcCompat; if (dbgLabel) *dbgLabel = idsCompat; }
        } else if (!idsAll.empty()) {
            mappedFrag = true; add_class(t, idsAll); t.numFwd += fwAll; t.numRC += rcAll; if (dbgLabel) *dbgLabel = idsAll;
        }
    }
    if (bc.seq && t.bias.size() < t.fld.size() + 1) t.bias.push_back(-1);
    t.fld.push_back(-1);
    t.mapped += mappedFrag ? 1 : 0;                                    // :628-631
    t.fragHits += joint.size();
    t.observed += 1;
}

}  // namespace

struct orc_run {
    const orc_index* ix;
    orc_map_opts o;
    std::map<std::vector<uint32_t>, uint64_t> classes;   // ordered: canonical (label-lexicographic) export
    uint64_t counters[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> fld;
    int32_t remainingFLOps;
    BiasCfg bias;
    int32_t remainingBiasSamples = 0;                     // sfOpts.numBiasSamples (:270,283)
    std::vector<uint32_t> readBias, observedGC;           // with their pseudo-counts of 1 (ReadKmerDist.hpp:20-24, ReadExperiment.hpp:50)
    Work work;
    std::vector<std::vector<uint32_t>> lastLabels;        // debug: label per read of the last batch
    bool keepLabels = false;
};

extern "C" orc_run* orc_run_create(const orc_index* ix, const orc_map_opts* o) {
    orc_run* r = new orc_run();
    r->ix = ix; r->o = *o;
    r->fld.assign(o->max_frag_len, 0);
    r->remainingFLOps = o->num_frag_samples;
    return r;
}
extern "C" void orc_run_free(orc_run* r) { delete r; }
// --biasCorrect / --gcBiasCorrect: collect the read-start 6-mer contexts (first num_bias_samples successes in read order, the
// reference at -p 1) and the observed fragment GC histogram while mapping
extern "C" void orc_run_set_bias(orc_run* r, int seq_bias, int gc_bias, int32_t num_bias_samples) {
    r->bias.seq = seq_bias != 0; r->bias.gc = gc_bias != 0;
    r->remainingBiasSamples = num_bias_samples;
    r->readBias.assign(4096, 1u); r->observedGC.assign(101, 1u);
}
extern "C" int orc_map_finish_bias(const orc_run* r, uint32_t* read_bias /*4096*/, uint32_t* observed_gc /*101*/) {
    if (r->readBias.empty()) return -1;
    std::memcpy(read_bias, r->readBias.data(), 4096 * 4); std::memcpy(observed_gc, r->observedGC.data(), 101 * 4);
    return 0;
}
// the two per-hit functions on their own, for tests of the device-side helpers (tests/bias_core_test.cpp)
extern "C" int32_t orc_bias_context_index(const orc_index* ix, uint32_t tid, int32_t pos, int fwd, uint32_t read_len) {
    Hit h{tid, pos, fwd != 0, read_len, 0};
    return bias_context_index(*ix, h);
}
extern "C" int32_t orc_gc_frac(const orc_index* ix, uint32_t tid, int32_t s, int32_t e) { return gc_frac(*ix, tid, s, e); }
extern "C" void orc_run_keep_labels(orc_run* r, int on) { r->keepLabels = on != 0; }

extern "C" int orc_map_batch(orc_run* r, const char* bases1, const uint64_t* off1, const char* bases2,
                             const uint64_t* off2, uint64_t n_reads, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    const bool paired = bases2 != nullptr;
    std::vector<ThreadOut> outs(n_threads);
    std::vector<Work> works(n_threads);
    if (r->keepLabels) r->lastLabels.assign(n_reads, std::vector<uint32_t>());
    orc::Pool pool(n_threads);
    const size_t per = (n_reads + n_threads - 1) / n_threads;
    pool.run_each([&](size_t ti) {
        {
            ThreadOut& t = outs[ti];
            Mapper mp(*r->ix, r->o, t.work);
            std::vector<uint8_t> s1, s2;
            const size_t b = per * ti, e = std::min<size_t>(n_reads, b + per);
            for (size_t i = b; i < e; ++i) {
                std::vector<uint32_t>* dbg = r->keepLabels ? &r->lastLabels[i] : nullptr;
                encode(bases1 + off1[i], off1[i + 1] - off1[i], s1);
                if (paired) {
                    encode(bases2 + off2[i], off2[i + 1] - off2[i], s2);
                    process_pair(mp, r->o, r->bias, s1, s2, t, dbg);
                } else {
                    process_single(mp, r->o, r->bias, s1, t, dbg);
                }
            }
        }
    });
    for (int ti = 0; ti < n_threads; ++ti) {
        ThreadOut& t = outs[ti];
        for (auto& kv : t.classes) {
            std::vector<uint32_t> lab(kv.first.size() / 4);
            std::memcpy(lab.data(), kv.first.data(), kv.first.size());
            r->classes[lab] += kv.second;
        }
        r->counters[0] += t.observed; r->counters[1] += t.mapped; r->counters[2] += t.fragHits; r->counters[3] += t.ubHits;
        r->counters[4] += static_cast<uint64_t>(t.numFwd); r->counters[5] += static_cast<uint64_t>(t.numRC);
        for (int32_t fl : t.fld) {                                    // first num_frag_samples eligible fragments in read order
            if (fl >= 0 && r->remainingFLOps > 0) { r->fld[fl]++; r->remainingFLOps--; }
        }
        for (int32_t idx : t.bias) {                                  // first numBiasSamples successful reads in read order
            if (idx >= 0 && r->remainingBiasSamples > 0) { r->readBias[idx]++; r->remainingBiasSamples--; }
        }
        if (r->bias.gc) for (int g = 0; g < 101; ++g) r->observedGC[g] += t.gc[g];
        r->work.P += t.work.P; r->work.S += t.work.S; r->work.X += t.work.X;
    }
    return 0;
}

extern "C" int orc_map_finish(orc_run* r, uint64_t counters[6], uint32_t* fld_hist, uint64_t* n_classes, uint64_t* nnz) {
    if (counters) std::memcpy(counters, r->counters, sizeof(r->counters));
    if (fld_hist) std::memcpy(fld_hist, r->fld.data(), 4 * r->fld.size());
    uint64_t z = 0;
    for (auto& kv : r->classes) z += kv.first.size();
    if (n_classes) *n_classes = r->classes.size();
    if (nnz) *nnz = z;
    return 0;
}

extern "C" int orc_eq_export(const orc_run* r, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts) {
    uint64_t e = 0, z = 0;
    row_ptr[0] = 0;
    for (auto& kv : r->classes) {
        for (uint32_t t : kv.first) labels[z++] = t;
        counts[e] = kv.second;
        row_ptr[++e] = z;
    }
    return 0;
}

extern "C" void orc_map_work(const orc_run* r, uint64_t work[3]) { work[0] = r->work.P; work[1] = r->work.S; work[2] = r->work.X; }

extern "C" int orc_last_label(const orc_run* r, uint64_t i, uint32_t* out, int cap) {
    if (i >= r->lastLabels.size() || r->lastLabels[i].empty()) return -1;
    const auto& l = r->lastLabels[i];
    for (int j = 0; j < cap && j < static_cast<int>(l.size()); ++j) out[j] = l[j];
    return static_cast<int>(l.size());
}
