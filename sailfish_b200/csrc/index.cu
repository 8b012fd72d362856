// index.cu -- builds the device index of "mapping spec v1"  from transcript sequences, on the GPU.
//
// Stands in for what ReadExperiment takes from RapMapSAIndex<IndexT> after SailfishIndex::load (reference
// include/SailfishIndex.hpp:28-43,104-144; include/ReadExperiment.hpp:103-116).  RapMap's builder and its on-disk
// format are not part of the reference tree (scripts/fetchRapMap.sh:20), so the index is specified by this repo:
//   words    2-bit text of all transcripts concatenated (32 bases / u64, base p at bits 2*(p%32))
//   sa       every position whose k-mer lies inside one transcript, sorted by (k-mer value, position): {transcript, position
//            inside it, text position, bases left to the transcript's end}
//   table    open-addressing k-mer table {k-mer, first entry, entry count}, slot = mix(k-mer) & mask, linear probing
//   mfilter  presence filter: bitmap over the m-mers of the text (kmer_filter.hpp)
// Index construction is a "next" row, outside the timed path: the big sort / compaction primitives
// are CUB's (part of the CUDA toolkit); everything else is hand-written.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint32_t txp_of(const uint64_t* __restrict__ start, uint32_t n_txp, uint64_t p) {
    // last t with start[t] <= p  (transcripts of length 0 share a start: take the last one, as upper_bound - 1 does)
    uint32_t lo = 0, hi = n_txp;          // invariant: start[lo] <= p < start[hi]
    while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (start[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}

// one thread packs one 32-base word
__global__ void k_pack_text(const char* __restrict__ seq, const uint64_t* __restrict__ txp_off,
                            const uint64_t* __restrict__ txp_start, uint32_t n_txp, uint64_t text_len, uint64_t n_words,
                            uint64_t* __restrict__ words) {
    const uint64_t wi = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (wi >= n_words) return;
    uint64_t w = 0;
    const uint64_t p0 = wi * 32;
    if (p0 < text_len) {
        uint32_t t = txp_of(txp_start, n_txp, p0);
        for (int j = 0; j < 32; ++j) {
            const uint64_t p = p0 + j;
            if (p >= text_len) break;
            while (p >= txp_start[t + 1]) ++t;
            const unsigned char ch = static_cast<unsigned char>(seq[txp_off[t] + (p - txp_start[t])]);
            const unsigned char up = ch & 0xDF;
            uint64_t code;
            if (up == 'A' || up == 'C' || up == 'G' || up == 'T') code = ((ch >> 1) ^ (ch >> 2)) & 3;
            else code = splitmix64(p) >> 62;            // spec v1: non-ACGT -> deterministic pseudo-random base
            w |= code << (2 * j);
        }
    }
    words[wi] = w;
}

// entry i (in position order) -> (k-mer, position)
__global__ void k_emit_kmers(const uint64_t* __restrict__ words, const uint64_t* __restrict__ txp_start,
                             const uint64_t* __restrict__ vstart, uint32_t n_txp, int k, uint64_t n_sa,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ pos) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_sa) return;
    const uint32_t t = txp_of(vstart, n_txp, i);
    const uint64_t p = txp_start[t] + (i - vstart[t]);
    const uint64_t idx = p >> 5, sh = 2 * (p & 31);
    uint64_t v = words[idx] >> sh;
    if (sh) v |= words[idx + 1] << (64 - sh);
    keys[i] = v & ((1ULL << (2 * k)) - 1);
    pos[i] = static_cast<uint32_t>(p);
}

__global__ void k_tid_and_heads(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos,
                                const uint64_t* __restrict__ txp_start, uint32_t n_txp, uint64_t n_sa,
                                uint4* __restrict__ sa, uint8_t* __restrict__ head) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_sa) return;
    const uint32_t p = pos[i];
    const uint32_t t = txp_of(txp_start, n_txp, p);
    sa[i] = make_uint4(t, (uint32_t)(p - txp_start[t]), p, (uint32_t)(txp_start[t + 1] - p));
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void k_txp_end(const uint64_t* __restrict__ txp_start, const uint32_t* __restrict__ txp_len, uint32_t n_txp,
                          uint64_t* __restrict__ txp_end) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_txp) txp_end[t] = txp_start[t] + txp_len[t];
}

__global__ void k_table_insert(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ heads, uint64_t n_kmers,
                               uint64_t n_sa, uint4* __restrict__ table, uint64_t mask, unsigned int* __restrict__ max_bucket) {
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n_kmers) return;
    const uint32_t lb = heads[j];
    const uint32_t cnt = static_cast<uint32_t>((j + 1 < n_kmers ? heads[j + 1] : n_sa) - lb);
    const uint64_t km = keys[lb];
    const uint64_t hh = sfb_kmer_mix(km);
    uint64_t h = (hh & (mask >> 1)) << 1;            // probing starts on an even slot: slots h, h+1 share a 32-byte sector
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(table);
    for (;;) {
        const unsigned long long prev = atomicCAS(&slots[2 * h], ~0ULL, (unsigned long long)km);
        if (prev == ~0ULL) {
            uint32_t* pl = reinterpret_cast<uint32_t*>(&slots[2 * h + 1]);
            pl[0] = lb; pl[1] = cnt;
            break;
        }
        h = (h + 1) & mask;
    }
    atomicMax(max_bucket, cnt);
}

// one thread per text position: the m-mer that starts there
__global__ void k_mfilter_build(const uint64_t* __restrict__ words, uint64_t n_pos, int m, uint32_t* __restrict__ bits) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    const uint64_t idx = p >> 5, sh = 2 * (p & 31);
    uint64_t v = words[idx] >> sh;
    if (sh) v |= words[idx + 1] << (64 - sh);
    v &= (1ULL << (2 * m)) - 1;
    const uint32_t bit = 1u << (v & 31);
    uint32_t* w = bits + (v >> 5);
    if (!(*w & bit)) atomicOr(w, bit);
}

__global__ void k_table_clear(uint4* __restrict__ table, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) table[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
}

inline unsigned gridn(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int sfb200_index_build(sfb200_ctx* c, const char* seq, const uint64_t* txp_off, const uint32_t* txp_len,
                                  uint32_t n_txp, int k) {
    if (!c || !seq || !txp_off || !txp_len || n_txp == 0) return SFB200_EINVAL;
    if (k < 1 || k > 31 || (k % 2) == 0) SFB_FAIL(c, SFB200_EINVAL, "k must be odd and <= 31 (SailfishIndexer.cpp:199-205)");
    cudaSetDevice(c->device);
    DevIndex& ix = c->index;
    ix.ready = false;
    cudaStream_t s = c->stream;

    std::vector<uint64_t> start(n_txp + 1), vstart(n_txp + 1);
    uint64_t tot = 0, nsa = 0, src_end = 0;
    for (uint32_t t = 0; t < n_txp; ++t) {
        start[t] = tot; vstart[t] = nsa;
        tot += txp_len[t];
        if (txp_len[t] >= static_cast<uint32_t>(k)) nsa += txp_len[t] - k + 1;
        if (txp_off[t] + txp_len[t] > src_end) src_end = txp_off[t] + txp_len[t];
    }
    start[n_txp] = tot; vstart[n_txp] = nsa;
    if (tot >= 0xFFFFFFF0ull) SFB_FAIL(c, SFB200_EINVAL, "transcriptome longer than 2^32 bases is not supported by the 32-bit position index");
    ix.k = k; ix.n_txp = n_txp; ix.text_len = tot; ix.n_sa = nsa;
    const uint64_t n_words = tot / 32 + 2;

    DevBuf<char> d_seq; DevBuf<uint64_t> d_off, d_vstart, d_keys, d_keys2; DevBuf<uint32_t> d_pos2, d_heads, d_pos_sorted; DevBuf<uint8_t> d_head, d_tmp;
    DevBuf<unsigned int> d_scalar;
    auto cleanup = [&]() { d_seq.release(); d_off.release(); d_vstart.release(); d_keys.release(); d_keys2.release();
                           d_pos2.release(); d_heads.release(); d_head.release(); d_tmp.release(); d_scalar.release(); d_pos_sorted.release(); };
#define IDX_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); c->err = std::string(#call) + ": " + cudaGetErrorString(e__); return SFB200_ECUDA; } } while (0)

    IDX_CUDA(d_seq.reserve(src_end + 1));
    IDX_CUDA(d_off.reserve(n_txp));
    IDX_CUDA(d_vstart.reserve(n_txp + 1));
    IDX_CUDA(ix.words.reserve(n_words));
    IDX_CUDA(ix.txp_start.reserve(n_txp + 1));
    IDX_CUDA(ix.txp_len.reserve(n_txp));
    IDX_CUDA(cudaMemcpyAsync(d_seq.p, seq, src_end, cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaMemcpyAsync(d_off.p, txp_off, n_txp * 8ull, cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaMemcpyAsync(d_vstart.p, vstart.data(), (n_txp + 1) * 8ull, cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaMemcpyAsync(ix.txp_start.p, start.data(), (n_txp + 1) * 8ull, cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaMemcpyAsync(ix.txp_len.p, txp_len, n_txp * 4ull, cudaMemcpyHostToDevice, s));
    k_pack_text<<<gridn(n_words, 256), 256, 0, s>>>(d_seq.p, d_off.p, ix.txp_start.p, n_txp, tot, n_words, ix.words.p);
    c->launches++;
    IDX_CUDA(cudaGetLastError());
    IDX_CUDA(cudaStreamSynchronize(s));
    d_seq.release();

    ix.n_kmers = 0; ix.max_bucket = 0;
    IDX_CUDA(ix.sa.reserve(nsa));
    IDX_CUDA(ix.txp_end.reserve(n_txp));
    k_txp_end<<<gridn(n_txp, 256), 256, 0, s>>>(ix.txp_start.p, ix.txp_len.p, n_txp, ix.txp_end.p);
    c->launches++;
    if (nsa > 0) {
        IDX_CUDA(d_pos_sorted.reserve(nsa));
        IDX_CUDA(d_keys.reserve(nsa)); IDX_CUDA(d_keys2.reserve(nsa)); IDX_CUDA(d_pos2.reserve(nsa));
        k_emit_kmers<<<gridn(nsa, 256), 256, 0, s>>>(ix.words.p, ix.txp_start.p, d_vstart.p, n_txp, k, nsa, d_keys.p, d_pos2.p);
        c->launches++;
        // stable LSD radix sort by k-mer value: entries were emitted in position order, so ties stay position-sorted
        size_t tmp_bytes = 0;
        IDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.p, d_keys2.p, d_pos2.p, d_pos_sorted.p, nsa, 0, 2 * k, s));
        IDX_CUDA(d_tmp.reserve(tmp_bytes));
        IDX_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.p, d_keys2.p, d_pos2.p, d_pos_sorted.p, nsa, 0, 2 * k, s));
        c->launches++;
        IDX_CUDA(cudaStreamSynchronize(s));
        d_keys.release(); d_pos2.release();
        IDX_CUDA(d_head.reserve(nsa));
        k_tid_and_heads<<<gridn(nsa, 256), 256, 0, s>>>(d_keys2.p, d_pos_sorted.p, ix.txp_start.p, n_txp, nsa, ix.sa.p, d_head.p);
        c->launches++;
        // bucket heads = indices whose k-mer differs from the previous entry's
        IDX_CUDA(d_heads.reserve(nsa));
        IDX_CUDA(d_scalar.reserve(4));
        size_t tmp2 = 0;
        cub::CountingInputIterator<uint32_t> iota(0);
        IDX_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp2, iota, d_head.p, d_heads.p, d_scalar.p, nsa, s));
        IDX_CUDA(d_tmp.reserve(tmp2));
        IDX_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tmp2, iota, d_head.p, d_heads.p, d_scalar.p, nsa, s));
        c->launches++;
        unsigned int n_heads = 0;
        IDX_CUDA(cudaMemcpyAsync(&n_heads, d_scalar.p, 4, cudaMemcpyDeviceToHost, s));
        IDX_CUDA(cudaStreamSynchronize(s));
        ix.n_kmers = n_heads;
    }
    uint64_t slots = 1024;
    while (slots < 4 * ix.n_kmers) slots <<= 1;      // load <= 25%: most lookups end in the first slot
    if (slots * 16 > (32ull << 30) && (slots >> 1) * 2 >= 5 * ix.n_kmers) slots >>= 1;   // metatranscriptome scale: up to 40% rather than 64+ GB
    ix.table_slots = slots;
    IDX_CUDA(ix.table.reserve(slots));
    // presence filter: every m-mer window of the packed text (windows that span two transcripts set a few spare bits: harmless)
    ix.mf_m = sfb_mfilter_m(k, tot);
    const uint64_t mf_words = sfb_mfilter_words(ix.mf_m);
    IDX_CUDA(ix.mfilter.reserve(mf_words));
    IDX_CUDA(cudaMemsetAsync(ix.mfilter.p, 0, mf_words * 4, s));
    if (tot >= (uint64_t)ix.mf_m) {
        k_mfilter_build<<<gridn(tot - ix.mf_m + 1, 256), 256, 0, s>>>(ix.words.p, tot - ix.mf_m + 1, ix.mf_m, ix.mfilter.p);
        c->launches++;
    }
    k_table_clear<<<gridn(slots, 256), 256, 0, s>>>(ix.table.p, slots);
    c->launches++;
    if (ix.n_kmers) {
        IDX_CUDA(cudaMemsetAsync(d_scalar.p + 1, 0, 4, s));
        k_table_insert<<<gridn(ix.n_kmers, 256), 256, 0, s>>>(d_keys2.p, d_heads.p, ix.n_kmers, nsa, ix.table.p, slots - 1, d_scalar.p + 1);
        c->launches++;
        unsigned int mb = 0;
        IDX_CUDA(cudaMemcpyAsync(&mb, d_scalar.p + 1, 4, cudaMemcpyDeviceToHost, s));
        IDX_CUDA(cudaStreamSynchronize(s));
        ix.max_bucket = mb;
    }
    IDX_CUDA(cudaGetLastError());
    IDX_CUDA(cudaStreamSynchronize(s));
    cleanup();
#undef IDX_CUDA
    ix.ready = true;
    return SFB200_OK;
}

extern "C" int sfb200_index_stats(const sfb200_ctx* c, uint64_t stats[8]) {
    if (!c || !stats) return SFB200_EINVAL;
    const DevIndex& ix = c->index;
    std::memset(stats, 0, 8 * sizeof(uint64_t));
    if (!ix.ready) return SFB200_EINVAL;
    stats[0] = ix.text_len; stats[1] = ix.n_sa; stats[2] = ix.n_kmers; stats[3] = ix.table_slots;
    stats[4] = ix.hbm_bytes(); stats[5] = ix.max_bucket; stats[6] = ix.k; stats[7] = ix.n_txp;
    return SFB200_OK;
}

extern "C" int sfb200_index_export(sfb200_ctx* c, uint64_t* words, uint32_t* sa_pos, uint32_t* sa_tid) {
    if (!c) return SFB200_EINVAL;
    DevIndex& ix = c->index;
    if (!ix.ready) SFB_FAIL(c, SFB200_EINVAL, "index_export: no index");
    cudaSetDevice(c->device);
    if (words) SFB_CUDA(c, cudaMemcpyAsync(words, ix.words.p, (ix.text_len / 32 + 2) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (sa_pos && ix.n_sa) SFB_CUDA(c, cudaMemcpy2DAsync(sa_pos, 4, &ix.sa.p->z, 16, 4, ix.n_sa, cudaMemcpyDeviceToHost, c->stream));
    if (sa_tid && ix.n_sa) SFB_CUDA(c, cudaMemcpy2DAsync(sa_tid, 4, &ix.sa.p->x, 16, 4, ix.n_sa, cudaMemcpyDeviceToHost, c->stream));
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFB200_OK;
}

extern "C" int sfb200_index_export_table(sfb200_ctx* c, void* table16) {
    if (!c || !table16) return SFB200_EINVAL;
    DevIndex& ix = c->index;
    if (!ix.ready) SFB_FAIL(c, SFB200_EINVAL, "index_export_table: no index");
    cudaSetDevice(c->device);
    SFB_CUDA(c, cudaMemcpyAsync(table16, ix.table.p, ix.table_slots * 16, cudaMemcpyDeviceToHost, c->stream));
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFB200_OK;
}

// ---- the index as a file ---------------------------------------------------------------------------------------------------------------
namespace {
struct IndexFileHeader {
    char magic[8];                 // "SFB200IX"
    uint32_t version, k;
    uint32_t n_txp, mf_m, max_bucket, pad_;
    uint64_t text_len, n_sa, n_kmers, table_slots;
    uint64_t bytes[7];             // words, txp_start, txp_len, txp_end, sa, table, mfilter
};
constexpr uint32_t INDEX_FILE_VERSION = 2;       // 2: 16-byte suffix entries, m-mer bitmap
constexpr size_t IO_PIECE = 64u << 20;

int stream_out(sfb200_ctx* c, FILE* f, const void* d_ptr, uint64_t bytes, std::vector<char>& buf) {
    for (uint64_t at = 0; at < bytes; at += IO_PIECE) {
        const size_t n = (size_t)std::min<uint64_t>(IO_PIECE, bytes - at);
        SFB_CUDA(c, cudaMemcpyAsync(buf.data(), static_cast<const char*>(d_ptr) + at, n, cudaMemcpyDeviceToHost, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (fwrite(buf.data(), 1, n, f) != n) SFB_FAIL(c, SFB200_EINVAL, "index_save: write failed");
    }
    return SFB200_OK;
}
int stream_in(sfb200_ctx* c, FILE* f, void* d_ptr, uint64_t bytes, std::vector<char>& buf) {
    for (uint64_t at = 0; at < bytes; at += IO_PIECE) {
        const size_t n = (size_t)std::min<uint64_t>(IO_PIECE, bytes - at);
        if (fread(buf.data(), 1, n, f) != n) SFB_FAIL(c, SFB200_EINVAL, "index_load: the file is truncated");
        SFB_CUDA(c, cudaMemcpyAsync(static_cast<char*>(d_ptr) + at, buf.data(), n, cudaMemcpyHostToDevice, c->stream));
        SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return SFB200_OK;
}
}  // namespace

extern "C" int sfb200_index_save(sfb200_ctx* c, const char* path) {
    if (!c || !path) return SFB200_EINVAL;
    DevIndex& ix = c->index;
    if (!ix.ready) SFB_FAIL(c, SFB200_EINVAL, "index_save: no index");
    cudaSetDevice(c->device);
    IndexFileHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "SFB200IX", 8);
    h.version = INDEX_FILE_VERSION; h.k = (uint32_t)ix.k; h.n_txp = ix.n_txp; h.mf_m = (uint32_t)ix.mf_m; h.max_bucket = ix.max_bucket;
    h.text_len = ix.text_len; h.n_sa = ix.n_sa; h.n_kmers = ix.n_kmers; h.table_slots = ix.table_slots;
    const void* ptr[7] = {ix.words.p, ix.txp_start.p, ix.txp_len.p, ix.txp_end.p, ix.sa.p, ix.table.p, ix.mfilter.p};
    h.bytes[0] = (ix.text_len / 32 + 2) * 8; h.bytes[1] = (ix.n_txp + 1ull) * 8; h.bytes[2] = ix.n_txp * 4ull; h.bytes[3] = ix.n_txp * 8ull;
    h.bytes[4] = ix.n_sa * 16; h.bytes[5] = ix.table_slots * 16; h.bytes[6] = sfb_mfilter_words(ix.mf_m) * 4;
    FILE* f = fopen(path, "wb");
    if (!f) SFB_FAIL(c, SFB200_EINVAL, std::string("index_save: cannot open ") + path);
    std::vector<char> buf(IO_PIECE);
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? SFB200_OK : SFB200_EINVAL;
    for (int i = 0; i < 7 && rc == SFB200_OK; ++i) rc = stream_out(c, f, ptr[i], h.bytes[i], buf);
    if (fclose(f) != 0 && rc == SFB200_OK) { c->err = "index_save: close failed"; rc = SFB200_EINVAL; }
    return rc;
}

extern "C" int sfb200_index_load(sfb200_ctx* c, const char* path) {
    if (!c || !path) return SFB200_EINVAL;
    cudaSetDevice(c->device);
    DevIndex& ix = c->index;
    ix.ready = false;
    FILE* f = fopen(path, "rb");
    if (!f) SFB_FAIL(c, SFB200_EINVAL, std::string("index_load: cannot open ") + path);
    IndexFileHeader h;
    auto fail = [&](const char* msg) { fclose(f); c->err = msg; return SFB200_EINVAL; };
    if (fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "SFB200IX", 8) != 0) return fail("index_load: not an sfb200 index file");
    if (h.version != INDEX_FILE_VERSION) return fail("index_load: the file was written by another format version (rebuild the index)");
    if (h.k < 1 || h.k > 31 || h.n_txp == 0 || h.bytes[0] != (h.text_len / 32 + 2) * 8 || h.bytes[4] != h.n_sa * 16 ||
        h.bytes[5] != h.table_slots * 16 || (h.table_slots & (h.table_slots - 1)) != 0 || h.bytes[6] != sfb_mfilter_words((int)h.mf_m) * 4)
        return fail("index_load: inconsistent header");
    ix.k = (int)h.k; ix.n_txp = h.n_txp; ix.mf_m = (int)h.mf_m; ix.max_bucket = h.max_bucket;
    ix.text_len = h.text_len; ix.n_sa = h.n_sa; ix.n_kmers = h.n_kmers; ix.table_slots = h.table_slots;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = ix.words.reserve(h.bytes[0] / 8);
    if (e == cudaSuccess) e = ix.txp_start.reserve(h.n_txp + 1ull);
    if (e == cudaSuccess) e = ix.txp_len.reserve(h.n_txp);
    if (e == cudaSuccess) e = ix.txp_end.reserve(h.n_txp);
    if (e == cudaSuccess) e = ix.sa.reserve(std::max<uint64_t>(1, h.n_sa));
    if (e == cudaSuccess) e = ix.table.reserve(h.table_slots);
    if (e == cudaSuccess) e = ix.mfilter.reserve(h.bytes[6] / 4);
    if (e != cudaSuccess) { fclose(f); c->err = std::string("index_load: ") + cudaGetErrorString(e); return SFB200_ECUDA; }
    void* ptr[7] = {ix.words.p, ix.txp_start.p, ix.txp_len.p, ix.txp_end.p, ix.sa.p, ix.table.p, ix.mfilter.p};
    std::vector<char> buf(IO_PIECE);
    int rc = SFB200_OK;
    for (int i = 0; i < 7 && rc == SFB200_OK; ++i) rc = stream_in(c, f, ptr[i], h.bytes[i], buf);
    fclose(f);
    if (rc == SFB200_OK) ix.ready = true;
    return rc;
}
