// fastq.cu -- read ingestion on the device: a block of FASTQ TEXT goes to the GPU as it is in the file, the
// sequence lines are found and copied into the `bases` / `off` form there, and the block is mapped like any device-resident batch.
//
// Why: the mapping kernels consume ~45 GB/s of FASTQ text, eight host threads index and copy 1 - 2.5 GB/s (host/fastx_reader.hpp).
// Finding newlines and copying 100-byte lines is bandwidth work the GPU does at HBM speed; what stays on the host is reading the
// file (or inflating it) and carrying the incomplete record at the end of a block over to the next one.
//
//   k_fq_count    newlines per 512-byte chunk (8 bytes per load)     -> exclusive scan (CUB) -> index of every chunk's first newline
//   k_fq_mark     newline j ends line j of the block; record r = lines 4r .. 4r+3: sequence start / length, record end, format checks
//   k_fq_copy     a warp per record copies its bases to bases[off[r] ..), off = exclusive scan of the lengths
//   then sfb200_map_batch_device on the extracted arrays (k_pack_reads -> k_scan_reads -> k_finalize_reads, map.cu)
// The per-chunk bodies are in fastq_core.inl, which tests/fastq_core_test.cpp compiles as host code.
//
// The copy and the extraction of a block run on a stream of their own, into one of two staging sets, while the mapping kernels of the
// previous block still run on the context's stream: a call returns as soon as its mapping kernels are enqueued.
// Parity: tests/test_gpu_map.py::test_map_fastq_equals_map_batch; `sfb200-quant --deviceParse` opts in.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace {

#define SFB_FQ __device__ __forceinline__
#define SFB_FQ_OR(p, v) atomicOr((p), (v))
#define SFB_FQ_LD8(p) (*reinterpret_cast<const unsigned long long*>(p))
#define SFB_FQ_POPC(x) __popcll(x)
#define SFB_FQ_CTZ(x) (__ffsll((long long)(x)) - 1)
#include "fastq_core.inl"
#undef SFB_FQ
#undef SFB_FQ_OR
#undef SFB_FQ_LD8
#undef SFB_FQ_POPC
#undef SFB_FQ_CTZ

struct FqMate {
    DevBuf<char> text, bases;
    DevBuf<uint32_t> cnt, len;                   // newlines per chunk (then their exclusive prefix); bases per record
    DevBuf<uint64_t> seq_start, rec_end, off;
    uint64_t n_text = 0, n_chunks = 0, n_newlines = 0;
};
struct FqState {
    FqMate sets[2][2];                           // two staging sets x two mates
    unsigned parity = 0;
    DevBuf<unsigned char> tmp;
    DevBuf<uint32_t> err;
    cudaStream_t xs = nullptr;                   // copy + extraction stream
    cudaEvent_t extracted = nullptr, mapped[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    void release() {
        for (auto& st : sets) for (FqMate& x : st) { x.text.release(); x.bases.release(); x.cnt.release(); x.len.release(); x.seq_start.release(); x.rec_end.release(); x.off.release(); }
        tmp.release(); err.release();
        if (xs) cudaStreamDestroy(xs);
        if (extracted) cudaEventDestroy(extracted);
        for (cudaEvent_t& e : mapped) if (e) cudaEventDestroy(e);
        xs = nullptr; extracted = nullptr; mapped[0] = mapped[1] = nullptr;
    }
};

__global__ void k_fq_count(const char* __restrict__ text, uint64_t n, uint64_t n_chunks, uint32_t* __restrict__ cnt) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < n_chunks) cnt[c] = fq_count_newlines(text, n, c);
}
__global__ void k_fq_mark(const char* __restrict__ text, uint64_t n, uint64_t n_chunks, const uint32_t* __restrict__ nl_base, uint64_t n_rec,
                          uint32_t lines_per_rec, uint64_t* __restrict__ seq_start, uint32_t* __restrict__ seq_len, uint64_t* __restrict__ rec_end,
                          uint32_t* __restrict__ err) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < n_chunks) fq_mark_chunk(text, n, c, nl_base[c], n_rec, lines_per_rec, seq_start, seq_len, rec_end, err);
}
__global__ void k_fq_len_to_u64(const uint32_t* __restrict__ len, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? len[i] : 0;                                  // n + 1 entries: the scan's last one is the total
}
__global__ void k_fq_copy(const char* __restrict__ text, const uint64_t* __restrict__ seq_start, const uint32_t* __restrict__ seq_len,
                          const uint64_t* __restrict__ off, uint64_t n_rec, char* __restrict__ bases) {
    const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (r < n_rec) fq_copy_record(text, seq_start[r], seq_len[r], bases + off[r], threadIdx.x & 31u, 32u);
}

inline unsigned fq_grid(uint64_t n, unsigned th) { return (unsigned)((n + th - 1) / th); }

// text -> device, newline counts, their exclusive scan; leaves the number of newlines in x.n_newlines
int fq_stage(sfb200_ctx* c, FqState* st, FqMate& x, const char* text, uint64_t n) {
    cudaStream_t s = st->xs;
    x.n_text = n; x.n_chunks = (n + FQ_CHUNK - 1) / FQ_CHUNK; x.n_newlines = 0;
    if (n == 0) return SFB200_OK;
    if (n >= (1ull << 31)) SFB_FAIL(c, SFB200_EINVAL, "map_fastq: block too large (at most 2 GB of text per call: newline counts and record indices are 32-bit)");
    SFB_CUDA(c, x.text.reserve(n + 1)); SFB_CUDA(c, x.cnt.reserve(x.n_chunks + 1));
    SFB_CUDA(c, cudaMemcpyAsync(x.text.p, text, n, cudaMemcpyHostToDevice, s));
    k_fq_count<<<fq_grid(x.n_chunks, 256), 256, 0, s>>>(x.text.p, n, x.n_chunks, x.cnt.p);
    c->launches++;
    SFB_CUDA(c, cudaMemsetAsync(x.cnt.p + x.n_chunks, 0, 4, s));
    size_t tmp = 0;
    SFB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp, x.cnt.p, x.cnt.p, (int)(x.n_chunks + 1), s));
    SFB_CUDA(c, st->tmp.reserve(tmp));
    SFB_CUDA(c, cub::DeviceScan::ExclusiveSum(st->tmp.p, tmp, x.cnt.p, x.cnt.p, (int)(x.n_chunks + 1), s));
    c->launches++;
    uint32_t total = 0;
    SFB_CUDA(c, cudaMemcpyAsync(&total, x.cnt.p + x.n_chunks, 4, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    x.n_newlines = total;
    return SFB200_OK;
}

// records [0, n_rec) of a staged mate -> bases / off; *consumed = bytes of text they cover
int fq_extract(sfb200_ctx* c, FqState* st, FqMate& x, uint64_t n_rec, uint32_t lines_per_rec, uint64_t* consumed) {
    cudaStream_t s = st->xs;
    SFB_CUDA(c, x.seq_start.reserve(n_rec)); SFB_CUDA(c, x.len.reserve(n_rec)); SFB_CUDA(c, x.rec_end.reserve(n_rec)); SFB_CUDA(c, x.off.reserve(n_rec + 1));
    k_fq_mark<<<fq_grid(x.n_chunks, 256), 256, 0, s>>>(x.text.p, x.n_text, x.n_chunks, x.cnt.p, n_rec, lines_per_rec, x.seq_start.p, x.len.p, x.rec_end.p, st->err.p);
    c->launches++;
    k_fq_len_to_u64<<<fq_grid(n_rec + 1, 256), 256, 0, s>>>(x.len.p, n_rec, x.off.p);
    c->launches++;
    size_t tmp = 0;
    SFB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp, x.off.p, x.off.p, (int)(n_rec + 1), s));
    SFB_CUDA(c, st->tmp.reserve(tmp));
    SFB_CUDA(c, cub::DeviceScan::ExclusiveSum(st->tmp.p, tmp, x.off.p, x.off.p, (int)(n_rec + 1), s));
    c->launches++;
    uint64_t tail[2] = {0, 0};                                               // total bases, end of the last record
    SFB_CUDA(c, cudaMemcpyAsync(&tail[0], x.off.p + n_rec, 8, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaMemcpyAsync(&tail[1], x.rec_end.p + (n_rec - 1), 8, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    SFB_CUDA(c, x.bases.reserve(tail[0] + 8));
    k_fq_copy<<<fq_grid(n_rec * 32, 256), 256, 0, s>>>(x.text.p, x.seq_start.p, x.len.p, x.off.p, n_rec, x.bases.p);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    *consumed = tail[1];
    return SFB200_OK;
}

}  // namespace

void sfb_fastq_free(sfb200_ctx* c) {
    FqState* st = static_cast<FqState*>(c->fastq);
    if (!st) return;
    st->release();
    delete st;
    c->fastq = nullptr;
}

extern "C" int sfb200_map_fastq(sfb200_ctx* c, const char* text1, uint64_t n1, const char* text2, uint64_t n2, uint64_t max_records,
                                uint64_t* n_records, uint64_t* consumed1, uint64_t* consumed2) {
    if (!c || !text1 || !n_records || !consumed1 || (text2 && !consumed2)) return SFB200_EINVAL;
    if (!c->map) SFB_FAIL(c, SFB200_EINVAL, "map_fastq: call map_begin first");
    cudaSetDevice(c->device);
    if (!c->fastq) c->fastq = new FqState();
    FqState* st = static_cast<FqState*>(c->fastq);
    if (!st->xs) {
        SFB_CUDA(c, cudaStreamCreateWithFlags(&st->xs, cudaStreamNonBlocking));
        SFB_CUDA(c, cudaEventCreateWithFlags(&st->extracted, cudaEventDisableTiming));
        for (cudaEvent_t& e : st->mapped) SFB_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t s = st->xs;
    const bool paired = text2 != nullptr;
    *n_records = 0; *consumed1 = 0;
    if (consumed2) *consumed2 = 0;
    const unsigned set = st->parity;
    st->parity ^= 1u;
    FqMate* m = st->sets[set];
    // this staging set was last read by the mapping kernels of the call before the previous one
    if (st->used[set]) SFB_CUDA(c, cudaEventSynchronize(st->mapped[set]));
    SFB_CUDA(c, st->err.reserve(1));
    SFB_CUDA(c, cudaMemsetAsync(st->err.p, 0, 4, s));
    { const int rc = fq_stage(c, st, m[0], text1, n1); if (rc) return rc; }
    if (paired) { const int rc = fq_stage(c, st, m[1], text2, n2); if (rc) return rc; }
    // FASTQ ('@': four lines per record) or FASTA reads ('>': two); both mates in the same format
    const uint32_t lpr = (n1 > 0 && text1[0] == '>') ? 2u : 4u;
    uint64_t n_rec = m[0].n_newlines / lpr;
    if (paired) n_rec = std::min<uint64_t>(n_rec, m[1].n_newlines / lpr);
    if (max_records) n_rec = std::min<uint64_t>(n_rec, max_records);
    if (n_rec == 0) return SFB200_OK;
    { const int rc = fq_extract(c, st, m[0], n_rec, lpr, consumed1); if (rc) return rc; }
    if (paired) { const int rc = fq_extract(c, st, m[1], n_rec, lpr, consumed2); if (rc) return rc; }
    uint32_t err = 0;
    SFB_CUDA(c, cudaMemcpyAsync(&err, st->err.p, 4, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(c, cudaEventRecord(st->extracted, s));
    SFB_CUDA(c, cudaStreamSynchronize(s));
    if (err & FQ_ERR_HEADER) SFB_FAIL(c, SFB200_EINVAL, "map_fastq: a record does not start with '@' / '>' (four-line FASTQ or two-line FASTA records expected, both mates alike)");
    if (err & FQ_ERR_PLUS) SFB_FAIL(c, SFB200_EINVAL, "map_fastq: the line after a sequence does not start with '+' (four-line FASTQ records expected)");
    if (err & FQ_ERR_LONG) SFB_FAIL(c, SFB200_EINVAL, "map_fastq: sequence line longer than 16 M bases");
    // the mapping kernels run on the context's stream, behind those of the previous block; the caller's text is already on the device,
    // so the call returns without waiting for them
    SFB_CUDA(c, cudaStreamWaitEvent(c->stream, st->extracted, 0));
    const int rc = sfb200_map_batch_device(c, m[0].bases.p, m[0].off.p, paired ? m[1].bases.p : nullptr, paired ? m[1].off.p : nullptr, n_rec);
    if (rc) return rc;
    SFB_CUDA(c, cudaEventRecord(st->mapped[set], c->stream));
    st->used[set] = true;
    *n_records = n_rec;
    return SFB200_OK;
}
