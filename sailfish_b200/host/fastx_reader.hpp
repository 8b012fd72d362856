// fastx_reader.hpp -- read ingestion for the quantification driver.
//
// Replaces, for this path, what the reference gets from Jellyfish's stream_manager + whole_sequence_parser and its own
// PairSequenceParser (include/PairSequenceParser.hpp:28-191; call sites src/SailfishQuantify.cpp:882-898,996-1005): FASTA /
// FASTQ text (plain or gzip, through zlib) -> batches of reads as ONE contiguous byte array + offsets, which is the layout
// sfb200_map_batch takes (a parser job's std::string per mate, concatenated).  Qualities and names are dropped, as
// processReadsQuasi only ever touches `seq` (SailfishQuantify.cpp:192-202,526-528).
//
// FASTQ is parsed block-wise: a block of text (2 MB per parser thread) is cut at a record boundary (line count multiple of 4;
// the previous block ended on one, so no guessing); the threads of a small persistent pool each pread their slice of the
// block (plain files; gzip goes through zlib on one thread), index the newlines of their slice with memchr, and -- after a
// serial prefix sum over the sequence-line lengths -- copy the sequence lines of their share of the records straight into the
// caller's batch.  Buffers are plain malloc'ed memory reused from block to block and from batch to batch (no zero-filling, no
// page faults after warm-up).  FASTA reads (may be multi-line) take a simple serial path.  The driver parses the two mate
// files side by side and one batch ahead of the device.
// Measured in the build container (2.7 GB of FASTQ from tmpfs): 0.43 GB/s for the first version (32 MB blocks, std::vector
// buffers), 1.7 GB/s on one thread, 2.4 GB/s on eight -- that container moves ~5.3 GB/s through memchr however many threads
// ask (microbenchmark), so the scaling of the parallel phases has to be measured on the GPU host.
// Header-only, C++11, needs -lz -pthread.
#ifndef SFB200_FASTX_READER_HPP
#define SFB200_FASTX_READER_HPP

#include <fcntl.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace sfb200 {

// growable byte / word buffers without value-initialisation (std::vector::resize would memset every block)
template <typename T>
class RawBuf {
public:
    RawBuf() {}
    ~RawBuf() { std::free(p_); }
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    void clear() { n_ = 0; }
    void reserve(size_t cap) {
        if (cap <= cap_) return;
        size_t c = cap_ ? cap_ : 1024;
        while (c < cap) c += c / 2 + 1024;
        T* q = static_cast<T*>(std::realloc(p_, c * sizeof(T)));
        if (!q) throw std::bad_alloc();
        p_ = q; cap_ = c;
    }
    void resize_uninit(size_t n) { reserve(n); n_ = n; }
    void push_back(const T& v) { reserve(n_ + 1); p_[n_++] = v; }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }
    T& back() { return p_[n_ - 1]; }
    T* begin() { return p_; }
    T* end() { return p_ + n_; }
    const T* begin() const { return p_; }
    const T* end() const { return p_ + n_; }

private:
    T* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

// a few persistent threads that run `fn(task)` for task = 0..n-1 and wait: the parser's parallel phases are short (a few MB of
// text each), so spawning threads per phase costs more than the phase
class WorkerPool {
public:
    explicit WorkerPool(unsigned n_threads) {
        for (unsigned t = 1; t < n_threads; ++t) th_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    unsigned size() const { return (unsigned)th_.size() + 1; }
    // fn must not throw
    void run(unsigned n_tasks, const std::function<void(unsigned)>& fn) {
        if (n_tasks == 0) return;
        if (th_.empty() || n_tasks == 1) { for (unsigned i = 0; i < n_tasks; ++i) fn(i); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_tasks_ = n_tasks; next_ = 0; pending_ = n_tasks; ++gen_;
        }
        cv_.notify_all();
        work();                                                   // the calling thread takes tasks too
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            unsigned i;
            const std::function<void(unsigned)>* f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= n_tasks_) return;
                i = next_++; f = fn_;
            }
            (*f)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned)>* fn_ = nullptr;
    unsigned n_tasks_ = 0, next_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// a batch of reads: read i = bases[off[i] .. off[i+1])
struct ReadBatch {
    RawBuf<char> bases;
    RawBuf<uint64_t> off;
    size_t size() const { return off.empty() ? 0 : off.size() - 1; }
    void clear() { bases.clear(); off.clear(); off.push_back(0); }
};

class FastxReader {
public:
    // block_bytes = 0: 2 MB per thread (each thread's share of a block stays in its cache between the newline index and the copy)
    explicit FastxReader(const std::string& path, unsigned threads = 4, size_t block_bytes = 0)
        : path_(path), threads_(threads ? threads : 1), block_(block_bytes == 0 ? (size_t)(threads ? threads : 1) * (2u << 20) : (block_bytes < 16 ? 16 : block_bytes)),
          pool_(threads ? threads : 1) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("cannot open " + path);
        unsigned char magic[2] = {0, 0};
        const ssize_t got = ::pread(fd_, magic, 2, 0);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {              // gzip: hand the descriptor to zlib
            gz_ = gzdopen(fd_, "rb");
            if (!gz_) { ::close(fd_); throw std::runtime_error("cannot open " + path); }
            gzbuffer(gz_, 1u << 20);
        }
    }
    ~FastxReader() { if (gz_) gzclose(gz_); else if (fd_ >= 0) ::close(fd_); }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    // Appends up to max_records reads to `out` (which the caller cleared); returns how many.  0 = end of file.
    size_t next(ReadBatch& out, size_t max_records) {
        if (out.off.empty()) out.off.push_back(0);
        size_t got = 0;
        while (got < max_records) {
            if (rec_pos_ == n_rec_) {
                if (!refill()) break;
                continue;
            }
            const size_t take = std::min(max_records - got, n_rec_ - rec_pos_);
            if (fmt_ == 'q') emit_fastq(out, rec_pos_, take); else emit_fasta(out, rec_pos_, take);
            rec_pos_ += take; got += take;
        }
        return got;
    }
    uint64_t records_read() const { return n_records_; }
    // seconds spent in: [0] reading, [1] newline index, [2] lengths + offsets, [3] copying bases, [4] moving the tail
    const double* phase_seconds() const { return tm_; }

private:
    size_t read_some(char* dst, size_t want) {
        if (!gz_ && threads_ > 1 && want >= (8u << 20)) {             // plain file, big block: every thread preads its slice
            const unsigned nt = threads_;
            std::vector<long> got(nt, 0);
            const uint64_t pos = file_pos_;
            pool_.run(nt, [&](unsigned t) {
                const size_t a = want * t / nt, b = want * (t + 1) / nt;
                size_t n = 0;
                while (n < b - a) {
                    const ssize_t r = ::pread(fd_, dst + a + n, b - a - n, (off_t)(pos + a + n));
                    if (r < 0) { got[t] = -1; return; }
                    if (r == 0) break;
                    n += (size_t)r;
                }
                got[t] = (long)n;
            });
            size_t n = 0;
            for (unsigned t = 0; t < nt; ++t) {
                if (got[t] < 0) throw std::runtime_error("read error in " + path_);
                n += (size_t)got[t];
                if ((size_t)got[t] < want * (t + 1) / nt - want * t / nt) { eof_ = true; break; }   // a short slice: end of file
            }
            file_pos_ += n;
            return n;
        }
        size_t n = 0;
        while (n < want) {
            long r;
            if (gz_) r = gzread(gz_, dst + n, (unsigned)std::min<size_t>(want - n, 1u << 30));
            else r = (long)::pread(fd_, dst + n, std::min<size_t>(want - n, 1u << 30), (off_t)(file_pos_ + n));
            if (r < 0) throw std::runtime_error("read error in " + path_);
            if (r == 0) { eof_ = true; break; }
            n += (size_t)r;
        }
        file_pos_ += n;
        return n;
    }

    // make the next block of complete records available (index built, nothing copied yet); false at end of input
    bool refill() {
        // drop what the previous block's records covered; keep the tail (an incomplete record) at the front
        if (consumed_ > 0) {
            const double t0 = now();
            const size_t tail = buf_.size() - consumed_;
            if (tail) std::memmove(buf_.data(), buf_.data() + consumed_, tail);
            buf_.resize_uninit(tail);
            consumed_ = 0;
            tm_[4] += now() - t0;
        }
        n_rec_ = 0; rec_pos_ = 0;
        for (;;) {
            if (eof_ && buf_.empty()) return false;
            if (!eof_) {
                const size_t old = buf_.size();
                buf_.reserve(old + block_);
                const double t0 = now();
                const size_t n = read_some(buf_.data() + old, block_);
                buf_.resize_uninit(old + n);
                tm_[0] += now() - t0;
            }
            if (buf_.empty()) return false;
            if (fmt_ == 0) {
                size_t i = 0;
                while (i < buf_.size() && (buf_[i] == '\n' || buf_[i] == '\r')) ++i;
                if (i == buf_.size()) { buf_.clear(); continue; }
                if (buf_[i] == '@') fmt_ = 'q'; else if (buf_[i] == '>') fmt_ = 'a';
                else throw std::runtime_error(path_ + ": neither FASTA nor FASTQ");
                if (i) { std::memmove(buf_.data(), buf_.data() + i, buf_.size() - i); buf_.resize_uninit(buf_.size() - i); }
            }
            { const double t0 = now(); if (fmt_ == 'q') index_fastq(); else index_fasta(); tm_[1] += now() - t0; }
            if (n_rec_ > 0) break;
            if (eof_) {
                for (size_t i = 0; i < buf_.size(); ++i)
                    if (buf_[i] != '\n' && buf_[i] != '\r' && buf_[i] != ' ') throw std::runtime_error(path_ + ": truncated record at end of file");
                buf_.clear();
                return false;
            }
        }
        n_records_ += n_rec_;
        return true;
    }

    // ---- FASTQ, four lines per record (the form every sequencer and the reference's test data use) ----------------------
    // ls_[i] = start of line i plus one sentinel, so that ls_[i + 1] - 1 is one past the last character of line i
    void index_fastq() {
        const char* p = buf_.data();
        const size_t len = buf_.size();
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, len >> 20));
        if (parts_.size() < nt) parts_.resize(nt);
        pool_.run(nt, [&](unsigned t) {
            std::vector<size_t>& v = parts_[t];
            v.clear();
            size_t pos = len * t / nt;
            const size_t end = len * (t + 1) / nt;
            // a line starts after every newline.  Lines are ~80 bytes: one memchr call per line costs more than the scan itself
            // (0.17 s per 0.67 GB on eight threads); 16 bytes at a time with SSE2 the newline positions fall out of a bit mask
#if defined(__SSE2__)
            const __m128i nlv = _mm_set1_epi8('\n');
            while (pos + 16 <= end) {
                unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + pos)), nlv));
                while (m) { v.push_back(pos + (size_t)__builtin_ctz(m) + 1); m &= m - 1; }
                pos += 16;
            }
#endif
            while (pos < end) {
                const void* nl = std::memchr(p + pos, '\n', end - pos);
                if (!nl) break;
                pos = (size_t)((const char*)nl - p) + 1;
                v.push_back(pos);
            }
        });
        size_t total = 1;
        std::vector<size_t> part_off(nt);
        for (unsigned t = 0; t < nt; ++t) { part_off[t] = total; total += parts_[t].size(); }
        ls_.resize_uninit(total + 1);
        ls_[0] = 0;
        pool_.run(nt, [&](unsigned t) { if (!parts_[t].empty()) std::memcpy(ls_.data() + part_off[t], parts_[t].data(), parts_[t].size() * sizeof(size_t)); });
        const size_t k = total;
        // the last entry is `len` if the text ended with a newline (then it starts no line); otherwise the last line is unterminated
        const bool terminated = ls_[k - 1] == len && k > 1;
        size_t n_lines;
        if (terminated) n_lines = k - 1;                       // ls_[k-1] == len is already the sentinel
        else { n_lines = eof_ ? k : k - 1; ls_[k] = len + 1; } // an unterminated last line counts only when the input is exhausted
        n_rec_ = n_lines / 4;
        consumed_ = n_rec_ == 0 ? 0 : std::min(ls_[4 * n_rec_], len);
    }
    bool record_ok(const char* p, size_t r) const { return p[ls_[4 * r]] == '@' && p[ls_[4 * r + 2]] == '+'; }
    size_t seq_end(const char* p, size_t r) const {            // one past the last base of record r (no '\n', no '\r')
        size_t e = ls_[4 * r + 2] - 1;
        if (e > ls_[4 * r + 1] && p[e - 1] == '\r') --e;
        return e;
    }
    void emit_fastq(ReadBatch& out, size_t r0, size_t n) {
        const char* p = buf_.data();
        const size_t o0 = out.off.size() - 1;                  // reads already in the batch
        const uint64_t base = out.off[o0];
        out.off.resize_uninit(o0 + 1 + n);
        uint64_t* off = out.off.data() + o0;
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, n >> 16));
        // lengths in parallel, prefix sum serial (a few hundred microseconds per million reads), copy in parallel
        std::atomic<size_t> bad(SIZE_MAX);                     // worker threads must not throw: remember the first bad record
        auto lens = [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                if (!record_ok(p, r0 + i)) { size_t cur = bad.load(); while (r0 + i < cur && !bad.compare_exchange_weak(cur, r0 + i)) { } }
                off[i + 1] = seq_end(p, r0 + i) - ls_[4 * (r0 + i) + 1];
            }
        };
        auto copy = [&](size_t a, size_t b) {
            char* dst = out.bases.data();
            for (size_t i = a; i < b; ++i) std::memcpy(dst + off[i], p + ls_[4 * (r0 + i) + 1], off[i + 1] - off[i]);
        };
        const double t0 = now();
        run_parallel(pool_, nt, n, lens);
        if (bad.load() != SIZE_MAX)
            throw std::runtime_error(path_ + ": malformed FASTQ record " + std::to_string(n_records_ - n_rec_ + bad.load()) + " (multi-line FASTQ is not supported)");
        uint64_t acc = base;
        for (size_t i = 0; i < n; ++i) { const uint64_t l = off[i + 1]; off[i + 1] = acc + l; acc += l; }
        out.bases.resize_uninit(acc);
        const double t1 = now();
        run_parallel(pool_, nt, n, copy);
        tm_[2] += t1 - t0; tm_[3] += now() - t1;
    }
    template <typename F>
    static void run_parallel(WorkerPool& pool, unsigned nt, size_t n, F f) {
        if (nt <= 1) { f(0, n); return; }
        pool.run(nt, [&](unsigned t) { f(n * t / nt, n * (t + 1) / nt); });
    }

    // ---- FASTA: '>' header line, then sequence lines up to the next header; a record is complete when the next header (or
    //      the end of the input) has been seen.  fa_[2r], fa_[2r+1] = first byte after the header line, start of the next header
    void index_fasta() {
        const char* p = buf_.data();
        const size_t len = buf_.size();
        fa_.clear();
        size_t pos = 0, consumed = 0;
        while (pos < len) {
            size_t q = pos, next_hdr = len;
            bool complete = false;
            for (;;) {
                const void* nl = std::memchr(p + q, '\n', len - q);
                if (!nl) { complete = eof_; break; }
                q = (size_t)((const char*)nl - p) + 1;
                if (q < len && p[q] == '>') { complete = true; next_hdr = q; break; }
                if (q >= len) { complete = eof_; break; }
            }
            if (!complete) break;
            const void* h_end = std::memchr(p + pos, '\n', next_hdr - pos);
            fa_.push_back(h_end ? (size_t)((const char*)h_end - p) + 1 : next_hdr);
            fa_.push_back(next_hdr);
            pos = next_hdr; consumed = next_hdr;
        }
        n_rec_ = fa_.size() / 2;
        consumed_ = consumed;
    }
    void emit_fasta(ReadBatch& out, size_t r0, size_t n) {
        const char* p = buf_.data();
        for (size_t r = r0; r < r0 + n; ++r) {
            for (size_t s = fa_[2 * r]; s < fa_[2 * r + 1]; ++s) { const char ch = p[s]; if (ch != '\n' && ch != '\r') out.bases.push_back(ch); }
            out.off.push_back(out.bases.size());
        }
    }

    std::string path_;
    unsigned threads_;
    size_t block_;
    WorkerPool pool_;
    int fd_ = -1;
    gzFile gz_ = nullptr;
    uint64_t file_pos_ = 0;           // plain files: offset of the next unread byte
    double tm_[5] = {0, 0, 0, 0, 0};
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    bool eof_ = false;
    char fmt_ = 0;                 // 'q' FASTQ, 'a' FASTA
    RawBuf<char> buf_;             // text: [0, consumed_) is covered by the indexed records, the rest is the next block's head
    size_t consumed_ = 0;
    RawBuf<size_t> ls_;            // FASTQ line starts of the current block
    std::vector<std::vector<size_t>> parts_;
    std::vector<size_t> fa_;       // FASTA record extents of the current block
    size_t n_rec_ = 0, rec_pos_ = 0;
    uint64_t n_records_ = 0;
};

// whole-file FASTA of the transcripts: name = header up to the first white space (what RapMap's indexer keeps), sequence
// lines concatenated
inline void read_transcripts(const std::string& path, std::vector<std::string>& names, std::string& seq, std::vector<uint64_t>& off,
                             std::vector<uint32_t>& lens) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1u << 20);
    std::vector<char> line(1u << 16);
    names.clear(); seq.clear(); off.clear(); lens.clear();
    std::string cur_line;
    bool in_record = false;
    auto close_record = [&]() { if (in_record) lens.push_back((uint32_t)(seq.size() - off.back())); };
    while (gzgets(f, line.data(), (int)line.size())) {
        cur_line.assign(line.data());
        while (!cur_line.empty() && cur_line.back() != '\n' && gzgets(f, line.data(), (int)line.size())) cur_line.append(line.data());
        while (!cur_line.empty() && (cur_line.back() == '\n' || cur_line.back() == '\r')) cur_line.pop_back();
        if (cur_line.empty()) continue;
        if (cur_line[0] == '>') {
            close_record();
            size_t e = 1;
            while (e < cur_line.size() && cur_line[e] != ' ' && cur_line[e] != '\t') ++e;
            names.push_back(cur_line.substr(1, e - 1));
            off.push_back(seq.size());
            in_record = true;
        } else if (in_record) {
            seq.append(cur_line);
        } else {
            gzclose(f);
            throw std::runtime_error(path + ": sequence before the first FASTA header");
        }
    }
    close_record();
    gzclose(f);
}

}  // namespace sfb200
#endif
