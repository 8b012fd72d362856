// fastx_reader.hpp -- read ingestion for the quantification driver.
//
// Replaces, for this path, what the reference gets from Jellyfish's stream_manager + whole_sequence_parser and its own
// PairSequenceParser (include/PairSequenceParser.hpp:28-191; call sites src/SailfishQuantify.cpp:882-898,996-1005): FASTA /
// FASTQ text (plain or gzip, through zlib) -> batches of reads as ONE contiguous byte array + offsets, which is the layout
// sfb200_map_batch takes (a parser job's std::string per mate, concatenated).  Qualities and names are dropped, as
// processReadsQuasi only ever touches `seq` (SailfishQuantify.cpp:192-202,526-528).
//
// FASTQ is parsed block-wise: a block is cut at a record boundary (line count multiple of 4), line starts are found with
// memchr, offsets are a prefix sum, and the sequence lines are copied into the batch by several threads.  FASTA reads (may be
// multi-line) take a simple serial path.  Header-only, C++11, needs -lz -pthread.
#ifndef SFB200_FASTX_READER_HPP
#define SFB200_FASTX_READER_HPP

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace sfb200 {

// a batch of reads: read i = bases[off[i] .. off[i+1])
struct ReadBatch {
    std::vector<char> bases;
    std::vector<uint64_t> off;
    size_t size() const { return off.empty() ? 0 : off.size() - 1; }
    void clear() { bases.clear(); off.assign(1, 0); }
};

class FastxReader {
public:
    explicit FastxReader(const std::string& path, unsigned copy_threads = 4, size_t block_bytes = 32u << 20)
        : path_(path), threads_(copy_threads ? copy_threads : 1), block_(block_bytes) {
        f_ = gzopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("cannot open " + path);
        gzbuffer(f_, 1u << 20);
        stage_.clear();
    }
    ~FastxReader() { if (f_) gzclose(f_); }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    // Appends up to max_records reads to `out` (which the caller cleared); returns how many.  0 = end of file.
    size_t next(ReadBatch& out, size_t max_records) {
        if (out.off.empty()) out.off.assign(1, 0);
        size_t got = 0;
        while (got < max_records) {
            if (stage_pos_ == stage_.size()) {
                if (!refill()) break;
                continue;
            }
            const size_t take = std::min(max_records - got, stage_.size() - stage_pos_);
            const uint64_t b0 = stage_.off[stage_pos_], b1 = stage_.off[stage_pos_ + take];
            const uint64_t base = out.off.back();
            out.bases.insert(out.bases.end(), stage_.bases.begin() + b0, stage_.bases.begin() + b1);
            for (size_t i = 1; i <= take; ++i) out.off.push_back(base + (stage_.off[stage_pos_ + i] - b0));
            stage_pos_ += take; got += take;
        }
        return got;
    }
    uint64_t records_read() const { return n_records_; }

private:
    // read the next block, parse every complete record in it into stage_; false at end of input
    bool refill() {
        stage_.clear(); stage_pos_ = 0;
        while (stage_.size() == 0) {
            if (eof_ && buf_.empty()) return false;
            if (!eof_) {
                const size_t old = buf_.size();
                buf_.resize(old + block_);
                size_t n = 0;
                while (n < block_) {                       // gzread takes an unsigned count
                    const int r = gzread(f_, buf_.data() + old + n, (unsigned)std::min<size_t>(block_ - n, 1u << 30));
                    if (r < 0) throw std::runtime_error("read error in " + path_);
                    if (r == 0) { eof_ = true; break; }
                    n += (size_t)r;
                }
                buf_.resize(old + n);
            }
            if (buf_.empty()) return false;
            if (fmt_ == 0) {
                size_t i = 0;
                while (i < buf_.size() && (buf_[i] == '\n' || buf_[i] == '\r')) ++i;
                if (i == buf_.size()) { buf_.clear(); continue; }
                if (buf_[i] == '@') fmt_ = 'q'; else if (buf_[i] == '>') fmt_ = 'a';
                else throw std::runtime_error(path_ + ": neither FASTA nor FASTQ");
                buf_.erase(buf_.begin(), buf_.begin() + i);
            }
            const size_t used = fmt_ == 'q' ? parse_fastq() : parse_fasta();
            buf_.erase(buf_.begin(), buf_.begin() + used);
            if (eof_ && used == 0 && stage_.size() == 0) {
                if (!buf_.empty()) {
                    bool blank = true;
                    for (char ch : buf_) if (ch != '\n' && ch != '\r' && ch != ' ') { blank = false; break; }
                    if (!blank) throw std::runtime_error(path_ + ": truncated record at end of file");
                }
                buf_.clear();
                return false;
            }
        }
        n_records_ += stage_.size();
        return true;
    }

    // FASTQ, four lines per record (the form every sequencer and the reference's test data use)
    size_t parse_fastq() {
        const char* p = buf_.data();
        const size_t len = buf_.size();
        ls_.clear();
        size_t pos = 0;
        while (pos < len) {
            ls_.push_back(pos);
            const void* nl = std::memchr(p + pos, '\n', len - pos);
            if (!nl) { pos = len + 1; break; }              // unterminated last line
            pos = (size_t)((const char*)nl - p) + 1;
        }
        // ls_[i] = start of line i, plus a sentinel so that ls_[i + 1] - 1 is always one past line i's last character;
        // the last line is complete if it ended with '\n' or the input is exhausted
        const bool last_unterminated = pos == len + 1;
        size_t n_lines = ls_.size();
        ls_.push_back(last_unterminated ? len + 1 : len);
        if (last_unterminated && !eof_) n_lines -= 1;
        const size_t n_rec = n_lines / 4;
        if (n_rec == 0) return 0;
        auto line_end = [&](size_t i) -> size_t {            // one past the last character of line i (no '\n', no '\r')
            size_t e = ls_[i + 1] - 1;
            if (e > ls_[i] && p[e - 1] == '\r') --e;
            return e;
        };
        stage_.off.resize(n_rec + 1);
        stage_.off[0] = 0;
        for (size_t r = 0; r < n_rec; ++r) {
            if (p[ls_[4 * r]] != '@' || p[ls_[4 * r + 2]] != '+')
                throw std::runtime_error(path_ + ": malformed FASTQ record " + std::to_string(n_records_ + r) + " (multi-line FASTQ is not supported)");
            stage_.off[r + 1] = stage_.off[r] + (line_end(4 * r + 1) - ls_[4 * r + 1]);
        }
        stage_.bases.resize(stage_.off[n_rec]);
        const unsigned nt = (unsigned)std::min<size_t>(threads_, (n_rec + 65535) / 65536);
        auto copy_range = [&](size_t a, size_t b) {
            for (size_t r = a; r < b; ++r) std::memcpy(stage_.bases.data() + stage_.off[r], p + ls_[4 * r + 1], stage_.off[r + 1] - stage_.off[r]);
        };
        if (nt <= 1) copy_range(0, n_rec);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(copy_range, n_rec * t / nt, n_rec * (t + 1) / nt);
            for (auto& x : th) x.join();
        }
        return std::min(ls_[4 * n_rec], len);
    }

    // FASTA: '>' header line, then sequence lines up to the next header; a record is complete when the next header (or the
    // end of the input) has been seen
    size_t parse_fasta() {
        const char* p = buf_.data();
        const size_t len = buf_.size();
        size_t pos = 0, consumed = 0;
        stage_.off.assign(1, 0);
        while (pos < len) {
            // pos is at a '>' ; find the end of the record
            size_t q = pos;
            bool complete = false;
            size_t next_hdr = len;
            while (true) {
                const void* nl = std::memchr(p + q, '\n', len - q);
                if (!nl) { complete = eof_; next_hdr = len; break; }
                q = (size_t)((const char*)nl - p) + 1;
                if (q < len && p[q] == '>') { complete = true; next_hdr = q; break; }
                if (q >= len) { complete = eof_; next_hdr = len; break; }
            }
            if (!complete) break;
            const void* h_end = std::memchr(p + pos, '\n', next_hdr - pos);
            size_t s = h_end ? (size_t)((const char*)h_end - p) + 1 : next_hdr;
            for (; s < next_hdr; ++s) { const char ch = p[s]; if (ch != '\n' && ch != '\r') stage_.bases.push_back(ch); }
            stage_.off.push_back(stage_.bases.size());
            pos = next_hdr; consumed = next_hdr;
        }
        return consumed;
    }

    std::string path_;
    unsigned threads_;
    size_t block_;
    gzFile f_ = nullptr;
    bool eof_ = false;
    char fmt_ = 0;                 // 'q' FASTQ, 'a' FASTA
    std::vector<char> buf_;
    std::vector<size_t> ls_;
    ReadBatch stage_;
    size_t stage_pos_ = 0;
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
