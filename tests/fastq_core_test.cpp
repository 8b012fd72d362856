// CPU check of the device-side FASTQ extraction (sailfish_b200/csrc/fastq_core.inl, the text fastq.cu compiles for the GPU): the three
// passes are replayed serially over random FASTQ text -- LF and CRLF, ragged and empty reads, '@' and '+' as first quality
// characters, blocks cut in the middle of a record, a record limit -- and compared with a plain line-by-line parse.
// Built and run by tests/test_parser_fuzz.py.
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#define SFB_FQ static inline
#define SFB_FQ_OR(p, v) (*(p) |= (v))
static inline uint64_t ld8(const char* p) { uint64_t v; __builtin_memcpy(&v, p, 8); return v; }
#define SFB_FQ_LD8(p) ld8(p)
#define SFB_FQ_POPC(x) __builtin_popcountll(x)
#define SFB_FQ_CTZ(x) __builtin_ctzll(x)
#include "../sailfish_b200/csrc/fastq_core.inl"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

struct Extracted { std::vector<std::string> seqs; uint64_t consumed = 0; uint32_t err = 0; uint64_t n_rec = 0; };

// what fastq.cu does with one mate's block, serially
static Extracted extract(const std::string& text, uint64_t max_records, uint32_t lpr = 4) {
    Extracted out;
    const uint64_t n = text.size(), n_chunks = (n + FQ_CHUNK - 1) / FQ_CHUNK;
    std::vector<uint32_t> cnt(n_chunks + 1, 0);
    for (uint64_t c = 0; c < n_chunks; ++c) cnt[c] = fq_count_newlines(text.data(), n, c);
    uint32_t run = 0;
    for (uint64_t c = 0; c <= n_chunks; ++c) { const uint32_t v = cnt[c]; cnt[c] = run; run += v; }      // exclusive scan
    uint64_t n_rec = cnt[n_chunks] / lpr;
    if (max_records && n_rec > max_records) n_rec = max_records;
    out.n_rec = n_rec;
    if (!n_rec) return out;
    std::vector<uint64_t> seq_start(n_rec, ~0ull), rec_end(n_rec, ~0ull), off(n_rec + 1, 0);
    std::vector<uint32_t> len(n_rec, ~0u);
    // chunks in a scrambled order: the passes must not depend on it
    std::vector<uint64_t> order(n_chunks);
    for (uint64_t c = 0; c < n_chunks; ++c) order[c] = (c * 7919) % n_chunks;
    if (n_chunks % 7919 == 0) for (uint64_t c = 0; c < n_chunks; ++c) order[c] = c;
    for (uint64_t c : order) fq_mark_chunk(text.data(), n, c, cnt[c], n_rec, lpr, seq_start.data(), len.data(), rec_end.data(), &out.err);
    for (uint64_t r = 0; r < n_rec; ++r) off[r + 1] = off[r] + len[r];
    std::string bases(off[n_rec], '?');
    for (uint64_t r = 0; r < n_rec; ++r) for (uint32_t lane = 0; lane < 32; ++lane) fq_copy_record(text.data(), seq_start[r], len[r], &bases[off[r]], lane, 32);
    for (uint64_t r = 0; r < n_rec; ++r) out.seqs.push_back(bases.substr(off[r], len[r]));
    out.consumed = rec_end[n_rec - 1];
    return out;
}

int main() {
    std::mt19937_64 rng(2024);
    const char* alpha = "ACGTN";
    for (int rep = 0; rep < 400; ++rep) {
        const bool crlf = rep % 5 == 1;
        const std::string eol = crlf ? "\r\n" : "\n";
        const size_t n_rec = rep < 8 ? rep : 1 + rng() % 300;
        std::vector<std::string> want;
        std::string text;
        std::vector<uint64_t> ends;
        for (size_t r = 0; r < n_rec; ++r) {
            const size_t L = (rng() % 20 == 0) ? 0 : (rng() % 10 == 0 ? 300 + rng() % 900 : 20 + rng() % 130);
            std::string s(L, 'A'), q(L, 'I');
            for (auto& ch : s) ch = alpha[rng() % 5];
            for (auto& ch : q) ch = (char)(33 + rng() % 60);
            if (L && rng() % 3 == 0) q[0] = '@';                       // a quality line may start like a header ...
            if (L && rng() % 7 == 0) q[0] = '+';                       // ... or like a separator
            text += "@read" + std::to_string(r) + (rng() % 2 ? " extra words" : "") + eol + s + eol + "+" + (rng() % 4 == 0 ? "read" + std::to_string(r) : "") + eol + q + eol;
            want.push_back(s);
            ends.push_back(text.size());
        }
        // the whole text, then prefixes cut in the middle of a record, then with a record limit
        for (int variant = 0; variant < 4; ++variant) {
            std::string blk = text;
            uint64_t limit = 0;
            if (variant == 1 && !text.empty()) blk = text.substr(0, rng() % text.size());
            if (variant == 2 && !text.empty()) blk = text.substr(0, text.size() - 1);                 // last newline missing: that record is incomplete
            if (variant == 3) limit = 1 + rng() % (n_rec + 1);
            size_t complete = 0;
            while (complete < n_rec && ends[complete] <= blk.size()) ++complete;
            if (limit && complete > limit) complete = limit;
            const Extracted got = extract(blk, limit);
            CHECK(got.err == 0, "rep %d variant %d: error flags %u", rep, variant, got.err);
            CHECK(got.n_rec == complete, "rep %d variant %d: %llu records, expected %zu", rep, variant, (unsigned long long)got.n_rec, complete);
            CHECK(got.consumed == (complete ? ends[complete - 1] : 0), "rep %d variant %d: consumed %llu", rep, variant, (unsigned long long)got.consumed);
            for (size_t r = 0; r < complete; ++r) CHECK(got.seqs[r] == want[r], "rep %d variant %d record %zu: '%s' vs '%s'", rep, variant, r, got.seqs[r].c_str(), want[r].c_str());
        }
    }
    // FASTA reads: two lines per record
    for (int rep = 0; rep < 100; ++rep) {
        const std::string eol = rep % 4 == 1 ? "\r\n" : "\n";
        const size_t n_rec = 1 + rng() % 200;
        std::vector<std::string> want;
        std::vector<uint64_t> ends;
        std::string text;
        for (size_t r = 0; r < n_rec; ++r) {
            std::string s(rng() % 15 == 0 ? 0 : 20 + rng() % 400, 'A');
            for (auto& ch : s) ch = alpha[rng() % 5];
            text += ">read" + std::to_string(r) + eol + s + eol;
            want.push_back(s); ends.push_back(text.size());
        }
        const std::string blk = rep % 2 ? text.substr(0, rng() % text.size()) : text;
        size_t complete = 0;
        while (complete < n_rec && ends[complete] <= blk.size()) ++complete;
        const Extracted got = extract(blk, 0, 2);
        CHECK(got.err == 0 && got.n_rec == complete && got.consumed == (complete ? ends[complete - 1] : 0), "FASTA rep %d: err %u, %llu records (expected %zu)", rep, got.err,
              (unsigned long long)got.n_rec, complete);
        for (size_t r = 0; r < complete; ++r) CHECK(got.seqs[r] == want[r], "FASTA rep %d record %zu", rep, r);
    }
    // malformed input is reported, not mis-parsed silently
    {
        const std::string fasta = ">a\nACGT\n>b\nGGCC\n";
        CHECK(extract(fasta, 0).err & FQ_ERR_HEADER, "FASTA text read as FASTQ must raise the header flag");
        const std::string wrapped = ">a\nACGT\nACGT\n>b\nGGCC\nGG\n";
        CHECK(extract(wrapped, 0, 2).err & FQ_ERR_HEADER, "wrapped FASTA sequences are refused");
        CHECK(extract(std::string("@a\nAC\n+\nII\n@b\nAC\n+\nII\n"), 0, 2).err & FQ_ERR_HEADER, "FASTQ text read as FASTA");
        const std::string noplus = "@a\nACGT\nACGT\nIIII\n@b\nAC\n+\nII\n";
        CHECK(extract(noplus, 0).err & FQ_ERR_PLUS, "missing '+' line");
        const std::string shifted = "@a\nACGT\n+\nIIII\n\n@b\nAC\n+\nII\n@c\nA\n+\nI\n";           // a blank line between records
        CHECK(extract(shifted, 0).err != 0, "blank line between records");
    }
    printf("fastq core ok\n");
    return 0;
}
