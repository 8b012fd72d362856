// fastq_core.inl -- FASTQ text -> read sequences ON THE DEVICE (the host parser is the end-to-end limit, the GPU maps
// ~40x more text per second than eight host threads can index and copy).  The text of a block of four-line records goes to the GPU
// as it is in the file; three passes find the sequence lines and copy them into the contiguous `bases` / `off` form
// sfb200_map_batch_device takes (fastq.cu).  The per-chunk bodies below are written against two macros so that the same text is
// device code and plain host code for the CPU check (tests/fastq_core_test.cpp):
//   SFB_FQ            function qualifiers
//   SFB_FQ_OR(p, v)   atomic OR on a uint32_t
//   SFB_FQ_LD8(p)     the 8 bytes at p as a little-endian u64 (p is 8-byte aligned on the device: chunks start at multiples of 512)
//   SFB_FQ_POPC(x) / SFB_FQ_CTZ(x)   population count / index of the lowest set bit of a non-zero u64
//
// A block starts at a record boundary.  Newline j (0-based) of the block ends line j; record r owns lines 4r .. 4r+3 (header,
// sequence, '+', qualities -- the form the reference's parser reads, include/PairSequenceParser.hpp, and every sequencer writes),
// or, for FASTA reads, lines 2r and 2r+1 ('>' header, the sequence on ONE line; wrapped sequences are refused).
// Only complete records are extracted: the caller carries the rest of the text over to the front of its next block.

constexpr uint32_t FQ_CHUNK = 512;                    // bytes of text per thread
constexpr uint32_t FQ_ERR_HEADER = 1, FQ_ERR_PLUS = 2, FQ_ERR_LONG = 4;

// 0x80 in every byte of v that is '\n', 0 elsewhere (exact: no carries between bytes), so the text is scanned 8 bytes at a time
SFB_FQ uint64_t fq_newline_mask(uint64_t v) {
    v ^= 0x0a0a0a0a0a0a0a0aULL;
    const uint64_t t = (v & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL;
    return ~(t | v | 0x7f7f7f7f7f7f7f7fULL);
}

// pass 1: newlines in chunk c of text[0, n)
SFB_FQ uint32_t fq_count_newlines(const char* __restrict__ text, uint64_t n, uint64_t c) {
    const uint64_t a = c * FQ_CHUNK, b = a + FQ_CHUNK < n ? a + FQ_CHUNK : n;
    uint32_t k = 0;
    uint64_t p = a;
    for (; p + 8 <= b; p += 8) k += (uint32_t)SFB_FQ_POPC(fq_newline_mask(SFB_FQ_LD8(text + p)));
    for (; p < b; ++p) k += text[p] == '\n';
    return k;
}

// pass 2: chunk c, whose first newline has index nl_base; records [0, n_rec) are wanted.
// seq_start[r] = first base, seq_len[r] = bases (a '\r' before the newline is not one), rec_end[r] = one past the record's last newline.
// seq_len is written by the thread that sees the END of the sequence line and needs its start: the start is at most max_line bytes
// back, found by scanning for the previous newline (the header's) -- records do not straddle threads otherwise.
// lines_per_rec = 4 (FASTQ, header character '@') or 2 (FASTA, '>').
SFB_FQ void fq_mark_chunk(const char* __restrict__ text, uint64_t n, uint64_t c, uint64_t nl_base, uint64_t n_rec, uint32_t lines_per_rec,
                          uint64_t* __restrict__ seq_start, uint32_t* __restrict__ seq_len, uint64_t* __restrict__ rec_end, uint32_t* __restrict__ err) {
    const uint64_t a = c * FQ_CHUNK, b = a + FQ_CHUNK < n ? a + FQ_CHUNK : n;
    const bool fastq = lines_per_rec == 4;
    const char hdr = fastq ? '@' : '>';
    uint64_t j = nl_base;
    if (c == 0 && n_rec > 0 && text[0] != hdr) SFB_FQ_OR(err, FQ_ERR_HEADER);
    uint64_t word = 0, wbase = a;                                    // pending newline bits of the 8 bytes at wbase
    uint64_t q = a;                                                  // next byte not yet loaded
    for (;;) {
        uint64_t p;
        if (word) {
            p = wbase + (SFB_FQ_CTZ(word) >> 3);
            word &= word - 1;
        } else if (q + 8 <= b) {
            word = fq_newline_mask(SFB_FQ_LD8(text + q)); wbase = q; q += 8;
            continue;
        } else if (q < b) {
            p = q++;
            if (text[p] != '\n') continue;
        } else {
            return;
        }
        const uint64_t r = fastq ? j >> 2 : j >> 1;
        const uint32_t k = (uint32_t)(fastq ? j & 3 : j & 1);
        ++j;
        if (r >= n_rec) return;
        if (k == 0) {
            seq_start[r] = p + 1;
        } else if (k == 1) {
            uint64_t s = p;                                          // back to the newline that ended the header line
            while (s > 0 && text[s - 1] != '\n') --s;
            const uint64_t e = (p > s && text[p - 1] == '\r') ? p - 1 : p;
            if (e - s > 0xFFFFFFu) SFB_FQ_OR(err, FQ_ERR_LONG);
            seq_len[r] = (uint32_t)(e - s);
            if (fastq) {
                if (p + 1 < n && text[p + 1] != '+') SFB_FQ_OR(err, FQ_ERR_PLUS);
            } else {
                rec_end[r] = p + 1;
                if (r + 1 < n_rec && text[p + 1] != '>') SFB_FQ_OR(err, FQ_ERR_HEADER);  // a wrapped sequence (or anything else) is not read as bases
            }
        } else if (k == 3) {
            rec_end[r] = p + 1;
            if (r + 1 < n_rec && text[p + 1] != '@') SFB_FQ_OR(err, FQ_ERR_HEADER);   // the next wanted record's header
        }
    }
}

// pass 3: bases of record r, bytes [i0, i0 + step, ...) -- a warp copies one record, lane = i0, step = 32
SFB_FQ void fq_copy_record(const char* __restrict__ text, uint64_t start, uint32_t len, char* __restrict__ out, uint32_t i0, uint32_t step) {
    for (uint32_t i = i0; i < len; i += step) out[i] = text[start + i];
}
