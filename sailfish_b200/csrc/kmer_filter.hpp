// kmer_filter.hpp -- hashing of the k-mer table and addressing of the presence filter (index.cu builds them, map.cu probes them).
// No CUDA dependency: tests/kmer_filter_test.cpp compiles this file with g++ and measures locality and false-positive rate.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define SFB_HD __host__ __device__ __forceinline__
#else
#define SFB_HD inline
#endif

// ---- k-mer table hashing -----------------------------------------------------------------------------------------------
// The k-mer table and the presence filter are this repo's own structures (RapMap's index is not in the reference
// tree), so their hash is free to choose.  XXH64 of the 8-byte k-mer costs five 64-bit multiplies per lookup -- measured
// at ~40% of the mapping kernel's instructions -- so lookups use a two-multiply mixer; XXH64 stays where the reference
// fixes it: the equivalence-class label hash (TranscriptGroup.cpp:9-12).
SFB_HD uint64_t sfb_kmer_mix(uint64_t x) {
    x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 29;
    return x;
}
// ---- presence filter: a direct-addressed bitmap over m-mers ---------------------------------------------------------------------------
// A read is scanned k-mer by k-mer, and on the strand that does not match every one of its L-k+1 k-mers is absent from the index.
// A k-mer can only be present if every m-mer inside it occurs in the transcriptome, and the LAST m-mer of the k-mer at read position i
// (bases i+k-m .. i+k-1) lies inside the k-mers of positions i .. i+k-m as well: if that one m-mer is absent, k-m+1 positions are
// decided by one bit.  The bitmap has 4^m bits, addressed by the m-mer's 2m bits directly (no hashing): m = 15 -> 128 MB, which a
// transcriptome of a few hundred Mnt fills to ~25%; larger texts take m = 16 (512 MB).  k <= m: m = k (the bitmap is then exact).
SFB_HD int sfb_mfilter_m(int k, uint64_t text_len) {
    const int m = text_len > 600000000ULL ? 16 : 15;
    return m < k ? m : k;
}
SFB_HD uint64_t sfb_mfilter_words(int m) { return ((1ULL << (2 * m)) + 31) / 32; }      // 32-bit words
// m-mer that starts at base `off` of the k-mer km (base i at bits 2i)
SFB_HD uint64_t sfb_mfilter_key(uint64_t km, int off, int m) { return (km >> (2 * off)) & ((1ULL << (2 * m)) - 1); }
