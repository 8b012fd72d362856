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
// Presence filter with locality.  A read is scanned k-mer by k-mer, and on the strand that does not match every one of its ~L-k
// k-mers is looked up and found absent: with a hashed filter word per k-mer that is one random 32-byte sector per position.  Here the
// SECTOR (4 words) is chosen by an anchor the neighbouring k-mers share: the first position i < k-14 of the k-mer where
// base[i] == base[i+1] and base[i+2] is A or C (one position in eight; if there is none, the first doubled base; if none, 0), and the
// 15-mer that starts there.  While the scan moves on by one base the anchor stays where it is until it leaves the window, so the
// k-mers of a read touch ~L/8 sectors instead of L-k.  The word inside the sector and the three bits come from the k-mer's own mix.
// k < 19 has no room for a window: the low bases of the k-mer choose the sector.
struct SfbBloomGeom { uint64_t wmask; int shift; };      // wmask: bit 2i set for anchor positions i; shift = 64 - log2(sectors)
SFB_HD SfbBloomGeom sfb_bloom_geom(int k, uint64_t n_words) {
    SfbBloomGeom g;
    const int w = k >= 19 ? k - 14 : 0;
    g.wmask = w ? (0x5555555555555555ULL & ((1ULL << (2 * w)) - 1)) : 0ULL;
    int lg = 0;
    while ((4ULL << lg) < n_words) ++lg;
    g.shift = 64 - (lg ? lg : 1);                // bloom_words >= 64, so lg >= 4
    return g;
}
SFB_HD uint64_t sfb_bloom_word(uint64_t km, uint64_t h, const SfbBloomGeom& g, uint64_t n_words) {
    const uint64_t e = km ^ (km >> 2);
    const uint64_t q = ~(e | (e >> 1)) & g.wmask;
    const uint64_t a1 = q & ~(km >> 5);
    const uint64_t a = a1 ? a1 : q;
#ifdef __CUDA_ARCH__
    const int sh = a ? (__ffsll((long long)a) - 1) : 0;
#else
    const int sh = a ? __builtin_ctzll(a) : 0;
#endif
    const uint64_t mm = (km >> sh) & 0x3FFFFFFFULL;
    const uint64_t sector = (mm * 0x9E3779B97F4A7C15ULL) >> g.shift;
    return ((sector << 2) | ((h >> 6) & 3)) & (n_words - 1);
}
SFB_HD uint64_t sfb_bloom_mask(uint64_t h) {
    return (1ULL << ((h >> 8) & 63)) | (1ULL << ((h >> 14) & 63)) | (1ULL << ((h >> 20) & 63));
}

