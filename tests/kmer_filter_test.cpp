// CPU check of the presence filter's addressing (sailfish_b200/csrc/kmer_filter.hpp): no false negatives, the false-positive
// rate at the size index.cu chooses, and the locality the scan kernel relies on (distinct 32-byte sectors touched by the successive
// k-mers of a read).  Prints "ok <fpr> <sectors per 46 k-mers>"; exit code 1 on failure.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../sailfish_b200/csrc/kmer_filter.hpp"

static uint64_t kmer_at(const std::vector<uint8_t>& t, size_t p, int k) {
    uint64_t v = 0;
    for (int i = 0; i < k; ++i) v |= (uint64_t)t[p + i] << (2 * i);
    return v;
}

int main(int argc, char** argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 31;
    const size_t N = 2000000;
    std::mt19937_64 rng(12345);
    std::vector<uint8_t> text(N + k);
    for (auto& b : text) b = rng() & 3;
    // the filter as index.cu sizes it: >= 16 bits per k-mer
    uint64_t words = 64;
    while (words * 64 < 16 * N) words <<= 1;
    const SfbBloomGeom g = sfb_bloom_geom(k, words);
    std::vector<uint64_t> bloom(words, 0);
    for (size_t p = 0; p < N; ++p) {
        const uint64_t km = kmer_at(text, p, k), h = sfb_kmer_mix(km);
        bloom[sfb_bloom_word(km, h, g, words)] |= sfb_bloom_mask(h);
    }
    // no false negatives
    for (size_t p = 0; p < N; p += 7) {
        const uint64_t km = kmer_at(text, p, k), h = sfb_kmer_mix(km);
        const uint64_t need = sfb_bloom_mask(h);
        if ((bloom[sfb_bloom_word(km, h, g, words)] & need) != need) { printf("false negative at %zu\n", p); return 1; }
    }
    // false positives: k-mers of an unrelated random text (absent with overwhelming probability for k >= 19)
    std::vector<uint8_t> other(400000 + k);
    for (auto& b : other) b = rng() & 3;
    size_t fp = 0, probes = 0;
    double sectors = 0; size_t windows = 0;
    for (size_t r = 0; r + 76 <= other.size(); r += 76) {        // "reads" of 76 bases: 46 k-mers each for k = 31
        std::set<uint64_t> sec;
        for (size_t i = 0; i + k <= 76; ++i) {
            const uint64_t km = kmer_at(other, r + i, k), h = sfb_kmer_mix(km);
            const uint64_t w = sfb_bloom_word(km, h, g, words), need = sfb_bloom_mask(h);
            if (w >= words) { printf("word out of range\n"); return 1; }
            sec.insert(w >> 2);
            ++probes;
            if ((bloom[w] & need) == need) ++fp;
        }
        sectors += sec.size(); ++windows;
    }
    const double fpr = (double)fp / probes, per_read = sectors / windows;
    printf("ok %.5f %.2f\n", fpr, per_read);
    if (k >= 19 && fpr > 0.02) return 1;
    if (k == 31 && per_read > 12.0) return 1;
    return 0;
}
