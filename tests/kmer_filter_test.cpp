// CPU check of the m-mer presence bitmap (sailfish_b200/csrc/kmer_filter.hpp) and of the skipping rule the scan kernel builds on it
// (map.cu: scan_step): with the bitmap of a text, "the last m-mer of the k-mer at i is absent => positions i .. i+k-m hold no indexed
// k-mer", "the m-mer at offset h is absent => positions i .. i+h hold none".  The skipping scan must report exactly the positions a
// one-position-at-a-time scan reports.  Prints "ok <positions probed per read position>"; exit code 1 on failure.
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
    const int k = argc > 1 ? atoi(argv[1]) : 11, m = argc > 2 ? atoi(argv[2]) : 6;
    const size_t N = argc > 3 ? atol(argv[3]) : 3000;
    if (sfb_mfilter_m(31, 300000000ULL) != 15 || sfb_mfilter_m(31, 1700000000ULL) != 16 || sfb_mfilter_m(15, 1000) != 15 ||
        sfb_mfilter_m(11, 1000) != 11 || sfb_mfilter_words(15) != (1ULL << 30) / 32) { printf("geometry\n"); return 1; }
    std::mt19937_64 rng(99 + k * 131 + m);
    std::vector<uint8_t> text(N);
    for (auto& b : text) b = rng() & 3;
    std::vector<uint32_t> bits(sfb_mfilter_words(m), 0);
    std::set<uint64_t> kmers;
    for (size_t p = 0; p + m <= N; ++p) { const uint64_t v = kmer_at(text, p, m); bits[v >> 5] |= 1u << (v & 31); }
    for (size_t p = 0; p + k <= N; ++p) kmers.insert(kmer_at(text, p, k));
    auto test = [&](uint64_t key) { return (bits[key >> 5] >> (key & 31)) & 1u; };
    const uint32_t J = k - m + 1, half = (J - 1) >> 1;
    size_t probes = 0, positions = 0;
    for (int trial = 0; trial < 400; ++trial) {
        // a "read": partly a copy of the text (with a substitution), partly random
        const size_t L = 60 + rng() % 60;
        std::vector<uint8_t> r(L);
        for (auto& b : r) b = rng() & 3;
        if (trial & 1) { const size_t s = rng() % (N - L); for (size_t i = 0; i < L; ++i) r[i] = text[s + i]; r[rng() % L] ^= 1; }
        std::vector<size_t> want, got;
        for (size_t i = 0; i + k <= L; ++i) if (kmers.count(kmer_at(r, i, k))) want.push_back(i);
        size_t i = 0;
        while (i + k <= L) {
            const uint64_t km = kmer_at(r, i, k);
            const bool ahead = i + J + k <= L;
            const bool b_last = test(sfb_mfilter_key(km, k - m, m));
            const bool b_next = ahead ? test(sfb_mfilter_key(kmer_at(r, i + J, k), k - m, m)) : true;
            const bool b_mid = half ? test(sfb_mfilter_key(km, half, m)) : true;
            ++probes;
            if (!b_last) { i += (ahead && !b_next) ? 2 * J : J; continue; }
            if (!b_mid) { i += half + 1; continue; }
            if (kmers.count(km)) got.push_back(i);                 // the table look-up
            i += 1;
        }
        positions += L - k + 1;
        if (got != want) { printf("trial %d: skipping scan differs (%zu vs %zu present positions)\n", trial, got.size(), want.size()); return 1; }
    }
    printf("ok %.3f\n", (double)probes / positions);
    return 0;
}
