// TEST INFRASTRUCTURE ONLY (see oracle.h).  CPU restatement of XXH64 and the library-format logic.
#include "oracle.h"
#include <cstring>
#include <string>
#include <cctype>

// ---------------------------------------------------------------------------------------------
// XXH64 -- follows /root/reference/src/xxhash.c:346-455 (XXH64_endian_align, little-endian reads).
// ---------------------------------------------------------------------------------------------
namespace {
const uint64_t P1 = 11400714785074694791ULL;
const uint64_t P2 = 14029467366897019727ULL;
const uint64_t P3 = 1609587929392839161ULL;
const uint64_t P4 = 9650029242287828579ULL;
const uint64_t P5 = 2870177450012600261ULL;

inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t lane(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl(acc, 31); return acc * P1; }
inline uint64_t fold(uint64_t h, uint64_t v) {           // xxhash.c:395-417
    v *= P2; v = rotl(v, 31); v *= P1; h ^= v; return h * P1 + P4;
}
}  // namespace

extern "C" uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {                                      // xxhash.c:362-418
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t* limit = end - 32;
        do {
            v1 = lane(v1, rd64(p)); v2 = lane(v2, rd64(p + 8));
            v3 = lane(v3, rd64(p + 16)); v4 = lane(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = fold(h, v1); h = fold(h, v2); h = fold(h, v3); h = fold(h, v4);
    } else {
        h = seed + P5;                                    // :420-423
    }
    h += static_cast<uint64_t>(len);                      // :425
    while (p + 8 <= end) {                                // :427-436
        uint64_t k1 = rd64(p); k1 *= P2; k1 = rotl(k1, 31); k1 *= P1;
        h ^= k1; h = rotl(h, 27) * P1 + P4; p += 8;
    }
    if (p + 4 <= end) {                                   // :438-443
        h ^= static_cast<uint64_t>(rd32(p)) * P1; h = rotl(h, 23) * P2 + P3; p += 4;
    }
    while (p < end) {                                     // :445-450
        h ^= (*p) * P5; h = rotl(h, 11) * P1; ++p;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;   // :452-456
    return h;
}

// ---------------------------------------------------------------------------------------------
// Library formats -- include/LibraryFormat.hpp:7-9,89-98 ; src/SailfishUtils.cpp:63-97,157-289
// ---------------------------------------------------------------------------------------------
namespace {
enum { SE = 0, PE = 1 };
enum { O_SAME = 0, O_AWAY = 1, O_TOWARD = 2, O_NONE = 3 };
enum { S_SA = 0, S_AS = 1, S_S = 2, S_A = 3, S_U = 4 };
inline int mkfmt(int type, int orient, int strand) { return (type & 1) | ((orient & 3) << 1) | ((strand & 7) << 3); }
inline int f_type(int id) { return id & 1; }
inline int f_orient(int id) { return (id >> 1) & 3; }
inline int f_strand(int id) { return (id >> 3) & 7; }
}  // namespace

extern "C" int orc_parse_libtype(const char* s) {          // SailfishUtils.cpp:63-97
    std::string f(s ? s : "");
    for (auto& c : f) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    struct { const char* n; int id; } tab[] = {
        {"IU", mkfmt(PE, O_TOWARD, S_U)},  {"ISF", mkfmt(PE, O_TOWARD, S_SA)}, {"ISR", mkfmt(PE, O_TOWARD, S_AS)},
        {"OU", mkfmt(PE, O_AWAY, S_U)},    {"OSF", mkfmt(PE, O_AWAY, S_SA)},   {"OSR", mkfmt(PE, O_AWAY, S_AS)},
        {"MU", mkfmt(PE, O_SAME, S_U)},    {"MSF", mkfmt(PE, O_SAME, S_S)},    {"MSR", mkfmt(PE, O_SAME, S_A)},
        {"U", mkfmt(SE, O_NONE, S_U)},     {"SF", mkfmt(SE, O_NONE, S_S)},     {"SR", mkfmt(SE, O_NONE, S_A)}};
    for (auto& e : tab) if (f == e.n) return e.id;
    return -1;
}

extern "C" int orc_compat_single(int expected, int32_t /*start*/, int is_fwd, int ms) {   // SailfishUtils.cpp:157-211
    const int es = f_strand(expected);
    switch (ms) {
        case 0:  // SINGLE_END
            return is_fwd ? (es == S_U || es == S_S) : (es == S_U || es == S_A);
        case 1:  // PAIRED_END_LEFT
            if (f_orient(expected) == O_SAME) return es == S_U || (es == S_S && is_fwd) || (es == S_A && !is_fwd);
            return is_fwd ? (es == S_U || es == S_S) : (es == S_U || es == S_A);
        case 2:  // PAIRED_END_RIGHT
            if (f_orient(expected) == O_SAME) return es == S_U || (es == S_S && is_fwd) || (es == S_A && !is_fwd);
            return is_fwd ? (es == S_U || es == S_A) : (es == S_U || es == S_S);
        default:
            return 0;
    }
}

extern "C" int orc_compat_paired(int expected, int observed) {                          // SailfishUtils.cpp:215-239
    if (f_type(observed) != PE) return 0;
    if (f_orient(expected) != f_orient(observed)) return 0;
    return f_strand(expected) == S_U || f_strand(expected) == f_strand(observed);
}

extern "C" int orc_hit_type(int32_t e1, int fwd1, uint32_t len1, int32_t e2, int fwd2, uint32_t len2, int dovetail) {
    if ((fwd1 != 0) != (fwd2 != 0)) {                                                   // SailfishUtils.cpp:243-289
        if (fwd1) {
            int32_t stretch = dovetail ? static_cast<int32_t>(len2) : 0;
            return (e1 <= e2 + stretch) ? mkfmt(PE, O_TOWARD, S_SA) : mkfmt(PE, O_AWAY, S_SA);
        }
        int32_t stretch = dovetail ? static_cast<int32_t>(len1) : 0;
        return (e2 <= e1 + stretch) ? mkfmt(PE, O_TOWARD, S_AS) : mkfmt(PE, O_AWAY, S_AS);
    }
    return fwd1 ? mkfmt(PE, O_SAME, S_S) : mkfmt(PE, O_SAME, S_A);
}
