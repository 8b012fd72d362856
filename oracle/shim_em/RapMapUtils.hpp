// Stand-in for RapMap's RapMapUtils.hpp (RapMap sf-v0.10.1 is not in /root/reference): only the names the
// reference's headers mention.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
namespace rapmap { namespace utils {
enum class MateStatus : uint8_t { SINGLE_END = 0, PAIRED_END_LEFT = 1, PAIRED_END_RIGHT = 2, PAIRED_END_PAIRED = 3 };
struct HitCounters {};
}}  // namespace rapmap::utils
