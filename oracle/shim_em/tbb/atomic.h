// Stand-in for tbb/atomic.h (TBB is not in /root/reference): just enough of tbb::atomic<T> for the reference's
// CollapsedEMOptimizer.cpp / Transcript.hpp to compile unmodified.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <atomic>
namespace tbb {
template <typename T>
class atomic {
  public:
    atomic() : v_(T()) {}
    atomic(T v) : v_(v) {}
    atomic(const atomic& o) : v_(o.load()) {}
    atomic& operator=(const atomic& o) { store(o.load()); return *this; }
    atomic& operator=(T v) { store(v); return *this; }
    operator T() const { return load(); }
    T load() const { return v_.load(std::memory_order_relaxed); }
    void store(T v) { v_.store(v, std::memory_order_relaxed); }
    // returns the value seen before the operation (TBB semantics)
    T compare_and_swap(T desired, T expected) {
        v_.compare_exchange_strong(expected, desired);
        return expected;
    }
    atomic& operator+=(T x) { T o = load(); while (!v_.compare_exchange_weak(o, o + x)) {} return *this; }
    T operator++(int) { T o = load(); while (!v_.compare_exchange_weak(o, o + 1)) {} return o; }
  private:
    std::atomic<T> v_;
};
}  // namespace tbb
