#pragma once
namespace boost {
template <typename T>
class irange_t {
  public:
    struct it { T v; T operator*() const { return v; } it& operator++() { ++v; return *this; } bool operator!=(const it& o) const { return v != o.v; } };
    irange_t(T b, T e) : b_(b), e_(e) {}
    it begin() const { return it{b_}; }
    it end() const { return it{e_}; }
  private:
    T b_, e_;
};
template <typename T> irange_t<T> irange(T b, T e) { return irange_t<T>(b, e); }
}  // namespace boost
