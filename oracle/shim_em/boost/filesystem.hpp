// Stand-in for boost/filesystem.hpp: the few members the reference's headers touch.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <ostream>
#include <sys/stat.h>
#ifndef BOOST_LIKELY
#define BOOST_LIKELY(x) __builtin_expect(!!(x), 1)
#define BOOST_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif
namespace boost { namespace filesystem {
class path {
  public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
    const char* c_str() const { return s_.c_str(); }
    path extension() const { auto p = s_.rfind('.'); return p == std::string::npos ? path() : path(s_.substr(p)); }
    path operator/(const path& o) const { return path(s_.empty() || s_.back() == '/' ? s_ + o.s_ : s_ + "/" + o.s_); }
  private:
    std::string s_;
};
inline std::ostream& operator<<(std::ostream& os, const path& p) { return os << p.string(); }
inline bool exists(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_regular_file(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
inline bool is_empty(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && st.st_size == 0; }
inline bool create_directories(const path& p) { return ::mkdir(p.c_str(), 0755) == 0; }
}}  // namespace boost::filesystem
