// Minimal stand-in for <boost/format.hpp> (not installed here) so the REFERENCE tracker kernels compile unmodified:
// their only use is a debug Print in numeric_cuda.h:226.  Test infrastructure only (oracle/ref_build.mk).
#pragma once
#include <sstream>
#include <string>
namespace boost {
class format {
 public:
  explicit format(const char* f) : f_(f) {}
  template <typename T> format& operator%(const T& v) { std::ostringstream o; o << v; args_ += o.str() + " "; return *this; }
  std::string str() const { return f_ + ": " + args_; }
 private:
  std::string f_, args_;
};
}  // namespace boost
