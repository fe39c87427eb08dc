// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for <glog/logging.h> (google-glog is not in this image) that provides just
// the CHECK family, so that reference source files whose ONLY external dependency is glog can be compiled where they lie
// (oracle/Makefile, target _ref) and run as the real-reference pin of the oracle.  A failed check prints the message
// and aborts, like glog's.  Not part of the product; nothing under pilotguru_b200/ includes it.
#pragma once
#include <cstdlib>
#include <iostream>
#include <limits>   // the real header pulls these in; align_time_series.cc relies on it for std::numeric_limits
#include <string>
#include <vector>
#include <sstream>

namespace pgo_glog_shim {
struct Fatal {
  std::ostringstream s;
  Fatal(const char* file, int line, const char* what) { s << file << ":" << line << "] Check failed: " << what << " "; }
  template <typename T> Fatal& operator<<(const T& v) { s << v; return *this; }
  [[noreturn]] ~Fatal() { std::cerr << s.str() << std::endl; std::abort(); }
};
struct Voidify { void operator&(const Fatal&) {} };
template <typename T> T&& NotNull(const char* file, int line, const char* what, T&& p) {
  if (p == nullptr) Fatal(file, line, what);
  return static_cast<T&&>(p);
}
}  // namespace pgo_glog_shim

#define CHECK(cond) (cond) ? (void)0 : ::pgo_glog_shim::Voidify() & ::pgo_glog_shim::Fatal(__FILE__, __LINE__, #cond)
#define PGO_CHECK_OP(a, op, b) CHECK((a)op(b))
#define CHECK_EQ(a, b) PGO_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) PGO_CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) PGO_CHECK_OP(a, <, b)
#define CHECK_LE(a, b) PGO_CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) PGO_CHECK_OP(a, >, b)
#define CHECK_GE(a, b) PGO_CHECK_OP(a, >=, b)
#define CHECK_NOTNULL(p) ::pgo_glog_shim::NotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))

// LOG(severity) << ...: swallowed (velocity.cc logs its input, gradient and loss at INFO on every evaluation).
namespace pgo_glog_shim {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
};
}  // namespace pgo_glog_shim
#define LOG(severity) ::pgo_glog_shim::NullStream()
