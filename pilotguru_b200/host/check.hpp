// glog-style CHECK for the host binaries: the reference aborts with a message on a failed CHECK
// (google::InstallFailureSignalHandler, fit_motion.cc:299-313); so do these.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace pgbhost {
class FatalMessage {
 public:
  FatalMessage(const char* file, int line, const char* cond) { ss_ << file << ":" << line << "] Check failed: " << cond << " "; }
  [[noreturn]] ~FatalMessage() {
    std::cerr << "F " << ss_.str() << std::endl;
    std::abort();
  }
  std::ostream& stream() { return ss_; }

 private:
  std::ostringstream ss_;
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace pgbhost

#define PGB_CHECK(cond) (cond) ? (void)0 : pgbhost::Voidify() & pgbhost::FatalMessage(__FILE__, __LINE__, #cond).stream()
// status check of a libpgb200 call
#define PGB_CALL(expr) PGB_CHECK((expr) == 0) << pgb_last_error()
