// Stand-in for <glog/logging.h> (glog is not installed in this image) so that the UNMODIFIED
// reference sources under /root/reference compile for oracle/_ref. Test infrastructure only.
// Semantics kept: CHECK* / LOG(FATAL) print and abort; DCHECK* compile to nothing under NDEBUG;
// every other severity is swallowed.
#ifndef REF_SHIM_GLOG_LOGGING_H
#define REF_SHIM_GLOG_LOGGING_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <atomic>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace ref_shim {

class FatalMessage {
 public:
  FatalMessage(const char* file, int line) { ss_ << "[reference FATAL] " << file << ":" << line << " "; }
  ~FatalMessage() {
      std::cerr << ss_.str() << std::endl;
      std::abort();
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
};

struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};

struct Voidify {
  void operator&(std::ostream&) {}
  void operator&(const NullStream&) {}
};

template <typename T>
inline T check_notnull(const char* file, int line, const char* expr, T ptr) {
    if (ptr == nullptr) {
        FatalMessage(file, line).stream() << "'" << expr << "' must be non-null";
    }
    return ptr;
}

}  // namespace ref_shim


#define REF_SHIM_FATAL_STREAM() ::ref_shim::FatalMessage(__FILE__, __LINE__).stream()
#define REF_SHIM_NULL_STREAM() ::ref_shim::NullStream()

#define REF_SHIM_LOG_INFO REF_SHIM_NULL_STREAM()
#define REF_SHIM_LOG_WARNING REF_SHIM_NULL_STREAM()
#define REF_SHIM_LOG_ERROR REF_SHIM_NULL_STREAM()
#define REF_SHIM_LOG_FATAL REF_SHIM_FATAL_STREAM()

#define LOG(severity) REF_SHIM_LOG_##severity
#define VLOG(level) REF_SHIM_NULL_STREAM()
#define DLOG(severity) REF_SHIM_NULL_STREAM()
#define LOG_IF(severity, cond) REF_SHIM_NULL_STREAM()
#define LOG_EVERY_N(severity, n) REF_SHIM_NULL_STREAM()
#define DLOG_EVERY_N(severity, n) REF_SHIM_NULL_STREAM()
#define LOG_IF_EVERY_N(severity, cond, n) REF_SHIM_NULL_STREAM()
#define LOG_FIRST_N(severity, n) REF_SHIM_NULL_STREAM()

#define CHECK(cond) \
  (cond) ? (void) 0 : ::ref_shim::Voidify() & REF_SHIM_FATAL_STREAM() << "Check failed: " #cond " "

#define REF_SHIM_CHECK_OP(a, b, op) \
  ((a) op (b)) ? (void) 0 : ::ref_shim::Voidify() & REF_SHIM_FATAL_STREAM() \
      << "Check failed: " #a " " #op " " #b " "

#define CHECK_EQ(a, b) REF_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) REF_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) REF_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) REF_SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) REF_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) REF_SHIM_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(ptr) ::ref_shim::check_notnull(__FILE__, __LINE__, #ptr, (ptr))

#ifdef NDEBUG
#define REF_SHIM_DCHECK_OFF(expr) \
  true ? (void) 0 : ::ref_shim::Voidify() & REF_SHIM_NULL_STREAM()
#define DCHECK(cond) REF_SHIM_DCHECK_OFF(cond)
#define DCHECK_EQ(a, b) REF_SHIM_DCHECK_OFF((a) == (b))
#define DCHECK_NE(a, b) REF_SHIM_DCHECK_OFF((a) != (b))
#define DCHECK_LT(a, b) REF_SHIM_DCHECK_OFF((a) < (b))
#define DCHECK_LE(a, b) REF_SHIM_DCHECK_OFF((a) <= (b))
#define DCHECK_GT(a, b) REF_SHIM_DCHECK_OFF((a) > (b))
#define DCHECK_GE(a, b) REF_SHIM_DCHECK_OFF((a) >= (b))
#define DCHECK_NOTNULL(ptr) (ptr)
#else
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_NOTNULL(ptr) CHECK_NOTNULL(ptr)
#endif

#endif  // REF_SHIM_GLOG_LOGGING_H
