// cuNVSM/base.h — basic typedefs of the C++ façade over libnvsm_b200 (include/nvsm_b200.h).
// Mirrors the names a caller of the reference's include/cuNVSM/base.h relies on
// (reference: include/cuNVSM/base.h:22-36): `int32` is `long`, RNG is std::minstd_rand0.
#ifndef CUNVSM_B200_BASE_H
#define CUNVSM_B200_BASE_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <sstream>
#include <string>

enum ParamIdentifier { WORD_REPRS, TRANSFORM, ENTITY_REPRS };

typedef unsigned short uint16;
typedef long int32;            // sic: 64-bit on LP64, exactly like the reference
typedef unsigned long uint32;
typedef long long int64;
typedef unsigned long long uint64;
typedef float float32;
typedef double float64;

typedef std::minstd_rand0 RNG;

#ifndef FLOATING_POINT_TYPE
#define FLOATING_POINT_TYPE float32   // the release build of the reference (cpp/CMakeLists.txt:17)
#endif

// The reference aborts on any error (glog CHECK / LOG(FATAL)); the façade does the same.
#define NVSM_CHECK(cond, msg)                                                        \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "Check failed: %s %s (%s:%d)\n", #cond, (msg), __FILE__, __LINE__); \
            std::abort();                                                            \
        }                                                                            \
    } while (0)

namespace nvsm_detail {
inline unsigned long rng_get_state(const RNG& rng) {
    std::ostringstream ss;
    ss << rng;
    return std::stoul(ss.str());
}
inline void rng_set_state(RNG* rng, unsigned long state) { rng->seed(state); }
}  // namespace nvsm_detail

#endif  // CUNVSM_B200_BASE_H
