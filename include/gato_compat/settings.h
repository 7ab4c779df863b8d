// Source-compatibility shim for code written against the reference's gato/settings.h (gato/settings.h:7-41).
// The plant and horizon stay compile-time macros for such callers (KNOT_POINTS, PLANT_INDY7 / PLANT_IIWA14,
// CMakeLists.txt:71-75) and are forwarded to gato_b200's runtime selectors.
#pragma once
#include <cstdint>
namespace sqp {
typedef float T;  // the B200 path is fp32 only (the reference's Python modules register float only, bindings.cu:244-265)
constexpr uint32_t NUM_ALPHAS = 8;
constexpr float    RHO_INIT = 1e-3, RHO_FACTOR = 1.2, RHO_MIN = 1e-8, RHO_MAX = 10;
}  // namespace sqp
#if defined(PLANT_INDY7)
namespace grid { constexpr int NUM_JOINTS = 6, NQ = 6, NX = 12, NU = 6, EE_POS_SIZE = 6; }
#define GATO_COMPAT_PLANT 0
#elif defined(PLANT_IIWA14)
namespace grid { constexpr int NUM_JOINTS = 7, NQ = 7, NX = 14, NU = 7, EE_POS_SIZE = 6; }
#define GATO_COMPAT_PLANT 1
#else
#error "Plant type must be defined: PLANT_INDY7 or PLANT_IIWA14"
#endif
#ifndef KNOT_POINTS
#error "KNOT_POINTS must be defined"
#endif
