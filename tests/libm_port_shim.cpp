// Host build of wrf_b200/csrc/libm_port.h for tests/test_libm_port.py: compares the
// restated glibc routines with the system libm over large argument sets.  Test
// infrastructure; the product compiles the same header with nvcc.
#include "../wrf_b200/csrc/libm_port.h"
#include "../wrf_b200/csrc/exact_div.h"
#include <cmath>
#include <cstdint>
#include <cstring>
static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fl(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
extern "C" {
float port_powf(float x, float y) { return wg_libm::powf_(x, y); }
float port_logf(float x) { return wg_libm::logf_(x); }
float port_expf(float x) { return wg_libm::expf_(x); }
float port_sinf(float x) { return wg_libm::sinf_(x); }
// scan the float bit patterns [lo, hi) with stride `step`; which: 0 logf, 1 expf, 2 sinf, 3 powf(x, y)
// returns the number of arguments whose result bits differ from libm's (NaNs compare equal)
// exact division by a constant: which = 60, 12 or 981 (g)
long long div_scan(int which, uint32_t lo, uint32_t hi, uint32_t step, uint32_t *first_bad) {
  long long bad = 0;
  for (uint64_t u = lo; u < hi; u += step) {
    const float x = fl((uint32_t)u);
    if (x != x) continue;
    float a, b;
    if (which == 60) { a = WG_DIV60(x); b = x / 60.0f; }
    else if (which == 12) { a = WG_DIV12(x); b = x / 12.0f; }
    else { a = WG_DIVG(x); b = x / 9.81f; }
    if (bits(a) != bits(b)) { if (!bad && first_bad) *first_bad = (uint32_t)u; bad++; }
  }
  return bad;
}
long long port_scan(int which, uint32_t lo, uint32_t hi, uint32_t step, float y, uint32_t *first_bad) {
  long long bad = 0;
  for (uint64_t u = lo; u < hi; u += step) {
    const float x = fl((uint32_t)u);
    float a, b;
    switch (which) {
      case 0: a = wg_libm::logf_(x); b = logf(x); break;
      case 1: a = wg_libm::expf_(x); b = expf(x); break;
      case 2: a = wg_libm::sinf_(x); b = sinf(x); break;
      default: a = wg_libm::powf_(x, y); b = powf(x, y); break;
    }
    if (bits(a) != bits(b) && !(a != a && b != b)) {
      if (!bad && first_bad) *first_bad = (uint32_t)u;
      bad++;
    }
  }
  return bad;
}
}
