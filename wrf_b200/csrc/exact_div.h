// exact_div.h -- IEEE-exact single-precision division by a compile-time constant in three
// instructions.  The reference's flux operators divide by literals (module_advect_em.F:202-213:
// (7*(q_i+q_im1) - ...)/12., (37*(...) - ...)/60.) and by g = 9.81; a multiplication by the
// reciprocal is not the same number, and nvcc's correctly rounded `/` costs MUFU.RCP + FCHK + four
// FFMA + a branch -- twelve such divisions per point made the stencil kernels issue-bound.
// With c = RN(1/d):   q0 = x*c;  r = fma(-d, q0, x);  q = fma(r, c, q0)
// is the correctly rounded quotient whenever it is a normal number (Markstein).  Checked
// EXHAUSTIVELY over all 2^32 bit patterns of x for d = 60, 12, 9.81 (host, fmaf): the only
// mismatches are |x| < 2^-103, which take the plain division below (zero keeps its sign through
// x*c).  tests/test_libm_port.py re-runs a strided version of that scan.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define WG_EXD __host__ __device__ __forceinline__
#else
#define WG_EXD static inline
#endif

// Device form: through fp64, exact for EVERY finite x (tiny ones and signed zeros included) with no branch at all:
//   q0 = x*C;  r = fma(d, q0, -x);  q = fma(-r, C, q0)   in double, C = RN(1/d), then one rounding to float.
// q is within a few 2^-53 of x/d; a float rounding boundary is never that close to x/d unless x/d sits on it exactly, and
// then the correction step returns it exactly (x/d is a double), so the final conversion rounds like the division itself
// (53 >= 2*24+2 bits: double rounding is innocuous).  Checked EXHAUSTIVELY on the host over all 2^32 bit patterns of x for
// d = 60, 12, 9.81: zero bitwise mismatches against x / d.  Five instructions (F2F, DMUL, 2 DFMA, F2F) on the otherwise
// idle fp64 / conversion pipes instead of FMUL + 2 FFMA + FSETP + BRA + BSSY + BSYNC: the branch around the tiny-|x| case
// kept the compiler from overlapping the four face fluxes of a point (ncu r2j: 30 % "wait", 7 % "branch resolving" stalls,
// 11 % of all instructions were BRA / BSSY / BSYNC).
WG_EXD float wg_div_const(float x, const float d, const float c) {
#if defined(__CUDA_ARCH__)
  (void)c;
  const double xd = (double)x, dd = (double)d, cd = 1.0 / (double)d;
  const double q0 = __dmul_rn(xd, cd);
  const double r = __fma_rn(dd, q0, -xd);
  return (float)__fma_rn(-r, cd, q0);
#else
  const float q0 = x * c;
  if (fabsf(x) < 0x1p-100f) return (x == 0.f) ? q0 : x / d;
  const float r = fmaf(-d, q0, x);
  return fmaf(r, c, q0);
#endif
}
#define WG_DIV60(x) wg_div_const((x), 60.0f, 1.0f / 60.0f)
#define WG_DIV12(x) wg_div_const((x), 12.0f, 1.0f / 12.0f)
#define WG_DIVG(x) wg_div_const((x), 9.81f, 1.0f / 9.81f)
