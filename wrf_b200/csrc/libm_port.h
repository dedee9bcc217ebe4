// libm_port.h -- single-precision powf / logf / expf / sinf evaluated exactly as the
// reference's runtime evaluates them.
//
// The reference is Fortran real*4 built with gfortran: `x**y`, LOG, EXP and SIN on the
// hot path (module_big_step_utilities_em.F:1048,1072,1085 calc_p_rho_phi; frame/libmassv.F
// vspow; module_small_step_em.F:1453; module_big_step_utilities_em.F:5930-6104;
// phys/module_mp_kessler.F:101-102,218-226) compile to calls of glibc's powf/logf/expf/
// sinf.  That libm is a third-party dependency which is not part of /root/reference:
// GNU libc >= 2.28 (2.39 in this image), whose single-precision routines are the ARM
// "optimized-routines" algorithms (Szabolcs Nagy, 2017): the argument is widened to
// double, a 16-entry 1/c,log(c) table plus a degree-3..5 polynomial gives log/log2, a
// 32-entry 2^(i/32) table plus a cubic gives exp2, and the double result is rounded to
// float ONCE.  The routines are not correctly rounded (up to 0.82 ulp for powf), so
// CUDA's powf (or even a correctly rounded double pow) differs from them in the last bit
// at several per cent of the arguments -- and one ulp of the 1e5 Pa total pressure is
// ~1e-2 Pa of p', enough to move w by 1e-3 relative after one step.  Restating the
// algorithm with its constants makes the device result bit-identical to the host libm,
// so whole-step GPU-vs-oracle parity can be held to bit equality instead of a
// libm-sized tolerance.
//
// x86-64 glibc selects its FMA build of these routines (ifunc, sysdeps/x86_64/fpu/
// multiarch) on every CPU with FMA3, where each `a*b+c` of the C source is one fused
// multiply-add; the fma() calls below reproduce that.  tests/test_libm_port.py compiles
// this header for the host and compares it with the system libm over millions of
// arguments.
//
// Only the argument ranges the path can produce take the fast route; everything else
// (negative bases, NaN/Inf, |x| >= 120 for sinf) is forwarded to the toolkit's own
// function so behaviour stays defined.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define WG_HD __host__ __device__ __forceinline__
#else
#define WG_HD static inline
#endif

namespace wg_libm {

#if defined(__CUDA_ARCH__)
WG_HD double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
WG_HD uint32_t asuint(float f) { return __float_as_uint(f); }
WG_HD float asfloat(uint32_t u) { return __uint_as_float(u); }
WG_HD uint64_t asuint64(double d) { return (uint64_t)__double_as_longlong(d); }
WG_HD double asdouble(uint64_t u) { return __longlong_as_double((long long)u); }
#define WG_TAB __constant__
#else
WG_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
WG_HD uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
WG_HD float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
WG_HD uint64_t asuint64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
WG_HD double asdouble(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#define WG_TAB static const
#endif

// ---- tables (glibc sysdeps/ieee754/flt-32: e_exp2f_data.c, e_logf_data.c, e_powf_log2_data.c)
// 2^(i/32) with the exponent field pre-decremented by i<<47 so that adding k<<47 assembles 2^(k/32)
#define WG_EXP2F_TAB                                                                                 \
  {0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,       \
   0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,       \
   0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,       \
   0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,       \
   0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,       \
   0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,       \
   0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,       \
   0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull}
// 1/c for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)
#define WG_LOG_INVC                                                                                   \
  {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010b0p+0, 0x1.3c995b0b80385p+0,            \
   0x1.30d190c8864a5p+0, 0x1.25e227b0b8ea0p+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,            \
   0x1.0953f419900a7p+0, 0x1.0000000000000p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aa0p-1,            \
   0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1}
// ln(c)
#define WG_LOGF_LOGC                                                                                  \
  {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,        \
   -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c8100p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,        \
   -0x1.252f438e10c1ep-5, 0x0.0p+0, 0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4,                        \
   0x1.526e57720db08p-3, 0x1.bc2860d224770p-3, 0x1.1058bc8a07ee1p-2, 0x1.4043057b6ee09p-2}
// log2(c)
#define WG_POWF_LOGC                                                                                  \
  {-0x1.efec65b963019p-2, -0x1.b0b6832d4fca4p-2, -0x1.7418b0a1fb77bp-2, -0x1.39de91a6dcf7bp-2,        \
   -0x1.01d9bf3f2b631p-2, -0x1.97c1d1b3b7af0p-3, -0x1.2f9e393af3c9fp-3, -0x1.960cbbf788d5cp-4,        \
   -0x1.a6f9db6475fcep-5, 0x0.0p+0, 0x1.338ca9f24f53dp-4, 0x1.476a9543891bap-3,                        \
   0x1.e840b4ac4e4d2p-3, 0x1.40645f0c6651cp-2, 0x1.88e9c2c1b9ff8p-2, 0x1.ce0a44eb17bccp-2}

#if defined(__CUDACC__)
__constant__ uint64_t c_exp2f_tab[32] = WG_EXP2F_TAB;
__constant__ double c_log_invc[16] = WG_LOG_INVC;
__constant__ double c_logf_logc[16] = WG_LOGF_LOGC;
__constant__ double c_powf_logc[16] = WG_POWF_LOGC;
#endif
static const uint64_t h_exp2f_tab[32] = WG_EXP2F_TAB;
static const double h_log_invc[16] = WG_LOG_INVC;
static const double h_logf_logc[16] = WG_LOGF_LOGC;
static const double h_powf_logc[16] = WG_POWF_LOGC;

#if defined(__CUDA_ARCH__)
#define T_EXP2F(i) c_exp2f_tab[i]
#define T_INVC(i) c_log_invc[i]
#define T_LOGF_LOGC(i) c_logf_logc[i]
#define T_POWF_LOGC(i) c_powf_logc[i]
#else
#define T_EXP2F(i) h_exp2f_tab[i]
#define T_INVC(i) h_log_invc[i]
#define T_LOGF_LOGC(i) h_logf_logc[i]
#define T_POWF_LOGC(i) h_powf_logc[i]
#endif

// ---- logf (e_logf.c) -------------------------------------------------------
WG_HD float logf_(float x) {
  uint32_t ix = asuint(x);
  if (ix == 0x3f800000u) return 0.f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0u) return -INFINITY;          // log(+-0)
    if (ix == 0x7f800000u) return x;              // log(inf)
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return ::logf(x);   // negative / NaN
    ix = asuint(x * 0x1p23f);                     // subnormal: normalise
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & (0x1ffu << 23));
  const double invc = T_INVC(i), logc = T_LOGF_LOGC(i);
  const double z = (double)asfloat(iz);
  const double r = fma_(z, invc, -1.0);
  const double y0 = fma_((double)k, 0x1.62e42fefa39efp-1, logc);
  const double r2 = r * r;
  double y = fma_(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
  y = fma_(-0x1.00ea348b88334p-2, r2, y);
  y = fma_(y, r2, y0 + r);
  return (float)y;
}

// ---- exp2 core shared by powf (e_powf.c: exp2_inline) ----------------------
WG_HD float exp2_core(double xd) {
  const double SHIFT = 0x1.8p+47;   // 0x1.8p52 / 32
  double kd = xd + SHIFT;
  const uint64_t ki = asuint64(kd);
  kd -= SHIFT;
  const double r = xd - kd;
  uint64_t t = T_EXP2F(ki & 31u);
  t += ki << 47;
  const double s = asdouble(t);
  const double z = fma_(0x1.c6af84b912394p-5, r, 0x1.ebfce50fac4f3p-3);
  const double r2 = r * r;
  double y = fma_(0x1.62e42ff0c52d6p-1, r, 1.0);
  y = fma_(z, r2, y);
  y = y * s;
  return (float)y;
}

// ---- powf (e_powf.c), positive finite x ------------------------------------
WG_HD float powf_(float x, float y) {
  uint32_t ix = asuint(x);
  const uint32_t iy = asuint(y);
  const bool y_special = (2u * iy - 1u >= 2u * 0x7f800000u - 1u);   // y is 0, inf or nan
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || y_special) {
    if (y_special || (ix & 0x80000000u) || ix >= 0x7f800000u) return ::powf(x, y);   // not on the path
    if (ix == 0u) return (iy & 0x80000000u) ? INFINITY : 0.f;                        // (+0)^y
    ix = asuint(x * 0x1p23f) & 0x7fffffffu;                                          // subnormal x
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const double invc = T_INVC(i), logc = T_POWF_LOGC(i);
  const double z = (double)asfloat(iz);
  const double r = fma_(z, invc, -1.0);
  const double y0 = logc + (double)k;
  double a = fma_(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
  const double p = fma_(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
  const double r2 = r * r;
  double q = fma_(r, 0x1.71547652ab82bp+0, y0);
  const double r4 = r2 * r2;
  q = fma_(r2, p, q);
  a = fma_(a, r4, q);
  const double ylogx = (double)y * a;
  if (((asuint64(ylogx) >> 47) & 0xffffu) >= (asuint64(126.0) >> 47)) {   // |y*log2(x)| >= 126
    if (ylogx > 0x1.fffffffd1d571p+6) return INFINITY;
    if (ylogx <= -150.0) return 0.f;
  }
  return exp2_core(ylogx);
}

// ---- expf (e_expf.c) -------------------------------------------------------
WG_HD float expf_(float x) {
  const uint32_t abstop = (asuint(x) >> 20) & 0x7ffu;
  if (abstop >= (asuint(88.0f) >> 20)) {
    if (asuint(x) == asuint(-INFINITY)) return 0.f;
    if (abstop >= (asuint(INFINITY) >> 20)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.f;
  }
  const double xd = (double)x;
  const double InvLn2N = 0x1.71547652b82fep+5;  // 32/ln2
  const double SHIFT = 0x1.8p+52;
  // z = InvLn2N*xd is never rounded in the FMA build: both uses are contracted
  double kd = fma_(InvLn2N, xd, SHIFT);
  const uint64_t ki = asuint64(kd);
  kd -= SHIFT;
  const double r = fma_(InvLn2N, xd, -kd);
  uint64_t t = T_EXP2F(ki & 31u);
  t += ki << 47;
  const double s = asdouble(t);
  const double zz = fma_(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
  const double r2 = r * r;
  double y = fma_(0x1.62e42ff0c52d6p-6, r, 1.0);
  y = fma_(zz, r2, y);
  y = y * s;
  return (float)y;
}

// ---- sinf (s_sinf.c + s_sincosf.h), |x| < 120 ------------------------------
WG_HD float sinf_(float yf) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
               C4 = 0x1.99343027bf8c3p-16;
  const uint32_t at = (asuint(yf) >> 20) & 0x7ffu;
  double x = (double)yf;
  if (at < (asuint(0x1.921FB6p-1f) >> 20)) {          // |y| < pi/4
    const double x2 = x * x;
    if (at < (asuint(0x1p-12f) >> 20)) return yf;
    const double x3 = x * x2;
    const double s1 = fma_(x2, S3, S2);
    const double x7 = x3 * x2;
    const double s = fma_(x3, S1, x);
    return (float)fma_(x7, s1, s);
  }
  if (at < (asuint(120.0f) >> 20)) {
    const double r = x * 0x1.45f306dc9c883p+23;       // 2/pi * 2^24
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = fma_(-(double)n, 0x1.921fb54442d18p+0, x);
    const double sgn = (n & 2) ? -1.0 : 1.0;          // table 1 is table 0 negated; sign[n&3] = {1,-1,-1,1}
    const double x2 = x * x;
    if ((n & 1) == 0) {
      // sin polynomial on x * sign[n&3]; with n even sign[n&3] = (n&2 ? -1 : 1) and table 1
      // negates nothing of S1..S3, so the two signs cancel into one
      const double sg = ((n & 3) == 0) ? 1.0 : -1.0;  // sign[0]=1, sign[2]=-1
      const double xs = x * sg;
      const double x3 = xs * x2;
      const double s1 = fma_(x2, S3, S2);
      const double x7 = x3 * x2;
      const double s = fma_(x3, S1, xs);
      return (float)fma_(x7, s1, s);
    } else {
      const double x4 = x2 * x2;
      const double c2 = fma_(x2, sgn * C4, sgn * C3);
      const double c1 = fma_(x2, sgn * C1, sgn * C0);
      const double x6 = x4 * x2;
      const double c = fma_(x4, sgn * C2, c1);
      return (float)fma_(x6, c2, c);
    }
  }
  return ::sinf(yf);
}

}  // namespace wg_libm
