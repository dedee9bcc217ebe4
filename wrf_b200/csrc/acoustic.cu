// acoustic.cu -- the split-explicit small step, dyn_em/module_small_step_em.F:
// small_step_prep :16, small_step_finish :295, calc_p_rho :438, calc_coef_w :570,
// advance_uv :654, advance_mu_t :969, advance_w :1178, sumflux :1473,
// calc_mu_uv_1 (module_big_step_utilities_em.F:184).
//
// Column kernels (calc_coef_w, advance_mu_t, advance_w) run one thread per (i,j)
// column: a warp covers 32 consecutive i, so every k level is one coalesced
// 128-byte access per field and the k-recurrences (mass-divergence prefix sum,
// Thomas forward/backward sweeps) live in registers.
#include "common.cuh"
#include "libm_port.h"
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace wg {

// ---------------------------------------------------------------------------
// small_step_prep, 2-D part (:123-215)
__global__ void k_ssp_2d(const Geo g, Box b, int rk_step, cfp mub, fp mu_1, fp mu_2, cfp muu, fp muus, cfp muv, fp muvs,
                         fp muts, fp mudf, fp mu_save) {
  TID2;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1);
  if (rk_step == 1) {
    if (im && jm) {
      const float m2 = A2(mu_2, i, j);
      A2(mu_1, i, j) = m2;
      A2(mudf, i, j) = 0.f;
      A2(muts, i, j) = A2(mub, i, j) + m2;
      A2(mu_save, i, j) = m2;
      A2(mu_2, i, j) = 0.f;
    }
    if (jm) A2(muus, i, j) = A2(muu, i, j);
    if (im) A2(muvs, i, j) = A2(muv, i, j);
  } else {
    if (im && jm) {
      A2(muts, i, j) = A2(mub, i, j) + A2(mu_1, i, j);
      A2(mu_save, i, j) = A2(mu_2, i, j);
      A2(mu_2, i, j) = A2(mu_1, i, j) - A2(mu_2, i, j);
    }
    if (jm) A2(muus, i, j) = 0.5f * (A2(mub, i, j) + A2(mu_1, i, j) + A2(mub, i - 1, j) + A2(mu_1, i - 1, j));
    if (im) A2(muvs, i, j) = 0.5f * (A2(mub, i, j) + A2(mu_1, i, j) + A2(mub, i, j - 1) + A2(mu_1, i, j - 1));
  }
}

struct SspArgs {
  fp u_1, u_2, v_1, v_2, w_1, w_2, t_1, t_2, ph_1, ph_2;
  fp u_save, v_save, w_save, t_save, ph_save, ww_save, c2a;
  cfp ww, pb, p, alt, muu, muus, muv, muvs, mut, muts;
};

// small_step_prep, 3-D part (:128-170, :218-290)
// zero_rest: also write the zeros rk_tendency leaves in the parts of u_save ... t_save this kernel does not
// define (module_em.F:370-428 zeroes the whole arrays first); wrfgpu_step then skips that zeroing pass
__global__ void k_ssp_3d(const Geo g, Box b, int rk_step, int zero_rest, SspArgs a) {
  TID3;
  const int kde = g.kde;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  if (km && jm) {   // u points
    const float u2 = A3(a.u_2, i, k, j);
    float u1;
    if (rk_step == 1) { u1 = u2; A3(a.u_1, i, k, j) = u2; } else u1 = A3(a.u_1, i, k, j);
    A3(a.u_save, i, k, j) = u2;
    A3(a.u_2, i, k, j) = ((c1h * A2(a.muus, i, j) + c2h) * u1 - (c1h * A2(a.muu, i, j) + c2h) * u2) / A2(g.msfuy, i, j);
  } else if (zero_rest) A3(a.u_save, i, k, j) = 0.f;
  if (km && im) {   // v points
    const float v2 = A3(a.v_2, i, k, j);
    float v1;
    if (rk_step == 1) { v1 = v2; A3(a.v_1, i, k, j) = v2; } else v1 = A3(a.v_1, i, k, j);
    A3(a.v_save, i, k, j) = v2;
    A3(a.v_2, i, k, j) = ((c1h * A2(a.muvs, i, j) + c2h) * v1 - (c1h * A2(a.muv, i, j) + c2h) * v2) * A2(g.msfvx_inv, i, j);
  } else if (zero_rest) A3(a.v_save, i, k, j) = 0.f;
  if (im && jm) {
    const float muts = A2(a.muts, i, j), mut = A2(a.mut, i, j);
    if (km) {
      const float t2 = A3(a.t_2, i, k, j);
      float t1;
      if (rk_step == 1) { t1 = t2; A3(a.t_1, i, k, j) = t2; } else t1 = A3(a.t_1, i, k, j);
      A3(a.c2a, i, k, j) = WG_CPOVCV * (A3(a.pb, i, k, j) + A3(a.p, i, k, j)) / A3(a.alt, i, k, j);
      A3(a.t_save, i, k, j) = t2;
      A3(a.t_2, i, k, j) = (c1h * muts + c2h) * t1 - (c1h * mut + c2h) * t2;
    } else if (zero_rest) A3(a.t_save, i, k, j) = 0.f;
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    const float w2 = A3(a.w_2, i, k, j), p2 = A3(a.ph_2, i, k, j);
    float w1, p1;
    if (rk_step == 1) { w1 = w2; p1 = p2; A3(a.w_1, i, k, j) = w2; A3(a.ph_1, i, k, j) = p2; }
    else { w1 = A3(a.w_1, i, k, j); p1 = A3(a.ph_1, i, k, j); }
    A3(a.w_save, i, k, j) = w2;
    A3(a.w_2, i, k, j) = ((c1f * muts + c2f) * w1 - (c1f * mut + c2f) * w2) / A2(g.msfty, i, j);
    A3(a.ph_save, i, k, j) = p2;
    A3(a.ph_2, i, k, j) = p1 - p2;
    A3(a.ww_save, i, k, j) = A3(a.ww, i, k, j);
  } else if (zero_rest) {
    A3(a.w_save, i, k, j) = 0.f; A3(a.ph_save, i, k, j) = 0.f; A3(a.t_save, i, k, j) = 0.f;
  }
}

// The same sweep with 16-byte accesses: a thread owns four consecutive columns (a 16-byte aligned group of the padded
// device row), loads every operand of the group with one LDG.128 and stores a result with one STG.128 when all four
// lanes take the same branch (all but the groups that straddle its, ite or the last mass column).  26 streams at 128-byte
// warp requests ran at 0.66 of the copy bandwidth; a warp request is now 512 contiguous bytes per stream.
__global__ void k_ssp_3d_v4(const Geo g, Box b, int rk_step, int zero_rest, SspArgs a) {
  // b.i0 / b.i1: first and last GROUP index of the row (column = ims + 4 * group + lane)
  const int gi = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int k = b.k0 + blockIdx.y * blockDim.y + threadIdx.y;
  const int j = b.j0 + blockIdx.z * blockDim.z + threadIdx.z;
  if (gi > b.i1 || k > b.k1 || j > b.j1) return;
  const int i = g.ims + 4 * gi;                 // column of lane 0
  const int kde = g.kde;
  const int i_m = min(g.ite, g.ide - 1);        // last mass column
  const bool jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  // lanes inside the tile, and inside the mass columns
  const int lo = max(g.its - i, 0), hi = min(g.ite - i, 3), him = min(i_m - i, 3);
  const float c1h = g.c1h[k], c2h = g.c2h[k], c1f = g.c1f[k], c2f = g.c2f[k];
  float zero[4] = {0.f, 0.f, 0.f, 0.f};
  if (km && jm) {   // u points: every column of the tile
    float u2[4], u1[4], muus[4], muu[4], my[4], out[4];
    ld4(&A3(a.u_2, i, k, j), u2); ld4(&A2(a.muus, i, j), muus); ld4(&A2(a.muu, i, j), muu); ld4(&A2(g.msfuy, i, j), my);
    if (rk_step == 1) { for (int q = 0; q < 4; q++) u1[q] = u2[q]; st4(&A3(a.u_1, i, k, j), u2, lo, hi); }
    else ld4(&A3(a.u_1, i, k, j), u1);
    st4(&A3(a.u_save, i, k, j), u2, lo, hi);
    for (int q = 0; q < 4; q++) out[q] = ((c1h * muus[q] + c2h) * u1[q] - (c1h * muu[q] + c2h) * u2[q]) / my[q];
    st4(&A3(a.u_2, i, k, j), out, lo, hi);
  } else if (zero_rest) st4(&A3(a.u_save, i, k, j), zero, lo, hi);
  if (km) {   // v points: the mass columns; the last u column (if any) only gets zero_rest's zero
    if (him >= lo) {
      float v2[4], v1[4], muvs[4], muv[4], mi[4], out[4];
      ld4(&A3(a.v_2, i, k, j), v2); ld4(&A2(a.muvs, i, j), muvs); ld4(&A2(a.muv, i, j), muv); ld4(&A2(g.msfvx_inv, i, j), mi);
      if (rk_step == 1) { for (int q = 0; q < 4; q++) v1[q] = v2[q]; st4(&A3(a.v_1, i, k, j), v2, lo, him); }
      else ld4(&A3(a.v_1, i, k, j), v1);
      st4(&A3(a.v_save, i, k, j), v2, lo, him);
      for (int q = 0; q < 4; q++) out[q] = ((c1h * muvs[q] + c2h) * v1[q] - (c1h * muv[q] + c2h) * v2[q]) * mi[q];
      st4(&A3(a.v_2, i, k, j), out, lo, him);
    }
    if (zero_rest && hi > him) st4(&A3(a.v_save, i, k, j), zero, max(him + 1, lo), hi);
  } else if (zero_rest) st4(&A3(a.v_save, i, k, j), zero, lo, hi);
  if (jm) {
    if (him >= lo) {
      float muts[4], mut[4];
      ld4(&A2(a.muts, i, j), muts); ld4(&A2(a.mut, i, j), mut);
      if (km) {
        float t2[4], t1[4], pb[4], p[4], alt[4], out[4];
        ld4(&A3(a.t_2, i, k, j), t2); ld4(&A3(a.pb, i, k, j), pb); ld4(&A3(a.p, i, k, j), p); ld4(&A3(a.alt, i, k, j), alt);
        if (rk_step == 1) { for (int q = 0; q < 4; q++) t1[q] = t2[q]; st4(&A3(a.t_1, i, k, j), t2, lo, him); }
        else ld4(&A3(a.t_1, i, k, j), t1);
        for (int q = 0; q < 4; q++) out[q] = WG_CPOVCV * (pb[q] + p[q]) / alt[q];
        st4(&A3(a.c2a, i, k, j), out, lo, him);
        st4(&A3(a.t_save, i, k, j), t2, lo, him);
        for (int q = 0; q < 4; q++) out[q] = (c1h * muts[q] + c2h) * t1[q] - (c1h * mut[q] + c2h) * t2[q];
        st4(&A3(a.t_2, i, k, j), out, lo, him);
      } else if (zero_rest) st4(&A3(a.t_save, i, k, j), zero, lo, him);
      float w2[4], p2[4], w1[4], p1[4], my[4], ww[4], out[4];
      ld4(&A3(a.w_2, i, k, j), w2); ld4(&A3(a.ph_2, i, k, j), p2); ld4(&A2(g.msfty, i, j), my); ld4(&A3(a.ww, i, k, j), ww);
      if (rk_step == 1) {
        for (int q = 0; q < 4; q++) { w1[q] = w2[q]; p1[q] = p2[q]; }
        st4(&A3(a.w_1, i, k, j), w2, lo, him); st4(&A3(a.ph_1, i, k, j), p2, lo, him);
      } else { ld4(&A3(a.w_1, i, k, j), w1); ld4(&A3(a.ph_1, i, k, j), p1); }
      st4(&A3(a.w_save, i, k, j), w2, lo, him);
      for (int q = 0; q < 4; q++) out[q] = ((c1f * muts[q] + c2f) * w1[q] - (c1f * mut[q] + c2f) * w2[q]) / my[q];
      st4(&A3(a.w_2, i, k, j), out, lo, him);
      st4(&A3(a.ph_save, i, k, j), p2, lo, him);
      for (int q = 0; q < 4; q++) out[q] = p1[q] - p2[q];
      st4(&A3(a.ph_2, i, k, j), out, lo, him);
      st4(&A3(a.ww_save, i, k, j), ww, lo, him);
    }
    if (zero_rest && hi > him) {
      const int l2 = max(him + 1, lo);
      st4(&A3(a.w_save, i, k, j), zero, l2, hi); st4(&A3(a.ph_save, i, k, j), zero, l2, hi); st4(&A3(a.t_save, i, k, j), zero, l2, hi);
    }
  } else if (zero_rest) {
    st4(&A3(a.w_save, i, k, j), zero, lo, hi); st4(&A3(a.ph_save, i, k, j), zero, lo, hi); st4(&A3(a.t_save, i, k, j), zero, lo, hi);
  }
}

// calc_p_rho (:438-566), non-hydrostatic
__global__ void k_calc_p_rho(const Geo g, Box b, int step, fp al, fp p, cfp ph, cfp alt, cfp t_2, cfp t_1, cfp c2a, fp pm1,
                             cfp mu, cfp mut) {
  TID3;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  const float m = A2(mu, i, j), mt = A2(mut, i, j);
  const float altv = A3(alt, i, k, j), t1 = A3(t_1, i, k, j);
  const float alv = -1.f / (c1h * mt + c2h) * (altv * (c1h * m) + g.rdnw[k] * (A3(ph, i, k + 1, j) - A3(ph, i, k, j)));
  float pv = A3(c2a, i, k, j) * (altv * (A3(t_2, i, k, j) - (c1h * m) * t1) / ((c1h * mt + c2h) * (WG_T0 + t1)) - alv);
  A3(al, i, k, j) = alv;
  if (step == 0) {
    A3(pm1, i, k, j) = pv;
  } else {
    const float ptmp = pv;
    const float pold = A3(pm1, i, k, j);
    pv = pv + g.smdiv * (pv - pold);
    A3(pm1, i, k, j) = after_load(ptmp, pold);   // see common.cuh: keep the store behind the load
  }
  A3(p, i, k, j) = pv;
}

// calc_p_rho with 16-byte accesses (mass columns: lanes lo..him of a group)
__global__ void k_calc_p_rho_v4(const Geo g, Box b, int step, fp al, fp p, cfp ph, cfp alt, cfp t_2, cfp t_1, cfp c2a, fp pm1,
                                cfp mu, cfp mut) {
  TID3_V4;
  (void)hi;
  if (him < lo) return;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  float m[4], mt[4], altv[4], t1[4], ph1[4], ph0[4], c2[4], t2[4], alv[4], pv[4], pold[4], ptmp[4];
  ld4(&A2(mu, i, j), m); ld4(&A2(mut, i, j), mt); ld4(&A3(alt, i, k, j), altv); ld4(&A3(t_1, i, k, j), t1);
  ld4(&A3(ph, i, k + 1, j), ph1); ld4(&A3(ph, i, k, j), ph0); ld4(&A3(c2a, i, k, j), c2); ld4(&A3(t_2, i, k, j), t2);
  if (step != 0) ld4(&A3(pm1, i, k, j), pold);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    alv[q] = -1.f / (c1h * mt[q] + c2h) * (altv[q] * (c1h * m[q]) + g.rdnw[k] * (ph1[q] - ph0[q]));
    pv[q] = c2[q] * (altv[q] * (t2[q] - (c1h * m[q]) * t1[q]) / ((c1h * mt[q] + c2h) * (WG_T0 + t1[q])) - alv[q]);
    ptmp[q] = pv[q];
    if (step != 0) { pv[q] = pv[q] + g.smdiv * (pv[q] - pold[q]); ptmp[q] = after_load(ptmp[q], pold[q]); }
  }
  st4(&A3(al, i, k, j), alv, lo, him);
  st4(&A3(pm1, i, k, j), ptmp, lo, him);
  st4(&A3(p, i, k, j), pv, lo, him);
}

// calc_coef_w (:570-650): column recurrence for the tridiagonal factors
__global__ void k_calc_coef_w(const Geo g, Box b, float dts, fp a, fp alpha, fp gamma, cfp mut, cfp cqw, cfp c2a) {
  COEF_SMEM;
  TID2;
  const int kde = g.kde;
  const float lid_flag = g.top_lid ? 0.f : 1.f;
  const float x0 = .5f * dts * WG_G * (1.f + g.epssm);
  const float cof = x0 * x0;
  const float mt = A2(mut, i, j);
  {
    const int k = kde - 1;
    A3(a, i, 2, j) = 0.f;
    A3(a, i, kde, j) = -2.f * cof * (cf.rdnw[kde - 1] * cf.rdnw[kde - 1]) * A3(c2a, i, kde - 1, j) * lid_flag /
                       ((cf.c1h[k] * mt + cf.c2h[k]) * (cf.c1f[k] * mt + cf.c2f[k]));
    A3(gamma, i, 1, j) = 0.f;
  }
  float gam_prev = 0.f;   // gamma(k-1)
  float c2a_km1 = A3(c2a, i, 1, j);
  for (int k = 2; k <= kde - 1; k++) {
    const float cq = A3(cqw, i, k, j);
    const float c2a_k = A3(c2a, i, k, j);
    float ak;
    if (k == 2) ak = 0.f;
    else ak = -cq * cof * cf.rdn[k] * cf.rdnw[k - 1] * c2a_km1 / ((cf.c1h[k - 1] * mt + cf.c2h[k - 1]) * (cf.c1f[k - 1] * mt + cf.c2f[k - 1]));
    if (k > 2) A3(a, i, k, j) = ak;
    const float bb = 1.f + cq * cof * cf.rdn[k] *
                               (cf.rdnw[k] * c2a_k / ((cf.c1h[k] * mt + cf.c2h[k]) * (cf.c1f[k] * mt + cf.c2f[k])) +
                                cf.rdnw[k - 1] * c2a_km1 / ((cf.c1h[k - 1] * mt + cf.c2h[k - 1]) * (cf.c1f[k] * mt + cf.c2f[k])));
    const float cc = -cq * cof * cf.rdn[k] * cf.rdnw[k] * c2a_k / ((cf.c1h[k] * mt + cf.c2h[k]) * (cf.c1f[k + 1] * mt + cf.c2f[k + 1]));
    const float al = 1.f / (bb - ak * gam_prev);
    gam_prev = cc * al;
    A3(alpha, i, k, j) = al;
    A3(gamma, i, k, j) = gam_prev;
    c2a_km1 = c2a_k;
  }
  {
    const int k = kde;
    const float bb = 1.f + 2.f * cof * (cf.rdnw[kde - 1] * cf.rdnw[kde - 1]) * A3(c2a, i, kde - 1, j) /
                               ((cf.c1h[k - 1] * mt + cf.c2h[k - 1]) * (cf.c1f[k] * mt + cf.c2f[k]));
    const float cc = 0.f;
    const float al = 1.f / (bb - A3(a, i, kde, j) * gam_prev);
    A3(alpha, i, kde, j) = al;
    A3(gamma, i, kde, j) = cc * al;
  }
}

// calc_coef_w with four columns per thread: 16-byte accesses and four independent recurrences in flight
__global__ void k_calc_coef_w_v4(const Geo g, Box b, float dts, fp a, fp alpha, fp gamma, cfp mut, cfp cqw, cfp c2a) {
  COEF_SMEM;
  TID2_V4;
  if (him < lo) return;
  const int kde = g.kde;
  const float lid_flag = g.top_lid ? 0.f : 1.f;
  const float x0 = .5f * dts * WG_G * (1.f + g.epssm);
  const float cof = x0 * x0;
  float mt[4], c2top[4], out[4], zero[4] = {0.f, 0.f, 0.f, 0.f}, a_top[4];
  ld4(&A2(mut, i, j), mt); ld4(&A3(c2a, i, kde - 1, j), c2top);
  {
    const int k = kde - 1;
    st4(&A3(a, i, 2, j), zero, lo, him);
#pragma unroll
    for (int q = 0; q < 4; q++)
      a_top[q] = -2.f * cof * (cf.rdnw[kde - 1] * cf.rdnw[kde - 1]) * c2top[q] * lid_flag / ((cf.c1h[k] * mt[q] + cf.c2h[k]) * (cf.c1f[k] * mt[q] + cf.c2f[k]));
    st4(&A3(a, i, kde, j), a_top, lo, him);
    st4(&A3(gamma, i, 1, j), zero, lo, him);
  }
  float gam_prev[4] = {0.f, 0.f, 0.f, 0.f}, c2a_km1[4];
  ld4(&A3(c2a, i, 1, j), c2a_km1);
  for (int k = 2; k <= kde - 1; k++) {
    float cq[4], c2a_k[4], ak[4], al[4];
    ld4(&A3(cqw, i, k, j), cq); ld4(&A3(c2a, i, k, j), c2a_k);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (k == 2) ak[q] = 0.f;
      else ak[q] = -cq[q] * cof * cf.rdn[k] * cf.rdnw[k - 1] * c2a_km1[q] / ((cf.c1h[k - 1] * mt[q] + cf.c2h[k - 1]) * (cf.c1f[k - 1] * mt[q] + cf.c2f[k - 1]));
      const float bb = 1.f + cq[q] * cof * cf.rdn[k] *
                                 (cf.rdnw[k] * c2a_k[q] / ((cf.c1h[k] * mt[q] + cf.c2h[k]) * (cf.c1f[k] * mt[q] + cf.c2f[k])) +
                                  cf.rdnw[k - 1] * c2a_km1[q] / ((cf.c1h[k - 1] * mt[q] + cf.c2h[k - 1]) * (cf.c1f[k] * mt[q] + cf.c2f[k])));
      const float cc = -cq[q] * cof * cf.rdn[k] * cf.rdnw[k] * c2a_k[q] / ((cf.c1h[k] * mt[q] + cf.c2h[k]) * (cf.c1f[k + 1] * mt[q] + cf.c2f[k + 1]));
      al[q] = 1.f / (bb - ak[q] * gam_prev[q]);
      gam_prev[q] = cc * al[q];
      c2a_km1[q] = c2a_k[q];
    }
    if (k > 2) st4(&A3(a, i, k, j), ak, lo, him);
    st4(&A3(alpha, i, k, j), al, lo, him);
    st4(&A3(gamma, i, k, j), gam_prev, lo, him);
  }
  {
    const int k = kde;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float bb = 1.f + 2.f * cof * (cf.rdnw[kde - 1] * cf.rdnw[kde - 1]) * c2top[q] / ((cf.c1h[k - 1] * mt[q] + cf.c2h[k - 1]) * (cf.c1f[k] * mt[q] + cf.c2f[k]));
      const float cc = 0.f;
      const float al = 1.f / (bb - a_top[q] * gam_prev[q]);
      out[q] = al;
      gam_prev[q] = cc * al;
    }
    st4(&A3(alpha, i, kde, j), out, lo, him);
    st4(&A3(gamma, i, kde, j), gam_prev, lo, him);
  }
}

// ---------------------------------------------------------------------------
struct UvArgs {
  fp u, v;
  cfp ru_tend, rv_tend, p, pb, ph, php, alt, al, mu, muu, cqu, muv, cqv, mudf;
};
__device__ __forceinline__ float dpn_uv(const Geo &g, cfp p, int i, int kk, int j, int di, int dj) {
  const int i2 = i - di, j2 = j - dj;
  if (kk == 1)
    return .5f * (g.cf1 * (A3(p, i, 1, j) + A3(p, i2, 1, j2)) + g.cf2 * (A3(p, i, 2, j) + A3(p, i2, 2, j2)) + g.cf3 * (A3(p, i, 3, j) + A3(p, i2, 3, j2)));
  if (kk == g.kde) {
    if (!g.top_lid) return 0.f;
    return .5f * (g.cf1 * (A3(p, i2, g.kde - 1, j2) + A3(p, i, g.kde - 1, j)) + g.cf2 * (A3(p, i2, g.kde - 2, j2) + A3(p, i, g.kde - 2, j)) +
                  g.cf3 * (A3(p, i2, g.kde - 3, j2) + A3(p, i, g.kde - 3, j)));
  }
  return .5f * (g.fnm[kk] * (A3(p, i, kk, j) + A3(p, i2, kk, j2)) + g.fnp[kk] * (A3(p, i, kk - 1, j) + A3(p, i2, kk - 1, j2)));
}

// advance_uv (:654-965) as a k-marching column kernel: one thread per (i,j) owns the u point and
// the v point of that column and walks k = 1..kde-1.  Everything with a vertical stencil -- the
// pressure at w levels dpn(k), dpn(k+1) (built from p at k-1,k,k+1 of the two columns either side
// of the point) and the geopotential differences at k, k+1 -- lives in rolling registers, so each
// level costs 24 loads instead of the ~50 a point-per-thread formulation needs (measured:
// 188 M L1 sectors per launch for 1.1 GB of DRAM traffic, LSU-bound at 20 % DRAM; profiles/r1a).
// Loads of level k+1 are issued before level k is computed (software prefetch: the loop has no
// other way to cover HBM latency).  Expressions are the reference's, operand for operand.
//
// FUSE: also does the u- and v-part of sumflux (:1473-1633) -- ru_m += u'', rv_m += v'' -- which in
// the reference runs after advance_w but reads the u'', v'' produced here (nothing in between
// changes them on owned points).  Only wrfgpu_step uses the fused form; the per-routine entry
// points keep the reference's split.
struct UvLevel {      // everything read at one level: c = (i,j), w = (i-1,j), s = (i,j-1)
  float alt_c, alt_w, alt_s, al_c, al_w, al_s, pb_c, pb_w, pb_s, php_c, php_w, php_s;
  float cqu, cqv, rut, rvt, u, v;
  float rum, rvm, ulin, vlin;   // FUSE: the sumflux operands, prefetched with the level (a load issued at its
                                // point of use costs one full memory round trip per level: ncu r1c, 31 % of the samples)
};
struct UvPp { float p_c, p_w, p_s, ph_c, ph_w, ph_s; };

struct UvFuse { int iteration, nsmall; fp ru_m, rv_m; cfp u_lin, v_lin; };

template <bool FUSE>
__device__ __forceinline__ UvLevel uv_load_level(const Geo &g, const UvArgs &a, const UvFuse &fz, int i, int k, int j) {
  UvLevel L;
  L.rum = 0.f; L.rvm = 0.f; L.ulin = 0.f; L.vlin = 0.f;
  if (FUSE) {
    if (fz.iteration != 1) { L.rum = A3(fz.ru_m, i, k, j); L.rvm = A3(fz.rv_m, i, k, j); }
    if (fz.iteration == fz.nsmall) { L.ulin = A3(fz.u_lin, i, k, j); L.vlin = A3(fz.v_lin, i, k, j); }
  }
  L.alt_c = A3(a.alt, i, k, j); L.alt_w = A3(a.alt, i - 1, k, j); L.alt_s = A3(a.alt, i, k, j - 1);
  L.al_c = A3(a.al, i, k, j); L.al_w = A3(a.al, i - 1, k, j); L.al_s = A3(a.al, i, k, j - 1);
  L.pb_c = A3(a.pb, i, k, j); L.pb_w = A3(a.pb, i - 1, k, j); L.pb_s = A3(a.pb, i, k, j - 1);
  L.php_c = A3(a.php, i, k, j); L.php_w = A3(a.php, i - 1, k, j); L.php_s = A3(a.php, i, k, j - 1);
  L.cqu = A3(a.cqu, i, k, j); L.cqv = A3(a.cqv, i, k, j);
  L.rut = A3(a.ru_tend, i, k, j); L.rvt = A3(a.rv_tend, i, k, j);
  L.u = A3(a.u, i, k, j); L.v = A3(a.v, i, k, j);
  return L;
}
__device__ __forceinline__ UvPp uv_load_pp(const Geo &g, const UvArgs &a, int i, int k, int j) {
  UvPp q;
  q.p_c = A3(a.p, i, k, j); q.p_w = A3(a.p, i - 1, k, j); q.p_s = A3(a.p, i, k, j - 1);
  q.ph_c = A3(a.ph, i, k, j); q.ph_w = A3(a.ph, i - 1, k, j); q.ph_s = A3(a.ph, i, k, j - 1);
  return q;
}

template <bool FUSE>
#ifndef UV_MINB
#define UV_MINB 5
#endif
#ifndef W_MINB
#define W_MINB 7
#endif
#ifndef MT_MINB
#define MT_MINB 8
#endif
__global__ void __launch_bounds__(128, UV_MINB) k_advance_uv(const Geo g, Box b, float dts, UvArgs a, UvFuse fz) {
  COEF_CONST;
  TID2;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, kde = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  int i_start = its, i_end = min(ite, ide - 1), j_start = jts, j_end = min(jte, jde - 1), i_endu = ite, j_endv = jte;
  if (g.spec) {   // :734-751: the specified rows are not advanced here (spec_bdyupdate does it)
    const int sz = g.spec_zone;
    i_start = max(its, ids + sz); i_end = min(ite, ide - sz - 1);
    j_start = max(jts, jds + sz); j_end = min(jte, jde - sz - 1);
    i_endu = min(ite, ide - sz); j_endv = min(jte, jde - sz);
    if (g.periodic_x) { i_start = its; i_end = min(ite, ide - 1); i_endu = ite; }
  }
  const float dx = 1.f / g.rdx, dy = 1.f / g.rdy;
  // ---- which of the reference's loop ranges this column is in (:734-797)
  int i_start_up = i_start, i_end_up = i_endu, i_s_t = i_start, i_e_t = i_endu;
  if ((g.open_xs || g.sym_xs) && its == ids) i_start_up++;
  if ((g.open_xe || g.sym_xe) && ite == ide) i_end_up--;
  if (g.sym_xs && its == ids) i_s_t++;
  if (g.sym_xe && ite == ide) i_e_t--;
  int j_start_vp = j_start, j_end_vp = j_endv, j_s_t = j_start, j_e_t = j_endv;
  if ((g.open_ys || g.sym_ys) && jts == jds) j_start_vp++;
  if ((g.open_ye || g.sym_ye) && jte == jde) j_end_vp--;
  if (g.sym_ys && jts == jds) j_s_t++;
  if (g.sym_ye && jte == jde) j_e_t--;
  const bool u_row = j >= j_start && j <= j_end, v_col = i >= i_start && i <= i_end;
  const bool u_tend = u_row && i >= i_s_t && i <= i_e_t, u_pg = u_row && i >= i_start_up && i <= i_end_up;
  const bool v_tend = v_col && j >= j_s_t && j <= j_e_t, v_pg = v_tend && j >= j_start_vp && j <= j_end_vp;
  if (!FUSE && !u_tend && !u_pg && !v_tend) return;
  // ---- column constants
  const float msfux = A2(g.msfux, i, j), msfuy = A2(g.msfuy, i, j), msfvx = A2(g.msfvx, i, j), msfvy = A2(g.msfvy, i, j);
  const float msfvx_inv = A2(g.msfvx_inv, i, j);
  const float mu_c = A2(a.mu, i, j), mu_w = A2(a.mu, i - 1, j), mu_s = A2(a.mu, i, j - 1);
  const float muu = A2(a.muu, i, j), muv = A2(a.muv, i, j);
  const float mudf_c = A2(a.mudf, i, j), mudf_w = A2(a.mudf, i - 1, j), mudf_s = A2(a.mudf, i, j - 1);
  const float mudf_x = -g.emdiv * dx * (mudf_c - mudf_w) / msfuy;
  const float mudf_y = -g.emdiv * dy * (mudf_c - mudf_s) * msfvx_inv;
  const float mr_u = msfux / msfuy, mr_v = msfvy / msfvx;
  const float nst = (float)fz.nsmall;
  // ---- prime the pipeline: p, ph at levels 1, 2 (and p at 3 for the surface extrapolation)
  UvPp q1 = uv_load_pp(g, a, i, 1, j), q2 = uv_load_pp(g, a, i, 2, j);
  const float p3_c = A3(a.p, i, 3, j), p3_w = A3(a.p, i - 1, 3, j), p3_s = A3(a.p, i, 3, j - 1);
  UvLevel L = uv_load_level<FUSE>(g, a, fz, i, 1, j);
  // dpn(1) (:828-836, :902-910): cf1..3 extrapolation of p to the surface
  float dpn_u = .5f * (g.cf1 * (q1.p_c + q1.p_w) + g.cf2 * (q2.p_c + q2.p_w) + g.cf3 * (p3_c + p3_w));
  float dpn_v = .5f * (g.cf1 * (q1.p_c + q1.p_s) + g.cf2 * (q2.p_c + q2.p_s) + g.cf3 * (p3_c + p3_s));
  UvPp qk = q1, qk1 = q2;           // levels k and k+1
  for (int k = 1; k <= kde - 1; k++) {
    // prefetch: the rest of level k+1 and p, ph of level k+2
    UvLevel Ln;
    UvPp qk2;
    if (k + 1 <= kde - 1) Ln = uv_load_level<FUSE>(g, a, fz, i, k + 1, j);
    if (k + 2 <= kde) qk2 = uv_load_pp(g, a, i, k + 2, j);
    const float c1h = cf.c1h[k], c2h = cf.c2h[k];
    // dpn(k+1)
    float dpn_u1, dpn_v1;
    if (k + 1 == kde) {
      if (!g.top_lid) { dpn_u1 = 0.f; dpn_v1 = 0.f; }
      else {
        cfp p = a.p;
        dpn_u1 = .5f * (g.cf1 * (A3(p, i - 1, kde - 1, j) + A3(p, i, kde - 1, j)) + g.cf2 * (A3(p, i - 1, kde - 2, j) + A3(p, i, kde - 2, j)) +
                        g.cf3 * (A3(p, i - 1, kde - 3, j) + A3(p, i, kde - 3, j)));
        dpn_v1 = .5f * (g.cf1 * (A3(p, i, kde - 1, j - 1) + A3(p, i, kde - 1, j)) + g.cf2 * (A3(p, i, kde - 2, j - 1) + A3(p, i, kde - 2, j)) +
                        g.cf3 * (A3(p, i, kde - 3, j - 1) + A3(p, i, kde - 3, j)));
      }
    } else {
      dpn_u1 = .5f * (cf.fnm[k + 1] * (qk1.p_c + qk1.p_w) + cf.fnp[k + 1] * (qk.p_c + qk.p_w));
      dpn_v1 = .5f * (cf.fnm[k + 1] * (qk1.p_c + qk1.p_s) + cf.fnp[k + 1] * (qk.p_c + qk.p_s));
    }
    // ---- u (:799-870)
    float uv = L.u;
    if (u_tend) uv = uv + dts * L.rut;
    if (u_pg) {
      float dpxy = mr_u * .5f * g.rdx * (c1h * muu + c2h) *
                   (((qk1.ph_c - qk1.ph_w) + (qk.ph_c - qk.ph_w)) + (L.alt_c + L.alt_w) * (qk.p_c - qk.p_w) +
                    (L.al_c + L.al_w) * (L.pb_c - L.pb_w));
      dpxy = dpxy + mr_u * g.rdx * (L.php_c - L.php_w) * (cf.rdnw[k] * (dpn_u1 - dpn_u) - .5f * ((c1h * mu_w) + (c1h * mu_c)));
      uv = uv - dts * L.cqu * dpxy + (c1h * mudf_x);
    }
    if (u_tend || u_pg) A3(a.u, i, k, j) = uv;
    // ---- v (:872-944)
    float vv = L.v;
    if (v_tend) {
      vv = vv + dts * L.rvt;
      if (v_pg) {
        float dpxy = mr_v * .5f * g.rdy * (c1h * muv + c2h) *
                     (((qk1.ph_c - qk1.ph_s) + (qk.ph_c - qk.ph_s)) + (L.alt_c + L.alt_s) * (qk.p_c - qk.p_s) +
                      (L.al_c + L.al_s) * (L.pb_c - L.pb_s));
        dpxy = dpxy + mr_v * g.rdy * (L.php_c - L.php_s) * (cf.rdnw[k] * (dpn_v1 - dpn_v) - .5f * ((c1h * mu_s) + (c1h * mu_c)));
        vv = vv - dts * L.cqv * dpxy + (c1h * mudf_y);
      }
      A3(a.v, i, k, j) = vv;
    }
    if (FUSE) {   // sumflux, u and v parts
      float ra = L.rum, rb = L.rvm;
      if (u_row) ra = ra + uv;
      if (v_col) rb = rb + vv;
      if (fz.iteration == fz.nsmall) {
        if (u_row) ra = ra / nst + (c1h * muu + c2h) * L.ulin / msfuy;
        if (v_col) rb = rb / nst + (c1h * muv + c2h) * L.vlin * msfvx_inv;
      }
      A3(fz.ru_m, i, k, j) = ra; A3(fz.rv_m, i, k, j) = rb;
    }
    dpn_u = dpn_u1; dpn_v = dpn_v1;
    L = Ln; qk = qk1; qk1 = qk2;
  }
  (void)msfvx; (void)msfvy;
}

// ---------------------------------------------------------------------------
struct MuTArgs {
  fp ww, mu, muave, muts, mudf, t, t_ave;
  cfp ww_1, u, u_1, v, v_1, mut, muu, muv, t_1, ft, mu_tend;
};
// advance_mu_t (:969-1174), column kernel.  Both k loops prefetch the next level's operands into
// registers before the current level is computed: the loops carry a recurrence (dmdt, ww, wdtn), so
// nothing else keeps enough loads in flight to cover HBM latency (ncu r1a: 26 % DRAM, latency-bound).
struct Mt1 { float vn, vsn, v0, vs0, ue, use, u0, us0; };
struct Mt2 { float dv, ww1, t1u, t, ft, vn, v0, t1n, t1s, ue, u0, t1e, t1w; };
__device__ __forceinline__ Mt1 mt_load1(const Geo &g, const MuTArgs &a, int i, int k, int j) {
  Mt1 m;
  m.vn = A3(a.v, i, k, j + 1); m.vsn = A3(a.v_1, i, k, j + 1); m.v0 = A3(a.v, i, k, j); m.vs0 = A3(a.v_1, i, k, j);
  m.ue = A3(a.u, i + 1, k, j); m.use = A3(a.u_1, i + 1, k, j); m.u0 = A3(a.u, i, k, j); m.us0 = A3(a.u_1, i, k, j);
  return m;
}
__device__ __forceinline__ Mt2 mt_load2(const Geo &g, const MuTArgs &a, int i, int k, int j, int k_end) {
  Mt2 m;
  m.dv = 0.f; m.ww1 = 0.f; m.t1u = 0.f;
  if (k + 1 <= k_end) { m.dv = A3(a.ww, i, k + 1, j); m.ww1 = A3(a.ww_1, i, k + 1, j); m.t1u = A3(a.t_1, i, k + 1, j); }
  m.t = A3(a.t, i, k, j); m.ft = A3(a.ft, i, k, j);
  m.vn = A3(a.v, i, k, j + 1); m.v0 = A3(a.v, i, k, j); m.t1n = A3(a.t_1, i, k, j + 1); m.t1s = A3(a.t_1, i, k, j - 1);
  m.ue = A3(a.u, i + 1, k, j); m.u0 = A3(a.u, i, k, j); m.t1e = A3(a.t_1, i + 1, k, j); m.t1w = A3(a.t_1, i - 1, k, j);
  return m;
}
// FINAL_AVE (the form wrfgpu_step uses, paired with k_advance_w<true>): t_ave receives the finished time average
// t_2ave of advance_w (:1314-1317) instead of the old theta'' -- this thread holds every operand of it (old and new
// theta'', t_1, muave, muts), and advance_w then reads one field in its upward sweep instead of reading t_2, t_1,
// t_2ave and writing t_2ave back: 12 B per point and launch less DRAM traffic in the DRAM-bound kernel.
// ld1(k) / ld2(k) deliver the operands of level k of the two sweeps (Mt1 / Mt2): straight from global memory
// (k_advance_mu_t) or from the shared-memory ring the bulk-copy engine fills (k_advance_mu_t_st).  `active` = false:
// a thread that only keeps a block-wide pipeline in step (it computes a copy of a neighbouring column and stores nothing).
template <bool FINAL_AVE, class Ld1, class Ld2, class Mid>
__device__ __forceinline__ void advance_mu_t_column(const Geo &g, const Coef &cf, float dts, const MuTArgs &a, int i, int j, bool active, Ld1 ld1,
                                                    Ld2 ld2, Mid between_sweeps) {
  const int k_end = g.kde - 1;
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float muu0 = A2(a.muu, i, j), muu1 = A2(a.muu, i + 1, j), muv0 = A2(a.muv, i, j), muv1 = A2(a.muv, i, j + 1);
  const float my0 = A2(g.msfuy, i, j), my1 = A2(g.msfuy, i + 1, j), mvi0 = A2(g.msfvx_inv, i, j), mvi1 = A2(g.msfvx_inv, i, j + 1);
  const float mu_tend = A2(a.mu_tend, i, j);
  // pass 1: dvdxi(k) (parked in ww(k+1)) and its column integral dmdt
  float dmdt = 0.f;
  Mt1 c1 = ld1(1);
  for (int k = 1; k <= k_end; k++) {
    Mt1 n1;
    if (k < k_end) n1 = ld1(k + 1);
    const float c1h = cf.c1h[k], c2h = cf.c2h[k];
    const float dvdxi = msftx * msfty *
                        (g.rdy * ((c1.vn + (c1h * muv1 + c2h) * c1.vsn * mvi1) - (c1.v0 + (c1h * muv0 + c2h) * c1.vs0 * mvi0)) +
                         g.rdx * ((c1.ue + (c1h * muu1 + c2h) * c1.use / my1) - (c1.u0 + (c1h * muu0 + c2h) * c1.us0 / my0)));
    dmdt = dmdt + cf.dnw[k] * dvdxi;
    if (k < k_end && active) A3(a.ww, i, k + 1, j) = dvdxi;
    c1 = n1;
  }
  float muts_new, muave_new;
  {
    const float mu_old = A2(a.mu, i, j);
    const float mu_new = mu_old + dts * (dmdt + mu_tend);
    muts_new = A2(a.mut, i, j) + mu_new;
    muave_new = .5f * ((1.f + g.epssm) * mu_new + (1.f - g.epssm) * mu_old);
    if (active) {
      A2(a.mu, i, j) = mu_new;
      A2(a.mudf, i, j) = (dmdt + mu_tend);
      A2(a.muts, i, j) = muts_new;
      A2(a.muave, i, j) = muave_new;
    }
  }
  between_sweeps();
  // pass 2: ww recurrence (:1109-1121), then the theta update which needs ww(k), ww(k+1)
  // ww(1) is not touched by the recurrence; ww(k) -= ww_1(k) for k = 1..k_end
  float w_rec = A3(a.ww, i, 1, j);             // ww(i,1,j) before the subtraction
  if (active) A3(a.ww, i, 1, j) = w_rec - A3(a.ww_1, i, 1, j);
  float wdtn_k = 0.f;                          // wdtn(1) = 0
  float t1_k = A3(a.t_1, i, 1, j);
  Mt2 c2 = ld2(1);
  for (int k = 1; k <= k_end; k++) {
    Mt2 n2;
    if (k < k_end) n2 = ld2(k + 1);
    float wdtn_kp1 = 0.f;
    if (k + 1 <= k_end) {
      w_rec = w_rec - cf.dnw[k] * ((cf.c1h[k] * dmdt) + c2.dv + (cf.c1h[k] * mu_tend)) / msfty;
      const float wkp1 = w_rec - c2.ww1;
      if (active) A3(a.ww, i, k + 1, j) = wkp1;
      wdtn_kp1 = wkp1 * (cf.fnm[k + 1] * c2.t1u + cf.fnp[k + 1] * t1_k);
    }   // else wdtn(kde) = 0
    const float told = c2.t;
    if (!FINAL_AVE && active) A3(a.t_ave, i, k, j) = told;
    float tn = told + msfty * dts * c2.ft;
    tn = tn - dts * msfty *
                  (msftx * (.5f * g.rdy * (c2.vn * (c2.t1n + t1_k) - c2.v0 * (t1_k + c2.t1s)) +
                            .5f * g.rdx * (c2.ue * (c2.t1e + t1_k) - c2.u0 * (t1_k + c2.t1w))) +
                   cf.rdnw[k] * (wdtn_kp1 - wdtn_k));
    if (active) A3(a.t, i, k, j) = tn;
    if (FINAL_AVE) {   // t_2ave(k) of advance_w (:1314-1317)
      float tv = .5f * ((1.f + g.epssm) * tn + (1.f - g.epssm) * told);
      tv = (tv + (cf.c1h[k] * muave_new) * WG_T0) / ((cf.c1h[k] * muts_new + cf.c2h[k]) * (WG_T0 + t1_k));
      if (active) A3(a.t_ave, i, k, j) = tv;
    }
    wdtn_k = wdtn_kp1;
    t1_k = c2.t1u;
    c2 = n2;
  }
}
#ifndef MT_COEF
#define MT_COEF COEF_SMEM     // r1u: 46.8 vs 47.7 ms per step with the constant bank
#endif
template <bool FINAL_AVE>
__global__ void __launch_bounds__(128, MT_MINB) k_advance_mu_t(const Geo g, Box b, float dts, MuTArgs a) {
  MT_COEF;
  TID2;
  const int k_end = g.kde - 1;
  advance_mu_t_column<FINAL_AVE>(
      g, cf, dts, a, i, j, true, [&](int k) { return mt_load1(g, a, i, k, j); }, [&](int k) { return mt_load2(g, a, i, k, j, k_end); }, [] {});
}

// ---------------------------------------------------------------------------
// The same column march with its operands staged by the bulk-copy engine (cp.async.bulk, the 1-D form of TMA): a block is
// 128 columns of one j row; per level one thread issues one 544-byte row copy per operand stream (6 in the first sweep,
// 10 in the second) into a ring of MTS_STAGES levels, an mbarrier per slot counts the bytes in, and every thread then
// takes its operands -- its own column and the i-1 / i+1 neighbours -- from shared memory.  The register prefetch of
// k_advance_mu_t keeps one level of loads in flight per warp and pays for it in registers; the ring keeps MTS_STAGES - 1
// levels in flight without holding a register.  Arithmetic: advance_mu_t_column, unchanged.
// Measured at 2000x2000x80 (profiles/README.md, mts1-mts8): the register-prefetch kernel 46.1 ms per step; staged with a
// block-wide barrier per level 42.4; with a producer warp and full / empty mbarriers (below) 38.9 at 3 stages, 37.2 at 4
// (the default: 8 consumer warps per block, 4 blocks per SM), 40.3 at 5 or 6 stages (3 blocks), 37.9 with 16 consumer warps.
#ifndef MTS_STAGES
#define MTS_STAGES 4
#endif
#ifndef MTS_CW
#define MTS_CW 8            // consumer warps per block (32 columns each); one more warp issues the copies
#endif
#define MTS_NCOL (32 * MTS_CW)
#define MTS_RW (MTS_NCOL + 8)   // floats per staged row: the columns + the i-1 / i+1 neighbours + the slack of a 16-byte aligned start
#define MTS_ROWS 10
#ifndef MTS_MINB
#define MTS_MINB 4
#endif
__device__ __forceinline__ uint32_t mts_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mts_bulk(uint32_t dst, const float *src, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"((uint32_t)(MTS_RW * sizeof(float))), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mts_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// bounded wait: a lost completion becomes a trap (an error the host reports), never a hung GPU
__device__ __forceinline__ void mts_wait(uint32_t bar, uint32_t parity) {
#ifndef MTS_UNROLL_WAIT
#pragma unroll 1
#endif
  for (int spin = 0; spin < (1 << 24); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// Roles: warps 0 .. MTS_CW-1 march their 32 columns each; the last warp's lane 0 is the producer.  The two sweeps are one
// sequence of 2 k_end "events" (event e < k_end: level e+1 of the first sweep, then the levels of the second); event e
// lives in slot e % MTS_STAGES.  full[q] (count 1 + the bytes of the copies) tells the consumers that a slot has landed,
// empty[q] (count MTS_CW: lane 0 of every consumer warp) tells the producer that they are done with it, `sweep` that the
// first sweep's stores of dvdxi into ww are visible to the copy engine.  No block-wide barrier inside the march: the warps
// drift by up to a ring's depth (the first version synchronised the block once per level and spent a third of its stall
// cycles there, ncu mtsp).
template <bool FINAL_AVE>
__global__ void __launch_bounds__(32 * (MTS_CW + 1), MTS_MINB) k_advance_mu_t_st(const Geo g, Box b, float dts, MuTArgs a) {
  MT_COEF;
  extern __shared__ __align__(128) float mts_dyn[];   // the ring: more than the 48 KB a static array may take
  float (*ring)[MTS_ROWS][MTS_RW] = reinterpret_cast<float (*)[MTS_ROWS][MTS_RW]>(mts_dyn);
  __shared__ __align__(8) unsigned long long full[MTS_STAGES], empty[MTS_STAGES], sweep;
  const int tx = threadIdx.x, lane = tx & 31;
  const int i0 = b.i0 + blockIdx.x * MTS_NCOL, j = b.j0 + blockIdx.y;
  const int s0 = (i0 - 1 - g.ims) & ~3;              // 16-byte aligned start of the staged rows, in floats from the row start
  const int k_end = g.kde - 1;
  if (tx == 0) {
    for (int q = 0; q < MTS_STAGES; q++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mts_smem(&full[q])), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mts_smem(&empty[q])), "r"((uint32_t)MTS_CW) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mts_smem(&sweep)), "r"((uint32_t)MTS_CW) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tx >= MTS_NCOL) {   // ---- producer
    if (lane != 0) return;
    auto row = [&](const float *f, int L, int jj) { return &A3(f, g.ims + s0, L, jj); };
    for (int e = 0; e < 2 * k_end; e++) {
      const int q = e % MTS_STAGES, n = e / MTS_STAGES;
      if (n > 0) mts_wait(mts_smem(&empty[q]), (uint32_t)((n - 1) & 1));
      if (e == k_end) mts_wait(mts_smem(&sweep), 0u);
      const uint32_t bar = mts_smem(&full[q]);
      if (e < k_end) {
        const int L = e + 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(6 * MTS_RW * sizeof(float))) : "memory");
        mts_bulk(mts_smem(ring[q][0]), row(a.v, L, j), bar);     mts_bulk(mts_smem(ring[q][1]), row(a.v, L, j + 1), bar);
        mts_bulk(mts_smem(ring[q][2]), row(a.v_1, L, j), bar);   mts_bulk(mts_smem(ring[q][3]), row(a.v_1, L, j + 1), bar);
        mts_bulk(mts_smem(ring[q][4]), row(a.u, L, j), bar);     mts_bulk(mts_smem(ring[q][5]), row(a.u_1, L, j), bar);
      } else {
        const int L = e - k_end + 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(10 * MTS_RW * sizeof(float))) : "memory");
        mts_bulk(mts_smem(ring[q][0]), row(a.ww, L, j), bar);      mts_bulk(mts_smem(ring[q][1]), row(a.ww_1, L, j), bar);
        mts_bulk(mts_smem(ring[q][2]), row(a.t_1, L, j - 1), bar); mts_bulk(mts_smem(ring[q][3]), row(a.t_1, L, j), bar);
        mts_bulk(mts_smem(ring[q][4]), row(a.t_1, L, j + 1), bar); mts_bulk(mts_smem(ring[q][5]), row(a.t, L, j), bar);
        mts_bulk(mts_smem(ring[q][6]), row(a.ft, L, j), bar);      mts_bulk(mts_smem(ring[q][7]), row(a.v, L, j), bar);
        mts_bulk(mts_smem(ring[q][8]), row(a.v, L, j + 1), bar);   mts_bulk(mts_smem(ring[q][9]), row(a.u, L, j), bar);
      }
    }
    return;
  }
  // ---- consumers
  const bool active = i0 + tx <= b.i1;
  const int i = active ? i0 + tx : b.i1;             // idle threads shadow the last column (and store nothing)
  const int ix = (i - g.ims) - s0;                   // this column in a staged row (>= 1)
  auto wait_event = [&](int e) { mts_wait(mts_smem(&full[e % MTS_STAGES]), (uint32_t)((e / MTS_STAGES) & 1)); };
  auto release = [&](int e) {
    __syncwarp();
    if (lane == 0) mts_arrive(mts_smem(&empty[e % MTS_STAGES]));
  };
  auto ld1 = [&](int k) {
    const int e = k - 1, q = e % MTS_STAGES;
    wait_event(e);
    Mt1 m;
    m.v0 = ring[q][0][ix]; m.vn = ring[q][1][ix]; m.vs0 = ring[q][2][ix]; m.vsn = ring[q][3][ix];
    m.u0 = ring[q][4][ix]; m.ue = ring[q][4][ix + 1]; m.us0 = ring[q][5][ix]; m.use = ring[q][5][ix + 1];
    release(e);
    return m;
  };
  auto between = [&] {
    // the first sweep parked dvdxi in ww with ordinary stores; the second sweep's bulk copies read it back
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if (lane == 0) mts_arrive(mts_smem(&sweep));
  };
  auto ld2 = [&](int k) {
    const int e = k_end + k - 1, q = e % MTS_STAGES, qn = (e + 1) % MTS_STAGES;
    wait_event(e);
    if (k + 1 <= k_end) wait_event(e + 1);
    Mt2 m;
    m.dv = 0.f; m.ww1 = 0.f; m.t1u = 0.f;
    if (k + 1 <= k_end) { m.dv = ring[qn][0][ix]; m.ww1 = ring[qn][1][ix]; m.t1u = ring[qn][3][ix]; }
    m.t1s = ring[q][2][ix]; m.t1w = ring[q][3][ix - 1]; m.t1e = ring[q][3][ix + 1]; m.t1n = ring[q][4][ix];
    m.t = ring[q][5][ix]; m.ft = ring[q][6][ix]; m.v0 = ring[q][7][ix]; m.vn = ring[q][8][ix];
    m.u0 = ring[q][9][ix]; m.ue = ring[q][9][ix + 1];
    release(e);                                            // level k's slot (level k+1's is read again by the next call)
    return m;
  };
  advance_mu_t_column<FINAL_AVE>(g, cf, dts, a, i, j, active, ld1, ld2, between);
}

// ---------------------------------------------------------------------------
struct WArgs {
  fp w, ph, t_2ave;
  cfp rw_tend, ww, w_save, u, v, mut, muave, muts, t_2, t_1, ph_1, phb, ph_tend, c2a, cqw, alt, a, alpha, gamma;
};
// what the fused form adds: calc_p_rho (:438-566) and the ww part of sumflux (:1473-1633), both
// column-local, done in the backward sweep where ph(k+1), ph(k) have just become final
struct WFuse { int iteration, nsmall; fp al, p, pm1, ww_m, ru_m, rv_m; cfp mu, ww_lin; };

// advance_w (:1178-1469), column kernel: t_2ave, rhs, the explicit part of w and the Thomas forward
// elimination in ONE upward sweep with every vertically shifted operand (rhs(k-1..k+1), ph(k-1..k+1),
// the layer integrals of -omega dphi/deta, c2a*alt*t_2ave at k-1) rolling in registers and the next
// level's 15 operands prefetched; the explicit w and rhs are parked in w / ph and finished in the
// backward sweep (back substitution, damp_opt 3, ph).  Expressions are the reference's, operand for
// operand; a value that the reference recomputes at the next level is computed once and carried,
// which is bit-identical.
struct WUp { float w_kp1, pht_kp1, ph_kp1, ph1_kp2, phb_kp2, ww_kp2, rwt, cqw, c2a, alt, t2, t2a, t1, a, alpha; };
template <bool FUSE>
__device__ __forceinline__ WUp w_load_up(const Geo &g, const WArgs &a, int i, int k, int j) {
  WUp q;
  const int kde = g.kde;
  q.w_kp1 = A3(a.w, i, k + 1, j); q.pht_kp1 = A3(a.ph_tend, i, k + 1, j); q.ph_kp1 = A3(a.ph, i, k + 1, j);
  q.ph1_kp2 = 0.f; q.phb_kp2 = 0.f; q.ww_kp2 = 0.f;
  if (k + 2 <= kde) { q.ph1_kp2 = A3(a.ph_1, i, k + 2, j); q.phb_kp2 = A3(a.phb, i, k + 2, j); q.ww_kp2 = A3(a.ww, i, k + 2, j); }
  q.rwt = A3(a.rw_tend, i, k, j); q.cqw = A3(a.cqw, i, k, j); q.c2a = A3(a.c2a, i, k, j); q.alt = A3(a.alt, i, k, j);
  q.t2a = A3(a.t_2ave, i, k, j);
  q.t2 = 0.f; q.t1 = 0.f;
  if (!FUSE) { q.t2 = A3(a.t_2, i, k, j); q.t1 = A3(a.t_1, i, k, j); }   // FUSE: t_2ave is already the finished average (k_advance_mu_t<true>)
  q.a = A3(a.a, i, k, j); q.alpha = A3(a.alpha, i, k, j);
  return q;
}
struct WDn { float w, gamma, rhs, ph1, phb, wsave, alt, t2, t1, c2a, pm1, ww, wwm, wwlin; };
template <bool FUSE>
__device__ __forceinline__ WDn w_load_dn(const Geo &g, const WArgs &a, const WFuse &fz, int i, int k, int j) {
  WDn q;
  q.w = A3(a.w, i, k, j); q.gamma = A3(a.gamma, i, k, j); q.rhs = A3(a.ph, i, k, j);
  q.ph1 = 0.f; q.phb = 0.f; q.wsave = 0.f;
  if (g.damp_opt == 3) { q.ph1 = A3(a.ph_1, i, k, j); q.phb = A3(a.phb, i, k, j); q.wsave = A3(a.w_save, i, k, j); }
  if (FUSE) {
    q.alt = A3(a.alt, i, k, j); q.t2 = A3(a.t_2, i, k, j); q.t1 = A3(a.t_1, i, k, j); q.c2a = A3(a.c2a, i, k, j);
    q.pm1 = fz.iteration != 0 ? A3(fz.pm1, i, k, j) : 0.f;
    q.ww = A3(a.ww, i, k, j);
    q.wwm = fz.iteration != 1 ? A3(fz.ww_m, i, k, j) : 0.f;
    q.wwlin = fz.iteration == fz.nsmall ? A3(fz.ww_lin, i, k, j) : 0.f;
  }
  return q;
}

template <bool FUSE>
__device__ __forceinline__ void advance_w_column(const Geo &g, const Coef &cf, float dts, const WArgs &a, const WFuse &fz, int i, int j) {
  const int kde = g.kde, k_end = g.kde - 1;
  if (FUSE) {
    // sumflux writes its whole tile: outside the mass points ww_m is only reset on the first iteration,
    // and level kde of ru_m, rv_m likewise (k_advance_uv<true> handles levels 1..kde-1)
    if (fz.iteration == 1) { A3(fz.ru_m, i, kde, j) = 0.f; A3(fz.rv_m, i, kde, j) = 0.f; }
    if (i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) {
      if (fz.iteration == 1) for (int k = 1; k <= kde; k++) A3(fz.ww_m, i, k, j) = 0.f;
      return;
    }
  }
  const float epssm = g.epssm;
  const float msfty = A2(g.msfty, i, j), msftx = A2(g.msftx, i, j);
  const float msft_inv = 1.f / msfty;
  const float mut = A2(a.mut, i, j), muave = A2(a.muave, i, j), muts = A2(a.muts, i, j);
  // surface w (:1384-1394)
  const float w1 = msfty * .5f * g.rdy *
                       ((A2(g.ht, i, j + 1) - A2(g.ht, i, j)) * (g.cf1 * A3(a.v, i, 1, j + 1) + g.cf2 * A3(a.v, i, 2, j + 1) + g.cf3 * A3(a.v, i, 3, j + 1)) +
                        (A2(g.ht, i, j) - A2(g.ht, i, j - 1)) * (g.cf1 * A3(a.v, i, 1, j) + g.cf2 * A3(a.v, i, 2, j) + g.cf3 * A3(a.v, i, 3, j))) +
                   msftx * .5f * g.rdx *
                       ((A2(g.ht, i + 1, j) - A2(g.ht, i, j)) * (g.cf1 * A3(a.u, i + 1, 1, j) + g.cf2 * A3(a.u, i + 1, 2, j) + g.cf3 * A3(a.u, i + 1, 3, j)) +
                        (A2(g.ht, i, j) - A2(g.ht, i - 1, j)) * (g.cf1 * A3(a.u, i, 1, j) + g.cf2 * A3(a.u, i, 2, j) + g.cf3 * A3(a.u, i, 3, j)));
  // ---- level 1: t_2ave(1) (:1314-1317) and the carried products
  float t2ave_km1, c2a_km1, alt_km1;
  {
    float tv;
    if (FUSE) tv = A3(a.t_2ave, i, 1, j);
    else {
      const float t1v = A3(a.t_1, i, 1, j);
      tv = .5f * ((1.f + epssm) * A3(a.t_2, i, 1, j) + (1.f - epssm) * A3(a.t_2ave, i, 1, j));
      tv = (tv + (cf.c1h[1] * muave) * WG_T0) / ((cf.c1h[1] * muts + cf.c2h[1]) * (WG_T0 + t1v));
      A3(a.t_2ave, i, 1, j) = tv;
    }
    t2ave_km1 = tv; c2a_km1 = A3(a.c2a, i, 1, j); alt_km1 = A3(a.alt, i, 1, j);
  }
  // layer integral of -omega dphi/deta: wdwn(kk+1) of the reference (:1318-1345)
  auto layer = [&](int kk, float ph1_hi, float ph1_lo, float phb_hi, float phb_lo, float ww_hi, float ww_lo) -> float {
    const float dphi = (ph1_hi - ph1_lo + phb_hi - phb_lo);
    if (g.phi_adv_z == 2) return cf.rdnw[kk] * dphi;
    return .5f * (ww_hi + ww_lo) * cf.rdnw[kk] * dphi;
  };
  // rhs(kk) for kk >= 2 from its operands (:1347-1371)
  auto rhs_of = [&](int kk, float pht, float wv, float phv, float ww_kk, float lay_kk, float lay_kkm1) -> float {
    float r = dts * (pht + .5f * WG_G * (1.f - epssm) * wv);
    if (kk <= k_end) {
      if (g.phi_adv_z == 2) r = r - dts * ww_kk * (cf.fnm[kk] * lay_kk + cf.fnp[kk] * lay_kkm1);
      else r = r - dts * (cf.fnm[kk] * lay_kk + cf.fnp[kk] * lay_kkm1);
    }
    r = phv + msfty * r / (cf.c1f[kk] * mut + cf.c2f[kk]);
    if (g.top_lid && kk == k_end + 1) r = 0.f;
    return r;
  };
  const float ph_1st = A3(a.ph, i, 1, j);
  float ph1_k, phb_k, ww_k, ph1_kp1, phb_kp1, ww_kp1, lay_km1;
  {
    const float ph1_1 = A3(a.ph_1, i, 1, j), phb_1 = A3(a.phb, i, 1, j), ww_1 = A3(a.ww, i, 1, j);
    ph1_k = A3(a.ph_1, i, 2, j); phb_k = A3(a.phb, i, 2, j); ww_k = A3(a.ww, i, 2, j);
    ph1_kp1 = A3(a.ph_1, i, 3, j); phb_kp1 = A3(a.phb, i, 3, j); ww_kp1 = A3(a.ww, i, 3, j);
    lay_km1 = layer(1, ph1_k, ph1_1, phb_k, phb_1, ww_k, ww_1);                 // layer 1
  }
  float lay_k = layer(2, ph1_kp1, ph1_k, phb_kp1, phb_k, ww_kp1, ww_k);          // layer 2
  float w_k = A3(a.w, i, 2, j);
  float ph_km1 = ph_1st, ph_k = A3(a.ph, i, 2, j);
  float rhs_km1 = 0.f;                                                           // rhs(1)
  float rhs_k = rhs_of(2, A3(a.ph_tend, i, 2, j), w_k, ph_k, ww_k, lay_k, lay_km1);
  float cpre_km1 = c2a_km1 * cf.rdnw[1] / (cf.c1h[1] * mut + cf.c2h[1]);
  float cat_km1 = c2a_km1 * alt_km1 * t2ave_km1;
  float wprev = w1;                                                              // w(k-1) after elimination
  WUp q = w_load_up<FUSE>(g, a, i, 2, j);
  for (int k = 2; k <= k_end; k++) {
    WUp qn;
    if (k < k_end) qn = w_load_up<FUSE>(g, a, i, k + 1, j);
    // t_2ave(k) (:1314-1317)
    float tv;
    if (FUSE) tv = q.t2a;
    else {
      tv = .5f * ((1.f + epssm) * q.t2 + (1.f - epssm) * q.t2a);
      tv = (tv + (cf.c1h[k] * muave) * WG_T0) / ((cf.c1h[k] * muts + cf.c2h[k]) * (WG_T0 + q.t1));
      A3(a.t_2ave, i, k, j) = tv;
    }
    // rhs(k+1): needs layer k+1 (levels k+1, k+2) unless k+1 = kde
    float lay_kp1 = 0.f;
    if (k + 1 <= k_end) lay_kp1 = layer(k + 1, q.ph1_kp2, ph1_kp1, q.phb_kp2, phb_kp1, q.ww_kp2, ww_kp1);
    const float rhs_kp1 = rhs_of(k + 1, q.pht_kp1, q.w_kp1, q.ph_kp1, ww_kp1, lay_kp1, lay_k);
    const float ph_kp1 = q.ph_kp1;
    const float cpre_k = q.c2a * cf.rdnw[k] / (cf.c1h[k] * mut + cf.c2h[k]);
    const float cat_k = q.c2a * q.alt * tv;
    float wv = w_k + dts * q.rwt +
               msft_inv * q.cqw *
                   (+.5f * dts * WG_G * cf.rdn[k] *
                    (cpre_k * ((1.f + epssm) * (rhs_kp1 - rhs_k) + (1.f - epssm) * (ph_kp1 - ph_k)) -
                     cpre_km1 * ((1.f + epssm) * (rhs_k - rhs_km1) + (1.f - epssm) * (ph_k - ph_km1)))) +
               dts * WG_G * msft_inv * (cf.rdn[k] * (cat_k - cat_km1) - (cf.c1f[k] * muave));
    wv = (wv - q.a * wprev) * q.alpha;
    A3(a.w, i, k, j) = wv;
    A3(a.ph, i, k, j) = rhs_k;         // park rhs(k); ph(k) itself is no longer needed below k+1
    wprev = wv;
    rhs_km1 = rhs_k; rhs_k = rhs_kp1;
    ph_km1 = ph_k; ph_k = ph_kp1;
    w_k = q.w_kp1;
    lay_km1 = lay_k; lay_k = lay_kp1;
    ph1_k = ph1_kp1; phb_k = phb_kp1; ww_k = ww_kp1;
    ph1_kp1 = q.ph1_kp2; phb_kp1 = q.phb_kp2; ww_kp1 = q.ww_kp2;
    cpre_km1 = cpre_k; cat_km1 = cat_k;
    c2a_km1 = q.c2a; alt_km1 = q.alt; t2ave_km1 = tv;
    q = qn;
  }
  // top level k = kde (:1417-1431)
  float wtop;
  {
    const int k = kde;
    wtop = w_k + dts * A3(a.rw_tend, i, k, j) +
           msft_inv * (-.5f * dts * WG_G / (cf.c1h[k - 1] * mut + cf.c2h[k - 1]) * (cf.rdnw[k - 1] * cf.rdnw[k - 1]) * 2.f * c2a_km1 *
                           ((1.f + epssm) * (rhs_k - rhs_km1) + (1.f - epssm) * (ph_k - ph_km1)) -
                       dts * WG_G * (2.f * cf.rdnw[k - 1] * c2a_km1 * alt_km1 * t2ave_km1 + (cf.c1f[k] * muave)));
    if (g.top_lid) wtop = 0.f;
    wtop = (wtop - A3(a.a, i, k, j) * wprev) * A3(a.alpha, i, k, j);
  }
  // backward sweep (:1439-1443) + damp_opt 3 (:1445-1458) + ph (:1460-1465)
  const float pi = 0x1.921fb6p+1f;   // 4.*atan(1.) in real*4
  const float dampmag = dts * g.dampcoef, hdepth = g.zdamp;
  const float htop = (A3(a.ph_1, i, kde, j) + A3(a.phb, i, kde, j)) / WG_G;
  // returns the final ph(k)
  auto finish = [&](int k, float wv, float rhs, float ph1v, float phbv, float wsave) -> float {
    if (g.damp_opt == 3) {
      const float hk = (ph1v + phbv) / WG_G;
      const float hbot = htop - hdepth;
      float dampwt = 0.f;
      if (hk >= hbot) dampwt = dampmag * wg_libm::sinf_(0.5f * pi * (hk - hbot) / hdepth) * wg_libm::sinf_(0.5f * pi * (hk - hbot) / hdepth);
      wv = (wv - dampwt * (cf.c1f[k] * mut + cf.c2f[k]) * wsave) / (1.f + dampwt);
    }
    A3(a.w, i, k, j) = wv;
    const float phn = rhs + msfty * .5f * dts * WG_G * (1.f + epssm) * wv / (cf.c1f[k] * muts + cf.c2f[k]);
    A3(a.ph, i, k, j) = phn;
    return phn;
  };
  // calc_p_rho at level k from the final ph(k+1), ph(k) (:522-567); sumflux ww_m (:1520-1596)
  const float mu_new = FUSE ? A2(fz.mu, i, j) : 0.f;
  auto p_rho = [&](int k, const WDn &d, float ph_hi, float ph_lo) {
    const float c1h = cf.c1h[k], c2h = cf.c2h[k];
    const float alv = -1.f / (c1h * muts + c2h) * (d.alt * (c1h * mu_new) + cf.rdnw[k] * (ph_hi - ph_lo));
    float pv = d.c2a * (d.alt * (d.t2 - (c1h * mu_new) * d.t1) / ((c1h * muts + c2h) * (WG_T0 + d.t1)) - alv);
    A3(fz.al, i, k, j) = alv;
    if (fz.iteration == 0) {
      A3(fz.pm1, i, k, j) = pv;
    } else {
      const float ptmp = pv;
      pv = pv + g.smdiv * (pv - d.pm1);
      A3(fz.pm1, i, k, j) = after_load(ptmp, d.pm1);
    }
    A3(fz.p, i, k, j) = pv;
    float cq = d.wwm + d.ww;
    if (fz.iteration == fz.nsmall) cq = cq / (float)fz.nsmall + d.wwlin;
    A3(fz.ww_m, i, k, j) = cq;
  };
  A3(a.w, i, 1, j) = w1;
  float wnext = wtop;       // un-damped w(k+1) for the back substitution
  float ph_hi;
  {
    float ph1v = 0.f, phbv = 0.f, wsv = 0.f;
    if (g.damp_opt == 3) { ph1v = A3(a.ph_1, i, kde, j); phbv = A3(a.phb, i, kde, j); wsv = A3(a.w_save, i, kde, j); }
    ph_hi = finish(kde, wtop, rhs_k, ph1v, phbv, wsv);
    if (FUSE) {   // ww_m(kde) (sumflux covers every level)
      float cq = (fz.iteration != 1 ? A3(fz.ww_m, i, kde, j) : 0.f) + A3(a.ww, i, kde, j);
      if (fz.iteration == fz.nsmall) cq = cq / (float)fz.nsmall + A3(fz.ww_lin, i, kde, j);
      A3(fz.ww_m, i, kde, j) = cq;
    }
  }
  WDn d = w_load_dn<FUSE>(g, a, fz, i, k_end, j);
  for (int k = k_end; k >= 2; k--) {
    WDn dn;
    if (k > 2) dn = w_load_dn<FUSE>(g, a, fz, i, k - 1, j);
    const float wv = d.w - d.gamma * wnext;
    wnext = wv;
    const float ph_lo = finish(k, wv, d.rhs, d.ph1, d.phb, d.wsave);
    if (FUSE) p_rho(k, d, ph_hi, ph_lo);
    ph_hi = ph_lo;
    d = dn;
  }
  if (FUSE) {   // level 1: ph(1) is not touched by advance_w
    WDn d1;
    d1.alt = A3(a.alt, i, 1, j); d1.t2 = A3(a.t_2, i, 1, j); d1.t1 = A3(a.t_1, i, 1, j); d1.c2a = A3(a.c2a, i, 1, j);
    d1.pm1 = fz.iteration != 0 ? A3(fz.pm1, i, 1, j) : 0.f;
    d1.ww = A3(a.ww, i, 1, j);
    d1.wwm = fz.iteration != 1 ? A3(fz.ww_m, i, 1, j) : 0.f;
    d1.wwlin = fz.iteration == fz.nsmall ? A3(fz.ww_lin, i, 1, j) : 0.f;
    p_rho(1, d1, ph_hi, ph_1st);
  }
}
template <bool FUSE>
__global__ void __launch_bounds__(128, W_MINB) k_advance_w(const Geo g, Box b, float dts, WArgs a, WFuse fz) {
  COEF_SMEM;
  TID2;
  advance_w_column<FUSE>(g, cf, dts, a, fz, i, j);
}

// sumflux (:1473-1633)
__global__ void k_sumflux(const Geo g, Box b, int iteration, int nsmall, cfp ru, cfp rv, cfp ww, cfp u_lin, cfp v_lin, cfp ww_lin,
                          cfp muu, cfp muv, fp ru_m, fp rv_m, fp ww_m) {
  TID3;
  const int mini = min(g.ide - 1, g.ite), minj = min(g.jde - 1, g.jte), mink = g.kde - 1;
  const float nst = (float)nsmall;
  const bool du = (j <= minj && k <= mink), dv = (i <= mini && k <= mink), dw = (i <= mini && j <= minj);
  float a = A3(ru_m, i, k, j), bq = A3(rv_m, i, k, j), cq = A3(ww_m, i, k, j);
  if (iteration == 1) { a = 0.f; bq = 0.f; cq = 0.f; }
  if (du) a = a + A3(ru, i, k, j);
  if (dv) bq = bq + A3(rv, i, k, j);
  if (dw) cq = cq + A3(ww, i, k, j);
  if (iteration == nsmall) {
    if (du) a = a / nst + (g.c1h[k] * A2(muu, i, j) + g.c2h[k]) * A3(u_lin, i, k, j) / A2(g.msfuy, i, j);
    if (dv) bq = bq / nst + (g.c1h[k] * A2(muv, i, j) + g.c2h[k]) * A3(v_lin, i, k, j) * A2(g.msfvx_inv, i, j);
    if (dw) cq = cq / nst + A3(ww_lin, i, k, j);
  }
  A3(ru_m, i, k, j) = a; A3(rv_m, i, k, j) = bq; A3(ww_m, i, k, j) = cq;
}

// calc_mu_uv_1 (module_big_step_utilities_em.F:184-321)
__global__ void k_calc_mu_uv_1(const Geo g, Box b, cfp mu, fp muu, fp muv) {
  TID2;
  if (j <= min(g.jte, g.jde - 1)) {
    float r;
    if (i == g.ids) { const int q = g.periodic_x ? g.ids - 1 : g.ids; r = 0.5f * (A2(mu, i, j) + A2(mu, q, j)); }
    else if (i == g.ide) { const int q = g.periodic_x ? g.ide : g.ide - 1; r = 0.5f * (A2(mu, i - 1, j) + A2(mu, q, j)); }
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i - 1, j));
    A2(muu, i, j) = r;
  }
  if (i <= min(g.ite, g.ide - 1)) {
    float r;
    if (j == g.jds) { const int q = g.periodic_y ? g.jds - 1 : g.jds; r = 0.5f * (A2(mu, i, j) + A2(mu, i, q)); }
    else if (j == g.jde) { const int q = g.periodic_y ? g.jde : g.jde - 1; r = 0.5f * (A2(mu, i, j - 1) + A2(mu, i, q)); }
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i, j - 1));
    A2(muv, i, j) = r;
  }
}

struct SsfArgs {
  fp u_2, v_2, w_2, t_2, ph_2, ww, mu_2;
  cfp ww1, mut, muts, muu, muus, muv, muvs, u_save, v_save, w_save, t_save, ph_save, mu_save, h_diabatic;
};
// small_step_finish (:295-434)
__global__ void k_ssf(const Geo g, Box b, int last, float dts, int nsmall, SsfArgs a) {
  TID3;
  const int kde = g.kde;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  if (km && im)
    A3(a.v_2, i, k, j) = (A2(g.msfvx, i, j) * A3(a.v_2, i, k, j) + A3(a.v_save, i, k, j) * (c1h * A2(a.muv, i, j) + c2h)) / (c1h * A2(a.muvs, i, j) + c2h);
  if (km && jm)
    A3(a.u_2, i, k, j) = (A2(g.msfuy, i, j) * A3(a.u_2, i, k, j) + A3(a.u_save, i, k, j) * (c1h * A2(a.muu, i, j) + c2h)) / (c1h * A2(a.muus, i, j) + c2h);
  if (im && jm) {
    const float mut = A2(a.mut, i, j), muts = A2(a.muts, i, j);
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    A3(a.w_2, i, k, j) = (A2(g.msfty, i, j) * A3(a.w_2, i, k, j) + A3(a.w_save, i, k, j) * (c1f * mut + c2f)) / (c1f * muts + c2f);
    A3(a.ph_2, i, k, j) = A3(a.ph_2, i, k, j) + A3(a.ph_save, i, k, j);
    A3(a.ww, i, k, j) = A3(a.ww, i, k, j) + A3(a.ww1, i, k, j);
    if (km) {
      if (!last)
        A3(a.t_2, i, k, j) = (A3(a.t_2, i, k, j) + A3(a.t_save, i, k, j) * (c1h * mut + c2h)) / (c1h * muts + c2h);
      else
        A3(a.t_2, i, k, j) = (A3(a.t_2, i, k, j) - dts * (float)nsmall * (c1h * mut + c2h) * A3(a.h_diabatic, i, k, j) +
                              A3(a.t_save, i, k, j) * (c1h * mut + c2h)) / (c1h * muts + c2h);
    }
    if (k == 1) A2(a.mu_2, i, j) = A2(a.mu_2, i, j) + A2(a.mu_save, i, j);
  }
}

// k_ssf with 16-byte accesses (see k_ssp_3d_v4): b.i0 / b.i1 count groups of four columns from the start of the memory row
__global__ void k_ssf_v4(const Geo g, Box b, int last, float dts, int nsmall, SsfArgs a) {
  const int gi = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int k = b.k0 + blockIdx.y * blockDim.y + threadIdx.y;
  const int j = b.j0 + blockIdx.z * blockDim.z + threadIdx.z;
  if (gi > b.i1 || k > b.k1 || j > b.j1) return;
  const int i = g.ims + 4 * gi;
  const int kde = g.kde;
  const bool jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  const int lo = max(g.its - i, 0), hi = min(g.ite - i, 3), him = min(min(g.ite, g.ide - 1) - i, 3);
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  float x[4], sv[4], m1[4], m2[4], ms[4], out[4];
  if (km && him >= lo) {   // v
    ld4(&A3(a.v_2, i, k, j), x); ld4(&A3(a.v_save, i, k, j), sv); ld4(&A2(a.muv, i, j), m1); ld4(&A2(a.muvs, i, j), m2); ld4(&A2(g.msfvx, i, j), ms);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = (ms[q] * x[q] + sv[q] * (c1h * m1[q] + c2h)) / (c1h * m2[q] + c2h);
    st4(&A3(a.v_2, i, k, j), out, lo, him);
  }
  if (km && jm) {   // u
    ld4(&A3(a.u_2, i, k, j), x); ld4(&A3(a.u_save, i, k, j), sv); ld4(&A2(a.muu, i, j), m1); ld4(&A2(a.muus, i, j), m2); ld4(&A2(g.msfuy, i, j), ms);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = (ms[q] * x[q] + sv[q] * (c1h * m1[q] + c2h)) / (c1h * m2[q] + c2h);
    st4(&A3(a.u_2, i, k, j), out, lo, hi);
  }
  if (jm && him >= lo) {
    float mut[4], muts[4];
    ld4(&A2(a.mut, i, j), mut); ld4(&A2(a.muts, i, j), muts);
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    ld4(&A3(a.w_2, i, k, j), x); ld4(&A3(a.w_save, i, k, j), sv); ld4(&A2(g.msfty, i, j), ms);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = (ms[q] * x[q] + sv[q] * (c1f * mut[q] + c2f)) / (c1f * muts[q] + c2f);
    st4(&A3(a.w_2, i, k, j), out, lo, him);
    ld4(&A3(a.ph_2, i, k, j), x); ld4(&A3(a.ph_save, i, k, j), sv);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = x[q] + sv[q];
    st4(&A3(a.ph_2, i, k, j), out, lo, him);
    ld4(&A3(a.ww, i, k, j), x); ld4(&A3(a.ww1, i, k, j), sv);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = x[q] + sv[q];
    st4(&A3(a.ww, i, k, j), out, lo, him);
    if (km) {
      ld4(&A3(a.t_2, i, k, j), x); ld4(&A3(a.t_save, i, k, j), sv);
      if (!last) {
#pragma unroll
        for (int q = 0; q < 4; q++) out[q] = (x[q] + sv[q] * (c1h * mut[q] + c2h)) / (c1h * muts[q] + c2h);
      } else {
        float hd[4];
        ld4(&A3(a.h_diabatic, i, k, j), hd);
#pragma unroll
        for (int q = 0; q < 4; q++)
          out[q] = (x[q] - dts * (float)nsmall * (c1h * mut[q] + c2h) * hd[q] + sv[q] * (c1h * mut[q] + c2h)) / (c1h * muts[q] + c2h);
      }
      st4(&A3(a.t_2, i, k, j), out, lo, him);
    }
    if (k == 1) {
      ld4(&A2(a.mu_2, i, j), x); ld4(&A2(a.mu_save, i, j), sv);
#pragma unroll
      for (int q = 0; q < 4; q++) out[q] = x[q] + sv[q];
      st4(&A2(a.mu_2, i, j), out, lo, him);
    }
  }
}

// ---------------------------------------------------------------------------
void Solver::small_step_prep(int rk_step, bool zero_rest) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  Box tile2{q.its, q.ite, 1, 1, q.jts, q.jte};
  LAUNCH2(this, k_ssp_2d, tile2, rk_step, o("mub"), o("mu_1"), o("mu_2"), o("muu"), o("muus"), o("muv"), o("muvs"), o("muts"),
          o("mudf"), o("mu_save"));
  SspArgs a;
  a.u_1 = o("u_1"); a.u_2 = o("u_2"); a.v_1 = o("v_1"); a.v_2 = o("v_2"); a.w_1 = o("w_1"); a.w_2 = o("w_2");
  a.t_1 = o("t_1"); a.t_2 = o("t_2"); a.ph_1 = o("ph_1"); a.ph_2 = o("ph_2");
  a.u_save = o("u_save"); a.v_save = o("v_save"); a.w_save = o("w_save"); a.t_save = o("t_save"); a.ph_save = o("ph_save");
  a.ww_save = o("ww1"); a.c2a = o("c2a"); a.ww = o("ww"); a.pb = o("pb"); a.p = o("p"); a.alt = o("alt");
  a.muu = o("muu"); a.muus = o("muus"); a.muv = o("muv"); a.muvs = o("muvs"); a.mut = o("mut"); a.muts = o("muts");
  static const bool vec4 = [] { const char *e = getenv("WRFGPU_SSP"); return !(e && !strcmp(e, "scalar")); }();
  if (vec4)   // four columns per thread, 16-byte accesses; the box counts groups of four from the start of the memory row
    LAUNCH3(this, k_ssp_3d_v4, (Box{(q.its - q.ims) / 4, (q.ite - q.ims) / 4, 1, q.kde, q.jts, q.jte}), rk_step, zero_rest ? 1 : 0, a);
  else
    LAUNCH3(this, k_ssp_3d, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), rk_step, zero_rest ? 1 : 0, a);
  const int im = std::min(q.ite, q.ide - 1), jm = std::min(q.jte, q.jde - 1);
  if (vec4) {
    Box bx = box_v4(q, 1, q.kde - 1);
    bx.j1 = jm;
    LAUNCH3(this, k_calc_p_rho_v4, bx, 0, o("al"), o("p"), o("ph_2"), o("alt"), o("t_2"), o("t_save"), o("c2a"), o("pm1"), o("mu_2"), o("muts"));
  } else
    LAUNCH3(this, k_calc_p_rho, (Box{q.its, im, 1, q.kde - 1, q.jts, jm}), 0, o("al"), o("p"), o("ph_2"), o("alt"), o("t_2"),
            o("t_save"), o("c2a"), o("pm1"), o("mu_2"), o("muts"));
  if (vec4) {
    Box bx = box_v4(q, 1, 1);
    bx.j1 = jm;
    LAUNCH2(this, k_calc_coef_w_v4, bx, dts_rk, o("a"), o("alpha"), o("gamma"), o("mut"), o("cqw"), o("c2a"));
  } else
    LAUNCH2(this, k_calc_coef_w, (Box{q.its, im, 1, 1, q.jts, jm}), dts_rk, o("a"), o("alpha"), o("gamma"), o("mut"), o("cqw"),
            o("c2a"));
  halo("HALO_EM_B");
  bc({{"ru_tend", 'u'}, {"rv_tend", 'v'}, {"ph_2", 'w'}, {"al", 'p'}, {"p", 'p'}, {"t_1", 'p'}, {"t_save", 't'},
      {"mu_1", 't'}, {"mu_2", 't'}, {"mudf", 't'}});
}

static UvArgs uv_args(Solver *s) {
  UvArgs a;
  a.u = s->o("u_2"); a.v = s->o("v_2"); a.ru_tend = s->o("ru_tend"); a.rv_tend = s->o("rv_tend"); a.p = s->o("p"); a.pb = s->o("pb");
  a.ph = s->o("ph_2"); a.php = s->o("php"); a.alt = s->o("alt"); a.al = s->o("al"); a.mu = s->o("mu_2"); a.muu = s->o("muu");
  a.cqu = s->o("cqu"); a.muv = s->o("muv"); a.cqv = s->o("cqv"); a.mudf = s->o("mudf");
  return a;
}

void Solver::advance_uv(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  UvFuse fz{0, 0, nullptr, nullptr, nullptr, nullptr};
  LAUNCH2(this, k_advance_uv<false>, (Box{q.its, q.ite, 1, 1, q.jts, q.jte}), dts_rk, uv_args(this), fz);
  if (spec()) spec_update_uv(dts_rk);   // solve_em.F:1344-1365
}

// advance_uv + the u, v part of sumflux(iteration); the ww part is k_sumflux_w (after advance_w)
void Solver::advance_uv_fused(int rk_step, int iteration) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  UvFuse fz{iteration, ns, o("ru_m"), o("rv_m"), o("u_save"), o("v_save")};
  const UvArgs ua = uv_args(this);
  // HALO_EM_C (solve_em.F:1382) overlaps the interior of the kernel that produces u, v
  overlapped(Box{q.its, q.ite, 1, 1, q.jts, q.jte}, "HALO_EM_C", [&](const Box &bx) { LAUNCH2(this, k_advance_uv<true>, bx, dts_rk, ua, fz); });
}

static MuTArgs mu_t_args(Solver *s) {
  MuTArgs a;
  a.ww = s->o("ww"); a.mu = s->o("mu_2"); a.muave = s->o("muave"); a.muts = s->o("muts"); a.mudf = s->o("mudf"); a.t = s->o("t_2");
  a.t_ave = s->o("t_2save"); a.ww_1 = s->o("ww1"); a.u = s->o("u_2"); a.u_1 = s->o("u_save"); a.v = s->o("v_2"); a.v_1 = s->o("v_save");
  a.mut = s->o("mut"); a.muu = s->o("muu"); a.muv = s->o("muv"); a.t_1 = s->o("t_save"); a.ft = s->o("t_tend"); a.mu_tend = s->o("mu_tend");
  return a;
}

void Solver::advance_mu_t(int rk_step, bool final_ave) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  MuTArgs a = mu_t_args(this);
  const Box mass = spec() ? spec_mass_box() : Box{q.its, std::min(q.ite, q.ide - 1), 1, 1, q.jts, std::min(q.jte, q.jde - 1)};
  static const bool staged = [] { const char *e = getenv("WRFGPU_MT"); return !(e && !strcmp(e, "column")); }();
  if (staged && !empty(mass) && mass.i1 - mass.i0 >= 4) {   // operands staged by the bulk-copy engine (k_advance_mu_t_st); WRFGPU_MT=column: the register-prefetch form
    const dim3 gr((unsigned)((mass.i1 - mass.i0 + MTS_NCOL) / MTS_NCOL), (unsigned)(mass.j1 - mass.j0 + 1));
    const int nthr = 32 * (MTS_CW + 1);
    const size_t ring_bytes = sizeof(float) * MTS_STAGES * MTS_ROWS * MTS_RW;
    if (!mt_smem_set) {   // per device, hence per Solver
      CUDA_CHECK(cudaFuncSetAttribute(k_advance_mu_t_st<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
      CUDA_CHECK(cudaFuncSetAttribute(k_advance_mu_t_st<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
      mt_smem_set = true;
    }
    if (coef_dirty) sync_coef();
    kpre(final_ave ? "k_advance_mu_t_st<true>" : "k_advance_mu_t_st<false>");
    if (final_ave) k_advance_mu_t_st<true><<<gr, nthr, ring_bytes, stream>>>(g, mass, dts_rk, a);
    else k_advance_mu_t_st<false><<<gr, nthr, ring_bytes, stream>>>(g, mass, dts_rk, a);
    kpost();
    launches++;
  } else if (final_ave) LAUNCH2(this, k_advance_mu_t<true>, mass, dts_rk, a);    // only together with advance_w_fused
  else LAUNCH2(this, k_advance_mu_t<false>, mass, dts_rk, a);
  if (spec()) spec_update_tmu(dts_rk);   // solve_em.F:1459-1490
}

static WArgs w_args(Solver *s) {
  WArgs a;
  a.w = s->o("w_2"); a.ph = s->o("ph_2"); a.t_2ave = s->o("t_2save"); a.rw_tend = s->o("rw_tend"); a.ww = s->o("ww"); a.w_save = s->o("w_save");
  a.u = s->o("u_2"); a.v = s->o("v_2"); a.mut = s->o("mut"); a.muave = s->o("muave"); a.muts = s->o("muts"); a.t_2 = s->o("t_2");
  a.t_1 = s->o("t_save"); a.ph_1 = s->o("ph_save"); a.phb = s->o("phb"); a.ph_tend = s->o("ph_tend"); a.c2a = s->o("c2a"); a.cqw = s->o("cqw");
  a.alt = s->o("alt"); a.a = s->o("a"); a.alpha = s->o("alpha"); a.gamma = s->o("gamma");
  return a;
}

void Solver::advance_w(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  WFuse fz{};
  const Box mass = spec() ? spec_mass_box() : Box{q.its, std::min(q.ite, q.ide - 1), 1, 1, q.jts, std::min(q.jte, q.jde - 1)};
  LAUNCH2(this, k_advance_w<false>, mass, dts_rk, w_args(this), fz);
}

// the columns advance_mu_t and advance_w update with specified / nested boundaries (module_small_step_em.F:1054-1063, :1273-1282):
// one row in from every side, whatever spec_zone is
Box Solver::spec_mass_box() const {
  const Geo &q = g;
  Box bx{q.its, std::min(q.ite, q.ide - 1), 1, 1, std::max(q.jts, q.jds + 1), std::min(q.jte, q.jde - 2)};
  if (!q.periodic_x) { bx.i0 = std::max(q.its, q.ids + 1); bx.i1 = std::min(q.ite, q.ide - 2); }
  return bx;
}

// advance_w + sumflux (ww part) + calc_p_rho(iteration) in one column kernel, then the exchange and
// boundary fills that follow calc_p_rho in solve_em.F:1659-1713
void Solver::advance_w_fused(int rk_step, int iteration) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  if (q.kde < 5) throw std::runtime_error("wrfgpu: the fused column kernel needs at least 5 levels");
  WFuse fz{iteration, ns, o("al"), o("p"), o("pm1"), o("ww_m"), o("ru_m"), o("rv_m"), o("mu_2"), o("ww1")};
  const WArgs wa = w_args(this);
  // HALO_EM_C2 (solve_em.F:1659) overlaps the interior of the column kernel that produces ph, al, p
  overlapped(Box{q.its, q.ite, 1, 1, q.jts, q.jte}, "HALO_EM_C2", [&](const Box &bx) { LAUNCH2(this, k_advance_w<true>, bx, dts_rk, wa, fz); });
  bc({{"ph_2", 'w'}, {"al", 'p'}, {"p", 'p'}, {"muts", 't'}, {"mu_2", 't'}, {"mudf", 't'}});
}

void Solver::sumflux(int rk_step, int iteration) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  LAUNCH3(this, k_sumflux, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), iteration, ns, o("u_2"), o("v_2"), o("ww"), o("u_save"),
          o("v_save"), o("ww1"), o("muu"), o("muv"), o("ru_m"), o("rv_m"), o("ww_m"));
  if (spec()) spec_update_phw(dts_rk);   // solve_em.F:1582-1620, between sumflux and calc_p_rho
}

// the part of sumflux the fused advance_uv leaves: ww_m (all levels) and the level-kde reset of ru_m, rv_m
__global__ void k_sumflux_w(const Geo g, Box b, int iteration, int nsmall, cfp ww, cfp ww_lin, fp ru_m, fp rv_m, fp ww_m) {
  TID3;
  const int mini = min(g.ide - 1, g.ite), minj = min(g.jde - 1, g.jte);
  const bool dw = (i <= mini && j <= minj);
  float cq = 0.f;
  if (iteration != 1) cq = A3(ww_m, i, k, j);
  if (dw) cq = cq + A3(ww, i, k, j);
  if (iteration == nsmall && dw) cq = cq / (float)nsmall + A3(ww_lin, i, k, j);
  A3(ww_m, i, k, j) = cq;
  if (k == g.kde && iteration == 1) { A3(ru_m, i, k, j) = 0.f; A3(rv_m, i, k, j) = 0.f; }
}

void Solver::sumflux_w(int rk_step, int iteration) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  LAUNCH3(this, k_sumflux_w, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), iteration, ns, o("ww"), o("ww1"), o("ru_m"), o("rv_m"), o("ww_m"));
}

void Solver::calc_p_rho(int iteration) {
  const Geo &q = g;
  LAUNCH3(this, k_calc_p_rho, (Box{q.its, std::min(q.ite, q.ide - 1), 1, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)}), iteration,
          o("al"), o("p"), o("ph_2"), o("alt"), o("t_2"), o("t_save"), o("c2a"), o("pm1"), o("mu_2"), o("muts"));
  halo("HALO_EM_C2");
  bc({{"ph_2", 'w'}, {"al", 'p'}, {"p", 'p'}, {"muts", 't'}, {"mu_2", 't'}, {"mudf", 't'}});
}

void Solver::small_step_finish(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  LAUNCH2(this, k_calc_mu_uv_1, (Box{q.its, q.ite, 1, 1, q.jts, q.jte}), o("muts"), o("muus"), o("muvs"));
  SsfArgs a;
  a.u_2 = o("u_2"); a.v_2 = o("v_2"); a.w_2 = o("w_2"); a.t_2 = o("t_2"); a.ph_2 = o("ph_2"); a.ww = o("ww"); a.mu_2 = o("mu_2");
  a.ww1 = o("ww1"); a.mut = o("mut"); a.muts = o("muts"); a.muu = o("muu"); a.muus = o("muus"); a.muv = o("muv"); a.muvs = o("muvs");
  a.u_save = o("u_save"); a.v_save = o("v_save"); a.w_save = o("w_save"); a.t_save = o("t_save"); a.ph_save = o("ph_save");
  a.mu_save = o("mu_save"); a.h_diabatic = o("h_diabatic");
  static const bool vec4 = [] { const char *e = getenv("WRFGPU_SSP"); return !(e && !strcmp(e, "scalar")); }();
  if (vec4) LAUNCH3(this, k_ssf_v4, (Box{(q.its - q.ims) / 4, (q.ite - q.ims) / 4, 1, q.kde, q.jts, q.jte}), rk_step == c.rk_ord ? 1 : 0, dts_rk, ns, a);
  else LAUNCH3(this, k_ssf, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), rk_step == c.rk_ord ? 1 : 0, dts_rk, ns, a);
  if (rk_step == c.rk_ord) bc({{"ru_m", 'u'}, {"rv_m", 'v'}, {"ww_m", 'w'}, {"mut", 't'}, {"muts", 't'}});
}

}  // namespace wg
