// scalar.cu -- scalar transport of solve_em.F:1850-2955: rk_update_scalar_pd
// (module_em.F:1803), rk_scalar_tend (:1096) with advect_scalar or, on the last RK
// substep under *_adv_opt = 1, the positive-definite advect_scalar_pd
// (module_advect_em.F:6069-7885), and rk_update_scalar (module_em.F:1587).
//
// advect_scalar_pd in the reference materialises eight tile-sized 3-D flux arrays
// ("memory intensive", :6085-6088).  Here nothing but a per-cell limiter factor
// `scale` is stored: kernel 1 evaluates a cell's six face fluxes (high-order minus
// upwind, and upwind) and writes scale = max(0, ph_low/(flux_out+eps)) where the
// outgoing correction exceeds the low-order mass (else 1); kernel 2 re-evaluates
// the face fluxes, multiplies each high-order correction by the scale of its DONOR
// cell (:7764-7773) and writes the tendency.  Flux values are pure functions of the
// inputs, so recomputation is bit-identical to the stored version.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include "weno.cuh"

namespace wg {

__global__ void k_advect_scalar(const Geo g, Box b, cfp q, cfp ru, cfp rv, cfp rom, fp tend);   // tend.cu

// advect_scalar_weno on the mass points (*_adv_opt = 3, last RK substep, module_em.F:1300-1312)
__global__ void k_advect_scalar_weno(const Geo g, Box b, cfp q, cfp ru, cfp rv, cfp rom, fp tend) {
  TID3;
  A3(tend, i, k, j) = advect_scalar_weno_pt(g, q, q, ru, rv, rom, i, k, j);
}

// statement functions of advect_scalar_pd (:6169-6186) -- note the pre-divided constants
__device__ __forceinline__ float pflux4(float qm2, float qm1, float q0, float qp1) {
  return (7.f / 12.f) * (q0 + qm1) - (1.f / 12.f) * (qp1 + qm2);
}
__device__ __forceinline__ float pflux3(float qm2, float qm1, float q0, float qp1, float ua, float sg) {
  return pflux4(qm2, qm1, q0, qp1) + sg * sign1(ua) * (1.f / 12.f) * ((qp1 - qm2) - 3.f * (q0 - qm1));
}
__device__ __forceinline__ float pflux6(float qm3, float qm2, float qm1, float q0, float qp1, float qp2) {
  return (37.f / 60.f) * (q0 + qm1) - (2.f / 15.f) * (qp1 + qm2) + (1.f / 60.f) * (qp2 + qm3);
}
__device__ __forceinline__ float pflux5(float qm3, float qm2, float qm1, float q0, float qp1, float qp2, float ua, float sg) {
  return pflux6(qm3, qm2, qm1, q0, qp1, qp2) - sg * sign1(ua) * (1.f / 60.f) * ((qp2 - qm3) - 5.f * (qp1 - qm2) + 10.f * (q0 - qm1));
}
__device__ __forceinline__ float flux_upwind(float qm1, float q0, float cr) {
  return 0.5f * fminf(1.0f, (cr + fabsf(cr))) * qm1 + 0.5f * fmaxf(-1.0f, (cr - fabsf(cr))) * q0;
}

struct PdIn {
  cfp q, q_old, ru, rv, rom, mut, mub, mu_old;
  float dt;
  cfp romI;    // wwI under IEVA (advect_scalar_mono's ieva_corr, module_advect_em.F:10383), else null
  bool weno;   // advect_scalar_wenopd (*_adv_opt = 4): WENO in place of the 5th-order fluxes, and on k = 4..ktf-2
};
struct Face { float hi, lo; bool ok; };   // hi = high-order minus low, lo = upwind flux

__device__ __forceinline__ int pd_order(int f, bool dlo, bool dhi, int lo_edge, int hi_edge) {
  // faces lo_edge+1 / hi_edge-1 are 2nd order, lo_edge+2 / hi_edge-2 3rd order
  if ((!dlo || f >= lo_edge + 3) && (!dhi || f <= hi_edge - 3)) return 5;
  if (f == lo_edge + 1) return 2;
  if (f == lo_edge + 2) return 3;
  if (f == hi_edge - 1) return 2;
  if (f == hi_edge - 2) return 3;
  return 0;
}

// x-face ii of row j, level k (:6640-6770)
__device__ __forceinline__ Face pd_face_x(const Geo &g, const PdIn &s, int ii, int k, int j, bool dxs, bool dxe) {
  Face f;
  const int ord = pd_order(ii, dxs, dxe, g.ids, g.ide);
  const float sgh = s.weno ? g.sgn_ts : g.sg_hs;   // h_sca_adv_order = 6: flux6 / flux4, i.e. no upwind term
  f.ok = ord != 0;
  if (!f.ok) { f.hi = 0.f; f.lo = 0.f; return f; }
  const float dx = 2.f / (A2(g.msfty, ii, j) + A2(g.msfty, ii - 1, j)) / g.rdx;
  const float mu = 0.5f * ((g.c1h[k] * A2(s.mut, ii, j) + g.c2h[k]) + (g.c1h[k] * A2(s.mut, ii - 1, j) + g.c2h[k]));
  float vel, cr;
  if (ord == 2 && ii == g.ids + 1) { vel = A3(s.ru, ii, k, j) / mu; cr = vel * s.dt / dx; }   // :6683-6684
  else { vel = A3(s.ru, ii, k, j); cr = vel * s.dt / dx / mu; }
  f.lo = mu * (dx / s.dt) * flux_upwind(A3(s.q_old, ii - 1, k, j), A3(s.q_old, ii, k, j), cr);
  float h;
  if (ord == 5 && s.weno) h = weno5(vel, vel * g.sgn_ts >= 0.0f, A3(s.q, ii - 3, k, j), A3(s.q, ii - 2, k, j), A3(s.q, ii - 1, k, j), A3(s.q, ii, k, j), A3(s.q, ii + 1, k, j), A3(s.q, ii + 2, k, j));
  else if (ord == 5) h = vel * pflux5(A3(s.q, ii - 3, k, j), A3(s.q, ii - 2, k, j), A3(s.q, ii - 1, k, j), A3(s.q, ii, k, j), A3(s.q, ii + 1, k, j), A3(s.q, ii + 2, k, j), vel, sgh);
  else if (ord == 3) h = vel * pflux3(A3(s.q, ii - 2, k, j), A3(s.q, ii - 1, k, j), A3(s.q, ii, k, j), A3(s.q, ii + 1, k, j), vel, sgh);
  else h = 0.5f * (A3(s.ru, ii, k, j)) * (A3(s.q, ii, k, j) + A3(s.q, ii - 1, k, j));
  f.hi = h - f.lo;
  return f;
}
// y-face jj of column i (:6540-6630)
__device__ __forceinline__ Face pd_face_y(const Geo &g, const PdIn &s, int i, int k, int jj, bool dys, bool dye) {
  Face f;
  const int ord = pd_order(jj, dys, dye, g.jds, g.jde);
  const float sgh = s.weno ? g.sgn_ts : g.sg_hs;
  f.ok = ord != 0;
  if (!f.ok) { f.hi = 0.f; f.lo = 0.f; return f; }
  const float dy = 2.f / (A2(g.msftx, i, jj) + A2(g.msftx, i, jj - 1)) / g.rdy;
  const float mu = 0.5f * ((g.c1h[k] * A2(s.mut, i, jj) + g.c2h[k]) + (g.c1h[k] * A2(s.mut, i, jj - 1) + g.c2h[k]));
  const float vel = A3(s.rv, i, k, jj);
  const float cr = vel * s.dt / dy / mu;
  f.lo = mu * (dy / s.dt) * flux_upwind(A3(s.q_old, i, k, jj - 1), A3(s.q_old, i, k, jj), cr);
  float h;
  if (ord == 5 && s.weno) h = weno5(vel, vel * g.sgn_ts >= 0.0f, A3(s.q, i, k, jj - 3), A3(s.q, i, k, jj - 2), A3(s.q, i, k, jj - 1), A3(s.q, i, k, jj), A3(s.q, i, k, jj + 1), A3(s.q, i, k, jj + 2));
  else if (ord == 5) h = vel * pflux5(A3(s.q, i, k, jj - 3), A3(s.q, i, k, jj - 2), A3(s.q, i, k, jj - 1), A3(s.q, i, k, jj), A3(s.q, i, k, jj + 1), A3(s.q, i, k, jj + 2), vel, sgh);
  else if (ord == 3) h = vel * pflux3(A3(s.q, i, k, jj - 2), A3(s.q, i, k, jj - 1), A3(s.q, i, k, jj), A3(s.q, i, k, jj + 1), vel, sgh);
  else h = 0.5f * A3(s.rv, i, k, jj) * (A3(s.q, i, k, jj) + A3(s.q, i, k, jj - 1));
  f.hi = h - f.lo;
  return f;
}
// z-face kk of column (i,j) (:7595-7645)
__device__ __forceinline__ Face pd_face_z(const Geo &g, const PdIn &s, int i, int kk, int j) {
  Face f;
  f.ok = true;
  if (kk == 1 || kk == g.kde) { f.hi = 0.f; f.lo = 0.f; return f; }
  const int ktf = g.kde - 1;
  const float sgv = s.weno ? g.sgn_ts : g.sg_vs;   // v_sca_adv_order = 4: flux4
  const float dz = 2.f / (g.rdnw[kk] + g.rdnw[kk - 1]);
  const float mu = 0.5f * ((g.c1h[kk] * A2(s.mut, i, j) + g.c2h[kk]) + (g.c1h[kk] * A2(s.mut, i, j) + g.c2h[kk]));
  const float vel = A3(s.rom, i, kk, j);
  const float cr = vel * s.dt / dz / mu;
  f.lo = mu * (dz / s.dt) * flux_upwind(A3(s.q_old, i, kk - 1, j), A3(s.q_old, i, kk, j), cr);
  float h;
  if (kk == 2 || kk == ktf) h = A3(s.rom, i, kk, j) * (g.fnm[kk] * A3(s.q, i, kk, j) + g.fnp[kk] * A3(s.q, i, kk - 1, j));
  else if (s.weno && kk >= 4 && kk <= ktf - 2)   // :9217-9262
    h = weno5(vel, -vel * g.sgn_ts >= 0.0f, A3(s.q, i, kk - 3, j), A3(s.q, i, kk - 2, j), A3(s.q, i, kk - 1, j), A3(s.q, i, kk, j), A3(s.q, i, kk + 1, j), A3(s.q, i, kk + 2, j));
  else h = vel * pflux3(A3(s.q, i, kk - 2, j), A3(s.q, i, kk - 1, j), A3(s.q, i, kk, j), A3(s.q, i, kk + 1, j), -vel, sgv);
  f.hi = h - f.lo;
  return f;
}

// kernel 1: limiter factor on the tile + 1 ring (:7681-7776)
__global__ void k_pd_scale(const Geo g, Box b, PdIn s, fp scale) {
  TID3;
  const bool dxs = !(g.periodic_x || g.sym_xs), dxe = !(g.periodic_x || g.sym_xe);
  const bool dys = !(g.periodic_y || g.sym_ys), dye = !(g.periodic_y || g.sym_ye);
  const Face x0 = pd_face_x(g, s, i, k, j, dxs, dxe), x1 = pd_face_x(g, s, i + 1, k, j, dxs, dxe);
  const Face y0 = pd_face_y(g, s, i, k, j, dys, dye), y1 = pd_face_y(g, s, i, k, j + 1, dys, dye);
  const Face z0 = pd_face_z(g, s, i, k, j), z1 = pd_face_z(g, s, i, k + 1, j);
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float ph_low = ((g.c1h[k] * A2(s.mub, i, j) + g.c2h[k]) + (g.c1h[k] * A2(s.mu_old, i, j))) * A3(s.q_old, i, k, j) -
                       s.dt * (msftx * msfty * (g.rdx * (x1.lo - x0.lo) + g.rdy * (y1.lo - y0.lo)) + msfty * g.rdnw[k] * (z1.lo - z0.lo));
  const float flux_out = s.dt * ((msftx * msfty) * (g.rdx * (fmaxf(0.f, x1.hi) - fminf(0.f, x0.hi)) + g.rdy * (fmaxf(0.f, y1.hi) - fminf(0.f, y0.hi))) +
                                 msfty * g.rdnw[k] * (fminf(0.f, z1.hi) - fmaxf(0.f, z0.hi)));
  float sc = 1.f;
  if (flux_out > ph_low) sc = fmaxf(0.f, ph_low / (flux_out + 1.e-20f));
  A3(scale, i, k, j) = sc;
}

// kernel 2: limited flux divergence (:7778-7883) + open-boundary terms (:7255-7318)
__global__ void k_pd_tend(const Geo g, Box b, PdIn s, cfp scale, Box lim, fp tend) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const bool dxs = !(g.periodic_x || g.sym_xs), dxe = !(g.periodic_x || g.sym_xe);
  const bool dys = !(g.periodic_y || g.sym_ys), dye = !(g.periodic_y || g.sym_ye);
  cfp q = s.q, q_old = s.q_old, ru = s.ru, rv = s.rv;
  // scale of a cell, 1 outside the limiter box
  auto sc_at = [&](int ii, int kk, int jj) -> float {
    if (ii < lim.i0 || ii > lim.i1 || jj < lim.j0 || jj > lim.j1 || kk < lim.k0 || kk > lim.k1) return 1.f;
    return A3(scale, ii, kk, jj);
  };
  float t = 0.f;
  // open boundary terms come first in the reference (they are added to `tendency` before the divergence)
  const int i_end = min(ite, ide - 1), j_end = min(jte, jde - 1);
  if (g.open_xs && i == ids) {
    const float ub = fminf(0.5f * (A3(ru, ids, k, j) + A3(ru, ids + 1, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, ids + 1, k, j) - A3(q_old, ids, k, j)) + A3(q, ids, k, j) * (A3(ru, ids + 1, k, j) - A3(ru, ids, k, j)));
  }
  if (g.open_xe && ite == ide && i == i_end) {
    const float ub = fmaxf(0.5f * (A3(ru, ide - 1, k, j) + A3(ru, ide, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, i_end, k, j) - A3(q_old, i_end - 1, k, j)) + A3(q, i_end, k, j) * (A3(ru, ide, k, j) - A3(ru, ide - 1, k, j)));
  }
  if (g.open_ys && j == jds) {
    const float vb = fminf(0.5f * (A3(rv, i, k, jds) + A3(rv, i, k, jds + 1)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, jds + 1) - A3(q_old, i, k, jds)) + A3(q, i, k, jds) * (A3(rv, i, k, jds + 1) - A3(rv, i, k, jds)));
  }
  if (g.open_ye && jte == jde && j == j_end) {
    const float vb = fmaxf(0.5f * (A3(rv, i, k, jde - 1) + A3(rv, i, k, jde)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, j_end) - A3(q_old, i, k, j_end - 1)) + A3(q, i, k, j_end) * (A3(rv, i, k, jde) - A3(rv, i, k, jde - 1)));
  }
  // limited high-order correction of a face: scaled by the donor cell (:7764-7773)
  // x: fqx(i+1)>0 -> scale(i) ; fqx(i)<0 -> scale(i).  For face ii the donor is ii-1 if hi>0, ii if hi<0.
  // z: sign convention is reversed (fqz(k+1)<0 and fqz(k)>0 are outgoing): donor is kk-1 if hi<0, kk if hi>0.
  {
    Face z0 = pd_face_z(g, s, i, k, j), z1 = pd_face_z(g, s, i, k + 1, j);
    float h0 = z0.hi, h1 = z1.hi;
    if (h0 > 0.f) h0 = sc_at(i, k, j) * h0; else if (h0 < 0.f) h0 = sc_at(i, k - 1, j) * h0;
    if (h1 > 0.f) h1 = sc_at(i, k + 1, j) * h1; else if (h1 < 0.f) h1 = sc_at(i, k, j) * h1;
    t = t - g.rdnw[k] * (h1 - h0 + z1.lo - z0.lo);
  }
  {
    const int i_lo = dxs ? max(its, ids + 1) : its, i_hi = dxe ? min(ite, ide - 2) : min(ite, ide - 1);
    if (i >= i_lo && i <= i_hi) {
      Face x0 = pd_face_x(g, s, i, k, j, dxs, dxe), x1 = pd_face_x(g, s, i + 1, k, j, dxs, dxe);
      float h0 = x0.hi, h1 = x1.hi;
      if (h0 > 0.f) h0 = sc_at(i - 1, k, j) * h0; else if (h0 < 0.f) h0 = sc_at(i, k, j) * h0;
      if (h1 > 0.f) h1 = sc_at(i, k, j) * h1; else if (h1 < 0.f) h1 = sc_at(i + 1, k, j) * h1;
      t = t - A2(g.msftx, i, j) * (g.rdx * (h1 - h0 + x1.lo - x0.lo));
    }
  }
  {
    const int j_lo = dys ? max(jts, jds + 1) : jts, j_hi = dye ? min(jte, jde - 2) : min(jte, jde - 1);
    if (j >= j_lo && j <= j_hi) {
      Face y0 = pd_face_y(g, s, i, k, j, dys, dye), y1 = pd_face_y(g, s, i, k, j + 1, dys, dye);
      float h0 = y0.hi, h1 = y1.hi;
      if (h0 > 0.f) h0 = sc_at(i, k, j - 1) * h0; else if (h0 < 0.f) h0 = sc_at(i, k, j) * h0;
      if (h1 > 0.f) h1 = sc_at(i, k, j) * h1; else if (h1 < 0.f) h1 = sc_at(i, k, j + 1) * h1;
      t = t - A2(g.msftx, i, j) * (g.rdy * (h1 - h0 + y1.lo - y0.lo));
    }
  }
  A3(tend, i, k, j) = t;
}


// ===========================================================================================
// advect_scalar_mono (module_advect_em.F:9495-10560): the monotonic limiter, *_adv_opt = 2.
// Same two-pass structure as the positive-definite path above: pass 1 writes the two per-cell factors
// scale_in / scale_out, pass 2 re-evaluates the face fluxes, scales each high-order part by
// min(scale_in(receiver), scale_out(donor)) (:10385-10416) and writes the tendency.
__device__ __forceinline__ float mflux4(float qm2, float qm1, float q0, float qp1) {
  return (7.f / 12.f) * (q0 + qm1) - (1.f / 12.f) * (qp1 + qm2);
}
__device__ __forceinline__ float mflux3(float qm2, float qm1, float q0, float qp1, float ua) {   // no sign(time_step) here (:9590-9592)
  return mflux4(qm2, qm1, q0, qp1) + sign1(ua) * (1.f / 12.f) * ((qp1 - qm2) - 3.f * (q0 - qm1));
}
__device__ __forceinline__ float mflux6(float qm3, float qm2, float qm1, float q0, float qp1, float qp2) {
  return (37.f / 60.f) * (q0 + qm1) - (2.f / 15.f) * (qp1 + qm2) + (1.f / 60.f) * (qp2 + qm3);
}
__device__ __forceinline__ float mflux5(float qm3, float qm2, float qm1, float q0, float qp1, float qp2, float ua) {
  return mflux6(qm3, qm2, qm1, q0, qp1, qp2) - sign1(ua) * (1.f / 60.f) * ((qp2 - qm3) - 5.f * (qp1 - qm2) + 10.f * (q0 - qm1));
}
__device__ __forceinline__ float mflux_upwind(float qm1, float q0, float cr) {   // :9602-9603
  return 0.5f * (1.f + sign1(cr)) * qm1 + 0.5f * (1.f - sign1(cr)) * q0;
}
struct MFace { float hi, lo, cr; bool ok; };   // hi = high-order minus upwind, lo = upwind flux, cr: cr > 0 means flow from the lower-index cell

__device__ __forceinline__ MFace mono_face_x(const Geo &g, const PdIn &s, int ii, int k, int j, bool dxs, bool dxe) {
  MFace f;
  const int ord = pd_order(ii, dxs, dxe, g.ids, g.ide);
  f.ok = ord != 0;
  if (!f.ok) { f.hi = 0.f; f.lo = 0.f; f.cr = 0.f; return f; }
  const float vel = A3(s.ru, ii, k, j);
  f.cr = vel;
  f.lo = vel * mflux_upwind(A3(s.q_old, ii - 1, k, j), A3(s.q_old, ii, k, j), f.cr);
  float h;
  if (ord == 5) h = vel * mflux5(A3(s.q, ii - 3, k, j), A3(s.q, ii - 2, k, j), A3(s.q, ii - 1, k, j), A3(s.q, ii, k, j), A3(s.q, ii + 1, k, j), A3(s.q, ii + 2, k, j), vel);
  else if (ord == 3) h = vel * mflux3(A3(s.q, ii - 2, k, j), A3(s.q, ii - 1, k, j), A3(s.q, ii, k, j), A3(s.q, ii + 1, k, j), vel);
  else h = 0.5f * (A3(s.ru, ii, k, j)) * (A3(s.q, ii, k, j) + A3(s.q, ii - 1, k, j));
  f.hi = h - f.lo;
  return f;
}
__device__ __forceinline__ MFace mono_face_y(const Geo &g, const PdIn &s, int i, int k, int jj, bool dys, bool dye) {
  MFace f;
  const int ord = pd_order(jj, dys, dye, g.jds, g.jde);
  f.ok = ord != 0;
  if (!f.ok) { f.hi = 0.f; f.lo = 0.f; f.cr = 0.f; return f; }
  const float vel = A3(s.rv, i, k, jj);
  f.cr = vel;
  f.lo = vel * mflux_upwind(A3(s.q_old, i, k, jj - 1), A3(s.q_old, i, k, jj), f.cr);
  float h;
  if (ord == 5) h = vel * mflux5(A3(s.q, i, k, jj - 3), A3(s.q, i, k, jj - 2), A3(s.q, i, k, jj - 1), A3(s.q, i, k, jj), A3(s.q, i, k, jj + 1), A3(s.q, i, k, jj + 2), vel);
  else if (ord == 3) h = vel * mflux3(A3(s.q, i, k, jj - 2), A3(s.q, i, k, jj - 1), A3(s.q, i, k, jj), A3(s.q, i, k, jj + 1), vel);
  else h = 0.5f * A3(s.rv, i, k, jj) * (A3(s.q, i, k, jj) + A3(s.q, i, k, jj - 1));
  f.hi = h - f.lo;
  return f;
}
__device__ __forceinline__ MFace mono_face_z(const Geo &g, const PdIn &s, int i, int kk, int j) {
  MFace f;
  const int ktf = g.kde - 1;
  f.ok = kk >= 2 && kk <= ktf;          // faces 1 and kde carry no flux and no bound (:10096-10101)
  if (!f.ok) { f.hi = 0.f; f.lo = 0.f; f.cr = 0.f; return f; }
  const float vel = A3(s.rom, i, kk, j);
  f.cr = -vel;
  f.lo = vel * mflux_upwind(A3(s.q_old, i, kk - 1, j), A3(s.q_old, i, kk, j), f.cr);
  float h;
  if (kk == 2 || kk == ktf) h = A3(s.rom, i, kk, j) * (g.fnm[kk] * A3(s.q, i, kk, j) + g.fnp[kk] * A3(s.q, i, kk - 1, j));
  else h = vel * mflux3(A3(s.q, i, kk - 2, j), A3(s.q, i, kk - 1, j), A3(s.q, i, kk, j), A3(s.q, i, kk + 1, j), -vel);
  f.hi = h - f.lo;
  return f;
}

// pass 1: scale_in, scale_out of a cell of the limiter box (:10322-10383)
__global__ void k_mono_scale(const Geo g, Box b, PdIn s, fp scale_in, fp scale_out) {
  TID3;
  const bool dxs = !(g.periodic_x || g.sym_xs), dxe = !(g.periodic_x || g.sym_xe);
  const bool dys = !(g.periodic_y || g.sym_ys), dye = !(g.periodic_y || g.sym_ye);
  const MFace x0 = mono_face_x(g, s, i, k, j, dxs, dxe), x1 = mono_face_x(g, s, i + 1, k, j, dxs, dxe);
  const MFace y0 = mono_face_y(g, s, i, k, j, dys, dye), y1 = mono_face_y(g, s, i, k, j + 1, dys, dye);
  const MFace z0 = mono_face_z(g, s, i, k, j), z1 = mono_face_z(g, s, i, k + 1, j);
  // bounds: the cell's own old value and the old value of every neighbour that is upwind of a computed face
  const float qo = A3(s.q_old, i, k, j);
  float qmax = qo, qmin = qo;
  auto take = [&](float v) { qmax = fmaxf(qmax, v); qmin = fminf(qmin, v); };
  if (x0.ok && x0.cr > 0.f) take(A3(s.q_old, i - 1, k, j));
  if (x1.ok && !(x1.cr > 0.f)) take(A3(s.q_old, i + 1, k, j));
  if (y0.ok && y0.cr > 0.f) take(A3(s.q_old, i, k, j - 1));
  if (y1.ok && !(y1.cr > 0.f)) take(A3(s.q_old, i, k, j + 1));
  if (z0.ok && z0.cr > 0.f) take(A3(s.q_old, i, k - 1, j));
  if (z1.ok && !(z1.cr > 0.f)) take(A3(s.q_old, i, k + 1, j));
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  // romI = wwI, identically 0 without implicit vertical advection
  const float dromI = s.romI ? (A3(s.romI, i, k + 1, j) - A3(s.romI, i, k, j)) : (0.f - 0.f);
  const float ieva_corr = (c1h * A2(s.mut, i, j) + c2h) + s.dt * msfty * g.rdnw[k] * dromI;
  const float ph_upwind = ((c1h * A2(s.mub, i, j) + c2h) + (c1h * A2(s.mu_old, i, j))) * qo -
                          s.dt * (msftx * msfty * (g.rdx * (x1.lo - x0.lo) + g.rdy * (y1.lo - y0.lo)) + msfty * g.rdnw[k] * (z1.lo - z0.lo));
  const float flux_in = -s.dt * ((msftx * msfty) * (g.rdx * (fminf(0.f, x1.hi) - fmaxf(0.f, x0.hi)) + g.rdy * (fminf(0.f, y1.hi) - fmaxf(0.f, y0.hi))) +
                                 msfty * g.rdnw[k] * (fmaxf(0.f, z1.hi) - fminf(0.f, z0.hi)));
  const float ph_hi = ieva_corr * qmax - ph_upwind;
  float si = 1.f, so = 1.f;
  if (flux_in > ph_hi) si = fmaxf(0.f, ph_hi / (flux_in + 1.e-20f));
  const float flux_out = s.dt * ((msftx * msfty) * (g.rdx * (fmaxf(0.f, x1.hi) - fminf(0.f, x0.hi)) + g.rdy * (fmaxf(0.f, y1.hi) - fminf(0.f, y0.hi))) +
                                 msfty * g.rdnw[k] * (fminf(0.f, z1.hi) - fmaxf(0.f, z0.hi)));
  const float ph_low = ph_upwind - ieva_corr * qmin;
  if (flux_out > ph_low) so = fmaxf(0.f, ph_low / (flux_out + 1.e-20f));
  A3(scale_in, i, k, j) = si;
  A3(scale_out, i, k, j) = so;
}

// pass 2: limited flux divergence (:10385-10556) + open-boundary terms (:10019-10083)
__global__ void k_mono_tend(const Geo g, Box b, PdIn s, cfp scale_in, cfp scale_out, Box lim, fp tend) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const bool dxs = !(g.periodic_x || g.sym_xs), dxe = !(g.periodic_x || g.sym_xe);
  const bool dys = !(g.periodic_y || g.sym_ys), dye = !(g.periodic_y || g.sym_ye);
  cfp q = s.q, q_old = s.q_old, ru = s.ru, rv = s.rv;
  // factors of a cell, 1 outside the limiter box (the reference initialises both arrays to 1)
  auto in_lim = [&](int ii, int kk, int jj) { return !(ii < lim.i0 || ii > lim.i1 || jj < lim.j0 || jj > lim.j1 || kk < lim.k0 || kk > lim.k1); };
  auto sin_at = [&](int ii, int kk, int jj) -> float { return in_lim(ii, kk, jj) ? A3(scale_in, ii, kk, jj) : 1.f; };
  auto sout_at = [&](int ii, int kk, int jj) -> float { return in_lim(ii, kk, jj) ? A3(scale_out, ii, kk, jj) : 1.f; };
  float t = 0.f;
  const int i_end = min(ite, ide - 1), j_end = min(jte, jde - 1);
  if (g.open_xs && i == ids) {
    const float ub = fminf(0.5f * (A3(ru, ids, k, j) + A3(ru, ids + 1, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, ids + 1, k, j) - A3(q_old, ids, k, j)) + A3(q, ids, k, j) * (A3(ru, ids + 1, k, j) - A3(ru, ids, k, j)));
  }
  if (g.open_xe && ite == ide && i == i_end) {
    const float ub = fmaxf(0.5f * (A3(ru, ide - 1, k, j) + A3(ru, ide, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, i_end, k, j) - A3(q_old, i_end - 1, k, j)) + A3(q, i_end, k, j) * (A3(ru, ide, k, j) - A3(ru, ide - 1, k, j)));
  }
  if (g.open_ys && j == jds) {
    const float vb = fminf(0.5f * (A3(rv, i, k, jds) + A3(rv, i, k, jds + 1)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, jds + 1) - A3(q_old, i, k, jds)) + A3(q, i, k, jds) * (A3(rv, i, k, jds + 1) - A3(rv, i, k, jds)));
  }
  if (g.open_ye && jte == jde && j == j_end) {
    const float vb = fmaxf(0.5f * (A3(rv, i, k, jde - 1) + A3(rv, i, k, jde)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, j_end) - A3(q_old, i, k, j_end - 1)) + A3(q, i, k, j_end) * (A3(rv, i, k, jde) - A3(rv, i, k, jde - 1)));
  }
  // the reference scales the faces in its own loops (x: i_start..i_end+1, y: j_start..j_end+1, z: k = 2..ktf of the
  // limiter box); a face outside them keeps its value, which is what the factor 1 of sin_at / sout_at gives
  {
    const MFace z0 = mono_face_z(g, s, i, k, j), z1 = mono_face_z(g, s, i, k + 1, j);
    float h0 = z0.hi, h1 = z1.hi;
    if (z0.ok) h0 = (h0 < 0.f ? fminf(sin_at(i, k, j), sout_at(i, k - 1, j)) : fminf(sout_at(i, k, j), sin_at(i, k - 1, j))) * h0;
    if (z1.ok) h1 = (h1 < 0.f ? fminf(sin_at(i, k + 1, j), sout_at(i, k, j)) : fminf(sout_at(i, k + 1, j), sin_at(i, k, j))) * h1;
    t = t - g.rdnw[k] * (h1 - h0 + z1.lo - z0.lo);
  }
  {
    const int i_lo = dxs ? max(its, ids + 1) : its, i_hi = dxe ? min(ite, ide - 2) : min(ite, ide - 1);
    if (i >= i_lo && i <= i_hi) {
      const MFace x0 = mono_face_x(g, s, i, k, j, dxs, dxe), x1 = mono_face_x(g, s, i + 1, k, j, dxs, dxe);
      float h0 = x0.hi, h1 = x1.hi;
      h0 = (h0 > 0.f ? fminf(sin_at(i, k, j), sout_at(i - 1, k, j)) : fminf(sout_at(i, k, j), sin_at(i - 1, k, j))) * h0;
      h1 = (h1 > 0.f ? fminf(sin_at(i + 1, k, j), sout_at(i, k, j)) : fminf(sout_at(i + 1, k, j), sin_at(i, k, j))) * h1;
      t = t - A2(g.msftx, i, j) * (g.rdx * (h1 - h0 + x1.lo - x0.lo));
    }
  }
  {
    const int j_lo = dys ? max(jts, jds + 1) : jts, j_hi = dye ? min(jte, jde - 2) : min(jte, jde - 1);
    if (j >= j_lo && j <= j_hi) {
      const MFace y0 = mono_face_y(g, s, i, k, j, dys, dye), y1 = mono_face_y(g, s, i, k, j + 1, dys, dye);
      float h0 = y0.hi, h1 = y1.hi;
      h0 = (h0 > 0.f ? fminf(sin_at(i, k, j), sout_at(i, k, j - 1)) : fminf(sout_at(i, k, j), sin_at(i, k, j - 1))) * h0;
      h1 = (h1 > 0.f ? fminf(sin_at(i, k, j + 1), sout_at(i, k, j)) : fminf(sout_at(i, k, j + 1), sin_at(i, k, j))) * h1;
      t = t - A2(g.msftx, i, j) * (g.rdy * (h1 - h0 + y1.lo - y0.lo));
    }
  }
  A3(tend, i, k, j) = t;
}

// rk_update_scalar_pd (module_em.F:1803-1916)
__global__ void k_update_scalar_pd(const Geo g, Box b, float dt, fp scalar, fp sc_tend, cfp mu_old, cfp mu_new, cfp mu_base) {
  TID3;
  const float tend_in = A3(sc_tend, i, k, j);
  const float s_in = A3(scalar, i, k, j);
  float tendency = 0.f;
  tendency = tendency + tend_in;
  const float muold = A2(mu_old, i, j) + A2(mu_base, i, j), munew = A2(mu_new, i, j) + A2(mu_base, i, j);
  A3(scalar, i, k, j) = ((g.c1h[k] * muold + g.c2h[k]) * s_in + dt * tendency) / (g.c1h[k] * munew + g.c2h[k]);
  A3(sc_tend, i, k, j) = after_load(0.f, tend_in);   // see common.cuh: keep the store behind the load
}

// rk_update_scalar (module_em.F:1587-1799)
__global__ void k_update_scalar(const Geo g, Box b, int rk_step, float dt, fp scalar_1, fp scalar_2, cfp sc_tend, cfp advect_tend,
                                cfp mu_old, cfp mu_new, cfp mu_base) {
  TID3;
  // the advective tendency is not used in the specified rows (:1671-1678); sc_tend (there: the boundary file's) is
  const int sz = g.spec_zone;
  const bool adv = !g.spec || ((g.periodic_x || (i >= g.ids + sz && i <= g.ide - sz - 1)) && j >= g.jds + sz && j <= g.jde - sz - 1);
  const float adv_t = A3(advect_tend, i, k, j) * A2(g.msfty, i, j);   // loaded unconditionally: a guarded load cost 25 % of this streaming kernel
  float tendency = adv ? adv_t : 0.f;
  tendency = tendency + A3(sc_tend, i, k, j);
  const float muold = A2(mu_old, i, j) + A2(mu_base, i, j), munew = A2(mu_new, i, j) + A2(mu_base, i, j);
  float s1;
  if (rk_step == 1) { s1 = A3(scalar_2, i, k, j); A3(scalar_1, i, k, j) = s1; }
  else s1 = A3(scalar_1, i, k, j);
  A3(scalar_2, i, k, j) = ((g.c1h[k] * muold + g.c2h[k]) * s1 + dt * tendency) / (g.c1h[k] * munew + g.c2h[k]);
}

// the two updates with 16-byte accesses (mass columns: lanes lo..him of a group of four)
__global__ void k_update_scalar_pd_v4(const Geo g, Box b, float dt, fp scalar, fp sc_tend, cfp mu_old, cfp mu_new, cfp mu_base) {
  TID3_V4;
  (void)hi;
  if (him < lo) return;
  float tend_in[4], s_in[4], mo[4], mn[4], mb[4], out[4], z[4];
  ld4(&A3(sc_tend, i, k, j), tend_in); ld4(&A3(scalar, i, k, j), s_in);
  ld4(&A2(mu_old, i, j), mo); ld4(&A2(mu_new, i, j), mn); ld4(&A2(mu_base, i, j), mb);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float tendency = 0.f;
    tendency = tendency + tend_in[q];
    const float muold = mo[q] + mb[q], munew = mn[q] + mb[q];
    out[q] = ((g.c1h[k] * muold + g.c2h[k]) * s_in[q] + dt * tendency) / (g.c1h[k] * munew + g.c2h[k]);
    z[q] = after_load(0.f, tend_in[q]);
  }
  st4(&A3(scalar, i, k, j), out, lo, him);
  st4(&A3(sc_tend, i, k, j), z, lo, him);
}
__global__ void k_update_scalar_v4(const Geo g, Box b, int rk_step, float dt, fp scalar_1, fp scalar_2, cfp sc_tend, cfp advect_tend,
                                   cfp mu_old, cfp mu_new, cfp mu_base) {
  TID3_V4;
  (void)hi;
  if (him < lo) return;
  const int sz = g.spec_zone;
  const bool adv_j = !g.spec || (j >= g.jds + sz && j <= g.jde - sz - 1);
  float at[4], my[4], st[4], mo[4], mn[4], mb[4], s1[4], out[4];
  ld4(&A3(advect_tend, i, k, j), at); ld4(&A2(g.msfty, i, j), my); ld4(&A3(sc_tend, i, k, j), st);
  ld4(&A2(mu_old, i, j), mo); ld4(&A2(mu_new, i, j), mn); ld4(&A2(mu_base, i, j), mb);
  if (rk_step == 1) { ld4(&A3(scalar_2, i, k, j), s1); st4(&A3(scalar_1, i, k, j), s1, lo, him); }
  else ld4(&A3(scalar_1, i, k, j), s1);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const bool adv = adv_j && (!g.spec || g.periodic_x || (i + q >= g.ids + sz && i + q <= g.ide - sz - 1));
    const float adv_t = at[q] * my[q];
    float tendency = adv ? adv_t : 0.f;
    tendency = tendency + st[q];
    const float muold = mo[q] + mb[q], munew = mn[q] + mb[q];
    out[q] = ((g.c1h[k] * muold + g.c2h[k]) * s1[q] + dt * tendency) / (g.c1h[k] * munew + g.c2h[k]);
  }
  st4(&A3(scalar_2, i, k, j), out, lo, him);
}

// bound_tke (module_em.F:2616-2646), after the update of the tke species (solve_em.F:2437-2445)
__global__ void k_bound_tke(const Geo g, Box b, float tke_upper_bound, fp tke) {
  TID3;
  A3(tke, i, k, j) = fminf(tke_upper_bound, fmaxf(A3(tke, i, k, j), 0.f));
}

void Solver::rk_scalar_step(int rk_step) {
  const std::string tke_name = tke_index() ? "scalar_" + std::to_string(tke_index()) : std::string("-");
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  const int im = std::min(q.ite, q.ide - 1), jm = std::min(q.jte, q.jde - 1), ktf = q.kde - 1;
  Box mass{q.its, im, 1, ktf, q.jts, jm};
  Box mass4 = box_v4(q, 1, ktf);   // the same points as groups of four columns (the *_v4 kernels mask the lanes)
  mass4.j1 = jm;
  static const bool vec4 = [] { const char *e = getenv("WRFGPU_SSP"); return !(e && !strcmp(e, "scalar")); }();
  const bool last = rk_step == c.rk_ord;
  if (last) {
    std::vector<BcItem> items;
    for (auto &s : species)
      if (s.adv_opt == 1 || s.adv_opt == 2 || s.adv_opt == 4) {   // /= ORIGINAL and /= WENO_SCALAR (solve_em.F:1850-1977)
        if (vec4) LAUNCH3(this, k_update_scalar_pd_v4, mass4, dt_rk, o(s.o), o(s.t), o("mu_1"), o("mu_1"), o("mub"));
        else LAUNCH3(this, k_update_scalar_pd, mass, dt_rk, o(s.o), o(s.t), o("mu_1"), o("mu_1"), o("mub"));
        items.push_back({s.o, 'p'});
      }
    if (!items.empty()) { halo("HALO_EM_MOIST_OLD_E_7"); bc(items); }
  }
  halo("HALO_EM_D");
  // IEVA on the last substep (module_em.F:1216-1243): the explicit schemes see wwE, the implicit solve takes wwI
  const bool ieva = c.zadvect_implicit > 0 && last;
  if (ieva && !species.empty()) {   // see HALO_EM_IEVA in halo.cu: current u_1, v_1 around the tile for WW_SPLIT
    halo("HALO_EM_IEVA");
    bc({{"u_1", 'U'}, {"v_1", 'V'}});
  }
  if (c.zadvect_implicit > 0 && !species.empty()) ieva_split(false, o("u_1"), o("v_1"), o("ww_m"), o("muts"), ieva);
  cfp rom = ieva ? o("wwE") : o("ww_m");
  const bool dflt_s = c.h_sca_adv_order == 5 && c.v_sca_adv_order == 3;   // the k-marching kernels exist for these
  const bool dxs = !(c.periodic_x || c.symmetric_xs), dxe = !(c.periodic_x || c.symmetric_xe);
  const bool dys = !(c.periodic_y || c.symmetric_ys), dye = !(c.periodic_y || c.symmetric_ye);
  // what follows rk_scalar_tend for one species (solve_em.F:2261-2355, :2437-2456, :2808-2843)
  auto update = [&](const Species &s) {
    const bool moist = s.f.rfind("moist_", 0) == 0, is_tke = s.f == tke_name;
    if (spec() && rk_step == 1 && has_bdy(s.f)) bdy_scalar(s);   // relax_bdy_scalar + spec_bdy_scalar: qv, or every species with boundary arrays
    if (vec4) LAUNCH3(this, k_update_scalar_v4, mass4, rk_step, dt_rk, o(s.o), o(s.f), o(s.t), o("advect_tend"), o("mu_1"), o("mu_2"), o("mub"));
    else LAUNCH3(this, k_update_scalar, mass, rk_step, dt_rk, o(s.o), o(s.f), o(s.t), o("advect_tend"), o("mu_1"), o("mu_2"), o("mub"));
    if (is_tke) LAUNCH3(this, k_bound_tke, mass, c.tke_upper_bound, o(s.f));
    // zero gradient on outflow / zero on inflow: moist species without boundary arrays (specified only), tke
    if ((moist && c.specified && !c.have_bcs_moist && s.f != "moist_qv") || (is_tke && spec())) flow_dep_bdy(s.f);
  };
  for (auto &s : species) {
    if (last && (s.adv_opt == 1 || s.adv_opt == 2 || s.adv_opt == 4)) {
      PdIn in;
      in.q = o(s.f); in.q_old = o(s.o); in.ru = o("ru_m"); in.rv = o("rv_m"); in.rom = rom; in.romI = ieva ? o("wwI") : nullptr;
      in.mut = o("muts"); in.mub = o("mub"); in.mu_old = o("mu_1"); in.dt = dt_rk; in.weno = s.adv_opt == 4;
      // limiter box (:7683-7710)
      Box lim{q.its - 1, im + 1, 1, ktf, q.jts - 1, jm + 1};
      if (dxs) lim.i0 = std::max(q.its - 1, q.ids);
      if (dxe) lim.i1 = std::min(q.ite + 1, q.ide - 1);
      if (dys) lim.j0 = std::max(q.jts - 1, q.jds);
      if (dye) lim.j1 = std::min(q.jte + 1, q.jde - 1);
      if (spec()) {   // :7697-7702
        if (dxs) lim.i0 = std::max(q.its - 1, q.ids + 1);
        if (dxe) lim.i1 = std::min(q.ite + 1, q.ide - 2);
        if (dys) lim.j0 = std::max(q.jts - 1, q.jds + 1);
        if (dye) lim.j1 = std::min(q.jte + 1, q.jde - 2);
      }
      if (c.open_xs && dxs) lim.i0 = std::max(q.its - 1, q.ids + 1);
      if (c.open_xe && dxe) lim.i1 = std::min(q.ite + 1, q.ide - 2);
      if (c.open_ys && dys) lim.j0 = std::max(q.jts - 1, q.jds + 1);
      if (c.open_ye && dye) lim.j1 = std::min(q.jte + 1, q.jde - 2);
      // "a" is free between calc_coef_w uses only until the next small_step_prep: use advect scratch "alpha"?  No:
      // the limiter factor gets its own scratch so that a, alpha, gamma stay comparable with the oracle.
      if (s.adv_opt == 2) {   // monotonic limiter: gather kernels on the whole limiter box / mass points
        LAUNCH3(this, k_mono_scale, lim, in, o("pd_scale"), o("pd_scale2"));
        LAUNCH3(this, k_mono_tend, mass, in, o("pd_scale"), o("pd_scale2"), lim, o("advect_tend"));
        if (ieva) ieva_implicit(0, o("advect_tend"), o(s.o), o("mut_old"), o("muts"), dt_rk);
        update(s);
        continue;
      }
      if (s.adv_opt == 4 || ieva || !dflt_s) {   // WENO inside the positive-definite limiter, or IEVA's wwE: the gather kernels everywhere
        LAUNCH3(this, k_pd_scale, lim, in, o("pd_scale"));
        LAUNCH3(this, k_pd_tend, mass, in, o("pd_scale"), lim, o("advect_tend"));
        if (ieva) ieva_implicit(0, o("advect_tend"), o(s.o), o("mut_old"), o("muts"), dt_rk);
        update(s);
        continue;
      }
      // k-marching kernels (pd_fast.cu) where every face is 5th order, the gather kernels on the strips around
      Box sfast = lim, tfast = interior_box(q);
      tfast.k1 = ktf;
      if (!c.periodic_x) { sfast.i0 = std::max(lim.i0, q.ids + 4); sfast.i1 = std::min(lim.i1, q.ide - 5); }
      if (!c.periodic_y) { sfast.j0 = std::max(lim.j0, q.jds + 4); sfast.j1 = std::min(lim.j1, q.jde - 5); }
      Box strips[4];
      launch_pd_fast(this, 1, sfast, in.q, in.q_old, dt_rk, o("pd_scale"));
      int nst = edge_strips(lim, sfast, strips);
      for (int m = 0; m < nst; m++) LAUNCH3(this, k_pd_scale, strips[m], in, o("pd_scale"));
      launch_pd_fast(this, 2, tfast, in.q, in.q_old, dt_rk, o("advect_tend"));
      nst = edge_strips(mass, tfast, strips);
      for (int m = 0; m < nst; m++) LAUNCH3(this, k_pd_tend, strips[m], in, o("pd_scale"), lim, o("advect_tend"));
    } else if (last && s.adv_opt == 3) {   // WENO_SCALAR (module_em.F:1300-1312)
      LAUNCH3(this, k_advect_scalar_weno, mass, o(s.f), o("ru_m"), o("rv_m"), rom, o("advect_tend"));
    } else if (!dflt_s) {   // orders other than (5, 3): the gather kernel everywhere
      LAUNCH3(this, k_advect_scalar, mass, o(s.f), o("ru_m"), o("rv_m"), rom, o("advect_tend"));
    } else {
      Box in = interior_box(q);
      in.k1 = ktf;
      launch_advect_scalar_fast(this, false, in, o(s.f), o("ru_m"), o("rv_m"), rom, nullptr, nullptr, nullptr, o("advect_tend"));
      Box strips[4];
      const int nst = edge_strips(mass, in, strips);
      for (int m = 0; m < nst; m++) LAUNCH3(this, k_advect_scalar, strips[m], o(s.f), o("ru_m"), o("rv_m"), rom, o("advect_tend"));
    }
    if (ieva) ieva_implicit(0, o("advect_tend"), o(s.o), o("mut_old"), o("muts"), dt_rk);   // module_em.F:1346-1364: mut_new = mut (= muts)
    if (rk_step == 1 && c.diff_opt == 1) {   // rk_step_1: diff_opt1 (module_em.F:1378-1419): K = xkhh, kvdq = kvdif/prandtl
      const float prandtl = 1.f / 3.0f;
      launch_diffuse(this, s.f == "moist_qv" ? 4 : 5, o(s.f), o(s.t), c.khdif / prandtl, c.kvdif / prandtl, c.km_opt == 1 ? nullptr : o("xkhh"));
    }
    if (rk_step == 1 && c.diff_6th_opt != 0 && !(s.f.rfind("moist_", 0) == 0 ? c.moist_mix6_off : c.scalar_mix6_off))
      launch_diff6(this, 5, o(s.f), o(s.t), o("muts"), dt_rk);   // module_em.F:1421-1437 (dt_step = dt_rk)
    update(s);
  }
}

}  // namespace wg
