// advect.cpp -- oracle restatement of dyn_em/module_advect_em.F for
// h_*_adv_order = 5, v_*_adv_order = 3 (the only orders every shipped ideal
// namelist uses): advect_u :126, advect_v :1530, advect_scalar :3029,
// advect_w :4364, advect_scalar_pd :6069.  The loop structure (degrade flags from
// tile bounds, rolling y-flux rows, separate x sweep, open-BC terms, vertical)
// follows the reference.  Test infrastructure only; see oracle.h header.
#include "oracle.h"
namespace orc {

namespace {
// 5th-order WENO face flux (Jiang and Shu 1996) of advect_scalar_weno / advect_scalar_wenopd
// (module_advect_em.F:8095-8124, 8803-8835).  m3..p2 are the six values around the face (q(-3) .. q(+2)),
// `upw` = the wind blows from the low-index side.  The candidate fluxes and smoothness indicators are REAL
// expressions assigned to DOUBLE PRECISION variables; the weights are double precision (pw = 2; eps = 1.d-40 for scalars,
// 1.d-18 in advect_weno_u/v/w :10950).
inline float weno5(float vel, bool upw, float m3, float m2, float m1, float p0, float p1, float p2, double eps = 1.0e-40) {
  float qim2, qim1, qi, qip1, qip2;
  if (upw) { qip2 = p1; qip1 = p0; qi = m1; qim1 = m2; qim2 = m3; }
  else { qip2 = m2; qip1 = m1; qi = p0; qim1 = p1; qim2 = p2; }
  const double f0 = (double)(1.f / 3.f * qim2 - 7.f / 6.f * qim1 + 11.f / 6.f * qi);
  const double f1 = (double)(-1.f / 6.f * qim1 + 5.f / 6.f * qi + 1.f / 3.f * qip1);
  const double f2 = (double)(1.f / 3.f * qi + 5.f / 6.f * qip1 - 1.f / 6.f * qip2);
  const float a0 = qim2 - 2.f * qim1 + qi, b0 = qim2 - 4.f * qim1 + 3.f * qi;
  const float a1 = qim1 - 2.f * qi + qip1, b1 = qim1 - qip1;
  const float a2 = qi - 2.f * qip1 + qip2, b2 = qip2 - 4.f * qip1 + 3.f * qi;
  const double beta0 = (double)(13.f / 12.f * (a0 * a0) + 1.f / 4.f * (b0 * b0));
  const double beta1 = (double)(13.f / 12.f * (a1 * a1) + 1.f / 4.f * (b1 * b1));
  const double beta2 = (double)(13.f / 12.f * (a2 * a2) + 1.f / 4.f * (b2 * b2));
  const double gi0 = 1.0 / 10.0, gi1 = 6.0 / 10.0, gi2 = 3.0 / 10.0;
  const double wi0 = gi0 / ((eps + beta0) * (eps + beta0));
  const double wi1 = gi1 / ((eps + beta1) * (eps + beta1));
  const double wi2 = gi2 / ((eps + beta2) * (eps + beta2));
  const double sumwk = wi0 + wi1 + wi2;
  return (float)((double)vel * (wi0 * f0 + wi1 * f1 + wi2 * f2) / sumwk);
}
struct Flux {
  float sgn_ts;  // sign(1,time_step)
  bool weno = false;
  // h_*_adv_order = 6 / v_*_adv_order = 4: "EXACTLY the same [code] EXCEPT the 5th order and 3rd operators are used in
  // place of the 6th and 4th order operators" (module_advect_em.F:488-491, 829-830, and the vert_order 4 / 3 blocks)
  bool h_even = false, v_even = false;
  // horizontal 3rd / 4th-order face value next to a boundary, vertical 3rd / 4th-order face value
  inline float h3(float q_im2, float q_im1, float q_i, float q_ip1, float ua) const {
    return (h_even && !weno) ? flux4(q_im2, q_im1, q_i, q_ip1) : flux3(q_im2, q_im1, q_i, q_ip1, ua);
  }
  inline float v3(float q_im2, float q_im1, float q_i, float q_ip1, float ua) const {
    return (v_even && !weno) ? flux4(q_im2, q_im1, q_i, q_ip1) : flux3(q_im2, q_im1, q_i, q_ip1, ua);
  }   // advect_weno_u/v/w (:10876-13041, momentum_adv_opt = 3): WENO in place of flux5, vert_order 5 structure
  // vel * flux5(...) of a horizontal face, or the WENO face flux
  inline float h5(float vel, float m3, float m2, float m1, float p0, float p1, float p2) const {
    if (weno) return weno5(vel, vel * sgn_ts >= 0.0f, m3, m2, m1, p0, p1, p2, 1.0e-18);
    return vel * flux5(m3, m2, m1, p0, p1, p2, vel);   // flux6 when h_even
  }
  // vertical face k: flux3 under vert_order 3; with weno the 5th-order structure, WENO on k = kts+3 .. k5max
  inline float v35(int k, int k5min, int k5max, float vel, float m3, float m2, float m1, float p0, float p1, float p2) const {
    if (weno && k >= k5min && k <= k5max) return weno5(vel, -vel * sgn_ts >= 0.0f, m3, m2, m1, p0, p1, p2, 1.0e-18);
    return vel * v3(m2, m1, p0, p1, -vel);
  }
  // module_advect_em.F:202-213 statement functions
  inline float flux4(float q_im2, float q_im1, float q_i, float q_ip1) const {
    return (7.f * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0f;
  }
  inline float flux3(float q_im2, float q_im1, float q_i, float q_ip1, float ua) const {
    return flux4(q_im2, q_im1, q_i, q_ip1) + sgn_ts * sign1(ua) * ((q_ip1 - q_im2) - 3.f * (q_i - q_im1)) / 12.0f;
  }
  inline float flux6(float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2) const {
    return (37.f * (q_i + q_im1) - 8.f * (q_ip1 + q_im2) + (q_ip2 + q_im3)) / 60.0f;
  }
  inline float flux5(float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2, float ua) const {
    if (h_even) return flux6(q_im3, q_im2, q_im1, q_i, q_ip1, q_ip2);
    return flux6(q_im3, q_im2, q_im1, q_i, q_ip1, q_ip2) -
           sgn_ts * sign1(ua) * ((q_ip2 - q_im3) - 5.f * (q_ip1 - q_im2) + 10.f * (q_i - q_im1)) / 60.0f;
  }
};
struct Row {  // (its-1:ite+2, kts:kte) style scratch
  std::vector<float> d; int i0, ni;
  Row(int its, int ite, int kte) : d((size_t)(ite - its + 6) * (kte + 2)), i0(its - 2), ni(ite - its + 6) {}
  inline float &operator()(int i, int k) { return d[(i - i0) + (size_t)k * ni]; }
};
inline void degrade_flags(const Cx &x, int xe_off, int ye_off, bool &dxs, bool &dxe, bool &dys, bool &dye) {
  dxs = dxe = dys = dye = true;
  if (x.c.periodic_x || x.c.symmetric_xs || (x.its > x.ids + 3)) dxs = false;
  if (x.c.periodic_x || x.c.symmetric_xe || (x.ite < x.ide - xe_off)) dxe = false;
  if (x.c.periodic_y || x.c.symmetric_ys || (x.jts > x.jds + 3)) dys = false;
  if (x.c.periodic_y || x.c.symmetric_ye || (x.jte < x.jde - ye_off)) dye = false;
}
}  // namespace

// ---------------------------------------------------------------------------
// module_advect_em.F:126-1526, horz_order 5 (:486-745), open BC (:1240-1334),
// vert_order 3 (:1470-1498)
void advect_u(const Cx &x, A3 u, A3 u_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, x.kde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy;
  Flux F{(float)(x.time_step >= 0 ? 1 : -1), weno, x.c.h_mom_adv_order == 6, x.c.v_mom_adv_order == 4};
  bool degrade_xs, degrade_xe, degrade_ys, degrade_ye;
  degrade_flags(x, 2, 4, degrade_xs, degrade_xe, degrade_ys, degrade_ye);
  Row fq0(its, ite, kte), fq1(its, ite, kte), fqx(its, ite, kte), vflux(its, ite, kte);

  int i_start = its, i_end = ite;
  if (x.oxs()) i_start = std::max(ids + 1, its);
  if (x.oxe()) i_end = std::min(ide - 1, ite);
  if (x.c.periodic_x) { i_start = its; i_end = ite; }
  int j_start = jts, j_end = std::min(jte, jde - 1);
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_ys) { j_start = std::max(jts, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte, jde - 2); j_end_f = jde - 3; }
  Row *jp1 = &fq1, *jp0 = &fq0;
  for (int j = j_start; j <= j_end + 1; j++) {
    Row &fqy = *jp1;
    if (j >= j_start_f && j <= j_end_f) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i - 1, k, j));
          fqy(i, k) = F.h5(vel, u(i, k, j - 3), u(i, k, j - 2), u(i, k, j - 1), u(i, k, j), u(i, k, j + 1), u(i, k, j + 2));
        }
    } else if (j == jds + 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.25f * (rv(i, k, j) + rv(i - 1, k, j)) * (u(i, k, j) + u(i, k, j - 1));
    } else if (j == jds + 2) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i - 1, k, j));
          fqy(i, k) = vel * F.h3(u(i, k, j - 2), u(i, k, j - 1), u(i, k, j), u(i, k, j + 1), vel);
        }
    } else if (j == jde - 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.25f * (rv(i, k, j) + rv(i - 1, k, j)) * (u(i, k, j) + u(i, k, j - 1));
    } else if (j == jde - 2) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i - 1, k, j));
          fqy(i, k) = vel * F.h3(u(i, k, j - 2), u(i, k, j - 1), u(i, k, j), u(i, k, j + 1), vel);
        }
    }
    if (j > j_start)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float mrdy = x.msfux(i, j - 1) * rdy;
          tendency(i, k, j - 1) = tendency(i, k, j - 1) - mrdy * ((*jp1)(i, k) - (*jp0)(i, k));
        }
    std::swap(jp1, jp0);
  }

  // x flux divergence
  i_start = its; i_end = ite; j_start = jts; j_end = std::min(jte, jde - 1);
  int i_start_f = i_start, i_end_f = i_end + 1;
  if (degrade_xs) { i_start = std::max(ids + 1, its); i_start_f = ids + 3; }
  if (degrade_xe) { i_end = std::min(ide - 1, ite); i_end_f = ide - 2; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) {
        float vel = 0.5f * (ru(i, k, j) + ru(i - 1, k, j));
        fqx(i, k) = F.h5(vel, u(i - 3, k, j), u(i - 2, k, j), u(i - 1, k, j), u(i, k, j), u(i + 1, k, j), u(i + 2, k, j));
      }
    if (degrade_xs) {
      if (i_start == ids + 1) {
        int i = ids + 1;
        for (int k = kts; k <= ktf; k++) {
          float ub = u(i - 1, k, j);
          if (x.spec && u(i, k, j) < 0.f) ub = u(i, k, j);   // specified uses upstream normal wind at boundaries (:698)
          fqx(i, k) = 0.25f * (ru(i, k, j) + ru(i - 1, k, j)) * (u(i, k, j) + ub);
        }
      }
      int i = ids + 2;
      for (int k = kts; k <= ktf; k++) {
        float vel = 0.5f * (ru(i, k, j) + ru(i - 1, k, j));
        fqx(i, k) = vel * F.h3(u(i - 2, k, j), u(i - 1, k, j), u(i, k, j), u(i + 1, k, j), vel);
      }
    }
    if (degrade_xe) {
      if (i_end == ide - 1) {
        int i = ide;
        for (int k = kts; k <= ktf; k++) {
          float ub = u(i, k, j);
          if (x.spec && u(i - 1, k, j) > 0.f) ub = u(i - 1, k, j);   // :720
          fqx(i, k) = 0.25f * (ru(i, k, j) + ru(i - 1, k, j)) * (u(i - 1, k, j) + ub);
        }
      }
      for (int k = kts; k <= ktf; k++) {
        int i = ide - 1;
        float vel = 0.5f * (ru(i, k, j) + ru(i - 1, k, j));
        fqx(i, k) = vel * F.h3(u(i - 2, k, j), u(i - 1, k, j), u(i, k, j), u(i + 1, k, j), vel);
      }
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mrdx = x.msfux(i, j) * rdx;
        tendency(i, k, j) = tendency(i, k, j) - mrdx * (fqx(i + 1, k) - fqx(i, k));
      }
  }

  // radiative lateral boundary condition in x for normal velocity (:1244-1268)
  if (x.c.open_xs && its == ids) {
    for (int j = jts; j <= std::min(jte, jde - 1); j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::min(ru(its, k, j) - cb * (c1(k) * mut(its, j) + c2(k)), 0.f);
        tendency(its, k, j) = tendency(its, k, j) - rdx * ub * (u_old(its + 1, k, j) - u_old(its, k, j));
      }
  }
  if (x.c.open_xe && ite == ide) {
    for (int j = jts; j <= std::min(jte, jde - 1); j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::max(ru(ite, k, j) + cb * (c1(k) * mut(ite - 1, j) + c2(k)), 0.f);
        tendency(ite, k, j) = tendency(ite, k, j) - rdx * ub * (u_old(ite, k, j) - u_old(ite - 1, k, j));
      }
  }
  i_start = its; i_end = std::min(ite, ide);
  int imin = ids, imax = ide - 1;
  if (x.c.open_xs) { i_start = std::max(ids + 1, its); imin = ids; }
  if (x.c.open_xe) { i_end = std::min(ite, ide - 1); imax = ide - 1; }
  if (x.c.open_ys && jts == jds) {
    for (int i = i_start; i <= i_end; i++) {
      float mrdy = x.msfux(i, jts) * rdy;
      int ip = std::min(imax, i), im = std::max(imin, i - 1);
      for (int k = kts; k <= ktf; k++) {
        float vw = 0.5f * (rv(ip, k, jts) + rv(im, k, jts));
        float vb = std::min(vw, 0.f);
        float dvm = rv(ip, k, jts + 1) - rv(ip, k, jts);
        float dvp = rv(im, k, jts + 1) - rv(im, k, jts);
        tendency(i, k, jts) = tendency(i, k, jts) - mrdy * (vb * (u_old(i, k, jts + 1) - u_old(i, k, jts)) + 0.5f * u(i, k, jts) * (dvm + dvp));
      }
    }
  }
  if (x.c.open_ye && jte == jde) {
    for (int i = i_start; i <= i_end; i++) {
      float mrdy = x.msfux(i, jte - 1) * rdy;
      int ip = std::min(imax, i), im = std::max(imin, i - 1);
      for (int k = kts; k <= ktf; k++) {
        float vw = 0.5f * (rv(ip, k, jte) + rv(im, k, jte));
        float vb = std::max(vw, 0.f);
        float dvm = rv(ip, k, jte) - rv(ip, k, jte - 1);
        float dvp = rv(im, k, jte) - rv(im, k, jte - 1);
        tendency(i, k, jte - 1) = tendency(i, k, jte - 1) -
                                  mrdy * (vb * (u_old(i, k, jte - 1) - u_old(i, k, jte - 2)) + 0.5f * u(i, k, jte - 1) * (dvm + dvp));
      }
    }
  }

  // vertical advection (:1336-1498); note the reference tests open_ys/open_ye here
  i_start = its; i_end = ite; j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.oys()) i_start = std::max(ids + 1, its);
  if (x.oye()) i_end = std::min(ide - 1, ite);
  if (x.c.periodic_x) { i_start = its; i_end = ite; }
  for (int i = i_start; i <= i_end; i++) { vflux(i, kts) = 0.f; vflux(i, kte) = 0.f; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vel = 0.5f * (rom(i - 1, k, j) + rom(i, k, j));
        vflux(i, k) = F.v35(k, kts + 3, ktf - 2, vel, u(i, k - 3, j), u(i, k - 2, j), u(i, k - 1, j), u(i, k, j), u(i, k + 1, j), u(i, k + 2, j));
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      vflux(i, k) = 0.5f * (rom(i, k, j) + rom(i - 1, k, j)) * (fzm(k) * u(i, k, j) + fzp(k) * u(i, k - 1, j));
      k = ktf;
      vflux(i, k) = 0.5f * (rom(i, k, j) + rom(i - 1, k, j)) * (fzm(k) * u(i, k, j) + fzp(k) * u(i, k - 1, j));
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tendency(i, k, j) - rdzw(k) * (vflux(i, k + 1) - vflux(i, k));
  }
}

// ---------------------------------------------------------------------------
// module_advect_em.F:1530-3024, horz_order 5 (:1904-2171), open BC (:2690-2800),
// vert_order 3 (:2961-2995)
void advect_v(const Cx &x, A3 v, A3 v_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, x.kde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy;
  Flux F{(float)(x.time_step >= 0 ? 1 : -1), weno, x.c.h_mom_adv_order == 6, x.c.v_mom_adv_order == 4};
  bool degrade_xs, degrade_xe, degrade_ys, degrade_ye;
  degrade_flags(x, 3, 3, degrade_xs, degrade_xe, degrade_ys, degrade_ye);
  Row fq0(its, ite, kte), fq1(its, ite, kte), fqx(its, ite, kte), vflux(its, ite, kte);

  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = jte;
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_ys) { j_start = std::max(jts, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte, jde - 1); j_end_f = jde - 2; }
  Row *jp1 = &fq1, *jp0 = &fq0;
  for (int j = j_start; j <= j_end + 1; j++) {
    Row &fqy = *jp1;
    if (j >= j_start_f && j <= j_end_f) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i, k, j - 1));
          fqy(i, k) = F.h5(vel, v(i, k, j - 3), v(i, k, j - 2), v(i, k, j - 1), v(i, k, j), v(i, k, j + 1), v(i, k, j + 2));
        }
    } else if (j == jds + 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vb = v(i, k, j - 1);
          if (x.spec && v(i, k, j) < 0.f) vb = v(i, k, j);   // specified uses upstream normal wind at boundaries (:1985)
          fqy(i, k) = 0.25f * (rv(i, k, j) + rv(i, k, j - 1)) * (v(i, k, j) + vb);
        }
    } else if (j == jds + 2) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i, k, j - 1));
          fqy(i, k) = vel * F.h3(v(i, k, j - 2), v(i, k, j - 1), v(i, k, j), v(i, k, j + 1), vel);
        }
    } else if (j == jde) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vb = v(i, k, j);
          if (x.spec && v(i, k, j - 1) > 0.f) vb = v(i, k, j - 1);   // :2007
          fqy(i, k) = 0.25f * (rv(i, k, j) + rv(i, k, j - 1)) * (vb + v(i, k, j - 1));
        }
    } else if (j == jde - 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = 0.5f * (rv(i, k, j) + rv(i, k, j - 1));
          fqy(i, k) = vel * F.h3(v(i, k, j - 2), v(i, k, j - 1), v(i, k, j), v(i, k, j + 1), vel);
        }
    }
    if (j > j_start)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float mrdy = x.msfvy(i, j - 1) * rdy;
          tendency(i, k, j - 1) = tendency(i, k, j - 1) - mrdy * ((*jp1)(i, k) - (*jp0)(i, k));
        }
    std::swap(jp1, jp0);
  }

  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = jte;
  if (x.oys()) j_start = std::max(jds + 1, jts);
  if (x.oye()) j_end = std::min(jde - 1, jte);
  int i_start_f = i_start, i_end_f = i_end + 1;
  if (degrade_xs) { i_start = std::max(ids + 1, its); i_start_f = std::min(i_start + 2, ids + 3); }
  if (degrade_xe) { i_end = std::min(ide - 2, ite); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) {
        float vel = 0.5f * (ru(i, k, j) + ru(i, k, j - 1));
        fqx(i, k) = F.h5(vel, v(i - 3, k, j), v(i - 2, k, j), v(i - 1, k, j), v(i, k, j), v(i + 1, k, j), v(i + 2, k, j));
      }
    if (degrade_xs) {
      for (int i = i_start; i <= i_start_f - 1; i++) {
        if (i == ids + 1)
          for (int k = kts; k <= ktf; k++) fqx(i, k) = 0.25f * (ru(i, k, j) + ru(i, k, j - 1)) * (v(i, k, j) + v(i - 1, k, j));
        if (i == ids + 2)
          for (int k = kts; k <= ktf; k++) {
            float vel = 0.5f * (ru(i, k, j) + ru(i, k, j - 1));
            fqx(i, k) = vel * F.h3(v(i - 2, k, j), v(i - 1, k, j), v(i, k, j), v(i + 1, k, j), vel);
          }
      }
    }
    if (degrade_xe) {
      for (int i = i_end_f + 1; i <= i_end + 1; i++) {
        if (i == ide - 1)
          for (int k = kts; k <= ktf; k++)
            fqx(i, k) = 0.25f * (ru(i_end + 1, k, j) + ru(i_end + 1, k, j - 1)) * (v(i_end + 1, k, j) + v(i_end, k, j));
        if (i == ide - 2)
          for (int k = kts; k <= ktf; k++) {
            float vel = 0.5f * (ru(i, k, j) + ru(i, k, j - 1));
            fqx(i, k) = vel * F.h3(v(i - 2, k, j), v(i - 1, k, j), v(i, k, j), v(i + 1, k, j), vel);
          }
      }
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mrdx = x.msfvy(i, j) * rdx;
        tendency(i, k, j) = tendency(i, k, j) - mrdx * (fqx(i + 1, k) - fqx(i, k));
      }
  }

  // radiative lateral boundary condition in y for normal velocity (:2712-2740)
  if (x.c.open_ys && jts == jds) {
    for (int i = its; i <= std::min(ite, ide - 1); i++)
      for (int k = kts; k <= ktf; k++) {
        float vb = std::min(rv(i, k, jts) - cb * (c1(k) * mut(i, jts) + c2(k)), 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) - rdy * vb * (v_old(i, k, jts + 1) - v_old(i, k, jts));
      }
  }
  if (x.c.open_ye && jte == jde) {
    for (int i = its; i <= std::min(ite, ide - 1); i++)
      for (int k = kts; k <= ktf; k++) {
        float vb = std::max(rv(i, k, jte) + cb * (c1(k) * mut(i, jte - 1) + c2(k)), 0.f);
        tendency(i, k, jte) = tendency(i, k, jte) - rdy * vb * (v_old(i, k, jte) - v_old(i, k, jte - 1));
      }
  }
  j_start = jts; j_end = std::min(jte, jde);
  int jmin = jds, jmax = jde - 1;
  if (x.c.open_ys) { j_start = std::max(jds + 1, jts); jmin = jds; }
  if (x.c.open_ye) { j_end = std::min(jte, jde - 1); jmax = jde - 1; }
  if (x.c.open_xs && its == ids) {
    for (int j = j_start; j <= j_end; j++) {
      float mrdx = x.msfvy(its, j) * rdx;
      int jp = std::min(jmax, j), jm = std::max(jmin, j - 1);
      for (int k = kts; k <= ktf; k++) {
        float uw = 0.5f * (ru(its, k, jp) + ru(its, k, jm));
        float ub = std::min(uw, 0.f);
        float dup = ru(its + 1, k, jp) - ru(its, k, jp);
        float dum = ru(its + 1, k, jm) - ru(its, k, jm);
        tendency(its, k, j) = tendency(its, k, j) - mrdx * (ub * (v_old(its + 1, k, j) - v_old(its, k, j)) + 0.5f * v(its, k, j) * (dup + dum));
      }
    }
  }
  if (x.c.open_xe && ite == ide) {
    for (int j = j_start; j <= j_end; j++) {
      float mrdx = x.msfvy(ite - 1, j) * rdx;
      int jp = std::min(jmax, j), jm = std::max(jmin, j - 1);
      for (int k = kts; k <= ktf; k++) {
        float uw = 0.5f * (ru(ite, k, jp) + ru(ite, k, jm));
        float ub = std::max(uw, 0.f);
        float dup = ru(ite, k, jp) - ru(ite - 1, k, jp);
        float dum = ru(ite, k, jm) - ru(ite - 1, k, jm);
        tendency(ite - 1, k, j) = tendency(ite - 1, k, j) -
                                  mrdx * (ub * (v_old(ite - 1, k, j) - v_old(ite - 2, k, j)) + 0.5f * v(ite - 1, k, j) * (dup + dum));
      }
    }
  }

  // vertical (:2802-2995)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = jte;
  for (int i = i_start; i <= i_end; i++) { vflux(i, kts) = 0.f; vflux(i, kte) = 0.f; }
  if (x.oys()) j_start = std::max(jds + 1, jts);
  if (x.oye()) j_end = std::min(jde - 1, jte);
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vel = 0.5f * (rom(i, k, j) + rom(i, k, j - 1));
        vflux(i, k) = F.v35(k, kts + 3, ktf - 2, vel, v(i, k - 3, j), v(i, k - 2, j), v(i, k - 1, j), v(i, k, j), v(i, k + 1, j), v(i, k + 2, j));
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      vflux(i, k) = 0.5f * (rom(i, k, j) + rom(i, k, j - 1)) * (fzm(k) * v(i, k, j) + fzp(k) * v(i, k - 1, j));
      k = ktf;
      vflux(i, k) = 0.5f * (rom(i, k, j) + rom(i, k, j - 1)) * (fzm(k) * v(i, k, j) + fzp(k) * v(i, k - 1, j));
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - (x.msfvy(i, j) / x.msfvx(i, j)) * rdzw(k) * (vflux(i, k + 1) - vflux(i, k));
  }
}

// ---------------------------------------------------------------------------
// module_advect_em.F:3029-4359, horz_order 5 (:3394-3650), open BC (:4104-4177),
// vert_order 3 (:4306-4333)
void advect_scalar(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut) {
  (void)mut;
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, x.kde - 1);
  const A1 &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy;
  Flux F{(float)(x.time_step >= 0 ? 1 : -1), false, x.c.h_sca_adv_order == 6, x.c.v_sca_adv_order == 4};
  bool degrade_xs, degrade_xe, degrade_ys, degrade_ye;
  degrade_flags(x, 3, 4, degrade_xs, degrade_xe, degrade_ys, degrade_ye);
  Row fq0(its, ite, kte), fq1(its, ite, kte), fqx(its, ite, kte), vflux(its, ite, kte);

  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_ys) { j_start = std::max(jts, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte, jde - 2); j_end_f = jde - 3; }
  Row *jp1 = &fq1, *jp0 = &fq0;
  for (int j = j_start; j <= j_end + 1; j++) {
    Row &fqy = *jp1;
    if (j >= j_start_f && j <= j_end_f) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = rv(i, k, j);
          fqy(i, k) = vel * F.flux5(field(i, k, j - 3), field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1),
                                    field(i, k, j + 2), vel);
        }
    } else if (j == jds + 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.5f * rv(i, k, j) * (field(i, k, j) + field(i, k, j - 1));
    } else if (j == jds + 2) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = rv(i, k, j);
          fqy(i, k) = vel * F.h3(field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), vel);
        }
    } else if (j == jde - 1) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.5f * rv(i, k, j) * (field(i, k, j) + field(i, k, j - 1));
    } else if (j == jde - 2) {
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = rv(i, k, j);
          fqy(i, k) = vel * F.h3(field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), vel);
        }
    }
    if (j > j_start)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float mrdy = x.msftx(i, j - 1) * rdy;
          tendency(i, k, j - 1) = tendency(i, k, j - 1) - mrdy * ((*jp1)(i, k) - (*jp0)(i, k));
        }
    std::swap(jp1, jp0);
  }

  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  int i_start_f = i_start, i_end_f = i_end + 1;
  if (degrade_xs) { i_start = std::max(ids + 1, its); i_start_f = std::min(i_start + 2, ids + 3); }
  if (degrade_xe) { i_end = std::min(ide - 2, ite); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) {
        float vel = ru(i, k, j);
        fqx(i, k) = vel * F.flux5(field(i - 3, k, j), field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j),
                                  field(i + 2, k, j), vel);
      }
    if (degrade_xs)
      for (int i = i_start; i <= i_start_f - 1; i++) {
        if (i == ids + 1)
          for (int k = kts; k <= ktf; k++) fqx(i, k) = 0.5f * (ru(i, k, j)) * (field(i, k, j) + field(i - 1, k, j));
        if (i == ids + 2)
          for (int k = kts; k <= ktf; k++) {
            float vel = ru(i, k, j);
            fqx(i, k) = vel * F.h3(field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), vel);
          }
      }
    if (degrade_xe)
      for (int i = i_end_f + 1; i <= i_end + 1; i++) {
        if (i == ide - 1)
          for (int k = kts; k <= ktf; k++) fqx(i, k) = 0.5f * (ru(i, k, j)) * (field(i, k, j) + field(i - 1, k, j));
        if (i == ide - 2)
          for (int k = kts; k <= ktf; k++) {
            float vel = ru(i, k, j);
            fqx(i, k) = vel * F.h3(field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), vel);
          }
      }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mrdx = x.msftx(i, j) * rdx;
        tendency(i, k, j) = tendency(i, k, j) - mrdx * (fqx(i + 1, k) - fqx(i, k));
      }
  }

  // open boundary terms
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::min(0.5f * (ru(its, k, j) + ru(its + 1, k, j)), 0.f);
        tendency(its, k, j) = tendency(its, k, j) -
                              rdx * (ub * (field_old(its + 1, k, j) - field_old(its, k, j)) + field(its, k, j) * (ru(its + 1, k, j) - ru(its, k, j)));
      }
  if (x.c.open_xe && ite == ide)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::max(0.5f * (ru(ite - 1, k, j) + ru(ite, k, j)), 0.f);
        tendency(i_end, k, j) = tendency(i_end, k, j) -
                                rdx * (ub * (field_old(i_end, k, j) - field_old(i_end - 1, k, j)) + field(i_end, k, j) * (ru(ite, k, j) - ru(ite - 1, k, j)));
      }
  if (x.c.open_ys && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::min(0.5f * (rv(i, k, jts) + rv(i, k, jts + 1)), 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) -
                              rdy * (vb * (field_old(i, k, jts + 1) - field_old(i, k, jts)) + field(i, k, jts) * (rv(i, k, jts + 1) - rv(i, k, jts)));
      }
  if (x.c.open_ye && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::max(0.5f * (rv(i, k, jte - 1) + rv(i, k, jte)), 0.f);
        tendency(i, k, j_end) = tendency(i, k, j_end) -
                                rdy * (vb * (field_old(i, k, j_end) - field_old(i, k, j_end - 1)) + field(i, k, j_end) * (rv(i, k, jte) - rv(i, k, jte - 1)));
      }

  // vertical
  for (int i = i_start; i <= i_end; i++) { vflux(i, kts) = 0.f; vflux(i, kte) = 0.f; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vel = rom(i, k, j);
        vflux(i, k) = vel * F.v3(field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), -vel);
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      vflux(i, k) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
      k = ktf;
      vflux(i, k) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tendency(i, k, j) - rdzw(k) * (vflux(i, k + 1) - vflux(i, k));
  }
}

// ---------------------------------------------------------------------------
// module_advect_em.F:4364-6065, horz_order 5 (:4784-5096), open BC (:5689-5820),
// vert_order 3 (:5996-6031).  rdzu = rdn.
void advect_w(const Cx &x, A3 w, A3 w_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno) {
  (void)mut;
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, x.kde - 1);
  const A1 &fzm = x.fnm, &fzp = x.fnp, &rdzu = x.rdn;
  const float rdx = x.rdx, rdy = x.rdy;
  // sic: w is advected with the SCALAR orders -- advect_w (:4463-4464) and advect_weno_w (:12339-12340) read h_sca_adv_order / v_sca_adv_order
  Flux F{(float)(x.time_step >= 0 ? 1 : -1), weno, x.c.h_sca_adv_order == 6, x.c.v_sca_adv_order == 4};
  bool degrade_xs, degrade_xe, degrade_ys, degrade_ye;
  degrade_flags(x, 3, 4, degrade_xs, degrade_xe, degrade_ys, degrade_ye);
  Row fq0(its, ite, kte), fq1(its, ite, kte), fqx(its, ite, kte), vflux(its, ite, kte);
  auto vely = [&](int i, int k, int j) {
    return (k <= ktf) ? fzm(k) * rv(i, k, j) + fzp(k) * rv(i, k - 1, j) : (2.f - fzm(k - 1)) * rv(i, k - 1, j) - fzp(k - 1) * rv(i, k - 2, j);
  };
  auto velx = [&](int i, int k, int j) {
    return (k <= ktf) ? fzm(k) * ru(i, k, j) + fzp(k) * ru(i, k - 1, j) : (2.f - fzm(k - 1)) * ru(i, k - 1, j) - fzp(k - 1) * ru(i, k - 2, j);
  };

  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_ys) { j_start = std::max(jts, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte, jde - 2); j_end_f = jde - 3; }
  Row *jp1 = &fq1, *jp0 = &fq0;
  for (int j = j_start; j <= j_end + 1; j++) {
    Row &fqy = *jp1;
    if (j >= j_start_f && j <= j_end_f) {
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = vely(i, k, j);
          fqy(i, k) = F.h5(vel, w(i, k, j - 3), w(i, k, j - 2), w(i, k, j - 1), w(i, k, j), w(i, k, j + 1), w(i, k, j + 2));
        }
    } else if (j == jds + 1) {
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.5f * (vely(i, k, j)) * (w(i, k, j) + w(i, k, j - 1));
    } else if (j == jds + 2) {
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = vely(i, k, j);
          fqy(i, k) = vel * F.h3(w(i, k, j - 2), w(i, k, j - 1), w(i, k, j), w(i, k, j + 1), vel);
        }
    } else if (j == jde - 1) {
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) fqy(i, k) = 0.5f * (vely(i, k, j)) * (w(i, k, j) + w(i, k, j - 1));
    } else if (j == jde - 2) {
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) {
          float vel = vely(i, k, j);
          fqy(i, k) = vel * F.h3(w(i, k, j - 2), w(i, k, j - 1), w(i, k, j), w(i, k, j + 1), vel);
        }
    }
    if (j > j_start)
      for (int k = kts + 1; k <= ktf + 1; k++)
        for (int i = i_start; i <= i_end; i++) {
          float mrdy = x.msftx(i, j - 1) * rdy;
          tendency(i, k, j - 1) = tendency(i, k, j - 1) - mrdy * ((*jp1)(i, k) - (*jp0)(i, k));
        }
    std::swap(jp1, jp0);
  }

  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  int i_start_f = i_start, i_end_f = i_end + 1;
  if (degrade_xs) { i_start = std::max(ids + 1, its); i_start_f = std::min(i_start + 2, ids + 3); }
  if (degrade_xe) { i_end = std::min(ide - 2, ite); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 1; k <= ktf + 1; k++)
      for (int i = i_start_f; i <= i_end_f; i++) {
        float vel = velx(i, k, j);
        fqx(i, k) = F.h5(vel, w(i - 3, k, j), w(i - 2, k, j), w(i - 1, k, j), w(i, k, j), w(i + 1, k, j), w(i + 2, k, j));
      }
    if (degrade_xs)
      for (int i = i_start; i <= i_start_f - 1; i++) {
        if (i == ids + 1)
          for (int k = kts + 1; k <= ktf + 1; k++) fqx(i, k) = 0.5f * (velx(i, k, j)) * (w(i, k, j) + w(i - 1, k, j));
        if (i == ids + 2)
          for (int k = kts + 1; k <= ktf + 1; k++) {
            float vel = velx(i, k, j);
            fqx(i, k) = vel * F.h3(w(i - 2, k, j), w(i - 1, k, j), w(i, k, j), w(i + 1, k, j), vel);
          }
      }
    if (degrade_xe)
      for (int i = i_end_f + 1; i <= i_end + 1; i++) {
        if (i == ide - 1)
          for (int k = kts + 1; k <= ktf + 1; k++) fqx(i, k) = 0.5f * (velx(i, k, j)) * (w(i, k, j) + w(i - 1, k, j));
        if (i == ide - 2)
          for (int k = kts + 1; k <= ktf + 1; k++) {
            float vel = velx(i, k, j);
            fqx(i, k) = vel * F.h3(w(i - 2, k, j), w(i - 1, k, j), w(i, k, j), w(i + 1, k, j), vel);
          }
      }
    for (int k = kts + 1; k <= ktf + 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mrdx = x.msftx(i, j) * rdx;
        tendency(i, k, j) = tendency(i, k, j) - mrdx * (fqx(i + 1, k) - fqx(i, k));
      }
  }

  // open boundary terms (:5689-5820)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++) {
        float uw = 0.5f * (fzm(k) * (ru(its, k, j) + ru(its + 1, k, j)) + fzp(k) * (ru(its, k - 1, j) + ru(its + 1, k - 1, j)));
        float ub = std::min(uw, 0.f);
        tendency(its, k, j) = tendency(its, k, j) -
                              rdx * (ub * (w_old(its + 1, k, j) - w_old(its, k, j)) +
                                     w(its, k, j) * (fzm(k) * (ru(its + 1, k, j) - ru(its, k, j)) + fzp(k) * (ru(its + 1, k - 1, j) - ru(its, k - 1, j))));
      }
    int k = ktf + 1;
    for (int j = j_start; j <= j_end; j++) {
      float uw = 0.5f * ((2.f - fzm(k - 1)) * (ru(its, k - 1, j) + ru(its + 1, k - 1, j)) - fzp(k - 1) * (ru(its, k - 2, j) + ru(its + 1, k - 2, j)));
      float ub = std::min(uw, 0.f);
      tendency(its, k, j) = tendency(its, k, j) -
                            rdx * (ub * (w_old(its + 1, k, j) - w_old(its, k, j)) +
                                   w(its, k, j) * ((2.f - fzm(k - 1)) * (ru(its + 1, k - 1, j) - ru(its, k - 1, j)) -
                                                   fzp(k - 1) * (ru(its + 1, k - 2, j) - ru(its, k - 2, j))));
    }
  }
  if (x.c.open_xe && ite == ide) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++) {
        float uw = 0.5f * (fzm(k) * (ru(ite - 1, k, j) + ru(ite, k, j)) + fzp(k) * (ru(ite - 1, k - 1, j) + ru(ite, k - 1, j)));
        float ub = std::max(uw, 0.f);
        tendency(i_end, k, j) = tendency(i_end, k, j) -
                                rdx * (ub * (w_old(i_end, k, j) - w_old(i_end - 1, k, j)) +
                                       w(i_end, k, j) * (fzm(k) * (ru(ite, k, j) - ru(ite - 1, k, j)) + fzp(k) * (ru(ite, k - 1, j) - ru(ite - 1, k - 1, j))));
      }
    int k = ktf + 1;
    for (int j = j_start; j <= j_end; j++) {
      float uw = 0.5f * ((2.f - fzm(k - 1)) * (ru(ite - 1, k - 1, j) + ru(ite, k - 1, j)) - fzp(k - 1) * (ru(ite - 1, k - 2, j) + ru(ite, k - 2, j)));
      float ub = std::max(uw, 0.f);
      tendency(i_end, k, j) = tendency(i_end, k, j) -
                              rdx * (ub * (w_old(i_end, k, j) - w_old(i_end - 1, k, j)) +
                                     w(i_end, k, j) * ((2.f - fzm(k - 1)) * (ru(ite, k - 1, j) - ru(ite - 1, k - 1, j)) -
                                                       fzp(k - 1) * (ru(ite, k - 2, j) - ru(ite - 1, k - 2, j))));
    }
  }
  if (x.c.open_ys && jts == jds) {
    for (int i = i_start; i <= i_end; i++)
      for (int k = kts + 1; k <= ktf; k++) {
        float vw = 0.5f * (fzm(k) * (rv(i, k, jts) + rv(i, k, jts + 1)) + fzp(k) * (rv(i, k - 1, jts) + rv(i, k - 1, jts + 1)));
        float vb = std::min(vw, 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) -
                              rdy * (vb * (w_old(i, k, jts + 1) - w_old(i, k, jts)) +
                                     w(i, k, jts) * (fzm(k) * (rv(i, k, jts + 1) - rv(i, k, jts)) + fzp(k) * (rv(i, k - 1, jts + 1) - rv(i, k - 1, jts))));
      }
    int k = ktf + 1;
    for (int i = i_start; i <= i_end; i++) {
      float vw = 0.5f * ((2.f - fzm(k - 1)) * (rv(i, k - 1, jts) + rv(i, k - 1, jts + 1)) - fzp(k - 1) * (rv(i, k - 2, jts) + rv(i, k - 2, jts + 1)));
      float vb = std::min(vw, 0.f);
      tendency(i, k, jts) = tendency(i, k, jts) -
                            rdy * (vb * (w_old(i, k, jts + 1) - w_old(i, k, jts)) +
                                   w(i, k, jts) * ((2.f - fzm(k - 1)) * (rv(i, k - 1, jts + 1) - rv(i, k - 1, jts)) -
                                                   fzp(k - 1) * (rv(i, k - 2, jts + 1) - rv(i, k - 2, jts))));
    }
  }
  if (x.c.open_ye && jte == jde) {
    for (int i = i_start; i <= i_end; i++)
      for (int k = kts + 1; k <= ktf; k++) {
        float vw = 0.5f * (fzm(k) * (rv(i, k, jte - 1) + rv(i, k, jte)) + fzp(k) * (rv(i, k - 1, jte - 1) + rv(i, k - 1, jte)));
        float vb = std::max(vw, 0.f);
        tendency(i, k, j_end) = tendency(i, k, j_end) -
                                rdy * (vb * (w_old(i, k, j_end) - w_old(i, k, j_end - 1)) +
                                       w(i, k, j_end) * (fzm(k) * (rv(i, k, jte) - rv(i, k, jte - 1)) + fzp(k) * (rv(i, k - 1, jte) - rv(i, k - 1, jte - 1))));
      }
    int k = ktf + 1;
    for (int i = i_start; i <= i_end; i++) {
      float vw = 0.5f * ((2.f - fzm(k - 1)) * (rv(i, k - 1, jte - 1) + rv(i, k - 1, jte)) - fzp(k - 1) * (rv(i, k - 2, jte - 1) + rv(i, k - 2, jte)));
      float vb = std::max(vw, 0.f);
      tendency(i, k, j_end) = tendency(i, k, j_end) -
                              rdy * (vb * (w_old(i, k, j_end) - w_old(i, k, j_end - 1)) +
                                     w(i, k, j_end) * ((2.f - fzm(k - 1)) * (rv(i, k - 1, jte) - rv(i, k - 1, jte - 1)) -
                                                       fzp(k - 1) * (rv(i, k - 2, jte) - rv(i, k - 2, jte - 1))));
    }
  }

  // vertical (:5822-6031)
  for (int i = i_start; i <= i_end; i++) { vflux(i, kts) = 0.f; vflux(i, kte) = 0.f; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 2; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vel = 0.5f * (rom(i, k, j) + rom(i, k - 1, j));
        vflux(i, k) = F.v35(k, kts + 3, ktf - 1, vel, w(i, k - 3, j), w(i, k - 2, j), w(i, k - 1, j), w(i, k, j), w(i, k + 1, j), w(i, k + 2, j));
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      vflux(i, k) = 0.25f * (rom(i, k, j) + rom(i, k - 1, j)) * (w(i, k, j) + w(i, k - 1, j));
      k = ktf + 1;
      vflux(i, k) = 0.25f * (rom(i, k, j) + rom(i, k - 1, j)) * (w(i, k, j) + w(i, k - 1, j));
    }
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tendency(i, k, j) - rdzu(k) * (vflux(i, k + 1) - vflux(i, k));
    int k = ktf + 1;
    for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tendency(i, k, j) + 2.f * rdzu(k - 1) * (vflux(i, k));
  }
}

// ---------------------------------------------------------------------------
// module_advect_em.F:6069-7885, horz_order 5 (:6505-6779), vert_order 3
// (:7595-7645), limiter (:7681-7776), flux divergence (:7778-7883)
// `weno` = advect_scalar_wenopd (:8573-9490, *_adv_opt = 4): the same routine with the 5th-order fluxes replaced by
// the WENO flux and, in the vertical, WENO on k = kts+3..ktf-2 (flux3 stays at kts+2 and ktf-1) (:9217-9262).
void advect_scalar_pd(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom,
                      A2 mut, A2 mub, A2 mu_old, float dt, bool weno) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, kde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const A2 &msftx = x.msftx, &msfty = x.msfty;
  const float rdx = x.rdx, rdy = x.rdy;
  const float sgn = (float)(x.time_step >= 0 ? 1 : -1);
  const float eps = 1.e-20f;
  // statement functions :6169-6186 (note: different rounding from advect_scalar)
  auto flux4 = [](float q_im2, float q_im1, float q_i, float q_ip1) {
    return (7.f / 12.f) * (q_i + q_im1) - (1.f / 12.f) * (q_ip1 + q_im2);
  };
  // h_sca_adv_order = 6 / v_sca_adv_order = 4: the same code with flux6 / flux4 (:6229-6503, :7544-7593); the WENO variant
  // keeps its own structure whatever the orders say
  const bool h_even = x.c.h_sca_adv_order == 6 && !weno, v_even = x.c.v_sca_adv_order == 4 && !weno;
  auto flux3o = [&](float q_im2, float q_im1, float q_i, float q_ip1, float ua) {
    return flux4(q_im2, q_im1, q_i, q_ip1) + sgn * sign1(ua) * (1.f / 12.f) * ((q_ip1 - q_im2) - 3.f * (q_i - q_im1));
  };
  auto flux3 = [&](float q_im2, float q_im1, float q_i, float q_ip1, float ua) {   // horizontal faces
    return h_even ? flux4(q_im2, q_im1, q_i, q_ip1) : flux3o(q_im2, q_im1, q_i, q_ip1, ua);
  };
  auto vflux3 = [&](float q_im2, float q_im1, float q_i, float q_ip1, float ua) {  // vertical faces
    return v_even ? flux4(q_im2, q_im1, q_i, q_ip1) : flux3o(q_im2, q_im1, q_i, q_ip1, ua);
  };
  auto flux6 = [](float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2) {
    return (37.f / 60.f) * (q_i + q_im1) - (2.f / 15.f) * (q_ip1 + q_im2) + (1.f / 60.f) * (q_ip2 + q_im3);
  };
  auto flux5 = [&](float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2, float ua) {
    if (h_even) return flux6(q_im3, q_im2, q_im1, q_i, q_ip1, q_ip2);
    return flux6(q_im3, q_im2, q_im1, q_i, q_ip1, q_ip2) -
           sgn * sign1(ua) * (1.f / 60.f) * ((q_ip2 - q_im3) - 5.f * (q_ip1 - q_im2) + 10.f * (q_i - q_im1));
  };
  auto flux_upwind = [](float q_im1, float q_i, float cr) {
    return 0.5f * std::min(1.0f, (cr + std::fabs(cr))) * q_im1 + 0.5f * std::max(-1.0f, (cr - std::fabs(cr))) * q_i;
  };
  // (its-1:ite+2, kts:kte, jts-1:jte+2)
  const int ni = ite - its + 4, nk = kte - kts + 1, nj = jte - jts + 4;
  const size_t n = (size_t)ni * nk * nj;
  std::vector<float> b_fqx(n, 0.f), b_fqy(n, 0.f), b_fqz(n, 0.f), b_fqxl(n, 0.f), b_fqyl(n, 0.f), b_fqzl(n, 0.f), b_fo(n, 0.f), b_pl(n, 0.f);
  auto idx = [&](int i, int k, int j) { return (size_t)(i - its + 1) + (size_t)(k - kts) * ni + (size_t)(j - jts + 1) * ni * nk; };
#define L3(name, buf) auto name = [&](int i, int k, int j) -> float & { return buf[idx(i, k, j)]; }
  L3(fqx, b_fqx); L3(fqy, b_fqy); L3(fqz, b_fqz); L3(fqxl, b_fqxl); L3(fqyl, b_fqyl); L3(fqzl, b_fqzl); L3(flux_out, b_fo); L3(ph_low, b_pl);
#undef L3
  bool degrade_xs = true, degrade_xe = true, degrade_ys = true, degrade_ye = true;
  if (x.c.periodic_x || x.c.symmetric_xs || (its > ids + 3)) degrade_xs = false;
  if (x.c.periodic_x || x.c.symmetric_xe || (ite < ide - 4)) degrade_xe = false;
  if (x.c.periodic_y || x.c.symmetric_ys || (jts > jds + 3)) degrade_ys = false;
  if (x.c.periodic_y || x.c.symmetric_ye || (jte < jde - 4)) degrade_ye = false;

  // y fluxes
  int i_start = its - 1, i_end = std::min(ite, ide - 1) + 1, j_start = jts - 1, j_end = std::min(jte, jde - 1) + 1;
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) { j_start = std::max(jts - 1, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte + 1, jde - 2); j_end_f = jde - 3; }
  for (int j = j_start; j <= j_end + 1; j++) {
    int mode;  // 5, 3, 2 or 0 (none)
    if (j >= j_start_f && j <= j_end_f) mode = 5;
    else if (j == jds + 1) mode = 2;
    else if (j == jds + 2) mode = 3;
    else if (j == jde - 1) mode = 2;
    else if (j == jde - 2) mode = 3;
    else mode = 0;
    if (!mode) continue;
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float dy = 2.f / (msftx(i, j) + msftx(i, j - 1)) / rdy;
        float mu = 0.5f * ((c1(k) * mut(i, j) + c2(k)) + (c1(k) * mut(i, j - 1) + c2(k)));
        float vel = rv(i, k, j);
        float cr = vel * dt / dy / mu;
        fqyl(i, k, j) = mu * (dy / dt) * flux_upwind(field_old(i, k, j - 1), field_old(i, k, j), cr);
        if (mode == 5 && weno)
          fqy(i, k, j) = weno5(vel, vel * sgn >= 0.0f, field(i, k, j - 3), field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), field(i, k, j + 2));
        else if (mode == 5)
          fqy(i, k, j) = vel * flux5(field(i, k, j - 3), field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), field(i, k, j + 2), vel);
        else if (mode == 2)
          fqy(i, k, j) = 0.5f * rv(i, k, j) * (field(i, k, j) + field(i, k, j - 1));
        else
          fqy(i, k, j) = vel * flux3(field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), vel);
        fqy(i, k, j) = fqy(i, k, j) - fqyl(i, k, j);
      }
  }
  // x fluxes
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1;
  int i_start_f = i_start, i_end_f = i_end + 1;
  j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  if (degrade_xs) { i_start = std::max(ids + 1, its - 1); i_start_f = ids + 3; }
  if (degrade_xe) { i_end = std::min(ide - 2, ite + 1); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    auto doface = [&](int i, int k, int mode) {
      float dx = 2.f / (msfty(i, j) + msfty(i - 1, j)) / rdx;
      float mu = 0.5f * ((c1(k) * mut(i, j) + c2(k)) + (c1(k) * mut(i - 1, j) + c2(k)));
      float vel, cr;
      if (mode == 2 && i == ids + 1) {   // :6683-6684 (west 2nd-order face divides by mu first)
        vel = ru(i, k, j) / mu;
        cr = vel * dt / dx;
      } else {
        vel = ru(i, k, j);
        cr = vel * dt / dx / mu;
      }
      fqxl(i, k, j) = mu * (dx / dt) * flux_upwind(field_old(i - 1, k, j), field_old(i, k, j), cr);
      if (mode == 5 && weno)
        fqx(i, k, j) = weno5(vel, vel * sgn >= 0.0f, field(i - 3, k, j), field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), field(i + 2, k, j));
      else if (mode == 5)
        fqx(i, k, j) = vel * flux5(field(i - 3, k, j), field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), field(i + 2, k, j), vel);
      else if (mode == 2)
        fqx(i, k, j) = 0.5f * (ru(i, k, j)) * (field(i, k, j) + field(i - 1, k, j));
      else
        fqx(i, k, j) = vel * flux3(field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), vel);
      fqx(i, k, j) = fqx(i, k, j) - fqxl(i, k, j);
    };
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) doface(i, k, 5);
    if (degrade_xs)
      for (int i = i_start; i <= i_start_f - 1; i++) {
        if (i == ids + 1) for (int k = kts; k <= ktf; k++) doface(i, k, 2);
        if (i == ids + 2) for (int k = kts; k <= ktf; k++) doface(i, k, 3);
      }
    if (degrade_xe)
      for (int i = i_end_f + 1; i <= i_end + 1; i++) {
        if (i == ide - 1) for (int k = kts; k <= ktf; k++) doface(i, k, 2);
        if (i == ide - 2) for (int k = kts; k <= ktf; k++) doface(i, k, 3);
      }
  }

  // open boundary terms (:7255-7318)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::min(0.5f * (ru(its, k, j) + ru(its + 1, k, j)), 0.f);
        tendency(its, k, j) = tendency(its, k, j) -
                              rdx * (ub * (field_old(its + 1, k, j) - field_old(its, k, j)) + field(its, k, j) * (ru(its + 1, k, j) - ru(its, k, j)));
      }
  if (x.c.open_xe && ite == ide)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::max(0.5f * (ru(ite - 1, k, j) + ru(ite, k, j)), 0.f);
        tendency(i_end, k, j) = tendency(i_end, k, j) -
                                rdx * (ub * (field_old(i_end, k, j) - field_old(i_end - 1, k, j)) + field(i_end, k, j) * (ru(ite, k, j) - ru(ite - 1, k, j)));
      }
  if (x.c.open_ys && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::min(0.5f * (rv(i, k, jts) + rv(i, k, jts + 1)), 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) -
                              rdy * (vb * (field_old(i, k, jts + 1) - field_old(i, k, jts)) + field(i, k, jts) * (rv(i, k, jts + 1) - rv(i, k, jts)));
      }
  if (x.c.open_ye && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::max(0.5f * (rv(i, k, jte - 1) + rv(i, k, jte)), 0.f);
        tendency(i, k, j_end) = tendency(i, k, j_end) -
                                rdy * (vb * (field_old(i, k, j_end) - field_old(i, k, j_end - 1)) + field(i, k, j_end) * (rv(i, k, jte) - rv(i, k, jte - 1)));
      }

  // vertical fluxes
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1; j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  for (int j = j_start; j <= j_end; j++) {
    for (int i = i_start; i <= i_end; i++) {
      fqz(i, 1, j) = 0.f; fqzl(i, 1, j) = 0.f; fqz(i, kde, j) = 0.f; fqzl(i, kde, j) = 0.f;
    }
    auto dolow = [&](int i, int k) {
      float dz = 2.f / (rdzw(k) + rdzw(k - 1));
      float mu = 0.5f * ((c1(k) * mut(i, j) + c2(k)) + (c1(k) * mut(i, j) + c2(k)));
      float vel = rom(i, k, j);
      float cr = vel * dt / dz / mu;
      fqzl(i, k, j) = mu * (dz / dt) * flux_upwind(field_old(i, k - 1, j), field_old(i, k, j), cr);
    };
    for (int k = kts + 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        dolow(i, k);
        float vel = rom(i, k, j);
        if (weno && k >= kts + 3 && k <= ktf - 2)
          fqz(i, k, j) = weno5(vel, -vel * sgn >= 0.0f, field(i, k - 3, j), field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), field(i, k + 2, j));
        else
          fqz(i, k, j) = vel * vflux3(field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), -vel);
        fqz(i, k, j) = fqz(i, k, j) - fqzl(i, k, j);
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      dolow(i, k);
      fqz(i, k, j) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
      fqz(i, k, j) = fqz(i, k, j) - fqzl(i, k, j);
      k = ktf;
      dolow(i, k);
      fqz(i, k, j) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
      fqz(i, k, j) = fqz(i, k, j) - fqzl(i, k, j);
    }
  }

  // positive definite filter
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1; j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  if (x.spec) {   // :7697-7702
    if (degrade_xs) i_start = std::max(its - 1, ids + 1);
    if (degrade_xe) i_end = std::min(ite + 1, ide - 2);
    if (degrade_ys) j_start = std::max(jts - 1, jds + 1);
    if (degrade_ye) j_end = std::min(jte + 1, jde - 2);
  }
  if (x.c.open_xs && degrade_xs) i_start = std::max(its - 1, ids + 1);
  if (x.c.open_xe && degrade_xe) i_end = std::min(ite + 1, ide - 2);
  if (x.c.open_ys && degrade_ys) j_start = std::max(jts - 1, jds + 1);
  if (x.c.open_ye && degrade_ye) j_end = std::min(jte + 1, jde - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        ph_low(i, k, j) = ((c1(k) * mub(i, j) + c2(k)) + (c1(k) * mu_old(i, j))) * field_old(i, k, j) -
                          dt * (msftx(i, j) * msfty(i, j) * (rdx * (fqxl(i + 1, k, j) - fqxl(i, k, j)) + rdy * (fqyl(i, k, j + 1) - fqyl(i, k, j))) +
                                msfty(i, j) * rdzw(k) * (fqzl(i, k + 1, j) - fqzl(i, k, j)));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        flux_out(i, k, j) = dt * ((msftx(i, j) * msfty(i, j)) * (rdx * (std::max(0.f, fqx(i + 1, k, j)) - std::min(0.f, fqx(i, k, j))) +
                                                                 rdy * (std::max(0.f, fqy(i, k, j + 1)) - std::min(0.f, fqy(i, k, j)))) +
                                  msfty(i, j) * rdzw(k) * (std::min(0.f, fqz(i, k + 1, j)) - std::max(0.f, fqz(i, k, j))));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        if (flux_out(i, k, j) > ph_low(i, k, j)) {
          float scale = std::max(0.f, ph_low(i, k, j) / (flux_out(i, k, j) + eps));
          if (fqx(i + 1, k, j) > 0.f) fqx(i + 1, k, j) = scale * fqx(i + 1, k, j);
          if (fqx(i, k, j) < 0.f) fqx(i, k, j) = scale * fqx(i, k, j);
          if (fqy(i, k, j + 1) > 0.f) fqy(i, k, j + 1) = scale * fqy(i, k, j + 1);
          if (fqy(i, k, j) < 0.f) fqy(i, k, j) = scale * fqy(i, k, j);
          if (fqz(i, k + 1, j) < 0.f) fqz(i, k + 1, j) = scale * fqz(i, k + 1, j);
          if (fqz(i, k, j) > 0.f) fqz(i, k, j) = scale * fqz(i, k, j);
        }

  // flux divergence
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - rdzw(k) * (fqz(i, k + 1, j) - fqz(i, k, j) + fqzl(i, k + 1, j) - fqzl(i, k, j));
  if (degrade_xs) i_start = std::max(its, ids + 1);
  if (degrade_xe) i_end = std::min(ite, ide - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - msftx(i, j) * (rdx * (fqx(i + 1, k, j) - fqx(i, k, j) + fqxl(i + 1, k, j) - fqxl(i, k, j)));
  i_start = its; i_end = std::min(ite, ide - 1);
  if (degrade_ys) j_start = std::max(jts, jds + 1);
  if (degrade_ye) j_end = std::min(jte, jde - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - msftx(i, j) * (rdy * (fqy(i, k, j + 1) - fqy(i, k, j) + fqyl(i, k, j + 1) - fqyl(i, k, j)));
}

// module_advect_em.F:7889-8569: 5th-order WENO in all three directions, "adapted from COMMAS";
// *_adv_opt = 3 (WENO_SCALAR) on the last RK substep, and theta with it (module_em.F:603-621).  Loop structure,
// degraded faces next to non-periodic boundaries and the open-boundary terms are advect_scalar's; horz_order and
// vert_order are forced to 5 (:8009-8010): WENO for k = kts+3..ktf-2, flux3 at kts+2 and ktf-1, 2nd order at
// kts+1 and ktf.  The statement functions are the pre-divided ones (:7974-7981), flux3 without sign(time_step).
void advect_scalar_weno(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, x.kde - 1);
  const A1 &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy;
  const float sgn_ts = (float)(x.time_step >= 0 ? 1 : -1);
  auto flux3 = [](float q_im2, float q_im1, float q_i, float q_ip1, float ua) {
    return ((7.f / 12.f) * (q_i + q_im1) - (1.f / 12.f) * (q_ip1 + q_im2)) + sign1(ua) * (1.f / 12.f) * ((q_ip1 - q_im2) - 3.f * (q_i - q_im1));
  };
  bool degrade_xs, degrade_xe, degrade_ys, degrade_ye;
  degrade_flags(x, 3, 4, degrade_xs, degrade_xe, degrade_ys, degrade_ye);   // :8017-8031
  Row fq0(its, ite, kte), fq1(its, ite, kte), fqx(its, ite, kte), vflux(its, ite, kte);

  // y
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_ys) { j_start = std::max(jts, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte, jde - 2); j_end_f = jde - 3; }
  Row *jp1 = &fq1, *jp0 = &fq0;
  for (int j = j_start; j <= j_end + 1; j++) {
    Row &fqy = *jp1;
    const bool full = j >= j_start_f && j <= j_end_f;
    const bool second = !full && (j == jds + 1 || (j == jde - 1 && j != jds + 2)), third = !full && !second && (j == jds + 2 || j == jde - 2);
    if (full || second || third)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          if (full) {
            float vel = 0.5f * (rv(i, k, j) + rv(i, k, j));     // is = js = ks = 0
            fqy(i, k) = weno5(vel, vel * sgn_ts >= 0.0f, field(i, k, j - 3), field(i, k, j - 2), field(i, k, j - 1), field(i, k, j),
                              field(i, k, j + 1), field(i, k, j + 2));
          } else if (second) {
            fqy(i, k) = 0.5f * rv(i, k, j) * (field(i, k, j) + field(i, k, j - 1));
          } else {
            float vel = rv(i, k, j);
            fqy(i, k) = vel * flux3(field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), vel);
          }
        }
    if (j > j_start)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          float mrdy = x.msftx(i, j - 1) * rdy;
          tendency(i, k, j - 1) = tendency(i, k, j - 1) - mrdy * ((*jp1)(i, k) - (*jp0)(i, k));
        }
    std::swap(jp1, jp0);
  }

  // x
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  int i_start_f = i_start, i_end_f = i_end + 1;
  if (degrade_xs) { i_start = std::max(ids + 1, its); i_start_f = std::min(i_start + 2, ids + 3); }
  if (degrade_xe) { i_end = std::min(ide - 2, ite); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) {
        float vel = 0.5f * (ru(i, k, j) + ru(i, k, j));
        fqx(i, k) = weno5(vel, vel * sgn_ts >= 0.0f, field(i - 3, k, j), field(i - 2, k, j), field(i - 1, k, j), field(i, k, j),
                          field(i + 1, k, j), field(i + 2, k, j));
      }
    auto low = [&](int i) {
      if (i == ids + 1 || i == ide - 1)
        for (int k = kts; k <= ktf; k++) fqx(i, k) = 0.5f * (ru(i, k, j)) * (field(i, k, j) + field(i - 1, k, j));
      if (i == ids + 2 || i == ide - 2)
        for (int k = kts; k <= ktf; k++) {
          float vel = ru(i, k, j);
          fqx(i, k) = vel * flux3(field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), vel);
        }
    };
    if (degrade_xs)
      for (int i = i_start; i <= i_start_f - 1; i++) if (i == ids + 1 || i == ids + 2) low(i);
    if (degrade_xe)
      for (int i = i_end_f + 1; i <= i_end + 1; i++) if (i == ide - 1 || i == ide - 2) low(i);
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mrdx = x.msftx(i, j) * rdx;
        tendency(i, k, j) = tendency(i, k, j) - mrdx * (fqx(i + 1, k) - fqx(i, k));
      }
  }

  // open boundary terms (:8402-8466)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::min(0.5f * (ru(its, k, j) + ru(its + 1, k, j)), 0.f);
        tendency(its, k, j) = tendency(its, k, j) -
                              rdx * (ub * (field_old(its + 1, k, j) - field_old(its, k, j)) + field(its, k, j) * (ru(its + 1, k, j) - ru(its, k, j)));
      }
  if (x.c.open_xe && ite == ide)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::max(0.5f * (ru(ite - 1, k, j) + ru(ite, k, j)), 0.f);
        tendency(i_end, k, j) = tendency(i_end, k, j) -
                                rdx * (ub * (field_old(i_end, k, j) - field_old(i_end - 1, k, j)) + field(i_end, k, j) * (ru(ite, k, j) - ru(ite - 1, k, j)));
      }
  if (x.c.open_ys && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::min(0.5f * (rv(i, k, jts) + rv(i, k, jts + 1)), 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) -
                              rdy * (vb * (field_old(i, k, jts + 1) - field_old(i, k, jts)) + field(i, k, jts) * (rv(i, k, jts + 1) - rv(i, k, jts)));
      }
  if (x.c.open_ye && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::max(0.5f * (rv(i, k, jte - 1) + rv(i, k, jte)), 0.f);
        tendency(i, k, j_end) = tendency(i, k, j_end) -
                                rdy * (vb * (field_old(i, k, j_end) - field_old(i, k, j_end - 1)) + field(i, k, j_end) * (rv(i, k, jte) - rv(i, k, jte - 1)));
      }

  // vertical (:8476-8565)
  for (int i = i_start; i <= i_end; i++) { vflux(i, kts) = 0.f; vflux(i, kte) = 0.f; }
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts + 3; k <= ktf - 2; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vel = 0.5f * (rom(i, k, j) + rom(i, k, j));
        vflux(i, k) = weno5(vel, -vel * sgn_ts >= 0.0f, field(i, k - 3, j), field(i, k - 2, j), field(i, k - 1, j), field(i, k, j),
                            field(i, k + 1, j), field(i, k + 2, j));
      }
    for (int i = i_start; i <= i_end; i++) {
      int k = kts + 1;
      vflux(i, k) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
      k = kts + 2;
      float vel = rom(i, k, j);
      vflux(i, k) = vel * flux3(field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), -vel);
      k = ktf - 1;
      vel = rom(i, k, j);
      vflux(i, k) = vel * flux3(field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), -vel);
      k = ktf;
      vflux(i, k) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tendency(i, k, j) - rdzw(k) * (vflux(i, k + 1) - vflux(i, k));
  }
}

}  // namespace orc

// the flux operators by their order, for tests/test_reference_tables.py (which evaluates the reference's own statement
// functions, module_advect_em.F:202-213, from their source text): q = q_im3..q_ip2 (orders 5, 6) or q_im2..q_ip1 (3, 4)
extern "C" float oracle_flux(int order, const float *q, float ua, int time_step) {
  orc::Flux F{(float)(time_step >= 0 ? 1 : -1), false, false, false};
  switch (order) {
    case 3: return F.flux3(q[0], q[1], q[2], q[3], ua);
    case 4: return F.flux4(q[0], q[1], q[2], q[3]);
    case 5: return F.flux5(q[0], q[1], q[2], q[3], q[4], q[5], ua);
    default: return F.flux6(q[0], q[1], q[2], q[3], q[4], q[5]);
  }
}
