// tend.cu -- rk_tendency + rk_addtend_dry (dyn_em/module_em.F:190-1092) as ONE
// gather kernel per tendency field: each thread owns one (i,k,j) point of
// ru_tend / rv_tend / rw_tend / t_tend / ph_tend, evaluates every contribution in
// the reference's accumulation order in registers and writes the result once.
// The reference zeroes 10 arrays and then makes ~12 read-modify-write sweeps
// (advect_*, rhs_ph, horizontal_pressure_gradient, pg_buoy_w, w_damp, coriolis,
// curvature, rk_addtend_dry); here each tendency is written exactly once and the
// stencil neighbours come from L1/L2.
//
// Face fluxes follow dyn_em/module_advect_em.F for h order 5 / v order 3: a face
// uses flux5 unless it lies within two points of a non-periodic, non-symmetric
// domain edge, where the order degrades to 3 and 2 (:568-605, :692-734).  The
// rule is evaluated with DOMAIN indices so that any patch decomposition gives
// the same bits as the single-patch run.
#include "common.cuh"
#include "libm_port.h"
#include "flux.cuh"
#include "weno.cuh"

namespace wg {

struct Deg { bool xs, xe, ys, ye; };
__device__ __forceinline__ Deg degrade(const Geo &g) {
  Deg d;
  d.xs = !(g.periodic_x || g.sym_xs);
  d.xe = !(g.periodic_x || g.sym_xe);
  d.ys = !(g.periodic_y || g.sym_ys);
  d.ye = !(g.periodic_y || g.sym_ye);
  return d;
}
// A point at least 4 cells inside every non-periodic edge sees none of the boundary logic: every
// face is 5th order, every loop range of the reference contains it and no open-boundary term
// applies.  The kernels instantiate their body twice and take the branch-free one there (ncu r1b:
// the gather kernels were issue-bound on the integer pipe by exactly this logic).
__device__ __forceinline__ bool interior_pt(const Geo &g, int i, int j) {
  const bool ix = g.periodic_x ? (i <= g.ide - 1) : (i >= g.ids + 4 && i <= g.ide - 5);
  const bool jy = g.periodic_y ? (j <= g.jde - 1) : (j >= g.jds + 4 && j <= g.jde - 5);
  return ix && jy;
}
// order of the flux at a face index `f`; lo2/lo3/hi2/hi3 are the face indices that
// degrade to 2nd / 3rd order at the low / high edge.  returns 5, 3, 2
__device__ __forceinline__ int face_order(int f, bool dlo, bool dhi, int lo2, int lo3, int hi2, int hi3) {
  const int lo5 = lo3 + 1, hi5 = hi3 - 1;
  if ((!dlo || f >= lo5) && (!dhi || f <= hi5)) return 5;
  if (f == lo2) return 2;
  if (f == lo3) return 3;
  if (f == hi2) return 2;
  if (f == hi3) return 3;
  return 5;   // unreachable for the faces the callers ask for
}

// horizontal flux of q through x-face ii (between ii-1 and ii) at level k, row j
#define XSTENCIL(q, ii, k, j) A3(q, (ii)-3, k, j), A3(q, (ii)-2, k, j), A3(q, (ii)-1, k, j), A3(q, ii, k, j), A3(q, (ii) + 1, k, j), A3(q, (ii) + 2, k, j)
#define YSTENCIL(q, i, k, jj) A3(q, i, k, (jj)-3), A3(q, i, k, (jj)-2), A3(q, i, k, (jj)-1), A3(q, i, k, jj), A3(q, i, k, (jj) + 1), A3(q, i, k, (jj) + 2)

// ---------------------------------------------------------------------------
// dpn: pressure at w levels averaged to a u (di=1) or v (dj=1) point,
// module_big_step_utilities_em.F:2247-2266 / :2340-2359
__device__ __forceinline__ float dpn_at(const Geo &g, cfp p, int i, int kk, int j, int di, int dj) {
  const int i2 = i - di, j2 = j - dj;
  if (kk == 1)
    return .5f * (g.cf1 * (A3(p, i2, 1, j2) + A3(p, i, 1, j)) + g.cf2 * (A3(p, i2, 2, j2) + A3(p, i, 2, j)) +
                  g.cf3 * (A3(p, i2, 3, j2) + A3(p, i, 3, j)));
  if (kk == g.kde) {
    if (!g.top_lid) return 0.f;
    return .5f * (g.cfn * (A3(p, i2, g.kde - 1, j2) + A3(p, i, g.kde - 1, j)) + g.cfn1 * (A3(p, i2, g.kde - 2, j2) + A3(p, i, g.kde - 2, j)));
  }
  return .5f * (g.fnm[kk] * (A3(p, i2, kk, j2) + A3(p, i, kk, j)) + g.fnp[kk] * (A3(p, i2, kk - 1, j2) + A3(p, i, kk - 1, j)));
}

// vxgm of curvature (:4230-4290) with the open-boundary copies folded into index clamps
template <bool FAST>
__device__ __forceinline__ float vxgm_at(const Geo &g, cfp u, cfp v, int i, int k, int j) {
  if (!FAST && (g.open_xs || (g.spec && !g.periodic_x)) && i == g.ids - 1) i = g.ids;
  if (!FAST && (g.open_xe || (g.spec && !g.periodic_x)) && i == g.ide) i = g.ide - 1;
  if (!FAST && g.oys() && j == g.jds - 1) j = g.jds;
  if (!FAST && g.oye() && j == g.jde) j = g.jde - 1;
  return 0.5f * (A3(u, i, k, j) + A3(u, i + 1, k, j)) * (A2(g.msfvx, i, j + 1) - A2(g.msfvx, i, j)) * g.rdy -
         0.5f * (A3(v, i, k, j) + A3(v, i, k, j + 1)) * (A2(g.msfuy, i + 1, j) - A2(g.msfuy, i, j)) * g.rdx;
}

struct TendIn {
  cfp u, v, w, t, ph, phb, ru, rv, rw, ww;
  cfp p, pb, al, alt, php, cqu, cqv, cqw, h_diabatic;
  cfp mut, muu, muv, mu, mub;
  cfp ru_tendf, rv_tendf, rw_tendf, ph_tendf, t_tendf;
  cfp rom;   // the vertical mass flux the explicit advection sees: ww, or wwE under IEVA (module_ieva_em.F:49)
};

// perturbation_coriolis (:3854-4173): ru, rv minus the mass-coupled base-state wind interpolated
// to the height of the u / v point.  The reference fills tile-local arrays first; evaluating the
// same expression at the point of use gives the same bits.
__device__ __forceinline__ float pert_wind(const Geo &g, cfp r_in, cfp ph, cfp phb, cfp mass2d, const float *base, int i, int k, int j,
                                           int di, int dj) {
  const int ktf = g.kde - 1;
  const float z = 0.25f * (A3(phb, i, k, j) + A3(phb, i, k + 1, j) + A3(phb, i - di, k, j - dj) + A3(phb, i - di, k + 1, j - dj) +
                           A3(ph, i, k, j) + A3(ph, i, k + 1, j) + A3(ph, i - di, k, j - dj) + A3(ph, i - di, k + 1, j - dj)) / WG_G;
  const float *zb = g.z_base;
  const float mass = (g.c1h[k] * A2(mass2d, i, j) + g.c2h[k]);
  float b;
  if (k == 1) {
    const float wkp1 = fminf(1.f, fmaxf(0.f, z - zb[k]) / (zb[k + 1] - zb[k]));
    const float wk = 1.f - wkp1;
    b = (wk * base[k] + wkp1 * base[k + 1]);
  } else if (k == ktf) {
    const float wkm1 = fminf(1.f, fmaxf(0.f, zb[k] - z) / (zb[k] - zb[k - 1]));
    const float wk = 1.f - wkm1;
    b = (wkm1 * base[k - 1] + wk * base[k]);
  } else {
    const float wkp1 = fminf(1.f, fmaxf(0.f, z - zb[k]) / (zb[k + 1] - zb[k]));
    const float wkm1 = fminf(1.f, fmaxf(0.f, zb[k] - z) / (zb[k] - zb[k - 1]));
    const float wk = 1.f - wkp1 - wkm1;
    b = (wkm1 * base[k - 1] + wk * base[k] + wkp1 * base[k + 1]);
  }
  return A3(r_in, i, k, j) - mass * b;
}
#define PRV(i, k, j) pert_wind(g, s.rv, s.ph, s.phb, s.muv, g.v_base, i, k, j, 0, 1)
#define PRU(i, k, j) pert_wind(g, s.ru, s.ph, s.phb, s.muu, g.u_base, i, k, j, 1, 0)

// ===========================================================================
// ru_tend: advect_u (:126) + pgf x (:2320-2400) + coriolis (:3700-3735) + curvature + addtend
template <bool FAST, bool WENO = false, int PART = 0>   // PART (IEVA): 1 = advection only, 2 = everything after it, starting from t0
__device__ __forceinline__ float tend_u_pt(const Geo &g, const TendIn &s, int i, int k, int j, float t0 = 0.f) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const Deg dg = degrade(g);
  const float sg = g.sgn_ts;
  // even orders (h 6 / v 4) are the odd ones without the upwind term: its factor is 0 then; WENO keeps sign(time_step)
  const float sgh = WENO ? sg : g.sg_hm, sgv = WENO ? sg : g.sg_vm;
  cfp u = s.u, ru = s.ru, rv = s.rv, rom = s.rom;
  float t = t0;
  if (PART != 2) {
  // --- y flux divergence
  {
    int i_s = its, i_e = ite;
    if (g.oxs()) i_s = max(ids + 1, its);
    if (g.oxe()) i_e = min(ide - 1, ite);
    if (g.periodic_x) { i_s = its; i_e = ite; }
    const int j_lo = dg.ys ? max(jts, jds + 1) : jts, j_hi = dg.ye ? min(jte, jde - 2) : min(jte, jde - 1);
    if (FAST || (i >= i_s && i <= i_e && j >= j_lo && j <= j_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int jj = j + m;
        const int ord = FAST ? 5 : face_order(jj, dg.ys, dg.ye, jds + 1, jds + 2, jde - 1, jde - 2);
        if (ord == 5) {
          const float vel = 0.5f * (A3(rv, i, k, jj) + A3(rv, i - 1, k, jj));
          fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, YSTENCIL(u, i, k, jj), 1.0e-18) : vel * flux5(YSTENCIL(u, i, k, jj), vel, sgh);
        } else if (ord == 3) {
          const float vel = 0.5f * (A3(rv, i, k, jj) + A3(rv, i - 1, k, jj));
          fq[m] = vel * flux3(A3(u, i, k, jj - 2), A3(u, i, k, jj - 1), A3(u, i, k, jj), A3(u, i, k, jj + 1), vel, sgh);
        } else {
          fq[m] = 0.25f * (A3(rv, i, k, jj) + A3(rv, i - 1, k, jj)) * (A3(u, i, k, jj) + A3(u, i, k, jj - 1));
        }
      }
      const float mrdy = A2(g.msfux, i, j) * g.rdy;
      t = t - mrdy * (fq[1] - fq[0]);
    }
  }
  // --- x flux divergence
  {
    const int i_lo = dg.xs ? max(ids + 1, its) : its, i_hi = dg.xe ? min(ide - 1, ite) : ite;
    if (FAST || (i >= i_lo && i <= i_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int ii = i + m;
        const int ord = FAST ? 5 : face_order(ii, dg.xs, dg.xe, ids + 1, ids + 2, ide, ide - 1);
        if (ord == 5) {
          const float vel = 0.5f * (A3(ru, ii, k, j) + A3(ru, ii - 1, k, j));
          fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, XSTENCIL(u, ii, k, j), 1.0e-18) : vel * flux5(XSTENCIL(u, ii, k, j), vel, sgh);
        } else if (ord == 3) {
          const float vel = 0.5f * (A3(ru, ii, k, j) + A3(ru, ii - 1, k, j));
          fq[m] = vel * flux3(A3(u, ii - 2, k, j), A3(u, ii - 1, k, j), A3(u, ii, k, j), A3(u, ii + 1, k, j), vel, sgh);
        } else {
          float ua = A3(u, ii, k, j), ub = A3(u, ii - 1, k, j);
          if (g.spec) {   // specified uses upstream normal wind at boundaries (:698, :720)
            if (ii == ids + 1) { if (ua < 0.f) ub = ua; }
            else if (ub > 0.f) ua = ub;
          }
          fq[m] = 0.25f * (A3(ru, ii, k, j) + A3(ru, ii - 1, k, j)) * (ua + ub);
        }
      }
      const float mrdx = A2(g.msfux, i, j) * g.rdx;
      t = t - mrdx * (fq[1] - fq[0]);
    }
  }
  // --- radiative / open boundary terms (:1244-1334)
  if (!FAST && g.open_xs && i == ids) {
    const float ub = fminf(A3(ru, ids, k, j) - WG_CB * (g.c1h[k] * A2(s.mut, ids, j) + g.c2h[k]), 0.f);
    t = t - g.rdx * ub * (A3(u, ids + 1, k, j) - A3(u, ids, k, j));
  }
  if (!FAST && g.open_xe && i == ide) {
    const float ub = fmaxf(A3(ru, ide, k, j) + WG_CB * (g.c1h[k] * A2(s.mut, ide - 1, j) + g.c2h[k]), 0.f);
    t = t - g.rdx * ub * (A3(u, ide, k, j) - A3(u, ide - 1, k, j));
  }
  {
    int i_s = its, i_e = min(ite, ide);
    if (g.open_xs) i_s = max(ids + 1, its);
    if (g.open_xe) i_e = min(ite, ide - 1);
    const int imin = ids, imax = ide - 1;
    if (FAST || (i >= i_s && i <= i_e)) {
      const int ip = min(imax, i), im = max(imin, i - 1);
      if (!FAST && g.open_ys && j == jds) {
        const float mrdy = A2(g.msfux, i, jds) * g.rdy;
        const float vw = 0.5f * (A3(rv, ip, k, jds) + A3(rv, im, k, jds));
        const float vb = fminf(vw, 0.f);
        const float dvm = A3(rv, ip, k, jds + 1) - A3(rv, ip, k, jds);
        const float dvp = A3(rv, im, k, jds + 1) - A3(rv, im, k, jds);
        t = t - mrdy * (vb * (A3(u, i, k, jds + 1) - A3(u, i, k, jds)) + 0.5f * A3(u, i, k, jds) * (dvm + dvp));
      }
      if (!FAST && g.open_ye && j == jde - 1) {
        const float mrdy = A2(g.msfux, i, jde - 1) * g.rdy;
        const float vw = 0.5f * (A3(rv, ip, k, jde) + A3(rv, im, k, jde));
        const float vb = fmaxf(vw, 0.f);
        const float dvm = A3(rv, ip, k, jde) - A3(rv, ip, k, jde - 1);
        const float dvp = A3(rv, im, k, jde) - A3(rv, im, k, jde - 1);
        t = t - mrdy * (vb * (A3(u, i, k, jde - 1) - A3(u, i, k, jde - 2)) + 0.5f * A3(u, i, k, jde - 1) * (dvm + dvp));
      }
    }
  }
  // --- vertical (:1336-1498; the reference tests open_ys/open_ye for the i range)
  {
    int i_s = its, i_e = ite;
    if (g.oys()) i_s = max(ids + 1, its);
    if (g.oye()) i_e = min(ide - 1, ite);
    if (g.periodic_x) { i_s = its; i_e = ite; }
    if (FAST || (i >= i_s && i <= i_e)) {
      float vf[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int kk = k + m;
        if (kk == 1 || kk == g.kde) vf[m] = 0.f;
        else if (kk == 2 || kk == ktf)
          vf[m] = 0.5f * (A3(rom, i, kk, j) + A3(rom, i - 1, kk, j)) * (g.fnm[kk] * A3(u, i, kk, j) + g.fnp[kk] * A3(u, i, kk - 1, j));
        else {
          const float vel = 0.5f * (A3(rom, i - 1, kk, j) + A3(rom, i, kk, j));
          if (WENO && kk >= 4 && kk <= ktf - 2)   // vert_order 5 structure of advect_weno_u
            vf[m] = weno5(vel, -vel * sg >= 0.0f, A3(u, i, kk - 3, j), A3(u, i, kk - 2, j), A3(u, i, kk - 1, j), A3(u, i, kk, j), A3(u, i, kk + 1, j), A3(u, i, kk + 2, j), 1.0e-18);
          else
            vf[m] = vel * flux3(A3(u, i, kk - 2, j), A3(u, i, kk - 1, j), A3(u, i, kk, j), A3(u, i, kk + 1, j), -vel, sgv);
        }
      }
      t = t - g.rdnw[k] * (vf[1] - vf[0]);
    }
  }
  }
  if (PART == 1) return t;
  // --- horizontal pressure gradient, x (:2320-2400)
  {
    int i_s = its, i_e = ite;
    if (g.oxs() && its == ids) i_s = its + 1;
    if (g.oxe() && ite == ide) i_e = ite - 1;
    if (g.periodic_x) { i_s = its; i_e = ite; }
    if (FAST || (i >= i_s && i <= i_e)) {
      const float mr = A2(g.msfux, i, j) / A2(g.msfuy, i, j);
      float dpx = mr * .5f * g.rdx * (g.c1h[k] * A2(s.muu, i, j) + g.c2h[k]) *
                  ((A3(s.ph, i, k + 1, j) - A3(s.ph, i - 1, k + 1, j) + A3(s.ph, i, k, j) - A3(s.ph, i - 1, k, j)) +
                   (A3(s.alt, i, k, j) + A3(s.alt, i - 1, k, j)) * (A3(s.p, i, k, j) - A3(s.p, i - 1, k, j)) +
                   (A3(s.al, i, k, j) + A3(s.al, i - 1, k, j)) * (A3(s.pb, i, k, j) - A3(s.pb, i - 1, k, j)));
      const float d1 = dpn_at(g, s.p, i, k + 1, j, 1, 0), d0 = dpn_at(g, s.p, i, k, j, 1, 0);
      dpx = dpx + mr * g.rdx * (A3(s.php, i, k, j) - A3(s.php, i - 1, k, j)) *
                      (g.rdnw[k] * (d1 - d0) - .5f * ((g.c1h[k] * A2(s.mu, i - 1, j)) + (g.c1h[k] * A2(s.mu, i, j))));
      t = t - A3(s.cqu, i, k, j) * dpx;
    }
  }
  // --- coriolis (:3700-3735) and curvature (:4292-4330)
  {
    int i_s = its, i_e = ite;
    if (g.oxs()) i_s = max(ids + 1, its);
    if (g.oxe()) i_e = min(ide - 1, ite);
    if (g.periodic_x) { i_s = its; i_e = ite; }
    if (FAST || (i >= i_s && i <= i_e)) {
      const float rvavg = 0.25f * (A3(rv, i - 1, k, j + 1) + A3(rv, i, k, j + 1) + A3(rv, i - 1, k, j) + A3(rv, i, k, j));
      const float rwsum = (A3(s.rw, i - 1, k + 1, j) + A3(s.rw, i - 1, k, j) + A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j));
      if (!g.pert_coriolis) {
        t = t + (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) * 0.5f * (A2(g.f, i, j) + A2(g.f, i - 1, j)) * 0.25f *
                    (A3(rv, i - 1, k, j + 1) + A3(rv, i, k, j + 1) + A3(rv, i - 1, k, j) + A3(rv, i, k, j)) -
            0.5f * (A2(g.e, i, j) + A2(g.e, i - 1, j)) * 0.5f * (A2(g.cosa, i, j) + A2(g.cosa, i - 1, j)) * 0.25f * rwsum;
      } else {
        t = t + (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) * 0.5f * (A2(g.f, i, j) + A2(g.f, i - 1, j)) * 0.25f *
                    (PRV(i - 1, k, j + 1) + PRV(i, k, j + 1) + PRV(i - 1, k, j) + PRV(i, k, j)) -
            0.5f * (A2(g.e, i, j) + A2(g.e, i - 1, j)) * 0.5f * (A2(g.cosa, i, j) + A2(g.cosa, i - 1, j)) * 0.25f * rwsum;
      }
      t = t + 0.5f * (vxgm_at<FAST>(g, s.u, s.v, i, k, j) + vxgm_at<FAST>(g, s.u, s.v, i - 1, k, j)) * rvavg -
          A3(u, i, k, j) * WG_RERADIUS * 0.25f * rwsum;
    }
  }
  // perturbation_coriolis keeps the open-boundary columns the plain routine dropped (:3975-4003)
  if (!FAST && g.pert_coriolis && !g.periodic_x) {
    if (!FAST && g.open_xs && i == ids) {
      t = t + (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) * 0.5f * (A2(g.f, i, j) + A2(g.f, i, j)) * 0.25f *
                  (PRV(i, k, j + 1) + PRV(i, k, j + 1) + PRV(i, k, j) + PRV(i, k, j)) -
          0.5f * (A2(g.e, i, j) + A2(g.e, i, j)) * 0.5f * (A2(g.cosa, i, j) + A2(g.cosa, i, j)) * 0.25f *
              (A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j) + A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j));
    }
    if (!FAST && g.open_xe && i == ide) {
      t = t + (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) * 0.5f * (A2(g.f, i - 1, j) + A2(g.f, i - 1, j)) * 0.25f *
                  (PRV(i - 1, k, j + 1) + PRV(i - 1, k, j + 1) + PRV(i - 1, k, j) + PRV(i - 1, k, j)) -
          0.5f * (A2(g.e, i - 1, j) + A2(g.e, i - 1, j)) * 0.5f * (A2(g.cosa, i - 1, j) + A2(g.cosa, i - 1, j)) * 0.25f *
              (A3(s.rw, i - 1, k + 1, j) + A3(s.rw, i - 1, k, j) + A3(s.rw, i - 1, k + 1, j) + A3(s.rw, i - 1, k, j));
    }
  }
  // --- rk_addtend_dry (module_em.F:1037-1043)
  t = t + A3(s.ru_tendf, i, k, j) / A2(g.msfuy, i, j);
  return t;
}
__global__ void k_tend_u(const Geo g, Box b, TendIn s, fp ru_tend) {
  TID3;
  const int jde = g.jde, ktf = g.kde - 1, jte = g.jte;
  if (k > ktf || j > min(jte, jde - 1)) { A3(ru_tend, i, k, j) = 0.f; return; }
  A3(ru_tend, i, k, j) = interior_pt(g, i, j) ? tend_u_pt<true>(g, s, i, k, j) : tend_u_pt<false>(g, s, i, k, j);
}

// momentum_adv_opt = 3 on the last RK substep: advect_weno_u (module_advect_em.F:10876) in place of advect_u
__global__ void k_tend_u_weno(const Geo g, Box b, TendIn s, fp ru_tend) {
  TID3;
  if (k > g.kde - 1 || j > min(g.jte, g.jde - 1)) { A3(ru_tend, i, k, j) = 0.f; return; }
  A3(ru_tend, i, k, j) = tend_u_pt<false, true>(g, s, i, k, j);
}

// ===========================================================================
// rv_tend: advect_v (:1530) + pgf y + coriolis + curvature + addtend
template <bool FAST, bool WENO = false, int PART = 0>
__device__ __forceinline__ float tend_v_pt(const Geo &g, const TendIn &s, int i, int k, int j, float t0 = 0.f) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const Deg dg = degrade(g);
  const float sg = g.sgn_ts;
  // even orders (h 6 / v 4) are the odd ones without the upwind term: its factor is 0 then; WENO keeps sign(time_step)
  const float sgh = WENO ? sg : g.sg_hm, sgv = WENO ? sg : g.sg_vm;
  cfp v = s.v, ru = s.ru, rv = s.rv, rom = s.rom;
  float t = t0;
  if (PART != 2) {
  // --- y flux divergence
  {
    const int j_lo = dg.ys ? max(jts, jds + 1) : jts, j_hi = dg.ye ? min(jte, jde - 1) : jte;
    if (FAST || (j >= j_lo && j <= j_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int jj = j + m;
        const int ord = FAST ? 5 : face_order(jj, dg.ys, dg.ye, jds + 1, jds + 2, jde, jde - 1);
        if (ord == 5) {
          const float vel = 0.5f * (A3(rv, i, k, jj) + A3(rv, i, k, jj - 1));
          fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, YSTENCIL(v, i, k, jj), 1.0e-18) : vel * flux5(YSTENCIL(v, i, k, jj), vel, sgh);
        } else if (ord == 3) {
          const float vel = 0.5f * (A3(rv, i, k, jj) + A3(rv, i, k, jj - 1));
          fq[m] = vel * flux3(A3(v, i, k, jj - 2), A3(v, i, k, jj - 1), A3(v, i, k, jj), A3(v, i, k, jj + 1), vel, sgh);
        } else {
          float va = A3(v, i, k, jj), vb = A3(v, i, k, jj - 1);
          if (g.spec) {   // specified uses upstream normal wind at boundaries (:1985, :2007)
            if (jj == jds + 1) { if (va < 0.f) vb = va; }
            else if (vb > 0.f) va = vb;
          }
          fq[m] = 0.25f * (A3(rv, i, k, jj) + A3(rv, i, k, jj - 1)) * (va + vb);
        }
      }
      const float mrdy = A2(g.msfvy, i, j) * g.rdy;
      t = t - mrdy * (fq[1] - fq[0]);
    }
  }
  // --- x flux divergence
  {
    int j_s = jts, j_e = jte;
    if (g.oys()) j_s = max(jds + 1, jts);
    if (g.oye()) j_e = min(jde - 1, jte);
    const int i_lo = dg.xs ? max(ids + 1, its) : its, i_hi = dg.xe ? min(ide - 2, ite) : min(ite, ide - 1);
    if (FAST || (j >= j_s && j <= j_e && i >= i_lo && i <= i_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int ii = i + m;
        const int ord = FAST ? 5 : face_order(ii, dg.xs, dg.xe, ids + 1, ids + 2, ide - 1, ide - 2);
        if (ord == 5) {
          const float vel = 0.5f * (A3(ru, ii, k, j) + A3(ru, ii, k, j - 1));
          fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, XSTENCIL(v, ii, k, j), 1.0e-18) : vel * flux5(XSTENCIL(v, ii, k, j), vel, sgh);
        } else if (ord == 3) {
          const float vel = 0.5f * (A3(ru, ii, k, j) + A3(ru, ii, k, j - 1));
          fq[m] = vel * flux3(A3(v, ii - 2, k, j), A3(v, ii - 1, k, j), A3(v, ii, k, j), A3(v, ii + 1, k, j), vel, sgh);
        } else {
          fq[m] = 0.25f * (A3(ru, ii, k, j) + A3(ru, ii, k, j - 1)) * (A3(v, ii, k, j) + A3(v, ii - 1, k, j));
        }
      }
      const float mrdx = A2(g.msfvy, i, j) * g.rdx;
      t = t - mrdx * (fq[1] - fq[0]);
    }
  }
  // --- open boundary terms (:2712-2800)
  if (!FAST && g.open_ys && j == jds) {
    const float vb = fminf(A3(rv, i, k, jds) - WG_CB * (g.c1h[k] * A2(s.mut, i, jds) + g.c2h[k]), 0.f);
    t = t - g.rdy * vb * (A3(v, i, k, jds + 1) - A3(v, i, k, jds));
  }
  if (!FAST && g.open_ye && j == jde) {
    const float vb = fmaxf(A3(rv, i, k, jde) + WG_CB * (g.c1h[k] * A2(s.mut, i, jde - 1) + g.c2h[k]), 0.f);
    t = t - g.rdy * vb * (A3(v, i, k, jde) - A3(v, i, k, jde - 1));
  }
  {
    int j_s = jts, j_e = min(jte, jde);
    if (g.open_ys) j_s = max(jds + 1, jts);
    if (g.open_ye) j_e = min(jte, jde - 1);
    const int jmin = jds, jmax = jde - 1;
    if (FAST || (j >= j_s && j <= j_e)) {
      const int jp = min(jmax, j), jm = max(jmin, j - 1);
      if (!FAST && g.open_xs && i == ids) {
        const float mrdx = A2(g.msfvy, ids, j) * g.rdx;
        const float uw = 0.5f * (A3(ru, ids, k, jp) + A3(ru, ids, k, jm));
        const float ub = fminf(uw, 0.f);
        const float dup = A3(ru, ids + 1, k, jp) - A3(ru, ids, k, jp);
        const float dum = A3(ru, ids + 1, k, jm) - A3(ru, ids, k, jm);
        t = t - mrdx * (ub * (A3(v, ids + 1, k, j) - A3(v, ids, k, j)) + 0.5f * A3(v, ids, k, j) * (dup + dum));
      }
      if (!FAST && g.open_xe && i == ide - 1) {
        const float mrdx = A2(g.msfvy, ide - 1, j) * g.rdx;
        const float uw = 0.5f * (A3(ru, ide, k, jp) + A3(ru, ide, k, jm));
        const float ub = fmaxf(uw, 0.f);
        const float dup = A3(ru, ide, k, jp) - A3(ru, ide - 1, k, jp);
        const float dum = A3(ru, ide, k, jm) - A3(ru, ide - 1, k, jm);
        t = t - mrdx * (ub * (A3(v, ide - 1, k, j) - A3(v, ide - 2, k, j)) + 0.5f * A3(v, ide - 1, k, j) * (dup + dum));
      }
    }
  }
  // --- vertical (:2802-2995)
  {
    int j_s = jts, j_e = jte;
    if (g.oys()) j_s = max(jds + 1, jts);
    if (g.oye()) j_e = min(jde - 1, jte);
    if (FAST || (j >= j_s && j <= j_e)) {
      float vf[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int kk = k + m;
        if (kk == 1 || kk == g.kde) vf[m] = 0.f;
        else if (kk == 2 || kk == ktf)
          vf[m] = 0.5f * (A3(rom, i, kk, j) + A3(rom, i, kk, j - 1)) * (g.fnm[kk] * A3(v, i, kk, j) + g.fnp[kk] * A3(v, i, kk - 1, j));
        else {
          const float vel = 0.5f * (A3(rom, i, kk, j) + A3(rom, i, kk, j - 1));
          if (WENO && kk >= 4 && kk <= ktf - 2)   // vert_order 5 structure of advect_weno_v
            vf[m] = weno5(vel, -vel * sg >= 0.0f, A3(v, i, kk - 3, j), A3(v, i, kk - 2, j), A3(v, i, kk - 1, j), A3(v, i, kk, j), A3(v, i, kk + 1, j), A3(v, i, kk + 2, j), 1.0e-18);
          else
            vf[m] = vel * flux3(A3(v, i, kk - 2, j), A3(v, i, kk - 1, j), A3(v, i, kk, j), A3(v, i, kk + 1, j), -vel, sgv);
        }
      }
      t = t - (A2(g.msfvy, i, j) / A2(g.msfvx, i, j)) * g.rdnw[k] * (vf[1] - vf[0]);
    }
  }
  }
  if (PART == 1) return t;
  // --- horizontal pressure gradient, y (:2230-2300)
  {
    int j_s = jts, j_e = jte;
    if (g.oys() && jts == jds) j_s = jts + 1;
    if (g.oye() && jte == jde) j_e = jte - 1;
    if (FAST || (j >= j_s && j <= j_e)) {
      const float mr = A2(g.msfvy, i, j) / A2(g.msfvx, i, j);
      float dpy = mr * .5f * g.rdy * (g.c1h[k] * A2(s.muv, i, j) + g.c2h[k]) *
                  ((A3(s.ph, i, k + 1, j) - A3(s.ph, i, k + 1, j - 1) + A3(s.ph, i, k, j) - A3(s.ph, i, k, j - 1)) +
                   (A3(s.alt, i, k, j) + A3(s.alt, i, k, j - 1)) * (A3(s.p, i, k, j) - A3(s.p, i, k, j - 1)) +
                   (A3(s.al, i, k, j) + A3(s.al, i, k, j - 1)) * (A3(s.pb, i, k, j) - A3(s.pb, i, k, j - 1)));
      const float d1 = dpn_at(g, s.p, i, k + 1, j, 0, 1), d0 = dpn_at(g, s.p, i, k, j, 0, 1);
      dpy = dpy + mr * g.rdy * (A3(s.php, i, k, j) - A3(s.php, i, k, j - 1)) *
                      (g.rdnw[k] * (d1 - d0) - .5f * ((g.c1h[k] * A2(s.mu, i, j - 1)) + (g.c1h[k] * A2(s.mu, i, j))));
      t = t - A3(s.cqv, i, k, j) * dpy;
    }
  }
  // --- coriolis (:3745-3765), curvature (:4340-4380)
  {
    int j_s = jts, j_e = jte;
    if (g.oys()) j_s = max(jds + 1, jts);
    if (g.oye()) j_e = min(jde - 1, jte);
    if (FAST || (j >= j_s && j <= j_e)) {
      const float rwsum = (A3(s.rw, i, k + 1, j - 1) + A3(s.rw, i, k, j - 1) + A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j));
      const float ruavg = 0.25f * (A3(ru, i, k, j) + A3(ru, i + 1, k, j) + A3(ru, i, k, j - 1) + A3(ru, i + 1, k, j - 1));
      const float mr = A2(g.msfvy, i, j) / A2(g.msfvx, i, j);
      if (!g.pert_coriolis) {
        t = t - mr * 0.5f * (A2(g.f, i, j) + A2(g.f, i, j - 1)) * 0.25f *
                    (A3(ru, i, k, j) + A3(ru, i + 1, k, j) + A3(ru, i, k, j - 1) + A3(ru, i + 1, k, j - 1)) +
            mr * 0.5f * (A2(g.e, i, j) + A2(g.e, i, j - 1)) * 0.5f * (A2(g.sina, i, j) + A2(g.sina, i, j - 1)) * 0.25f * rwsum;
      } else {
        t = t - mr * 0.5f * (A2(g.f, i, j) + A2(g.f, i, j - 1)) * 0.25f *
                    (PRU(i, k, j) + PRU(i + 1, k, j) + PRU(i, k, j - 1) + PRU(i + 1, k, j - 1)) +
            mr * 0.5f * (A2(g.e, i, j) + A2(g.e, i, j - 1)) * 0.5f * (A2(g.sina, i, j) + A2(g.sina, i, j - 1)) * 0.25f * rwsum;
      }
      t = t - 0.5f * (vxgm_at<FAST>(g, s.u, s.v, i, k, j) + vxgm_at<FAST>(g, s.u, s.v, i, k, j - 1)) * ruavg -
          mr * A3(v, i, k, j) * WG_RERADIUS * 0.25f * rwsum;
    }
  }
  if (!FAST && g.pert_coriolis) {   // open-boundary rows of perturbation_coriolis (:4066-4085, :4104-4123)
    const float mr = A2(g.msfvy, i, j) / A2(g.msfvx, i, j);
    if (!FAST && g.open_ys && j == jds) {
      t = t - mr * 0.5f * (A2(g.f, i, j) + A2(g.f, i, j)) * 0.25f * (PRU(i, k, j) + PRU(i + 1, k, j) + PRU(i, k, j) + PRU(i + 1, k, j)) +
          mr * 0.5f * (A2(g.e, i, j) + A2(g.e, i, j)) * 0.5f * (A2(g.sina, i, j) + A2(g.sina, i, j)) * 0.25f *
              (A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j) + A3(s.rw, i, k + 1, j) + A3(s.rw, i, k, j));
    }
    if (!FAST && g.open_ye && j == jde) {
      t = t - mr * 0.5f * (A2(g.f, i, j - 1) + A2(g.f, i, j - 1)) * 0.25f *
                  (PRU(i, k, j - 1) + PRU(i + 1, k, j - 1) + PRU(i, k, j - 1) + PRU(i + 1, k, j - 1)) +
          mr * 0.5f * (A2(g.e, i, j - 1) + A2(g.e, i, j - 1)) * 0.5f * (A2(g.sina, i, j - 1) + A2(g.sina, i, j - 1)) * 0.25f *
              (A3(s.rw, i, k + 1, j - 1) + A3(s.rw, i, k, j - 1) + A3(s.rw, i, k + 1, j - 1) + A3(s.rw, i, k, j - 1));
    }
  }
  t = t + A3(s.rv_tendf, i, k, j) * A2(g.msfvx_inv, i, j);
  return t;
}
__global__ void k_tend_v(const Geo g, Box b, TendIn s, fp rv_tend) {
  TID3;
  const int ide = g.ide, ktf = g.kde - 1, ite = g.ite;
  if (k > ktf || i > min(ite, ide - 1)) { A3(rv_tend, i, k, j) = 0.f; return; }
  A3(rv_tend, i, k, j) = interior_pt(g, i, j) ? tend_v_pt<true>(g, s, i, k, j) : tend_v_pt<false>(g, s, i, k, j);
}

__global__ void k_tend_v_weno(const Geo g, Box b, TendIn s, fp rv_tend) {   // advect_weno_v (:11523)
  TID3;
  if (k > g.kde - 1 || i > min(g.ite, g.ide - 1)) { A3(rv_tend, i, k, j) = 0.f; return; }
  A3(rv_tend, i, k, j) = tend_v_pt<false, true>(g, s, i, k, j);
}

// ===========================================================================
// rw_tend: advect_w (:4364) + pg_buoy_w (:2419) + w_damp (:2503) + coriolis + curvature + addtend
// cqw is read here un-inverted; k_invert_cqw applies cqw <- 1/(1+cqw) (:2489) afterwards.
template <bool FAST, bool WENO = false, int PART = 0>
__device__ __forceinline__ float tend_w_pt(const Geo &g, const TendIn &s, int i, int k, int j, float t0 = 0.f) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, kde = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const Deg dg = degrade(g);
  const float sg = g.sgn_ts;
  // even orders (h 6 / v 4) are the odd ones without the upwind term: its factor is 0 then; WENO keeps sign(time_step)
  // sic: advect_w (:4463-4464) and advect_weno_w (:12339-12340) read h_sca_adv_order / v_sca_adv_order, not the momentum orders
  const float sgh = WENO ? sg : g.sg_hs, sgv = WENO ? sg : g.sg_vs;
  cfp w = s.w, ru = s.ru, rv = s.rv, rom = s.rom;
  float t = t0;
  if (k >= 2) {
    const bool top = (k == kde);
    const float fa = top ? (2.f - g.fnm[k - 1]) : g.fnm[k];
    const float fb = top ? g.fnp[k - 1] : g.fnp[k];
    const int ka = top ? k - 1 : k, kb = top ? k - 2 : k - 1;
    // vel at a face: interior  fzm(k)*r(k)+fzp(k)*r(k-1);  top  (2-fzm(k-1))*r(k-1)-fzp(k-1)*r(k-2)
#define WVEL(r, ii, jj) (top ? (fa * A3(r, ii, ka, jj) - fb * A3(r, ii, kb, jj)) : (fa * A3(r, ii, ka, jj) + fb * A3(r, ii, kb, jj)))
    if (PART != 2) {
    // --- y
    {
      const int j_lo = dg.ys ? max(jts, jds + 1) : jts, j_hi = dg.ye ? min(jte, jde - 2) : min(jte, jde - 1);
      if (FAST || (j >= j_lo && j <= j_hi)) {
        float fq[2];
#pragma unroll
        for (int m = 0; m < 2; m++) {
          const int jj = j + m;
          const int ord = FAST ? 5 : face_order(jj, dg.ys, dg.ye, jds + 1, jds + 2, jde - 1, jde - 2);
          const float vel = WVEL(rv, i, jj);
          if (ord == 5) fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, YSTENCIL(w, i, k, jj), 1.0e-18) : vel * flux5(YSTENCIL(w, i, k, jj), vel, sgh);
          else if (ord == 3) fq[m] = vel * flux3(A3(w, i, k, jj - 2), A3(w, i, k, jj - 1), A3(w, i, k, jj), A3(w, i, k, jj + 1), vel, sgh);
          else fq[m] = 0.5f * (vel) * (A3(w, i, k, jj) + A3(w, i, k, jj - 1));
        }
        const float mrdy = A2(g.msftx, i, j) * g.rdy;
        t = t - mrdy * (fq[1] - fq[0]);
      }
    }
    // --- x
    {
      const int i_lo = dg.xs ? max(ids + 1, its) : its, i_hi = dg.xe ? min(ide - 2, ite) : min(ite, ide - 1);
      if (FAST || (i >= i_lo && i <= i_hi)) {
        float fq[2];
#pragma unroll
        for (int m = 0; m < 2; m++) {
          const int ii = i + m;
          const int ord = FAST ? 5 : face_order(ii, dg.xs, dg.xe, ids + 1, ids + 2, ide - 1, ide - 2);
          const float vel = WVEL(ru, ii, j);
          if (ord == 5) fq[m] = WENO ? weno5(vel, vel * sg >= 0.0f, XSTENCIL(w, ii, k, j), 1.0e-18) : vel * flux5(XSTENCIL(w, ii, k, j), vel, sgh);
          else if (ord == 3) fq[m] = vel * flux3(A3(w, ii - 2, k, j), A3(w, ii - 1, k, j), A3(w, ii, k, j), A3(w, ii + 1, k, j), vel, sgh);
          else fq[m] = 0.5f * (vel) * (A3(w, ii, k, j) + A3(w, ii - 1, k, j));
        }
        const float mrdx = A2(g.msftx, i, j) * g.rdx;
        t = t - mrdx * (fq[1] - fq[0]);
      }
    }
    // --- open boundary terms (:5689-5820)
    // uw = 0.5*(fa*(r(a,ka)+r(b,ka)) +/- fb*(r(a,kb)+r(b,kb))), divergence term fa*(dr(ka)) +/- fb*(dr(kb))
#define WSUM(r, i1, j1, i2, j2) (top ? (fa * (A3(r, i1, ka, j1) + A3(r, i2, ka, j2)) - fb * (A3(r, i1, kb, j1) + A3(r, i2, kb, j2))) \
                                     : (fa * (A3(r, i1, ka, j1) + A3(r, i2, ka, j2)) + fb * (A3(r, i1, kb, j1) + A3(r, i2, kb, j2))))
#define WDIF(r, i1, j1, i2, j2) (top ? (fa * (A3(r, i1, ka, j1) - A3(r, i2, ka, j2)) - fb * (A3(r, i1, kb, j1) - A3(r, i2, kb, j2))) \
                                     : (fa * (A3(r, i1, ka, j1) - A3(r, i2, ka, j2)) + fb * (A3(r, i1, kb, j1) - A3(r, i2, kb, j2))))
    const int i_end = min(ite, ide - 1), j_end = min(jte, jde - 1);
    if (!FAST && g.open_xs && i == ids) {
      const float uw = 0.5f * WSUM(ru, ids, j, ids + 1, j);
      const float ub = fminf(uw, 0.f);
      t = t - g.rdx * (ub * (A3(w, ids + 1, k, j) - A3(w, ids, k, j)) + A3(w, ids, k, j) * WDIF(ru, ids + 1, j, ids, j));
    }
    if (!FAST && g.open_xe && ite == ide && i == i_end) {
      const float uw = 0.5f * WSUM(ru, ide - 1, j, ide, j);
      const float ub = fmaxf(uw, 0.f);
      t = t - g.rdx * (ub * (A3(w, i_end, k, j) - A3(w, i_end - 1, k, j)) + A3(w, i_end, k, j) * WDIF(ru, ide, j, ide - 1, j));
    }
    if (!FAST && g.open_ys && j == jds) {
      const float vw = 0.5f * WSUM(rv, i, jds, i, jds + 1);
      const float vb = fminf(vw, 0.f);
      t = t - g.rdy * (vb * (A3(w, i, k, jds + 1) - A3(w, i, k, jds)) + A3(w, i, k, jds) * WDIF(rv, i, jds + 1, i, jds));
    }
    if (!FAST && g.open_ye && jte == jde && j == j_end) {
      const float vw = 0.5f * WSUM(rv, i, jde - 1, i, jde);
      const float vb = fmaxf(vw, 0.f);
      t = t - g.rdy * (vb * (A3(w, i, k, j_end) - A3(w, i, k, j_end - 1)) + A3(w, i, k, j_end) * WDIF(rv, i, jde, i, jde - 1));
    }
#undef WVEL
#undef WSUM
#undef WDIF
    // --- vertical (:5996-6031), rdzu = rdn
    {
      auto vflux = [&](int kk) -> float {
        if (kk == 2 || kk == kde) return 0.25f * (A3(rom, i, kk, j) + A3(rom, i, kk - 1, j)) * (A3(w, i, kk, j) + A3(w, i, kk - 1, j));
        const float vel = 0.5f * (A3(rom, i, kk, j) + A3(rom, i, kk - 1, j));
        if (WENO && kk >= 4 && kk <= kde - 2)   // vert_order 5 structure of advect_weno_w (:12959-13037)
          return weno5(vel, -vel * sg >= 0.0f, A3(w, i, kk - 3, j), A3(w, i, kk - 2, j), A3(w, i, kk - 1, j), A3(w, i, kk, j), A3(w, i, kk + 1, j), A3(w, i, kk + 2, j), 1.0e-18);
        return vel * flux3(A3(w, i, kk - 2, j), A3(w, i, kk - 1, j), A3(w, i, kk, j), A3(w, i, kk + 1, j), -vel, sgv);
      };
      if (!top) t = t - g.rdn[k] * (vflux(k + 1) - vflux(k));
      else t = t + 2.f * g.rdn[k - 1] * (vflux(k));
    }
    }
    if (PART == 1) return t;
    // --- pg_buoy_w (:2419-2499)
    {
      const float mu = A2(s.mu, i, j), mub = A2(s.mub, i, j), rmsf = (1.f / A2(g.msfty, i, j));
      if (top) {
        const float cq1 = 1.f / (1.f + A3(s.cqw, i, k - 1, j));
        const float cq2 = A3(s.cqw, i, k - 1, j) * cq1;
        t = t + rmsf * WG_G * (cq1 * 2.f * g.rdnw[k - 1] * (-A3(s.p, i, k - 1, j)) - (g.c1f[k] * mu) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
      } else {
        const float cq1 = 1.f / (1.f + A3(s.cqw, i, k, j));
        const float cq2 = A3(s.cqw, i, k, j) * cq1;
        t = t + rmsf * WG_G * (cq1 * g.rdn[k] * (A3(s.p, i, k, j) - A3(s.p, i, k - 1, j)) - (g.c1f[k] * mu) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
      }
    }
    if (!top) {
      // --- w_damp (:2635-2695); the CFL maxima are only consumed by adaptive dt (solve_em.F:914-930)
      if (g.w_damping == 1) {
        const float mass = (g.c1f[k] * A2(s.mut, i, j) + g.c2f[k]);
        const float vert_cfl = fabsf(A3(s.ww, i, k, j) / mass * g.rdnw[k] * g.dt);
        if (vert_cfl > g.w_damp_on) t = t - sign1(A3(w, i, k, j)) * 0.3f * (vert_cfl - g.w_crit_cfl) * mass;
      }
      // --- coriolis (:3839-3844) and curvature (:4459-4463)
      const float fzm = g.fnm[k], fzp = g.fnp[k];
      const float mr = A2(g.msftx, i, j) / A2(g.msfty, i, j);
      if (!g.pert_coriolis)
        t = t + A2(g.e, i, j) * (A2(g.cosa, i, j) * 0.5f * (fzm * (A3(ru, i, k, j) + A3(ru, i + 1, k, j)) + fzp * (A3(ru, i, k - 1, j) + A3(ru, i + 1, k - 1, j))) -
                                 mr * A2(g.sina, i, j) * 0.5f * (fzm * (A3(rv, i, k, j) + A3(rv, i, k, j + 1)) + fzp * (A3(rv, i, k - 1, j) + A3(rv, i, k - 1, j + 1))));
      else
        t = t + A2(g.e, i, j) * (A2(g.cosa, i, j) * 0.5f * (fzm * (PRU(i, k, j) + PRU(i + 1, k, j)) + fzp * (PRU(i, k - 1, j) + PRU(i + 1, k - 1, j))) -
                                 mr * A2(g.sina, i, j) * 0.5f * (fzm * (PRV(i, k, j) + PRV(i, k, j + 1)) + fzp * (PRV(i, k - 1, j) + PRV(i, k - 1, j + 1))));
      t = t + WG_RERADIUS * (0.5f * (fzm * (A3(ru, i, k, j) + A3(ru, i + 1, k, j)) + fzp * (A3(ru, i, k - 1, j) + A3(ru, i + 1, k - 1, j))) * 0.5f *
                                 (fzm * (A3(s.u, i, k, j) + A3(s.u, i + 1, k, j)) + fzp * (A3(s.u, i, k - 1, j) + A3(s.u, i + 1, k - 1, j))) +
                             mr * 0.5f * (fzm * (A3(rv, i, k, j) + A3(rv, i, k, j + 1)) + fzp * (A3(rv, i, k - 1, j) + A3(rv, i, k - 1, j + 1))) * 0.5f *
                                 (fzm * (A3(s.v, i, k, j) + A3(s.v, i, k, j + 1)) + fzp * (A3(s.v, i, k - 1, j) + A3(s.v, i, k - 1, j + 1))));
    }
  }
  if (PART == 1) return t;
  t = t + A3(s.rw_tendf, i, k, j) / A2(g.msfty, i, j);
  return t;
}
__global__ void k_tend_w(const Geo g, Box b, TendIn s, fp rw_tend) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1, kde = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  (void)ids; (void)jds; (void)ktf; (void)kde; (void)its; (void)jts;
  if (i > min(ite, ide - 1) || j > min(jte, jde - 1)) { A3(rw_tend, i, k, j) = 0.f; return; }
  A3(rw_tend, i, k, j) = interior_pt(g, i, j) ? tend_w_pt<true>(g, s, i, k, j) : tend_w_pt<false>(g, s, i, k, j);
}

__global__ void k_tend_w_weno(const Geo g, Box b, TendIn s, fp rw_tend) {   // advect_weno_w (:12218)
  TID3;
  if (i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { A3(rw_tend, i, k, j) = 0.f; return; }
  A3(rw_tend, i, k, j) = tend_w_pt<false, true>(g, s, i, k, j);
}

// cqw <- 1/(1+cqw), module_big_step_utilities_em.F:2489
__global__ void k_invert_cqw(const Geo g, Box b, fp cqw) {
  TID3;
  A3(cqw, i, k, j) = 1.f / (1.f + A3(cqw, i, k, j));
}

// ===========================================================================
// advect_scalar (:3029) evaluated at one point; shared by t_tend and the scalar path
template <bool FAST>
__device__ __forceinline__ float advect_scalar_pt(const Geo &g, cfp q, cfp q_old, cfp ru, cfp rv, cfp rom, int i, int k, int j) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const Deg dg = degrade(g);
  const float sgh = g.sg_hs, sgv = g.sg_vs;   // 0 for the even orders (h 6 / v 4): no upwind term
  float t = 0.f;
  {
    const int j_lo = dg.ys ? max(jts, jds + 1) : jts, j_hi = dg.ye ? min(jte, jde - 2) : min(jte, jde - 1);
    if (FAST || (j >= j_lo && j <= j_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int jj = j + m;
        const int ord = FAST ? 5 : face_order(jj, dg.ys, dg.ye, jds + 1, jds + 2, jde - 1, jde - 2);
        const float vel = A3(rv, i, k, jj);
        if (ord == 5) fq[m] = vel * flux5(YSTENCIL(q, i, k, jj), vel, sgh);
        else if (ord == 3) fq[m] = vel * flux3(A3(q, i, k, jj - 2), A3(q, i, k, jj - 1), A3(q, i, k, jj), A3(q, i, k, jj + 1), vel, sgh);
        else fq[m] = 0.5f * vel * (A3(q, i, k, jj) + A3(q, i, k, jj - 1));
      }
      const float mrdy = A2(g.msftx, i, j) * g.rdy;
      t = t - mrdy * (fq[1] - fq[0]);
    }
  }
  {
    const int i_lo = dg.xs ? max(ids + 1, its) : its, i_hi = dg.xe ? min(ide - 2, ite) : min(ite, ide - 1);
    if (FAST || (i >= i_lo && i <= i_hi)) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int ii = i + m;
        const int ord = FAST ? 5 : face_order(ii, dg.xs, dg.xe, ids + 1, ids + 2, ide - 1, ide - 2);
        const float vel = A3(ru, ii, k, j);
        if (ord == 5) fq[m] = vel * flux5(XSTENCIL(q, ii, k, j), vel, sgh);
        else if (ord == 3) fq[m] = vel * flux3(A3(q, ii - 2, k, j), A3(q, ii - 1, k, j), A3(q, ii, k, j), A3(q, ii + 1, k, j), vel, sgh);
        else fq[m] = 0.5f * (vel) * (A3(q, ii, k, j) + A3(q, ii - 1, k, j));
      }
      const float mrdx = A2(g.msftx, i, j) * g.rdx;
      t = t - mrdx * (fq[1] - fq[0]);
    }
  }
  const int i_end = min(ite, ide - 1), j_end = min(jte, jde - 1);
  if (!FAST && g.open_xs && i == ids) {
    const float ub = fminf(0.5f * (A3(ru, ids, k, j) + A3(ru, ids + 1, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, ids + 1, k, j) - A3(q_old, ids, k, j)) + A3(q, ids, k, j) * (A3(ru, ids + 1, k, j) - A3(ru, ids, k, j)));
  }
  if (!FAST && g.open_xe && ite == ide && i == i_end) {
    const float ub = fmaxf(0.5f * (A3(ru, ide - 1, k, j) + A3(ru, ide, k, j)), 0.f);
    t = t - g.rdx * (ub * (A3(q_old, i_end, k, j) - A3(q_old, i_end - 1, k, j)) + A3(q, i_end, k, j) * (A3(ru, ide, k, j) - A3(ru, ide - 1, k, j)));
  }
  if (!FAST && g.open_ys && j == jds) {
    const float vb = fminf(0.5f * (A3(rv, i, k, jds) + A3(rv, i, k, jds + 1)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, jds + 1) - A3(q_old, i, k, jds)) + A3(q, i, k, jds) * (A3(rv, i, k, jds + 1) - A3(rv, i, k, jds)));
  }
  if (!FAST && g.open_ye && jte == jde && j == j_end) {
    const float vb = fmaxf(0.5f * (A3(rv, i, k, jde - 1) + A3(rv, i, k, jde)), 0.f);
    t = t - g.rdy * (vb * (A3(q_old, i, k, j_end) - A3(q_old, i, k, j_end - 1)) + A3(q, i, k, j_end) * (A3(rv, i, k, jde) - A3(rv, i, k, jde - 1)));
  }
  {
    float vf[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const int kk = k + m;
      if (kk == 1 || kk == g.kde) vf[m] = 0.f;
      else if (kk == 2 || kk == ktf) vf[m] = A3(rom, i, kk, j) * (g.fnm[kk] * A3(q, i, kk, j) + g.fnp[kk] * A3(q, i, kk - 1, j));
      else {
        const float vel = A3(rom, i, kk, j);
        vf[m] = vel * flux3(A3(q, i, kk - 2, j), A3(q, i, kk - 1, j), A3(q, i, kk, j), A3(q, i, kk + 1, j), -vel, sgv);
      }
    }
    t = t - g.rdnw[k] * (vf[1] - vf[0]);
  }
  return t;
}

// t_tend: advect_scalar(theta) + rk_addtend_dry (module_em.F:1068-1078)
__global__ void k_tend_t(const Geo g, Box b, TendIn s, fp t_tend) {
  TID3;
  if (k > g.kde - 1 || i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { A3(t_tend, i, k, j) = 0.f; return; }
  float t = interior_pt(g, i, j) ? advect_scalar_pt<true>(g, s.t, s.t, s.ru, s.rv, s.rom, i, k, j)
                               : advect_scalar_pt<false>(g, s.t, s.t, s.ru, s.rv, s.rom, i, k, j);
  const float msfty = A2(g.msfty, i, j);
  t = t + A3(s.t_tendf, i, k, j) / msfty + (g.c1h[k] * A2(s.mut, i, j) + g.c2h[k]) * A3(s.h_diabatic, i, k, j) / msfty;
  A3(t_tend, i, k, j) = t;
}

// t_tend on the last RK substep when a scalar family uses WENO (module_em.F:603-621): advect_scalar_weno(theta)
__global__ void k_tend_t_weno(const Geo g, Box b, TendIn s, fp t_tend) {
  TID3;
  if (k > g.kde - 1 || i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { A3(t_tend, i, k, j) = 0.f; return; }
  float t = advect_scalar_weno_pt(g, s.t, s.t, s.ru, s.rv, s.rom, i, k, j);
  const float msfty = A2(g.msfty, i, j);
  t = t + A3(s.t_tendf, i, k, j) / msfty + (g.c1h[k] * A2(s.mut, i, j) + g.c2h[k]) * A3(s.h_diabatic, i, k, j) / msfty;
  A3(t_tend, i, k, j) = t;
}

// external entry for the scalar path (scalar.cu)
__global__ void k_advect_scalar(const Geo g, Box b, cfp q, cfp ru, cfp rv, cfp rom, fp tend) {
  TID3;
  A3(tend, i, k, j) = interior_pt(g, i, j) ? advect_scalar_pt<true>(g, q, q, ru, rv, rom, i, k, j)
                                           : advect_scalar_pt<false>(g, q, q, ru, rv, rom, i, k, j);
}

// ===========================================================================
// ph_tend: rhs_ph (:1365-2181, 6th-order branch) + rk_addtend_dry
template <bool FAST, int PART = 0>
__device__ __forceinline__ float tend_ph_pt(const Geo &g, const TendIn &s, int i, int k, int j, float t0 = 0.f) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, kde = g.kde, kte = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  cfp ph = s.ph, phb = s.phb, u = s.u, v = s.v, ww = s.rom;
  const float msfty = A2(g.msfty, i, j);
  float t = t0;
  if (k >= 2) {
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    if (k <= kte - 1) {
      auto wdwn = [&](int kk) -> float {
        if (g.phi_adv_z == 2) return g.rdnw[kk - 1] * (A3(ph, i, kk, j) - A3(ph, i, kk - 1, j) + A3(phb, i, kk, j) - A3(phb, i, kk - 1, j));
        return .5f * (A3(ww, i, kk, j) + A3(ww, i, kk - 1, j)) * g.rdnw[kk - 1] *
               (A3(ph, i, kk, j) - A3(ph, i, kk - 1, j) + A3(phb, i, kk, j) - A3(phb, i, kk - 1, j));
      };
      if (g.phi_adv_z == 2) t = t - A3(ww, i, k, j) * (g.fnm[k] * wdwn(k + 1) + g.fnp[k] * wdwn(k));
      else t = t - (g.fnm[k] * wdwn(k + 1) + g.fnp[k] * wdwn(k));
    }
    // "gw" term (:1496-1508); ph_tend(:,kde,:) is reset to 0 before it
    t = t + (c1f * A2(s.mut, i, j) + c2f) * WG_G * A3(s.w, i, k, j) / msfty;
    const bool top = (k == kte);
    const float fac = top ? 0.5f : 0.25f;
    // --- y advection
    {
      int j6s = jts, j6e = min(jte, jde - 1);
      if (g.oys()) j6s = max(jts, jds + 3);
      if (g.oye()) j6e = min(j6e, jde - 4);
      int mode = 0;
      if (FAST || (j >= j6s && j <= j6e)) mode = 6;
      else if (g.oys() && j == jds + 2) mode = 4;
      else if (g.oye() && j == jde - 3) mode = 4;
      else if (g.oys() && j == jds + 1) mode = 2;
      else if (g.oye() && j == jde - 2) mode = 2;
      // a row can match more than one of the reference's blocks only on tiny domains
      if (mode) {
        const float va = top ? (g.cfn * A3(v, i, k - 1, j + 1) + g.cfn1 * A3(v, i, k - 2, j + 1)) : (A3(v, i, k, j + 1) + A3(v, i, k - 1, j + 1));
        const float vb = top ? (g.cfn * A3(v, i, k - 1, j) + g.cfn1 * A3(v, i, k - 2, j)) : (A3(v, i, k, j) + A3(v, i, k - 1, j));
        const float ma = (c1f * A2(s.muv, i, j + 1) + c2f), mb = (c1f * A2(s.muv, i, j) + c2f);
        const float ya = A2(g.msfvy, i, j + 1), yb = A2(g.msfvy, i, j);
        if (mode == 6)
          t = t - (fac * g.rdy / msfty) *
                      ((ma * va * ya + mb * vb * yb) * (1.f / 60.f) *
                       (45.f * (A3(ph, i, k, j + 1) - A3(ph, i, k, j - 1)) - 9.f * (A3(ph, i, k, j + 2) - A3(ph, i, k, j - 2)) +
                        (A3(ph, i, k, j + 3) - A3(ph, i, k, j - 3)) + 45.f * (A3(phb, i, k, j + 1) - A3(phb, i, k, j - 1)) -
                        9.f * (A3(phb, i, k, j + 2) - A3(phb, i, k, j - 2)) + (A3(phb, i, k, j + 3) - A3(phb, i, k, j - 3))));
        else if (mode == 4)
          t = t - (fac * g.rdy / msfty) *
                      ((ma * va * ya + mb * vb * yb) * (1.f / 12.f) *
                       (8.f * (A3(ph, i, k, j + 1) - A3(ph, i, k, j - 1)) - (A3(ph, i, k, j + 2) - A3(ph, i, k, j - 2)) +
                        8.f * (A3(phb, i, k, j + 1) - A3(phb, i, k, j - 1)) - (A3(phb, i, k, j + 2) - A3(phb, i, k, j - 2))));
        else
          t = t - (fac * g.rdy / msfty) *
                      (ma * va * ya * (A3(phb, i, k, j + 1) - A3(phb, i, k, j) + A3(ph, i, k, j + 1) - A3(ph, i, k, j)) +
                       mb * vb * yb * (A3(phb, i, k, j) - A3(phb, i, k, j - 1) + A3(ph, i, k, j) - A3(ph, i, k, j - 1)));
      }
    }
    // --- x advection
    {
      int i6s = its, i6e = min(ite, ide - 1);
      if (g.oxs()) i6s = max(its, ids + 3);
      if (g.oxe()) i6e = min(i6e, ide - 4);
      int mode = 0;
      if (FAST || (i >= i6s && i <= i6e)) mode = 6;
      else if (g.open_xs && i == ids + 2) mode = 4;
      else if (g.open_xe && i == ide - 3) mode = 4;
      else if (g.oxs() && i == ids + 1) mode = 2;
      else if (g.oxe() && i == ide - 2) mode = 2;
      if (mode) {
        const float ua = top ? (g.cfn * A3(u, i + 1, k - 1, j) + g.cfn1 * A3(u, i + 1, k - 2, j)) : (A3(u, i + 1, k, j) + A3(u, i + 1, k - 1, j));
        const float ub = top ? (g.cfn * A3(u, i, k - 1, j) + g.cfn1 * A3(u, i, k - 2, j)) : (A3(u, i, k, j) + A3(u, i, k - 1, j));
        const float ma = (c1f * A2(s.muu, i + 1, j) + c2f), mb = (c1f * A2(s.muu, i, j) + c2f);
        const float xa = A2(g.msfux, i + 1, j), xb = A2(g.msfux, i, j);
        if (mode == 6)
          t = t - (fac * g.rdx / msfty) *
                      ((ma * ua * xa + mb * ub * xb) * (1.f / 60.f) *
                       (45.f * (A3(ph, i + 1, k, j) - A3(ph, i - 1, k, j)) - 9.f * (A3(ph, i + 2, k, j) - A3(ph, i - 2, k, j)) +
                        (A3(ph, i + 3, k, j) - A3(ph, i - 3, k, j)) + 45.f * (A3(phb, i + 1, k, j) - A3(phb, i - 1, k, j)) -
                        9.f * (A3(phb, i + 2, k, j) - A3(phb, i - 2, k, j)) + (A3(phb, i + 3, k, j) - A3(phb, i - 3, k, j))));
        else if (mode == 4)
          t = t - (fac * g.rdx / msfty) *
                      ((ma * ua * xa + mb * ub * xb) * (1.f / 12.f) *
                       (8.f * (A3(ph, i + 1, k, j) - A3(ph, i - 1, k, j)) - (A3(ph, i + 2, k, j) - A3(ph, i - 2, k, j)) +
                        8.f * (A3(phb, i + 1, k, j) - A3(phb, i - 1, k, j)) - (A3(phb, i + 2, k, j) - A3(phb, i - 2, k, j))));
        else
          t = t - (fac * g.rdx / msfty) *
                      (ma * ua * xa * (A3(phb, i + 1, k, j) - A3(phb, i, k, j) + A3(ph, i + 1, k, j) - A3(ph, i, k, j)) +
                       mb * ub * xb * (A3(phb, i, k, j) - A3(phb, i - 1, k, j) + A3(ph, i, k, j) - A3(ph, i - 1, k, j)));
      }
    }
    // --- lateral open boundary terms (:2065-2179); ph_old == ph in rk_tendency
    const float mass = (c1f * A2(s.mut, i, j) + c2f);
    if (!FAST && g.open_ys && j == jds) {
      const int kz = min(k, kde - 1);
      const float vb = .5f * (g.fnm[kz] * (A3(v, i, kz, j + 1) + A3(v, i, kz, j)) + g.fnp[kz] * (A3(v, i, kz - 1, j + 1) + A3(v, i, kz - 1, j)));
      const float vl = fminf(vb, 0.f);
      t = t - g.rdy * mass * (+vl * (A3(ph, i, k, j + 1) - A3(ph, i, k, j)));
    }
    if (!FAST && g.open_ye && jte == jde && j == jde - 1) {
      const int kz = min(k, kde - 1);
      const float vb = .5f * (g.fnm[kz] * (A3(v, i, kz, j + 1) + A3(v, i, kz, j)) + g.fnp[kz] * (A3(v, i, kz - 1, j + 1) + A3(v, i, kz - 1, j)));
      const float vr = fmaxf(vb, 0.f);
      t = t - g.rdy * mass * (+vr * (A3(ph, i, k, j) - A3(ph, i, k, j - 1)));
    }
    if (!FAST && g.open_xs && i == ids) {
      const int kz = k;   // the reference uses kz = k even at k = kde (:2139-2146)
      const float ub = .5f * (g.fnm[kz] * (A3(u, i + 1, kz, j) + A3(u, i, kz, j)) + g.fnp[kz] * (A3(u, i + 1, kz - 1, j) + A3(u, i, kz - 1, j)));
      const float ul = fminf(ub, 0.f);
      t = t - (A2(g.msftx, i, j) / msfty) * g.rdx * mass * (+ul * (A3(ph, i + 1, k, j) - A3(ph, i, k, j)));
    }
    if (!FAST && g.open_xe && ite == ide && i == ide - 1) {
      const int kz = (k == kde) ? k - 1 : k;
      const float ub = .5f * (g.fnm[kz] * (A3(u, i + 1, kz, j) + A3(u, i, kz, j)) + g.fnp[kz] * (A3(u, i + 1, kz - 1, j) + A3(u, i, kz - 1, j)));
      const float ur = fmaxf(ub, 0.f);
      t = t - (A2(g.msftx, i, j) / msfty) * g.rdx * mass * (+ur * (A3(ph, i, k, j) - A3(ph, i - 1, k, j)));
    }
  }
  if (PART == 1) return t;
  t = t + A3(s.ph_tendf, i, k, j) / msfty;
  return t;
}
__global__ void k_tend_ph(const Geo g, Box b, TendIn s, fp ph_tend) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1, kde = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  (void)ids; (void)jds; (void)ktf; (void)kde; (void)its; (void)jts;
  if (i > min(ite, ide - 1) || j > min(jte, jde - 1)) { A3(ph_tend, i, k, j) = 0.f; return; }
  A3(ph_tend, i, k, j) = interior_pt(g, i, j) ? tend_ph_pt<true>(g, s, i, k, j) : tend_ph_pt<false>(g, s, i, k, j);
}

// ===========================================================================
// IEVA (module_ieva_em.F), last RK substep: each tendency is built in three launches -- PART 1 the explicit advection
// (with wwE), then the implicit column solve on the stored tendency (ieva.cu), then PART 2 everything the fused kernels
// add after the advection, starting from the stored value.  Gather kernels on the whole tile.
template <int PART>
__global__ void k_tend_u_part(const Geo g, Box b, TendIn s, fp ru_tend, int weno) {
  TID3;
  if (k > g.kde - 1 || j > min(g.jte, g.jde - 1)) { if (PART == 1) A3(ru_tend, i, k, j) = 0.f; return; }
  const float t0 = PART == 2 ? A3(ru_tend, i, k, j) : 0.f;
  A3(ru_tend, i, k, j) = weno ? tend_u_pt<false, true, PART>(g, s, i, k, j, t0) : tend_u_pt<false, false, PART>(g, s, i, k, j, t0);
}
template <int PART>
__global__ void k_tend_v_part(const Geo g, Box b, TendIn s, fp rv_tend, int weno) {
  TID3;
  if (k > g.kde - 1 || i > min(g.ite, g.ide - 1)) { if (PART == 1) A3(rv_tend, i, k, j) = 0.f; return; }
  const float t0 = PART == 2 ? A3(rv_tend, i, k, j) : 0.f;
  A3(rv_tend, i, k, j) = weno ? tend_v_pt<false, true, PART>(g, s, i, k, j, t0) : tend_v_pt<false, false, PART>(g, s, i, k, j, t0);
}
template <int PART>
__global__ void k_tend_w_part(const Geo g, Box b, TendIn s, fp rw_tend, int weno) {
  TID3;
  if (i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { if (PART == 1) A3(rw_tend, i, k, j) = 0.f; return; }
  const float t0 = PART == 2 ? A3(rw_tend, i, k, j) : 0.f;
  A3(rw_tend, i, k, j) = weno ? tend_w_pt<false, true, PART>(g, s, i, k, j, t0) : tend_w_pt<false, false, PART>(g, s, i, k, j, t0);
}
template <int PART>
__global__ void k_tend_ph_part(const Geo g, Box b, TendIn s, fp ph_tend) {
  TID3;
  if (i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { if (PART == 1) A3(ph_tend, i, k, j) = 0.f; return; }
  if (PART == 1) A3(ph_tend, i, k, j) = tend_ph_pt<false, 1>(g, s, i, k, j);
  else A3(ph_tend, i, k, j) = A3(ph_tend, i, k, j) + A3(s.ph_tendf, i, k, j) / A2(g.msfty, i, j);
}
template <int PART>
__global__ void k_tend_t_part(const Geo g, Box b, TendIn s, fp t_tend, int weno) {
  TID3;
  if (k > g.kde - 1 || i > min(g.ite, g.ide - 1) || j > min(g.jte, g.jde - 1)) { if (PART == 1) A3(t_tend, i, k, j) = 0.f; return; }
  if (PART == 1) {
    A3(t_tend, i, k, j) = weno ? advect_scalar_weno_pt(g, s.t, s.t, s.ru, s.rv, s.rom, i, k, j)
                               : advect_scalar_pt<false>(g, s.t, s.t, s.ru, s.rv, s.rom, i, k, j);
  } else {
    const float msfty = A2(g.msfty, i, j);
    float t = A3(t_tend, i, k, j);
    t = t + A3(s.t_tendf, i, k, j) / msfty + (g.c1h[k] * A2(s.mut, i, j) + g.c2h[k]) * A3(s.h_diabatic, i, k, j) / msfty;
    A3(t_tend, i, k, j) = t;
  }
}

// ===========================================================================
// rk_rayleigh_damp (:5836-6110): one thread per column and per variable kind
// (kind 0: u points, 1: v points, 2: mass points -> w and theta)
__global__ void k_rayleigh(const Geo g, Box b, int kind, cfp u, cfp v, cfp w, cfp t, cfp ph, cfp phb, cfp mut, cfp muu, cfp muv,
                           fp ru_tendf, fp rv_tendf, fp rw_tendf, fp t_tendf) {
  TID2;
  const int kde = g.kde, ktf = g.kde - 1;
  const float zdamp = g.zdamp, dampcoef = g.dampcoef;
  const float pii = 0x1.921fb6p+1f;   // 2.*asin(1.) in real*4
  const int di = kind == 0 ? 1 : 0, dj = kind == 1 ? 1 : 0;
  auto zmid = [&](int k1) -> float {
    if (kind == 2) return 0.5f * (A3(phb, i, k1, j) + A3(phb, i, k1 + 1, j) + A3(ph, i, k1, j) + A3(ph, i, k1 + 1, j)) / WG_G;
    return 0.25f * (A3(phb, i, k1, j) + A3(phb, i, k1 + 1, j) + A3(phb, i - di, k1, j - dj) + A3(phb, i - di, k1 + 1, j - dj) +
                    A3(ph, i, k1, j) + A3(ph, i, k1 + 1, j) + A3(ph, i - di, k1, j - dj) + A3(ph, i - di, k1 + 1, j - dj)) / WG_G;
  };
  float ztop;
  if (kind == 2) ztop = (A3(phb, i, kde, j) + A3(ph, i, kde, j)) / WG_G;
  else ztop = 0.5f * (A3(phb, i, kde, j) + A3(phb, i - di, kde, j - dj) + A3(ph, i, kde, j) + A3(ph, i - di, kde, j - dj)) / WG_G;
  if (kind == 2) {
    for (int k = 1; k <= kde; k++) {
      const float z = (A3(phb, i, k, j) + A3(ph, i, k, j)) / WG_G;
      if (z >= (ztop - zdamp)) {
        float dcoef = 1.0f - fminf(1.0f, (ztop - z) / zdamp);
        const float sn = wg_libm::sinf_(0.5f * pii * dcoef);
        dcoef = sn * sn;
        A3(rw_tendf, i, k, j) = A3(rw_tendf, i, k, j) - (g.c1f[k] * A2(mut, i, j) + g.c2f[k]) * (dcoef * dampcoef) * A3(w, i, k, j);
      }
    }
  }
  // find the bottom of the damping layer (DO WHILE, :5930-5945)
  int k1 = ktf;
  float z = ztop;
  while (z >= (ztop - zdamp)) {
    z = zmid(k1);
    k1 = k1 - 1;
  }
  k1 = k1 + 2;
  const float *base = kind == 0 ? g.u_base : (kind == 1 ? g.v_base : g.t_base);
  const float *zb = g.z_base;
  for (int k = k1; k <= ktf; k++) {
    const float z00 = zmid(k);
    int k2 = ktf;
    while (zb[k2] > z00) k2 = k2 - 1;
    float b00;
    if (k2 + 1 > ktf) b00 = base[k2] + (base[k2] - base[k2 - 1]) * (z00 - zb[k2]) / (zb[k2] - zb[k2 - 1]);
    else b00 = base[k2] + (base[k2 + 1] - base[k2]) * (z00 - zb[k2]) / (zb[k2 + 1] - zb[k2]);
    float dcoef = 1.0f - fminf(1.0f, (ztop - z00) / zdamp);
    const float sn = wg_libm::sinf_(0.5f * pii * dcoef);
    dcoef = sn * sn;
    if (kind == 0)
      A3(ru_tendf, i, k, j) = A3(ru_tendf, i, k, j) - (g.c1h[k] * A2(muu, i, j) + g.c2h[k]) * (dcoef * dampcoef) * (A3(u, i, k, j) - b00);
    else if (kind == 1)
      A3(rv_tendf, i, k, j) = A3(rv_tendf, i, k, j) - (g.c1h[k] * A2(muv, i, j) + g.c2h[k]) * (dcoef * dampcoef) * (A3(v, i, k, j) - b00);
    else
      A3(t_tendf, i, k, j) = A3(t_tendf, i, k, j) - (g.c1h[k] * A2(mut, i, j) + g.c2h[k]) * (dcoef * dampcoef) * (A3(t, i, k, j) - b00);
  }
}

// w_damp's CFL maxima (module_big_step_utilities_em.F:2503-2711): max over j = jts..jtf, k = 2..kde-1, i = its..itf of
// vert_cfl and horiz_cfl, the same real*4 expressions.  The maximum of non-negative floats is order independent and their
// bit patterns order like integers: warp shuffle, then one atomicMax per warp into cfl[0] (vertical), cfl[1] (horizontal).
__global__ void k_cfl_max(const Geo g, Box b, cfp u, cfp v, cfp ww, cfp mut, unsigned *cfl) {
  const int i = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int k = b.k0 + blockIdx.y * blockDim.y + threadIdx.y;
  const int j = b.j0 + blockIdx.z * blockDim.z + threadIdx.z;
  float vc = 0.f, hc = 0.f;
  if (i <= b.i1 && k <= b.k1 && j <= b.j1) {
    vc = fabsf(A3(ww, i, k, j) / (g.c1f[k] * A2(mut, i, j) + g.c2f[k]) * g.rdnw[k] * g.dt);
    hc = fmaxf(fabsf(A3(u, i, k, j) * g.rdx * A2(g.msfux, i, j) * g.dt), fabsf(A3(v, i, k, j) * g.rdy * A2(g.msfvy, i, j) * g.dt));
    if (!(vc >= 0.f)) vc = 0.f;   // NaN never wins a Fortran '>' comparison
    if (!(hc >= 0.f)) hc = 0.f;
  }
  for (int m = 16; m > 0; m >>= 1) {
    vc = fmaxf(vc, __shfl_xor_sync(0xffffffffu, vc, m));
    hc = fmaxf(hc, __shfl_xor_sync(0xffffffffu, hc, m));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(cfl, __float_as_uint(vc));
    atomicMax(cfl + 1, __float_as_uint(hc));
  }
}

__global__ void k_zero3_v4(const Geo g, Box b, fp a0, fp a1, fp a2, fp a3, fp a4) {   // 16-byte stores, four columns per thread
  TID3_V4;
  (void)him;
  const float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (a0) st4(&A3(a0, i, k, j), z, lo, hi);
  if (a1) st4(&A3(a1, i, k, j), z, lo, hi);
  if (a2) st4(&A3(a2, i, k, j), z, lo, hi);
  if (a3) st4(&A3(a3, i, k, j), z, lo, hi);
  if (a4) st4(&A3(a4, i, k, j), z, lo, hi);
}
__global__ void k_zero2(const Geo g, Box b, fp a0, fp a1) {
  TID2;
  if (a0) A2(a0, i, j) = 0.f;
  if (a1) A2(a1, i, j) = 0.f;
}

// module_em.F:1920 init_zero_tendency (first_rk_step_part1)
void Solver::init_zero_tendency() {
  const Geo &q = g;
  Box tile{q.its, q.ite, 1, q.kde, q.jts, q.jte};
  LAUNCH3(this, k_zero3_v4, box_v4(q, 1, q.kde), o("ru_tendf"), o("rv_tendf"), o("rw_tendf"), o("ph_tendf"), o("t_tendf"));
  LAUNCH2(this, k_zero2, tile, o("mu_tendf"), (float *)nullptr);
  for (size_t s = 0; s < species.size(); s += 5) {
    float *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (size_t m = 0; m < 5 && s + m < species.size(); m++) p[m] = o(species[s + m].t);
    LAUNCH3(this, k_zero3_v4, box_v4(q, 1, q.kde), p[0], p[1], p[2], p[3], p[4]);
  }
}

void Solver::rk_tendency(int rk_step, bool zero_save) {
  rk_tendency_main(rk_step, zero_save);
  if (spec()) spec_bdy_dry();   // solve_em.F:982-1007, after rk_addtend_dry
}

void Solver::rk_tendency_main(int rk_step, bool zero_save) {
  const Geo &q = g;
  Box tile{q.its, q.ite, 1, q.kde, q.jts, q.jte};
  // the *_save arrays and mu_tend/mu_save are zeroed by rk_tendency (module_em.F:370-428)
  if (zero_save) LAUNCH3(this, k_zero3_v4, box_v4(q, 1, q.kde), o("u_save"), o("v_save"), o("w_save"), o("ph_save"), o("t_save"));
  LAUNCH2(this, k_zero2, tile, o("mu_tend"), o("mu_save"));
  if (rk_step == 1 && c.diff_opt == 1) {   // forward_step: diff_opt1 (module_em.F:802-878), before rk_rayleigh_damp
    const float prandtl = 1.f / 3.0f;    // share/module_model_constants.F:86
    const float xkmh = c.khdif, xkhh = c.khdif / prandtl;
    cfp kmh = c.km_opt == 1 ? nullptr : o("xkmh"), khh = c.km_opt == 1 ? nullptr : o("xkhh");   // K arrays of diffusion_prep (deform.cu)
    launch_diffuse(this, 0, o("u_2"), o("ru_tendf"), xkmh, c.kvdif, kmh);
    launch_diffuse(this, 1, o("v_2"), o("rv_tendf"), xkmh, c.kvdif, kmh);
    launch_diffuse(this, 2, o("w_2"), o("rw_tendf"), xkmh, c.kvdif, kmh);
    launch_diffuse(this, 3, o("t_2"), o("t_tendf"), xkhh, 3.f * c.kvdif, khh);
  }
  if (rk_step == 1 && c.diff_6th_opt != 0) {   // module_em.F:880-930
    launch_diff6(this, 0, o("u_2"), o("ru_tendf"), o("mut"), c.dt);
    launch_diff6(this, 1, o("v_2"), o("rv_tendf"), o("mut"), c.dt);
    launch_diff6(this, 2, o("w_2"), o("rw_tendf"), o("mut"), c.dt);
    launch_diff6(this, 3, o("t_2"), o("t_tendf"), o("mut"), c.dt);
  }
  if (rk_step == 1 && c.damp_opt == 2) {
    const int jm = std::min(q.jte, q.jde - 1), im = std::min(q.ite, q.ide - 1);
    for (int kind = 0; kind < 3; kind++) {
      Box bx{q.its, kind == 0 ? std::min(q.ite, q.ide) : im, 1, 1, q.jts, kind == 1 ? std::min(q.jte, q.jde) : jm};
      LAUNCH2(this, k_rayleigh, bx, kind, o("u_2"), o("v_2"), o("w_2"), o("t_2"), o("ph_2"), o("phb"), o("mut"), o("muu"),
              o("muv"), o("ru_tendf"), o("rv_tendf"), o("rw_tendf"), o("t_tendf"));
    }
  }
  // relax_bdy_dry (solve_em.F:937-968) + the first-substep half of rk_addtend_dry: it only reads the state, so it can go
  // ahead of the tendency kernels, which add *_tendf last; it must follow every other contribution to *_tendf
  if (spec() && rk_step == 1) relax_bdy_dry(zero_save);
  TendIn s;
  s.u = o("u_2"); s.v = o("v_2"); s.w = o("w_2"); s.t = o("t_2"); s.ph = o("ph_2"); s.phb = o("phb");
  s.ru = o("ru"); s.rv = o("rv"); s.rw = o("rw"); s.ww = o("ww"); s.rom = o("ww");
  s.p = o("p"); s.pb = o("pb"); s.al = o("al"); s.alt = o("alt"); s.php = o("php");
  s.cqu = o("cqu"); s.cqv = o("cqv"); s.cqw = o("cqw"); s.h_diabatic = o("h_diabatic");
  s.mut = o("mut"); s.muu = o("muu"); s.muv = o("muv"); s.mu = o("mu_2"); s.mub = o("mub");
  s.ru_tendf = o("ru_tendf"); s.rv_tendf = o("rv_tendf"); s.rw_tendf = o("rw_tendf"); s.ph_tendf = o("ph_tendf");
  s.t_tendf = o("t_tendf");
  if (c.use_adaptive_time_step) {   // solve_em.F:914-922: the maxima are only consumed by the adaptive time step
    const Box cb{q.its, std::min(q.ite, q.ide - 1), 2, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)};
    LAUNCH3(this, k_cfl_max, cb, s.u, s.v, s.ww, s.mut, (unsigned *)d_cfl);
  }
  const Box in = interior_box(q);
  Box strips[4];
  const int ns = edge_strips(tile, in, strips);
  const bool weno_m = c.momentum_adv_opt == 3 && rk_step == c.rk_ord;   // WENO_MOM, module_em.F:491-600
  const bool weno_t = (c.scalar_adv_opt == 3 || c.scalar_adv_opt == 4 || c.moist_adv_opt == 3 || c.moist_adv_opt == 4) && rk_step == c.rk_ord;
  if (c.zadvect_implicit > 0 && rk_step != c.rk_ord) ieva_split(true, s.u, s.v, o("ww"), s.mut, false);
  if (c.zadvect_implicit > 0 && rk_step == c.rk_ord) {   // IEVA, CHK_IEVA module_ieva_em.F:13-43; order of module_em.F:438-715
    const float dt_step = c.dt / (float)(c.rk_ord - rk_step + 1);
    ieva_split(true, s.u, s.v, o("ww"), s.mut);
    s.rom = o("wwE");
    LAUNCH3(this, k_tend_u_part<1>, tile, s, o("ru_tend"), (int)weno_m);
    ieva_implicit(1, o("ru_tend"), o("u_1"), o("muu_old"), o("muu_new"), dt_step);
    LAUNCH3(this, k_tend_v_part<1>, tile, s, o("rv_tend"), (int)weno_m);
    ieva_implicit(2, o("rv_tend"), o("v_1"), o("muv_old"), o("muv_new"), dt_step);
    LAUNCH3(this, k_tend_w_part<1>, tile, s, o("rw_tend"), (int)weno_m);
    LAUNCH3(this, k_tend_t_part<1>, tile, s, o("t_tend"), (int)weno_t);
    ieva_implicit(0, o("t_tend"), o("t_1"), o("mut_old"), o("mut_new"), dt_step);
    LAUNCH3(this, k_tend_ph_part<1>, tile, s, o("ph_tend"));
    ieva_implicit(3, o("ph_tend"), nullptr, nullptr, nullptr, dt_step);
    ieva_implicit(4, o("rw_tend"), nullptr, nullptr, nullptr, dt_step);
    LAUNCH3(this, k_tend_u_part<2>, tile, s, o("ru_tend"), (int)weno_m);
    LAUNCH3(this, k_tend_v_part<2>, tile, s, o("rv_tend"), (int)weno_m);
    LAUNCH3(this, k_tend_w_part<2>, tile, s, o("rw_tend"), (int)weno_m);
    LAUNCH3(this, k_tend_t_part<2>, tile, s, o("t_tend"), (int)weno_t);
    LAUNCH3(this, k_tend_ph_part<2>, tile, s, o("ph_tend"));
    const Box cq{q.its, std::min(q.ite, q.ide - 1), 2, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)};
    LAUNCH3(this, k_invert_cqw, cq, o("cqw"));
    return;
  }
  // the k-marching interior kernels exist for the default orders (5, 3); other orders take the gather kernels everywhere
  const bool dflt_m = c.h_mom_adv_order == 5 && c.v_mom_adv_order == 3, dflt_s = c.h_sca_adv_order == 5 && c.v_sca_adv_order == 3;
  // default orders, no WENO, no perturbation Coriolis: the five tendencies of the interior in one tile-staged kernel
  // (tend_tile.cu), the gather kernels on the edge strips
  if (!weno_m && !weno_t && dflt_m && dflt_s && !c.pert_coriolis && !empty(in) && tend_tile_enabled()) {
    launch_tend_tile(this, in);
    for (int m = 0; m < ns; m++) {
      LAUNCH3(this, k_tend_u, strips[m], s, o("ru_tend"));
      LAUNCH3(this, k_tend_v, strips[m], s, o("rv_tend"));
      LAUNCH3(this, k_tend_w, strips[m], s, o("rw_tend"));
      LAUNCH3(this, k_tend_t, strips[m], s, o("t_tend"));
      LAUNCH3(this, k_tend_ph, strips[m], s, o("ph_tend"));
      const Box cq{q.its, std::min(q.ite, q.ide - 1), 2, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)};
      Box sb = strips[m];
      sb.i1 = std::min(sb.i1, cq.i1); sb.j1 = std::min(sb.j1, cq.j1); sb.k0 = cq.k0; sb.k1 = cq.k1;
      LAUNCH3(this, k_invert_cqw, sb, o("cqw"));
    }
    return;
  }
  // w is advected with the SCALAR orders (module_advect_em.F:4463-4464): its k-marching kernel (orders 5 / 3) needs dflt_s, not dflt_m
  const bool fast_w = !weno_m && dflt_s && !c.pert_coriolis && !empty(in);
  if (weno_m) {
    LAUNCH3(this, k_tend_u_weno, tile, s, o("ru_tend"));
    LAUNCH3(this, k_tend_v_weno, tile, s, o("rv_tend"));
    LAUNCH3(this, k_tend_w_weno, tile, s, o("rw_tend"));
  } else {
    // momentum: k-marching kernels on the interior (mom_fast.cu), the gather kernels on the edge strips.  u and v follow the
    // momentum orders, w the scalar ones (fast_w above)
    if (dflt_m && !c.pert_coriolis && !empty(in)) {
      for (int which = 0; which < 2; which++) launch_mom_fast(this, which, in);
      for (int m = 0; m < ns; m++) {
        LAUNCH3(this, k_tend_u, strips[m], s, o("ru_tend"));
        LAUNCH3(this, k_tend_v, strips[m], s, o("rv_tend"));
      }
    } else {
      LAUNCH3(this, k_tend_u, tile, s, o("ru_tend"));
      LAUNCH3(this, k_tend_v, tile, s, o("rv_tend"));
    }
    if (fast_w) {
      launch_mom_fast(this, 2, in);
      for (int m = 0; m < ns; m++) LAUNCH3(this, k_tend_w, strips[m], s, o("rw_tend"));
    } else {
      LAUNCH3(this, k_tend_w, tile, s, o("rw_tend"));
    }
  }
  if (weno_t) {   // theta follows the scalars' WENO option on the last substep: gather kernel on the whole tile
    LAUNCH3(this, k_tend_t_weno, tile, s, o("t_tend"));
  } else if (!dflt_s) {
    LAUNCH3(this, k_tend_t, tile, s, o("t_tend"));
  } else {   // theta: k-marching flux-sharing kernel on the interior, the gather kernel on the edge strips
    launch_advect_scalar_fast(this, true, in, s.t, s.ru, s.rv, s.rom, s.mut, s.t_tendf, s.h_diabatic, o("t_tend"));
    for (int m = 0; m < ns; m++) LAUNCH3(this, k_tend_t, strips[m], s, o("t_tend"));
  }
  if (!empty(in)) {   // ph: k-marching kernel on the interior (mom_fast.cu), the gather kernel on the edge strips
    launch_mom_fast(this, 3, in);
    for (int m = 0; m < ns; m++) LAUNCH3(this, k_tend_ph, strips[m], s, o("ph_tend"));
  } else {
    LAUNCH3(this, k_tend_ph, tile, s, o("ph_tend"));
  }
  {   // cqw <- 1/(1+cqw): done by k_mom_w_fast on the interior, here on what is left of the mass points
    const Box cq{q.its, std::min(q.ite, q.ide - 1), 2, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)};
    if (fast_w) {
      for (int m = 0; m < ns; m++) {
        Box sb = strips[m];
        sb.i1 = std::min(sb.i1, cq.i1); sb.j1 = std::min(sb.j1, cq.j1); sb.k0 = cq.k0; sb.k1 = cq.k1;
        LAUNCH3(this, k_invert_cqw, sb, o("cqw"));
      }
    } else {
      LAUNCH3(this, k_invert_cqw, cq, o("cqw"));
    }
  }
}

}  // namespace wg
