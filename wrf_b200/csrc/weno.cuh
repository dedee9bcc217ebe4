// weno.cuh -- 5th-order WENO scalar advection (advect_scalar_weno, dyn_em/module_advect_em.F:7889-8569),
// *_adv_opt = 3 and, inside the positive-definite limiter, *_adv_opt = 4 (advect_scalar_wenopd, :8573-9490).
// Shared by tend.cu (theta on the last RK substep, module_em.F:603-621) and scalar.cu (moist / scalar species).
//
// The reference evaluates the candidate fluxes and smoothness indicators in REAL, stores them in DOUBLE PRECISION
// variables and forms the non-linear weights in double precision (:7966-7969).  fp64 add / mul / div are IEEE on the
// device and -fmad=false keeps them uncontracted, so the face flux is bit-identical to the host evaluation.  The
// weights cost three fp64 divisions and one more for the normalisation per face; the scheme runs on the last RK
// substep only.
#pragma once
#include "common.cuh"

namespace wg {

// m3 .. p2 are the six values around the face (q(-3) .. q(+2)); upw = the wind blows from the low-index side
__device__ __forceinline__ float weno5(float vel, bool upw, float m3, float m2, float m1, float p0, float p1, float p2, const double eps = 1.0e-40) {
  const float qim2 = upw ? m3 : p2, qim1 = upw ? m2 : p1, qi = upw ? m1 : p0, qip1 = upw ? p0 : m1, qip2 = upw ? p1 : m2;
  const double f0 = (double)(1.f / 3.f * qim2 - 7.f / 6.f * qim1 + 11.f / 6.f * qi);
  const double f1 = (double)(-1.f / 6.f * qim1 + 5.f / 6.f * qi + 1.f / 3.f * qip1);
  const double f2 = (double)(1.f / 3.f * qi + 5.f / 6.f * qip1 - 1.f / 6.f * qip2);
  const float a0 = qim2 - 2.f * qim1 + qi, b0 = qim2 - 4.f * qim1 + 3.f * qi;
  const float a1 = qim1 - 2.f * qi + qip1, b1 = qim1 - qip1;
  const float a2 = qi - 2.f * qip1 + qip2, b2 = qip2 - 4.f * qip1 + 3.f * qi;
  const double beta0 = (double)(13.f / 12.f * (a0 * a0) + 1.f / 4.f * (b0 * b0));
  const double beta1 = (double)(13.f / 12.f * (a1 * a1) + 1.f / 4.f * (b1 * b1));
  const double beta2 = (double)(13.f / 12.f * (a2 * a2) + 1.f / 4.f * (b2 * b2));
  const double gi0 = 1.0 / 10.0, gi1 = 6.0 / 10.0, gi2 = 3.0 / 10.0;   // eps: 1.d-40 for scalars (:7968), 1.d-18 for momentum (:10950)
  const double wi0 = gi0 / ((eps + beta0) * (eps + beta0));
  const double wi1 = gi1 / ((eps + beta1) * (eps + beta1));
  const double wi2 = gi2 / ((eps + beta2) * (eps + beta2));
  const double sumwk = wi0 + wi1 + wi2;
  return (float)((double)vel * (wi0 * f0 + wi1 * f1 + wi2 * f2) / sumwk);
}

// statement function flux3 of advect_scalar_weno (:7974-7981): pre-divided constants, no sign(time_step) factor
__device__ __forceinline__ float weno_flux3(float qm2, float qm1, float q0, float qp1, float ua) {
  return ((7.f / 12.f) * (q0 + qm1) - (1.f / 12.f) * (qp1 + qm2)) + sign1(ua) * (1.f / 12.f) * ((qp1 - qm2) - 3.f * (q0 - qm1));
}

// order of the flux at face f: faces lo_edge+1 / hi_edge-1 are 2nd order, lo_edge+2 / hi_edge-2 3rd order next to a
// boundary that is neither periodic nor symmetric (:8017-8031), everything else takes the full stencil
__device__ __forceinline__ int weno_order(int f, bool dlo, bool dhi, int lo_edge, int hi_edge) {
  if ((!dlo || f >= lo_edge + 3) && (!dhi || f <= hi_edge - 3)) return 5;
  if (f == lo_edge + 1) return 2;
  if (f == lo_edge + 2) return 3;
  if (f == hi_edge - 1) return 2;
  return 3;
}

// advect_scalar_weno evaluated at one mass point: y, x, open-boundary and vertical contributions in the
// reference's order of accumulation (:8059-8565)
__device__ __forceinline__ float advect_scalar_weno_pt(const Geo &g, cfp q, cfp q_old, cfp ru, cfp rv, cfp rom, int i, int k, int j) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, ktf = g.kde - 1;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const bool dxs = !(g.periodic_x || g.sym_xs), dxe = !(g.periodic_x || g.sym_xe);
  const bool dys = !(g.periodic_y || g.sym_ys), dye = !(g.periodic_y || g.sym_ye);
  const float sg = g.sgn_ts;
  float t = 0.f;
  {
    const int j_lo = dys ? max(jts, jds + 1) : jts, j_hi = dye ? min(jte, jde - 2) : min(jte, jde - 1);
    if (j >= j_lo && j <= j_hi) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int jj = j + m;
        const int ord = weno_order(jj, dys, dye, jds, jde);
        if (ord == 5) {
          const float vel = 0.5f * (A3(rv, i, k, jj) + A3(rv, i, k, jj));
          fq[m] = weno5(vel, vel * sg >= 0.0f, A3(q, i, k, jj - 3), A3(q, i, k, jj - 2), A3(q, i, k, jj - 1), A3(q, i, k, jj),
                        A3(q, i, k, jj + 1), A3(q, i, k, jj + 2));
        } else if (ord == 3) {
          const float vel = A3(rv, i, k, jj);
          fq[m] = vel * weno_flux3(A3(q, i, k, jj - 2), A3(q, i, k, jj - 1), A3(q, i, k, jj), A3(q, i, k, jj + 1), vel);
        } else {
          fq[m] = 0.5f * A3(rv, i, k, jj) * (A3(q, i, k, jj) + A3(q, i, k, jj - 1));
        }
      }
      const float mrdy = A2(g.msftx, i, j) * g.rdy;
      t = t - mrdy * (fq[1] - fq[0]);
    }
  }
  {
    const int i_lo = dxs ? max(ids + 1, its) : its, i_hi = dxe ? min(ide - 2, ite) : min(ite, ide - 1);
    if (i >= i_lo && i <= i_hi) {
      float fq[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const int ii = i + m;
        const int ord = weno_order(ii, dxs, dxe, ids, ide);
        if (ord == 5) {
          const float vel = 0.5f * (A3(ru, ii, k, j) + A3(ru, ii, k, j));
          fq[m] = weno5(vel, vel * sg >= 0.0f, A3(q, ii - 3, k, j), A3(q, ii - 2, k, j), A3(q, ii - 1, k, j), A3(q, ii, k, j),
                        A3(q, ii + 1, k, j), A3(q, ii + 2, k, j));
        } else if (ord == 3) {
          const float vel = A3(ru, ii, k, j);
          fq[m] = vel * weno_flux3(A3(q, ii - 2, k, j), A3(q, ii - 1, k, j), A3(q, ii, k, j), A3(q, ii + 1, k, j), vel);
        } else {
          fq[m] = 0.5f * (A3(ru, ii, k, j)) * (A3(q, ii, k, j) + A3(q, ii - 1, k, j));
        }
      }
      const float mrdx = A2(g.msftx, i, j) * g.rdx;
      t = t - mrdx * (fq[1] - fq[0]);
    }
  }
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
  {
    float vf[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const int kk = k + m;
      if (kk == 1 || kk == g.kde) {
        vf[m] = 0.f;
      } else if (kk == 2 || kk == ktf) {
        vf[m] = A3(rom, i, kk, j) * (g.fnm[kk] * A3(q, i, kk, j) + g.fnp[kk] * A3(q, i, kk - 1, j));
      } else if (kk == 3 || kk == ktf - 1) {
        const float vel = A3(rom, i, kk, j);
        vf[m] = vel * weno_flux3(A3(q, i, kk - 2, j), A3(q, i, kk - 1, j), A3(q, i, kk, j), A3(q, i, kk + 1, j), -vel);
      } else {
        const float vel = 0.5f * (A3(rom, i, kk, j) + A3(rom, i, kk, j));
        vf[m] = weno5(vel, -vel * sg >= 0.0f, A3(q, i, kk - 3, j), A3(q, i, kk - 2, j), A3(q, i, kk - 1, j), A3(q, i, kk, j),
                      A3(q, i, kk + 1, j), A3(q, i, kk + 2, j));
      }
    }
    t = t - g.rdnw[k] * (vf[1] - vf[0]);
  }
  return t;
}

}  // namespace wg
