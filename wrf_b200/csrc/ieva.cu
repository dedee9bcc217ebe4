// ieva.cu -- implicit-explicit vertical advection (dyn_em/module_ieva_em.F, zadvect_implicit > 0), active on the
// last RK substep (CHK_IEVA :13-43).  WW_SPLIT (:49-323) splits ww into wwE, which the explicit flux-form advection
// sees, and wwI, which is advected by an implicit upwind scheme: one tridiagonal system per column and variable
// (advect_u/v/w/s/ph_implicit :560-1271, TRIDIAG2D :326-386).
//
// The reference assembles the systems in REAL(KIND=8) from REAL expressions and solves them in double precision;
// the kernels do the same operations in the same order (fp64 add/mul/div are IEEE, -fmad=false), one thread per
// column, coefficients generated level by level, gam and bb in per-thread local arrays.  The path runs once per
// time step and is latency-, not bandwidth-shaped; it is kept simple.
#include "common.cuh"

namespace wg {

#define IEVA_ALPHA_MAX 1.1f
#define IEVA_ALPHA_MIN 0.8f

// WW_SPLIT, IEVA branch (:210-300), on the tile + 1 ring
__global__ void k_ww_split_ieva(const Geo g, Box b, cfp u, cfp v, cfp ww, cfp mut, float dt, fp wwE, fp wwI) {
  TID3;
  const int ktf = g.kde;
  const float w = A3(ww, i, k, j);
  if (k == 1 || k == ktf) { A3(wwE, i, k, j) = w; A3(wwI, i, k, j) = 0.0f; return; }
  const float cmnx_ratio = IEVA_ALPHA_MIN / IEVA_ALPHA_MAX;
  const float cutoff = 2.0f - cmnx_ratio;
  const float r4cmx = 1.0f / (4.0f - 4.0f * cmnx_ratio);
  const float Ceps = 0.9f;
  const float cx = 0.25f * g.rdx * (A3(u, i + 1, k, j) + A3(u, i, k, j) + A3(u, i + 1, k - 1, j) + A3(u, i, k - 1, j));
  const float cy = 0.25f * g.rdy * (A3(v, i, k, j + 1) + A3(v, i, k, j) + A3(v, i, k - 1, j + 1) + A3(v, i, k - 1, j));
  const float cw_max = fmaxf(IEVA_ALPHA_MAX - dt * Ceps * sqrtf(cx * cx + cy * cy), 0.0f);
  const float cr = w * dt * g.rdnw[k] / (g.c1f[k] * A2(mut, i, j) + g.c2f[k]);
  if (cw_max > 0.0f) {
    const float cw_max2 = cw_max * cw_max;
    const float cw_min = cw_max * cmnx_ratio;
    const float cw = fabsf(cr);
    float cff;
    if (cw < cw_min) cff = cw_max2;
    else if (cw < cutoff * cw_min) cff = cw_max2 + r4cmx * ((cw - cw_min) * (cw - cw_min));
    else cff = cw_max * cw;
    float wfrac = cw_max2 / cff;
    wfrac = fmaxf(fminf(wfrac, 1.0f), 0.0f);
    A3(wwE, i, k, j) = w * wfrac;
    A3(wwI, i, k, j) = w * (1.0f - wfrac);
  } else {
    A3(wwE, i, k, j) = 0.0f;
    A3(wwI, i, k, j) = w;
  }
}

// WW_SPLIT, copy branch (:302-320): what the RK substeps before the last one leave in wwE / wwI.  The explicit kernels
// read ww itself on those substeps; the copy only keeps the two arrays as the reference has them.
__global__ void k_ww_split_copy(const Geo g, Box b, cfp ww, fp wwE, fp wwI) {
  TID3;
  A3(wwE, i, k, j) = A3(ww, i, k, j);
  A3(wwI, i, k, j) = 0.0f;
}

// calculate_full(mut_old, mub, mu_old) (module_em.F:457-460), one point wider than the tile
__global__ void k_mut_old(const Geo g, Box b, cfp mub, cfp mu_old, fp mut_old) {
  TID2;
  A2(mut_old, i, j) = A2(mub, i, j) + A2(mu_old, i, j);
}

// CALC_MUT_NEW (:431-557)
__global__ void k_calc_mut_new(const Geo g, Box b, cfp u, cfp v, cfp mut_old, cfp muu, cfp muv, float dt, fp mut_new) {
  TID2;
  const int ktf = g.kde - 1;
  float dmdt = 0.f;
  for (int k = 1; k <= ktf; k++) {
    const float divv = (A2(g.msftx, i, j) / g.rdnw[k]) *
                       (g.rdx * ((g.c1h[k] * A2(muu, i + 1, j) + g.c2h[k]) * A3(u, i + 1, k, j) / A2(g.msfuy, i + 1, j) -
                                 (g.c1h[k] * A2(muu, i, j) + g.c2h[k]) * A3(u, i, k, j) / A2(g.msfuy, i, j)) +
                        g.rdy * ((g.c1h[k] * A2(muv, i, j + 1) + g.c2h[k]) * A3(v, i, k, j + 1) * A2(g.msfvx_inv, i, j + 1) -
                                 (g.c1h[k] * A2(muv, i, j) + g.c2h[k]) * A3(v, i, k, j) * A2(g.msfvx_inv, i, j)));
    dmdt = dmdt + divv;
  }
  A2(mut_new, i, j) = A2(mut_old, i, j) - dt * dmdt;
}

// calc_mu_uv_1 (module_big_step_utilities_em.F:184-321): staggered column mass from a full mu
__global__ void k_mu_uv_1(const Geo g, Box b, cfp mu, fp muu, fp muv) {
  TID2;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1);
  if (jm) {
    float r;
    if (i == g.ids) r = 0.5f * (A2(mu, i, j) + A2(mu, g.periodic_x ? g.ids - 1 : g.ids, j));
    else if (i == g.ide) r = 0.5f * (A2(mu, i - 1, j) + A2(mu, g.periodic_x ? g.ide : g.ide - 1, j));
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i - 1, j));
    A2(muu, i, j) = r;
  }
  if (im) {
    float r;
    if (j == g.jds) r = 0.5f * (A2(mu, i, j) + A2(mu, i, g.periodic_y ? g.jds - 1 : g.jds));
    else if (j == g.jde) r = 0.5f * (A2(mu, i, j - 1) + A2(mu, i, g.periodic_y ? g.jde : g.jde - 1));
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i, j - 1));
    A2(muv, i, j) = r;
  }
}

// TRIDIAG2D (:326-386) for one column.  coef(k, a, b, c, r) yields row k; out(k, bb) consumes the solution.
template <class Coef, class Out>
__device__ __forceinline__ void tridiag_column(int kstart, int kend, Coef coef, Out out) {
  double gam[WG_NK], bb[WG_NK];
  double a, bcoef, c, r, bet, c_prev;
  coef(kstart, a, bcoef, c, r);
  bet = bcoef;
  bb[kstart] = r / bet;
  c_prev = c;
  for (int k = kstart + 1; k <= kend; k++) {
    coef(k, a, bcoef, c, r);
    gam[k] = c_prev / bet;
    bet = bcoef - a * gam[k];
    bb[k] = (r - a * bb[k - 1]) / bet;
    c_prev = c;
  }
  for (int k = kend - 1; k >= kstart; k--) bb[k] = bb[k] - gam[k + 1] * bb[k + 1];
  for (int k = kstart; k <= kend; k++) out(k, bb[k]);
}

// advect_ph_implicit (:560-702)
__global__ void k_ieva_ph(const Geo g, Box b, cfp pho, cfp phb, cfp wwI, cfp mut, float dt_rk, fp tendency) {
  TID2;
  const int ktf = g.kde - 1;
  const float msfty = A2(g.msfty, i, j), mu = A2(mut, i, j);
  tridiag_column(2, ktf,
    [&](int k, double &a, double &bc, double &c, double &r) {
      const float wiC = 0.5f * A3(wwI, i, k, j) * (g.rdnw[k - 1] + g.rdnw[k]) * msfty / (g.c1f[k] * mu + g.c2f[k]);
      a = (double)(-dt_rk * fmaxf(wiC, 0.0f));
      c = (double)(dt_rk * fminf(wiC, 0.0f));
      const double btmp = -a - c;
      bc = (double)1.0f + btmp;
      r = (double)(A3(tendency, i, k, j) * dt_rk * msfty / (g.c1f[k] * mu + g.c2f[k])) -
          a * (double)A3(pho, i, k - 1, j) - btmp * (double)A3(pho, i, k, j) - c * (double)A3(pho, i, k + 1, j) -
          a * (double)A3(phb, i, k - 1, j) - btmp * (double)A3(phb, i, k, j) - c * (double)A3(phb, i, k + 1, j);
    },
    [&](int k, double bb) { A3(tendency, i, k, j) = (float)bb * (g.c1f[k] * mu + g.c2f[k]) / dt_rk / msfty; });
}

// advect_s_implicit (:705-835), advect_u_implicit (:838-977), advect_v_implicit (:980-1116): the same system on a
// mass, u or v column.  di/dj select the neighbour column of the averaged wwI; msf_w scales wiL/wiR (1 for scalars),
// msf_r the right-hand side and the result.
template <int KIND>   // 0 scalar, 1 u, 2 v
__global__ void k_ieva_huv(const Geo g, Box b, cfp q_old, cfp rom, cfp mu_old, cfp mu_new, float dt_rk, fp tendency) {
  TID2;
  const int ktf = g.kde - 1;
  const float mo = A2(mu_old, i, j), mn = A2(mu_new, i, j);
  const float msf_w = KIND == 1 ? A2(g.msfuy, i, j) : (KIND == 2 ? A2(g.msfvy, i, j) : 1.f);
  const float msf_r = KIND == 1 ? A2(g.msfuy, i, j) : (KIND == 2 ? A2(g.msfvx, i, j) : 1.f);
  tridiag_column(1, ktf,
    [&](int k, double &a, double &bc, double &c, double &r) {
      const int km1 = k == 1 ? 1 : k - 1, kp1 = k == ktf ? ktf : k + 1;
      float wiL, wiR;
      if (KIND == 0) {
        wiL = A3(rom, i, k, j) * g.rdnw[k] / (g.c1h[k] * mn + g.c2h[k]);
        wiR = A3(rom, i, k + 1, j) * g.rdnw[k] / (g.c1h[k] * mn + g.c2h[k]);
      } else if (KIND == 1) {
        wiL = 0.5f * (A3(rom, i - 1, k, j) + A3(rom, i, k, j)) * g.rdnw[k] * msf_w / (g.c1h[k] * mn + g.c2h[k]);
        wiR = 0.5f * (A3(rom, i - 1, k + 1, j) + A3(rom, i, k + 1, j)) * g.rdnw[k] * msf_w / (g.c1h[k] * mn + g.c2h[k]);
      } else {
        wiL = 0.5f * (A3(rom, i, k, j - 1) + A3(rom, i, k, j)) * g.rdnw[k] * msf_w / (g.c1h[k] * mn + g.c2h[k]);
        wiR = 0.5f * (A3(rom, i, k + 1, j - 1) + A3(rom, i, k + 1, j)) * g.rdnw[k] * msf_w / (g.c1h[k] * mn + g.c2h[k]);
      }
      a = (double)(-dt_rk * fmaxf(wiL, 0.0f));
      c = (double)(dt_rk * fminf(wiR, 0.0f));
      const double btmp = (double)(dt_rk * (fmaxf(wiR, 0.0f) - fminf(wiL, 0.0f)));
      bc = (double)1.0f + btmp;
      const float rhs = KIND == 0 ? dt_rk * A3(tendency, i, k, j) : dt_rk * A3(tendency, i, k, j) * msf_r;
      r = (double)rhs - (double)(g.c1h[k] * mo + g.c2h[k]) *
                            (a * (double)A3(q_old, i, km1, j) + btmp * (double)A3(q_old, i, k, j) + c * (double)A3(q_old, i, kp1, j));
    },
    [&](int k, double bb) {
      A3(tendency, i, k, j) = KIND == 0 ? (float)bb / dt_rk : (float)bb / dt_rk / msf_r;
    });
}

// advect_w_implicit (:1120-1271)
__global__ void k_ieva_w(const Geo g, Box b, cfp w_old, cfp utend, cfp vtend, cfp rom, cfp ph_new, cfp ph_old, cfp ph_tend,
                         cfp mut_old, cfp mut, cfp mut_new, float dt_rk, fp tendency) {
  TID2;
  const int ktf = g.kde - 1;
  const float mo = A2(mut_old, i, j), mu = A2(mut, i, j), mn = A2(mut_new, i, j);
  const float msfty = A2(g.msfty, i, j);
  tridiag_column(2, ktf,
    [&](int k, double &a, double &bc, double &c, double &r) {
      const float wiL = 0.5f * (A3(rom, i, k - 1, j) + A3(rom, i, k, j)) * g.rdn[k] * msfty / (g.c1f[k] * mn + g.c2f[k]);
      const float wiR = 0.5f * (A3(rom, i, k + 1, j) + A3(rom, i, k, j)) * g.rdn[k] * msfty / (g.c1f[k] * mn + g.c2f[k]);
      a = (double)(-dt_rk * fmaxf(wiL, 0.0f));
      c = (double)(dt_rk * fminf(wiR, 0.0f));
      const double btmp = (double)(dt_rk * (fmaxf(wiR, 0.0f) - fminf(wiL, 0.0f)));
      bc = (double)1.0f + btmp;
      r = (double)(dt_rk * A3(tendency, i, k, j) * msfty) -
          (double)(g.c1f[k] * mo + g.c2f[k]) * (a * (double)A3(w_old, i, k - 1, j) + btmp * (double)A3(w_old, i, k, j) + c * (double)A3(w_old, i, k + 1, j));
      if (k == 2) {
        const float cf1 = g.cf1, cf2 = g.cf2, cf3 = g.cf3;
        const float dw = msfty * .5f * g.rdy *
                             ((A2(g.ht, i, j + 1) - A2(g.ht, i, j)) * (cf1 * A3(vtend, i, 1, j + 1) + cf2 * A3(vtend, i, 2, j + 1) + cf3 * A3(vtend, i, 3, j + 1)) +
                              (A2(g.ht, i, j) - A2(g.ht, i, j - 1)) * (cf1 * A3(vtend, i, 1, j) + cf2 * A3(vtend, i, 2, j) + cf3 * A3(vtend, i, 3, j))) +
                         A2(g.msftx, i, j) * .5f * g.rdx *
                             ((A2(g.ht, i + 1, j) - A2(g.ht, i, j)) * (cf1 * A3(utend, i + 1, 1, j) + cf2 * A3(utend, i + 1, 2, j) + cf3 * A3(utend, i + 1, 3, j)) +
                              (A2(g.ht, i, j) - A2(g.ht, i - 1, j)) * (cf1 * A3(utend, i, 1, j) + cf2 * A3(utend, i, 2, j) + cf3 * A3(utend, i, 3, j)));
        r = r - (double)(g.c1f[k] * mu + g.c2f[k]) * a * (double)dt_rk * (double)dw;
      }
      if (k == ktf) {
        const float dw = msfty * ((A3(ph_new, i, k + 1, j) - A3(ph_old, i, k + 1, j)) / dt_rk - A3(ph_tend, i, k + 1, j) / (g.c1f[k] * mu + g.c2f[k]) / WG_G);
        r = r - (double)(g.c1f[k] * mu + g.c2f[k]) * c * (double)(dw - A3(w_old, i, k + 1, j));
      }
    },
    [&](int k, double bb) { A3(tendency, i, k, j) = (float)bb / dt_rk / msfty; });
}

// ---------------------------------------------------------------------------------------------------------------
// host side

// WW_SPLIT + the mass estimates of rk_tendency (module_em.F:438-480) / rk_scalar_tend (:1222-1243)
void Solver::ieva_split(bool dry, cfp u, cfp v, cfp ww, cfp mut, bool active) {
  const Geo &q = g;
  const int im = std::min(q.ite, q.ide - 1), jm = std::min(q.jte, q.jde - 1);
  Box ring{q.its - 1, im + 1, 1, q.kde, q.jts - 1, jm + 1};
  if (!active) { LAUNCH3(this, k_ww_split_copy, ring, ww, o("wwE"), o("wwI")); return; }
  LAUNCH3(this, k_ww_split_ieva, ring, u, v, ww, mut, c.dt, o("wwE"), o("wwI"));
  Box ring2{q.its - 1, im + 1, 1, 1, q.jts - 1, jm + 1};
  LAUNCH2(this, k_mut_old, ring2, o("mub"), o("mu_1"), o("mut_old"));
  if (!dry) return;
  Box lo{q.its - 1, im, 1, 1, q.jts - 1, jm};
  LAUNCH2(this, k_calc_mut_new, lo, u, v, o("mut_old"), o("muu"), o("muv"), c.dt, o("mut_new"));
  Box tile{q.its, q.ite, 1, 1, q.jts, q.jte};
  LAUNCH2(this, k_mu_uv_1, tile, o("mut_old"), o("muu_old"), o("muv_old"));
  LAUNCH2(this, k_mu_uv_1, tile, o("mut_new"), o("muu_new"), o("muv_new"));
}

// which: 0 scalar on mass points (tendency, old value, old and new column mass given), 1 u, 2 v, 3 ph, 4 w
void Solver::ieva_implicit(int which, fp tendency, cfp q_old, cfp mu_old, cfp mu_new, float dt_rk) {
  const Geo &q = g;
  const int im = std::min(q.ite, q.ide - 1), jm = std::min(q.jte, q.jde - 1);
  Box mass{q.its, im, 1, 1, q.jts, jm};
  if (which == 0) {
    LAUNCH2(this, k_ieva_huv<0>, mass, q_old, o("wwI"), mu_old, mu_new, dt_rk, tendency);
  } else if (which == 1) {
    Box bx{q.its, q.ite, 1, 1, q.jts, jm};
    if (c.open_ys) bx.i0 = std::max(q.ids + 1, q.its);     // sic (:905-906)
    if (c.open_ye) bx.i1 = std::min(q.ide - 1, q.ite);
    if (c.periodic_x) { bx.i0 = q.its; bx.i1 = q.ite; }
    LAUNCH2(this, k_ieva_huv<1>, bx, q_old, o("wwI"), mu_old, mu_new, dt_rk, tendency);
  } else if (which == 2) {
    Box bx{q.its, im, 1, 1, q.jts, q.jte};
    if (c.open_ys) bx.j0 = std::max(q.jds + 1, q.jts);
    if (c.open_ye) bx.j1 = std::min(q.jde - 1, q.jte);
    LAUNCH2(this, k_ieva_huv<2>, bx, q_old, o("wwI"), mu_old, mu_new, dt_rk, tendency);
  } else if (which == 3) {
    LAUNCH2(this, k_ieva_ph, mass, o("ph_1"), o("phb"), o("wwI"), o("mut"), dt_rk, tendency);
  } else {
    LAUNCH2(this, k_ieva_w, mass, o("w_1"), o("ru_tend"), o("rv_tend"), o("wwI"), o("ph_2"), o("ph_1"), o("ph_tend"),
            o("mut_old"), o("mut"), o("mut_new"), dt_rk, tendency);
  }
}

}  // namespace wg
