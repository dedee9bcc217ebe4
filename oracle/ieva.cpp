// ieva.cpp -- oracle restatement of dyn_em/module_ieva_em.F: implicit-explicit vertical advection
// (zadvect_implicit > 0, Wicker and Skamarock 2020), active on the last RK substep (CHK_IEVA :13-43).
// ww is split into an explicit part wwE (what the explicit flux-form advection sees) and an implicit part wwI
// (WW_SPLIT :49-323); wwI is advected with an upwind implicit scheme, one tridiagonal solve per column and variable
// (advect_u/v/w/s/ph_implicit :560-1271).  The tridiagonal systems are REAL(KIND=8) (:326-386) with coefficients
// that are REAL expressions assigned to double variables: the casts below mark exactly those assignments.
// Test infrastructure only; see oracle.h header.
#include "oracle.h"
namespace orc {

namespace {
constexpr float alpha_max = 1.1f, alpha_min = 0.8f;   // :7-8

// TRIDIAG2D (:326-386) for one column: a, b, c, r, bb are (kstart..kend)
struct Tri {
  std::vector<double> a, b, c, r, bb, gam;
  explicit Tri(int n) : a(n), b(n), c(n), r(n), bb(n), gam(n + 1) {}
  void solve(int kstart, int kend) {
    double bet = b[kstart];
    bb[kstart] = r[kstart] / bet;
    for (int k = kstart + 1; k <= kend; k++) {
      gam[k] = c[k - 1] / bet;
      bet = b[k] - a[k] * gam[k];
      bb[k] = (r[k] - a[k] * bb[k - 1]) / bet;
    }
    for (int k = kend - 1; k >= kstart; k--) bb[k] = bb[k] - gam[k + 1] * bb[k + 1];
  }
};
}  // namespace

// WW_SPLIT, IEVA branch (:210-300); the copy branch is orc::ww_split (big_step.cpp)
void ww_split_ieva(const Cx &x, A3 wwE, A3 wwI, A3 u, A3 v, A3 ww, A2 mut, float dt) {
  const int ktf = x.kte;
  const int i_start = x.its - 1, i_end = std::min(x.ite, x.ide - 1) + 1;
  const int j_start = x.jts - 1, j_end = std::min(x.jte, x.jde - 1) + 1;
  const float cmnx_ratio = alpha_min / alpha_max;
  const float cutoff = 2.0f - cmnx_ratio;
  const float r4cmx = 1.0f / (4.0f - 4.0f * cmnx_ratio);
  const float Ceps = 0.9f;
  for (int j = j_start; j <= j_end; j++) {
    for (int k = x.kts + 1; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) { wwE(i, k, j) = 0.f; wwI(i, k, j) = 0.f; }
    for (int i = i_start; i <= i_end; i++) {
      wwE(i, 1, j) = ww(i, 1, j); wwI(i, 1, j) = 0.0f;
      wwE(i, ktf, j) = ww(i, ktf, j); wwI(i, ktf, j) = 0.0f;
    }
    for (int k = 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float cx = 0.25f * x.rdx * (u(i + 1, k, j) + u(i, k, j) + u(i + 1, k - 1, j) + u(i, k - 1, j));
        const float cy = 0.25f * x.rdy * (v(i, k, j + 1) + v(i, k, j) + v(i, k - 1, j + 1) + v(i, k - 1, j));
        const float cw_max = std::max(alpha_max - dt * Ceps * std::sqrt(cx * cx + cy * cy), 0.0f);
        const float cr = ww(i, k, j) * dt * x.rdnw(k) / (x.c1f(k) * mut(i, j) + x.c2f(k));
        if (cw_max > 0.0f) {
          const float cw_max2 = cw_max * cw_max;
          const float cw_min = cw_max * cmnx_ratio;
          const float cw = std::fabs(cr);
          float cff;
          if (cw < cw_min) cff = cw_max2;
          else if (cw < cutoff * cw_min) cff = cw_max2 + r4cmx * ((cw - cw_min) * (cw - cw_min));
          else cff = cw_max * cw;
          float wfrac = cw_max2 / cff;
          wfrac = std::max(std::min(wfrac, 1.0f), 0.0f);
          wwE(i, k, j) = ww(i, k, j) * wfrac;
          wwI(i, k, j) = ww(i, k, j) * (1.0f - wfrac);
        } else {
          wwE(i, k, j) = 0.0f;
          wwI(i, k, j) = ww(i, k, j);
        }
      }
  }
}

// CALC_MUT_NEW (:431-557): column mass after one forward step of the mass divergence
void calc_mut_new(const Cx &x, A3 u, A3 v, A2 mut_old, A2 muu, A2 muv, A2 mut_new, float dt) {
  const int ktf = std::min(x.kte, x.kde - 1);
  const int i_start = x.its - 1, i_end = std::min(x.ite, x.ide - 1);
  const int j_start = x.jts - 1, j_end = std::min(x.jte, x.jde - 1);
  const A1 &c1h = x.c1h, &c2h = x.c2h;
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      float dmdt = 0.f;
      for (int k = x.kts; k <= ktf; k++) {
        const float divv = (x.msftx(i, j) / x.rdnw(k)) *
                           (x.rdx * ((c1h(k) * muu(i + 1, j) + c2h(k)) * u(i + 1, k, j) / x.msfuy(i + 1, j) - (c1h(k) * muu(i, j) + c2h(k)) * u(i, k, j) / x.msfuy(i, j)) +
                            x.rdy * ((c1h(k) * muv(i, j + 1) + c2h(k)) * v(i, k, j + 1) * x.msfvx_inv(i, j + 1) - (c1h(k) * muv(i, j) + c2h(k)) * v(i, k, j) * x.msfvx_inv(i, j)));
        dmdt = dmdt + divv;
      }
      mut_new(i, j) = mut_old(i, j) - dt * dmdt;
    }
}

// calculate_full(mut_old, mub, mu_old) of rk_tendency (module_em.F:457-460) on the rows and columns CALC_MUT_NEW and
// calc_mu_uv_1 read.  The reference fills its..ite / jts..jte of a tile-local array and reads one row and column
// beyond (uninitialised stack memory there); here the array is patch-sized and filled one point wider.
void calc_mut_old(const Cx &x, A2 mut_old, A2 mub, A2 mu_old) {
  for (int j = x.jts - 1; j <= std::min(x.jte, x.jde - 1) + 1; j++)
    for (int i = x.its - 1; i <= std::min(x.ite, x.ide - 1) + 1; i++) mut_old(i, j) = mub(i, j) + mu_old(i, j);
}

// advect_ph_implicit (:560-702); c1 = c1f, c2 = c2f, rdzw = rdnw
void advect_ph_implicit(const Cx &x, A3 pho, A3 tendency, A3 phb, A3 wwI, A2 mut, float dt_rk) {
  const int ktf = x.kte - 1, kts = x.kts;
  const int i_end = std::min(x.ite, x.ide - 1), j_end = std::min(x.jte, x.jde - 1);
  const A1 &c1 = x.c1f, &c2 = x.c2f, &rdzw = x.rdnw;
  Tri T(x.kte + 2);
  for (int j = x.jts; j <= j_end; j++)
    for (int i = x.its; i <= i_end; i++) {
      for (int k = kts + 1; k <= ktf; k++) {
        const float wiC = 0.5f * wwI(i, k, j) * (rdzw(k - 1) + rdzw(k)) * x.msfty(i, j) / (c1(k) * mut(i, j) + c2(k));
        T.a[k] = (double)(-dt_rk * std::max(wiC, 0.0f));
        T.c[k] = (double)(dt_rk * std::min(wiC, 0.0f));
        const double btmp = -T.a[k] - T.c[k];
        T.b[k] = (double)1.0f + btmp;
        T.r[k] = (double)(tendency(i, k, j) * dt_rk * x.msfty(i, j) / (c1(k) * mut(i, j) + c2(k))) -
                 T.a[k] * (double)pho(i, k - 1, j) - btmp * (double)pho(i, k, j) - T.c[k] * (double)pho(i, k + 1, j) -
                 T.a[k] * (double)phb(i, k - 1, j) - btmp * (double)phb(i, k, j) - T.c[k] * (double)phb(i, k + 1, j);
      }
      T.solve(kts + 1, ktf);
      for (int k = kts + 1; k <= ktf; k++) tendency(i, k, j) = (float)T.bb[k] * (c1(k) * mut(i, j) + c2(k)) / dt_rk / x.msfty(i, j);
    }
}

// advect_s_implicit (:705-835)
void advect_s_implicit(const Cx &x, A3 s_old, A3 tendency, A3 rom, A2 mut_old, A2 mut_new, float dt_rk) {
  const int ktf = std::min(x.kte, x.kde - 1), kts = x.kts;
  const int i_end = std::min(x.ite, x.ide - 1), j_end = std::min(x.jte, x.jde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &rdzw = x.rdnw;
  Tri T(x.kte + 2);
  for (int j = x.jts; j <= j_end; j++)
    for (int i = x.its; i <= i_end; i++) {
      for (int k = kts; k <= ktf; k++) {
        const int km1 = k == kts ? kts : k - 1, kp1 = k == ktf ? ktf : k + 1;
        const float wiL = rom(i, k, j) * rdzw(k) / (c1(k) * mut_new(i, j) + c2(k));
        const float wiR = rom(i, k + 1, j) * rdzw(k) / (c1(k) * mut_new(i, j) + c2(k));
        T.a[k] = (double)(-dt_rk * std::max(wiL, 0.0f));
        T.c[k] = (double)(dt_rk * std::min(wiR, 0.0f));
        const double btmp = (double)(dt_rk * (std::max(wiR, 0.0f) - std::min(wiL, 0.0f)));
        T.b[k] = (double)1.0f + btmp;
        T.r[k] = (double)(dt_rk * tendency(i, k, j)) -
                 (double)(c1(k) * mut_old(i, j) + c2(k)) * (T.a[k] * (double)s_old(i, km1, j) + btmp * (double)s_old(i, k, j) + T.c[k] * (double)s_old(i, kp1, j));
      }
      T.solve(kts, ktf);
      for (int k = kts; k <= ktf; k++) tendency(i, k, j) = (float)T.bb[k] / dt_rk;
    }
}

// advect_u_implicit (:838-977)
void advect_u_implicit(const Cx &x, A3 u_old, A3 tendency, A3 rom, A2 muu_old, A2 muu_new, float dt_rk) {
  const int ktf = std::min(x.kte, x.kde - 1), kts = x.kts;
  int i_start = x.its, i_end = x.ite;
  const int j_end = std::min(x.jte, x.jde - 1);
  if (x.c.open_ys) i_start = std::max(x.ids + 1, x.its);     // sic: the reference tests open_ys / open_ye here
  if (x.c.open_ye) i_end = std::min(x.ide - 1, x.ite);
  if (x.c.periodic_x) { i_start = x.its; i_end = x.ite; }
  const A1 &c1 = x.c1h, &c2 = x.c2h, &rdzw = x.rdnw;
  Tri T(x.kte + 2);
  for (int j = x.jts; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      for (int k = kts; k <= ktf; k++) {
        const int km1 = k == kts ? kts : k - 1, kp1 = k == ktf ? ktf : k + 1;
        const float wiL = 0.5f * (rom(i - 1, k, j) + rom(i, k, j)) * rdzw(k) * x.msfuy(i, j) / (c1(k) * muu_new(i, j) + c2(k));
        const float wiR = 0.5f * (rom(i - 1, k + 1, j) + rom(i, k + 1, j)) * rdzw(k) * x.msfuy(i, j) / (c1(k) * muu_new(i, j) + c2(k));
        T.a[k] = (double)(-dt_rk * std::max(wiL, 0.0f));
        T.c[k] = (double)(dt_rk * std::min(wiR, 0.0f));
        const double btmp = (double)(dt_rk * (std::max(wiR, 0.0f) - std::min(wiL, 0.0f)));
        T.b[k] = (double)1.0f + btmp;
        T.r[k] = (double)(dt_rk * tendency(i, k, j) * x.msfuy(i, j)) -
                 (double)(c1(k) * muu_old(i, j) + c2(k)) * (T.a[k] * (double)u_old(i, km1, j) + btmp * (double)u_old(i, k, j) + T.c[k] * (double)u_old(i, kp1, j));
      }
      T.solve(kts, ktf);
      for (int k = kts; k <= ktf; k++) tendency(i, k, j) = (float)T.bb[k] / dt_rk / x.msfuy(i, j);
    }
}

// advect_v_implicit (:980-1116)
void advect_v_implicit(const Cx &x, A3 v_old, A3 tendency, A3 rom, A2 muv_old, A2 muv_new, float dt_rk) {
  const int ktf = std::min(x.kte, x.kde - 1), kts = x.kts;
  const int i_end = std::min(x.ite, x.ide - 1);
  int j_start = x.jts, j_end = x.jte;
  if (x.c.open_ys) j_start = std::max(x.jds + 1, x.jts);
  if (x.c.open_ye) j_end = std::min(x.jde - 1, x.jte);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &rdzw = x.rdnw;
  Tri T(x.kte + 2);
  for (int j = j_start; j <= j_end; j++)
    for (int i = x.its; i <= i_end; i++) {
      for (int k = kts; k <= ktf; k++) {
        const int km1 = k == kts ? kts : k - 1, kp1 = k == ktf ? ktf : k + 1;
        const float wiL = 0.5f * (rom(i, k, j - 1) + rom(i, k, j)) * rdzw(k) * x.msfvy(i, j) / (c1(k) * muv_new(i, j) + c2(k));
        const float wiR = 0.5f * (rom(i, k + 1, j - 1) + rom(i, k + 1, j)) * rdzw(k) * x.msfvy(i, j) / (c1(k) * muv_new(i, j) + c2(k));
        T.a[k] = (double)(-dt_rk * std::max(wiL, 0.0f));
        T.c[k] = (double)(dt_rk * std::min(wiR, 0.0f));
        const double btmp = (double)(dt_rk * (std::max(wiR, 0.0f) - std::min(wiL, 0.0f)));
        T.b[k] = (double)1.0f + btmp;
        T.r[k] = (double)(dt_rk * tendency(i, k, j) * x.msfvx(i, j)) -
                 (double)(c1(k) * muv_old(i, j) + c2(k)) * (T.a[k] * (double)v_old(i, km1, j) + btmp * (double)v_old(i, k, j) + T.c[k] * (double)v_old(i, kp1, j));
      }
      T.solve(kts, ktf);
      for (int k = kts; k <= ktf; k++) tendency(i, k, j) = (float)T.bb[k] / dt_rk / x.msfvx(i, j);
    }
}

// advect_w_implicit (:1120-1271); c1 = c1f, c2 = c2f, rdzu = rdn.  The lower boundary row carries the terrain-following
// w of the tendencies of u and v (:1225-1237), the upper one the w implied by the geopotential tendency (:1239-1243).
void advect_w_implicit(const Cx &x, A3 w_old, A3 tendency, A3 utend, A3 vtend, A3 rom, A3 ph_new, A3 ph_old, A3 ph_tend,
                       A2 mut_old, A2 mut, A2 mut_new, float dt_rk) {
  const int ktf = std::min(x.kte, x.kde - 1), kts = x.kts;
  const int i_end = std::min(x.ite, x.ide - 1), j_end = std::min(x.jte, x.jde - 1);
  const A1 &c1 = x.c1f, &c2 = x.c2f, &rdzu = x.rdn;
  const A2 &ht = x.ht;
  const float cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  Tri T(x.kte + 2);
  for (int j = x.jts; j <= j_end; j++)
    for (int i = x.its; i <= i_end; i++) {
      for (int k = kts + 1; k <= ktf; k++) {
        const float wiL = 0.5f * (rom(i, k - 1, j) + rom(i, k, j)) * rdzu(k) * x.msfty(i, j) / (c1(k) * mut_new(i, j) + c2(k));
        const float wiR = 0.5f * (rom(i, k + 1, j) + rom(i, k, j)) * rdzu(k) * x.msfty(i, j) / (c1(k) * mut_new(i, j) + c2(k));
        T.a[k] = (double)(-dt_rk * std::max(wiL, 0.0f));
        T.c[k] = (double)(dt_rk * std::min(wiR, 0.0f));
        const double btmp = (double)(dt_rk * (std::max(wiR, 0.0f) - std::min(wiL, 0.0f)));
        T.b[k] = (double)1.0f + btmp;
        T.r[k] = (double)(dt_rk * tendency(i, k, j) * x.msfty(i, j)) -
                 (double)(c1(k) * mut_old(i, j) + c2(k)) * (T.a[k] * (double)w_old(i, k - 1, j) + btmp * (double)w_old(i, k, j) + T.c[k] * (double)w_old(i, k + 1, j));
        if (k == kts + 1) {
          const float dw = x.msfty(i, j) * .5f * x.rdy *
                               ((ht(i, j + 1) - ht(i, j)) * (cf1 * vtend(i, 1, j + 1) + cf2 * vtend(i, 2, j + 1) + cf3 * vtend(i, 3, j + 1)) +
                                (ht(i, j) - ht(i, j - 1)) * (cf1 * vtend(i, 1, j) + cf2 * vtend(i, 2, j) + cf3 * vtend(i, 3, j))) +
                           x.msftx(i, j) * .5f * x.rdx *
                               ((ht(i + 1, j) - ht(i, j)) * (cf1 * utend(i + 1, 1, j) + cf2 * utend(i + 1, 2, j) + cf3 * utend(i + 1, 3, j)) +
                                (ht(i, j) - ht(i - 1, j)) * (cf1 * utend(i, 1, j) + cf2 * utend(i, 2, j) + cf3 * utend(i, 3, j)));
          T.r[k] = T.r[k] - (double)(c1(k) * mut(i, j) + c2(k)) * T.a[k] * (double)dt_rk * (double)dw;
        }
        if (k == ktf) {
          const float dw = x.msfty(i, j) * ((ph_new(i, k + 1, j) - ph_old(i, k + 1, j)) / dt_rk - ph_tend(i, k + 1, j) / (c1(k) * mut(i, j) + c2(k)) / g);
          T.r[k] = T.r[k] - (double)(c1(k) * mut(i, j) + c2(k)) * T.c[k] * (double)(dw - w_old(i, k + 1, j));
        }
      }
      T.solve(kts + 1, ktf);
      for (int k = kts + 1; k <= ktf; k++) tendency(i, k, j) = (float)T.bb[k] / dt_rk / x.msfty(i, j);
    }
}

}  // namespace orc
