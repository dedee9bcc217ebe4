// diffusion2.cpp -- diffusion in physical space, diff_opt = 2: vertical_diffusion_2
// (dyn_em/module_diffusion_em.F:4004-4457) with vertical_diffusion_u_2 / _v_2 / _w_2 / _s (:4463-4906),
// horizontal_diffusion_2 (:2864-3112) with horizontal_diffusion_u_2 / _v_2 / _w_2 / _s (:3118-3998), the stresses
// cal_titau_11_22_33 / _12_21 / _13_31 / _23_32 (:5331-5895), and the air density of phy_prep
// (module_big_step_utilities_em.F:4856), as called from first_rk_step_part2
// (module_first_rk_step_part2.F:1019-1122: "vertical diffusion first and then horizontal").
// sfs_opt = 0 and m_opt = 0 (no NBA stresses, no stress output); no TKE species (km_opt 2 is not restated).
// TEST INFRASTRUCTURE (see oracle.h).
#include "oracle.h"

namespace orc {

namespace {
struct Loc3 {   // automatic array (its-1:ite+1, kts:kte, jts-1:jte+1), zero-filled
  std::vector<float> v; A3 a;
  Loc3(const Cx &x) {
    const long ni = x.ite - x.its + 3, nk = x.kte - x.kts + 1, nj = x.jte - x.jts + 3;
    v.assign((size_t)(ni * nk * nj), 0.f);
    a.p = v.data(); a.ims = x.its - 1; a.kms = x.kts; a.jms = x.jts - 1; a.sk = ni; a.sj = ni * nk;
  }
};
struct Ext { int is, ie, js, je; };
// the loop ranges the four stress routines and the eight diffusion routines share, by staggering
struct R4 { int i_start, i_end, j_start, j_end; };
// 'c': corner points (cal_titau_11_22_33 on mass points uses the same un-staggered i/j limits, :5363-5380)
R4 range(const Cx &x, char kind) {
  const wrfgpu_config &cf = x.c;
  const bool spec = cf.specified || cf.nested;
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  R4 r;
  switch (kind) {
    case 'c':   // cal_titau_11_22_33, cal_titau_12_21
      r = {its, ite, jts, jte};
      if (cf.open_xs || spec) r.i_start = std::max(ids + 1, its);
      if (cf.open_xe || spec) r.i_end = std::min(ide - 1, ite);
      if (cf.open_ys || spec) r.j_start = std::max(jds + 1, jts);
      if (cf.open_ye || spec) r.j_end = std::min(jde - 1, jte);
      if (cf.periodic_x) { r.i_start = its; r.i_end = ite; }
      break;
    case 'u':   // cal_titau_13_31, horizontal_diffusion_u_2, vertical_diffusion_u_2
      r = {its, ite, jts, std::min(jte, jde - 1)};
      if (cf.open_xs || spec) r.i_start = std::max(ids + 1, its);
      if (cf.open_xe || spec) r.i_end = std::min(ide - 1, ite);
      if (cf.open_ys || spec) r.j_start = std::max(jds + 1, jts);
      if (cf.open_ye || spec) r.j_end = std::min(jde - 2, jte);
      if (cf.periodic_x) { r.i_start = its; r.i_end = ite; }
      break;
    case 'v':   // cal_titau_23_32, horizontal_diffusion_v_2, vertical_diffusion_v_2
      r = {its, std::min(ite, ide - 1), jts, jte};
      if (cf.open_xs || spec) r.i_start = std::max(ids + 1, its);
      if (cf.open_xe || spec) r.i_end = std::min(ide - 2, ite);
      if (cf.open_ys || spec) r.j_start = std::max(jds + 1, jts);
      if (cf.open_ye || spec) r.j_end = std::min(jde - 1, jte);
      if (cf.periodic_x) { r.i_start = its; r.i_end = std::min(ite, ide - 1); }
      break;
    default:    // 'm': horizontal_diffusion_w_2 / _s, vertical_diffusion_w_2 / _s
      r = {its, std::min(ite, ide - 1), jts, std::min(jte, jde - 1)};
      if (cf.open_xs || spec) r.i_start = std::max(ids + 1, its);
      if (cf.open_xe || spec) r.i_end = std::min(ide - 2, ite);
      if (cf.open_ys || spec) r.j_start = std::max(jds + 1, jts);
      if (cf.open_ye || spec) r.j_end = std::min(jde - 2, jte);
      if (cf.periodic_x) { r.i_start = its; r.i_end = std::min(ite, ide - 1); }
      break;
  }
  return r;
}
R4 extend(R4 r, const Ext &e) { return {r.i_start - e.is, r.i_end + e.ie, r.j_start - e.js, r.j_end + e.je}; }

// :5331-5450
void cal_titau_11_22_33(const Cx &x, A3 titau, A3 xkx, A3 defor, A3 rho, const Ext &e) {
  const int ktf = std::min(x.kte, x.kde - 1);
  const R4 r = extend(range(x, 'c'), e);
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) titau(i, k, j) = -rho(i, k, j) * xkx(i, k, j) * defor(i, k, j);
}
// :5456-5593
void cal_titau_12_21(const Cx &x, A3 titau, A3 xkx, A3 defor, A3 rho, const Ext &e) {
  const int ktf = std::min(x.kte, x.kde - 1);
  const R4 r = extend(range(x, 'c'), e);
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float rhoavg = 0.25f * (rho(i - 1, k, j) + rho(i, k, j) + rho(i - 1, k, j - 1) + rho(i, k, j - 1));
        const float xkxavg = rhoavg * 0.25f * (xkx(i - 1, k, j) + xkx(i, k, j) + xkx(i - 1, k, j - 1) + xkx(i, k, j - 1));
        titau(i, k, j) = -xkxavg * defor(i, k, j);
      }
}
// :5598-5743
void cal_titau_13_31(const Cx &x, A3 titau, A3 defor, A3 xkx, A3 rho, const Ext &e) {
  const int ktf = std::min(x.kte, x.kde - 1);
  const R4 r = extend(range(x, 'u'), e);
  const A1 &fnm = x.fnm, &fnp = x.fnp;
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = x.kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float rhoavg = 0.5f * (fnm(k) * (rho(i - 1, k, j) + rho(i, k, j)) + fnp(k) * (rho(i - 1, k - 1, j) + rho(i, k - 1, j)));
        const float xkxavg = rhoavg * 0.5f * (fnm(k) * (xkx(i, k, j) + xkx(i - 1, k, j)) + fnp(k) * (xkx(i, k - 1, j) + xkx(i - 1, k - 1, j)));
        titau(i, k, j) = -xkxavg * defor(i, k, j);
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      titau(i, x.kts, j) = 0.0f;
      titau(i, ktf + 1, j) = 0.0f;
    }
}
// :5749-5895
void cal_titau_23_32(const Cx &x, A3 titau, A3 defor, A3 xkx, A3 rho, const Ext &e) {
  const int ktf = std::min(x.kte, x.kde - 1);
  const R4 r = extend(range(x, 'v'), e);
  const A1 &fnm = x.fnm, &fnp = x.fnp;
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = x.kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float rhoavg = 0.5f * (fnm(k) * (rho(i, k, j) + rho(i, k, j - 1)) + fnp(k) * (rho(i, k - 1, j) + rho(i, k - 1, j - 1)));
        const float xkxavg = rhoavg * 0.5f * (fnm(k) * (xkx(i, k, j) + xkx(i, k, j - 1)) + fnp(k) * (xkx(i, k - 1, j) + xkx(i, k - 1, j - 1)));
        titau(i, k, j) = -xkxavg * defor(i, k, j);
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      titau(i, x.kts, j) = 0.0f;
      titau(i, ktf + 1, j) = 0.0f;
    }
}
}  // namespace

// phy_prep's density (module_big_step_utilities_em.F:4856): rho = 1/alt (1 + qv) on the mass points of the tile
void phy_prep_rho(const Cx &x, A3 rho, A3 alt, const A3 *qv) {
  for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
    for (int k = x.kts; k <= std::min(x.kte, x.kde - 1); k++)
      for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) rho(i, k, j) = 1.f / alt(i, k, j) * (1.f + (qv ? (*qv)(i, k, j) : 0.f));
}

// :3118-3317
void horizontal_diffusion_u_2(const Cx &x, A3 tendency, A3 defor11, A3 defor12, A3 xkmh, A3 zx, A3 zy, A3 rdzw, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &fnm = x.fnm, &fnp = x.fnp, &dnw = x.dnw;
  const A2 &msfux = x.msfux, &msfuy = x.msfuy;
  const float rdx = x.rdx, rdy = x.rdy;
  const R4 r = range(x, 'u');
  Loc3 l1(x), l2(x), l3(x), l4(x), l5(x), l6(x);
  A3 titau1 = l1.a, titau2 = l2.a, titau1avg = l3.a, titau2avg = l4.a, zx_at_u = l5.a, zy_at_u = l6.a;
  cal_titau_11_22_33(x, titau1, xkmh, defor11, rho, Ext{1, 0, 0, 0});
  cal_titau_12_21(x, titau2, xkmh, defor12, rho, Ext{0, 0, 0, 1});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        titau1avg(i, k, j) = 0.5f * (fnm(k) * (titau1(i - 1, k, j) + titau1(i, k, j)) + fnp(k) * (titau1(i - 1, k - 1, j) + titau1(i, k - 1, j)));
        titau2avg(i, k, j) = 0.5f * (fnm(k) * (titau2(i, k, j + 1) + titau2(i, k, j)) + fnp(k) * (titau2(i, k - 1, j + 1) + titau2(i, k - 1, j)));
        zx_at_u(i, k, j) = 0.5f * (zx(i, k, j) + zx(i, k + 1, j));
        zy_at_u(i, k, j) = 0.125f * (zy(i - 1, k, j) + zy(i, k, j) + zy(i - 1, k, j + 1) + zy(i, k, j + 1) + zy(i - 1, k + 1, j) + zy(i, k + 1, j) +
                                     zy(i - 1, k + 1, j + 1) + zy(i, k + 1, j + 1));
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      titau1avg(i, kts, j) = 0.f;
      titau1avg(i, ktf + 1, j) = 0.f;
      titau2avg(i, kts, j) = 0.f;
      titau2avg(i, ktf + 1, j) = 0.f;
      zx_at_u(i, kts, j) = 0.5f * (zx(i, kts, j) + zx(i, kts + 1, j));
      zy_at_u(i, kts, j) = 0.125f * (zy(i - 1, kts, j) + zy(i, kts, j) + zy(i - 1, kts, j + 1) + zy(i, kts, j + 1) + zy(i - 1, kts + 1, j) +
                                     zy(i, kts + 1, j) + zy(i - 1, kts + 1, j + 1) + zy(i, kts + 1, j + 1));
    }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float mrdx = msfux(i, j) * rdx;
        const float mrdy = msfuy(i, j) * rdy;
        const float tmpdz = (1.f / rdzw(i, k, j) + 1.f / rdzw(i - 1, k, j)) / 2.f;
        tendency(i, k, j) = tendency(i, k, j) + g * tmpdz / dnw(k) *
                                                    (mrdx * (titau1(i, k, j) - titau1(i - 1, k, j)) + mrdy * (titau2(i, k, j + 1) - titau2(i, k, j)) -
                                                     msfux(i, j) * zx_at_u(i, k, j) * (titau1avg(i, k + 1, j) - titau1avg(i, k, j)) / tmpdz -
                                                     msfuy(i, j) * zy_at_u(i, k, j) * (titau2avg(i, k + 1, j) - titau2avg(i, k, j)) / tmpdz);
      }
}

// :3323-3513
void horizontal_diffusion_v_2(const Cx &x, A3 tendency, A3 defor12, A3 defor22, A3 xkmh, A3 zx, A3 zy, A3 rdzw, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &fnm = x.fnm, &fnp = x.fnp, &dnw = x.dnw;
  const A2 &msfvx = x.msfvx, &msfvy = x.msfvy;
  const float rdx = x.rdx, rdy = x.rdy;
  const R4 r = range(x, 'v');
  Loc3 l1(x), l2(x), l3(x), l4(x), l5(x), l6(x);
  A3 titau1 = l1.a, titau2 = l2.a, titau1avg = l3.a, titau2avg = l4.a, zx_at_v = l5.a, zy_at_v = l6.a;
  cal_titau_12_21(x, titau1, xkmh, defor12, rho, Ext{0, 1, 0, 0});
  cal_titau_11_22_33(x, titau2, xkmh, defor22, rho, Ext{0, 0, 1, 0});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        titau1avg(i, k, j) = 0.5f * (fnm(k) * (titau1(i + 1, k, j) + titau1(i, k, j)) + fnp(k) * (titau1(i + 1, k - 1, j) + titau1(i, k - 1, j)));
        titau2avg(i, k, j) = 0.5f * (fnm(k) * (titau2(i, k, j - 1) + titau2(i, k, j)) + fnp(k) * (titau2(i, k - 1, j - 1) + titau2(i, k - 1, j)));
        zx_at_v(i, k, j) = 0.125f * (zx(i, k, j) + zx(i + 1, k, j) + zx(i, k, j - 1) + zx(i + 1, k, j - 1) + zx(i, k + 1, j) + zx(i + 1, k + 1, j) +
                                     zx(i, k + 1, j - 1) + zx(i + 1, k + 1, j - 1));
        zy_at_v(i, k, j) = 0.5f * (zy(i, k, j) + zy(i, k + 1, j));
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      titau1avg(i, kts, j) = 0.f;
      titau1avg(i, ktf + 1, j) = 0.f;
      titau2avg(i, kts, j) = 0.f;
      titau2avg(i, ktf + 1, j) = 0.f;
      zx_at_v(i, kts, j) = 0.125f * (zx(i, kts, j) + zx(i + 1, kts, j) + zx(i, kts, j - 1) + zx(i + 1, kts, j - 1) + zx(i, kts + 1, j) +
                                     zx(i + 1, kts + 1, j) + zx(i, kts + 1, j - 1) + zx(i + 1, kts + 1, j - 1));
      zy_at_v(i, kts, j) = 0.5f * (zy(i, kts, j) + zy(i, kts + 1, j));
    }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float mrdx = msfvx(i, j) * rdx;
        const float mrdy = msfvy(i, j) * rdy;
        const float tmpdz = (1.f / rdzw(i, k, j) + 1.f / rdzw(i, k, j - 1)) / 2.f;
        tendency(i, k, j) = tendency(i, k, j) + g * tmpdz / dnw(k) *
                                                    (mrdy * (titau2(i, k, j) - titau2(i, k, j - 1)) + mrdx * (titau1(i + 1, k, j) - titau1(i, k, j)) -
                                                     msfvx(i, j) * zx_at_v(i, k, j) * (titau1avg(i, k + 1, j) - titau1avg(i, k, j)) / tmpdz -
                                                     msfvy(i, j) * zy_at_v(i, k, j) * (titau2avg(i, k + 1, j) - titau2avg(i, k, j)) / tmpdz);
      }
}

// :3519-3705
void horizontal_diffusion_w_2(const Cx &x, A3 tendency, A3 defor13, A3 defor23, A3 xkmv, A3 zx, A3 zy, A3 rdz, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &dn = x.dn;
  const A2 &msftx = x.msftx, &msfty = x.msfty;
  const float rdx = x.rdx, rdy = x.rdy;
  const R4 r = range(x, 'm');
  Loc3 l1(x), l2(x), l3(x), l4(x), l5(x), l6(x);
  A3 titau1 = l1.a, titau2 = l2.a, titau1avg = l3.a, titau2avg = l4.a, zx_at_w = l5.a, zy_at_w = l6.a;
  cal_titau_13_31(x, titau1, defor13, xkmv, rho, Ext{0, 1, 0, 0});
  cal_titau_23_32(x, titau2, defor23, xkmv, rho, Ext{0, 0, 0, 1});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        titau1avg(i, k, j) = 0.25f * (titau1(i + 1, k + 1, j) + titau1(i, k + 1, j) + titau1(i + 1, k, j) + titau1(i, k, j));
        titau2avg(i, k, j) = 0.25f * (titau2(i, k + 1, j + 1) + titau2(i, k + 1, j) + titau2(i, k, j + 1) + titau2(i, k, j));
        zx_at_w(i, k, j) = 0.5f * (zx(i, k, j) + zx(i + 1, k, j));
        zy_at_w(i, k, j) = 0.5f * (zy(i, k, j) + zy(i, k, j + 1));
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      titau1avg(i, ktf + 1, j) = 0.f;
      titau2avg(i, ktf + 1, j) = 0.f;
    }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float mrdx = msftx(i, j) * rdx;
        const float mrdy = msfty(i, j) * rdy;
        tendency(i, k, j) = tendency(i, k, j) + g / (dn(k) * rdz(i, k, j)) *
                                                    (mrdx * (titau1(i + 1, k, j) - titau1(i, k, j)) + mrdy * (titau2(i, k, j + 1) - titau2(i, k, j)) -
                                                     msfty(i, j) * rdz(i, k, j) *
                                                         (zx_at_w(i, k, j) * (titau1avg(i, k, j) - titau1avg(i, k - 1, j)) +
                                                          zy_at_w(i, k, j) * (titau2avg(i, k, j) - titau2avg(i, k - 1, j))));
      }
}

// :3711-3998 (doing_tke = .false.)
void horizontal_diffusion_s(const Cx &x, A3 tendency, A3 var, A3 xkhh, A3 zx, A3 zy, A3 rdzw, A3 rho, bool doing_tke) {
  const int kts = x.kts, kte = x.kte, ktf = std::min(x.kte, x.kde - 1);
  const int ktes1 = kte - 1, ktes2 = kte - 2;
  const A1 &fnm = x.fnm, &fnp = x.fnp, &dnw = x.dnw, &dn = x.dn;
  const A2 &msftx = x.msftx, &msfty = x.msfty, &msfux = x.msfux, &msfvy = x.msfvy;
  const float rdx = x.rdx, rdy = x.rdy, cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  const R4 r = range(x, 'm');
  const int i_start = r.i_start, i_end = r.i_end, j_start = r.j_start, j_end = r.j_end;
  Loc3 l1(x), l2(x), l3(x), l4(x), l5(x), l6(x), l7(x);
  A3 H1avg = l1.a, H2avg = l2.a, H1 = l3.a, H2 = l4.a, xkxavg = l5.a, zx_at_m = l6.a, zy_at_m = l7.a;
  Loc3 l8(x);
  A3 tmptendf = l8.a;
  if (doing_tke)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tmptendf(i, k, j) = tendency(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end + 1; i++) xkxavg(i, k, j) = 0.5f * (xkhh(i - 1, k, j) + xkhh(i, k, j)) * 0.5f * (rho(i - 1, k, j) + rho(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end + 1; i++)
        H1avg(i, k, j) = 0.5f * (fnm(k) * (var(i - 1, k, j) + var(i, k, j)) + fnp(k) * (var(i - 1, k - 1, j) + var(i, k - 1, j)));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end + 1; i++) {
      H1avg(i, kts, j) = 0.5f * (cf1 * var(i, 1, j) + cf2 * var(i, 2, j) + cf3 * var(i, 3, j) + cf1 * var(i - 1, 1, j) + cf2 * var(i - 1, 2, j) +
                                 cf3 * var(i - 1, 3, j));
      H1avg(i, ktf + 1, j) = 0.5f * (var(i, ktes1, j) + (var(i, ktes1, j) - var(i, ktes2, j)) * 0.5f * dnw(ktes1) / dn(ktes1) + var(i - 1, ktes1, j) +
                                     (var(i - 1, ktes1, j) - var(i - 1, ktes2, j)) * 0.5f * dnw(ktes1) / dn(ktes1));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end + 1; i++) {
        const float tmpzx = 0.5f * (zx(i, k, j) + zx(i, k + 1, j));
        const float rdzu = 2.f / (1.f / rdzw(i, k, j) + 1.f / rdzw(i - 1, k, j));
        H1(i, k, j) = -msfux(i, j) * xkxavg(i, k, j) * (rdx * (var(i, k, j) - var(i - 1, k, j)) - tmpzx * (H1avg(i, k + 1, j) - H1avg(i, k, j)) * rdzu);
      }
  for (int j = j_start; j <= j_end + 1; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) xkxavg(i, k, j) = 0.5f * (xkhh(i, k, j - 1) + xkhh(i, k, j)) * 0.5f * (rho(i, k, j - 1) + rho(i, k, j));
  for (int j = j_start; j <= j_end + 1; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        H2avg(i, k, j) = 0.5f * (fnm(k) * (var(i, k, j - 1) + var(i, k, j)) + fnp(k) * (var(i, k - 1, j - 1) + var(i, k - 1, j)));
  for (int j = j_start; j <= j_end + 1; j++)
    for (int i = i_start; i <= i_end; i++) {
      H2avg(i, kts, j) = 0.5f * (cf1 * var(i, 1, j) + cf2 * var(i, 2, j) + cf3 * var(i, 3, j) + cf1 * var(i, 1, j - 1) + cf2 * var(i, 2, j - 1) +
                                 cf3 * var(i, 3, j - 1));
      H2avg(i, ktf + 1, j) = 0.5f * (var(i, ktes1, j) + (var(i, ktes1, j) - var(i, ktes2, j)) * 0.5f * dnw(ktes1) / dn(ktes1) + var(i, ktes1, j - 1) +
                                     (var(i, ktes1, j - 1) - var(i, ktes2, j - 1)) * 0.5f * dnw(ktes1) / dn(ktes1));
    }
  for (int j = j_start; j <= j_end + 1; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float tmpzy = 0.5f * (zy(i, k, j) + zy(i, k + 1, j));
        const float rdzv = 2.f / (1.f / rdzw(i, k, j) + 1.f / rdzw(i, k, j - 1));
        H2(i, k, j) = -msfvy(i, j) * xkxavg(i, k, j) * (rdy * (var(i, k, j) - var(i, k, j - 1)) - tmpzy * (H2avg(i, k + 1, j) - H2avg(i, k, j)) * rdzv);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        H1avg(i, k, j) = 0.5f * (fnm(k) * (H1(i + 1, k, j) + H1(i, k, j)) + fnp(k) * (H1(i + 1, k - 1, j) + H1(i, k - 1, j)));
        H2avg(i, k, j) = 0.5f * (fnm(k) * (H2(i, k, j + 1) + H2(i, k, j)) + fnp(k) * (H2(i, k - 1, j + 1) + H2(i, k - 1, j)));
        zx_at_m(i, k, j) = 0.25f * (zx(i, k, j) + zx(i + 1, k, j) + zx(i, k + 1, j) + zx(i + 1, k + 1, j));
        zy_at_m(i, k, j) = 0.25f * (zy(i, k, j) + zy(i, k, j + 1) + zy(i, k + 1, j) + zy(i, k + 1, j + 1));
      }
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      H1avg(i, kts, j) = 0.f;
      H1avg(i, ktf + 1, j) = 0.f;
      H2avg(i, kts, j) = 0.f;
      H2avg(i, ktf + 1, j) = 0.f;
      zx_at_m(i, kts, j) = 0.25f * (zx(i, kts, j) + zx(i + 1, kts, j) + zx(i, kts + 1, j) + zx(i + 1, kts + 1, j));
      zy_at_m(i, kts, j) = 0.25f * (zy(i, kts, j) + zy(i, kts, j + 1) + zy(i, kts + 1, j) + zy(i, kts + 1, j + 1));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float mrdx = msftx(i, j) * rdx;
        const float mrdy = msfty(i, j) * rdy;
        tendency(i, k, j) = tendency(i, k, j) + g / (dnw(k) * rdzw(i, k, j)) *
                                                    (mrdx * (H1(i + 1, k, j) - H1(i, k, j)) + mrdy * (H2(i, k, j + 1) - H2(i, k, j)) -
                                                     msftx(i, j) * zx_at_m(i, k, j) * (H1avg(i, k + 1, j) - H1avg(i, k, j)) * rdzw(i, k, j) -
                                                     msfty(i, j) * zy_at_m(i, k, j) * (H2avg(i, k + 1, j) - H2avg(i, k, j)) * rdzw(i, k, j));
      }
  if (doing_tke)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tendency(i, k, j) = tmptendf(i, k, j) + 2.f * (tendency(i, k, j) - tmptendf(i, k, j));
}

// :4463-4570
void vertical_diffusion_u_2(const Cx &x, A3 tendency, A3 defor13, A3 xkmv, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &dnw = x.dnw;
  const R4 r = range(x, 'u');
  Loc3 l1(x);
  A3 titau3 = l1.a;
  cal_titau_13_31(x, titau3, defor13, xkmv, rho, Ext{0, 0, 0, 0});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float rdzu = -g / (dnw(k));
        tendency(i, k, j) = tendency(i, k, j) - rdzu * (titau3(i, k + 1, j) - titau3(i, k, j));
      }
  for (int j = r.j_start; j <= r.j_end; j++) {
    const int k = kts;
    for (int i = r.i_start; i <= r.i_end; i++) {
      const float rdzu = -g / dnw(k);
      tendency(i, k, j) = tendency(i, k, j) - rdzu * (titau3(i, k + 1, j));
    }
  }
}
// :4576-4683
void vertical_diffusion_v_2(const Cx &x, A3 tendency, A3 defor23, A3 xkmv, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &dnw = x.dnw;
  const R4 r = range(x, 'v');
  Loc3 l1(x);
  A3 titau3 = l1.a;
  cal_titau_23_32(x, titau3, defor23, xkmv, rho, Ext{0, 0, 0, 0});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float rdzv = -g / dnw(k);
        tendency(i, k, j) = tendency(i, k, j) - rdzv * (titau3(i, k + 1, j) - titau3(i, k, j));
      }
  for (int j = r.j_start; j <= r.j_end; j++) {
    const int k = kts;
    for (int i = r.i_start; i <= r.i_end; i++) {
      const float rdzv = -g / dnw(k);
      tendency(i, k, j) = tendency(i, k, j) - rdzv * (titau3(i, k + 1, j));
    }
  }
}
// :4688-4783 (titau33 with xkmh, "xkmh added Oct2013")
void vertical_diffusion_w_2(const Cx &x, A3 tendency, A3 defor33, A3 xkmh, A3 rho) {
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const A1 &dn = x.dn;
  const R4 r = range(x, 'm');
  Loc3 l1(x);
  A3 titau3 = l1.a;
  cal_titau_11_22_33(x, titau3, xkmh, defor33, rho, Ext{0, 0, 0, 0});
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) tendency(i, k, j) = tendency(i, k, j) + g * (titau3(i, k, j) - titau3(i, k - 1, j)) / dn(k);
}
// :4789-4906 (doing_tke = .false.); var(i,k,j) - base(k) when base is given (the var_mix copies of the driver, :4239-4256, :4312-4330)
void vertical_diffusion_s(const Cx &x, A3 tendency, A3 var_in, const A1 *base, A3 xkhv, A3 rdz, A3 rho, bool doing_tke) {
  const int kts = x.kts, kte = x.kte, ktf = std::min(x.kte, x.kde - 1);
  const A1 &fnm = x.fnm, &fnp = x.fnp, &dnw = x.dnw;
  const R4 r = range(x, 'm');
  Loc3 l1(x), l2(x), l3(x);
  A3 H3 = l1.a, var = l2.a, tmptendf = l3.a;
  if (doing_tke)
    for (int j = r.j_start; j <= r.j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = r.i_start; i <= r.i_end; i++) tmptendf(i, k, j) = tendency(i, k, j);
  for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
    for (int k = kts; k <= kte - 1; k++)
      for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) var(i, k, j) = base ? var_in(i, k, j) - (*base)(k) : var_in(i, k, j);
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        float xkxavg = fnm(k) * xkhv(i, k, j) + fnp(k) * xkhv(i, k - 1, j);
        xkxavg = xkxavg * (fnm(k) * rho(i, k, j) + fnp(k) * rho(i, k - 1, j));
        H3(i, k, j) = -xkxavg * (var(i, k, j) - var(i, k - 1, j)) * rdz(i, k, j);
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      H3(i, kts, j) = 0.f;
      H3(i, ktf + 1, j) = 0.f;
    }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) tendency(i, k, j) = tendency(i, k, j) + g * (H3(i, k + 1, j) - H3(i, k, j)) / dnw(k);
  if (doing_tke)
    for (int j = r.j_start; j <= r.j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = r.i_start; i <= r.i_end; i++) tendency(i, k, j) = tmptendf(i, k, j) + 2.f * (tendency(i, k, j) - tmptendf(i, k, j));
}

// the surface momentum and heat / moisture fluxes of vertical_diffusion_2 (:4160-4236, :4262-4300, :4338-4353)
void surface_drag(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 u_2, A3 v_2, A3 rho, A2 ust) {
  const int kts = x.kts;
  const A1 &dnw = x.dnw;
  const float epsilon = 1.e-10f;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1), j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const float cd0 = x.c.tke_drag_coefficient;
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= x.ite; i++) {
      const float vavg = (v_2(i, kts, j) + v_2(i, kts, j + 1) + v_2(i - 1, kts, j) + v_2(i - 1, kts, j + 1)) / 4;
      const float V0_u = std::sqrt((u_2(i, kts, j) * u_2(i, kts, j)) + (vavg * vavg)) + epsilon;
      float tao_xz;
      if (x.c.isfflx == 0) tao_xz = cd0 * V0_u * u_2(i, kts, j);
      else {
        const float ustar = 0.5f * (ust(i, j) + ust(i - 1, j));
        tao_xz = ustar * ustar * u_2(i, kts, j) / V0_u;
      }
      ru_tendf(i, kts, j) = ru_tendf(i, kts, j) + g * tao_xz * 0.5f * (rho(i, kts, j) + rho(i - 1, kts, j)) / dnw(kts);
    }
  for (int j = j_start; j <= x.jte; j++)
    for (int i = i_start; i <= i_end; i++) {
      const float uavg = (u_2(i, kts, j) + u_2(i, kts, j - 1) + u_2(i + 1, kts, j) + u_2(i + 1, kts, j - 1)) / 4;
      const float V0_v = std::sqrt((v_2(i, kts, j) * v_2(i, kts, j)) + (uavg * uavg)) + epsilon;
      float tao_yz;
      if (x.c.isfflx == 0) tao_yz = cd0 * V0_v * v_2(i, kts, j);
      else {
        const float ustar = 0.5f * (ust(i, j) + ust(i, j - 1));
        tao_yz = ustar * ustar * v_2(i, kts, j) / V0_v;
      }
      rv_tendf(i, kts, j) = rv_tendf(i, kts, j) + g * tao_yz * 0.5f * (rho(i, kts, j) + rho(i, kts, j - 1)) / dnw(kts);
    }
}
// theta = th_phy of phy_prep at kts: (t + t0) / (1 + R_v/R_d qv) with use_theta_m = 1 (module_big_step_utilities_em.F:4830-4846)
void surface_heat_flux(const Cx &x, A3 rt_tendf, A3 t_2, const A3 *qv, A3 rho, A2 hfx, A2 qfx) {
  const int kts = x.kts;
  const A1 &dnw = x.dnw;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1), j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      const float q = qv ? (*qv)(i, kts, j) : 0.f;
      const float cpm = cp * (1.f + 0.8f * q);
      const float theta = (x.c.use_theta_m == 1 && qv) ? (t_2(i, kts, j) + t0) / (1.f + r_v / r_d * q) : t_2(i, kts, j) + t0;
      float heat_flux;
      if (x.c.isfflx == 0 || x.c.isfflx == 2) {
        heat_flux = x.c.tke_heat_flux;
        hfx(i, j) = heat_flux * cpm * rho(i, kts, j);
        if (x.c.use_theta_m == 1)
          rt_tendf(i, kts, j) = rt_tendf(i, kts, j) - g * heat_flux * (1.f + rvovrd * q) * rho(i, kts, j) / dnw(kts) - g * 1.61f * theta * qfx(i, j) / dnw(kts);
        else
          rt_tendf(i, kts, j) = rt_tendf(i, kts, j) - g * heat_flux * rho(i, kts, j) / dnw(kts);
      } else {
        heat_flux = hfx(i, j) / cpm;
        if (x.c.use_theta_m == 1)
          rt_tendf(i, kts, j) = rt_tendf(i, kts, j) - g * heat_flux * (1.f + rvovrd * q) / dnw(kts) - g * 1.61f * theta * qfx(i, j) / dnw(kts);
        else
          rt_tendf(i, kts, j) = rt_tendf(i, kts, j) - g * heat_flux / dnw(kts);
      }
    }
}
void surface_moisture_flux(const Cx &x, A3 qv_tendf, A2 qfx) {
  const int kts = x.kts;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1), j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  if (x.c.isfflx == 0) return;
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) qv_tendf(i, kts, j) = qv_tendf(i, kts, j) - g * qfx(i, j) / x.dnw(kts);
}

}  // namespace orc
