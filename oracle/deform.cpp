// deform.cpp -- the diffusion preamble of the first RK substep: compute_diff_metrics
// (dyn_em/module_diffusion_em.F:6882-7129), cal_deform_and_div (:17-1189) and smag2d_km (:1934-2043), called from
// first_rk_step_part2 (dyn_em/module_first_rk_step_part2.F:427-680).
// TEST INFRASTRUCTURE (see oracle.h).
//
// The reference's tile-local work arrays tmp1, hat, hatavg (its-1:ite+1, kts:kte, jts-1:jte+1) are automatic arrays.
// One element of `hat` is read before it is written -- hat(i_start-1,k,jde) in the defor13 / defor23 blocks when
// periodic_y puts j_end on jde (:775-792 fill i_start-1 only up to MIN(jte,jde-1)) -- i.e. stack garbage in the
// reference.  The values it produces, defor13/23(i_start,k,jde), are overwritten by phy_bc / the periodic exchange
// before anything reads them.  Here `hat` is filled over the whole box it is read on, so the result is defined.
#include "oracle.h"

namespace orc {

namespace {
struct Loc3 {   // (its-1:ite+1, kts:kte, jts-1:jte+1), zero-filled
  std::vector<float> v; A3 a;
  Loc3(const Cx &x) {
    const long ni = x.ite - x.its + 3, nk = x.kte - x.kts + 1, nj = x.jte - x.jts + 3;
    v.assign((size_t)(ni * nk * nj), 0.f);
    a.p = v.data(); a.ims = x.its - 1; a.kms = x.kts; a.jms = x.jts - 1; a.sk = ni; a.sj = ni * nk;
  }
};
struct Loc2 {
  std::vector<float> v; A2 a;
  Loc2(const Cx &x) {
    const long ni = x.ite - x.its + 3, nj = x.jte - x.jts + 3;
    v.assign((size_t)(ni * nj), 0.f);
    a.p = v.data(); a.ims = x.its - 1; a.jms = x.jts - 1; a.sj = ni;
  }
};
}  // namespace

// module_diffusion_em.F:6882-7129
void compute_diff_metrics(const Cx &x, A3 ph, A3 phb, A3 z, A3 rdz, A3 rdzw, A3 zx, A3 zy) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kte = x.kte;
  const float rdx = x.rdx, rdy = x.rdy;
  const wrfgpu_config &cf = x.c;
  const int ktf = std::min(kte, kde - 1);
  int i_start, i_end, j_start, j_end;
  // z_at_w (its-1:ite, kts:kte, jts-1:jte)
  Loc3 zw(x);
  A3 z_at_w = zw.a;
  j_start = jts - 1;
  j_end = jte;
  for (int j = j_start; j <= j_end; j++) {
    if ((j_start >= jts) && (j_end <= std::min(jte, jde - 1))) {   // sic: never true (j_start = jts-1)
      i_start = its - 1;
      i_end = ite;
    } else {
      i_start = its;
      i_end = std::min(ite, ide - 1);
    }
    for (int k = 1; k <= kte; k++)
      for (int i = i_start; i <= i_end; i++) z_at_w(i, k, j) = (ph(i, k, j) + phb(i, k, j)) / g;
    for (int k = 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) rdzw(i, k, j) = 1.0f / (z_at_w(i, k + 1, j) - z_at_w(i, k, j));
    for (int k = 2; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) rdz(i, k, j) = 2.0f / (z_at_w(i, k + 1, j) - z_at_w(i, k - 1, j));
    for (int i = i_start; i <= i_end; i++) rdz(i, 1, j) = 2.f / (z_at_w(i, 2, j) - z_at_w(i, 1, j));
  }
  i_start = its;
  i_end = std::min(ite, ide - 1);
  j_start = jts;
  j_end = std::min(jte, jde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int k = 1; k <= kte; k++)
      for (int i = std::max(ids + 1, its); i <= i_end; i++) zx(i, k, j) = rdx * (phb(i, k, j) - phb(i - 1, k, j)) / g;
  for (int j = j_start; j <= j_end; j++)
    for (int k = 1; k <= kte; k++)
      for (int i = std::max(ids + 1, its); i <= i_end; i++) zx(i, k, j) = zx(i, k, j) + rdx * (ph(i, k, j) - ph(i - 1, k, j)) / g;
  for (int j = std::max(jds + 1, jts); j <= j_end; j++)
    for (int k = 1; k <= kte; k++)
      for (int i = i_start; i <= i_end; i++) zy(i, k, j) = rdy * (phb(i, k, j) - phb(i, k, j - 1)) / g;
  for (int j = std::max(jds + 1, jts); j <= j_end; j++)
    for (int k = 1; k <= kte; k++)
      for (int i = i_start; i <= i_end; i++) zy(i, k, j) = zy(i, k, j) + rdy * (ph(i, k, j) - ph(i, k, j - 1)) / g;
  if (!cf.periodic_x) {
    if (ite == ide)
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ide, k, j) = 0.0f;
    if (its == ids)
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ids, k, j) = 0.0f;
  } else {
    if (ite == ide) {
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ide, k, j) = rdx * (phb(ide, k, j) - phb(ide - 1, k, j)) / g;
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ide, k, j) = zx(ide, k, j) + rdx * (ph(ide, k, j) - ph(ide - 1, k, j)) / g;
    }
    if (its == ids) {
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ids, k, j) = rdx * (phb(ids, k, j) - phb(ids - 1, k, j)) / g;
      for (int j = j_start; j <= j_end; j++)
        for (int k = 1; k <= ktf; k++) zx(ids, k, j) = zx(ids, k, j) + rdx * (ph(ids, k, j) - ph(ids - 1, k, j)) / g;
    }
  }
  if (!cf.periodic_y) {
    if (jte == jde)
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jde) = 0.0f;
    if (jts == jds)
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jds) = 0.0f;
  } else {
    if (jte == jde) {
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jde) = rdy * (phb(i, k, jde) - phb(i, k, jde - 1)) / g;
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jde) = zy(i, k, jde) + rdy * (ph(i, k, jde) - ph(i, k, jde - 1)) / g;
    }
    if (jts == jds) {
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jds) = rdy * (phb(i, k, jds) - phb(i, k, jds - 1)) / g;
      for (int k = 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) zy(i, k, jds) = zy(i, k, jds) + rdy * (ph(i, k, jds) - ph(i, k, jds - 1)) / g;
    }
  }
  for (int j = j_start; j <= j_end; j++)
    for (int k = 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        z(i, k, j) = 0.5f * (ph(i, k, j) + phb(i, k, j) + ph(i, k + 1, j) + phb(i, k + 1, j)) / g;
}

// module_diffusion_em.F:17-1189 (sfs_opt = 0: no nba_rij; polar = 0)
void cal_deform_and_div(const Cx &x, A3 u, A3 v, A3 w, A3 div, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13,
                        A3 defor23, A3 rdz, A3 rdzw, A3 zx, A3 zy) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const float rdx = x.rdx, rdy = x.rdy, cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  const wrfgpu_config &cf = x.c;
  const A1 &fnm = x.fnm, &fnp = x.fnp, &dn = x.dn, &dnw = x.dnw, &u_base = x.u_base, &v_base = x.v_base;
  const A2 &msfux = x.msfux, &msfuy = x.msfuy, &msfvx = x.msfvx, &msfvy = x.msfvy, &msftx = x.msftx, &msfty = x.msfty;
  const bool spec = cf.specified || cf.nested;
  Loc3 l1(x), l2(x), l3(x);
  Loc2 lm(x);
  A3 tmp1 = l1.a, hat = l2.a, hatavg = l3.a;
  A2 mm = lm.a;
  float tmpzx, tmpzy;
  int i_start, i_end, j_start, j_end;

  const int ktes1 = kte - 1, ktes2 = kte - 2;
  const float cft2 = -0.5f * dnw(ktes1) / dn(ktes1);
  const float cft1 = 1.0f - cft2;
  const int ktf = std::min(kte, kde - 1);

  // ---- defor11, defor22, defor33, div (:205-609)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) mm(i, j) = msftx(i, j) * msfty(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end + 1; i++) hat(i, k, j) = u(i, k, j) / msfuy(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        hatavg(i, k, j) = 0.5f * (fnm(k) * (hat(i, k, j) + hat(i + 1, k, j)) + fnp(k) * (hat(i, k - 1, j) + hat(i + 1, k - 1, j)));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      hatavg(i, 1, j) = 0.5f * (cf1 * hat(i, 1, j) + cf2 * hat(i, 2, j) + cf3 * hat(i, 3, j) + cf1 * hat(i + 1, 1, j) + cf2 * hat(i + 1, 2, j) +
                                cf3 * hat(i + 1, 3, j));
      hatavg(i, kte, j) = 0.5f * (cft1 * (hat(i, ktes1, j) + hat(i + 1, ktes1, j)) + cft2 * (hat(i, ktes2, j) + hat(i + 1, ktes2, j)));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmpzx = 0.25f * (zx(i, k, j) + zx(i + 1, k, j) + zx(i, k + 1, j) + zx(i + 1, k + 1, j));
        tmp1(i, k, j) = (hatavg(i, k + 1, j) - hatavg(i, k, j)) * tmpzx * rdzw(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = mm(i, j) * (rdx * (hat(i + 1, k, j) - hat(i, k, j)) - tmp1(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor11(i, k, j) = 2.0f * tmp1(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) div(i, k, j) = tmp1(i, k, j);
  for (int j = j_start; j <= j_end + 1; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) hat(i, k, j) = v(i, k, j) / msfvx(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        hatavg(i, k, j) = 0.5f * (fnm(k) * (hat(i, k, j) + hat(i, k, j + 1)) + fnp(k) * (hat(i, k - 1, j) + hat(i, k - 1, j + 1)));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      hatavg(i, 1, j) = 0.5f * (cf1 * hat(i, 1, j) + cf2 * hat(i, 2, j) + cf3 * hat(i, 3, j) + cf1 * hat(i, 1, j + 1) + cf2 * hat(i, 2, j + 1) +
                                cf3 * hat(i, 3, j + 1));
      hatavg(i, kte, j) = 0.5f * (cft1 * (hat(i, ktes1, j) + hat(i, ktes1, j + 1)) + cft2 * (hat(i, ktes2, j) + hat(i, ktes2, j + 1)));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmpzy = 0.25f * (zy(i, k, j) + zy(i, k, j + 1) + zy(i, k + 1, j) + zy(i, k + 1, j + 1));
        tmp1(i, k, j) = (hatavg(i, k + 1, j) - hatavg(i, k, j)) * tmpzy * rdzw(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = mm(i, j) * (rdy * (hat(i, k, j + 1) - hat(i, k, j)) - tmp1(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor22(i, k, j) = 2.0f * tmp1(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) div(i, k, j) = div(i, k, j) + tmp1(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = (w(i, k + 1, j) - w(i, k, j)) * rdzw(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor33(i, k, j) = 2.0f * tmp1(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) div(i, k, j) = div(i, k, j) + tmp1(i, k, j);

  // ---- defor12 (:611-870)
  i_start = its; i_end = ite; j_start = jts; j_end = jte;
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_xe || spec) i_end = std::min(ide - 1, ite);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || spec) j_end = std::min(jde - 1, jte);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = ite;
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) mm(i, j) = 0.25f * (msfux(i, j - 1) + msfux(i, j)) * (msfvy(i - 1, j) + msfvy(i, j));
  for (int j = j_start - 1; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) hat(i, k, j) = u(i, k, j) / msfux(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        hatavg(i, k, j) = 0.5f * (fnm(k) * (hat(i, k, j - 1) + hat(i, k, j)) + fnp(k) * (hat(i, k - 1, j - 1) + hat(i, k - 1, j)));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      hatavg(i, 1, j) = 0.5f * (cf1 * hat(i, 1, j - 1) + cf2 * hat(i, 2, j - 1) + cf3 * hat(i, 3, j - 1) + cf1 * hat(i, 1, j) + cf2 * hat(i, 2, j) +
                                cf3 * hat(i, 3, j));
      hatavg(i, kte, j) = 0.5f * (cft1 * (hat(i, ktes1, j - 1) + hat(i, ktes1, j)) + cft2 * (hat(i, ktes2, j - 1) + hat(i, ktes2, j)));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmpzy = 0.25f * (zy(i - 1, k, j) + zy(i, k, j) + zy(i - 1, k + 1, j) + zy(i, k + 1, j));
        tmp1(i, k, j) = (hatavg(i, k + 1, j) - hatavg(i, k, j)) * 0.25f * tmpzy *
                        (rdzw(i, k, j) + rdzw(i - 1, k, j) + rdzw(i - 1, k, j - 1) + rdzw(i, k, j - 1));
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor12(i, k, j) = mm(i, j) * (rdy * (hat(i, k, j) - hat(i, k, j - 1)) - tmp1(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start - 1; i <= i_end; i++) hat(i, k, j) = v(i, k, j) / msfvy(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        hatavg(i, k, j) = 0.5f * (fnm(k) * (hat(i - 1, k, j) + hat(i, k, j)) + fnp(k) * (hat(i - 1, k - 1, j) + hat(i, k - 1, j)));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      hatavg(i, 1, j) = 0.5f * (cf1 * hat(i - 1, 1, j) + cf2 * hat(i - 1, 2, j) + cf3 * hat(i - 1, 3, j) + cf1 * hat(i, 1, j) + cf2 * hat(i, 2, j) +
                                cf3 * hat(i, 3, j));
      hatavg(i, kte, j) = 0.5f * (cft1 * (hat(i, ktes1, j) + hat(i - 1, ktes1, j)) + cft2 * (hat(i, ktes2, j) + hat(i - 1, ktes2, j)));
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmpzx = 0.25f * (zx(i, k, j - 1) + zx(i, k, j) + zx(i, k + 1, j - 1) + zx(i, k + 1, j));
        tmp1(i, k, j) = (hatavg(i, k + 1, j) - hatavg(i, k, j)) * 0.25f * tmpzx *
                        (rdzw(i, k, j) + rdzw(i, k, j - 1) + rdzw(i - 1, k, j - 1) + rdzw(i - 1, k, j));
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        defor12(i, k, j) = defor12(i, k, j) + mm(i, j) * (rdx * (hat(i, k, j) - hat(i - 1, k, j)) - tmp1(i, k, j));
  if (!cf.periodic_x && i_start == ids + 1)
    for (int j = jts; j <= jte; j++)
      for (int k = kts; k <= kte; k++) defor12(ids, k, j) = defor12(ids + 1, k, j);
  if (!cf.periodic_y && j_start == jds + 1)
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= ite; i++) defor12(i, k, jds) = defor12(i, k, jds + 1);
  if (!cf.periodic_x && i_end == ide - 1)
    for (int j = jts; j <= jte; j++)
      for (int k = kts; k <= kte; k++) defor12(ide, k, j) = defor12(ide - 1, k, j);
  if (!cf.periodic_y && j_end == jde - 1)
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= ite; i++) defor12(i, k, jde) = defor12(i, k, jde - 1);

  // ---- defor13 (:872-1010)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide);
  if (cf.periodic_y) j_end = std::min(jte, jde);
  for (int j = jts; j <= jte; j++)
    for (int i = its; i <= ite; i++) mm(i, j) = msfux(i, j) * msfuy(i, j);
  // hat = w / msfty on [i_start-1, i_end] x [j_start-1, j_end] (the reference fills the interior, the column i_start-1
  // for j <= MIN(jte,jde-1) and the row j_start-1 for i <= MIN(ite,ide-1); see the header for the one element it misses)
  for (int j = j_start - 1; j <= j_end; j++)
    for (int k = kts; k <= kte; k++)
      for (int i = i_start - 1; i <= i_end; i++) hat(i, k, j) = w(i, k, j) / msfty(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) hatavg(i, k, j) = 0.25f * (hat(i, k, j) + hat(i, k + 1, j) + hat(i - 1, k, j) + hat(i - 1, k + 1, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tmp1(i, k, j) = (hatavg(i, k, j) - hatavg(i, k - 1, j)) * zx(i, k, j) * 0.5f * (rdz(i, k, j) + rdz(i - 1, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor13(i, k, j) = mm(i, j) * (rdx * (hat(i, k, j) - hat(i - 1, k, j)) - tmp1(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      defor13(i, kts, j) = 0.0f;
      defor13(i, ktf + 1, j) = 0.0f;
    }
  if (cf.mix_full_fields) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = (u(i, k, j) - u(i, k - 1, j)) * 0.5f * (rdz(i, k, j) + rdz(i - 1, k, j));
  } else {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++)
          tmp1(i, k, j) = (u(i, k, j) - u_base(k) - u(i, k - 1, j) + u_base(k - 1)) * 0.5f * (rdz(i, k, j) + rdz(i - 1, k, j));
  }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor13(i, k, j) = defor13(i, k, j) + tmp1(i, k, j);

  // ---- defor23 (:1012-1187)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.periodic_y) j_end = std::min(jte, jde);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide - 1);
  for (int j = jts; j <= jte; j++)
    for (int i = its; i <= ite; i++) mm(i, j) = msfvx(i, j) * msfvy(i, j);
  for (int j = j_start - 1; j <= j_end; j++)
    for (int k = kts; k <= kte; k++)
      for (int i = i_start - 1; i <= i_end; i++) hat(i, k, j) = w(i, k, j) / msftx(i, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) hatavg(i, k, j) = 0.25f * (hat(i, k, j) + hat(i, k + 1, j) + hat(i, k, j - 1) + hat(i, k + 1, j - 1));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tmp1(i, k, j) = (hatavg(i, k, j) - hatavg(i, k - 1, j)) * zy(i, k, j) * 0.5f * (rdz(i, k, j) + rdz(i, k, j - 1));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor23(i, k, j) = mm(i, j) * (rdy * (hat(i, k, j) - hat(i, k, j - 1)) - tmp1(i, k, j));
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      defor23(i, kts, j) = 0.0f;
      defor23(i, ktf + 1, j) = 0.0f;
    }
  if (cf.mix_full_fields) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = (v(i, k, j) - v(i, k - 1, j)) * 0.5f * (rdz(i, k, j) + rdz(i, k, j - 1));
  } else {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++)
          tmp1(i, k, j) = (v(i, k, j) - v_base(k) - v(i, k - 1, j) + v_base(k - 1)) * 0.5f * (rdz(i, k, j) + rdz(i, k, j - 1));
  }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) defor23(i, k, j) = defor23(i, k, j) + tmp1(i, k, j);
  if (!cf.periodic_x && i_start == ids + 1)
    for (int j = jts; j <= jte; j++)
      for (int k = kts; k <= kte; k++) {
        defor13(ids, k, j) = defor13(ids + 1, k, j);
        defor23(ids, k, j) = defor23(ids + 1, k, j);
      }
  if (!cf.periodic_y && j_start == jds + 1)
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= ite; i++) {
        defor13(i, k, jds) = defor13(i, k, jds + 1);
        defor23(i, k, jds) = defor23(i, k, jds + 1);
      }
  if (!cf.periodic_x && i_end == ide - 1)
    for (int j = jts; j <= jte; j++)
      for (int k = kts; k <= kte; k++) {
        defor13(ide, k, j) = defor13(ide - 1, k, j);
        defor23(ide, k, j) = defor23(ide - 1, k, j);
      }
  if (!cf.periodic_y && j_end == jde - 1)
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= ite; i++) {
        defor13(i, k, jde) = defor13(i, k, jde - 1);
        defor23(i, k, jde) = defor23(i, k, jde - 1);
      }
}

// module_diffusion_em.F:1934-2043 (prandtl = 1/3, share/module_model_constants.F)
void smag2d_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 defor11, A3 defor22, A3 defor12, A3 rdzw, A3 zx, A3 zy) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const A2 &msftx = x.msftx, &msfty = x.msfty;
  const float dx = cf.dx, dy = cf.dy;
  const bool spec = cf.specified || cf.nested;
  Loc3 l1(x);
  A3 def2 = l1.a;
  const int ktf = std::min(kte, kde - 1);
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_xe || spec) i_end = std::min(ide - 2, ite);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || spec) j_end = std::min(jde - 2, jte);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide - 1);
  const float pr = 1.f / 3.f;
  const float c_s = cf.c_s;
  float tmp, mlen_h, dxm, dym, tmpzx, tmpzy, alpha, def_limit;
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        def2(i, k, j) = 0.25f * ((defor11(i, k, j) - defor22(i, k, j)) * (defor11(i, k, j) - defor22(i, k, j)));
        tmp = 0.25f * (defor12(i, k, j) + defor12(i, k, j + 1) + defor12(i + 1, k, j) + defor12(i + 1, k, j + 1));
        def2(i, k, j) = def2(i, k, j) + tmp * tmp;
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        mlen_h = std::sqrt(dx / msftx(i, j) * dy / msfty(i, j));
        tmp = std::sqrt(def2(i, k, j));
        xkmh(i, k, j) = c_s * c_s * mlen_h * mlen_h * tmp;
        xkmh(i, k, j) = std::min(xkmh(i, k, j), 10.f * mlen_h);
        xkmv(i, k, j) = xkmh(i, k, j);
        xkhh(i, k, j) = xkmh(i, k, j) / pr;
        xkhv(i, k, j) = 0.f;
        if (cf.diff_opt == 2) {
          dxm = dx / msftx(i, j);
          dym = dy / msfty(i, j);
          tmpzx = (0.25f * (std::fabs(zx(i, k, j)) + std::fabs(zx(i + 1, k, j)) + std::fabs(zx(i, k + 1, j)) + std::fabs(zx(i + 1, k + 1, j))) *
                   rdzw(i, k, j) * dxm);
          tmpzy = (0.25f * (std::fabs(zy(i, k, j)) + std::fabs(zy(i, k, j + 1)) + std::fabs(zy(i, k + 1, j)) + std::fabs(zy(i, k + 1, j + 1))) *
                   rdzw(i, k, j) * dym);
          alpha = std::max(std::sqrt(tmpzx * tmpzx + tmpzy * tmpzy), 1.0f);
          def_limit = std::max(10.f / mlen_h, 1.e-3f);
          if (tmp > def_limit) xkmh(i, k, j) = xkmh(i, k, j) / (alpha * alpha);
          else xkmh(i, k, j) = xkmh(i, k, j) / (alpha);
          xkhh(i, k, j) = xkmh(i, k, j) / pr;
          xkmv(i, k, j) = xkmh(i, k, j);
        }
      }
}

// the thermodynamic part of phy_prep that calculate_N2 reads (module_big_step_utilities_em.F:4818-4912): th_phy, p_phy,
// t_phy = th_phy * pi_phy on the tile's mass points, and p8w, t8w at the surface (level 1) as 2-D fields
void phy_prep_thermo(const Cx &x, A3 t, const A3 *qv, A3 p, A3 pb, A3 ph, A3 phb, A3 th_phy, A3 p_phy, A3 t_phy, A2 p8w_sfc, A2 t8w_sfc) {
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1), j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int k_start = x.kts, k_end = std::min(x.kte, x.kde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        if (x.c.use_theta_m == 1 && qv) th_phy(i, k, j) = (t(i, k, j) + t0) / (1.f + r_v / r_d * (*qv)(i, k, j));
        else th_phy(i, k, j) = t(i, k, j) + t0;
        p_phy(i, k, j) = p(i, k, j) + pb(i, k, j);
        const float pi_phy = std::pow(p_phy(i, k, j) / p1000mb, rcp);
        t_phy(i, k, j) = th_phy(i, k, j) * pi_phy;
      }
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      const float zw1 = (phb(i, 1, j) + ph(i, 1, j)) / g, zw2 = (phb(i, 2, j) + ph(i, 2, j)) / g, zw3 = (phb(i, 3, j) + ph(i, 3, j)) / g;
      const float z0 = zw1;
      const float z1 = 0.5f * (zw1 + zw2);
      const float z2 = 0.5f * (zw2 + zw3);
      const float w1 = (z0 - z2) / (z1 - z2);
      const float w2 = 1.f - w1;
      p8w_sfc(i, j) = w1 * p_phy(i, 1, j) + w2 * p_phy(i, 2, j);
      t8w_sfc(i, j) = w1 * t_phy(i, 1, j) + w2 * t_phy(i, 2, j);
    }
}

// module_diffusion_em.F:1485-1712; moist species qv, qc (no ice on this path: P_QI is not allocated)
void calculate_N2(const Cx &x, A3 BN2, const A3 *qv, const A3 *qc, A3 theta, A3 t, A3 p, A2 p8w_sfc, A2 t8w_sfc, A3 rdz, A3 rdzw) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const A1 &dnw = x.dnw, &dn = x.dn;
  const float cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  const bool spec = cf.specified || cf.nested;
  const float qc_cr = 0.00001f;
  const int ktf = std::min(kte, kde - 1), ktes1 = kte - 1, ktes2 = kte - 2;
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_xe || spec) i_end = std::min(ide - 2, ite);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || spec) j_end = std::min(jde - 2, jte);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide - 1);
  Loc3 l1(x), l2(x), l3(x);
  Loc2 m1(x), m2(x);
  A3 tmp1 = l1.a, qvs = l2.a, qctmp = l3.a;
  A2 tmp1sfc = m1.a, tmp1top = m2.a;
  auto QV = [&](int i, int k, int j) -> float { return qv ? (*qv)(i, k, j) : 0.f; };   // moist(:,:,:,P_QV): the zero dummy slot without qv
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) qctmp(i, k, j) = qc ? (*qc)(i, k, j) : 0.0f;
  for (const A3 *sp : {qv, qc}) {
    if (!sp) continue;
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tmp1(i, k, j) = tmp1(i, k, j) + (*sp)(i, k, j);
    for (int j = j_start; j <= j_end; j++)
      for (int i = i_start; i <= i_end; i++) {
        tmp1sfc(i, j) = tmp1sfc(i, j) + cf1 * (*sp)(i, 1, j) + cf2 * (*sp)(i, 2, j) + cf3 * (*sp)(i, 3, j);
        tmp1top(i, j) = tmp1top(i, j) + (*sp)(i, ktes1, j) + ((*sp)(i, ktes1, j) - (*sp)(i, ktes2, j)) * 0.5f * dnw(ktes1) / dn(ktes1);
      }
  }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float tc = t(i, k, j) - svpt0;
        const float es = 1000.0f * svp1 * std::exp(svp2 * tc / (t(i, k, j) - svp3));
        qvs(i, k, j) = ep_2 * es / (p(i, k, j) - es);
      }
  float tmpdz, xlvqv, coefa, thetaep1, thetaem1;
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmpdz = 1.0f / rdz(i, k, j) + 1.0f / rdz(i, k + 1, j);
        if (QV(i, k, j) >= qvs(i, k, j) || qctmp(i, k, j) >= qc_cr) {
          xlvqv = xlv * QV(i, k, j);
          coefa = (1.0f + xlvqv / r_d / t(i, k, j)) / (1.0f + xlv * xlvqv / cp / r_v / t(i, k, j) / t(i, k, j)) / theta(i, k, j);
          thetaep1 = theta(i, k + 1, j) * (1.0f + xlv * qvs(i, k + 1, j) / cp / t(i, k + 1, j));
          thetaem1 = theta(i, k - 1, j) * (1.0f + xlv * qvs(i, k - 1, j) / cp / t(i, k - 1, j));
          BN2(i, k, j) = g * (coefa * (thetaep1 - thetaem1) / tmpdz - (tmp1(i, k + 1, j) - tmp1(i, k - 1, j)) / tmpdz);
        } else {
          BN2(i, k, j) = g * ((theta(i, k + 1, j) - theta(i, k - 1, j)) / theta(i, k, j) / tmpdz + 1.61f * (QV(i, k + 1, j) - QV(i, k - 1, j)) / tmpdz -
                              (tmp1(i, k + 1, j) - tmp1(i, k - 1, j)) / tmpdz);
        }
      }
  const int k = kts;
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      tmpdz = 1.0f / rdz(i, k + 1, j) + 0.5f / rdzw(i, k, j);
      const float thetasfc = t8w_sfc(i, j) / std::pow(p8w_sfc(i, j) / p1000mb, r_d / cp);
      if (QV(i, k, j) >= qvs(i, k, j) || qctmp(i, k, j) >= qc_cr) {
        const float qvsfc = cf1 * qvs(i, 1, j) + cf2 * qvs(i, 2, j) + cf3 * qvs(i, 3, j);
        xlvqv = xlv * QV(i, k, j);
        coefa = (1.0f + xlvqv / r_d / t(i, k, j)) / (1.0f + xlv * xlvqv / cp / r_v / t(i, k, j) / t(i, k, j)) / theta(i, k, j);
        thetaep1 = theta(i, k + 1, j) * (1.0f + xlv * qvs(i, k + 1, j) / cp / t(i, k + 1, j));
        const float thetaesfc = thetasfc * (1.0f + xlv * qvsfc / cp / t8w_sfc(i, j));
        BN2(i, k, j) = g * (coefa * (thetaep1 - thetaesfc) / tmpdz - (tmp1(i, k + 1, j) - tmp1sfc(i, j)) / tmpdz);
      } else {
        const float qvsfc = cf1 * QV(i, 1, j) + cf2 * QV(i, 2, j) + cf3 * QV(i, 3, j);
        tmpdz = 1.f / rdzw(i, k, j);
        BN2(i, k, j) = g * ((theta(i, k + 1, j) - theta(i, k, j)) / theta(i, k, j) / tmpdz + 1.61f * (QV(i, k + 1, j) - qvsfc) / tmpdz -
                            (tmp1(i, k + 1, j) - tmp1sfc(i, j)) / tmpdz);
      }
    }
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) BN2(i, ktf, j) = BN2(i, ktf - 1, j);
}

// module_diffusion_em.F:1777-1928 (prandtl = 1/3)
void smag_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 BN2, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13, A3 defor23,
             A3 rdzw) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const A2 &msftx = x.msftx, &msfty = x.msfty;
  const float dx = cf.dx, dy = cf.dy, dt = cf.dt, mix_upper_bound = cf.mix_upper_bound;
  const bool spec = cf.specified || cf.nested;
  Loc3 l1(x);
  A3 def2 = l1.a;
  const int ktf = std::min(kte, kde - 1);
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_xs || spec) i_start = std::max(ids + 1, its);
  if (cf.open_xe || spec) i_end = std::min(ide - 2, ite);
  if (cf.open_ys || spec) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || spec) j_end = std::min(jde - 2, jte);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide - 1);
  const float pr = 1.f / 3.f, c_s = cf.c_s;
  float tmp;
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        def2(i, k, j) = 0.5f * (defor11(i, k, j) * defor11(i, k, j) + defor22(i, k, j) * defor22(i, k, j) + defor33(i, k, j) * defor33(i, k, j));
        tmp = 0.25f * (defor12(i, k, j) + defor12(i, k, j + 1) + defor12(i + 1, k, j) + defor12(i + 1, k, j + 1));
        def2(i, k, j) = def2(i, k, j) + tmp * tmp;
        tmp = 0.25f * (defor13(i, k + 1, j) + defor13(i, k, j) + defor13(i + 1, k + 1, j) + defor13(i + 1, k, j));
        def2(i, k, j) = def2(i, k, j) + tmp * tmp;
        tmp = 0.25f * (defor23(i, k + 1, j) + defor23(i, k, j) + defor23(i, k + 1, j + 1) + defor23(i, k, j + 1));
        def2(i, k, j) = def2(i, k, j) + tmp * tmp;
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        tmp = std::max(0.f, def2(i, k, j) - BN2(i, k, j) / pr);
        tmp = std::pow(tmp, 0.5f);   // tmp**0.5: gfortran calls powf unless -fno-signed-zeros; glibc's powf(x, .5) is not always sqrtf(x)
        if (cf.mix_isotropic == 0) {
          const float mlen_h = std::sqrt(dx / msftx(i, j) * dy / msfty(i, j));
          const float mlen_v = 1.f / rdzw(i, k, j);
          xkmh(i, k, j) = std::max(c_s * c_s * mlen_h * mlen_h * tmp, 1.0E-6f * mlen_h * mlen_h);
          xkmh(i, k, j) = std::min(xkmh(i, k, j), mix_upper_bound * mlen_h * mlen_h / dt);
          xkmv(i, k, j) = std::max(c_s * c_s * mlen_v * mlen_v * tmp, 1.0E-6f * mlen_v * mlen_v);
          xkmv(i, k, j) = std::min(xkmv(i, k, j), mix_upper_bound * mlen_v * mlen_v / dt);
          xkhh(i, k, j) = xkmh(i, k, j) / pr;
          xkhh(i, k, j) = std::min(xkhh(i, k, j), mix_upper_bound * mlen_h * mlen_h / dt);
          xkhv(i, k, j) = xkmv(i, k, j) / pr;
          xkhv(i, k, j) = std::min(xkhv(i, k, j), mix_upper_bound * mlen_v * mlen_v / dt);
        } else {
          const float deltas = std::pow(dx / msftx(i, j) * dy / msfty(i, j) / rdzw(i, k, j), 0.33333333f);
          xkmh(i, k, j) = std::max(c_s * c_s * deltas * deltas * tmp, 1.0E-6f * deltas * deltas);
          xkmh(i, k, j) = std::min(xkmh(i, k, j), mix_upper_bound * dx / msftx(i, j) * dy / msfty(i, j) / dt);
          xkmv(i, k, j) = xkmh(i, k, j);
          xkmv(i, k, j) = std::min(xkmv(i, k, j), mix_upper_bound / rdzw(i, k, j) / rdzw(i, k, j) / dt);
          xkhh(i, k, j) = xkmh(i, k, j) / pr;
          xkhh(i, k, j) = std::min(xkhh(i, k, j), mix_upper_bound * dx / msftx(i, j) * dy / msfty(i, j) / dt);
          xkhv(i, k, j) = xkmv(i, k, j) / pr;
          xkhv(i, k, j) = std::min(xkhv(i, k, j), mix_upper_bound / rdzw(i, k, j) / rdzw(i, k, j) / dt);
        }
      }
}

}  // namespace orc
