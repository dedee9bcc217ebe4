// small_step.cpp -- oracle restatement of dyn_em/module_small_step_em.F
// (test infrastructure only; see oracle.h header).
#include "oracle.h"
namespace orc {

// dyn_em/module_small_step_em.F:16-291
void small_step_prep(const Cx &x, A3 u_1, A3 u_2, A3 v_1, A3 v_2, A3 w_1, A3 w_2, A3 t_1, A3 t_2,
                     A3 ph_1, A3 ph_2, A2 mub, A2 mu_1, A2 mu_2, A2 muu, A2 muus, A2 muv, A2 muvs,
                     A2 mut, A2 muts, A2 mudf, A3 u_save, A3 v_save, A3 w_save, A3 t_save,
                     A3 ph_save, A2 mu_save, A3 ww, A3 ww_save, A3 c2a, A3 pb, A3 p, A3 alt,
                     int rk_step) {
  const int kde = x.kde;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  const int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int k_start = x.kts, k_end = std::min(x.kte, x.kde - 1);
  const int i_endu = x.ite, j_endv = x.jte;
  const A1 &c1h = x.c1h, &c2h = x.c2h, &c1f = x.c1f, &c2f = x.c2f;

  if (rk_step == 1) {
    for (int j = j_start; j <= j_end; j++)
      for (int i = i_start; i <= i_end; i++) {
        mu_1(i, j) = mu_2(i, j);
        ww_save(i, kde, j) = 0.f;
        ww_save(i, 1, j) = 0.f;
        mudf(i, j) = 0.f;
      }
    for (int j = j_start; j <= j_end; j++)
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_endu; i++) u_1(i, k, j) = u_2(i, k, j);
    for (int j = j_start; j <= j_endv; j++)
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++) v_1(i, k, j) = v_2(i, k, j);
    for (int j = j_start; j <= j_end; j++)
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++) t_1(i, k, j) = t_2(i, k, j);
    for (int j = j_start; j <= j_end; j++)
      for (int k = k_start; k <= std::min(kde, x.kte); k++)
        for (int i = i_start; i <= i_end; i++) {
          w_1(i, k, j) = w_2(i, k, j);
          ph_1(i, k, j) = ph_2(i, k, j);
        }
    for (int j = j_start; j <= j_end; j++) {
      for (int i = i_start; i <= i_end; i++) muts(i, j) = mub(i, j) + mu_2(i, j);
      for (int i = i_start; i <= i_endu; i++) muus(i, j) = muu(i, j);
    }
    for (int j = j_start; j <= j_endv; j++)
      for (int i = i_start; i <= i_end; i++) muvs(i, j) = muv(i, j);
    for (int j = j_start; j <= j_end; j++)
      for (int i = i_start; i <= i_end; i++) {
        mu_save(i, j) = mu_2(i, j);
        mu_2(i, j) = 0.f;
      }
  } else {
    for (int j = j_start; j <= j_end; j++) {
      for (int i = i_start; i <= i_end; i++) muts(i, j) = mub(i, j) + mu_1(i, j);
      for (int i = i_start; i <= i_endu; i++)
        muus(i, j) = 0.5f * (mub(i, j) + mu_1(i, j) + mub(i - 1, j) + mu_1(i - 1, j));
    }
    for (int j = j_start; j <= j_endv; j++)
      for (int i = i_start; i <= i_end; i++)
        muvs(i, j) = 0.5f * (mub(i, j) + mu_1(i, j) + mub(i, j - 1) + mu_1(i, j - 1));
    for (int j = j_start; j <= j_end; j++)
      for (int i = i_start; i <= i_end; i++) {
        mu_save(i, j) = mu_2(i, j);
        mu_2(i, j) = mu_1(i, j) - mu_2(i, j);
      }
  }

  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) {
      ww_save(i, kde, j) = 0.f;
      ww_save(i, 1, j) = 0.f;
    }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++)
        c2a(i, k, j) = cpovcv * (pb(i, k, j) + p(i, k, j)) / alt(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_endu; i++) {
        u_save(i, k, j) = u_2(i, k, j);
        u_2(i, k, j) = ((c1h(k) * muus(i, j) + c2h(k)) * u_1(i, k, j) -
                        (c1h(k) * muu(i, j) + c2h(k)) * u_2(i, k, j)) / x.msfuy(i, j);
      }
  for (int j = j_start; j <= j_endv; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        v_save(i, k, j) = v_2(i, k, j);
        v_2(i, k, j) = ((c1h(k) * muvs(i, j) + c2h(k)) * v_1(i, k, j) -
                        (c1h(k) * muv(i, j) + c2h(k)) * v_2(i, k, j)) * x.msfvx_inv(i, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        t_save(i, k, j) = t_2(i, k, j);
        t_2(i, k, j) = (c1h(k) * muts(i, j) + c2h(k)) * t_1(i, k, j) -
                       (c1h(k) * mut(i, j) + c2h(k)) * t_2(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= kde; k++)
      for (int i = i_start; i <= i_end; i++) {
        w_save(i, k, j) = w_2(i, k, j);
        w_2(i, k, j) = ((c1f(k) * muts(i, j) + c2f(k)) * w_1(i, k, j) -
                        (c1f(k) * mut(i, j) + c2f(k)) * w_2(i, k, j)) / x.msfty(i, j);
        ph_save(i, k, j) = ph_2(i, k, j);
        ph_2(i, k, j) = ph_1(i, k, j) - ph_2(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= kde; k++)
      for (int i = i_start; i <= i_end; i++) ww_save(i, k, j) = ww(i, k, j);
}

// dyn_em/module_small_step_em.F:295-434
void small_step_finish(const Cx &x, A3 u_2, A3 u_1, A3 v_2, A3 v_1, A3 w_2, A3 w_1, A3 t_2, A3 t_1,
                       A3 ph_2, A3 ph_1, A3 ww, A3 ww1, A2 mu_2, A2 mu_1, A2 mut, A2 muts, A2 muu,
                       A2 muus, A2 muv, A2 muvs, A3 u_save, A3 v_save, A3 w_save, A3 t_save,
                       A3 ph_save, A2 mu_save, A3 h_diabatic, int number_of_small_timesteps,
                       float dts, int rk_step, int rk_order) {
  (void)u_1; (void)v_1; (void)w_1; (void)t_1; (void)ph_1; (void)mu_1;
  const int kds = x.kds, kde = x.kde;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  const int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int i_endu = x.ite, j_endv = x.jte;
  const A1 &c1h = x.c1h, &c2h = x.c2h, &c1f = x.c1f, &c2f = x.c2f;
  for (int j = j_start; j <= j_endv; j++)
    for (int k = kds; k <= kde - 1; k++)
      for (int i = i_start; i <= i_end; i++)
        v_2(i, k, j) = (x.msfvx(i, j) * v_2(i, k, j) + v_save(i, k, j) * (c1h(k) * muv(i, j) + c2h(k))) /
                       (c1h(k) * muvs(i, j) + c2h(k));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kds; k <= kde - 1; k++)
      for (int i = i_start; i <= i_endu; i++)
        u_2(i, k, j) = (x.msfuy(i, j) * u_2(i, k, j) + u_save(i, k, j) * (c1h(k) * muu(i, j) + c2h(k))) /
                       (c1h(k) * muus(i, j) + c2h(k));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kds; k <= kde; k++)
      for (int i = i_start; i <= i_end; i++) {
        w_2(i, k, j) = (x.msfty(i, j) * w_2(i, k, j) + w_save(i, k, j) * (c1f(k) * mut(i, j) + c2f(k))) /
                       (c1f(k) * muts(i, j) + c2f(k));
        ph_2(i, k, j) = ph_2(i, k, j) + ph_save(i, k, j);
        ww(i, k, j) = ww(i, k, j) + ww1(i, k, j);
      }
  if (rk_step < rk_order) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kds; k <= kde - 1; k++)
        for (int i = i_start; i <= i_end; i++)
          t_2(i, k, j) = (t_2(i, k, j) + t_save(i, k, j) * (c1h(k) * mut(i, j) + c2h(k))) /
                         (c1h(k) * muts(i, j) + c2h(k));
  } else {
    for (int j = j_start; j <= j_end; j++)
      for (int k = kds; k <= kde - 1; k++)
        for (int i = i_start; i <= i_end; i++)
          t_2(i, k, j) = (t_2(i, k, j) -
                          dts * (float)number_of_small_timesteps * (c1h(k) * mut(i, j) + c2h(k)) * h_diabatic(i, k, j) +
                          t_save(i, k, j) * (c1h(k) * mut(i, j) + c2h(k))) /
                         (c1h(k) * muts(i, j) + c2h(k));
  }
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) mu_2(i, j) = mu_2(i, j) + mu_save(i, j);
}

// dyn_em/module_small_step_em.F:438-566 (non-hydrostatic branch)
void calc_p_rho(const Cx &x, A3 al, A3 p, A3 ph, A3 alt, A3 t_2, A3 t_1, A3 c2a, A3 pm1, A2 mu,
                A2 mut, float smdiv, int step) {
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  const int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int k_start = x.kts, k_end = std::min(x.kte, x.kde - 1);
  const A1 &c1h = x.c1h, &c2h = x.c2h, &rdnw = x.rdnw;
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        al(i, k, j) = -1.f / (c1h(k) * mut(i, j) + c2h(k)) *
                      (alt(i, k, j) * (c1h(k) * mu(i, j)) + rdnw(k) * (ph(i, k + 1, j) - ph(i, k, j)));
        p(i, k, j) = c2a(i, k, j) * (alt(i, k, j) * (t_2(i, k, j) - (c1h(k) * mu(i, j)) * t_1(i, k, j)) /
                                         ((c1h(k) * mut(i, j) + c2h(k)) * (t0 + t_1(i, k, j))) -
                                     al(i, k, j));
      }
  if (step == 0) {
    for (int j = j_start; j <= j_end; j++)
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++) pm1(i, k, j) = p(i, k, j);
  } else {
    for (int j = j_start; j <= j_end; j++)
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++) {
          float ptmp = p(i, k, j);
          p(i, k, j) = p(i, k, j) + smdiv * (p(i, k, j) - pm1(i, k, j));
          pm1(i, k, j) = ptmp;
        }
  }
}

// dyn_em/module_small_step_em.F:570-650
void calc_coef_w(const Cx &x, A3 a, A3 alpha, A3 gamma, A2 mut, A3 cqw, A3 c2a, float dts,
                 float epssm) {
  const int kde = x.kde;
  const int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  const int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const A1 &c1h = x.c1h, &c2h = x.c2h, &c1f = x.c1f, &c2f = x.c2f, &rdn = x.rdn, &rdnw = x.rdnw;
  const float lid_flag = x.c.top_lid ? 0.f : 1.f;   // INTEGER lid_flag promoted to REAL in the product
  for (int j = j_start; j <= j_end; j++) {
    int k = kde - 1;
    const float cofs = (.5f * dts * g * (1.f + epssm)) * (.5f * dts * g * (1.f + epssm));
    for (int i = i_start; i <= i_end; i++) {
      a(i, 2, j) = 0.f;
      a(i, kde, j) = -2.f * cofs * (rdnw(kde - 1) * rdnw(kde - 1)) * c2a(i, kde - 1, j) * lid_flag /
                     ((c1h(k) * mut(i, j) + c2h(k)) * (c1f(k) * mut(i, j) + c2f(k)));
      gamma(i, 1, j) = 0.f;
    }
    for (int kk = 3; kk <= kde - 1; kk++) {
      k = kk - 1;
      for (int i = i_start; i <= i_end; i++)
        a(i, kk, j) = -cqw(i, kk, j) * cofs * rdn(kk) * rdnw(kk - 1) * c2a(i, kk - 1, j) /
                      ((c1h(k) * mut(i, j) + c2h(k)) * (c1f(k) * mut(i, j) + c2f(k)));
    }
    for (k = 2; k <= kde - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        float b = 1.f + cqw(i, k, j) * cofs * rdn(k) *
                            (rdnw(k) * c2a(i, k, j) / ((c1h(k) * mut(i, j) + c2h(k)) * (c1f(k) * mut(i, j) + c2f(k))) +
                             rdnw(k - 1) * c2a(i, k - 1, j) /
                                 ((c1h(k - 1) * mut(i, j) + c2h(k - 1)) * (c1f(k) * mut(i, j) + c2f(k))));
        float c = -cqw(i, k, j) * cofs * rdn(k) * rdnw(k) * c2a(i, k, j) /
                  ((c1h(k) * mut(i, j) + c2h(k)) * (c1f(k + 1) * mut(i, j) + c2f(k + 1)));
        alpha(i, k, j) = 1.f / (b - a(i, k, j) * gamma(i, k - 1, j));
        gamma(i, k, j) = c * alpha(i, k, j);
      }
    k = kde;
    for (int i = i_start; i <= i_end; i++) {
      float b = 1.f + 2.f * cofs * (rdnw(kde - 1) * rdnw(kde - 1)) * c2a(i, kde - 1, j) /
                          ((c1h(k - 1) * mut(i, j) + c2h(k - 1)) * (c1f(k) * mut(i, j) + c2f(k)));
      float c = 0.f;
      alpha(i, kde, j) = 1.f / (b - a(i, kde, j) * gamma(i, kde - 1, j));
      gamma(i, kde, j) = c * alpha(i, kde, j);
    }
  }
}

// dyn_em/module_small_step_em.F:654-965 (no polar)
void advance_uv(const Cx &x, A3 u, A3 ru_tend, A3 v, A3 rv_tend, A3 p, A3 pb, A3 ph, A3 php,
                A3 alt, A3 al, A2 mu, A2 muu, A3 cqu, A2 muv, A3 cqv, A2 mudf, float dts) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  const A1 &c1h = x.c1h, &c2h = x.c2h, &fnm = x.fnm, &fnp = x.fnp, &rdnw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy, cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3, emdiv = x.c.emdiv;
  const bool top_lid = x.c.top_lid;
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  int k_start = x.kts, k_end = x.kte - 1;
  int i_endu = ite, j_endv = jte;
  if (x.spec) {   // :734-751
    const int sz = x.c.spec_zone;
    i_start = std::max(its, ids + sz); i_end = std::min(ite, ide - sz - 1);
    j_start = std::max(jts, jds + sz); j_end = std::min(jte, jde - sz - 1);
    k_end = std::min(x.kte, kde - 1);
    i_endu = std::min(ite, ide - sz); j_endv = std::min(jte, jde - sz);
    if (x.c.periodic_x) { i_start = its; i_end = std::min(ite, ide - 1); i_endu = ite; }
  }
  int i_start_up = i_start, i_end_up = i_endu;
  int j_start_vp = j_start, j_end_vp = j_endv;
  if ((x.c.open_xs || x.c.symmetric_xs) && its == ids) i_start_up++;
  if ((x.c.open_xe || x.c.symmetric_xe) && ite == ide) i_end_up--;
  if ((x.c.open_ys || x.c.symmetric_ys) && jts == jds) j_start_vp++;
  if ((x.c.open_ye || x.c.symmetric_ye) && jte == jde) j_end_vp--;
  int i_start_u_tend = i_start, i_end_u_tend = i_endu;
  int j_start_v_tend = j_start, j_end_v_tend = j_endv;
  if (x.c.symmetric_xs && its == ids) i_start_u_tend++;
  if (x.c.symmetric_xe && ite == ide) i_end_u_tend--;
  if (x.c.symmetric_ys && jts == jds) j_start_v_tend++;
  if (x.c.symmetric_ye && jte == jde) j_end_v_tend--;
  const float dx = 1.f / rdx, dy = 1.f / rdy;
  const int ni = ite - its + 2, nk = kde + 1;
  std::vector<float> dpn_(ni * nk), dpxy_(ni * nk), mudf_xy_(ni);
  auto dpn = [&](int i, int k) -> float & { return dpn_[(i - its) + (k - 1) * ni]; };
  auto dpxy = [&](int i, int k) -> float & { return dpxy_[(i - its) + (k - 1) * ni]; };
  auto mudf_xy = [&](int i) -> float & { return mudf_xy_[i - its]; };

  for (int j = j_start; j <= j_end; j++) {
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start_u_tend; i <= i_end_u_tend; i++) u(i, k, j) = u(i, k, j) + dts * ru_tend(i, k, j);
    for (int i = i_start_up; i <= i_end_up; i++)
      mudf_xy(i) = -emdiv * dx * (mudf(i, j) - mudf(i - 1, j)) / x.msfuy(i, j);
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start_up; i <= i_end_up; i++)
        dpxy(i, k) = (x.msfux(i, j) / x.msfuy(i, j)) * .5f * rdx * (c1h(k) * muu(i, j) + c2h(k)) *
                     (((ph(i, k + 1, j) - ph(i - 1, k + 1, j)) + (ph(i, k, j) - ph(i - 1, k, j))) +
                      (alt(i, k, j) + alt(i - 1, k, j)) * (p(i, k, j) - p(i - 1, k, j)) +
                      (al(i, k, j) + al(i - 1, k, j)) * (pb(i, k, j) - pb(i - 1, k, j)));
    // non_hydrostatic
    for (int i = i_start_up; i <= i_end_up; i++) {
      dpn(i, 1) = .5f * (cf1 * (p(i, 1, j) + p(i - 1, 1, j)) + cf2 * (p(i, 2, j) + p(i - 1, 2, j)) +
                         cf3 * (p(i, 3, j) + p(i - 1, 3, j)));
      dpn(i, kde) = 0.f;
    }
    if (top_lid)
      for (int i = i_start_up; i <= i_end_up; i++)
        dpn(i, kde) = .5f * (cf1 * (p(i - 1, kde - 1, j) + p(i, kde - 1, j)) +
                             cf2 * (p(i - 1, kde - 2, j) + p(i, kde - 2, j)) +
                             cf3 * (p(i - 1, kde - 3, j) + p(i, kde - 3, j)));
    for (int k = k_start + 1; k <= k_end; k++)
      for (int i = i_start_up; i <= i_end_up; i++)
        dpn(i, k) = .5f * (fnm(k) * (p(i, k, j) + p(i - 1, k, j)) + fnp(k) * (p(i, k - 1, j) + p(i - 1, k - 1, j)));
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start_up; i <= i_end_up; i++)
        dpxy(i, k) = dpxy(i, k) + (x.msfux(i, j) / x.msfuy(i, j)) * rdx * (php(i, k, j) - php(i - 1, k, j)) *
                                      (rdnw(k) * (dpn(i, k + 1) - dpn(i, k)) -
                                       .5f * ((c1h(k) * mu(i - 1, j)) + (c1h(k) * mu(i, j))));
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start_up; i <= i_end_up; i++)
        u(i, k, j) = u(i, k, j) - dts * cqu(i, k, j) * dpxy(i, k) + (c1h(k) * mudf_xy(i));
  }

  for (int j = j_start_v_tend; j <= j_end_v_tend; j++) {
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) v(i, k, j) = v(i, k, j) + dts * rv_tend(i, k, j);
    for (int i = i_start; i <= i_end; i++)
      mudf_xy(i) = -emdiv * dy * (mudf(i, j) - mudf(i, j - 1)) * x.msfvx_inv(i, j);
    if (j >= j_start_vp && j <= j_end_vp) {
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          dpxy(i, k) = (x.msfvy(i, j) / x.msfvx(i, j)) * .5f * rdy * (c1h(k) * muv(i, j) + c2h(k)) *
                       (((ph(i, k + 1, j) - ph(i, k + 1, j - 1)) + (ph(i, k, j) - ph(i, k, j - 1))) +
                        (alt(i, k, j) + alt(i, k, j - 1)) * (p(i, k, j) - p(i, k, j - 1)) +
                        (al(i, k, j) + al(i, k, j - 1)) * (pb(i, k, j) - pb(i, k, j - 1)));
      for (int i = i_start; i <= i_end; i++) {
        dpn(i, 1) = .5f * (cf1 * (p(i, 1, j) + p(i, 1, j - 1)) + cf2 * (p(i, 2, j) + p(i, 2, j - 1)) +
                           cf3 * (p(i, 3, j) + p(i, 3, j - 1)));
        dpn(i, kde) = 0.f;
      }
      if (top_lid)
        for (int i = i_start; i <= i_end; i++)
          dpn(i, kde) = .5f * (cf1 * (p(i, kde - 1, j - 1) + p(i, kde - 1, j)) +
                               cf2 * (p(i, kde - 2, j - 1) + p(i, kde - 2, j)) +
                               cf3 * (p(i, kde - 3, j - 1) + p(i, kde - 3, j)));
      for (int k = k_start + 1; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          dpn(i, k) = .5f * (fnm(k) * (p(i, k, j) + p(i, k, j - 1)) + fnp(k) * (p(i, k - 1, j) + p(i, k - 1, j - 1)));
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          dpxy(i, k) = dpxy(i, k) + (x.msfvy(i, j) / x.msfvx(i, j)) * rdy * (php(i, k, j) - php(i, k, j - 1)) *
                                        (rdnw(k) * (dpn(i, k + 1) - dpn(i, k)) -
                                         .5f * ((c1h(k) * mu(i, j - 1)) + (c1h(k) * mu(i, j))));
      for (int k = k_start; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          v(i, k, j) = v(i, k, j) - dts * cqv(i, k, j) * dpxy(i, k) + (c1h(k) * mudf_xy(i));
    }
  }
}

// dyn_em/module_small_step_em.F:969-1174
void advance_mu_t(const Cx &x, A3 ww, A3 ww_1, A3 u, A3 u_1, A3 v, A3 v_1, A2 mu, A2 mut,
                  A2 muave, A2 muts, A2 muu, A2 muv, A2 mudf, A3 t, A3 t_1, A3 t_ave, A3 ft,
                  A2 mu_tend, float dts, float epssm) {
  const int kde = x.kde;
  const int its = x.its;
  int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int k_start = x.kts, k_end = x.kte - 1;
  if (x.spec) {   // :1054-1063
    if (!x.c.periodic_x) { i_start = std::max(x.its, x.ids + 1); i_end = std::min(x.ite, x.ide - 2); }
    j_start = std::max(x.jts, x.jds + 1); j_end = std::min(x.jte, x.jde - 2);
  }
  const A1 &c1h = x.c1h, &c2h = x.c2h, &dnw = x.dnw, &fnm = x.fnm, &fnp = x.fnp, &rdnw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy;
  const int ni = x.ite - its + 2, nk = kde + 1;
  std::vector<float> wdtn_(ni * nk), dvdxi_(ni * nk), dmdt_(ni);
  auto wdtn = [&](int i, int k) -> float & { return wdtn_[(i - its) + (k - 1) * ni]; };
  auto dvdxi = [&](int i, int k) -> float & { return dvdxi_[(i - its) + (k - 1) * ni]; };
  auto dmdt = [&](int i) -> float & { return dmdt_[i - its]; };

  for (int j = j_start; j <= j_end; j++) {
    for (int i = i_start; i <= i_end; i++) dmdt(i) = 0.f;
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        dvdxi(i, k) = x.msftx(i, j) * x.msfty(i, j) *
                      (rdy * ((v(i, k, j + 1) + (c1h(k) * muv(i, j + 1) + c2h(k)) * v_1(i, k, j + 1) * x.msfvx_inv(i, j + 1)) -
                              (v(i, k, j) + (c1h(k) * muv(i, j) + c2h(k)) * v_1(i, k, j) * x.msfvx_inv(i, j))) +
                       rdx * ((u(i + 1, k, j) + (c1h(k) * muu(i + 1, j) + c2h(k)) * u_1(i + 1, k, j) / x.msfuy(i + 1, j)) -
                              (u(i, k, j) + (c1h(k) * muu(i, j) + c2h(k)) * u_1(i, k, j) / x.msfuy(i, j))));
        dmdt(i) = dmdt(i) + dnw(k) * dvdxi(i, k);
      }
    for (int i = i_start; i <= i_end; i++) {
      muave(i, j) = mu(i, j);
      mu(i, j) = mu(i, j) + dts * (dmdt(i) + mu_tend(i, j));
      mudf(i, j) = (dmdt(i) + mu_tend(i, j));
      muts(i, j) = mut(i, j) + mu(i, j);
      muave(i, j) = .5f * ((1.f + epssm) * mu(i, j) + (1.f - epssm) * muave(i, j));
    }
    for (int kk = 2; kk <= k_end; kk++) {
      int k = kk - 1;
      for (int i = i_start; i <= i_end; i++)
        ww(i, kk, j) = ww(i, kk - 1, j) -
                       dnw(kk - 1) * ((c1h(k) * dmdt(i)) + dvdxi(i, kk - 1) + (c1h(k) * mu_tend(i, j))) / x.msfty(i, j);
    }
    for (int k = 1; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) ww(i, k, j) = ww(i, k, j) - ww_1(i, k, j);
  }

  for (int j = j_start; j <= j_end; j++)
    for (int k = 1; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        t_ave(i, k, j) = t(i, k, j);
        t(i, k, j) = t(i, k, j) + x.msfty(i, j) * dts * ft(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++) {
    for (int i = i_start; i <= i_end; i++) {
      wdtn(i, 1) = 0.f;
      wdtn(i, kde) = 0.f;
    }
    for (int k = 2; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++)
        wdtn(i, k) = ww(i, k, j) * (fnm(k) * t_1(i, k, j) + fnp(k) * t_1(i, k - 1, j));
    for (int k = 1; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++)
        t(i, k, j) = t(i, k, j) -
                     dts * x.msfty(i, j) *
                         (x.msftx(i, j) * (.5f * rdy * (v(i, k, j + 1) * (t_1(i, k, j + 1) + t_1(i, k, j)) -
                                                        v(i, k, j) * (t_1(i, k, j) + t_1(i, k, j - 1))) +
                                           .5f * rdx * (u(i + 1, k, j) * (t_1(i + 1, k, j) + t_1(i, k, j)) -
                                                        u(i, k, j) * (t_1(i, k, j) + t_1(i - 1, k, j)))) +
                          rdnw(k) * (wdtn(i, k + 1) - wdtn(i, k)));
  }
}

// dyn_em/module_small_step_em.F:1178-1469
void advance_w(const Cx &x, A3 w, A3 rw_tend, A3 ww, A3 w_save, A3 u, A3 v, A2 mu1, A2 mut,
               A2 muave, A2 muts, A3 t_2ave, A3 t_2, A3 t_1, A3 ph, A3 ph_1, A3 phb, A3 ph_tend,
               A3 c2a, A3 cqw, A3 alt, A3 a, A3 alpha, A3 gamma, float dts, float epssm) {
  (void)mu1;
  const int kde = x.kde, its = x.its;
  int i_start = x.its, i_end = std::min(x.ite, x.ide - 1);
  int j_start = x.jts, j_end = std::min(x.jte, x.jde - 1);
  const int k_end = x.kte - 1;
  if (x.spec) {   // :1273-1282
    if (!x.c.periodic_x) { i_start = std::max(x.its, x.ids + 1); i_end = std::min(x.ite, x.ide - 2); }
    j_start = std::max(x.jts, x.jds + 1); j_end = std::min(x.jte, x.jde - 2);
  }
  const A1 &c1h = x.c1h, &c2h = x.c2h, &c1f = x.c1f, &c2f = x.c2f;
  const A1 &fnm = x.fnm, &fnp = x.fnp, &rdnw = x.rdnw, &rdn = x.rdn;
  const float rdx = x.rdx, rdy = x.rdy, cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  const bool top_lid = x.c.top_lid;
  const A2 &ht = x.ht, &msftx = x.msftx, &msfty = x.msfty;
  const int ni = x.ite - its + 2, nk = kde + 1;
  std::vector<float> rhs_(ni * nk), wdwn_(ni * nk), msft_inv_(ni);
  auto rhs = [&](int i, int k) -> float & { return rhs_[(i - its) + (k - 1) * ni]; };
  auto wdwn = [&](int i, int k) -> float & { return wdwn_[(i - its) + (k - 1) * ni]; };
  auto msft_inv = [&](int i) -> float & { return msft_inv_[i - its]; };
  const float pi = 4.f * std::atan(1.f);
  const float dampmag = dts * x.c.dampcoef;
  const float hdepth = x.c.zdamp;

  for (int i = i_start; i <= i_end; i++) rhs(i, 1) = 0.f;
  for (int j = j_start; j <= j_end; j++) {
    for (int i = i_start; i <= i_end; i++) msft_inv(i) = 1.f / msfty(i, j);
    for (int k = 1; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        t_2ave(i, k, j) = .5f * ((1.f + epssm) * t_2(i, k, j) + (1.f - epssm) * t_2ave(i, k, j));
        t_2ave(i, k, j) = (t_2ave(i, k, j) + (c1h(k) * muave(i, j)) * t0) /
                          ((c1h(k) * muts(i, j) + c2h(k)) * (t0 + t_1(i, k, j)));
        rhs(i, k + 1) = dts * (ph_tend(i, k + 1, j) + .5f * g * (1.f - epssm) * w(i, k + 1, j));
      }
    if (x.c.phi_adv_z == 2) {
      for (int k = 1; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          wdwn(i, k + 1) = rdnw(k) * (ph_1(i, k + 1, j) - ph_1(i, k, j) + phb(i, k + 1, j) - phb(i, k, j));
      for (int k = 2; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          rhs(i, k) = rhs(i, k) - dts * ww(i, k, j) * (fnm(k) * wdwn(i, k + 1) + fnp(k) * wdwn(i, k));
    } else {
      for (int k = 1; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          wdwn(i, k + 1) = .5f * (ww(i, k + 1, j) + ww(i, k, j)) * rdnw(k) *
                           (ph_1(i, k + 1, j) - ph_1(i, k, j) + phb(i, k + 1, j) - phb(i, k, j));
      for (int k = 2; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++)
          rhs(i, k) = rhs(i, k) - dts * (fnm(k) * wdwn(i, k + 1) + fnp(k) * wdwn(i, k));
    }
    for (int k = 2; k <= k_end + 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        rhs(i, k) = ph(i, k, j) + msfty(i, j) * rhs(i, k) / (c1f(k) * mut(i, j) + c2f(k));
        if (top_lid && k == k_end + 1) rhs(i, k) = 0.f;
      }
    for (int i = i_start; i <= i_end; i++)
      w(i, 1, j) = msfty(i, j) * .5f * rdy *
                       ((ht(i, j + 1) - ht(i, j)) * (cf1 * v(i, 1, j + 1) + cf2 * v(i, 2, j + 1) + cf3 * v(i, 3, j + 1)) +
                        (ht(i, j) - ht(i, j - 1)) * (cf1 * v(i, 1, j) + cf2 * v(i, 2, j) + cf3 * v(i, 3, j))) +
                   msftx(i, j) * .5f * rdx *
                       ((ht(i + 1, j) - ht(i, j)) * (cf1 * u(i + 1, 1, j) + cf2 * u(i + 1, 2, j) + cf3 * u(i + 1, 3, j)) +
                        (ht(i, j) - ht(i - 1, j)) * (cf1 * u(i, 1, j) + cf2 * u(i, 2, j) + cf3 * u(i, 3, j)));
    for (int k = 2; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++)
        w(i, k, j) = w(i, k, j) + dts * rw_tend(i, k, j) +
                     msft_inv(i) * cqw(i, k, j) *
                         (+.5f * dts * g * rdn(k) *
                          (c2a(i, k, j) * rdnw(k) / (c1h(k) * mut(i, j) + c2h(k)) *
                               ((1.f + epssm) * (rhs(i, k + 1) - rhs(i, k)) + (1.f - epssm) * (ph(i, k + 1, j) - ph(i, k, j))) -
                           c2a(i, k - 1, j) * rdnw(k - 1) / (c1h(k - 1) * mut(i, j) + c2h(k - 1)) *
                               ((1.f + epssm) * (rhs(i, k) - rhs(i, k - 1)) + (1.f - epssm) * (ph(i, k, j) - ph(i, k - 1, j))))) +
                     dts * g * msft_inv(i) *
                         (rdn(k) * (c2a(i, k, j) * alt(i, k, j) * t_2ave(i, k, j) -
                                    c2a(i, k - 1, j) * alt(i, k - 1, j) * t_2ave(i, k - 1, j)) -
                          (c1f(k) * muave(i, j)));
    {
      int k = k_end + 1;
      for (int i = i_start; i <= i_end; i++) {
        w(i, k, j) = w(i, k, j) + dts * rw_tend(i, k, j) +
                     msft_inv(i) *
                         (-.5f * dts * g / (c1h(k - 1) * mut(i, j) + c2h(k - 1)) * (rdnw(k - 1) * rdnw(k - 1)) * 2.f *
                              c2a(i, k - 1, j) *
                              ((1.f + epssm) * (rhs(i, k) - rhs(i, k - 1)) + (1.f - epssm) * (ph(i, k, j) - ph(i, k - 1, j))) -
                          dts * g * (2.f * rdnw(k - 1) * c2a(i, k - 1, j) * alt(i, k - 1, j) * t_2ave(i, k - 1, j) +
                                     (c1f(k) * muave(i, j))));
        if (top_lid) w(i, k, j) = 0.f;
      }
    }
    for (int k = 2; k <= k_end + 1; k++)
      for (int i = i_start; i <= i_end; i++) w(i, k, j) = (w(i, k, j) - a(i, k, j) * w(i, k - 1, j)) * alpha(i, k, j);
    for (int k = k_end; k >= 2; k--)
      for (int i = i_start; i <= i_end; i++) w(i, k, j) = w(i, k, j) - gamma(i, k, j) * w(i, k + 1, j);
    if (x.c.damp_opt == 3) {
      for (int k = k_end + 1; k >= 2; k--)
        for (int i = i_start; i <= i_end; i++) {
          float htop = (ph_1(i, k_end + 1, j) + phb(i, k_end + 1, j)) / g;
          float hk = (ph_1(i, k, j) + phb(i, k, j)) / g;
          float hbot = htop - hdepth;
          float dampwt = 0.f;
          if (hk >= hbot)
            dampwt = dampmag * std::sin(0.5f * pi * (hk - hbot) / hdepth) * std::sin(0.5f * pi * (hk - hbot) / hdepth);
          w(i, k, j) = (w(i, k, j) - dampwt * (c1f(k) * mut(i, j) + c2f(k)) * w_save(i, k, j)) / (1.f + dampwt);
        }
    }
    for (int k = k_end + 1; k >= 2; k--)
      for (int i = i_start; i <= i_end; i++)
        ph(i, k, j) = rhs(i, k) + msfty(i, j) * .5f * dts * g * (1.f + epssm) * w(i, k, j) / (c1f(k) * muts(i, j) + c2f(k));
  }
}

// dyn_em/module_small_step_em.F:1473-1633
void sumflux(const Cx &x, A3 ru, A3 rv, A3 ww, A3 u_lin, A3 v_lin, A3 ww_lin, A2 muu, A2 muv,
             A3 ru_m, A3 rv_m, A3 ww_m, int iteration, int number_of_small_timesteps) {
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const A1 &c1h = x.c1h, &c2h = x.c2h;
  const float nst = (float)number_of_small_timesteps;
  if (iteration == 1)
    for (int j = jts; j <= jte; j++)
      for (int k = kts; k <= kte; k++)
        for (int i = its; i <= ite; i++) {
          ru_m(i, k, j) = 0.f;
          rv_m(i, k, j) = 0.f;
          ww_m(i, k, j) = 0.f;
        }
  const int mini = std::min(x.ide - 1, ite), minj = std::min(x.jde - 1, jte), mink = std::min(x.kde - 1, kte);
  for (int j = jts; j <= minj; j++)
    for (int k = kts; k <= mink; k++)
      for (int i = its; i <= mini; i++) {
        ru_m(i, k, j) = ru_m(i, k, j) + ru(i, k, j);
        rv_m(i, k, j) = rv_m(i, k, j) + rv(i, k, j);
        ww_m(i, k, j) = ww_m(i, k, j) + ww(i, k, j);
      }
  if (ite > mini)
    for (int j = jts; j <= minj; j++)
      for (int k = kts; k <= mink; k++)
        for (int i = mini + 1; i <= ite; i++) ru_m(i, k, j) = ru_m(i, k, j) + ru(i, k, j);
  if (jte > minj)
    for (int j = minj + 1; j <= jte; j++)
      for (int k = kts; k <= mink; k++)
        for (int i = its; i <= mini; i++) rv_m(i, k, j) = rv_m(i, k, j) + rv(i, k, j);
  if (kte > mink)
    for (int j = jts; j <= minj; j++)
      for (int k = mink + 1; k <= kte; k++)
        for (int i = its; i <= mini; i++) ww_m(i, k, j) = ww_m(i, k, j) + ww(i, k, j);
  if (iteration == number_of_small_timesteps) {
    for (int j = jts; j <= minj; j++)
      for (int k = kts; k <= mink; k++)
        for (int i = its; i <= mini; i++) {
          ru_m(i, k, j) = ru_m(i, k, j) / nst + (c1h(k) * muu(i, j) + c2h(k)) * u_lin(i, k, j) / x.msfuy(i, j);
          rv_m(i, k, j) = rv_m(i, k, j) / nst + (c1h(k) * muv(i, j) + c2h(k)) * v_lin(i, k, j) * x.msfvx_inv(i, j);
          ww_m(i, k, j) = ww_m(i, k, j) / nst + ww_lin(i, k, j);
        }
    if (ite > mini)
      for (int j = jts; j <= minj; j++)
        for (int k = kts; k <= mink; k++)
          for (int i = mini + 1; i <= ite; i++)
            ru_m(i, k, j) = ru_m(i, k, j) / nst + (c1h(k) * muu(i, j) + c2h(k)) * u_lin(i, k, j) / x.msfuy(i, j);
    if (jte > minj)
      for (int j = minj + 1; j <= jte; j++)
        for (int k = kts; k <= mink; k++)
          for (int i = its; i <= mini; i++)
            rv_m(i, k, j) = rv_m(i, k, j) / nst + (c1h(k) * muv(i, j) + c2h(k)) * v_lin(i, k, j) * x.msfvx_inv(i, j);
    if (kte > mink)
      for (int j = jts; j <= minj; j++)
        for (int k = mink + 1; k <= kte; k++)
          for (int i = its; i <= mini; i++) ww_m(i, k, j) = ww_m(i, k, j) / nst + ww_lin(i, k, j);
  }
}

}  // namespace orc
