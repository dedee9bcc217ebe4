// tke.cpp -- the 1.5-order TKE closure (km_opt = 2 with diff_opt = 2): tke_km
// (dyn_em/module_diffusion_em.F:2049-2335, the km_opt = 2 branches), calc_l_scale (:2341-2405), tke_rhs (:6099-6228) with
// tke_shear (:6529-6876), tke_buoyancy (:6234-6378) and tke_dissip (:6384-6523), bound_tke (module_em.F:2616-2646), on the
// thermodynamic fields of phy_prep (module_big_step_utilities_em.F:4818-4912).
// TEST INFRASTRUCTURE (see oracle.h).
#include "oracle.h"

namespace orc {

namespace {
struct Loc3 {
  std::vector<float> v; A3 a;
  Loc3(const Cx &x) {
    const long ni = x.ite - x.its + 3, nk = x.kte - x.kts + 1, nj = x.jte - x.jts + 3;
    v.assign((size_t)(ni * nk * nj), 0.f);
    a.p = v.data(); a.ims = x.its - 1; a.kms = x.kts; a.jms = x.jts - 1; a.sk = ni; a.sj = ni * nk;
  }
};
struct R4 { int i_start, i_end, j_start, j_end; };
R4 mass_range(const Cx &x, bool open_restricted) {
  const wrfgpu_config &cf = x.c;
  const bool spec = cf.specified || cf.nested;
  R4 r{x.its, std::min(x.ite, x.ide - 1), x.jts, std::min(x.jte, x.jde - 1)};
  if (open_restricted) {
    if (cf.open_xs || spec) r.i_start = std::max(x.ids + 1, x.its);
    if (cf.open_xe || spec) r.i_end = std::min(x.ide - 2, x.ite);
    if (cf.open_ys || spec) r.j_start = std::max(x.jds + 1, x.jts);
    if (cf.open_ye || spec) r.j_end = std::min(x.jde - 2, x.jte);
  }
  if (cf.periodic_x) { r.i_start = x.its; r.i_end = std::min(x.ite, x.ide - 1); }
  return r;
}
}  // namespace

// p8w, t8w at the model top (level kde) as phy_prep extrapolates them (module_big_step_utilities_em.F:4904-4911)
void phy_prep_top(const Cx &x, A3 ph, A3 phb, A3 p_phy, A3 t_phy, A2 p8w_top, A2 t8w_top) {
  const int kte = x.kte, kde = x.kde, k_end = std::min(x.kte, x.kde - 1);
  for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
    for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) {
      auto zw = [&](int k) { return (phb(i, k, j) + ph(i, k, j)) / g; };
      const float z0 = zw(kte);
      const float z1 = 0.5f * (zw(k_end) + zw(k_end + 1));
      const float z2 = 0.5f * (zw(k_end - 1) + zw(k_end));
      const float w1 = (z0 - z2) / (z1 - z2);
      const float w2 = 1.f - w1;
      p8w_top(i, j) = std::exp(w1 * std::log(p_phy(i, kde - 1, j)) + w2 * std::log(p_phy(i, kde - 2, j)));
      t8w_top(i, j) = w1 * t_phy(i, kde - 1, j) + w2 * t_phy(i, kde - 2, j);
    }
}

// :2341-2405
static void calc_l_scale(const Cx &x, A3 tke, A3 BN2, A3 l_scale, const R4 &r, int ktf, A3 rdzw) {
  const float dx = x.c.dx, dy = x.c.dy;
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float deltas = std::pow(dx / x.msftx(i, j) * dy / x.msfty(i, j) / rdzw(i, k, j), 0.33333333f);
        l_scale(i, k, j) = deltas;
        if (BN2(i, k, j) > 1.0e-6f) {
          const float tmp = std::sqrt(std::max(tke(i, k, j), 1.0e-6f));
          l_scale(i, k, j) = 0.76f * tmp / std::sqrt(BN2(i, k, j));
          l_scale(i, k, j) = std::min(l_scale(i, k, j), deltas);
          l_scale(i, k, j) = std::max(l_scale(i, k, j), 0.001f * deltas);
        }
      }
}

// :2049-2335, km_opt = 2 (bl_pbl_physics = 0)
void tke_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 BN2, A3 tke, A2 p8w_sfc, A2 t8w_sfc, A2 p8w_top, A2 t8w_top, A3 theta, A3 rdz,
            A3 rdzw) {
  const wrfgpu_config &cf = x.c;
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const float dx = cf.dx, dy = cf.dy, dt = cf.dt, mix_upper_bound = cf.mix_upper_bound, c_k = cf.c_k;
  const float prandtl = 1.f / 3.0f, epsilon = 1.e-10f, tke_seed_value = 1.e-06f;
  const R4 r = mass_range(x, true);
  Loc3 l1(x), l2(x);
  A3 dthrdn = l1.a, l_scale = l2.a;
  float tke_seed = 0.f;
  if (cf.isfflx == 0) {
    if (cf.diff_opt == 2) {
      if (cf.tke_drag_coefficient < epsilon && cf.tke_heat_flux < epsilon) tke_seed = tke_seed_value;
    } else tke_seed = tke_seed_value;
  }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int k = kts + 1; k <= ktf - 1; k++)
      for (int i = r.i_start; i <= r.i_end; i++) {
        const float tmpdz = 1.0f / rdz(i, k + 1, j) + 1.0f / rdz(i, k, j);
        dthrdn(i, k, j) = (theta(i, k + 1, j) - theta(i, k - 1, j)) / tmpdz;
      }
  for (int j = r.j_start; j <= r.j_end; j++)
    for (int i = r.i_start; i <= r.i_end; i++) {
      int k = kts;
      float tmpdz = 1.0f / rdzw(i, k + 1, j) + 1.0f / rdzw(i, k, j);
      const float thetasfc = t8w_sfc(i, j) / std::pow(p8w_sfc(i, j) / p1000mb, r_d / cp);
      dthrdn(i, k, j) = (theta(i, k + 1, j) - thetasfc) / tmpdz;
      k = ktf;
      tmpdz = 1.0f / rdz(i, k, j) + 0.5f / rdzw(i, k, j);
      const float thetatop = t8w_top(i, j) / std::pow(p8w_top(i, j) / p1000mb, r_d / cp);
      dthrdn(i, k, j) = (thetatop - theta(i, k - 1, j)) / tmpdz;
    }
  if (cf.mix_isotropic == 0) {
    for (int j = r.j_start; j <= r.j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = r.i_start; i <= r.i_end; i++) {
          const float mlen_h = std::sqrt(dx / x.msftx(i, j) * dy / x.msfty(i, j));
          const float tmp = std::sqrt(std::max(tke(i, k, j), tke_seed));
          const float deltas = 1.0f / rdzw(i, k, j);
          float mlen_v = deltas;
          if (dthrdn(i, k, j) > 0.f) {
            const float mlen_s = 0.76f * tmp / std::pow(std::fabs(g / theta(i, k, j) * dthrdn(i, k, j)), 0.5f);   // (...)**0.5 = powf, see smag_km
            mlen_v = std::min(mlen_v, mlen_s);
          }
          xkmh(i, k, j) = std::max(c_k * tmp * mlen_h, 1.0E-6f * mlen_h * mlen_h);
          xkmh(i, k, j) = std::min(xkmh(i, k, j), mix_upper_bound * mlen_h * mlen_h / dt);
          xkmv(i, k, j) = std::max(c_k * tmp * mlen_v, 1.0E-6f * deltas * deltas);
          xkmv(i, k, j) = std::min(xkmv(i, k, j), mix_upper_bound * deltas * deltas / dt);
          const float pr_inv_h = 1.f / prandtl;
          const float pr_inv_v = 1.0f + 2.0f * mlen_v / deltas;
          xkhh(i, k, j) = xkmh(i, k, j) * pr_inv_h;
          xkhv(i, k, j) = xkmv(i, k, j) * pr_inv_v;
        }
  } else {
    calc_l_scale(x, tke, BN2, l_scale, r, ktf, rdzw);
    for (int j = r.j_start; j <= r.j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = r.i_start; i <= r.i_end; i++) {
          const float tmp = std::sqrt(std::max(tke(i, k, j), tke_seed));
          const float deltas = std::pow(dx / x.msftx(i, j) * dy / x.msfty(i, j) / rdzw(i, k, j), 0.33333333f);
          xkmh(i, k, j) = c_k * tmp * l_scale(i, k, j);
          xkmh(i, k, j) = std::min(mix_upper_bound * dx / x.msftx(i, j) * dy / x.msfty(i, j) / dt, xkmh(i, k, j));
          xkmv(i, k, j) = c_k * tmp * l_scale(i, k, j);
          xkmv(i, k, j) = std::min(mix_upper_bound / rdzw(i, k, j) / rdzw(i, k, j) / dt, xkmv(i, k, j));
          const float pr_inv = 1.0f + 2.0f * l_scale(i, k, j) / deltas;
          xkhh(i, k, j) = std::min(mix_upper_bound * dx / x.msftx(i, j) * dy / x.msfty(i, j) / dt, xkmh(i, k, j) * pr_inv);
          xkhv(i, k, j) = std::min(mix_upper_bound / rdzw(i, k, j) / rdzw(i, k, j) / dt, xkmv(i, k, j) * pr_inv);
        }
  }
}

// :6099-6228 tke_rhs = tke_shear + tke_buoyancy + tke_dissip + the limiter of the sink
void tke_rhs(const Cx &x, A3 tendency, A3 BN2, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13, A3 defor23, A3 u, A3 v, A3 tke, A2 mu,
             A3 theta, A3 xkmh, A3 xkmv, A3 xkhv, A3 rdzw, A2 hfx, const A3 *qv, A2 ust, A3 rho) {
  const wrfgpu_config &cf = x.c;
  const A1 &c1 = x.c1h, &c2 = x.c2h;
  const int kts = x.kts, ktf = std::min(x.kte, x.kde - 1);
  const float epsilon = 1.e-10f, dt = cf.dt;
  Loc3 l1(x), l2(x), l3(x);
  A3 avg = l1.a, tmp2 = l2.a, l_scale = l3.a;
  // ---- tke_shear (:6529-6876)
  {
    const R4 r = mass_range(x, true);
    const int i_start = r.i_start, i_end = r.i_end, j_start = r.j_start, j_end = r.j_end;
#define LOOP3(K0, K1) for (int j = j_start; j <= j_end; j++) for (int k = (K0); k <= (K1); k++) for (int i = i_start; i <= i_end; i++)
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + 0.5f * (c1(k) * mu(i, j) + c2(k)) * xkmh(i, k, j) * ((defor11(i, k, j)) * (defor11(i, k, j)));
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + 0.5f * (c1(k) * mu(i, j) + c2(k)) * xkmh(i, k, j) * ((defor22(i, k, j)) * (defor22(i, k, j)));
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + 0.5f * (c1(k) * mu(i, j) + c2(k)) * xkmv(i, k, j) * ((defor33(i, k, j)) * (defor33(i, k, j)));
    LOOP3(kts, ktf)
      avg(i, k, j) = 0.25f * ((defor12(i, k, j) * defor12(i, k, j)) + (defor12(i, k, j + 1) * defor12(i, k, j + 1)) + (defor12(i + 1, k, j) * defor12(i + 1, k, j)) +
                              (defor12(i + 1, k, j + 1) * defor12(i + 1, k, j + 1)));
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + (c1(k) * mu(i, j) + c2(k)) * xkmh(i, k, j) * avg(i, k, j);
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end + 1; i++) tmp2(i, k, j) = defor13(i, k, j);
    for (int j = j_start; j <= j_end; j++)
      for (int i = i_start; i <= i_end + 1; i++) { tmp2(i, kts, j) = 0.0f; tmp2(i, ktf + 1, j) = 0.0f; }
    LOOP3(kts, ktf)
      avg(i, k, j) = 0.25f * ((tmp2(i, k + 1, j) * tmp2(i, k + 1, j)) + (tmp2(i, k, j) * tmp2(i, k, j)) + (tmp2(i + 1, k + 1, j) * tmp2(i + 1, k + 1, j)) +
                              (tmp2(i + 1, k, j) * tmp2(i + 1, k, j)));
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + (c1(k) * mu(i, j) + c2(k)) * xkmv(i, k, j) * avg(i, k, j);
    auto absU_ = [&](int i, int k, int j) {
      const float a = u(i, k, j) + u(i + 1, k, j), b = v(i, k, j) + v(i, k, j + 1);
      return 0.5f * std::sqrt(a * a + b * b);
    };
    {
      const int k = kts;
      for (int j = j_start; j <= j_end; j++)
        for (int i = i_start; i <= i_end; i++) {
          float absU, Cd;
          if (cf.isfflx == 0) { absU = absU_(i, k, j); Cd = cf.tke_drag_coefficient; }
          else { absU = absU_(i, k, j) + epsilon; Cd = (ust(i, j) * ust(i, j)) / (absU * absU); }
          tendency(i, k, j) = tendency(i, k, j) + (c1(k) * mu(i, j) + c2(k)) * ((u(i, k, j) + u(i + 1, k, j)) * 0.5f * Cd * absU *
                                                                               (defor13(i, kts + 1, j) + defor13(i + 1, kts + 1, j)) * 0.5f);
        }
    }
    for (int j = j_start; j <= j_end + 1; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) tmp2(i, k, j) = defor23(i, k, j);
    for (int j = j_start; j <= j_end + 1; j++)
      for (int i = i_start; i <= i_end; i++) { tmp2(i, kts, j) = 0.0f; tmp2(i, ktf + 1, j) = 0.0f; }
    LOOP3(kts, ktf)
      avg(i, k, j) = 0.25f * ((tmp2(i, k + 1, j) * tmp2(i, k + 1, j)) + (tmp2(i, k, j) * tmp2(i, k, j)) + (tmp2(i, k + 1, j + 1) * tmp2(i, k + 1, j + 1)) +
                              (tmp2(i, k, j + 1) * tmp2(i, k, j + 1)));
    LOOP3(kts, ktf) tendency(i, k, j) = tendency(i, k, j) + (c1(k) * mu(i, j) + c2(k)) * xkmv(i, k, j) * avg(i, k, j);
    {
      const int k = kts;
      for (int j = j_start; j <= j_end; j++)
        for (int i = i_start; i <= i_end; i++) {
          float absU, Cd;
          if (cf.isfflx == 0) { absU = absU_(i, k, j); Cd = cf.tke_drag_coefficient; }
          else { absU = absU_(i, k, j) + epsilon; Cd = (ust(i, j) * ust(i, j)) / (absU * absU); }
          tendency(i, k, j) = tendency(i, k, j) + (c1(k) * mu(i, j) + c2(k)) * ((v(i, k, j) + v(i, k, j + 1)) * 0.5f * Cd * absU *
                                                                               (defor23(i, kts + 1, j) + defor23(i, kts + 1, j + 1)) * 0.5f);
        }
    }
  }
  // ---- tke_buoyancy (:6234-6378) and tke_dissip (:6384-6523): the same range
  {
    const R4 r = mass_range(x, true);
    const int i_start = r.i_start, i_end = r.i_end, j_start = r.j_start, j_end = r.j_end;
    LOOP3(kts + 1, ktf) tendency(i, k, j) = tendency(i, k, j) - (c1(k) * mu(i, j) + c2(k)) * xkhv(i, k, j) * BN2(i, k, j);
    {
      const int k = kts;
      for (int j = j_start; j <= j_end; j++)
        for (int i = i_start; i <= i_end; i++) {
          float heat_flux;
          if (cf.isfflx == 0 || cf.isfflx == 2) heat_flux = cf.tke_heat_flux;
          else {
            const float cpm = cp * (1.f + 0.8f * (qv ? (*qv)(i, k, j) : 0.f));
            heat_flux = (hfx(i, j) / cpm) / rho(i, k, j);
          }
          tendency(i, k, j) = tendency(i, k, j) - (c1(k) * mu(i, j) + c2(k)) * ((xkhv(i, k, j) * BN2(i, k, j)) - (g / theta(i, k, j)) * heat_flux) / 2.f;
        }
    }
    const float c_k = cf.c_k;
    const float ce1 = (c_k / 0.10f) * 0.19f;
    const float ce2 = std::max(0.0f, 0.93f - ce1);
    calc_l_scale(x, tke, BN2, l_scale, r, ktf, rdzw);
    LOOP3(kts, ktf) {
      const float deltas = std::pow(cf.dx / x.msftx(i, j) * cf.dy / x.msfty(i, j) / rdzw(i, k, j), 0.33333333f);
      const float tketmp = std::max(tke(i, k, j), 1.0e-6f);
      float coefc;
      if (k == kts || k == ktf) coefc = 3.9f;
      else coefc = ce1 + ce2 * l_scale(i, k, j) / deltas;
      tendency(i, k, j) = tendency(i, k, j) - (c1(k) * mu(i, j) + c2(k)) * coefc * std::pow(tketmp, 1.5f) / l_scale(i, k, j);
    }
  }
  {
    const R4 r = mass_range(x, true);
    const int i_start = r.i_start, i_end = r.i_end, j_start = r.j_start, j_end = r.j_end;
    LOOP3(kts, ktf) tendency(i, k, j) = std::max(tendency(i, k, j), -(c1(k) * mu(i, j) + c2(k)) * std::max(0.0f, tke(i, k, j)) / dt);
#undef LOOP3
  }
}

// module_em.F:2616-2646
void bound_tke(const Cx &x, A3 tke, float tke_upper_bound) {
  for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
    for (int k = x.kts; k <= x.kte - 1; k++)
      for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) tke(i, k, j) = std::min(tke_upper_bound, std::max(tke(i, k, j), 0.f));
}

}  // namespace orc
