// diffusion.cpp -- constant-K diffusion on coordinate surfaces, diff_opt = 1 with km_opt = 1
// (the "constant-K" part): horizontal_diffusion (module_big_step_utilities_em.F:2715),
// horizontal_diffusion_3dmp (:2954), vertical_diffusion (:3064), vertical_diffusion_mp (:3192),
// vertical_diffusion_3dmp (:3286), vertical_diffusion_u (:3382), vertical_diffusion_v (:3485).
// TEST INFRASTRUCTURE (see oracle.h).
//
// The eddy-viscosity arrays xkmh / xkhh the reference passes in are filled by isotropic_km
// (module_diffusion_em.F:1718-1772) with the constants khdif and khdif/prandtl on every mass point, then
// exchanged (HALO_EM_TKE_E) and extended by phy_bc's set_physical_bc3d: every element the routines below
// read holds the same number, so the restatement takes that number `xk` instead of an array and keeps
// the averaging expressions (0.25*(xk+xk+xk+xk) ...) as written.
#include "oracle.h"

namespace orc {

// name: 'u', 'v', 'w' or anything else for a mass-point field ("m")
void horizontal_diffusion(const Cx &x, char name, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, float xk) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const float rdx = x.rdx, rdy = x.rdy;
  const A2 &msfux = x.msfux, &msfuy = x.msfuy, &msfvx = x.msfvx, &msfvx_inv = x.msfvx_inv, &msfvy = x.msfvy, &msftx = x.msftx, &msfty = x.msfty;
  const bool specified = cf.specified || cf.nested;
  const int ktf = std::min(kte, kde - 1);
  int i_start, i_end, j_start, j_end;
  float mrdx, mkrdxm, mkrdxp, mrdy, mkrdym, mkrdyp;
  if (name == 'u') {
    i_start = its; i_end = ite; j_start = jts; j_end = std::min(jte, jde - 1);
    if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
    if (cf.open_xe || specified) i_end = std::min(ide - 1, ite);
    if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
    if (cf.open_ye || specified) j_end = std::min(jde - 2, jte);
    if (cf.periodic_x) i_start = its;
    if (cf.periodic_x) i_end = ite;
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          mkrdxm = (msftx(i - 1, j) / msfty(i - 1, j)) * (c1(k) * MUT(i - 1, j) + c2(k)) * xk * rdx;
          mkrdxp = (msftx(i, j) / msfty(i, j)) * (c1(k) * MUT(i, j) + c2(k)) * xk * rdx;
          mrdx = msfux(i, j) * msfuy(i, j) * rdx;
          mkrdym = ((msfuy(i, j) + msfuy(i, j - 1)) / (msfux(i, j) + msfux(i, j - 1))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdy;
          mkrdyp = ((msfuy(i, j) + msfuy(i, j + 1)) / (msfux(i, j) + msfux(i, j + 1))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i - 1, j + 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdy;
          mrdy = msfux(i, j) * msfuy(i, j) * rdy;
          tendency(i, k, j) = tendency(i, k, j) + (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j)) - mkrdxm * (field(i, k, j) - field(i - 1, k, j))) +
                                                   mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j)) - mkrdym * (field(i, k, j) - field(i, k, j - 1))));
        }
  } else if (name == 'v') {
    i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = jte;
    if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
    if (cf.open_xe || specified) i_end = std::min(ide - 2, ite);
    if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
    if (cf.open_ye || specified) j_end = std::min(jde - 1, jte);
    if (cf.periodic_x) i_start = its;
    if (cf.periodic_x) i_end = std::min(ite, ide - 1);
    if (cf.polar) j_start = std::max(jds + 1, jts);
    if (cf.polar) j_end = std::min(jde - 1, jte);
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          mkrdxm = ((msfvx(i, j) + msfvx(i - 1, j)) / (msfvy(i, j) + msfvy(i - 1, j))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdx;
          mkrdxp = ((msfvx(i, j) + msfvx(i + 1, j)) / (msfvy(i, j) + msfvy(i + 1, j))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i + 1, j - 1) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdx;
          mrdx = msfvx(i, j) * msfvy(i, j) * rdx;
          mkrdym = (msfty(i, j - 1) / msftx(i, j - 1)) * xk * rdy;
          mkrdyp = (msfty(i, j) / msftx(i, j)) * xk * rdy;
          mrdy = msfvx(i, j) * msfvy(i, j) * rdy;
          tendency(i, k, j) = tendency(i, k, j) + (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j)) - mkrdxm * (field(i, k, j) - field(i - 1, k, j))) +
                                                   mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j)) - mkrdym * (field(i, k, j) - field(i, k, j - 1))));
        }
  } else if (name == 'w') {
    i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
    if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
    if (cf.open_xe || specified) i_end = std::min(ide - 2, ite);
    if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
    if (cf.open_ye || specified) j_end = std::min(jde - 2, jte);
    if (cf.periodic_x) i_start = its;
    if (cf.periodic_x) i_end = std::min(ite, ide - 1);
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          mkrdxm = (msfux(i, j) / msfuy(i, j)) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdx;
          mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) *
                   0.25f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdx;
          mrdx = msftx(i, j) * msfty(i, j) * rdx;
          mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdy;
          mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) *
                   0.25f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) *
                   0.25f * (xk + xk + xk + xk) * rdy;
          mrdy = msftx(i, j) * msfty(i, j) * rdy;
          tendency(i, k, j) = tendency(i, k, j) + (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j)) - mkrdxm * (field(i, k, j) - field(i - 1, k, j))) +
                                                   mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j)) - mkrdym * (field(i, k, j) - field(i, k, j - 1))));
        }
  } else {
    i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
    if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
    if (cf.open_xe || specified) i_end = std::min(ide - 2, ite);
    if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
    if (cf.open_ye || specified) j_end = std::min(jde - 2, jte);
    if (cf.periodic_x) i_start = its;
    if (cf.periodic_x) i_end = std::min(ite, ide - 1);
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++) {
          mkrdxm = (msfux(i, j) / msfuy(i, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) * rdx;
          mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdx;
          mrdx = msftx(i, j) * msfty(i, j) * rdx;
          mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) * rdy;
          mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdy;
          mrdy = msftx(i, j) * msfty(i, j) * rdy;
          tendency(i, k, j) = tendency(i, k, j) + (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j)) - mkrdxm * (field(i, k, j) - field(i - 1, k, j))) +
                                                   mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j)) - mkrdym * (field(i, k, j) - field(i, k, j - 1))));
        }
  }
}

// :2954-3060, mass-point field minus a 3-D base state (theta - t_init)
void horizontal_diffusion_3dmp(const Cx &x, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, A3 base_3d, float xk) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const float rdx = x.rdx, rdy = x.rdy;
  const A2 &msfux = x.msfux, &msfuy = x.msfuy, &msfvx_inv = x.msfvx_inv, &msfvy = x.msfvy, &msftx = x.msftx, &msfty = x.msfty;
  const bool specified = cf.specified || cf.nested;
  const int ktf = std::min(kte, kde - 1);
  int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
  if (cf.open_xe || specified) i_end = std::min(ide - 2, ite);
  if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || specified) j_end = std::min(jde - 2, jte);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = std::min(ite, ide - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float mkrdxm = (msfux(i, j) / msfuy(i, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) * rdx;
        const float mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdx;
        const float mrdx = msftx(i, j) * msfty(i, j) * rdx;
        const float mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) * rdy;
        const float mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) * 0.5f * (xk + xk) * 0.5f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdy;
        const float mrdy = msftx(i, j) * msfty(i, j) * rdy;
        tendency(i, k, j) = tendency(i, k, j) +
                            (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j) - base_3d(i + 1, k, j) + base_3d(i, k, j)) -
                                     mkrdxm * (field(i, k, j) - field(i - 1, k, j) - base_3d(i, k, j) + base_3d(i - 1, k, j))) +
                             mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j) - base_3d(i, k, j + 1) + base_3d(i, k, j)) -
                                     mkrdym * (field(i, k, j) - field(i, k, j - 1) - base_3d(i, k, j) + base_3d(i, k, j - 1))));
      }
}

// :3064-3187, name 'w' or 'm'
void vertical_diffusion(const Cx &x, char name, A3 field, A3 tendency, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif) {
  const int ide = x.ide, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const A1 &rdn = x.rdn, &rdnw = x.rdnw;
  const int ktf = std::min(kte, kde - 1);
  const int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  std::vector<float> vfl((size_t)(i_end - i_start + 1 > 0 ? i_end - i_start + 1 : 0) * (kte + 2));
  auto vflux = [&](int i, int k) -> float & { return vfl[(size_t)(i - i_start) * (kte + 2) + k]; };
  if (name == 'w') {
    for (int j = j_start; j <= j_end; j++) {
      for (int k = kts; k <= ktf - 1; k++)
        for (int i = i_start; i <= i_end; i++) vflux(i, k) = (kvdif / alt(i, k, j)) * rdnw(k) * (field(i, k + 1, j) - field(i, k, j));
      for (int i = i_start; i <= i_end; i++) vflux(i, ktf) = 0.f;
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++)
          tendency(i, k, j) = tendency(i, k, j) + rdn(k) * g * g / (c1(k) * MUT(i, j) + c2(k)) / (0.5f * (alt(i, k, j) + alt(i, k - 1, j))) * (vflux(i, k) - vflux(i, k - 1));
    }
  } else {
    for (int j = j_start; j <= j_end; j++) {
      for (int k = kts; k <= ktf - 1; k++)
        for (int i = i_start; i <= i_end; i++)
          vflux(i, k) = kvdif * rdn(k + 1) / (0.5f * (alt(i, k, j) + alt(i, k + 1, j))) * (field(i, k + 1, j) - field(i, k, j));
      for (int i = i_start; i <= i_end; i++) vflux(i, 0) = vflux(i, 1);
      for (int i = i_start; i <= i_end; i++) vflux(i, ktf) = 0.f;
      for (int k = kts; k <= ktf; k++)
        for (int i = i_start; i <= i_end; i++)
          tendency(i, k, j) = tendency(i, k, j) + g * g / (c1(k) * MUT(i, j) + c2(k)) / alt(i, k, j) * rdnw(k) * (vflux(i, k) - vflux(i, k - 1));
    }
  }
}

// :3192-3281 (1-D base state: qv - qv_base) and :3286-3376 (3-D base state: theta - t_init)
static void vertical_diffusion_base(const Cx &x, A3 field, A3 tendency, const A1 *base, const A3 *base_3d, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif) {
  const int ide = x.ide, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const A1 &rdn = x.rdn, &rdnw = x.rdnw;
  const int ktf = std::min(kte, kde - 1);
  const int i_start = its, i_end = std::min(ite, ide - 1), j_start = jts, j_end = std::min(jte, jde - 1);
  std::vector<float> vfl((size_t)(i_end - i_start + 1 > 0 ? i_end - i_start + 1 : 0) * (kte + 2));
  auto vflux = [&](int i, int k) -> float & { return vfl[(size_t)(i - i_start) * (kte + 2) + k]; };
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) {
        if (base)
          vflux(i, k) = kvdif * rdn(k + 1) / (0.5f * (alt(i, k, j) + alt(i, k + 1, j))) * (field(i, k + 1, j) - field(i, k, j) - (*base)(k + 1) + (*base)(k));
        else
          vflux(i, k) = kvdif * rdn(k + 1) / (0.5f * (alt(i, k, j) + alt(i, k + 1, j))) *
                        (field(i, k + 1, j) - field(i, k, j) - (*base_3d)(i, k + 1, j) + (*base_3d)(i, k, j));
      }
    for (int i = i_start; i <= i_end; i++) vflux(i, 0) = vflux(i, 1);
    for (int i = i_start; i <= i_end; i++) vflux(i, ktf) = 0.f;
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) + g * g / (c1(k) * MUT(i, j) + c2(k)) / alt(i, k, j) * rdnw(k) * (vflux(i, k) - vflux(i, k - 1));
  }
}
void vertical_diffusion_mp(const Cx &x, A3 field, A3 tendency, A1 base, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif) {
  vertical_diffusion_base(x, field, tendency, &base, nullptr, c1, c2, alt, MUT, kvdif);
}
void vertical_diffusion_3dmp(const Cx &x, A3 field, A3 tendency, A3 base_3d, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif) {
  vertical_diffusion_base(x, field, tendency, nullptr, &base_3d, c1, c2, alt, MUT, kvdif);
}

// :3382-3480
void vertical_diffusion_u(const Cx &x, A3 field, A3 tendency, A1 u_base, A1 c1h, A1 c2h, A3 alt, A2 muu, float kvdif) {
  const int ids = x.ids, ide = x.ide, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const A1 &rdn = x.rdn, &rdnw = x.rdnw;
  const bool specified = cf.specified || cf.nested;
  const int ktf = std::min(kte, kde - 1);
  int i_start = its, i_end = ite;
  const int j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_xs || specified) i_start = std::max(ids + 1, its);
  if (cf.open_xe || specified) i_end = std::min(ide - 1, ite);
  if (cf.periodic_x) i_start = its;
  if (cf.periodic_x) i_end = ite;
  std::vector<float> vfl((size_t)(i_end - i_start + 1 > 0 ? i_end - i_start + 1 : 0) * (kte + 2));
  auto vflux = [&](int i, int k) -> float & { return vfl[(size_t)(i - i_start) * (kte + 2) + k]; };
  for (int j = j_start; j <= j_end; j++) {
    for (int k = kts; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++)
        vflux(i, k) = kvdif * rdn(k + 1) / (0.25f * (alt(i, k, j) + alt(i - 1, k, j) + alt(i, k + 1, j) + alt(i - 1, k + 1, j))) *
                      (field(i, k + 1, j) - field(i, k, j) - u_base(k + 1) + u_base(k));
    for (int i = i_start; i <= i_end; i++) vflux(i, 0) = vflux(i, 1);
    for (int i = i_start; i <= i_end; i++) vflux(i, ktf) = 0.f;
    for (int k = kts; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) + g * g * rdnw(k) / (c1h(k) * muu(i, j) + c2h(k)) / (0.5f * (alt(i - 1, k, j) + alt(i, k, j))) * (vflux(i, k) - vflux(i, k - 1));
  }
}

// :3485-3581
void vertical_diffusion_v(const Cx &x, A3 field, A3 tendency, A1 v_base, A1 c1h, A1 c2h, A3 alt, A2 muv, float kvdif) {
  const int ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const wrfgpu_config &cf = x.c;
  const A1 &rdn = x.rdn, &rdnw = x.rdnw;
  const bool specified = cf.specified || cf.nested;
  const int ktf = std::min(kte, kde - 1);
  const int i_start = its, i_end = std::min(ite, ide - 1);
  int j_start = jts, j_end = std::min(jte, jde - 1);
  if (cf.open_ys || specified) j_start = std::max(jds + 1, jts);
  if (cf.open_ye || specified) j_end = std::min(jde - 1, jte);
  std::vector<float> vfl((size_t)(i_end - i_start + 1 > 0 ? i_end - i_start + 1 : 0) * (kte + 2));
  auto vflux = [&](int i, int k) -> float & { return vfl[(size_t)(i - i_start) * (kte + 2) + k]; };
  for (int j = j_start; j <= j_end; j++) {
    const int jm1 = j - 1;
    for (int k = kts; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++)
        vflux(i, k) = kvdif * rdn(k + 1) / (0.25f * (alt(i, k, j) + alt(i, k, jm1) + alt(i, k + 1, j) + alt(i, k + 1, jm1))) *
                      (field(i, k + 1, j) - field(i, k, j) - v_base(k + 1) + v_base(k));
    for (int i = i_start; i <= i_end; i++) vflux(i, 0) = vflux(i, 1);
    for (int i = i_start; i <= i_end; i++) vflux(i, ktf) = 0.f;
    for (int k = kts; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) + g * g * rdnw(k) / (c1h(k) * muv(i, j) + c2h(k)) / (0.5f * (alt(i, k, jm1) + alt(i, k, j))) * (vflux(i, k) - vflux(i, k - 1));
  }
}

}  // namespace orc
