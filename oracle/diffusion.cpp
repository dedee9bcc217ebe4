// diffusion.cpp -- constant-K diffusion on coordinate surfaces, diff_opt = 1 with km_opt = 1
// (the "constant-K" part): horizontal_diffusion (module_big_step_utilities_em.F:2715),
// horizontal_diffusion_3dmp (:2954), vertical_diffusion (:3064), vertical_diffusion_mp (:3192),
// vertical_diffusion_3dmp (:3286), vertical_diffusion_u (:3382), vertical_diffusion_v (:3485).
// TEST INFRASTRUCTURE (see oracle.h).
//
// The eddy-viscosity arrays xkmh / xkhh the reference passes in are filled by isotropic_km
// (module_diffusion_em.F:1718-1772) with the constants khdif and khdif/prandtl on every mass point, then
// exchanged (HALO_EM_TKE_E) and extended by phy_bc's set_physical_bc3d: every element the routines below
// read holds the same number: the routines take an `XK` (oracle.h), which is either that number or -- for the
// closures that compute K (smag2d_km, deform.cpp) -- the xkmh / xkhh array itself, indexed as in the reference.
#include "oracle.h"

namespace orc {

// name: 'u', 'v', 'w' or anything else for a mass-point field ("m")
void horizontal_diffusion(const Cx &x, char name, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, XK xkmhd) {
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
          mkrdxm = (msftx(i - 1, j) / msfty(i - 1, j)) * (c1(k) * MUT(i - 1, j) + c2(k)) * xkmhd(i - 1, k, j) * rdx;
          mkrdxp = (msftx(i, j) / msfty(i, j)) * (c1(k) * MUT(i, j) + c2(k)) * xkmhd(i, k, j) * rdx;
          mrdx = msfux(i, j) * msfuy(i, j) * rdx;
          mkrdym = ((msfuy(i, j) + msfuy(i, j - 1)) / (msfux(i, j) + msfux(i, j - 1))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1) + xkmhd(i - 1, k, j - 1) + xkmhd(i - 1, k, j)) * rdy;
          mkrdyp = ((msfuy(i, j) + msfuy(i, j + 1)) / (msfux(i, j) + msfux(i, j + 1))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i - 1, j + 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) *
                   0.25f * (xkmhd(i, k, j) + xkmhd(i, k, j + 1) + xkmhd(i - 1, k, j + 1) + xkmhd(i - 1, k, j)) * rdy;
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
                   0.25f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1) + xkmhd(i - 1, k, j - 1) + xkmhd(i - 1, k, j)) * rdx;
          mkrdxp = ((msfvx(i, j) + msfvx(i + 1, j)) / (msfvy(i, j) + msfvy(i + 1, j))) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i + 1, j - 1) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k))) *
                   0.25f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1) + xkmhd(i + 1, k, j - 1) + xkmhd(i + 1, k, j)) * rdx;
          mrdx = msfvx(i, j) * msfvy(i, j) * rdx;
          mkrdym = (msfty(i, j - 1) / msftx(i, j - 1)) * xkmhd(i, k, j - 1) * rdy;
          mkrdyp = (msfty(i, j) / msftx(i, j)) * xkmhd(i, k, j) * rdy;
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
                   0.25f * (xkmhd(i, k, j) + xkmhd(i - 1, k, j) + xkmhd(i, k - 1, j) + xkmhd(i - 1, k - 1, j)) * rdx;
          mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) *
                   0.25f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) *
                   0.25f * (xkmhd(i + 1, k, j) + xkmhd(i, k, j) + xkmhd(i + 1, k - 1, j) + xkmhd(i, k - 1, j)) * rdx;
          mrdx = msftx(i, j) * msfty(i, j) * rdx;
          mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) *
                   0.25f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) *
                   0.25f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1) + xkmhd(i, k - 1, j) + xkmhd(i, k - 1, j - 1)) * rdy;
          mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) *
                   0.25f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) *
                   0.25f * (xkmhd(i, k, j + 1) + xkmhd(i, k, j) + xkmhd(i, k - 1, j + 1) + xkmhd(i, k - 1, j)) * rdy;
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
          mkrdxm = (msfux(i, j) / msfuy(i, j)) * 0.5f * (xkmhd(i, k, j) + xkmhd(i - 1, k, j)) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) * rdx;
          mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) * 0.5f * (xkmhd(i + 1, k, j) + xkmhd(i, k, j)) * 0.5f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdx;
          mrdx = msftx(i, j) * msfty(i, j) * rdx;
          mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) * 0.5f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1)) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) * rdy;
          mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) * 0.5f * (xkmhd(i, k, j + 1) + xkmhd(i, k, j)) * 0.5f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdy;
          mrdy = msftx(i, j) * msfty(i, j) * rdy;
          tendency(i, k, j) = tendency(i, k, j) + (mrdx * (mkrdxp * (field(i + 1, k, j) - field(i, k, j)) - mkrdxm * (field(i, k, j) - field(i - 1, k, j))) +
                                                   mrdy * (mkrdyp * (field(i, k, j + 1) - field(i, k, j)) - mkrdym * (field(i, k, j) - field(i, k, j - 1))));
        }
  }
}

// :2954-3060, mass-point field minus a 3-D base state (theta - t_init)
void horizontal_diffusion_3dmp(const Cx &x, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, A3 base_3d, XK xkmhd) {
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
        const float mkrdxm = (msfux(i, j) / msfuy(i, j)) * 0.5f * (xkmhd(i, k, j) + xkmhd(i - 1, k, j)) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k))) * rdx;
        const float mkrdxp = (msfux(i + 1, j) / msfuy(i + 1, j)) * 0.5f * (xkmhd(i + 1, k, j) + xkmhd(i, k, j)) * 0.5f * ((c1(k) * MUT(i + 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdx;
        const float mrdx = msftx(i, j) * msfty(i, j) * rdx;
        const float mkrdym = (msfvy(i, j) * msfvx_inv(i, j)) * 0.5f * (xkmhd(i, k, j) + xkmhd(i, k, j - 1)) * 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k))) * rdy;
        const float mkrdyp = (msfvy(i, j + 1) * msfvx_inv(i, j + 1)) * 0.5f * (xkmhd(i, k, j + 1) + xkmhd(i, k, j)) * 0.5f * ((c1(k) * MUT(i, j + 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k))) * rdy;
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

// module_big_step_utilities_em.F:6220-6620 sixth_order_diffusion (Knievel et al. 2007): the 6th-order horizontal filter
// of diff_6th_opt = 1, 2 (2: fluxes that run up the gradient are dropped), optionally tapered over steep terrain
// (diff_6th_slopeopt), added to the *_tendf arrays on the first RK substep (module_em.F:880-930, 1421-1437).
void sixth_order_diffusion(const Cx &x, char name, A3 field, A3 tendency, A2 MUT, float dt, A1 c1, A1 c2, A3 phb) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts;
  const int ktf = std::min(x.kte, x.kde - 1);
  const int opt = x.c.diff_6th_opt;
  const float rdx = x.rdx, rdy = x.rdy;
  const A2 &msfux = x.msfux, &msfuy = x.msfuy, &msfvx = x.msfvx, &msfvy = x.msfvy, &msftx = x.msftx, &msfty = x.msfty;
  const float diff_6th_coef = x.c.diff_6th_factor * 0.015625f / (2.0f * dt);
  int i_start = its, i_end, j_start = jts, j_end, k_start = kts, k_end = ktf;
  if (name == 'u') {
    i_end = ite; j_end = std::min(jde - 1, jte);
    if (x.c.open_xs) i_start = std::max(its, ids + 3);
    if (x.c.open_ys) j_start = std::max(jts, jds + 3);
    if (x.c.open_xe) i_end = std::min(ite, ide - 3);
    if (x.c.open_ye) j_end = std::min(jte, jde - 4);
    if (x.spec) {   // :6354
      i_start = std::max(its, ids + 3); i_end = std::min(ite, ide - 3);
      j_start = std::max(jts, jds + 3); j_end = std::min(jte, jde - 4);
    }
  } else if (name == 'v') {
    i_end = std::min(ide - 1, ite); j_end = jte;
    if (x.c.open_xs) i_start = std::max(its, ids + 3);
    if (x.c.open_ys) j_start = std::max(jts, jds + 3);
    if (x.c.open_xe) i_end = std::min(ite, ide - 4);
    if (x.c.open_ye) j_end = std::min(jte, jde - 3);
    if (x.c.open_xs || x.spec) {   // sic (:6381): "open_xs .or. specified" restricts all four sides
      i_start = std::max(its, ids + 3); i_end = std::min(ite, ide - 4);
      j_start = std::max(jts, jds + 3); j_end = std::min(jte, jde - 3);
    }
  } else {
    i_end = std::min(ide - 1, ite); j_end = std::min(jde - 1, jte);
    if (x.c.open_xs) i_start = std::max(its, ids + 3);
    if (x.c.open_ys) j_start = std::max(jts, jds + 3);
    if (x.c.open_xe) i_end = std::min(ite, ide - 4);
    if (x.c.open_ye) j_end = std::min(jte, jde - 4);
    if (x.spec) {   // :6408, :6435
      i_start = std::max(its, ids + 3); i_end = std::min(ide - 4, ite);
      j_start = std::max(jts, jds + 3); j_end = std::min(jde - 4, jte);
    }
    if (name == 'w') k_start = kts + 1;
  }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        float dflux_x_p0 = (10.0f * (field(i, k, j) - field(i - 1, k, j)) - 5.0f * (field(i + 1, k, j) - field(i - 2, k, j)) + (field(i + 2, k, j) - field(i - 3, k, j)));
        float dflux_x_p1 = (10.0f * (field(i + 1, k, j) - field(i, k, j)) - 5.0f * (field(i + 2, k, j) - field(i - 1, k, j)) + (field(i + 3, k, j) - field(i - 2, k, j)));
        if (opt == 2) {
          if (dflux_x_p0 * (field(i, k, j) - field(i - 1, k, j)) <= 0.0f) dflux_x_p0 = 0.0f;
          if (dflux_x_p1 * (field(i + 1, k, j) - field(i, k, j)) <= 0.0f) dflux_x_p1 = 0.0f;
        }
        float slopedamp_p0 = 1.f, slopedamp_p1 = 1.f;
        if (x.c.diff_6th_slopeopt >= 1) {
          const float dx = 1.f / rdx;
          const float dzthresh = x.c.diff_6th_thresh * 9.81f * dx;
          float dzmax_p0, dzmax_p1;
          if (name == 'u') {
            dzmax_p0 = std::max(std::fabs(phb(i, k, j) - phb(i - 1, k, j)) * msfux(i, j), std::fabs(phb(i - 1, k, j) - phb(i - 2, k, j)) * msfux(i - 1, j));
            dzmax_p1 = std::max(std::fabs(phb(i + 1, k, j) - phb(i, k, j)) * msfux(i + 1, j), std::fabs(phb(i, k, j) - phb(i - 1, k, j)) * msfux(i, j));
          } else if (name == 'v') {
            dzmax_p0 = std::max(std::fabs(phb(i, k, j) - phb(i - 1, k, j)) * msfux(i, j), std::fabs(phb(i, k, j - 1) - phb(i - 1, k, j - 1)) * msfux(i, j - 1));
            dzmax_p1 = std::max(std::fabs(phb(i + 1, k, j) - phb(i, k, j)) * msfux(i + 1, j), std::fabs(phb(i + 1, k, j - 1) - phb(i, k, j - 1)) * msfux(i + 1, j - 1));
          } else {
            dzmax_p0 = std::fabs(phb(i, k, j) - phb(i - 1, k, j)) * msfux(i, j);
            dzmax_p1 = std::fabs(phb(i + 1, k, j) - phb(i, k, j)) * msfux(i + 1, j);
          }
          slopedamp_p0 = std::max(1.f - (dzmax_p0 / dzthresh), 0.f);
          slopedamp_p1 = std::max(1.f - (dzmax_p1 / dzthresh), 0.f);
        }
        float mu_avg_p0, mu_avg_p1, tendency_x, tendency_y;
        if (name == 'u') {
          mu_avg_p0 = (c1(k) * MUT(i - 1, j) + c2(k));
          mu_avg_p1 = (c1(k) * MUT(i, j) + c2(k));
          tendency_x = diff_6th_coef * msfux(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
        } else if (name == 'v') {
          mu_avg_p0 = 0.25f * ((c1(k) * MUT(i - 1, j - 1) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)));
          mu_avg_p1 = 0.25f * ((c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i + 1, j - 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k)));
          tendency_x = diff_6th_coef * msfvx(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
        } else {
          mu_avg_p0 = 0.5f * ((c1(k) * MUT(i - 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)));
          mu_avg_p1 = 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i + 1, j) + c2(k)));
          tendency_x = diff_6th_coef * msftx(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
        }
        float dflux_y_p0 = (10.0f * (field(i, k, j) - field(i, k, j - 1)) - 5.0f * (field(i, k, j + 1) - field(i, k, j - 2)) + (field(i, k, j + 2) - field(i, k, j - 3)));
        float dflux_y_p1 = (10.0f * (field(i, k, j + 1) - field(i, k, j)) - 5.0f * (field(i, k, j + 2) - field(i, k, j - 1)) + (field(i, k, j + 3) - field(i, k, j - 2)));
        if (opt == 2) {
          if (dflux_y_p0 * (field(i, k, j) - field(i, k, j - 1)) <= 0.0f) dflux_y_p0 = 0.0f;
          if (dflux_y_p1 * (field(i, k, j + 1) - field(i, k, j)) <= 0.0f) dflux_y_p1 = 0.0f;
        }
        slopedamp_p0 = 1.f; slopedamp_p1 = 1.f;
        if (x.c.diff_6th_slopeopt >= 1) {
          const float dy = 1.f / rdy;
          const float dzthresh = x.c.diff_6th_thresh * 9.81f * dy;
          float dzmax_p0, dzmax_p1;
          if (name == 'u') {
            dzmax_p0 = std::max(std::fabs(phb(i, k, j) - phb(i, k, j - 1)) * msfvy(i, j), std::fabs(phb(i - 1, k, j) - phb(i - 1, k, j - 1)) * msfvy(i - 1, j));
            dzmax_p1 = std::max(std::fabs(phb(i, k, j + 1) - phb(i, k, j)) * msfvy(i, j + 1), std::fabs(phb(i - 1, k, j + 1) - phb(i - 1, k, j)) * msfvy(i - 1, j + 1));
          } else if (name == 'v') {
            dzmax_p0 = std::max(std::fabs(phb(i, k, j) - phb(i, k, j - 1)) * msfvy(i, j), std::fabs(phb(i, k, j - 1) - phb(i, k, j - 2)) * msfvy(i, j - 1));
            dzmax_p1 = std::max(std::fabs(phb(i, k, j + 1) - phb(i, k, j)) * msfvy(i, j + 1), std::fabs(phb(i, k, j) - phb(i, k, j - 1)) * msfvy(i, j));
          } else {
            dzmax_p0 = std::fabs(phb(i, k, j) - phb(i, k, j - 1)) * msfvy(i, j);
            dzmax_p1 = std::fabs(phb(i, k, j + 1) - phb(i, k, j)) * msfvy(i, j + 1);
          }
          slopedamp_p0 = std::max(1.f - (dzmax_p0 / dzthresh), 0.f);
          slopedamp_p1 = std::max(1.f - (dzmax_p1 / dzthresh), 0.f);
        }
        if (name == 'u') {
          mu_avg_p0 = 0.25f * ((c1(k) * MUT(i - 1, j - 1) + c2(k)) + (c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i - 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)));
          mu_avg_p1 = 0.25f * ((c1(k) * MUT(i - 1, j) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i - 1, j + 1) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)));
          tendency_y = diff_6th_coef * msfuy(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
        } else if (name == 'v') {
          mu_avg_p0 = (c1(k) * MUT(i, j - 1) + c2(k));
          mu_avg_p1 = (c1(k) * MUT(i, j) + c2(k));
          tendency_y = diff_6th_coef * msfvy(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
        } else {
          mu_avg_p0 = 0.5f * ((c1(k) * MUT(i, j - 1) + c2(k)) + (c1(k) * MUT(i, j) + c2(k)));
          mu_avg_p1 = 0.5f * ((c1(k) * MUT(i, j) + c2(k)) + (c1(k) * MUT(i, j + 1) + c2(k)));
          tendency_y = diff_6th_coef * msfty(i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
        }
        tendency(i, k, j) = tendency(i, k, j) + tendency_x + tendency_y;
      }
}

}  // namespace orc
