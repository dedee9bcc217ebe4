// diffusion2.cuh -- point functions of the diffusion in physical space (diff_opt = 2):
// vertical_diffusion_2 and horizontal_diffusion_2 of dyn_em/module_diffusion_em.F (:4004-4906, :2864-3998) with the
// stresses cal_titau_* (:5331-5895), as called from first_rk_step_part2 (vertical first, then horizontal).
//
// The reference builds each stress / flux component in a tile-sized work array (titau1, titau2, titau1avg, H1, H2, ...)
// and differences it in a later loop nest; every element of those arrays is a pure function of a few neighbouring points
// of the deformation tensor, K, rho and the metrics, and the consumers only read elements inside the range the
// producers filled (checked routine by routine against the is_ext ... je_ext arguments).  The thread of a point therefore
// evaluates the elements it needs in registers, in the reference's order of operations, and adds the vertical term, the
// surface flux and the horizontal term to its tendency in the reference's order -- one launch per tendency array.
// __host__ __device__ so that tests/deform_shim.cu runs exactly this code on the CPU against the oracle.
#pragma once
#include "deform.cuh"

namespace wg {

struct D2In {
  cfp defor11, defor22, defor33, defor12, defor13, defor23, xkmh, xkmv, xkhh, xkhv, zx, zy, rdz, rdzw, rho;
  const float *dn;                  // origin shifted 1-D (dnw, fnm, fnp are in Geo)
  cfp u, v, t, qv;                  // surface fluxes: u_2, v_2, t_2, moist_qv (or null)
  cfp ustm, hfx_in, qfx;            // 2-D
  float *hfx_out;                   // isfflx 0 / 2 store the diagnosed heat flux
  int isfflx, use_theta_m;
  float cd0, heat_flux;
};
struct R4 { int i0, i1, j0, j1; };
// loop ranges by staggering: 'u', 'v', anything else = mass / w points (module_diffusion_em.F:3167-3181 and alike)
inline R4 d2_range(const GeoBase &g, char kind) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  auto mn = [](int a, int b) { return a < b ? a : b; };
  auto mx = [](int a, int b) { return a > b ? a : b; };
  R4 r;
  if (kind == 'u') {
    r = {its, ite, jts, mn(jte, jde - 1)};
    if (g.open_xs) r.i0 = mx(ids + 1, its);
    if (g.open_xe) r.i1 = mn(ide - 1, ite);
    if (g.open_ys) r.j0 = mx(jds + 1, jts);
    if (g.open_ye) r.j1 = mn(jde - 2, jte);
    if (g.periodic_x) { r.i0 = its; r.i1 = ite; }
  } else if (kind == 'v') {
    r = {its, mn(ite, ide - 1), jts, jte};
    if (g.open_xs) r.i0 = mx(ids + 1, its);
    if (g.open_xe) r.i1 = mn(ide - 2, ite);
    if (g.open_ys) r.j0 = mx(jds + 1, jts);
    if (g.open_ye) r.j1 = mn(jde - 1, jte);
    if (g.periodic_x) { r.i0 = its; r.i1 = mn(ite, ide - 1); }
  } else {
    r = {its, mn(ite, ide - 1), jts, mn(jte, jde - 1)};
    if (g.open_xs) r.i0 = mx(ids + 1, its);
    if (g.open_xe) r.i1 = mn(ide - 2, ite);
    if (g.open_ys) r.j0 = mx(jds + 1, jts);
    if (g.open_ye) r.j1 = mn(jde - 2, jte);
    if (g.periodic_x) { r.i0 = its; r.i1 = mn(ite, ide - 1); }
  }
  return r;
}
WG_HD bool in_r4(const R4 &r, int i, int j) { return i >= r.i0 && i <= r.i1 && j >= r.j0 && j <= r.j1; }

// ---- stresses (cal_titau_11_22_33 :5331, _12_21 :5456, _13_31 :5598, _23_32 :5749)
WG_HD float tau_n(const Geo &g, cfp xkx, cfp defor, cfp rho, int i, int k, int j) { return -A3(rho, i, k, j) * A3(xkx, i, k, j) * A3(defor, i, k, j); }
WG_HD float tau12(const Geo &g, cfp xkx, cfp defor, cfp rho, int i, int k, int j) {
  const float rhoavg = 0.25f * (A3(rho, i - 1, k, j) + A3(rho, i, k, j) + A3(rho, i - 1, k, j - 1) + A3(rho, i, k, j - 1));
  const float xkxavg = rhoavg * 0.25f * (A3(xkx, i - 1, k, j) + A3(xkx, i, k, j) + A3(xkx, i - 1, k, j - 1) + A3(xkx, i, k, j - 1));
  return -xkxavg * A3(defor, i, k, j);
}
WG_HD float tau13(const Geo &g, cfp defor, cfp xkx, cfp rho, int i, int k, int j, int ktf) {
  if (k == 1 || k == ktf + 1) return 0.0f;
  const float rhoavg = 0.5f * (g.fnm[k] * (A3(rho, i - 1, k, j) + A3(rho, i, k, j)) + g.fnp[k] * (A3(rho, i - 1, k - 1, j) + A3(rho, i, k - 1, j)));
  const float xkxavg = rhoavg * 0.5f * (g.fnm[k] * (A3(xkx, i, k, j) + A3(xkx, i - 1, k, j)) + g.fnp[k] * (A3(xkx, i, k - 1, j) + A3(xkx, i - 1, k - 1, j)));
  return -xkxavg * A3(defor, i, k, j);
}
WG_HD float tau23(const Geo &g, cfp defor, cfp xkx, cfp rho, int i, int k, int j, int ktf) {
  if (k == 1 || k == ktf + 1) return 0.0f;
  const float rhoavg = 0.5f * (g.fnm[k] * (A3(rho, i, k, j) + A3(rho, i, k, j - 1)) + g.fnp[k] * (A3(rho, i, k - 1, j) + A3(rho, i, k - 1, j - 1)));
  const float xkxavg = rhoavg * 0.5f * (g.fnm[k] * (A3(xkx, i, k, j) + A3(xkx, i, k, j - 1)) + g.fnp[k] * (A3(xkx, i, k - 1, j) + A3(xkx, i, k - 1, j - 1)));
  return -xkxavg * A3(defor, i, k, j);
}

// ---- u (vertical_diffusion_u_2 :4463, surface drag :4160-4236, horizontal_diffusion_u_2 :3118); thread box: the tile, k in 1..ktf
WG_HD void diff2_u_point(const Geo &g, const D2In &s, const R4 &r, int i, int k, int j, fp tendency) {
  const int ktf = g.kde - 1;
  const bool in = in_r4(r, i, j);
  const bool drag = k == 1 && i >= g.its && i <= g.ite && j >= g.jts && j <= min(g.jte, g.jde - 1);
  if (!in && !drag) return;
  float t = A3(tendency, i, k, j);
  if (in) {
    const float rdzu = -WG_G / (g.dnw[k]);
    if (k >= 2) t = t - rdzu * (tau13(g, s.defor13, s.xkmv, s.rho, i, k + 1, j, ktf) - tau13(g, s.defor13, s.xkmv, s.rho, i, k, j, ktf));
    else t = t - rdzu * (tau13(g, s.defor13, s.xkmv, s.rho, i, k + 1, j, ktf));
  }
  if (drag) {
    cfp u_2 = s.u, v_2 = s.v, rho = s.rho;
    const float vavg = (A3(v_2, i, 1, j) + A3(v_2, i, 1, j + 1) + A3(v_2, i - 1, 1, j) + A3(v_2, i - 1, 1, j + 1)) / 4;
    const float V0_u = sqrtf((A3(u_2, i, 1, j) * A3(u_2, i, 1, j)) + (vavg * vavg)) + 1.e-10f;
    float tao_xz;
    if (s.isfflx == 0) tao_xz = s.cd0 * V0_u * A3(u_2, i, 1, j);
    else {
      const float ustar = 0.5f * (A2(s.ustm, i, j) + A2(s.ustm, i - 1, j));
      tao_xz = ustar * ustar * A3(u_2, i, 1, j) / V0_u;
    }
    t = t + WG_G * tao_xz * 0.5f * (A3(rho, i, 1, j) + A3(rho, i - 1, 1, j)) / g.dnw[1];
  }
  if (in) {
    cfp zx = s.zx, zy = s.zy, rdzw = s.rdzw;
    auto t1avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (tau_n(g, s.xkmh, s.defor11, s.rho, i - 1, kk, j) + tau_n(g, s.xkmh, s.defor11, s.rho, i, kk, j)) +
                     g.fnp[kk] * (tau_n(g, s.xkmh, s.defor11, s.rho, i - 1, kk - 1, j) + tau_n(g, s.xkmh, s.defor11, s.rho, i, kk - 1, j)));
    };
    auto t2avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (tau12(g, s.xkmh, s.defor12, s.rho, i, kk, j + 1) + tau12(g, s.xkmh, s.defor12, s.rho, i, kk, j)) +
                     g.fnp[kk] * (tau12(g, s.xkmh, s.defor12, s.rho, i, kk - 1, j + 1) + tau12(g, s.xkmh, s.defor12, s.rho, i, kk - 1, j)));
    };
    const float zx_at_u = 0.5f * (A3(zx, i, k, j) + A3(zx, i, k + 1, j));
    const float zy_at_u = 0.125f * (A3(zy, i - 1, k, j) + A3(zy, i, k, j) + A3(zy, i - 1, k, j + 1) + A3(zy, i, k, j + 1) + A3(zy, i - 1, k + 1, j) +
                                    A3(zy, i, k + 1, j) + A3(zy, i - 1, k + 1, j + 1) + A3(zy, i, k + 1, j + 1));
    const float mrdx = A2(g.msfux, i, j) * g.rdx;
    const float mrdy = A2(g.msfuy, i, j) * g.rdy;
    const float tmpdz = (1.f / A3(rdzw, i, k, j) + 1.f / A3(rdzw, i - 1, k, j)) / 2.f;
    t = t + WG_G * tmpdz / g.dnw[k] *
                (mrdx * (tau_n(g, s.xkmh, s.defor11, s.rho, i, k, j) - tau_n(g, s.xkmh, s.defor11, s.rho, i - 1, k, j)) +
                 mrdy * (tau12(g, s.xkmh, s.defor12, s.rho, i, k, j + 1) - tau12(g, s.xkmh, s.defor12, s.rho, i, k, j)) -
                 A2(g.msfux, i, j) * zx_at_u * (t1avg(k + 1) - t1avg(k)) / tmpdz - A2(g.msfuy, i, j) * zy_at_u * (t2avg(k + 1) - t2avg(k)) / tmpdz);
  }
  A3(tendency, i, k, j) = t;
}

// ---- v (vertical_diffusion_v_2 :4576, surface drag, horizontal_diffusion_v_2 :3323)
WG_HD void diff2_v_point(const Geo &g, const D2In &s, const R4 &r, int i, int k, int j, fp tendency) {
  const int ktf = g.kde - 1;
  const bool in = in_r4(r, i, j);
  const bool drag = k == 1 && i >= g.its && i <= min(g.ite, g.ide - 1) && j >= g.jts && j <= g.jte;
  if (!in && !drag) return;
  float t = A3(tendency, i, k, j);
  if (in) {
    const float rdzv = -WG_G / g.dnw[k];
    if (k >= 2) t = t - rdzv * (tau23(g, s.defor23, s.xkmv, s.rho, i, k + 1, j, ktf) - tau23(g, s.defor23, s.xkmv, s.rho, i, k, j, ktf));
    else t = t - rdzv * (tau23(g, s.defor23, s.xkmv, s.rho, i, k + 1, j, ktf));
  }
  if (drag) {
    cfp u_2 = s.u, v_2 = s.v, rho = s.rho;
    const float uavg = (A3(u_2, i, 1, j) + A3(u_2, i, 1, j - 1) + A3(u_2, i + 1, 1, j) + A3(u_2, i + 1, 1, j - 1)) / 4;
    const float V0_v = sqrtf((A3(v_2, i, 1, j) * A3(v_2, i, 1, j)) + (uavg * uavg)) + 1.e-10f;
    float tao_yz;
    if (s.isfflx == 0) tao_yz = s.cd0 * V0_v * A3(v_2, i, 1, j);
    else {
      const float ustar = 0.5f * (A2(s.ustm, i, j) + A2(s.ustm, i, j - 1));
      tao_yz = ustar * ustar * A3(v_2, i, 1, j) / V0_v;
    }
    t = t + WG_G * tao_yz * 0.5f * (A3(rho, i, 1, j) + A3(rho, i, 1, j - 1)) / g.dnw[1];
  }
  if (in) {
    cfp zx = s.zx, zy = s.zy, rdzw = s.rdzw;
    auto t1avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (tau12(g, s.xkmh, s.defor12, s.rho, i + 1, kk, j) + tau12(g, s.xkmh, s.defor12, s.rho, i, kk, j)) +
                     g.fnp[kk] * (tau12(g, s.xkmh, s.defor12, s.rho, i + 1, kk - 1, j) + tau12(g, s.xkmh, s.defor12, s.rho, i, kk - 1, j)));
    };
    auto t2avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (tau_n(g, s.xkmh, s.defor22, s.rho, i, kk, j - 1) + tau_n(g, s.xkmh, s.defor22, s.rho, i, kk, j)) +
                     g.fnp[kk] * (tau_n(g, s.xkmh, s.defor22, s.rho, i, kk - 1, j - 1) + tau_n(g, s.xkmh, s.defor22, s.rho, i, kk - 1, j)));
    };
    const float zx_at_v = 0.125f * (A3(zx, i, k, j) + A3(zx, i + 1, k, j) + A3(zx, i, k, j - 1) + A3(zx, i + 1, k, j - 1) + A3(zx, i, k + 1, j) +
                                    A3(zx, i + 1, k + 1, j) + A3(zx, i, k + 1, j - 1) + A3(zx, i + 1, k + 1, j - 1));
    const float zy_at_v = 0.5f * (A3(zy, i, k, j) + A3(zy, i, k + 1, j));
    const float mrdx = A2(g.msfvx, i, j) * g.rdx;
    const float mrdy = A2(g.msfvy, i, j) * g.rdy;
    const float tmpdz = (1.f / A3(rdzw, i, k, j) + 1.f / A3(rdzw, i, k, j - 1)) / 2.f;
    t = t + WG_G * tmpdz / g.dnw[k] *
                (mrdy * (tau_n(g, s.xkmh, s.defor22, s.rho, i, k, j) - tau_n(g, s.xkmh, s.defor22, s.rho, i, k, j - 1)) +
                 mrdx * (tau12(g, s.xkmh, s.defor12, s.rho, i + 1, k, j) - tau12(g, s.xkmh, s.defor12, s.rho, i, k, j)) -
                 A2(g.msfvx, i, j) * zx_at_v * (t1avg(k + 1) - t1avg(k)) / tmpdz - A2(g.msfvy, i, j) * zy_at_v * (t2avg(k + 1) - t2avg(k)) / tmpdz);
  }
  A3(tendency, i, k, j) = t;
}

// ---- w (vertical_diffusion_w_2 :4688, horizontal_diffusion_w_2 :3519); thread box: the mass range, k in 2..ktf
WG_HD void diff2_w_point(const Geo &g, const D2In &s, int i, int k, int j, fp tendency) {
  const int ktf = g.kde - 1;
  cfp zx = s.zx, zy = s.zy, rdz = s.rdz;
  float t = A3(tendency, i, k, j);
  t = t + WG_G * (tau_n(g, s.xkmh, s.defor33, s.rho, i, k, j) - tau_n(g, s.xkmh, s.defor33, s.rho, i, k - 1, j)) / s.dn[k];
  auto t1avg = [&](int kk) -> float {   // kk in 1..ktf (ktf+1 is not read for k <= ktf)
    return 0.25f * (tau13(g, s.defor13, s.xkmv, s.rho, i + 1, kk + 1, j, ktf) + tau13(g, s.defor13, s.xkmv, s.rho, i, kk + 1, j, ktf) +
                    tau13(g, s.defor13, s.xkmv, s.rho, i + 1, kk, j, ktf) + tau13(g, s.defor13, s.xkmv, s.rho, i, kk, j, ktf));
  };
  auto t2avg = [&](int kk) -> float {
    return 0.25f * (tau23(g, s.defor23, s.xkmv, s.rho, i, kk + 1, j + 1, ktf) + tau23(g, s.defor23, s.xkmv, s.rho, i, kk + 1, j, ktf) +
                    tau23(g, s.defor23, s.xkmv, s.rho, i, kk, j + 1, ktf) + tau23(g, s.defor23, s.xkmv, s.rho, i, kk, j, ktf));
  };
  const float zx_at_w = 0.5f * (A3(zx, i, k, j) + A3(zx, i + 1, k, j));
  const float zy_at_w = 0.5f * (A3(zy, i, k, j) + A3(zy, i, k, j + 1));
  const float mrdx = A2(g.msftx, i, j) * g.rdx;
  const float mrdy = A2(g.msfty, i, j) * g.rdy;
  t = t + WG_G / (s.dn[k] * A3(rdz, i, k, j)) *
              (mrdx * (tau13(g, s.defor13, s.xkmv, s.rho, i + 1, k, j, ktf) - tau13(g, s.defor13, s.xkmv, s.rho, i, k, j, ktf)) +
               mrdy * (tau23(g, s.defor23, s.xkmv, s.rho, i, k, j + 1, ktf) - tau23(g, s.defor23, s.xkmv, s.rho, i, k, j, ktf)) -
               A2(g.msfty, i, j) * A3(rdz, i, k, j) * (zx_at_w * (t1avg(k) - t1avg(k - 1)) + zy_at_w * (t2avg(k) - t2avg(k - 1))));
  A3(tendency, i, k, j) = t;
}

// ---- scalars (vertical_diffusion_s :4789, the surface heat / moisture flux :4262-4353, horizontal_diffusion_s :3711)
// which: 0 theta (t_2; vertical part on t_2 - t_base unless mix_full_fields), 1 qv (vertical part on qv - qv_base unless
// mix_full_fields), 2 any other species.  base: origin-shifted 1-D array or null.  thread box: the tile's mass points, k in 1..ktf
// doing_tke: the tke species (K = xkmv / xkmh passed as xkhv / xkhh): each of the two terms is added twice (:4893-4903, :3985-3995)
WG_HD void diff2_s_point(const Geo &g, const D2In &s, const R4 &r, int which, cfp var, const float *base, cfp xkhv, cfp xkhh, int i, int k, int j,
                         fp tendency, int doing_tke = 0) {
  const int ktf = g.kde - 1, kte = g.kde;
  const bool in = in_r4(r, i, j);
  const bool flux = k == 1 && which != 2 && i <= min(g.ite, g.ide - 1) && j <= min(g.jte, g.jde - 1);
  if (!in && !flux) return;
  cfp rho = s.rho, rdz = s.rdz, rdzw = s.rdzw, zx = s.zx, zy = s.zy;
  float t = A3(tendency, i, k, j);
  if (in) {
    auto vm = [&](int kk) -> float { return base ? A3(var, i, kk, j) - base[kk] : A3(var, i, kk, j); };
    auto H3 = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      float xkxavg = g.fnm[kk] * A3(xkhv, i, kk, j) + g.fnp[kk] * A3(xkhv, i, kk - 1, j);
      xkxavg = xkxavg * (g.fnm[kk] * A3(rho, i, kk, j) + g.fnp[kk] * A3(rho, i, kk - 1, j));
      return -xkxavg * (vm(kk) - vm(kk - 1)) * A3(rdz, i, kk, j);
    };
    const float t0 = t;
    t = t + WG_G * (H3(k + 1) - H3(k)) / g.dnw[k];
    if (doing_tke) t = t0 + 2.f * (t - t0);
  }
  if (flux) {
    if (which == 0) {
      const float q = s.qv ? A3(s.qv, i, 1, j) : 0.f;
      const float cpm = WG_CP * (1.f + 0.8f * q);
      const float theta = (s.use_theta_m == 1 && s.qv) ? (A3(s.t, i, 1, j) + WG_T0) / (1.f + WG_RV / WG_RD * q) : A3(s.t, i, 1, j) + WG_T0;
      if (s.isfflx == 0 || s.isfflx == 2) {
        const float heat_flux = s.heat_flux;
        A2(s.hfx_out, i, j) = heat_flux * cpm * A3(rho, i, 1, j);
        if (s.use_theta_m == 1)
          t = t - WG_G * heat_flux * (1.f + WG_RVOVRD * q) * A3(rho, i, 1, j) / g.dnw[1] - WG_G * 1.61f * theta * A2(s.qfx, i, j) / g.dnw[1];
        else
          t = t - WG_G * heat_flux * A3(rho, i, 1, j) / g.dnw[1];
      } else {
        const float heat_flux = A2(s.hfx_in, i, j) / cpm;
        if (s.use_theta_m == 1) t = t - WG_G * heat_flux * (1.f + WG_RVOVRD * q) / g.dnw[1] - WG_G * 1.61f * theta * A2(s.qfx, i, j) / g.dnw[1];
        else t = t - WG_G * heat_flux / g.dnw[1];
      }
    } else if (s.isfflx != 0) {
      t = t - WG_G * A2(s.qfx, i, j) / g.dnw[1];
    }
  }
  if (in) {
    // var averaged to the w-level kk of the face between columns (ia,ja) and (ib,jb); B is the "minus" neighbour
    auto havg = [&](int ia, int ja, int ib, int jb, int kk) -> float {
      if (kk == 1)
        return 0.5f * (g.cf1 * A3(var, ia, 1, ja) + g.cf2 * A3(var, ia, 2, ja) + g.cf3 * A3(var, ia, 3, ja) + g.cf1 * A3(var, ib, 1, jb) +
                       g.cf2 * A3(var, ib, 2, jb) + g.cf3 * A3(var, ib, 3, jb));
      if (kk == ktf + 1)
        return 0.5f * (A3(var, ia, kte - 1, ja) + (A3(var, ia, kte - 1, ja) - A3(var, ia, kte - 2, ja)) * 0.5f * g.dnw[kte - 1] / s.dn[kte - 1] +
                       A3(var, ib, kte - 1, jb) + (A3(var, ib, kte - 1, jb) - A3(var, ib, kte - 2, jb)) * 0.5f * g.dnw[kte - 1] / s.dn[kte - 1]);
      return 0.5f * (g.fnm[kk] * (A3(var, ib, kk, jb) + A3(var, ia, kk, ja)) + g.fnp[kk] * (A3(var, ib, kk - 1, jb) + A3(var, ia, kk - 1, ja)));
    };
    auto H1 = [&](int ii, int kk) -> float {   // x-flux at the u-point ii
      const float xkxavg = 0.5f * (A3(xkhh, ii - 1, kk, j) + A3(xkhh, ii, kk, j)) * 0.5f * (A3(rho, ii - 1, kk, j) + A3(rho, ii, kk, j));
      const float tmpzx = 0.5f * (A3(zx, ii, kk, j) + A3(zx, ii, kk + 1, j));
      const float rdzu = 2.f / (1.f / A3(rdzw, ii, kk, j) + 1.f / A3(rdzw, ii - 1, kk, j));
      return -A2(g.msfux, ii, j) * xkxavg *
             (g.rdx * (A3(var, ii, kk, j) - A3(var, ii - 1, kk, j)) - tmpzx * (havg(ii, j, ii - 1, j, kk + 1) - havg(ii, j, ii - 1, j, kk)) * rdzu);
    };
    auto H2 = [&](int jj, int kk) -> float {   // y-flux at the v-point jj
      const float xkxavg = 0.5f * (A3(xkhh, i, kk, jj - 1) + A3(xkhh, i, kk, jj)) * 0.5f * (A3(rho, i, kk, jj - 1) + A3(rho, i, kk, jj));
      const float tmpzy = 0.5f * (A3(zy, i, kk, jj) + A3(zy, i, kk + 1, jj));
      const float rdzv = 2.f / (1.f / A3(rdzw, i, kk, jj) + 1.f / A3(rdzw, i, kk, jj - 1));
      return -A2(g.msfvy, i, jj) * xkxavg *
             (g.rdy * (A3(var, i, kk, jj) - A3(var, i, kk, jj - 1)) - tmpzy * (havg(i, jj, i, jj - 1, kk + 1) - havg(i, jj, i, jj - 1, kk)) * rdzv);
    };
    auto H1avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (H1(i + 1, kk) + H1(i, kk)) + g.fnp[kk] * (H1(i + 1, kk - 1) + H1(i, kk - 1)));
    };
    auto H2avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (H2(j + 1, kk) + H2(j, kk)) + g.fnp[kk] * (H2(j + 1, kk - 1) + H2(j, kk - 1)));
    };
    const float zx_at_m = 0.25f * (A3(zx, i, k, j) + A3(zx, i + 1, k, j) + A3(zx, i, k + 1, j) + A3(zx, i + 1, k + 1, j));
    const float zy_at_m = 0.25f * (A3(zy, i, k, j) + A3(zy, i, k, j + 1) + A3(zy, i, k + 1, j) + A3(zy, i, k + 1, j + 1));
    const float mrdx = A2(g.msftx, i, j) * g.rdx;
    const float mrdy = A2(g.msfty, i, j) * g.rdy;
    const float t0 = t;
    t = t + WG_G / (g.dnw[k] * A3(rdzw, i, k, j)) *
                (mrdx * (H1(i + 1, k) - H1(i, k)) + mrdy * (H2(j + 1, k) - H2(j, k)) -
                 A2(g.msftx, i, j) * zx_at_m * (H1avg(k + 1) - H1avg(k)) * A3(rdzw, i, k, j) -
                 A2(g.msfty, i, j) * zy_at_m * (H2avg(k + 1) - H2avg(k)) * A3(rdzw, i, k, j));
    if (doing_tke) t = t0 + 2.f * (t - t0);
  }
  A3(tendency, i, k, j) = t;
}

// phy_prep's density (module_big_step_utilities_em.F:4856) on the tile's mass points, k in 1..ktf
WG_HD void rho_point(const Geo &g, int i, int k, int j, cfp alt, cfp qv, fp rho) {
  A3(rho, i, k, j) = 1.f / A3(alt, i, k, j) * (1.f + (qv ? A3(qv, i, k, j) : 0.f));
}

}  // namespace wg
