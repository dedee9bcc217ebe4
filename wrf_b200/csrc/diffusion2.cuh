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
    if (g.oxs()) r.i0 = mx(ids + 1, its);
    if (g.oxe()) r.i1 = mn(ide - 1, ite);
    if (g.oys()) r.j0 = mx(jds + 1, jts);
    if (g.oye()) r.j1 = mn(jde - 2, jte);
    if (g.periodic_x) { r.i0 = its; r.i1 = ite; }
  } else if (kind == 'v') {
    r = {its, mn(ite, ide - 1), jts, jte};
    if (g.oxs()) r.i0 = mx(ids + 1, its);
    if (g.oxe()) r.i1 = mn(ide - 2, ite);
    if (g.oys()) r.j0 = mx(jds + 1, jts);
    if (g.oye()) r.j1 = mn(jde - 1, jte);
    if (g.periodic_x) { r.i0 = its; r.i1 = mn(ite, ide - 1); }
  } else {
    r = {its, mn(ite, ide - 1), jts, mn(jte, jde - 1)};
    if (g.oxs()) r.i0 = mx(ids + 1, its);
    if (g.oxe()) r.i1 = mn(ide - 2, ite);
    if (g.oys()) r.j0 = mx(jds + 1, jts);
    if (g.oye()) r.j1 = mn(jde - 2, jte);
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
//
// The pieces below are shared by the reference form (diff2_s_point: everything of a point evaluated on the spot) and by the
// k-marching kernel of diffusion2.cu (each face flux evaluated once per column and level and carried in registers).
// var averaged to the w-level kk of the face between columns (ia,ja) and (ib,jb); B is the "minus" neighbour
WG_HD float d2s_havg(const Geo &g, const D2In &s, cfp var, int ia, int ja, int ib, int jb, int kk) {
  const int ktf = g.kde - 1, kte = g.kde;
  if (kk == 1)
    return 0.5f * (g.cf1 * A3(var, ia, 1, ja) + g.cf2 * A3(var, ia, 2, ja) + g.cf3 * A3(var, ia, 3, ja) + g.cf1 * A3(var, ib, 1, jb) +
                   g.cf2 * A3(var, ib, 2, jb) + g.cf3 * A3(var, ib, 3, jb));
  if (kk == ktf + 1)
    return 0.5f * (A3(var, ia, kte - 1, ja) + (A3(var, ia, kte - 1, ja) - A3(var, ia, kte - 2, ja)) * 0.5f * g.dnw[kte - 1] / s.dn[kte - 1] +
                   A3(var, ib, kte - 1, jb) + (A3(var, ib, kte - 1, jb) - A3(var, ib, kte - 2, jb)) * 0.5f * g.dnw[kte - 1] / s.dn[kte - 1]);
  return 0.5f * (g.fnm[kk] * (A3(var, ib, kk, jb) + A3(var, ia, kk, ja)) + g.fnp[kk] * (A3(var, ib, kk - 1, jb) + A3(var, ia, kk - 1, ja)));
}
// x-flux at the u-point ii of level kk; hlo, hhi = d2s_havg(ii, j, ii-1, j, kk / kk+1)
WG_HD float d2s_H1(const Geo &g, const D2In &s, cfp var, cfp xkhh, int ii, int kk, int j, float hlo, float hhi) {
  cfp rho = s.rho, rdzw = s.rdzw, zx = s.zx;
  const float xkxavg = 0.5f * (A3(xkhh, ii - 1, kk, j) + A3(xkhh, ii, kk, j)) * 0.5f * (A3(rho, ii - 1, kk, j) + A3(rho, ii, kk, j));
  const float tmpzx = 0.5f * (A3(zx, ii, kk, j) + A3(zx, ii, kk + 1, j));
  const float rdzu = 2.f / (1.f / A3(rdzw, ii, kk, j) + 1.f / A3(rdzw, ii - 1, kk, j));
  return -A2(g.msfux, ii, j) * xkxavg * (g.rdx * (A3(var, ii, kk, j) - A3(var, ii - 1, kk, j)) - tmpzx * (hhi - hlo) * rdzu);
}
// y-flux at the v-point jj of level kk; hlo, hhi = d2s_havg(i, jj, i, jj-1, kk / kk+1)
WG_HD float d2s_H2(const Geo &g, const D2In &s, cfp var, cfp xkhh, int i, int kk, int jj, float hlo, float hhi) {
  cfp rho = s.rho, rdzw = s.rdzw, zy = s.zy;
  const float xkxavg = 0.5f * (A3(xkhh, i, kk, jj - 1) + A3(xkhh, i, kk, jj)) * 0.5f * (A3(rho, i, kk, jj - 1) + A3(rho, i, kk, jj));
  const float tmpzy = 0.5f * (A3(zy, i, kk, jj) + A3(zy, i, kk + 1, jj));
  const float rdzv = 2.f / (1.f / A3(rdzw, i, kk, jj) + 1.f / A3(rdzw, i, kk, jj - 1));
  return -A2(g.msfvy, i, jj) * xkxavg * (g.rdy * (A3(var, i, kk, jj) - A3(var, i, kk, jj - 1)) - tmpzy * (hhi - hlo) * rdzv);
}
// vertical flux at the w-level kk of the column
WG_HD float d2s_H3(const Geo &g, const D2In &s, cfp var, const float *base, cfp xkhv, int i, int kk, int j) {
  if (kk == 1 || kk == g.kde) return 0.f;
  cfp rho = s.rho, rdz = s.rdz;
  float xkxavg = g.fnm[kk] * A3(xkhv, i, kk, j) + g.fnp[kk] * A3(xkhv, i, kk - 1, j);
  xkxavg = xkxavg * (g.fnm[kk] * A3(rho, i, kk, j) + g.fnp[kk] * A3(rho, i, kk - 1, j));
  const float va = base ? A3(var, i, kk, j) - base[kk] : A3(var, i, kk, j);
  const float vb = base ? A3(var, i, kk - 1, j) - base[kk - 1] : A3(var, i, kk - 1, j);
  return -xkxavg * (va - vb) * A3(rdz, i, kk, j);
}
// the surface heat / moisture flux of level 1 added to t
WG_HD float d2s_surface_flux(const Geo &g, const D2In &s, int which, int i, int j, float t) {
  cfp rho = s.rho;
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
  return t;
}
// the horizontal term of a point from its face fluxes: h1w/h1e = H1(i / i+1, k), h2s/h2n = H2(j / j+1, k), a1lo/a1hi =
// H1avg(k / k+1), a2lo/a2hi = H2avg(k / k+1)
WG_HD float d2s_horizontal(const Geo &g, const D2In &s, int i, int k, int j, float h1w, float h1e, float h2s, float h2n, float a1lo,
                           float a1hi, float a2lo, float a2hi) {
  cfp rdzw = s.rdzw, zx = s.zx, zy = s.zy;
  const float zx_at_m = 0.25f * (A3(zx, i, k, j) + A3(zx, i + 1, k, j) + A3(zx, i, k + 1, j) + A3(zx, i + 1, k + 1, j));
  const float zy_at_m = 0.25f * (A3(zy, i, k, j) + A3(zy, i, k, j + 1) + A3(zy, i, k + 1, j) + A3(zy, i, k + 1, j + 1));
  const float mrdx = A2(g.msftx, i, j) * g.rdx;
  const float mrdy = A2(g.msfty, i, j) * g.rdy;
  return WG_G / (g.dnw[k] * A3(rdzw, i, k, j)) *
         (mrdx * (h1e - h1w) + mrdy * (h2n - h2s) - A2(g.msftx, i, j) * zx_at_m * (a1hi - a1lo) * A3(rdzw, i, k, j) -
          A2(g.msfty, i, j) * zy_at_m * (a2hi - a2lo) * A3(rdzw, i, k, j));
}
WG_HD void diff2_s_point(const Geo &g, const D2In &s, const R4 &r, int which, cfp var, const float *base, cfp xkhv, cfp xkhh, int i, int k, int j,
                         fp tendency, int doing_tke = 0) {
  const int ktf = g.kde - 1;
  const bool in = in_r4(r, i, j);
  const bool flux = k == 1 && which != 2 && i <= min(g.ite, g.ide - 1) && j <= min(g.jte, g.jde - 1);
  if (!in && !flux) return;
  float t = A3(tendency, i, k, j);
  if (in) {
    const float t0 = t;
    t = t + WG_G * (d2s_H3(g, s, var, base, xkhv, i, k + 1, j) - d2s_H3(g, s, var, base, xkhv, i, k, j)) / g.dnw[k];
    if (doing_tke) t = t0 + 2.f * (t - t0);
  }
  if (flux) t = d2s_surface_flux(g, s, which, i, j, t);
  if (in) {
    auto H1 = [&](int ii, int kk) -> float {
      return d2s_H1(g, s, var, xkhh, ii, kk, j, d2s_havg(g, s, var, ii, j, ii - 1, j, kk), d2s_havg(g, s, var, ii, j, ii - 1, j, kk + 1));
    };
    auto H2 = [&](int jj, int kk) -> float {
      return d2s_H2(g, s, var, xkhh, i, kk, jj, d2s_havg(g, s, var, i, jj, i, jj - 1, kk), d2s_havg(g, s, var, i, jj, i, jj - 1, kk + 1));
    };
    auto H1avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (H1(i + 1, kk) + H1(i, kk)) + g.fnp[kk] * (H1(i + 1, kk - 1) + H1(i, kk - 1)));
    };
    auto H2avg = [&](int kk) -> float {
      if (kk == 1 || kk == ktf + 1) return 0.f;
      return 0.5f * (g.fnm[kk] * (H2(j + 1, kk) + H2(j, kk)) + g.fnp[kk] * (H2(j + 1, kk - 1) + H2(j, kk - 1)));
    };
    const float t0 = t;
    t = t + d2s_horizontal(g, s, i, k, j, H1(i, k), H1(i + 1, k), H2(j, k), H2(j + 1, k), H1avg(k), H1avg(k + 1), H2avg(k), H2avg(k + 1));
    if (doing_tke) t = t0 + 2.f * (t - t0);
  }
  A3(tendency, i, k, j) = t;
}

// the same sums with a thread per column marching k: H1 / H2 of the level above are evaluated once (four per point and
// level instead of twelve) and the level's and the one below roll in registers, like the w-level averages they difference
WG_HD void diff2_s_column(const Geo &g, const D2In &s, const R4 &r, int which, cfp var, const float *base, cfp xkhv, cfp xkhh, int i, int j,
                          fp tendency, int doing_tke) {
  const int ktf = g.kde - 1;
  const bool in = in_r4(r, i, j);
  const bool flux = which != 2 && i <= min(g.ite, g.ide - 1) && j <= min(g.jte, g.jde - 1);
  if (!in) {
    if (flux) A3(tendency, i, 1, j) = d2s_surface_flux(g, s, which, i, j, A3(tendency, i, 1, j));
    return;
  }
  // w-level averages of the four faces at levels k (lo) and k+1 (hi), and the face fluxes at k-1 (m), k (c), k+1 (p)
  float aw_lo = d2s_havg(g, s, var, i, j, i - 1, j, 1), ae_lo = d2s_havg(g, s, var, i + 1, j, i, j, 1);
  float as_lo = d2s_havg(g, s, var, i, j, i, j - 1, 1), an_lo = d2s_havg(g, s, var, i, j + 1, i, j, 1);
  float aw_hi = d2s_havg(g, s, var, i, j, i - 1, j, 2), ae_hi = d2s_havg(g, s, var, i + 1, j, i, j, 2);
  float as_hi = d2s_havg(g, s, var, i, j, i, j - 1, 2), an_hi = d2s_havg(g, s, var, i, j + 1, i, j, 2);
  float h1w_m = 0.f, h1e_m = 0.f, h2s_m = 0.f, h2n_m = 0.f;
  float h1w_c = d2s_H1(g, s, var, xkhh, i, 1, j, aw_lo, aw_hi), h1e_c = d2s_H1(g, s, var, xkhh, i + 1, 1, j, ae_lo, ae_hi);
  float h2s_c = d2s_H2(g, s, var, xkhh, i, 1, j, as_lo, as_hi), h2n_c = d2s_H2(g, s, var, xkhh, i, 1, j + 1, an_lo, an_hi);
  float h3_lo = 0.f;   // H3(1) = 0
  for (int k = 1; k <= ktf; k++) {
    // level k+1 of the face fluxes (needed by H1avg(k+1) for k+1 <= ktf)
    float h1w_p = 0.f, h1e_p = 0.f, h2s_p = 0.f, h2n_p = 0.f;
    if (k + 1 <= ktf) {
      aw_lo = aw_hi; ae_lo = ae_hi; as_lo = as_hi; an_lo = an_hi;
      aw_hi = d2s_havg(g, s, var, i, j, i - 1, j, k + 2); ae_hi = d2s_havg(g, s, var, i + 1, j, i, j, k + 2);
      as_hi = d2s_havg(g, s, var, i, j, i, j - 1, k + 2); an_hi = d2s_havg(g, s, var, i, j + 1, i, j, k + 2);
      h1w_p = d2s_H1(g, s, var, xkhh, i, k + 1, j, aw_lo, aw_hi); h1e_p = d2s_H1(g, s, var, xkhh, i + 1, k + 1, j, ae_lo, ae_hi);
      h2s_p = d2s_H2(g, s, var, xkhh, i, k + 1, j, as_lo, as_hi); h2n_p = d2s_H2(g, s, var, xkhh, i, k + 1, j + 1, an_lo, an_hi);
    }
    float t = A3(tendency, i, k, j);
    {
      const float h3_hi = d2s_H3(g, s, var, base, xkhv, i, k + 1, j);
      const float t0 = t;
      t = t + WG_G * (h3_hi - h3_lo) / g.dnw[k];
      if (doing_tke) t = t0 + 2.f * (t - t0);
      h3_lo = h3_hi;
    }
    if (k == 1 && flux) t = d2s_surface_flux(g, s, which, i, j, t);
    {
      float a1lo = 0.f, a1hi = 0.f, a2lo = 0.f, a2hi = 0.f;
      if (k != 1) {
        a1lo = 0.5f * (g.fnm[k] * (h1e_c + h1w_c) + g.fnp[k] * (h1e_m + h1w_m));
        a2lo = 0.5f * (g.fnm[k] * (h2n_c + h2s_c) + g.fnp[k] * (h2n_m + h2s_m));
      }
      if (k + 1 != ktf + 1) {
        a1hi = 0.5f * (g.fnm[k + 1] * (h1e_p + h1w_p) + g.fnp[k + 1] * (h1e_c + h1w_c));
        a2hi = 0.5f * (g.fnm[k + 1] * (h2n_p + h2s_p) + g.fnp[k + 1] * (h2n_c + h2s_c));
      }
      const float t0 = t;
      t = t + d2s_horizontal(g, s, i, k, j, h1w_c, h1e_c, h2s_c, h2n_c, a1lo, a1hi, a2lo, a2hi);
      if (doing_tke) t = t0 + 2.f * (t - t0);
    }
    A3(tendency, i, k, j) = t;
    h1w_m = h1w_c; h1e_m = h1e_c; h2s_m = h2s_c; h2n_m = h2n_c;
    h1w_c = h1w_p; h1e_c = h1e_p; h2s_c = h2s_p; h2n_c = h2n_p;
  }
}

// phy_prep's density (module_big_step_utilities_em.F:4856) on the tile's mass points, k in 1..ktf
WG_HD void rho_point(const Geo &g, int i, int k, int j, cfp alt, cfp qv, fp rho) {
  A3(rho, i, k, j) = 1.f / A3(alt, i, k, j) * (1.f + (qv ? A3(qv, i, k, j) : 0.f));
}

}  // namespace wg
