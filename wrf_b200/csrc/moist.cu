// moist.cu -- calc_p_rho_phi (module_big_step_utilities_em.F:953-1223), the Kessler
// column kernel (phys/module_mp_kessler.F:8-244) fused with moist_physics_prep_em
// (:5394-5589) and moist_physics_finish_em (:5593-5784), set_w_surface
// (module_bc_em.F:1196-1270) and the end-of-step / start-of-run boundary fills.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include "libm_port.h"

namespace wg {

// LOG and x**y (vspow, frame/libmassv.F) are glibc's logf/powf in the reference build;
// libm_port.h restates those routines so the device result is bit-identical (see its header).
__device__ __forceinline__ float log_f(float x) { return wg_libm::logf_(x); }
__device__ __forceinline__ float pow_f(float x, float y) { return wg_libm::powf_(x, y); }
__device__ __forceinline__ float exp_f(float x) { return wg_libm::expf_(x); }

// full nonlinear diagnosis of al and p (non-hydrostatic branch)
__global__ void k_calc_p_rho_phi(const Geo g, Box b, fp al, cfp alb, cfp mu, cfp muts, cfp ph, cfp phb, fp p, cfp pb, cfp t, cfp qv) {
  TID3;
  const float mts = A2(muts, i, j);
  float alv;
  if (g.hypsometric_opt == 1) {
    alv = -1.f / (g.c1h[k] * mts + g.c2h[k]) * (A3(alb, i, k, j) * (g.c1h[k] * A2(mu, i, j)) + g.rdnw[k] * (A3(ph, i, k + 1, j) - A3(ph, i, k, j)));
  } else {
    const float pfu = g.c3f[k + 1] * mts + g.c4f[k + 1] + g.p_top;
    const float pfd = g.c3f[k] * mts + g.c4f[k] + g.p_top;
    const float phm = g.c3h[k] * mts + g.c4h[k] + g.p_top;
    alv = (A3(ph, i, k + 1, j) - A3(ph, i, k, j) + A3(phb, i, k + 1, j) - A3(phb, i, k, j)) / phm / log_f(pfd / pfu) - A3(alb, i, k, j);
  }
  A3(al, i, k, j) = alv;
  float pv;
  if (g.num_moist >= 1) {
    float temp;
    if (g.use_theta_m == 1) temp = (WG_RD * (WG_T0 + A3(t, i, k, j))) / (WG_P0 * (alv + A3(alb, i, k, j)));
    else {
      const float qvf = 1.f + WG_RVOVRD * A3(qv, i, k, j);
      temp = (WG_RD * (WG_T0 + A3(t, i, k, j)) * qvf) / (WG_P0 * (alv + A3(alb, i, k, j)));
    }
    pv = pow_f(temp, WG_CPOVCV) * WG_P0 - A3(pb, i, k, j);
  } else {
    pv = WG_P0 * pow_f((WG_RD * (WG_T0 + A3(t, i, k, j))) / (WG_P0 * (alv + A3(alb, i, k, j))), WG_CPOVCV) - A3(pb, i, k, j);
  }
  A3(p, i, k, j) = pv;
}

// calc_p_rho_phi with 16-byte accesses (mass columns: lanes lo..him of a group of four)
__global__ void k_calc_p_rho_phi_v4(const Geo g, Box b, fp al, cfp alb, cfp mu, cfp muts, cfp ph, cfp phb, fp p, cfp pb, cfp t, cfp qv) {
  TID3_V4;
  (void)hi;
  if (him < lo) return;
  float mts[4], m[4], ph1[4], ph0[4], pb1[4], pb0[4], ab[4], tv[4], q4[4], pbv[4], alv[4], pv[4];
  ld4(&A2(muts, i, j), mts); ld4(&A3(ph, i, k + 1, j), ph1); ld4(&A3(ph, i, k, j), ph0); ld4(&A3(alb, i, k, j), ab);
  ld4(&A3(t, i, k, j), tv); ld4(&A3(pb, i, k, j), pbv);
  if (g.hypsometric_opt == 1) ld4(&A2(mu, i, j), m);
  else { ld4(&A3(phb, i, k + 1, j), pb1); ld4(&A3(phb, i, k, j), pb0); }
  if (g.num_moist >= 1 && g.use_theta_m != 1) ld4(&A3(qv, i, k, j), q4);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (g.hypsometric_opt == 1) {
      alv[q] = -1.f / (g.c1h[k] * mts[q] + g.c2h[k]) * (ab[q] * (g.c1h[k] * m[q]) + g.rdnw[k] * (ph1[q] - ph0[q]));
    } else {
      const float pfu = g.c3f[k + 1] * mts[q] + g.c4f[k + 1] + g.p_top;
      const float pfd = g.c3f[k] * mts[q] + g.c4f[k] + g.p_top;
      const float phm = g.c3h[k] * mts[q] + g.c4h[k] + g.p_top;
      alv[q] = (ph1[q] - ph0[q] + pb1[q] - pb0[q]) / phm / log_f(pfd / pfu) - ab[q];
    }
    if (g.num_moist >= 1) {
      float temp;
      if (g.use_theta_m == 1) temp = (WG_RD * (WG_T0 + tv[q])) / (WG_P0 * (alv[q] + ab[q]));
      else {
        const float qvf = 1.f + WG_RVOVRD * q4[q];
        temp = (WG_RD * (WG_T0 + tv[q]) * qvf) / (WG_P0 * (alv[q] + ab[q]));
      }
      pv[q] = pow_f(temp, WG_CPOVCV) * WG_P0 - pbv[q];
    } else {
      pv[q] = WG_P0 * pow_f((WG_RD * (WG_T0 + tv[q])) / (WG_P0 * (alv[q] + ab[q])), WG_CPOVCV) - pbv[q];
    }
  }
  st4(&A3(al, i, k, j), alv, lo, him);
  st4(&A3(p, i, k, j), pv, lo, him);
}

void Solver::calc_p_rho_phi() {
  const Geo &q = g;
  static const bool vec4 = [] { const char *e = getenv("WRFGPU_SSP"); return !(e && !strcmp(e, "scalar")); }();
  if (vec4) {
    Box bx = box_v4(q, 1, q.kde - 1);
    bx.j1 = std::min(q.jte, q.jde - 1);
    LAUNCH3(this, k_calc_p_rho_phi_v4, bx, o("al"), o("alb"), o("mu_2"), o("muts"), o("ph_2"), o("phb"), o("p"), o("pb"), o("t_2"),
            c.num_moist >= 1 ? o("moist_qv") : o("t_2"));
    return;
  }
  LAUNCH3(this, k_calc_p_rho_phi, (Box{q.its, std::min(q.ite, q.ide - 1), 1, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)}), o("al"),
          o("alb"), o("mu_2"), o("muts"), o("ph_2"), o("phb"), o("p"), o("pb"), o("t_2"), c.num_moist >= 1 ? o("moist_qv") : o("t_2"));
}

// solve_em.F:3245-3343
void Solver::rk_phys_bc_dry_2() {
  halo("HALO_EM_D2_5");
  halo("HALO_EM_MOIST_E_7");
  std::vector<BcItem> it = {{"u_2", 'U'}, {"v_2", 'V'}, {"w_2", 'w'}, {"t_2", 'p'}, {"ph_2", 'w'}, {"mu_2", 't'}};
  for (auto &s : species) it.push_back({s.f, 'p'});
  bc(it);
}

// ---------------------------------------------------------------------------
struct MpArgs {
  fp t_2, qv, qc, qr, h_diabatic, qv_diabatic, qc_diabatic, rho, th_phy, pi_phy, z, dz8w, rainnc, rainncv;
  cfp al, alb, p, pb, ph, phb;
  float dt;
};

#define KMAX 256
// One thread per column: prep (rho, pi, z, dz8w, th) -> sedimentation -> per-level
// conversion/adjustment -> finish (theta_m reconstruction, h_diabatic).  rho, th_phy,
// pi_phy, z, dz8w are also written back so that they can be compared with the oracle.
__global__ void k_kessler(const Geo g, Box b, MpArgs a) {
  TID2;
  const int kte = g.kde - 1;   // microphysics runs on kts..kde-1
  const float dt = a.dt;
  const float c1 = .001f, c2 = .001f, c3 = 2.2f, c4 = .875f, max_cr_sedimentation = 0.75f;
  const float xlv = 2.5e6f, cp = WG_CP, svp1 = 0.6112f, svp2 = 17.67f, svp3 = 29.65f, svpt0 = 273.15f, rhowater = 1000.f;
  const float ep2 = WG_RD / WG_RV;
  const float f5 = svp2 * (svpt0 - svp3) * xlv / cp;
  const float rvrd = (WG_RV / WG_RD);
  // ---- moist_physics_prep_em
  float zw_k = (A3(a.ph, i, 1, j) + A3(a.phb, i, 1, j)) / WG_G;
  for (int k = 1; k <= kte; k++) {
    const float zw_kp1 = (A3(a.ph, i, k + 1, j) + A3(a.phb, i, k + 1, j)) / WG_G;
    A3(a.dz8w, i, k, j) = zw_kp1 - zw_k;
    const float qvv = A3(a.qv, i, k, j);
    float th;
    if (g.use_theta_m == 1) th = (A3(a.t_2, i, k, j) + WG_T0) / (1.f + rvrd * qvv);
    else th = A3(a.t_2, i, k, j) + WG_T0;
    A3(a.th_phy, i, k, j) = th;
    A3(a.h_diabatic, i, k, j) = th;
    A3(a.qv_diabatic, i, k, j) = qvv;
    A3(a.qc_diabatic, i, k, j) = A3(a.qc, i, k, j);
    A3(a.rho, i, k, j) = 1.f / (A3(a.al, i, k, j) + A3(a.alb, i, k, j));
    A3(a.pi_phy, i, k, j) = pow_f((A3(a.p, i, k, j) + A3(a.pb, i, k, j)) / WG_P0, WG_RCP);
    A3(a.z, i, k, j) = 0.5f * (zw_k + zw_kp1);
    zw_k = zw_kp1;
  }
  A3(a.dz8w, i, g.kde, j) = 0.f;
  // ---- kessler sedimentation (:93-194); column state in local memory
  float vt[KMAX], prodk[KMAX];
  const float rho1 = A3(a.rho, i, 1, j);
  float crmax = 0.f;
  for (int k = 1; k <= kte; k++) {
    const float qrk = A3(a.qr, i, k, j), rhok = A3(a.rho, i, k, j);
    prodk[k] = qrk;
    const float qrr = fmaxf(0.f, qrk * 0.001f * rhok);
    const float vtden = sqrtf(rho1 / rhok);
    vt[k] = 36.34f * pow_f(qrr, 0.1364f) * vtden;
    const float rdzw = 1.f / A3(a.dz8w, i, k, j);
    crmax = fmaxf(vt[k] * dt * rdzw, crmax);
  }
  int nfall = max(1, (int)lroundf(0.5f + crmax / max_cr_sedimentation));
  float dtfall = dt / (float)nfall;
  float time_sediment = dt;
  while (nfall > 0) {
    time_sediment = time_sediment - dtfall;
    {
      const float ppt = A3(a.rho, i, 1, j) * prodk[1] * vt[1] * dtfall / rhowater;
      A2(a.rainncv, i, j) = ppt * 1000.f;
      A2(a.rainnc, i, j) = A2(a.rainnc, i, j) + ppt * 1000.f;
    }
    for (int k = 1; k <= kte - 1; k++) {
      const float rhok = A3(a.rho, i, k, j), rhokp = A3(a.rho, i, k + 1, j);
      const float rdzk = 1.f / (A3(a.z, i, k + 1, j) - A3(a.z, i, k, j));
      const float factor = dtfall * rdzk / rhok;
      prodk[k] = prodk[k] - factor * (rhok * prodk[k] * vt[k] - rhokp * prodk[k + 1] * vt[k + 1]);
    }
    {
      const int k = kte;
      const float rdzk = 1.f / (A3(a.z, i, kte, j) - A3(a.z, i, kte - 1, j));
      const float factor = dtfall * rdzk;
      prodk[k] = prodk[k] - factor * prodk[k] * vt[k];
    }
    if (nfall > 1) {
      nfall = nfall - 1;
      crmax = 0.f;
      for (int k = 1; k <= kte; k++) {
        const float rhok = A3(a.rho, i, k, j);
        const float qrr = fmaxf(0.f, prodk[k] * 0.001f * rhok);
        const float vtden = sqrtf(rho1 / rhok);
        vt[k] = 36.34f * pow_f(qrr, 0.1364f) * vtden;
        const float rdzw = 1.f / A3(a.dz8w, i, k, j);
        crmax = fmaxf(vt[k] * time_sediment * rdzw, crmax);
      }
      const int nfall_new = max(1, (int)lroundf(0.5f + crmax / max_cr_sedimentation));
      if (nfall_new != nfall) { nfall = nfall_new; dtfall = time_sediment / (float)nfall; }
    } else {
      nfall = 0;
    }
  }
  // ---- conversion, saturation adjustment, evaporation (:202-236) + moist_physics_finish_em
  const float mp_tend_lim = 10.f;
  for (int k = 1; k <= kte; k++) {
    float qrv = A3(a.qr, i, k, j), qcv = A3(a.qc, i, k, j), qvv = A3(a.qv, i, k, j), tv = A3(a.th_phy, i, k, j);
    const float pii = A3(a.pi_phy, i, k, j);
    const float factorn = 1.0f / (1.f + c3 * dt * pow_f(fmaxf(0.f, qrv), c4));
    const float qrprod = qcv * (1.0f - factorn) + factorn * c1 * dt * fmaxf(qcv - c2, 0.f);
    const float rcgs = 0.001f * A3(a.rho, i, k, j);
    qcv = fmaxf(qcv - qrprod, 0.f);
    qrv = (qrv + prodk[k] - qrv);
    qrv = fmaxf(qrv + qrprod, 0.f);
    const float temp = pii * tv;
    const float pressure = 1.000e+05f * pow_f(pii, 1004.f / 287.f);
    const float gam = 2.5e+06f / (1004.f * pii);
    const float es = 1000.f * svp1 * exp_f(svp2 * (temp - svpt0) / (temp - svp3));
    const float qvs = ep2 * es / (pressure - es);
    const float pr = (qvv - qvs) / (1.f + pressure / (pressure - es) * qvs * f5 / ((temp - svp3) * (temp - svp3)));
    const float dimv = fmaxf(qvs - qvv, 0.f);
    const float ern = fminf(fminf(dt * (((1.6f + 124.9f * pow_f(rcgs * qrv, .2046f)) * pow_f(rcgs * qrv, .525f)) / (2.55e8f / (pressure * qvs) + 5.4e5f)) *
                                       (dimv / (rcgs * qvs)),
                                   fmaxf(-pr - qcv, 0.f)),
                            qrv);
    const float product = fmaxf(pr, -qcv);
    tv = tv + gam * (product - ern);
    qvv = fmaxf(qvv - product + ern, 0.f);
    qcv = qcv + product;
    qrv = qrv - ern;
    A3(a.th_phy, i, k, j) = tv;
    A3(a.qv, i, k, j) = qvv; A3(a.qc, i, k, j) = qcv; A3(a.qr, i, k, j) = qrv;
    // finish
    const float hd = A3(a.h_diabatic, i, k, j), qvd = A3(a.qv_diabatic, i, k, j), qcd = A3(a.qc_diabatic, i, k, j);
    float mpten = tv - hd;
    const float qvten = qvv - qvd, qcten = qcv - qcd;
    mpten = fminf(mp_tend_lim * dt, mpten);
    mpten = fmaxf(-mp_tend_lim * dt, mpten);
    if (g.use_theta_m == 1) {
      A3(a.t_2, i, k, j) = hd * (1.f + rvrd * qvd) + mpten * (1.f + rvrd * qvv) + rvrd * qvten * tv - WG_T0;
      A3(a.h_diabatic, i, k, j) = (mpten * (1.f + rvrd * qvv) + rvrd * qvten * tv) / dt;
    } else {
      A3(a.t_2, i, k, j) = A3(a.t_2, i, k, j) + mpten;
      A3(a.h_diabatic, i, k, j) = mpten / dt;
    }
    A3(a.qv_diabatic, i, k, j) = qvten / dt;
    A3(a.qc_diabatic, i, k, j) = qcten / dt;
  }
}

void Solver::microphysics() {
  if (c.mp_physics == 0) return;
  const Geo &q = g;
  if (q.kde > KMAX) throw std::runtime_error("wrfgpu: Kessler column kernel supports at most 256 levels");
  MpArgs a;
  a.t_2 = o("t_2"); a.qv = o("moist_qv"); a.qc = o("moist_qc"); a.qr = o("moist_qr"); a.h_diabatic = o("h_diabatic");
  a.qv_diabatic = o("qv_diabatic"); a.qc_diabatic = o("qc_diabatic"); a.rho = o("rho"); a.th_phy = o("th_phy");
  a.pi_phy = o("pi_phy"); a.z = o("z"); a.dz8w = o("dz8w"); a.rainnc = o("rainnc"); a.rainncv = o("rainncv");
  a.al = o("al"); a.alb = o("alb"); a.p = o("p"); a.pb = o("pb"); a.ph = o("ph_2"); a.phb = o("phb");
  a.dt = c.dt;
  // solve_em.F:3623-3645: the specified rows are left out (sz rows; all of x when periodic_x)
  const int sz = spec() ? c.spec_zone : 0, szx = c.periodic_x ? 0 : sz;
  LAUNCH2(this, k_kessler, (Box{std::max(q.its, q.ids + szx), std::min(q.ite, q.ide - 1 - szx), 1, 1, std::max(q.jts, q.jds + sz),
                                std::min(q.jte, q.jde - 1 - sz)}), a);
  halo("HALO_EM_THETAM");
  bc({{"h_diabatic", 'p'}});
}

// set_w_surface, module_bc_em.F:1196-1270 (fill_w_flag = .false.)
__global__ void k_set_w_surface(const Geo g, Box b, fp w, cfp u, cfp v) {
  TID2;
  int jm1_limit = g.jds, jp1_limit = g.jde - 1, im1_limit = g.ids, ip1_limit = g.ide - 1;
  if (g.periodic_x) { im1_limit = g.ids - 1; ip1_limit = g.ide; }
  if (g.periodic_y) { jm1_limit = g.jds - 1; jp1_limit = g.jde; }
  const int jm1 = max(j - 1, jm1_limit), jp1 = min(j + 1, jp1_limit), im1 = max(i - 1, im1_limit), ip1 = min(i + 1, ip1_limit);
  A3(w, i, 1, j) = A2(g.msfty, i, j) * .5f * g.rdy *
                       ((A2(g.ht, i, jp1) - A2(g.ht, i, j)) * (g.cf1 * A3(v, i, 1, j + 1) + g.cf2 * A3(v, i, 2, j + 1) + g.cf3 * A3(v, i, 3, j + 1)) +
                        (A2(g.ht, i, j) - A2(g.ht, i, jm1)) * (g.cf1 * A3(v, i, 1, j) + g.cf2 * A3(v, i, 2, j) + g.cf3 * A3(v, i, 3, j))) +
                   A2(g.msftx, i, j) * .5f * g.rdx *
                       ((A2(g.ht, ip1, j) - A2(g.ht, i, j)) * (g.cf1 * A3(u, i + 1, 1, j) + g.cf2 * A3(u, i + 1, 2, j) + g.cf3 * A3(u, i + 1, 3, j)) +
                        (A2(g.ht, i, j) - A2(g.ht, im1, j)) * (g.cf1 * A3(u, i, 1, j) + g.cf2 * A3(u, i, 2, j) + g.cf3 * A3(u, i, 3, j)));
}

// solve_em.F:4175-4759: calc_p_rho_phi, HALO_EM_D3_5, set_phys_bc_dry_2 + species, HALO_EM_C, set_w_surface
void Solver::end_of_step() {
  const Geo &q = g;
  // with mp_physics = 0 nothing changed al/p's inputs since the last RK substep: the
  // second calc_p_rho_phi (solve_em.F:4190) would rewrite identical values
  if (c.mp_physics) calc_p_rho_phi();
  halo("HALO_EM_D3_5");
  halo("HALO_EM_MOIST_E_7");
  std::vector<BcItem> it = {{"u_1", 'U'}, {"u_2", 'U'}, {"v_1", 'V'}, {"v_2", 'V'}, {"w_1", 'w'}, {"w_2", 'w'},
                            {"t_1", 'p'}, {"t_2", 'p'}, {"ph_1", 'w'}, {"ph_2", 'w'}, {"mu_1", 't'}, {"mu_2", 't'}};
  for (auto &s : species) it.push_back({s.f, 'p'});
  bc(it);
  if (spec()) spec_final();   // solve_em.F:4540-4722
  halo("HALO_EM_C");
  LAUNCH2(this, k_set_w_surface, (Box{q.its, std::min(q.ite, q.ide - 1), 1, 1, q.jts, std::min(q.jte, q.jde - 1)}), o("w_2"), o("u_2"),
          o("v_2"));
  halo("HALO_EM_E_5");
}

// the part of start_domain_em (dyn_em/start_em.F) the hot path needs before step 1
void Solver::start_domain() {
  halo("HALO_EM_INIT");
  std::vector<BcItem> it = {{"phb", 'w'}, {"pb", 't'}, {"alb", 't'}, {"p", 't'}, {"al", 't'}, {"alt", 't'}, {"t_init", 't'},
                            {"mub", 't'}, {"ht", 't'}, {"msftx", 't'}, {"msfty", 't'}, {"f", 't'}, {"e", 't'}, {"sina", 't'}, {"cosa", 't'},
                            {"msfux", 'u'}, {"msfuy", 'u'}, {"msfvx", 'v'}, {"msfvx_inv", 'v'}, {"msfvy", 'v'},
                            {"u_1", 'U'}, {"u_2", 'U'}, {"v_1", 'V'}, {"v_2", 'V'}, {"w_1", 'w'}, {"w_2", 'w'},
                            {"t_1", 'p'}, {"t_2", 'p'}, {"ph_1", 'w'}, {"ph_2", 'w'}, {"mu_1", 't'}, {"mu_2", 't'}};
  for (auto &s : species) { it.push_back({s.f, 'p'}); it.push_back({s.o, 'p'}); it.push_back({s.t, 'p'}); }
  bc(it);
}

}  // namespace wg
