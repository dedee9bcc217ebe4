// acoustic.cu -- the split-explicit small step, dyn_em/module_small_step_em.F:
// small_step_prep :16, small_step_finish :295, calc_p_rho :438, calc_coef_w :570,
// advance_uv :654, advance_mu_t :969, advance_w :1178, sumflux :1473,
// calc_mu_uv_1 (module_big_step_utilities_em.F:184).
//
// Column kernels (calc_coef_w, advance_mu_t, advance_w) run one thread per (i,j)
// column: a warp covers 32 consecutive i, so every k level is one coalesced
// 128-byte access per field and the k-recurrences (mass-divergence prefix sum,
// Thomas forward/backward sweeps) live in registers.
#include "common.cuh"
#include "libm_port.h"

namespace wg {

// ---------------------------------------------------------------------------
// small_step_prep, 2-D part (:123-215)
__global__ void k_ssp_2d(const Geo g, Box b, int rk_step, cfp mub, fp mu_1, fp mu_2, cfp muu, fp muus, cfp muv, fp muvs,
                         fp muts, fp mudf, fp mu_save) {
  TID2;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1);
  if (rk_step == 1) {
    if (im && jm) {
      const float m2 = A2(mu_2, i, j);
      A2(mu_1, i, j) = m2;
      A2(mudf, i, j) = 0.f;
      A2(muts, i, j) = A2(mub, i, j) + m2;
      A2(mu_save, i, j) = m2;
      A2(mu_2, i, j) = 0.f;
    }
    if (jm) A2(muus, i, j) = A2(muu, i, j);
    if (im) A2(muvs, i, j) = A2(muv, i, j);
  } else {
    if (im && jm) {
      A2(muts, i, j) = A2(mub, i, j) + A2(mu_1, i, j);
      A2(mu_save, i, j) = A2(mu_2, i, j);
      A2(mu_2, i, j) = A2(mu_1, i, j) - A2(mu_2, i, j);
    }
    if (jm) A2(muus, i, j) = 0.5f * (A2(mub, i, j) + A2(mu_1, i, j) + A2(mub, i - 1, j) + A2(mu_1, i - 1, j));
    if (im) A2(muvs, i, j) = 0.5f * (A2(mub, i, j) + A2(mu_1, i, j) + A2(mub, i, j - 1) + A2(mu_1, i, j - 1));
  }
}

struct SspArgs {
  fp u_1, u_2, v_1, v_2, w_1, w_2, t_1, t_2, ph_1, ph_2;
  fp u_save, v_save, w_save, t_save, ph_save, ww_save, c2a;
  cfp ww, pb, p, alt, muu, muus, muv, muvs, mut, muts;
};

// small_step_prep, 3-D part (:128-170, :218-290)
__global__ void k_ssp_3d(const Geo g, Box b, int rk_step, SspArgs a) {
  TID3;
  const int kde = g.kde;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  if (km && jm) {   // u points
    const float u2 = A3(a.u_2, i, k, j);
    float u1;
    if (rk_step == 1) { u1 = u2; A3(a.u_1, i, k, j) = u2; } else u1 = A3(a.u_1, i, k, j);
    A3(a.u_save, i, k, j) = u2;
    A3(a.u_2, i, k, j) = ((c1h * A2(a.muus, i, j) + c2h) * u1 - (c1h * A2(a.muu, i, j) + c2h) * u2) / A2(g.msfuy, i, j);
  }
  if (km && im) {   // v points
    const float v2 = A3(a.v_2, i, k, j);
    float v1;
    if (rk_step == 1) { v1 = v2; A3(a.v_1, i, k, j) = v2; } else v1 = A3(a.v_1, i, k, j);
    A3(a.v_save, i, k, j) = v2;
    A3(a.v_2, i, k, j) = ((c1h * A2(a.muvs, i, j) + c2h) * v1 - (c1h * A2(a.muv, i, j) + c2h) * v2) * A2(g.msfvx_inv, i, j);
  }
  if (im && jm) {
    const float muts = A2(a.muts, i, j), mut = A2(a.mut, i, j);
    if (km) {
      const float t2 = A3(a.t_2, i, k, j);
      float t1;
      if (rk_step == 1) { t1 = t2; A3(a.t_1, i, k, j) = t2; } else t1 = A3(a.t_1, i, k, j);
      A3(a.c2a, i, k, j) = WG_CPOVCV * (A3(a.pb, i, k, j) + A3(a.p, i, k, j)) / A3(a.alt, i, k, j);
      A3(a.t_save, i, k, j) = t2;
      A3(a.t_2, i, k, j) = (c1h * muts + c2h) * t1 - (c1h * mut + c2h) * t2;
    }
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    const float w2 = A3(a.w_2, i, k, j), p2 = A3(a.ph_2, i, k, j);
    float w1, p1;
    if (rk_step == 1) { w1 = w2; p1 = p2; A3(a.w_1, i, k, j) = w2; A3(a.ph_1, i, k, j) = p2; }
    else { w1 = A3(a.w_1, i, k, j); p1 = A3(a.ph_1, i, k, j); }
    A3(a.w_save, i, k, j) = w2;
    A3(a.w_2, i, k, j) = ((c1f * muts + c2f) * w1 - (c1f * mut + c2f) * w2) / A2(g.msfty, i, j);
    A3(a.ph_save, i, k, j) = p2;
    A3(a.ph_2, i, k, j) = p1 - p2;
    A3(a.ww_save, i, k, j) = A3(a.ww, i, k, j);
  }
}

// calc_p_rho (:438-566), non-hydrostatic
__global__ void k_calc_p_rho(const Geo g, Box b, int step, fp al, fp p, cfp ph, cfp alt, cfp t_2, cfp t_1, cfp c2a, fp pm1,
                             cfp mu, cfp mut) {
  TID3;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  const float m = A2(mu, i, j), mt = A2(mut, i, j);
  const float altv = A3(alt, i, k, j), t1 = A3(t_1, i, k, j);
  const float alv = -1.f / (c1h * mt + c2h) * (altv * (c1h * m) + g.rdnw[k] * (A3(ph, i, k + 1, j) - A3(ph, i, k, j)));
  float pv = A3(c2a, i, k, j) * (altv * (A3(t_2, i, k, j) - (c1h * m) * t1) / ((c1h * mt + c2h) * (WG_T0 + t1)) - alv);
  A3(al, i, k, j) = alv;
  if (step == 0) {
    A3(pm1, i, k, j) = pv;
  } else {
    const float ptmp = pv;
    const float pold = A3(pm1, i, k, j);
    pv = pv + g.smdiv * (pv - pold);
    A3(pm1, i, k, j) = after_load(ptmp, pold);   // see common.cuh: keep the store behind the load
  }
  A3(p, i, k, j) = pv;
}

// calc_coef_w (:570-650): column recurrence for the tridiagonal factors
__global__ void k_calc_coef_w(const Geo g, Box b, float dts, fp a, fp alpha, fp gamma, cfp mut, cfp cqw, cfp c2a) {
  TID2;
  const int kde = g.kde;
  const float lid_flag = g.top_lid ? 0.f : 1.f;
  const float x0 = .5f * dts * WG_G * (1.f + g.epssm);
  const float cof = x0 * x0;
  const float mt = A2(mut, i, j);
  {
    const int k = kde - 1;
    A3(a, i, 2, j) = 0.f;
    A3(a, i, kde, j) = -2.f * cof * (g.rdnw[kde - 1] * g.rdnw[kde - 1]) * A3(c2a, i, kde - 1, j) * lid_flag /
                       ((g.c1h[k] * mt + g.c2h[k]) * (g.c1f[k] * mt + g.c2f[k]));
    A3(gamma, i, 1, j) = 0.f;
  }
  float gam_prev = 0.f;   // gamma(k-1)
  float c2a_km1 = A3(c2a, i, 1, j);
  for (int k = 2; k <= kde - 1; k++) {
    const float cq = A3(cqw, i, k, j);
    const float c2a_k = A3(c2a, i, k, j);
    float ak;
    if (k == 2) ak = 0.f;
    else ak = -cq * cof * g.rdn[k] * g.rdnw[k - 1] * c2a_km1 / ((g.c1h[k - 1] * mt + g.c2h[k - 1]) * (g.c1f[k - 1] * mt + g.c2f[k - 1]));
    if (k > 2) A3(a, i, k, j) = ak;
    const float bb = 1.f + cq * cof * g.rdn[k] *
                               (g.rdnw[k] * c2a_k / ((g.c1h[k] * mt + g.c2h[k]) * (g.c1f[k] * mt + g.c2f[k])) +
                                g.rdnw[k - 1] * c2a_km1 / ((g.c1h[k - 1] * mt + g.c2h[k - 1]) * (g.c1f[k] * mt + g.c2f[k])));
    const float cc = -cq * cof * g.rdn[k] * g.rdnw[k] * c2a_k / ((g.c1h[k] * mt + g.c2h[k]) * (g.c1f[k + 1] * mt + g.c2f[k + 1]));
    const float al = 1.f / (bb - ak * gam_prev);
    gam_prev = cc * al;
    A3(alpha, i, k, j) = al;
    A3(gamma, i, k, j) = gam_prev;
    c2a_km1 = c2a_k;
  }
  {
    const int k = kde;
    const float bb = 1.f + 2.f * cof * (g.rdnw[kde - 1] * g.rdnw[kde - 1]) * A3(c2a, i, kde - 1, j) /
                               ((g.c1h[k - 1] * mt + g.c2h[k - 1]) * (g.c1f[k] * mt + g.c2f[k]));
    const float cc = 0.f;
    const float al = 1.f / (bb - A3(a, i, kde, j) * gam_prev);
    A3(alpha, i, kde, j) = al;
    A3(gamma, i, kde, j) = cc * al;
  }
}

// ---------------------------------------------------------------------------
struct UvArgs {
  fp u, v;
  cfp ru_tend, rv_tend, p, pb, ph, php, alt, al, mu, muu, cqu, muv, cqv, mudf;
};
__device__ __forceinline__ float dpn_uv(const Geo &g, cfp p, int i, int kk, int j, int di, int dj) {
  const int i2 = i - di, j2 = j - dj;
  if (kk == 1)
    return .5f * (g.cf1 * (A3(p, i, 1, j) + A3(p, i2, 1, j2)) + g.cf2 * (A3(p, i, 2, j) + A3(p, i2, 2, j2)) + g.cf3 * (A3(p, i, 3, j) + A3(p, i2, 3, j2)));
  if (kk == g.kde) {
    if (!g.top_lid) return 0.f;
    return .5f * (g.cf1 * (A3(p, i2, g.kde - 1, j2) + A3(p, i, g.kde - 1, j)) + g.cf2 * (A3(p, i2, g.kde - 2, j2) + A3(p, i, g.kde - 2, j)) +
                  g.cf3 * (A3(p, i2, g.kde - 3, j2) + A3(p, i, g.kde - 3, j)));
  }
  return .5f * (g.fnm[kk] * (A3(p, i, kk, j) + A3(p, i2, kk, j2)) + g.fnp[kk] * (A3(p, i, kk - 1, j) + A3(p, i2, kk - 1, j2)));
}

// advance_uv (:654-965), one thread per (i,k,j); u and v points handled by the same thread
__global__ void k_advance_uv(const Geo g, Box b, float dts, UvArgs a) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const int i_start = its, i_end = min(ite, ide - 1), j_start = jts, j_end = min(jte, jde - 1);
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  const float dx = 1.f / g.rdx, dy = 1.f / g.rdy;
  // ---- u
  if (j <= j_end) {
    int i_start_up = i_start, i_end_up = ite, i_s_t = i_start, i_e_t = ite;
    if ((g.open_xs || g.sym_xs) && its == ids) i_start_up++;
    if ((g.open_xe || g.sym_xe) && ite == ide) i_end_up--;
    if (g.sym_xs && its == ids) i_s_t++;
    if (g.sym_xe && ite == ide) i_e_t--;
    float uv = A3(a.u, i, k, j);
    bool wr = false;
    if (i >= i_s_t && i <= i_e_t) { uv = uv + dts * A3(a.ru_tend, i, k, j); wr = true; }
    if (i >= i_start_up && i <= i_end_up) {
      const float mudf_xy = -g.emdiv * dx * (A2(a.mudf, i, j) - A2(a.mudf, i - 1, j)) / A2(g.msfuy, i, j);
      const float mr = A2(g.msfux, i, j) / A2(g.msfuy, i, j);
      float dpxy = mr * .5f * g.rdx * (c1h * A2(a.muu, i, j) + c2h) *
                   (((A3(a.ph, i, k + 1, j) - A3(a.ph, i - 1, k + 1, j)) + (A3(a.ph, i, k, j) - A3(a.ph, i - 1, k, j))) +
                    (A3(a.alt, i, k, j) + A3(a.alt, i - 1, k, j)) * (A3(a.p, i, k, j) - A3(a.p, i - 1, k, j)) +
                    (A3(a.al, i, k, j) + A3(a.al, i - 1, k, j)) * (A3(a.pb, i, k, j) - A3(a.pb, i - 1, k, j)));
      const float d1 = dpn_uv(g, a.p, i, k + 1, j, 1, 0), d0 = dpn_uv(g, a.p, i, k, j, 1, 0);
      dpxy = dpxy + mr * g.rdx * (A3(a.php, i, k, j) - A3(a.php, i - 1, k, j)) *
                        (g.rdnw[k] * (d1 - d0) - .5f * ((c1h * A2(a.mu, i - 1, j)) + (c1h * A2(a.mu, i, j))));
      uv = uv - dts * A3(a.cqu, i, k, j) * dpxy + (c1h * mudf_xy);
      wr = true;
    }
    if (wr) A3(a.u, i, k, j) = uv;
  }
  // ---- v
  if (i <= i_end) {
    int j_start_vp = j_start, j_end_vp = jte, j_s_t = j_start, j_e_t = jte;
    if ((g.open_ys || g.sym_ys) && jts == jds) j_start_vp++;
    if ((g.open_ye || g.sym_ye) && jte == jde) j_end_vp--;
    if (g.sym_ys && jts == jds) j_s_t++;
    if (g.sym_ye && jte == jde) j_e_t--;
    if (j >= j_s_t && j <= j_e_t) {
      float vv = A3(a.v, i, k, j) + dts * A3(a.rv_tend, i, k, j);
      if (j >= j_start_vp && j <= j_end_vp) {
        const float mudf_xy = -g.emdiv * dy * (A2(a.mudf, i, j) - A2(a.mudf, i, j - 1)) * A2(g.msfvx_inv, i, j);
        const float mr = A2(g.msfvy, i, j) / A2(g.msfvx, i, j);
        float dpxy = mr * .5f * g.rdy * (c1h * A2(a.muv, i, j) + c2h) *
                     (((A3(a.ph, i, k + 1, j) - A3(a.ph, i, k + 1, j - 1)) + (A3(a.ph, i, k, j) - A3(a.ph, i, k, j - 1))) +
                      (A3(a.alt, i, k, j) + A3(a.alt, i, k, j - 1)) * (A3(a.p, i, k, j) - A3(a.p, i, k, j - 1)) +
                      (A3(a.al, i, k, j) + A3(a.al, i, k, j - 1)) * (A3(a.pb, i, k, j) - A3(a.pb, i, k, j - 1)));
        const float d1 = dpn_uv(g, a.p, i, k + 1, j, 0, 1), d0 = dpn_uv(g, a.p, i, k, j, 0, 1);
        dpxy = dpxy + mr * g.rdy * (A3(a.php, i, k, j) - A3(a.php, i, k, j - 1)) *
                          (g.rdnw[k] * (d1 - d0) - .5f * ((c1h * A2(a.mu, i, j - 1)) + (c1h * A2(a.mu, i, j))));
        vv = vv - dts * A3(a.cqv, i, k, j) * dpxy + (c1h * mudf_xy);
      }
      A3(a.v, i, k, j) = vv;
    }
  }
}

// ---------------------------------------------------------------------------
struct MuTArgs {
  fp ww, mu, muave, muts, mudf, t, t_ave;
  cfp ww_1, u, u_1, v, v_1, mut, muu, muv, t_1, ft, mu_tend;
};
// advance_mu_t (:969-1174), column kernel
__global__ void k_advance_mu_t(const Geo g, Box b, float dts, MuTArgs a) {
  TID2;
  const int k_end = g.kde - 1;
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float muu0 = A2(a.muu, i, j), muu1 = A2(a.muu, i + 1, j), muv0 = A2(a.muv, i, j), muv1 = A2(a.muv, i, j + 1);
  const float my0 = A2(g.msfuy, i, j), my1 = A2(g.msfuy, i + 1, j), mvi0 = A2(g.msfvx_inv, i, j), mvi1 = A2(g.msfvx_inv, i, j + 1);
  const float mu_tend = A2(a.mu_tend, i, j);
  // pass 1: dvdxi(k) (parked in ww(k+1)) and its column integral dmdt
  float dmdt = 0.f;
  for (int k = 1; k <= k_end; k++) {
    const float c1h = g.c1h[k], c2h = g.c2h[k];
    const float dvdxi = msftx * msfty *
                        (g.rdy * ((A3(a.v, i, k, j + 1) + (c1h * muv1 + c2h) * A3(a.v_1, i, k, j + 1) * mvi1) -
                                  (A3(a.v, i, k, j) + (c1h * muv0 + c2h) * A3(a.v_1, i, k, j) * mvi0)) +
                         g.rdx * ((A3(a.u, i + 1, k, j) + (c1h * muu1 + c2h) * A3(a.u_1, i + 1, k, j) / my1) -
                                  (A3(a.u, i, k, j) + (c1h * muu0 + c2h) * A3(a.u_1, i, k, j) / my0)));
    dmdt = dmdt + g.dnw[k] * dvdxi;
    if (k < k_end) A3(a.ww, i, k + 1, j) = dvdxi;
  }
  {
    const float mu_old = A2(a.mu, i, j);
    const float mu_new = mu_old + dts * (dmdt + mu_tend);
    A2(a.mu, i, j) = mu_new;
    A2(a.mudf, i, j) = (dmdt + mu_tend);
    A2(a.muts, i, j) = A2(a.mut, i, j) + mu_new;
    A2(a.muave, i, j) = .5f * ((1.f + g.epssm) * mu_new + (1.f - g.epssm) * mu_old);
  }
  // pass 2: ww recurrence (:1109-1121), then the theta update which needs ww(k), ww(k+1)
  // ww(1) is not touched by the recurrence; ww(k) -= ww_1(k) for k = 1..k_end
  float w_rec = A3(a.ww, i, 1, j);             // ww(i,1,j) before the subtraction
  float wk = w_rec - A3(a.ww_1, i, 1, j);      // final ww(1)
  A3(a.ww, i, 1, j) = wk;
  float wdtn_k = 0.f;                          // wdtn(1) = 0
  float t1_km1 = 0.f, t1_k = A3(a.t_1, i, 1, j);
  for (int k = 1; k <= k_end; k++) {
    // ww(k+1)
    float wkp1 = 0.f, wdtn_kp1 = 0.f;
    float t1_kp1 = 0.f;
    if (k + 1 <= k_end) {
      const float dvdxi = A3(a.ww, i, k + 1, j);
      w_rec = w_rec - g.dnw[k] * ((g.c1h[k] * dmdt) + dvdxi + (g.c1h[k] * mu_tend)) / msfty;
      wkp1 = w_rec - A3(a.ww_1, i, k + 1, j);
      A3(a.ww, i, k + 1, j) = wkp1;
      t1_kp1 = A3(a.t_1, i, k + 1, j);
      wdtn_kp1 = wkp1 * (g.fnm[k + 1] * t1_kp1 + g.fnp[k + 1] * t1_k);
    }   // else wdtn(kde) = 0
    const float told = A3(a.t, i, k, j);
    A3(a.t_ave, i, k, j) = told;
    float tn = told + msfty * dts * A3(a.ft, i, k, j);
    tn = tn - dts * msfty *
                  (msftx * (.5f * g.rdy * (A3(a.v, i, k, j + 1) * (A3(a.t_1, i, k, j + 1) + t1_k) - A3(a.v, i, k, j) * (t1_k + A3(a.t_1, i, k, j - 1))) +
                            .5f * g.rdx * (A3(a.u, i + 1, k, j) * (A3(a.t_1, i + 1, k, j) + t1_k) - A3(a.u, i, k, j) * (t1_k + A3(a.t_1, i - 1, k, j)))) +
                   g.rdnw[k] * (wdtn_kp1 - wdtn_k));
    A3(a.t, i, k, j) = tn;
    wdtn_k = wdtn_kp1;
    t1_km1 = t1_k; t1_k = t1_kp1;
  }
  (void)t1_km1; (void)wk;
}

// ---------------------------------------------------------------------------
struct WArgs {
  fp w, ph, t_2ave;
  cfp rw_tend, ww, w_save, u, v, mut, muave, muts, t_2, t_1, ph_1, phb, ph_tend, c2a, cqw, alt, a, alpha, gamma;
};
// advance_w (:1178-1469), column kernel.  rhs(k) is kept in a rolling register
// window; the explicit w and rhs are parked in w / ph during the forward sweep
// and finished in the backward sweep.
__global__ void k_advance_w(const Geo g, Box b, float dts, WArgs a) {
  TID2;
  const int kde = g.kde, k_end = g.kde - 1;
  const float epssm = g.epssm;
  const float msfty = A2(g.msfty, i, j), msftx = A2(g.msftx, i, j);
  const float msft_inv = 1.f / msfty;
  const float mut = A2(a.mut, i, j), muave = A2(a.muave, i, j), muts = A2(a.muts, i, j);
  // ---- t_2ave (:1314-1317)
  for (int k = 1; k <= k_end; k++) {
    float tv = .5f * ((1.f + epssm) * A3(a.t_2, i, k, j) + (1.f - epssm) * A3(a.t_2ave, i, k, j));
    tv = (tv + (g.c1h[k] * muave) * WG_T0) / ((g.c1h[k] * muts + g.c2h[k]) * (WG_T0 + A3(a.t_1, i, k, j)));
    A3(a.t_2ave, i, k, j) = tv;
  }
  // ---- rhs(k), k = 1..kde (:1318-1371).  rhs(1) = 0.
  auto phi_dz = [&](int kk) -> float {   // wdwn(kk+1) of the reference, layer kk
    const float dphi = (A3(a.ph_1, i, kk + 1, j) - A3(a.ph_1, i, kk, j) + A3(a.phb, i, kk + 1, j) - A3(a.phb, i, kk, j));
    if (g.phi_adv_z == 2) return g.rdnw[kk] * dphi;
    return .5f * (A3(a.ww, i, kk + 1, j) + A3(a.ww, i, kk, j)) * g.rdnw[kk] * dphi;
  };
  auto rhs_at = [&](int k) -> float {
    if (k == 1) return 0.f;
    float r = dts * (A3(a.ph_tend, i, k, j) + .5f * WG_G * (1.f - epssm) * A3(a.w, i, k, j));
    if (k <= k_end) {
      const float wa = phi_dz(k), wb = phi_dz(k - 1);   // wdwn(k+1), wdwn(k)
      if (g.phi_adv_z == 2) r = r - dts * A3(a.ww, i, k, j) * (g.fnm[k] * wa + g.fnp[k] * wb);
      else r = r - dts * (g.fnm[k] * wa + g.fnp[k] * wb);
    }
    r = A3(a.ph, i, k, j) + msfty * r / (g.c1f[k] * mut + g.c2f[k]);
    if (g.top_lid && k == k_end + 1) r = 0.f;
    return r;
  };
  // surface w (:1384-1394)
  const float w1 = msfty * .5f * g.rdy *
                       ((A2(g.ht, i, j + 1) - A2(g.ht, i, j)) * (g.cf1 * A3(a.v, i, 1, j + 1) + g.cf2 * A3(a.v, i, 2, j + 1) + g.cf3 * A3(a.v, i, 3, j + 1)) +
                        (A2(g.ht, i, j) - A2(g.ht, i, j - 1)) * (g.cf1 * A3(a.v, i, 1, j) + g.cf2 * A3(a.v, i, 2, j) + g.cf3 * A3(a.v, i, 3, j))) +
                   msftx * .5f * g.rdx *
                       ((A2(g.ht, i + 1, j) - A2(g.ht, i, j)) * (g.cf1 * A3(a.u, i + 1, 1, j) + g.cf2 * A3(a.u, i + 1, 2, j) + g.cf3 * A3(a.u, i + 1, 3, j)) +
                        (A2(g.ht, i, j) - A2(g.ht, i - 1, j)) * (g.cf1 * A3(a.u, i, 1, j) + g.cf2 * A3(a.u, i, 2, j) + g.cf3 * A3(a.u, i, 3, j)));
  // forward sweep: explicit w (:1402-1431) fused with the Thomas forward elimination (:1433-1437)
  float rhs_km1 = 0.f;                 // rhs(1)
  float rhs_k = rhs_at(2);             // rhs(2)
  float ph_km1 = A3(a.ph, i, 1, j);
  float ph_k = A3(a.ph, i, 2, j);
  float wprev = w1;                    // w(k-1) after elimination
  for (int k = 2; k <= k_end; k++) {
    const float rhs_kp1 = rhs_at(k + 1);
    const float ph_kp1 = A3(a.ph, i, k + 1, j);
    float wv = A3(a.w, i, k, j) + dts * A3(a.rw_tend, i, k, j) +
               msft_inv * A3(a.cqw, i, k, j) *
                   (+.5f * dts * WG_G * g.rdn[k] *
                    (A3(a.c2a, i, k, j) * g.rdnw[k] / (g.c1h[k] * mut + g.c2h[k]) *
                         ((1.f + epssm) * (rhs_kp1 - rhs_k) + (1.f - epssm) * (ph_kp1 - ph_k)) -
                     A3(a.c2a, i, k - 1, j) * g.rdnw[k - 1] / (g.c1h[k - 1] * mut + g.c2h[k - 1]) *
                         ((1.f + epssm) * (rhs_k - rhs_km1) + (1.f - epssm) * (ph_k - ph_km1)))) +
               dts * WG_G * msft_inv *
                   (g.rdn[k] * (A3(a.c2a, i, k, j) * A3(a.alt, i, k, j) * A3(a.t_2ave, i, k, j) -
                                A3(a.c2a, i, k - 1, j) * A3(a.alt, i, k - 1, j) * A3(a.t_2ave, i, k - 1, j)) -
                    (g.c1f[k] * muave));
    wv = (wv - A3(a.a, i, k, j) * wprev) * A3(a.alpha, i, k, j);
    A3(a.w, i, k, j) = wv;
    A3(a.ph, i, k, j) = rhs_k;         // park rhs(k); ph(k) itself is no longer needed below k+1
    wprev = wv;
    rhs_km1 = rhs_k; rhs_k = rhs_kp1;
    ph_km1 = ph_k; ph_k = ph_kp1;
  }
  // top level k = kde (:1417-1431)
  float wtop;
  {
    const int k = kde;
    wtop = A3(a.w, i, k, j) + dts * A3(a.rw_tend, i, k, j) +
           msft_inv * (-.5f * dts * WG_G / (g.c1h[k - 1] * mut + g.c2h[k - 1]) * (g.rdnw[k - 1] * g.rdnw[k - 1]) * 2.f * A3(a.c2a, i, k - 1, j) *
                           ((1.f + epssm) * (rhs_k - rhs_km1) + (1.f - epssm) * (ph_k - ph_km1)) -
                       dts * WG_G * (2.f * g.rdnw[k - 1] * A3(a.c2a, i, k - 1, j) * A3(a.alt, i, k - 1, j) * A3(a.t_2ave, i, k - 1, j) + (g.c1f[k] * muave)));
    if (g.top_lid) wtop = 0.f;
    wtop = (wtop - A3(a.a, i, k, j) * wprev) * A3(a.alpha, i, k, j);
  }
  // backward sweep (:1439-1443) + damp_opt 3 (:1445-1458) + ph (:1460-1465)
  const float pi = 0x1.921fb6p+1f;   // 4.*atan(1.) in real*4
  const float dampmag = dts * g.dampcoef, hdepth = g.zdamp;
  const float htop = (A3(a.ph_1, i, kde, j) + A3(a.phb, i, kde, j)) / WG_G;
  auto finish = [&](int k, float wv, float rhs) {
    if (g.damp_opt == 3) {
      const float hk = (A3(a.ph_1, i, k, j) + A3(a.phb, i, k, j)) / WG_G;
      const float hbot = htop - hdepth;
      float dampwt = 0.f;
      if (hk >= hbot) dampwt = dampmag * wg_libm::sinf_(0.5f * pi * (hk - hbot) / hdepth) * wg_libm::sinf_(0.5f * pi * (hk - hbot) / hdepth);
      wv = (wv - dampwt * (g.c1f[k] * mut + g.c2f[k]) * A3(a.w_save, i, k, j)) / (1.f + dampwt);
    }
    A3(a.w, i, k, j) = wv;
    A3(a.ph, i, k, j) = rhs + msfty * .5f * dts * WG_G * (1.f + epssm) * wv / (g.c1f[k] * muts + g.c2f[k]);
  };
  A3(a.w, i, 1, j) = w1;
  float wnext = wtop;       // un-damped w(k+1) for the back substitution
  finish(kde, wtop, rhs_k);
  for (int k = k_end; k >= 2; k--) {
    const float wv = A3(a.w, i, k, j) - A3(a.gamma, i, k, j) * wnext;
    const float rhs = A3(a.ph, i, k, j);
    wnext = wv;
    finish(k, wv, rhs);
  }
}

// sumflux (:1473-1633)
__global__ void k_sumflux(const Geo g, Box b, int iteration, int nsmall, cfp ru, cfp rv, cfp ww, cfp u_lin, cfp v_lin, cfp ww_lin,
                          cfp muu, cfp muv, fp ru_m, fp rv_m, fp ww_m) {
  TID3;
  const int mini = min(g.ide - 1, g.ite), minj = min(g.jde - 1, g.jte), mink = g.kde - 1;
  const float nst = (float)nsmall;
  const bool du = (j <= minj && k <= mink), dv = (i <= mini && k <= mink), dw = (i <= mini && j <= minj);
  float a = A3(ru_m, i, k, j), bq = A3(rv_m, i, k, j), cq = A3(ww_m, i, k, j);
  if (iteration == 1) { a = 0.f; bq = 0.f; cq = 0.f; }
  if (du) a = a + A3(ru, i, k, j);
  if (dv) bq = bq + A3(rv, i, k, j);
  if (dw) cq = cq + A3(ww, i, k, j);
  if (iteration == nsmall) {
    if (du) a = a / nst + (g.c1h[k] * A2(muu, i, j) + g.c2h[k]) * A3(u_lin, i, k, j) / A2(g.msfuy, i, j);
    if (dv) bq = bq / nst + (g.c1h[k] * A2(muv, i, j) + g.c2h[k]) * A3(v_lin, i, k, j) * A2(g.msfvx_inv, i, j);
    if (dw) cq = cq / nst + A3(ww_lin, i, k, j);
  }
  A3(ru_m, i, k, j) = a; A3(rv_m, i, k, j) = bq; A3(ww_m, i, k, j) = cq;
}

// calc_mu_uv_1 (module_big_step_utilities_em.F:184-321)
__global__ void k_calc_mu_uv_1(const Geo g, Box b, cfp mu, fp muu, fp muv) {
  TID2;
  if (j <= min(g.jte, g.jde - 1)) {
    float r;
    if (i == g.ids) { const int q = g.periodic_x ? g.ids - 1 : g.ids; r = 0.5f * (A2(mu, i, j) + A2(mu, q, j)); }
    else if (i == g.ide) { const int q = g.periodic_x ? g.ide : g.ide - 1; r = 0.5f * (A2(mu, i - 1, j) + A2(mu, q, j)); }
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i - 1, j));
    A2(muu, i, j) = r;
  }
  if (i <= min(g.ite, g.ide - 1)) {
    float r;
    if (j == g.jds) { const int q = g.periodic_y ? g.jds - 1 : g.jds; r = 0.5f * (A2(mu, i, j) + A2(mu, i, q)); }
    else if (j == g.jde) { const int q = g.periodic_y ? g.jde : g.jde - 1; r = 0.5f * (A2(mu, i, j - 1) + A2(mu, i, q)); }
    else r = 0.5f * (A2(mu, i, j) + A2(mu, i, j - 1));
    A2(muv, i, j) = r;
  }
}

struct SsfArgs {
  fp u_2, v_2, w_2, t_2, ph_2, ww, mu_2;
  cfp ww1, mut, muts, muu, muus, muv, muvs, u_save, v_save, w_save, t_save, ph_save, mu_save, h_diabatic;
};
// small_step_finish (:295-434)
__global__ void k_ssf(const Geo g, Box b, int last, float dts, int nsmall, SsfArgs a) {
  TID3;
  const int kde = g.kde;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1), km = k <= kde - 1;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  if (km && im)
    A3(a.v_2, i, k, j) = (A2(g.msfvx, i, j) * A3(a.v_2, i, k, j) + A3(a.v_save, i, k, j) * (c1h * A2(a.muv, i, j) + c2h)) / (c1h * A2(a.muvs, i, j) + c2h);
  if (km && jm)
    A3(a.u_2, i, k, j) = (A2(g.msfuy, i, j) * A3(a.u_2, i, k, j) + A3(a.u_save, i, k, j) * (c1h * A2(a.muu, i, j) + c2h)) / (c1h * A2(a.muus, i, j) + c2h);
  if (im && jm) {
    const float mut = A2(a.mut, i, j), muts = A2(a.muts, i, j);
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    A3(a.w_2, i, k, j) = (A2(g.msfty, i, j) * A3(a.w_2, i, k, j) + A3(a.w_save, i, k, j) * (c1f * mut + c2f)) / (c1f * muts + c2f);
    A3(a.ph_2, i, k, j) = A3(a.ph_2, i, k, j) + A3(a.ph_save, i, k, j);
    A3(a.ww, i, k, j) = A3(a.ww, i, k, j) + A3(a.ww1, i, k, j);
    if (km) {
      if (!last)
        A3(a.t_2, i, k, j) = (A3(a.t_2, i, k, j) + A3(a.t_save, i, k, j) * (c1h * mut + c2h)) / (c1h * muts + c2h);
      else
        A3(a.t_2, i, k, j) = (A3(a.t_2, i, k, j) - dts * (float)nsmall * (c1h * mut + c2h) * A3(a.h_diabatic, i, k, j) +
                              A3(a.t_save, i, k, j) * (c1h * mut + c2h)) / (c1h * muts + c2h);
    }
    if (k == 1) A2(a.mu_2, i, j) = A2(a.mu_2, i, j) + A2(a.mu_save, i, j);
  }
}

// ---------------------------------------------------------------------------
void Solver::small_step_prep(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  Box tile2{q.its, q.ite, 1, 1, q.jts, q.jte};
  LAUNCH2(this, k_ssp_2d, tile2, rk_step, o("mub"), o("mu_1"), o("mu_2"), o("muu"), o("muus"), o("muv"), o("muvs"), o("muts"),
          o("mudf"), o("mu_save"));
  SspArgs a;
  a.u_1 = o("u_1"); a.u_2 = o("u_2"); a.v_1 = o("v_1"); a.v_2 = o("v_2"); a.w_1 = o("w_1"); a.w_2 = o("w_2");
  a.t_1 = o("t_1"); a.t_2 = o("t_2"); a.ph_1 = o("ph_1"); a.ph_2 = o("ph_2");
  a.u_save = o("u_save"); a.v_save = o("v_save"); a.w_save = o("w_save"); a.t_save = o("t_save"); a.ph_save = o("ph_save");
  a.ww_save = o("ww1"); a.c2a = o("c2a"); a.ww = o("ww"); a.pb = o("pb"); a.p = o("p"); a.alt = o("alt");
  a.muu = o("muu"); a.muus = o("muus"); a.muv = o("muv"); a.muvs = o("muvs"); a.mut = o("mut"); a.muts = o("muts");
  LAUNCH3(this, k_ssp_3d, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), rk_step, a);
  const int im = std::min(q.ite, q.ide - 1), jm = std::min(q.jte, q.jde - 1);
  LAUNCH3(this, k_calc_p_rho, (Box{q.its, im, 1, q.kde - 1, q.jts, jm}), 0, o("al"), o("p"), o("ph_2"), o("alt"), o("t_2"),
          o("t_save"), o("c2a"), o("pm1"), o("mu_2"), o("muts"));
  LAUNCH2(this, k_calc_coef_w, (Box{q.its, im, 1, 1, q.jts, jm}), dts_rk, o("a"), o("alpha"), o("gamma"), o("mut"), o("cqw"),
          o("c2a"));
  halo("HALO_EM_B");
  bc({{"ru_tend", 'u'}, {"rv_tend", 'v'}, {"ph_2", 'w'}, {"al", 'p'}, {"p", 'p'}, {"t_1", 'p'}, {"t_save", 't'},
      {"mu_1", 't'}, {"mu_2", 't'}, {"mudf", 't'}});
}

void Solver::advance_uv(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  UvArgs a;
  a.u = o("u_2"); a.v = o("v_2"); a.ru_tend = o("ru_tend"); a.rv_tend = o("rv_tend"); a.p = o("p"); a.pb = o("pb");
  a.ph = o("ph_2"); a.php = o("php"); a.alt = o("alt"); a.al = o("al"); a.mu = o("mu_2"); a.muu = o("muu"); a.cqu = o("cqu");
  a.muv = o("muv"); a.cqv = o("cqv"); a.mudf = o("mudf");
  LAUNCH3(this, k_advance_uv, (Box{q.its, q.ite, 1, q.kde - 1, q.jts, q.jte}), dts_rk, a);
}

void Solver::advance_mu_t(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  MuTArgs a;
  a.ww = o("ww"); a.mu = o("mu_2"); a.muave = o("muave"); a.muts = o("muts"); a.mudf = o("mudf"); a.t = o("t_2");
  a.t_ave = o("t_2save"); a.ww_1 = o("ww1"); a.u = o("u_2"); a.u_1 = o("u_save"); a.v = o("v_2"); a.v_1 = o("v_save");
  a.mut = o("mut"); a.muu = o("muu"); a.muv = o("muv"); a.t_1 = o("t_save"); a.ft = o("t_tend"); a.mu_tend = o("mu_tend");
  LAUNCH2(this, k_advance_mu_t, (Box{q.its, std::min(q.ite, q.ide - 1), 1, 1, q.jts, std::min(q.jte, q.jde - 1)}), dts_rk, a);
}

void Solver::advance_w(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  WArgs a;
  a.w = o("w_2"); a.ph = o("ph_2"); a.t_2ave = o("t_2save"); a.rw_tend = o("rw_tend"); a.ww = o("ww"); a.w_save = o("w_save");
  a.u = o("u_2"); a.v = o("v_2"); a.mut = o("mut"); a.muave = o("muave"); a.muts = o("muts"); a.t_2 = o("t_2");
  a.t_1 = o("t_save"); a.ph_1 = o("ph_save"); a.phb = o("phb"); a.ph_tend = o("ph_tend"); a.c2a = o("c2a"); a.cqw = o("cqw");
  a.alt = o("alt"); a.a = o("a"); a.alpha = o("alpha"); a.gamma = o("gamma");
  LAUNCH2(this, k_advance_w, (Box{q.its, std::min(q.ite, q.ide - 1), 1, 1, q.jts, std::min(q.jte, q.jde - 1)}), dts_rk, a);
}

void Solver::sumflux(int rk_step, int iteration) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  LAUNCH3(this, k_sumflux, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), iteration, ns, o("u_2"), o("v_2"), o("ww"), o("u_save"),
          o("v_save"), o("ww1"), o("muu"), o("muv"), o("ru_m"), o("rv_m"), o("ww_m"));
}

void Solver::calc_p_rho(int iteration) {
  const Geo &q = g;
  LAUNCH3(this, k_calc_p_rho, (Box{q.its, std::min(q.ite, q.ide - 1), 1, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)}), iteration,
          o("al"), o("p"), o("ph_2"), o("alt"), o("t_2"), o("t_save"), o("c2a"), o("pm1"), o("mu_2"), o("muts"));
  halo("HALO_EM_C2");
  bc({{"ph_2", 'w'}, {"al", 'p'}, {"p", 'p'}, {"muts", 't'}, {"mu_2", 't'}, {"mudf", 't'}});
}

void Solver::small_step_finish(int rk_step) {
  const Geo &q = g;
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  LAUNCH2(this, k_calc_mu_uv_1, (Box{q.its, q.ite, 1, 1, q.jts, q.jte}), o("muts"), o("muus"), o("muvs"));
  SsfArgs a;
  a.u_2 = o("u_2"); a.v_2 = o("v_2"); a.w_2 = o("w_2"); a.t_2 = o("t_2"); a.ph_2 = o("ph_2"); a.ww = o("ww"); a.mu_2 = o("mu_2");
  a.ww1 = o("ww1"); a.mut = o("mut"); a.muts = o("muts"); a.muu = o("muu"); a.muus = o("muus"); a.muv = o("muv"); a.muvs = o("muvs");
  a.u_save = o("u_save"); a.v_save = o("v_save"); a.w_save = o("w_save"); a.t_save = o("t_save"); a.ph_save = o("ph_save");
  a.mu_save = o("mu_save"); a.h_diabatic = o("h_diabatic");
  LAUNCH3(this, k_ssf, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), rk_step == c.rk_ord ? 1 : 0, dts_rk, ns, a);
  if (rk_step == c.rk_ord) bc({{"ru_m", 'u'}, {"rv_m", 'v'}, {"ww_m", 'w'}, {"mut", 't'}, {"muts", 't'}});
}

}  // namespace wg
