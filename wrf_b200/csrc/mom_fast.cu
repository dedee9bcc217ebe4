// mom_fast.cu -- interior fast path of the momentum tendencies ru_tend, rv_tend, rw_tend:
// advect_u/v/w (dyn_em/module_advect_em.F:126, :1530, :4364, h5/v3), horizontal_pressure_gradient
// (module_big_step_utilities_em.F:2183), pg_buoy_w (:2419), w_damp (:2503), coriolis (:3640),
// curvature (:4175) and rk_addtend_dry (module_em.F:959) -- the same sums, operand for operand, as
// the gather kernels of tend.cu (tend_u_pt<true>, tend_v_pt<true>, tend_w_pt<true>), which stay
// the reference form and still do the boundary strips.
//
// ncu r1c: the gather kernels are issue-bound -- ~1000 instructions and ~100 loads per point, 55 %
// issue utilisation at 15-20 % DRAM.  Here a thread owns one (i,j) column and marches k:
//   * everything with a vertical stencil rolls in registers: the 4-point window of the advected
//     field, the vertical face flux of level k+1 (it is the lower face of the next level), the
//     pressure at w levels dpn(k+1) -> dpn(k), ph, p, rw, and for w the k-1 values of ru, rv, u, v;
//   * the west x-face flux is computed by the thread, the east one comes from lane+1 by warp
//     shuffle (warps overlap by one column: 31 outputs per 32 lanes);
//   * all boundary logic is gone: the launch box holds only points at least 4 cells inside every
//     non-periodic edge (interior_box), where every face is 5th order and every loop range of the
//     reference contains the point.
// ~40 loads and ~300 instructions per point.  perturbation_coriolis (em_les) stays on the gather path.
#include "common.cuh"
#include "flux.cuh"

namespace wg {

struct MomIn {
  cfp u, v, w, ph, ru, rv, rw, ww;
  cfp p, pb, al, alt, php, cqu, cqv, cqw;
  cfp mut, muu, muv, mu, mub;
  cfp ru_tendf, rv_tendf, rw_tendf;
};

// resident blocks per SM the register allocation aims at (measured, profiles/README.md r1o/r1p: more
// warps beat more registers until the spills start)
#ifndef PH_MINB
#define PH_MINB 7
#endif
#ifndef MOM_MINB
#define MOM_MINB 6
#endif
#ifndef MOMW_MINB
#define MOMW_MINB 7
#endif
#define MOM_THREAD_SETUP                                                     \
  const int lane = threadIdx.x;                                              \
  const int i_raw = b.i0 + blockIdx.x * 31 + lane;                           \
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;                \
  if (j > b.j1) return;                           /* whole warps leave */    \
  const int i = min(i_raw, b.i1 + 1);             /* lanes past the box only feed the shuffle */ \
  const bool st = lane < 31 && i_raw <= b.i1;                                \
  const int kde = g.kde, ktf = g.kde - 1;                                    \
  const float sg = g.sgn_ts

// ===========================================================================
// ru_tend at (i, 1..kde-1, j); level kde is zero (rk_tendency zeroes the whole array, module_em.F:370-428)
__global__ void __launch_bounds__(128, MOM_MINB) k_mom_u_fast(const Geo g, Box b, MomIn s, fp ru_tend) {
  MOM_THREAD_SETUP;
  cfp u = s.u, ru = s.ru, rv = s.rv, rom = s.ww;
  // ---- column constants
  const float msfux = A2(g.msfux, i, j), msfuy = A2(g.msfuy, i, j);
  const float mrdy = msfux * g.rdy, mrdx = msfux * g.rdx;
  const float mr = msfux / msfuy;
  const float muu = A2(s.muu, i, j), mu_c = A2(s.mu, i, j), mu_w = A2(s.mu, i - 1, j);
  const float pA = mr * .5f * g.rdx, pB = mr * g.rdx;
  const float cA = mr * 0.5f * (A2(g.f, i, j) + A2(g.f, i - 1, j)) * 0.25f;
  const float cB = 0.5f * (A2(g.e, i, j) + A2(g.e, i - 1, j)) * 0.5f * (A2(g.cosa, i, j) + A2(g.cosa, i - 1, j)) * 0.25f;
  // vxgm(i), vxgm(i-1) (curvature :4230-4290): the map-factor differences are column constants
  const float dvx_c = A2(g.msfvx, i, j + 1) - A2(g.msfvx, i, j), duy_c = A2(g.msfuy, i + 1, j) - A2(g.msfuy, i, j);
  const float dvx_w = A2(g.msfvx, i - 1, j + 1) - A2(g.msfvx, i - 1, j), duy_w = A2(g.msfuy, i, j) - A2(g.msfuy, i - 1, j);
  // ---- rolling state
  float um1 = 0.f, u0 = A3(u, i, 1, j), u1 = A3(u, i, 2, j);        // u(k-1), u(k), u(k+1)
  float vf = 0.f;                                                   // vertical face k (face 1 carries no flux)
  float ph_c = A3(s.ph, i, 1, j), ph_w = A3(s.ph, i - 1, 1, j);
  float p_c = A3(s.p, i, 1, j), p_w = A3(s.p, i - 1, 1, j);
  float ps_km1 = 0.f, ps_k = p_w + p_c;                             // p(i-1,k)+p(i,k) at k-1, k
  float rw_c = A3(s.rw, i, 1, j), rw_w = A3(s.rw, i - 1, 1, j);
  // dpn(1): cf1..3 extrapolation to the surface (:2247-2252)
  float d0 = .5f * (g.cf1 * ps_k + g.cf2 * (A3(s.p, i - 1, 2, j) + A3(s.p, i, 2, j)) + g.cf3 * (A3(s.p, i - 1, 3, j) + A3(s.p, i, 3, j)));
  for (int k = 1; k <= ktf; k++) {
    const float c1h = g.c1h[k], c2h = g.c2h[k];
    // ---- this level's operands
    const float uy0 = A3(u, i, k, j - 3), uy1 = A3(u, i, k, j - 2), uy2 = A3(u, i, k, j - 1);
    const float uy4 = A3(u, i, k, j + 1), uy5 = A3(u, i, k, j + 2), uy6 = A3(u, i, k, j + 3);
    const float ux0 = A3(u, i - 3, k, j), ux1 = A3(u, i - 2, k, j), ux2 = A3(u, i - 1, k, j);
    const float ux4 = A3(u, i + 1, k, j), ux5 = A3(u, i + 2, k, j);
    float u2 = 0.f;
    if (k + 2 <= ktf) u2 = A3(u, i, k + 2, j);
    const float ru_c = A3(ru, i, k, j), ru_w = A3(ru, i - 1, k, j);
    const float rv_c0 = A3(rv, i, k, j), rv_w0 = A3(rv, i - 1, k, j), rv_c1 = A3(rv, i, k, j + 1), rv_w1 = A3(rv, i - 1, k, j + 1);
    const float rom_c = A3(rom, i, k + 1, j), rom_w = A3(rom, i - 1, k + 1, j);
    const float ph1_c = A3(s.ph, i, k + 1, j), ph1_w = A3(s.ph, i - 1, k + 1, j);
    float p1_c = 0.f, p1_w = 0.f;
    if (k + 1 <= ktf) { p1_c = A3(s.p, i, k + 1, j); p1_w = A3(s.p, i - 1, k + 1, j); }
    const float alt_c = A3(s.alt, i, k, j), alt_w = A3(s.alt, i - 1, k, j);
    const float al_c = A3(s.al, i, k, j), al_w = A3(s.al, i - 1, k, j);
    const float pb_c = A3(s.pb, i, k, j), pb_w = A3(s.pb, i - 1, k, j);
    const float php_c = A3(s.php, i, k, j), php_w = A3(s.php, i - 1, k, j);
    const float cqu = A3(s.cqu, i, k, j), tendf = A3(s.ru_tendf, i, k, j);
    const float rw1_c = A3(s.rw, i, k + 1, j), rw1_w = A3(s.rw, i - 1, k + 1, j);
    const float v_c0 = A3(s.v, i, k, j), v_c1 = A3(s.v, i, k, j + 1), v_w0 = A3(s.v, i - 1, k, j), v_w1 = A3(s.v, i - 1, k, j + 1);
    float t = 0.f;
    // ---- y flux divergence (:553-647), faces j and j+1
    {
      const float vel0 = 0.5f * (rv_c0 + rv_w0), vel1 = 0.5f * (rv_c1 + rv_w1);
      const float fq0 = vel0 * flux5(uy0, uy1, uy2, u0, uy4, uy5, vel0, sg);
      const float fq1 = vel1 * flux5(uy1, uy2, u0, uy4, uy5, uy6, vel1, sg);
      t = t - mrdy * (fq1 - fq0);
    }
    // ---- x flux divergence (:649-745): west face here, east face from the next lane
    {
      const float vel = 0.5f * (ru_c + ru_w);
      const float fq0 = vel * flux5(ux0, ux1, ux2, u0, ux4, ux5, vel, sg);
      const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
      t = t - mrdx * (fq1 - fq0);
    }
    // ---- vertical flux divergence (:1336-1498), face k+1
    float vf1;
    {
      const int kk = k + 1;
      if (kk == kde) vf1 = 0.f;
      else if (kk == 2 || kk == ktf) vf1 = 0.5f * (rom_c + rom_w) * (g.fnm[kk] * u1 + g.fnp[kk] * u0);
      else {
        const float vel = 0.5f * (rom_w + rom_c);
        vf1 = vel * flux3(um1, u0, u1, u2, -vel, sg);
      }
      t = t - g.rdnw[k] * (vf1 - vf);
    }
    // ---- horizontal pressure gradient, x (:2320-2400)
    const float ps_k1 = p1_w + p1_c;
    float d1;
    if (k + 1 == kde) d1 = g.top_lid ? .5f * (g.cfn * ps_k + g.cfn1 * ps_km1) : 0.f;
    else d1 = .5f * (g.fnm[k + 1] * ps_k1 + g.fnp[k + 1] * ps_k);
    {
      float dpx = pA * (c1h * muu + c2h) * ((ph1_c - ph1_w + ph_c - ph_w) + (alt_c + alt_w) * (p_c - p_w) + (al_c + al_w) * (pb_c - pb_w));
      dpx = dpx + pB * (php_c - php_w) * (g.rdnw[k] * (d1 - d0) - .5f * ((c1h * mu_w) + (c1h * mu_c)));
      t = t - cqu * dpx;
    }
    // ---- coriolis (:3700-3735) and curvature (:4292-4330)
    {
      const float rvavg = 0.25f * (rv_w1 + rv_c1 + rv_w0 + rv_c0);
      const float rwsum = (rw1_w + rw_w + rw1_c + rw_c);
      t = t + cA * (rv_w1 + rv_c1 + rv_w0 + rv_c0) - cB * rwsum;
      const float vx_c = 0.5f * (u0 + ux4) * dvx_c * g.rdy - 0.5f * (v_c0 + v_c1) * duy_c * g.rdx;
      const float vx_w = 0.5f * (ux2 + u0) * dvx_w * g.rdy - 0.5f * (v_w0 + v_w1) * duy_w * g.rdx;
      t = t + 0.5f * (vx_c + vx_w) * rvavg - u0 * WG_RERADIUS * 0.25f * rwsum;
    }
    // ---- rk_addtend_dry (module_em.F:1037-1043)
    t = t + tendf / msfuy;
    if (st) A3(ru_tend, i, k, j) = t;
    // ---- roll
    vf = vf1; d0 = d1;
    um1 = u0; u0 = u1; u1 = u2;
    ph_c = ph1_c; ph_w = ph1_w; p_c = p1_c; p_w = p1_w;
    ps_km1 = ps_k; ps_k = ps_k1;
    rw_c = rw1_c; rw_w = rw1_w;
  }
  if (st) A3(ru_tend, i, kde, j) = 0.f;
}

// ===========================================================================
// rv_tend at (i, 1..kde-1, j)
__global__ void __launch_bounds__(128, MOM_MINB) k_mom_v_fast(const Geo g, Box b, MomIn s, fp rv_tend) {
  MOM_THREAD_SETUP;
  cfp v = s.v, ru = s.ru, rv = s.rv, rom = s.ww;
  const float msfvy = A2(g.msfvy, i, j), msfvx = A2(g.msfvx, i, j), msfvx_inv = A2(g.msfvx_inv, i, j);
  const float mrdy = msfvy * g.rdy, mrdx = msfvy * g.rdx;
  const float mr = msfvy / msfvx;
  const float muv = A2(s.muv, i, j), mu_c = A2(s.mu, i, j), mu_s = A2(s.mu, i, j - 1);
  const float pA = mr * .5f * g.rdy, pB = mr * g.rdy;
  const float cA = mr * 0.5f * (A2(g.f, i, j) + A2(g.f, i, j - 1)) * 0.25f;
  const float cB = mr * 0.5f * (A2(g.e, i, j) + A2(g.e, i, j - 1)) * 0.5f * (A2(g.sina, i, j) + A2(g.sina, i, j - 1)) * 0.25f;
  const float cR = mr;   // mr*v*reradius*0.25*rwsum
  const float dvx_c = A2(g.msfvx, i, j + 1) - A2(g.msfvx, i, j), duy_c = A2(g.msfuy, i + 1, j) - A2(g.msfuy, i, j);
  const float dvx_s = A2(g.msfvx, i, j) - A2(g.msfvx, i, j - 1), duy_s = A2(g.msfuy, i + 1, j - 1) - A2(g.msfuy, i, j - 1);
  float vm1 = 0.f, v0 = A3(v, i, 1, j), v1 = A3(v, i, 2, j);
  float vf = 0.f;
  float ph_c = A3(s.ph, i, 1, j), ph_s = A3(s.ph, i, 1, j - 1);
  float p_c = A3(s.p, i, 1, j), p_s = A3(s.p, i, 1, j - 1);
  float ps_km1 = 0.f, ps_k = p_s + p_c;
  float rw_c = A3(s.rw, i, 1, j), rw_s = A3(s.rw, i, 1, j - 1);
  float d0 = .5f * (g.cf1 * ps_k + g.cf2 * (A3(s.p, i, 2, j - 1) + A3(s.p, i, 2, j)) + g.cf3 * (A3(s.p, i, 3, j - 1) + A3(s.p, i, 3, j)));
  for (int k = 1; k <= ktf; k++) {
    const float c1h = g.c1h[k], c2h = g.c2h[k];
    const float vy0 = A3(v, i, k, j - 3), vy1 = A3(v, i, k, j - 2), vy2 = A3(v, i, k, j - 1);
    const float vy4 = A3(v, i, k, j + 1), vy5 = A3(v, i, k, j + 2), vy6 = A3(v, i, k, j + 3);
    const float vx0 = A3(v, i - 3, k, j), vx1 = A3(v, i - 2, k, j), vx2 = A3(v, i - 1, k, j);
    const float vx4 = A3(v, i + 1, k, j), vx5 = A3(v, i + 2, k, j);
    float v2 = 0.f;
    if (k + 2 <= ktf) v2 = A3(v, i, k + 2, j);
    const float rv_m = A3(rv, i, k, j - 1), rv_c = A3(rv, i, k, j), rv_p = A3(rv, i, k, j + 1);
    const float ru_c0 = A3(ru, i, k, j), ru_s0 = A3(ru, i, k, j - 1), ru_c1 = A3(ru, i + 1, k, j), ru_s1 = A3(ru, i + 1, k, j - 1);
    const float rom_c = A3(rom, i, k + 1, j), rom_s = A3(rom, i, k + 1, j - 1);
    const float ph1_c = A3(s.ph, i, k + 1, j), ph1_s = A3(s.ph, i, k + 1, j - 1);
    float p1_c = 0.f, p1_s = 0.f;
    if (k + 1 <= ktf) { p1_c = A3(s.p, i, k + 1, j); p1_s = A3(s.p, i, k + 1, j - 1); }
    const float alt_c = A3(s.alt, i, k, j), alt_s = A3(s.alt, i, k, j - 1);
    const float al_c = A3(s.al, i, k, j), al_s = A3(s.al, i, k, j - 1);
    const float pb_c = A3(s.pb, i, k, j), pb_s = A3(s.pb, i, k, j - 1);
    const float php_c = A3(s.php, i, k, j), php_s = A3(s.php, i, k, j - 1);
    const float cqv = A3(s.cqv, i, k, j), tendf = A3(s.rv_tendf, i, k, j);
    const float rw1_c = A3(s.rw, i, k + 1, j), rw1_s = A3(s.rw, i, k + 1, j - 1);
    const float u_c0 = A3(s.u, i, k, j), u_c1 = A3(s.u, i + 1, k, j), u_s0 = A3(s.u, i, k, j - 1), u_s1 = A3(s.u, i + 1, k, j - 1);
    float t = 0.f;
    // ---- y flux divergence (:1960-2060), faces j and j+1
    {
      const float vel0 = 0.5f * (rv_c + rv_m), vel1 = 0.5f * (rv_p + rv_c);
      const float fq0 = vel0 * flux5(vy0, vy1, vy2, v0, vy4, vy5, vel0, sg);
      const float fq1 = vel1 * flux5(vy1, vy2, v0, vy4, vy5, vy6, vel1, sg);
      t = t - mrdy * (fq1 - fq0);
    }
    // ---- x flux divergence (:2062-2160): west face here, east face from the next lane
    {
      const float vel = 0.5f * (ru_c0 + ru_s0);
      const float fq0 = vel * flux5(vx0, vx1, vx2, v0, vx4, vx5, vel, sg);
      const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
      t = t - mrdx * (fq1 - fq0);
    }
    // ---- vertical flux divergence (:2802-2995), face k+1
    float vf1;
    {
      const int kk = k + 1;
      if (kk == kde) vf1 = 0.f;
      else if (kk == 2 || kk == ktf) vf1 = 0.5f * (rom_c + rom_s) * (g.fnm[kk] * v1 + g.fnp[kk] * v0);
      else {
        const float vel = 0.5f * (rom_c + rom_s);
        vf1 = vel * flux3(vm1, v0, v1, v2, -vel, sg);
      }
      t = t - mr * g.rdnw[k] * (vf1 - vf);
    }
    // ---- horizontal pressure gradient, y (:2230-2300)
    const float ps_k1 = p1_s + p1_c;
    float d1;
    if (k + 1 == kde) d1 = g.top_lid ? .5f * (g.cfn * ps_k + g.cfn1 * ps_km1) : 0.f;
    else d1 = .5f * (g.fnm[k + 1] * ps_k1 + g.fnp[k + 1] * ps_k);
    {
      float dpy = pA * (c1h * muv + c2h) * ((ph1_c - ph1_s + ph_c - ph_s) + (alt_c + alt_s) * (p_c - p_s) + (al_c + al_s) * (pb_c - pb_s));
      dpy = dpy + pB * (php_c - php_s) * (g.rdnw[k] * (d1 - d0) - .5f * ((c1h * mu_s) + (c1h * mu_c)));
      t = t - cqv * dpy;
    }
    // ---- coriolis (:3745-3765), curvature (:4340-4380)
    {
      const float rwsum = (rw1_s + rw_s + rw1_c + rw_c);
      const float ruavg = 0.25f * (ru_c0 + ru_c1 + ru_s0 + ru_s1);
      t = t - cA * (ru_c0 + ru_c1 + ru_s0 + ru_s1) + cB * rwsum;
      const float vx_c = 0.5f * (u_c0 + u_c1) * dvx_c * g.rdy - 0.5f * (v0 + vy4) * duy_c * g.rdx;
      const float vx_s = 0.5f * (u_s0 + u_s1) * dvx_s * g.rdy - 0.5f * (vy2 + v0) * duy_s * g.rdx;
      t = t - 0.5f * (vx_c + vx_s) * ruavg - cR * v0 * WG_RERADIUS * 0.25f * rwsum;
    }
    t = t + tendf * msfvx_inv;
    if (st) A3(rv_tend, i, k, j) = t;
    vf = vf1; d0 = d1;
    vm1 = v0; v0 = v1; v1 = v2;
    ph_c = ph1_c; ph_s = ph1_s; p_c = p1_c; p_s = p1_s;
    ps_km1 = ps_k; ps_k = ps_k1;
    rw_c = rw1_c; rw_s = rw1_s;
  }
  if (st) A3(rv_tend, i, kde, j) = 0.f;
}

// ===========================================================================
// rw_tend at (i, 1..kde, j).  Also pg_buoy_w's side effect cqw <- 1/(1+cqw) at levels 2..kde-1 (:2489):
// the column's thread is the only reader of cqw(i,:,j) here, and it carries cqw(k-1) in a register.
__global__ void __launch_bounds__(128, MOMW_MINB) k_mom_w_fast(const Geo g, Box b, MomIn s, fp rw_tend, fp cqw_inout) {
  MOM_THREAD_SETUP;
  (void)ktf;
  cfp w = s.w, ru = s.ru, rv = s.rv, rom = s.ww;
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float mrdy = msftx * g.rdy, mrdx = msftx * g.rdx;
  const float mr = msftx / msfty;
  const float mu = A2(s.mu, i, j), mub = A2(s.mub, i, j), mut = A2(s.mut, i, j), rmsf = (1.f / msfty);
  const float e = A2(g.e, i, j), cosa = A2(g.cosa, i, j), sina = A2(g.sina, i, j);
  // level 1: only rk_addtend_dry
  if (st) A3(rw_tend, i, 1, j) = 0.f + A3(s.rw_tendf, i, 1, j) / msfty;
  // rolling state at k = 2: values of level k-1 (= 1) and the vertical window of w
  float w0 = A3(w, i, 1, j), w1 = A3(w, i, 2, j), w2 = A3(w, i, 3, j);   // w(k-1), w(k), w(k+1)
  float ru_c_m = A3(ru, i, 1, j), ru_e_m = A3(ru, i + 1, 1, j), rv_c_m = A3(rv, i, 1, j), rv_n_m = A3(rv, i, 1, j + 1);
  float ru_c_mm = 0.f, rv_c_mm = 0.f, rv_n_mm = 0.f;                              // level k-2 (top level only)
  float u_c_m = A3(s.u, i, 1, j), u_e_m = A3(s.u, i + 1, 1, j), v_c_m = A3(s.v, i, 1, j), v_n_m = A3(s.v, i, 1, j + 1);
  float rom_m = A3(rom, i, 1, j), rom_k = A3(rom, i, 2, j);
  float p_m = A3(s.p, i, 1, j), cqw_m = 0.f;
  // vertical face 2 (:5996-6031): 2nd order
  float vf = 0.25f * (rom_k + rom_m) * (w1 + w0);
  for (int k = 2; k <= kde; k++) {
    const bool top = (k == kde);
    float w3 = 0.f;
    if (k + 2 <= kde) w3 = A3(w, i, k + 2, j);
    const float wy0 = A3(w, i, k, j - 3), wy1 = A3(w, i, k, j - 2), wy2 = A3(w, i, k, j - 1);
    const float wy4 = A3(w, i, k, j + 1), wy5 = A3(w, i, k, j + 2), wy6 = A3(w, i, k, j + 3);
    const float wx0 = A3(w, i - 3, k, j), wx1 = A3(w, i - 2, k, j), wx2 = A3(w, i - 1, k, j);
    const float wx4 = A3(w, i + 1, k, j), wx5 = A3(w, i + 2, k, j);
    const float tendf = A3(s.rw_tendf, i, k, j);
    float ru_c = 0.f, ru_e = 0.f, rv_c = 0.f, rv_n = 0.f, u_c = 0.f, u_e = 0.f, v_c = 0.f, v_n = 0.f, p_k = 0.f, cqw = 0.f, rom_1 = 0.f;
    if (!top) {
      ru_c = A3(ru, i, k, j); ru_e = A3(ru, i + 1, k, j); rv_c = A3(rv, i, k, j); rv_n = A3(rv, i, k, j + 1);
      u_c = A3(s.u, i, k, j); u_e = A3(s.u, i + 1, k, j); v_c = A3(s.v, i, k, j); v_n = A3(s.v, i, k, j + 1);
      p_k = A3(s.p, i, k, j); cqw = A3(s.cqw, i, k, j);
      rom_1 = A3(rom, i, k + 1, j);
    }
    const float fa = top ? (2.f - g.fnm[k - 1]) : g.fnm[k];
    const float fb = top ? g.fnp[k - 1] : g.fnp[k];
    float t = 0.f;
    // ---- y flux divergence, faces j and j+1: vel = fzm*rv(k)+fzp*rv(k-1), extrapolated at the top (:4847,4856)
    {
      const float vel0 = top ? (fa * rv_c_m - fb * rv_c_mm) : (fa * rv_c + fb * rv_c_m);
      const float vel1 = top ? (fa * rv_n_m - fb * rv_n_mm) : (fa * rv_n + fb * rv_n_m);
      const float fq0 = vel0 * flux5(wy0, wy1, wy2, w1, wy4, wy5, vel0, sg);
      const float fq1 = vel1 * flux5(wy1, wy2, w1, wy4, wy5, wy6, vel1, sg);
      t = t - mrdy * (fq1 - fq0);
    }
    // ---- x flux divergence: west face here, east face from the next lane
    {
      const float vel = top ? (fa * ru_c_m - fb * ru_c_mm) : (fa * ru_c + fb * ru_c_m);
      const float fq0 = vel * flux5(wx0, wx1, wx2, w1, wx4, wx5, vel, sg);
      const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
      t = t - mrdx * (fq1 - fq0);
    }
    // ---- vertical (:5996-6031), rdzu = rdn: face k+1 between w(k) and w(k+1)
    float vf1 = 0.f;
    if (!top) {
      const int kk = k + 1;
      if (kk == kde) vf1 = 0.25f * (rom_1 + rom_k) * (w2 + w1);
      else {
        const float vel = 0.5f * (rom_1 + rom_k);
        vf1 = vel * flux3(w0, w1, w2, w3, -vel, sg);
      }
      t = t - g.rdn[k] * (vf1 - vf);
    } else {
      t = t + 2.f * g.rdn[k - 1] * (vf);
    }
    // ---- pg_buoy_w (:2419-2499); cqw is still un-inverted here (k_invert_cqw runs after the tendencies)
    if (top) {
      const float cq1 = 1.f / (1.f + cqw_m);
      const float cq2 = cqw_m * cq1;
      t = t + rmsf * WG_G * (cq1 * 2.f * g.rdnw[k - 1] * (-p_m) - (g.c1f[k] * mu) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
    } else {
      const float cq1 = 1.f / (1.f + cqw);
      const float cq2 = cqw * cq1;
      if (st) A3(cqw_inout, i, k, j) = cq1;
      t = t + rmsf * WG_G * (cq1 * g.rdn[k] * (p_k - p_m) - (g.c1f[k] * mu) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
    }
    if (!top) {
      // ---- w_damp (:2635-2695)
      if (g.w_damping == 1) {
        const float mass = (g.c1f[k] * mut + g.c2f[k]);
        const float vert_cfl = fabsf(rom_k / mass * g.rdnw[k] * g.dt);
        if (vert_cfl > g.w_damp_on) t = t - sign1(w1) * 0.3f * (vert_cfl - g.w_crit_cfl) * mass;
      }
      // ---- coriolis (:3839-3844) and curvature (:4459-4463)
      const float fzm = g.fnm[k], fzp = g.fnp[k];
      t = t + e * (cosa * 0.5f * (fzm * (ru_c + ru_e) + fzp * (ru_c_m + ru_e_m)) -
                   mr * sina * 0.5f * (fzm * (rv_c + rv_n) + fzp * (rv_c_m + rv_n_m)));
      t = t + WG_RERADIUS * (0.5f * (fzm * (ru_c + ru_e) + fzp * (ru_c_m + ru_e_m)) * 0.5f * (fzm * (u_c + u_e) + fzp * (u_c_m + u_e_m)) +
                             mr * 0.5f * (fzm * (rv_c + rv_n) + fzp * (rv_c_m + rv_n_m)) * 0.5f * (fzm * (v_c + v_n) + fzp * (v_c_m + v_n_m)));
    }
    t = t + tendf / msfty;
    if (st) A3(rw_tend, i, k, j) = t;
    // ---- roll
    vf = vf1;
    w0 = w1; w1 = w2; w2 = w3;
    ru_c_mm = ru_c_m; rv_c_mm = rv_c_m; rv_n_mm = rv_n_m;
    ru_c_m = ru_c; ru_e_m = ru_e; rv_c_m = rv_c; rv_n_m = rv_n;
    u_c_m = u_c; u_e_m = u_e; v_c_m = v_c; v_n_m = v_n;
    rom_m = rom_k; rom_k = rom_1;
    p_m = p_k; cqw_m = cqw;
  }
  (void)rom_m;
}


// ===========================================================================
// ph_tend at (i, 1..kde, j): rhs_ph (module_big_step_utilities_em.F:1365-2181, 6th-order branch) +
// rk_addtend_dry, the sums of tend_ph_pt<true> (tend.cu) with the vertical operands rolling in registers
struct PhIn { cfp ph, phb, u, v, w, ww, mut, muu, muv, ph_tendf; };
__global__ void __launch_bounds__(128, PH_MINB) k_ph_fast(const Geo g, Box b, PhIn s, fp ph_tend) {
  TID2;
  const int kde = g.kde;
  cfp ph = s.ph, phb = s.phb;
  const float msfty = A2(g.msfty, i, j), mut = A2(s.mut, i, j);
  const float muv_n = A2(s.muv, i, j + 1), muv_c = A2(s.muv, i, j), muu_e = A2(s.muu, i + 1, j), muu_c = A2(s.muu, i, j);
  const float ya = A2(g.msfvy, i, j + 1), yb = A2(g.msfvy, i, j), xa = A2(g.msfux, i + 1, j), xb = A2(g.msfux, i, j);
  A3(ph_tend, i, 1, j) = 0.f + A3(s.ph_tendf, i, 1, j) / msfty;
  float ph_k = A3(ph, i, 2, j), phb_k = A3(phb, i, 2, j), ww_k = A3(s.ww, i, 2, j);
  float wd_k;   // wdwn(k) (:1459-1475)
  {
    const float ph_1 = A3(ph, i, 1, j), phb_1 = A3(phb, i, 1, j), ww_1 = A3(s.ww, i, 1, j);
    if (g.phi_adv_z == 2) wd_k = g.rdnw[1] * (ph_k - ph_1 + phb_k - phb_1);
    else wd_k = .5f * (ww_k + ww_1) * g.rdnw[1] * (ph_k - ph_1 + phb_k - phb_1);
  }
  float u_c_m = A3(s.u, i, 1, j), u_e_m = A3(s.u, i + 1, 1, j), v_c_m = A3(s.v, i, 1, j), v_n_m = A3(s.v, i, 1, j + 1);
  float u_c_mm = 0.f, u_e_mm = 0.f, v_c_mm = 0.f, v_n_mm = 0.f;
  for (int k = 2; k <= kde; k++) {
    const bool top = (k == kde);
    const float c1f = g.c1f[k], c2f = g.c2f[k];
    const float py1 = A3(ph, i, k, j + 1), pym1 = A3(ph, i, k, j - 1), py2 = A3(ph, i, k, j + 2), pym2 = A3(ph, i, k, j - 2);
    const float py3 = A3(ph, i, k, j + 3), pym3 = A3(ph, i, k, j - 3);
    const float by1 = A3(phb, i, k, j + 1), bym1 = A3(phb, i, k, j - 1), by2 = A3(phb, i, k, j + 2), bym2 = A3(phb, i, k, j - 2);
    const float by3 = A3(phb, i, k, j + 3), bym3 = A3(phb, i, k, j - 3);
    const float px1 = A3(ph, i + 1, k, j), pxm1 = A3(ph, i - 1, k, j), px2 = A3(ph, i + 2, k, j), pxm2 = A3(ph, i - 2, k, j);
    const float px3 = A3(ph, i + 3, k, j), pxm3 = A3(ph, i - 3, k, j);
    const float bx1 = A3(phb, i + 1, k, j), bxm1 = A3(phb, i - 1, k, j), bx2 = A3(phb, i + 2, k, j), bxm2 = A3(phb, i - 2, k, j);
    const float bx3 = A3(phb, i + 3, k, j), bxm3 = A3(phb, i - 3, k, j);
    const float wk = A3(s.w, i, k, j), tendf = A3(s.ph_tendf, i, k, j);
    float ph1 = 0.f, phb1 = 0.f, ww1 = 0.f, u_c = 0.f, u_e = 0.f, v_c = 0.f, v_n = 0.f;
    if (!top) {
      ph1 = A3(ph, i, k + 1, j); phb1 = A3(phb, i, k + 1, j); ww1 = A3(s.ww, i, k + 1, j);
      u_c = A3(s.u, i, k, j); u_e = A3(s.u, i + 1, k, j); v_c = A3(s.v, i, k, j); v_n = A3(s.v, i, k, j + 1);
    }
    float t = 0.f;
    float wd_k1 = 0.f;
    if (!top) {   // -omega dphi/deta (:1459-1494)
      if (g.phi_adv_z == 2) {
        wd_k1 = g.rdnw[k] * (ph1 - ph_k + phb1 - phb_k);
        t = t - ww_k * (g.fnm[k] * wd_k1 + g.fnp[k] * wd_k);
      } else {
        wd_k1 = .5f * (ww1 + ww_k) * g.rdnw[k] * (ph1 - ph_k + phb1 - phb_k);
        t = t - (g.fnm[k] * wd_k1 + g.fnp[k] * wd_k);
      }
    }
    // "gw" term (:1496-1508)
    t = t + (c1f * mut + c2f) * WG_G * wk / msfty;
    const float fac = top ? 0.5f : 0.25f;
    {   // y advection, 6th-order centred difference of ph+phb (:1783-1815)
      const float va = top ? (g.cfn * v_n_m + g.cfn1 * v_n_mm) : (v_n + v_n_m);
      const float vb = top ? (g.cfn * v_c_m + g.cfn1 * v_c_mm) : (v_c + v_c_m);
      const float ma = (c1f * muv_n + c2f), mb = (c1f * muv_c + c2f);
      t = t - (fac * g.rdy / msfty) *
                  ((ma * va * ya + mb * vb * yb) * (1.f / 60.f) *
                   (45.f * (py1 - pym1) - 9.f * (py2 - pym2) + (py3 - pym3) + 45.f * (by1 - bym1) - 9.f * (by2 - bym2) + (by3 - bym3)));
    }
    {   // x advection (:1940-1970)
      const float ua = top ? (g.cfn * u_e_m + g.cfn1 * u_e_mm) : (u_e + u_e_m);
      const float ub = top ? (g.cfn * u_c_m + g.cfn1 * u_c_mm) : (u_c + u_c_m);
      const float ma = (c1f * muu_e + c2f), mb = (c1f * muu_c + c2f);
      t = t - (fac * g.rdx / msfty) *
                  ((ma * ua * xa + mb * ub * xb) * (1.f / 60.f) *
                   (45.f * (px1 - pxm1) - 9.f * (px2 - pxm2) + (px3 - pxm3) + 45.f * (bx1 - bxm1) - 9.f * (bx2 - bxm2) + (bx3 - bxm3)));
    }
    t = t + tendf / msfty;
    A3(ph_tend, i, k, j) = t;
    wd_k = wd_k1;
    ph_k = ph1; phb_k = phb1; ww_k = ww1;
    u_c_mm = u_c_m; u_e_mm = u_e_m; v_c_mm = v_c_m; v_n_mm = v_n_m;
    u_c_m = u_c; u_e_m = u_e; v_c_m = v_c; v_n_m = v_n;
  }
}

// ---------------------------------------------------------------------------
static MomIn mom_in(Solver *s) {
  MomIn m;
  m.u = s->o("u_2"); m.v = s->o("v_2"); m.w = s->o("w_2"); m.ph = s->o("ph_2");
  m.ru = s->o("ru"); m.rv = s->o("rv"); m.rw = s->o("rw"); m.ww = s->o("ww");
  m.p = s->o("p"); m.pb = s->o("pb"); m.al = s->o("al"); m.alt = s->o("alt"); m.php = s->o("php");
  m.cqu = s->o("cqu"); m.cqv = s->o("cqv"); m.cqw = s->o("cqw");
  m.mut = s->o("mut"); m.muu = s->o("muu"); m.muv = s->o("muv"); m.mu = s->o("mu_2"); m.mub = s->o("mub");
  m.ru_tendf = s->o("ru_tendf"); m.rv_tendf = s->o("rv_tendf"); m.rw_tendf = s->o("rw_tendf");
  return m;
}

// which: 0 = ru_tend, 1 = rv_tend, 2 = rw_tend, 3 = ph_tend, on the interior box `in`
void launch_mom_fast(Solver *s, int which, const Box &in) {
  if (empty(in)) return;
  if (which == 3) {
    PhIn p;
    p.ph = s->o("ph_2"); p.phb = s->o("phb"); p.u = s->o("u_2"); p.v = s->o("v_2"); p.w = s->o("w_2"); p.ww = s->o("ww");
    p.mut = s->o("mut"); p.muu = s->o("muu"); p.muv = s->o("muv"); p.ph_tendf = s->o("ph_tendf");
    LAUNCH2(s, k_ph_fast, in, p, s->o("ph_tend"));
    return;
  }
  const int ni = in.i1 - in.i0 + 1, nj = in.j1 - in.j0 + 1;
  dim3 block(32, 4), grid((unsigned)((ni + 30) / 31), (unsigned)((nj + 3) / 4));
  if (s->coef_dirty) s->sync_coef();
  const MomIn m = mom_in(s);
  if (which == 0) {
    s->kpre("k_mom_u_fast");
    k_mom_u_fast<<<grid, block, 0, s->stream>>>(s->g, in, m, s->o("ru_tend"));
  } else if (which == 1) {
    s->kpre("k_mom_v_fast");
    k_mom_v_fast<<<grid, block, 0, s->stream>>>(s->g, in, m, s->o("rv_tend"));
  } else {
    s->kpre("k_mom_w_fast");
    k_mom_w_fast<<<grid, block, 0, s->stream>>>(s->g, in, m, s->o("rw_tend"), s->o("cqw"));
  }
  s->kpost();
  s->launches++;
}

}  // namespace wg
