// pd_fast.cu -- interior fast path of the positive-definite scalar advection advect_scalar_pd
// (dyn_em/module_advect_em.F:6069-7885): the two passes of scalar.cu (k_pd_scale: the limiter factor
// of a cell, k_pd_tend: the limited flux divergence) as k-marching kernels.  The face fluxes are the
// expressions of pd_face_x / pd_face_y / pd_face_z (scalar.cu) with every face 5th order (3rd order
// in the vertical), which holds at least 4 cells inside every non-periodic edge.
//
// ncu r1c: the gather forms evaluate six faces per cell and pass, each with its own loads, two real
// divisions for the Courant number and the boundary-order logic: 1100-1400 instructions per point,
// 80 % issue utilisation at 7 % DRAM.  Here a thread owns one (i,j) column and marches k:
//   * the lower z face is the upper one of the level before (carried, with its limited value);
//   * the west x face is computed by the thread, the east one comes from lane+1 by warp shuffle
//     (warps overlap by one column: 31 outputs per 32 lanes);
//   * dx, dy, dx/dt, dy/dt and the 2-D masses are column constants.
// Four face evaluations per cell and pass instead of six, ~20 loads instead of ~60.
#include "common.cuh"

namespace wg {

// statement functions of advect_scalar_pd (:6169-6186), as in scalar.cu
__device__ __forceinline__ float qflux4(float qm2, float qm1, float q0, float qp1) {
  return (7.f / 12.f) * (q0 + qm1) - (1.f / 12.f) * (qp1 + qm2);
}
__device__ __forceinline__ float qflux3(float qm2, float qm1, float q0, float qp1, float ua, float sg) {
  return qflux4(qm2, qm1, q0, qp1) + sg * sign1(ua) * (1.f / 12.f) * ((qp1 - qm2) - 3.f * (q0 - qm1));
}
__device__ __forceinline__ float qflux6(float qm3, float qm2, float qm1, float q0, float qp1, float qp2) {
  return (37.f / 60.f) * (q0 + qm1) - (2.f / 15.f) * (qp1 + qm2) + (1.f / 60.f) * (qp2 + qm3);
}
__device__ __forceinline__ float qflux5(float qm3, float qm2, float qm1, float q0, float qp1, float qp2, float ua, float sg) {
  return qflux6(qm3, qm2, qm1, q0, qp1, qp2) - sg * sign1(ua) * (1.f / 60.f) * ((qp2 - qm3) - 5.f * (qp1 - qm2) + 10.f * (q0 - qm1));
}
__device__ __forceinline__ float qflux_upwind(float qm1, float q0, float cr) {
  return 0.5f * fminf(1.0f, (cr + fabsf(cr))) * qm1 + 0.5f * fmaxf(-1.0f, (cr - fabsf(cr))) * q0;
}

struct PdFastIn {
  cfp q, q_old, ru, rv, rom, mut, mub, mu_old;
  float dt;
};

// everything of one column that does not depend on k
struct PdCol {
  float mut_c, mut_w, mut_s, mut_n;
  float dx, dy0, dy1, dxdt, dy0dt, dy1dt;   // face i, faces j and j+1
};
__device__ __forceinline__ PdCol pd_col(const Geo &g, const PdFastIn &s, int i, int j) {
  PdCol c;
  c.mut_c = A2(s.mut, i, j); c.mut_w = A2(s.mut, i - 1, j); c.mut_s = A2(s.mut, i, j - 1); c.mut_n = A2(s.mut, i, j + 1);
  c.dx = 2.f / (A2(g.msfty, i, j) + A2(g.msfty, i - 1, j)) / g.rdx;
  c.dy0 = 2.f / (A2(g.msftx, i, j) + A2(g.msftx, i, j - 1)) / g.rdy;
  c.dy1 = 2.f / (A2(g.msftx, i, j + 1) + A2(g.msftx, i, j)) / g.rdy;
  c.dxdt = c.dx / s.dt; c.dy0dt = c.dy0 / s.dt; c.dy1dt = c.dy1 / s.dt;
  return c;
}

// one level's horizontal operands
struct PdLev {
  float qx0, qx1, qx2, qx4, qx5;          // q(i-3..i-1), q(i+1..i+2) at (k, j)
  float qy0, qy1, qy2, qy4, qy5, qy6;     // q(i, k, j-3..j-1), q(i, k, j+1..j+3)
  float qo_w, qo_s, qo_n;                 // q_old at (i-1), (j-1), (j+1)
  float ru, rv0, rv1;                     // ru(i,k,j), rv(i,k,j), rv(i,k,j+1)
};
__device__ __forceinline__ PdLev pd_lev(const Geo &g, const PdFastIn &s, int i, int k, int j) {
  PdLev L;
  L.qx0 = A3(s.q, i - 3, k, j); L.qx1 = A3(s.q, i - 2, k, j); L.qx2 = A3(s.q, i - 1, k, j);
  L.qx4 = A3(s.q, i + 1, k, j); L.qx5 = A3(s.q, i + 2, k, j);
  L.qy0 = A3(s.q, i, k, j - 3); L.qy1 = A3(s.q, i, k, j - 2); L.qy2 = A3(s.q, i, k, j - 1);
  L.qy4 = A3(s.q, i, k, j + 1); L.qy5 = A3(s.q, i, k, j + 2); L.qy6 = A3(s.q, i, k, j + 3);
  L.qo_w = A3(s.q_old, i - 1, k, j); L.qo_s = A3(s.q_old, i, k, j - 1); L.qo_n = A3(s.q_old, i, k, j + 1);
  L.ru = A3(s.ru, i, k, j); L.rv0 = A3(s.rv, i, k, j); L.rv1 = A3(s.rv, i, k, j + 1);
  return L;
}

struct HiLo { float hi, lo; };
// x face i (:6640-6770, 5th order)
__device__ __forceinline__ HiLo pd_x(const PdCol &c, const PdLev &L, float q0, float qo_c, float c1h, float c2h, float dt, float sg) {
  const float mu = 0.5f * ((c1h * c.mut_c + c2h) + (c1h * c.mut_w + c2h));
  const float vel = L.ru;
  const float cr = vel * dt / c.dx / mu;
  HiLo f;
  f.lo = mu * c.dxdt * qflux_upwind(L.qo_w, qo_c, cr);
  f.hi = vel * qflux5(L.qx0, L.qx1, L.qx2, q0, L.qx4, L.qx5, vel, sg) - f.lo;
  return f;
}
// y faces j (m = 0) and j+1 (m = 1) (:6540-6630)
__device__ __forceinline__ HiLo pd_y0(const PdCol &c, const PdLev &L, float q0, float qo_c, float c1h, float c2h, float dt, float sg) {
  const float mu = 0.5f * ((c1h * c.mut_c + c2h) + (c1h * c.mut_s + c2h));
  const float vel = L.rv0;
  const float cr = vel * dt / c.dy0 / mu;
  HiLo f;
  f.lo = mu * c.dy0dt * qflux_upwind(L.qo_s, qo_c, cr);
  f.hi = vel * qflux5(L.qy0, L.qy1, L.qy2, q0, L.qy4, L.qy5, vel, sg) - f.lo;
  return f;
}
__device__ __forceinline__ HiLo pd_y1(const PdCol &c, const PdLev &L, float q0, float qo_c, float c1h, float c2h, float dt, float sg) {
  const float mu = 0.5f * ((c1h * c.mut_n + c2h) + (c1h * c.mut_c + c2h));
  const float vel = L.rv1;
  const float cr = vel * dt / c.dy1 / mu;
  HiLo f;
  f.lo = mu * c.dy1dt * qflux_upwind(qo_c, L.qo_n, cr);
  f.hi = vel * qflux5(L.qy1, L.qy2, q0, L.qy4, L.qy5, L.qy6, vel, sg) - f.lo;
  return f;
}
// z face kk between levels kk-1 and kk (:7595-7645); window q(kk-2..kk+1), q_old(kk-1), q_old(kk)
__device__ __forceinline__ HiLo pd_z(const Geo &g, const Coef &cf, int kk, float mut_c, float rom, float qm2, float qm1, float q0, float qp1,
                                     float qo_m1, float qo_0, float dt, float sg) {
  HiLo f;
  if (kk == 1 || kk == g.kde) { f.hi = 0.f; f.lo = 0.f; return f; }
  const float dz = 2.f / (cf.rdnw[kk] + cf.rdnw[kk - 1]);
  const float mu = 0.5f * ((cf.c1h[kk] * mut_c + cf.c2h[kk]) + (cf.c1h[kk] * mut_c + cf.c2h[kk]));
  const float vel = rom;
  const float cr = vel * dt / dz / mu;
  f.lo = mu * (dz / dt) * qflux_upwind(qo_m1, qo_0, cr);
  float h;
  if (kk == 2 || kk == g.kde - 1) h = rom * (cf.fnm[kk] * q0 + cf.fnp[kk] * qm1);
  else h = vel * qflux3(qm2, qm1, q0, qp1, -vel, sg);
  f.hi = h - f.lo;
  return f;
}

#ifndef PD_MINB
#define PD_MINB 6
#endif
#define PD_THREAD_SETUP                                                      \
  const int lane = threadIdx.x;                                              \
  const int i_raw = b.i0 + blockIdx.x * 31 + lane;                           \
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;                \
  if (j > b.j1) return;                           /* whole warps leave */    \
  const int i = min(i_raw, b.i1 + 1);             /* lanes past the box only feed the shuffle */ \
  const bool st = lane < 31 && i_raw <= b.i1;                                \
  const int ktf = g.kde - 1;                                                 \
  const float sg = g.sgn_ts, dt = s.dt;                                      \
  const Coef &cf = g

// pass 1: scale(i, 1..ktf, j) = limiter factor of the cell (:7681-7776)
__global__ void __launch_bounds__(128, PD_MINB) k_pd_scale_fast(const Geo g, Box b, PdFastIn s, fp scale) {
  PD_THREAD_SETUP;
  const PdCol c = pd_col(g, s, i, j);
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float mub = A2(s.mub, i, j), mu_old = A2(s.mu_old, i, j);
  float qm1 = 0.f, q0 = A3(s.q, i, 1, j), q1 = A3(s.q, i, 2, j);     // q(k-1), q(k), q(k+1)
  float qo0 = A3(s.q_old, i, 1, j);                                  // q_old(k)
  HiLo z0{0.f, 0.f};                                                 // face 1
  for (int k = 1; k <= ktf; k++) {
    const float c1h = cf.c1h[k], c2h = cf.c2h[k];
    const PdLev L = pd_lev(g, s, i, k, j);
    float q2 = 0.f, qo1 = 0.f;
    if (k + 2 <= ktf) q2 = A3(s.q, i, k + 2, j);
    if (k + 1 <= ktf) qo1 = A3(s.q_old, i, k + 1, j);
    const float rom1 = A3(s.rom, i, k + 1, j);
    const HiLo x0 = pd_x(c, L, q0, qo0, c1h, c2h, dt, sg);
    HiLo x1;
    x1.hi = __shfl_down_sync(0xffffffffu, x0.hi, 1);
    x1.lo = __shfl_down_sync(0xffffffffu, x0.lo, 1);
    const HiLo y0 = pd_y0(c, L, q0, qo0, c1h, c2h, dt, sg), y1 = pd_y1(c, L, q0, qo0, c1h, c2h, dt, sg);
    const HiLo z1 = pd_z(g, cf, k + 1, c.mut_c, rom1, qm1, q0, q1, q2, qo0, qo1, dt, sg);
    const float ph_low = ((c1h * mub + c2h) + (c1h * mu_old)) * qo0 -
                         dt * (msftx * msfty * (g.rdx * (x1.lo - x0.lo) + g.rdy * (y1.lo - y0.lo)) + msfty * cf.rdnw[k] * (z1.lo - z0.lo));
    const float flux_out = dt * ((msftx * msfty) * (g.rdx * (fmaxf(0.f, x1.hi) - fminf(0.f, x0.hi)) + g.rdy * (fmaxf(0.f, y1.hi) - fminf(0.f, y0.hi))) +
                                 msfty * cf.rdnw[k] * (fminf(0.f, z1.hi) - fmaxf(0.f, z0.hi)));
    float sc = 1.f;
    if (flux_out > ph_low) sc = fmaxf(0.f, ph_low / (flux_out + 1.e-20f));
    if (st) A3(scale, i, k, j) = sc;
    z0 = z1;
    qm1 = q0; q0 = q1; q1 = q2; qo0 = qo1;
  }
}

// pass 2: tendency(i, 1..ktf, j) = - limited flux divergence (:7778-7883).  A face's high-order part is
// scaled by its donor cell (:7764-7773), so the limited value of a face is the same seen from either side.
__global__ void __launch_bounds__(128, PD_MINB) k_pd_tend_fast(const Geo g, Box b, PdFastIn s, cfp scale, fp tend) {
  PD_THREAD_SETUP;
  const PdCol c = pd_col(g, s, i, j);
  const float msftx = A2(g.msftx, i, j);
  float qm1 = 0.f, q0 = A3(s.q, i, 1, j), q1 = A3(s.q, i, 2, j);
  float qo0 = A3(s.q_old, i, 1, j);
  float sc_k = A3(scale, i, 1, j);                                   // scale(i,k,j)
  float z0h = 0.f, z0l = 0.f;                                        // limited hi and lo of face k
  for (int k = 1; k <= ktf; k++) {
    const float c1h = cf.c1h[k], c2h = cf.c2h[k];
    const PdLev L = pd_lev(g, s, i, k, j);
    float q2 = 0.f, qo1 = 0.f, sc_k1 = 1.f;
    if (k + 2 <= ktf) q2 = A3(s.q, i, k + 2, j);
    if (k + 1 <= ktf) { qo1 = A3(s.q_old, i, k + 1, j); sc_k1 = A3(scale, i, k + 1, j); }
    const float rom1 = A3(s.rom, i, k + 1, j);
    const float sc_w = A3(scale, i - 1, k, j), sc_s = A3(scale, i, k, j - 1), sc_n = A3(scale, i, k, j + 1);
    float t = 0.f;
    // z (the sign convention is reversed: a positive flux through face kk leaves cell kk)
    const HiLo z1 = pd_z(g, cf, k + 1, c.mut_c, rom1, qm1, q0, q1, q2, qo0, qo1, dt, sg);
    float z1h = z1.hi;
    if (z1h > 0.f) z1h = sc_k1 * z1h; else if (z1h < 0.f) z1h = sc_k * z1h;
    t = t - cf.rdnw[k] * (z1h - z0h + z1.lo - z0l);
    // x: the west face here, the east face from the next lane
    {
      const HiLo x0 = pd_x(c, L, q0, qo0, c1h, c2h, dt, sg);
      float h0 = x0.hi;
      if (h0 > 0.f) h0 = sc_w * h0; else if (h0 < 0.f) h0 = sc_k * h0;
      const float h1 = __shfl_down_sync(0xffffffffu, h0, 1), lo1 = __shfl_down_sync(0xffffffffu, x0.lo, 1);
      t = t - msftx * (g.rdx * (h1 - h0 + lo1 - x0.lo));
    }
    // y
    {
      const HiLo y0 = pd_y0(c, L, q0, qo0, c1h, c2h, dt, sg), y1 = pd_y1(c, L, q0, qo0, c1h, c2h, dt, sg);
      float h0 = y0.hi, h1 = y1.hi;
      if (h0 > 0.f) h0 = sc_s * h0; else if (h0 < 0.f) h0 = sc_k * h0;
      if (h1 > 0.f) h1 = sc_k * h1; else if (h1 < 0.f) h1 = sc_n * h1;
      t = t - msftx * (g.rdy * (h1 - h0 + y1.lo - y0.lo));
    }
    if (st) A3(tend, i, k, j) = t;
    z0h = z1h; z0l = z1.lo;
    qm1 = q0; q0 = q1; q1 = q2; qo0 = qo1; sc_k = sc_k1;
  }
}

// pass 1 on `sb` (inside the limiter box), pass 2 on `tb` (inside the mass points)
void launch_pd_fast(Solver *s, int pass, const Box &bx, cfp q, cfp q_old, float dt, fp out) {
  if (empty(bx)) return;
  PdFastIn in;
  in.q = q; in.q_old = q_old; in.ru = s->o("ru_m"); in.rv = s->o("rv_m"); in.rom = s->o("ww_m");
  in.mut = s->o("muts"); in.mub = s->o("mub"); in.mu_old = s->o("mu_1"); in.dt = dt;
  const int ni = bx.i1 - bx.i0 + 1, nj = bx.j1 - bx.j0 + 1;
  dim3 block(32, 4), grid((unsigned)((ni + 30) / 31), (unsigned)((nj + 3) / 4));
  if (s->coef_dirty) s->sync_coef();
  if (pass == 1) {
    s->kpre("k_pd_scale_fast");
    k_pd_scale_fast<<<grid, block, 0, s->stream>>>(s->g, bx, in, out);
  } else {
    s->kpre("k_pd_tend_fast");
    k_pd_tend_fast<<<grid, block, 0, s->stream>>>(s->g, bx, in, s->o("pd_scale"), out);
  }
  s->kpost();
  s->launches++;
}

}  // namespace wg
