// advect_fast.cu -- interior fast path of advect_scalar (dyn_em/module_advect_em.F:3029-4359, h5/v3).
//
// The gather kernels of tend.cu evaluate six face fluxes per point, each with its own loads: ncu
// (profiles/r1b) shows them issue-bound (integer pipe 48 %, 10 % DRAM).  Here a thread owns one (i,j)
// column and marches k:
//   * the vertical face flux of level k+1 is carried to the next level and the 4-point vertical
//     stencil rolls in registers;
//   * the west x-face is computed by the thread, the east one comes from lane+1 by warp shuffle
//     (warps overlap by one column: 31 outputs per 32 lanes).
// 4 flux evaluations and ~17 loads per point instead of 6 and ~40.  (A two-rows-per-thread form with an
// explicit one-level register prefetch shared the middle y face too, but at twice the registers it
// measured slower for theta, r1k.)  Only points at least 4 cells inside every non-periodic edge are
// done here (all faces 5th order, no boundary term); the edge strips go through the gather kernels,
// so the result is the same bits everywhere.
#include "common.cuh"
#include "flux.cuh"

namespace wg {

// vertical face kk = k+1 (:1470-1498 pattern): 0 at the lid, 2nd order next to the ends, 3rd order inside
__device__ __forceinline__ float adv_vface(const Geo &g, int kk, int ktf, float vel, float qm2, float qm1, float q0, float qp1, float sg) {
  if (kk == g.kde) return 0.f;
  if (kk == 2 || kk == ktf) return vel * (g.fnm[kk] * q0 + g.fnp[kk] * qm1);
  return vel * flux3(qm2, qm1, q0, qp1, -vel, sg);
}

// THETA: also rk_addtend_dry's part for theta (module_em.F:1068-1078): + t_tendf/msfty + (c1h*mut+c2h)*h_diabatic/msfty
#ifndef ADV_MINB
#define ADV_MINB 6
#endif
template <bool THETA>
__global__ void __launch_bounds__(128, ADV_MINB) k_advect_scalar_fast(const Geo g, Box b, cfp q, cfp ru, cfp rv, cfp rom, cfp mut, cfp t_tendf,
                                                           cfp h_diabatic, fp tend) {
  const int lane = threadIdx.x;
  const int i_raw = b.i0 + blockIdx.x * 31 + lane;
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j > b.j1) return;
  const int i = min(i_raw, b.i1 + 1);
  const bool st = lane < 31 && i_raw <= b.i1;
  const int ktf = g.kde - 1;
  const float sg = g.sgn_ts;
  const float msf = A2(g.msftx, i, j);
  const float mrdy = msf * g.rdy, mrdx = msf * g.rdx;
  float msfty = 1.f, mutv = 0.f;
  if (THETA) { msfty = A2(g.msfty, i, j); mutv = A2(mut, i, j); }
  float qm1 = 0.f, q0 = A3(q, i, 1, j), q1 = A3(q, i, 2, j);
  float vf = 0.f;
  for (int k = 1; k <= ktf; k++) {
    const float y0 = A3(q, i, k, j - 3), y1 = A3(q, i, k, j - 2), y2 = A3(q, i, k, j - 1);
    const float y4 = A3(q, i, k, j + 1), y5 = A3(q, i, k, j + 2), y6 = A3(q, i, k, j + 3);
    const float x0 = A3(q, i - 3, k, j), x1 = A3(q, i - 2, k, j), x2 = A3(q, i - 1, k, j), x4 = A3(q, i + 1, k, j), x5 = A3(q, i + 2, k, j);
    const float ruv = A3(ru, i, k, j), rv0 = A3(rv, i, k, j), rv1 = A3(rv, i, k, j + 1), romv = A3(rom, i, k + 1, j);
    float q2 = 0.f, tf = 0.f, hd = 0.f;
    if (k + 2 <= ktf) q2 = A3(q, i, k + 2, j);
    if (THETA) { tf = A3(t_tendf, i, k, j); hd = A3(h_diabatic, i, k, j); }
    const float fy0 = rv0 * flux5(y0, y1, y2, q0, y4, y5, rv0, sg);
    const float fy1 = rv1 * flux5(y1, y2, q0, y4, y5, y6, rv1, sg);
    const float fx0 = ruv * flux5(x0, x1, x2, q0, x4, x5, ruv, sg);
    const float fx1 = __shfl_down_sync(0xffffffffu, fx0, 1);
    const float vf1 = adv_vface(g, k + 1, ktf, romv, qm1, q0, q1, q2, sg);
    float t = 0.f;
    t = t - mrdy * (fy1 - fy0);
    t = t - mrdx * (fx1 - fx0);
    t = t - g.rdnw[k] * (vf1 - vf);
    if (THETA) t = t + tf / msfty + (g.c1h[k] * mutv + g.c2h[k]) * hd / msfty;
    if (st) A3(tend, i, k, j) = t;
    vf = vf1;
    qm1 = q0; q0 = q1; q1 = q2;
  }
  if (THETA && st) A3(tend, i, g.kde, j) = 0.f;
}

// the interior box of a tile: every non-periodic side loses 4 cells (see interior_pt in tend.cu)
Box interior_box(const Geo &q) {
  Box in{q.its, std::min(q.ite, q.ide - 1), 1, q.kde, q.jts, std::min(q.jte, q.jde - 1)};
  if (!q.periodic_x) { in.i0 = std::max(in.i0, q.ids + 4); in.i1 = std::min(in.i1, q.ide - 5); }
  if (!q.periodic_y) { in.j0 = std::max(in.j0, q.jds + 4); in.j1 = std::min(in.j1, q.jde - 5); }
  return in;
}

// the (up to four) strips of `tile` around `in`
int edge_strips(const Box &tile, const Box &in, Box out[4]) {
  if (empty(in)) { out[0] = tile; return 1; }
  int n = 0;
  if (in.j0 > tile.j0) out[n++] = Box{tile.i0, tile.i1, tile.k0, tile.k1, tile.j0, in.j0 - 1};
  if (in.j1 < tile.j1) out[n++] = Box{tile.i0, tile.i1, tile.k0, tile.k1, in.j1 + 1, tile.j1};
  if (in.i0 > tile.i0) out[n++] = Box{tile.i0, in.i0 - 1, tile.k0, tile.k1, in.j0, in.j1};
  if (in.i1 < tile.i1) out[n++] = Box{in.i1 + 1, tile.i1, tile.k0, tile.k1, in.j0, in.j1};
  return n;
}

void launch_advect_scalar_fast(Solver *s, bool theta, const Box &in, cfp q, cfp ru, cfp rv, cfp rom, cfp mut, cfp t_tendf,
                               cfp h_diabatic, fp tend) {
  if (empty(in)) return;
  const int ni = in.i1 - in.i0 + 1, nj = in.j1 - in.j0 + 1;
  if (s->coef_dirty) s->sync_coef();
  dim3 block(32, 4), grid((unsigned)((ni + 30) / 31), (unsigned)((nj + 3) / 4));
  s->kpre(theta ? "k_advect_scalar_fast<theta>" : "k_advect_scalar_fast");
  if (theta) k_advect_scalar_fast<true><<<grid, block, 0, s->stream>>>(s->g, in, q, ru, rv, rom, mut, t_tendf, h_diabatic, tend);
  else k_advect_scalar_fast<false><<<grid, block, 0, s->stream>>>(s->g, in, q, ru, rv, rom, mut, t_tendf, h_diabatic, tend);
  s->kpost();
  s->launches++;
}

}  // namespace wg
