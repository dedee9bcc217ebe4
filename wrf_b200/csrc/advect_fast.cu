// advect_fast.cu -- interior fast path of advect_scalar (dyn_em/module_advect_em.F:3029-4359, h5/v3).
//
// The gather kernels of tend.cu evaluate six face fluxes per point, each with its own loads: ncu
// (profiles/r1b) shows them issue-bound (integer pipe 48 %, 10 % DRAM).  Here every face flux is
// evaluated ONCE and shared:
//   * a thread owns one column i and TWO rows (j, j+1) and marches k: the vertical face flux
//     of level k+1 is carried to the next level and the 4-point vertical stencil rolls in registers;
//   * the three y-faces of the row pair share one 8-row column of q (1.5 y-fluxes per point);
//   * the west x-face is computed by the thread, the east one comes from lane+1 by warp shuffle
//     (warps overlap by one column: 31 outputs per 32 lanes);
//   * the next level's 25 operands are loaded before the current level is computed.
// 3.5 flux evaluations and ~13 loads per point instead of 6 and ~40.  Only points at least 4 cells
// inside every non-periodic edge are done here (all faces 5th order, no boundary term); the edge
// strips go through the gather kernels, so the result is the same bits everywhere.
#include "common.cuh"
#include "flux.cuh"

namespace wg {

struct AdvLevel {
  float y0, y1, y2, y5, y6, y7;      // q(i, k, ja-3..ja-1) and q(i, k, ja+2..ja+4); the two centres are carried
  float xa0, xa1, xa2, xa4, xa5;     // q(i-3..i-1, k, ja), q(i+1..i+2, k, ja)
  float xb0, xb1, xb2, xb4, xb5;     // same for row ja+1
  float rua, rub, rv0, rv1, rv2;     // ru(i,k,ja), ru(i,k,ja+1), rv(i,k,ja..ja+2)
  float roma, romb;                  // rom(i,k+1,ja), rom(i,k+1,ja+1)
  float qa2, qb2;                    // q(i,k+2,ja), q(i,k+2,ja+1)
  float tfa, tfb, hda, hdb;          // THETA only: t_tendf, h_diabatic of the two points
};

template <bool THETA>
__device__ __forceinline__ AdvLevel adv_load(const Geo &g, cfp q, cfp ru, cfp rv, cfp rom, cfp t_tendf, cfp h_diabatic, int i, int k, int ja,
                                             int ktf) {
  AdvLevel L;
  if (THETA) {
    L.tfa = A3(t_tendf, i, k, ja); L.tfb = A3(t_tendf, i, k, ja + 1);
    L.hda = A3(h_diabatic, i, k, ja); L.hdb = A3(h_diabatic, i, k, ja + 1);
  }
  L.y0 = A3(q, i, k, ja - 3); L.y1 = A3(q, i, k, ja - 2); L.y2 = A3(q, i, k, ja - 1);
  L.y5 = A3(q, i, k, ja + 2); L.y6 = A3(q, i, k, ja + 3); L.y7 = A3(q, i, k, ja + 4);
  L.xa0 = A3(q, i - 3, k, ja); L.xa1 = A3(q, i - 2, k, ja); L.xa2 = A3(q, i - 1, k, ja);
  L.xa4 = A3(q, i + 1, k, ja); L.xa5 = A3(q, i + 2, k, ja);
  L.xb0 = A3(q, i - 3, k, ja + 1); L.xb1 = A3(q, i - 2, k, ja + 1); L.xb2 = A3(q, i - 1, k, ja + 1);
  L.xb4 = A3(q, i + 1, k, ja + 1); L.xb5 = A3(q, i + 2, k, ja + 1);
  L.rua = A3(ru, i, k, ja); L.rub = A3(ru, i, k, ja + 1);
  L.rv0 = A3(rv, i, k, ja); L.rv1 = A3(rv, i, k, ja + 1); L.rv2 = A3(rv, i, k, ja + 2);
  L.roma = A3(rom, i, k + 1, ja); L.romb = A3(rom, i, k + 1, ja + 1);
  L.qa2 = 0.f; L.qb2 = 0.f;
  if (k + 2 <= ktf) { L.qa2 = A3(q, i, k + 2, ja); L.qb2 = A3(q, i, k + 2, ja + 1); }
  return L;
}

// vertical face kk = k+1 (:1470-1498 pattern): 0 at the lid, 2nd order next to the ends, 3rd order inside
__device__ __forceinline__ float adv_vface(const Geo &g, int kk, int ktf, float vel, float qm2, float qm1, float q0, float qp1, float sg) {
  if (kk == g.kde) return 0.f;
  if (kk == 2 || kk == ktf) return vel * (g.fnm[kk] * q0 + g.fnp[kk] * qm1);
  return vel * flux3(qm2, qm1, q0, qp1, -vel, sg);
}

// THETA: also rk_addtend_dry's part for theta (module_em.F:1068-1078): + t_tendf/msfty + (c1h*mut+c2h)*h_diabatic/msfty
template <bool THETA>
__global__ void __launch_bounds__(128) k_advect_scalar_fast(const Geo g, Box b, cfp q, cfp ru, cfp rv, cfp rom, cfp mut, cfp t_tendf,
                                                            cfp h_diabatic, fp tend) {
  const int lane = threadIdx.x;
  const int i_raw = b.i0 + blockIdx.x * 31 + lane;
  const int ja = b.j0 + (blockIdx.y * blockDim.y + threadIdx.y) * 2;
  if (ja > b.j1) return;                          // the whole warp leaves together (one row pair per warp)
  const int i = min(i_raw, b.i1 + 1);             // lanes past the box only feed the shuffle / stay in bounds
  const bool st_a = lane < 31 && i_raw <= b.i1, st_b = st_a && (ja + 1 <= b.j1);
  const int ktf = g.kde - 1;
  const float sg = g.sgn_ts;
  const float msfa = A2(g.msftx, i, ja), msfb = A2(g.msftx, i, ja + 1);
  const float mrdya = msfa * g.rdy, mrdxa = msfa * g.rdx, mrdyb = msfb * g.rdy, mrdxb = msfb * g.rdx;
  float msftya = 1.f, msftyb = 1.f, muta = 0.f, mutb = 0.f;
  if (THETA) { msftya = A2(g.msfty, i, ja); msftyb = A2(g.msfty, i, ja + 1); muta = A2(mut, i, ja); mutb = A2(mut, i, ja + 1); }
  // vertical windows: q at k-1, k, k+1 (k+2 arrives with the level's loads)
  float am1 = 0.f, a0 = A3(q, i, 1, ja), a1 = A3(q, i, 2, ja);
  float bm1 = 0.f, b0 = A3(q, i, 1, ja + 1), b1 = A3(q, i, 2, ja + 1);
  float vfa = 0.f, vfb = 0.f;                     // vertical face k (face 1 carries no flux)
  AdvLevel L = adv_load<THETA>(g, q, ru, rv, rom, t_tendf, h_diabatic, i, 1, ja, ktf);
  for (int k = 1; k <= ktf; k++) {
    AdvLevel N;
    if (k < ktf) N = adv_load<THETA>(g, q, ru, rv, rom, t_tendf, h_diabatic, i, k + 1, ja, ktf);
    // y faces ja, ja+1, ja+2 (:3413-3560 with every face 5th order)
    const float fy0 = L.rv0 * flux5(L.y0, L.y1, L.y2, a0, b0, L.y5, L.rv0, sg);
    const float fy1 = L.rv1 * flux5(L.y1, L.y2, a0, b0, L.y5, L.y6, L.rv1, sg);
    const float fy2 = L.rv2 * flux5(L.y2, a0, b0, L.y5, L.y6, L.y7, L.rv2, sg);
    // x faces: west here, east from the next lane
    const float fxa = L.rua * flux5(L.xa0, L.xa1, L.xa2, a0, L.xa4, L.xa5, L.rua, sg);
    const float fxb = L.rub * flux5(L.xb0, L.xb1, L.xb2, b0, L.xb4, L.xb5, L.rub, sg);
    const float fxa_e = __shfl_down_sync(0xffffffffu, fxa, 1);
    const float fxb_e = __shfl_down_sync(0xffffffffu, fxb, 1);
    // vertical face k+1
    const float vfa1 = adv_vface(g, k + 1, ktf, L.roma, am1, a0, a1, L.qa2, sg);
    const float vfb1 = adv_vface(g, k + 1, ktf, L.romb, bm1, b0, b1, L.qb2, sg);
    float ta = 0.f, tb = 0.f;
    ta = ta - mrdya * (fy1 - fy0);
    ta = ta - mrdxa * (fxa_e - fxa);
    ta = ta - g.rdnw[k] * (vfa1 - vfa);
    tb = tb - mrdyb * (fy2 - fy1);
    tb = tb - mrdxb * (fxb_e - fxb);
    tb = tb - g.rdnw[k] * (vfb1 - vfb);
    if (THETA) {
      ta = ta + L.tfa / msftya + (g.c1h[k] * muta + g.c2h[k]) * L.hda / msftya;
      tb = tb + L.tfb / msftyb + (g.c1h[k] * mutb + g.c2h[k]) * L.hdb / msftyb;
    }
    if (st_a) A3(tend, i, k, ja) = ta;
    if (st_b) A3(tend, i, k, ja + 1) = tb;
    vfa = vfa1; vfb = vfb1;
    am1 = a0; a0 = a1; a1 = L.qa2;
    bm1 = b0; b0 = b1; b1 = L.qb2;
    L = N;
  }
  if (THETA) {   // the theta tendency is defined on all levels: level kde is zero (module_em.F:370-428 zero_tend)
    if (st_a) A3(tend, i, g.kde, ja) = 0.f;
    if (st_b) A3(tend, i, g.kde, ja + 1) = 0.f;
  }
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
  dim3 block(32, 4), grid((unsigned)((ni + 30) / 31), (unsigned)((nj + 7) / 8));
  if (s->coef_dirty) s->sync_coef();
  s->kpre(theta ? "k_advect_scalar_fast<theta>" : "k_advect_scalar_fast");
  if (theta) k_advect_scalar_fast<true><<<grid, block, 0, s->stream>>>(s->g, in, q, ru, rv, rom, mut, t_tendf, h_diabatic, tend);
  else k_advect_scalar_fast<false><<<grid, block, 0, s->stream>>>(s->g, in, q, ru, rv, rom, mut, t_tendf, h_diabatic, tend);
  s->kpost();
  s->launches++;
}

}  // namespace wg
