// tend_tile.cu -- the whole rk_tendency phase of the patch interior as ONE kernel: ru_tend, rv_tend, rw_tend, ph_tend
// and t_tend from one pass over their operands, the horizontal stencil neighbourhoods staged in shared memory by TMA.
//
// Reference routines fused (the sums, operand for operand, of the k-marching kernels of mom_fast.cu / advect_fast.cu,
// which remain as the reference form of this kernel and as the path for configurations it does not cover):
//   advect_u / advect_v / advect_w / advect_scalar(theta)  dyn_em/module_advect_em.F:126, :1530, :4364, :3029 (h5 / v3)
//   rhs_ph                                                  module_big_step_utilities_em.F:1365
//   horizontal_pressure_gradient, pg_buoy_w (+ cqw <- 1/(1+cqw)), w_damp, coriolis, curvature   :2183, :2419, :2503, :3640, :4175
//   rk_addtend_dry                                          module_em.F:959
//
// Why: as five kernels the phase moved 233 B per point of DRAM traffic against 120 algorithmic (each kernel re-read ru, rv,
// ww, u, v, ph, p, ...), and every kernel was bound by instruction issue and L1 wavefronts: ~40 LDG per point per kernel,
// each with 64-bit address arithmetic and, rows not being 128-byte aligned, two L1 wavefronts (profiles/r1t_summary.md).
// Here
//   * a CTA owns a tile of 31 x 8 columns and marches k; per level ONE elected thread issues 15 TMA loads
//     (cp.async.bulk.tensor.3d, SASS UTMALDG): the 6 fields with +-3 stencils (u, v, w, theta, ph, phb) as 44 x 14 boxes,
//     the 9 fields read one point to the west / south / east / north (ru, rv, ww, rw, p, pb, al, alt, php) as 40 x 10 boxes,
//     three levels deep (k and k+1 are read, k+2 is in flight), completion signalled through mbarriers;
//   * every stencil operand is then one LDS with an immediate offset (no address arithmetic, one wavefront per warp);
//   * operands shared by the five tendencies are fetched once: 24 fields in, 5 tendencies + cqw out = 120 B per point;
//   * the vertical stencils roll in registers as before, the west x-face flux is computed by the thread and the east one
//     comes from lane + 1 by warp shuffle (31 outputs per 32 lanes).
// Only points at least 4 cells inside every non-periodic edge (interior_box) are done here; the strips go through the
// gather kernels of tend.cu, so the result is the same bits everywhere.
#include "common.cuh"
#include "flux.cuh"
#include <cuda.h>
#include <cstdint>

namespace wg {

#define TT_TX 32          // lanes along i; 31 outputs (lane 31 only feeds the shuffle)
#define TT_OX 31
#define TT_TY 8
// TMA wants the start coordinate of the contiguous dimension on a 16-byte boundary (measured: a start of 2 floats raises
// "illegal instruction", 4 works; tools/tma_probe2.cu), so a box starts at the tile's first column rounded DOWN to a
// multiple of 4 floats and is 3 columns wider than the stencil needs; the CTA adds its offset (0..3) when it reads.
#define TT_SW 44          // stencil tile: 38 columns i0-3 .. i0+34 (+3 for the alignment, 16-byte multiple), rows j0-3 .. j0+10
#define TT_SH 14
#define TT_HW 40          // one-point tile: 34 columns i0-1 .. i0+32 (+3), rows j0-1 .. j0+8
#define TT_HH 10
#define TT_NST 6
#define TT_NH1 9
#define TT_STF 640        // floats per stencil tile slot (44*14 = 616, padded to a 128-byte multiple)
#define TT_H1F 416        // floats per one-point tile slot (40*10 = 400, padded)
#define TT_STAGE_F (TT_NST * TT_STF + TT_NH1 * TT_H1F)      // 7584 floats = 30336 B
#define TT_STAGES 3
#define TT_TX_BYTES ((TT_NST * TT_SW * TT_SH + TT_NH1 * TT_HW * TT_HH) * 4)   // bytes one level's loads deliver
#define TT_SMEM (TT_STAGES * TT_STAGE_F * 4 + 64)

enum { F_U = 0, F_V, F_W, F_T, F_PH, F_PHB };
enum { H_RU = 0, H_RV, H_WW, H_RW, H_P, H_PB, H_AL, H_ALT, H_PHP };

struct TileMaps { CUtensorMap st[TT_NST]; CUtensorMap h1[TT_NH1]; };

struct TileIn {
  cfp u, v, w, t, ww;                        // centre-column reads two levels ahead
  cfp p, ru, rv;                             // column start-up / top-level reads
  cfp cqu, cqv, cqw, h_diabatic;
  cfp mut, muu, muv, mu, mub;
  cfp ru_tendf, rv_tendf, rw_tendf, ph_tendf, t_tendf;
  fp ru_tend, rv_tend, rw_tend, ph_tend, t_tend, cqw_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a lost completion becomes a trap (a CUDA error the host reports), never a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// x / msf with the map factor of the column: x / 1.0f is x for every x (IEEE), and the idealised cases have msf == 1
// everywhere.  Worth a test per column because the tendf / h_diabatic numerators are exactly zero until a physics scheme
// fills them, and a zero numerator sends nvcc's IEEE division down its ~70-instruction slow path (FCHK): five of those per
// point were 18 % of this kernel's instructions (ncu r2j).
__device__ __forceinline__ float div_msf(float x, float msf, bool msf_is_one) {
  float r = x;
  if (!msf_is_one) {
    asm volatile("" ::: "memory");   // keeps the division inside the branch (nvcc otherwise evaluates it speculatively and selects)
    r = x / msf;
  }
  return r;
}

#ifndef TT_MINB
#define TT_MINB 2
#endif

__global__ void __launch_bounds__(TT_TX *TT_TY, TT_MINB)
k_tend_tile(const Geo g, Box b, const __grid_constant__ TileMaps maps, TileIn s) {
  extern __shared__ __align__(128) float tt_smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(tt_smem + TT_STAGES * TT_STAGE_F);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TT_TX + tx;
  const int i0 = b.i0 + blockIdx.x * TT_OX, j0 = b.j0 + blockIdx.y * TT_TY;
  const int i_raw = i0 + tx, j_raw = j0 + ty;
  const int i = min(i_raw, b.i1 + 1), j = min(j_raw, b.j1);   // clamped: global reads of lanes outside the box stay in memory
  const bool st = tx < TT_OX && i_raw <= b.i1 && j_raw <= b.j1;
  const int kde = g.kde, ktf = g.kde - 1;
  const float sg = g.sgn_ts;
  // offsets of this thread's column inside a stage: the box's alignment shift (0..3) + its place in the tile
  const int st_off = ((i0 - 3 - g.ims) & 3) + ty * TT_SW + tx;
  const int h1_off = TT_NST * TT_STF + ((i0 - 1 - g.ims) & 3) + ty * TT_HW + tx;

  if (tid == 0) {
    for (int q = 0; q < TT_STAGES; q++) mbar_init(smem_u32(&bars[q]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // level L lives in slot L % 3; its n-th use of the slot (n = (L-1)/3) completes phase n
  auto issue_level = [&](int L) {
    const int slot = L % TT_STAGES;
    const uint32_t bar = smem_u32(&bars[slot]);
    float *base = tt_smem + slot * TT_STAGE_F;
    mbar_expect_tx(bar, TT_TX_BYTES);
    const int cs0 = (i0 - 3 - g.ims) & ~3, cs2 = j0 - 3 - g.jms, ch0 = (i0 - 1 - g.ims) & ~3, ch2 = j0 - 1 - g.jms, c1 = L - g.kms;
#pragma unroll
    for (int f = 0; f < TT_NST; f++) tma_load_3d(smem_u32(base + f * TT_STF), &maps.st[f], cs0, c1, cs2, bar);
#pragma unroll
    for (int f = 0; f < TT_NH1; f++) tma_load_3d(smem_u32(base + TT_NST * TT_STF + f * TT_H1F), &maps.h1[f], ch0, c1, ch2, bar);
  };
  if (tid == 0) { issue_level(1); issue_level(2); }

  // ---- column constants (2-D fields)
  const float msfux = A2(g.msfux, i, j), msfuy = A2(g.msfuy, i, j);
  const float msfvy = A2(g.msfvy, i, j), msfvx = A2(g.msfvx, i, j), msfvx_inv = A2(g.msfvx_inv, i, j);
  const float msftx = A2(g.msftx, i, j), msfty = A2(g.msfty, i, j);
  const float muu = A2(s.muu, i, j), muv = A2(s.muv, i, j), mut = A2(s.mut, i, j), mub = A2(s.mub, i, j);
  const float mu_c = A2(s.mu, i, j), mu_w = A2(s.mu, i - 1, j), mu_s = A2(s.mu, i, j - 1);
  const float f_c = A2(g.f, i, j), e_c = A2(g.e, i, j), cosa_c = A2(g.cosa, i, j), sina_c = A2(g.sina, i, j);
  // u (k_mom_u_fast)
  const float u_mrdy = msfux * g.rdy, u_mrdx = msfux * g.rdx;
  const float u_mr = msfux / msfuy;
  const float u_pA = u_mr * .5f * g.rdx, u_pB = u_mr * g.rdx;
  const float u_cA = u_mr * 0.5f * (f_c + A2(g.f, i - 1, j)) * 0.25f;
  const float u_cB = 0.5f * (e_c + A2(g.e, i - 1, j)) * 0.5f * (cosa_c + A2(g.cosa, i - 1, j)) * 0.25f;
  const float dvx_c = A2(g.msfvx, i, j + 1) - msfvx, duy_c = A2(g.msfuy, i + 1, j) - msfuy;
  const float dvx_w = A2(g.msfvx, i - 1, j + 1) - A2(g.msfvx, i - 1, j), duy_w = msfuy - A2(g.msfuy, i - 1, j);
  // v (k_mom_v_fast)
  const float v_mrdy = msfvy * g.rdy, v_mrdx = msfvy * g.rdx;
  const float v_mr = msfvy / msfvx;
  const float v_pA = v_mr * .5f * g.rdy, v_pB = v_mr * g.rdy;
  const float v_cA = v_mr * 0.5f * (f_c + A2(g.f, i, j - 1)) * 0.25f;
  const float v_cB = v_mr * 0.5f * (e_c + A2(g.e, i, j - 1)) * 0.5f * (sina_c + A2(g.sina, i, j - 1)) * 0.25f;
  const float dvx_s = msfvx - A2(g.msfvx, i, j - 1), duy_s = A2(g.msfuy, i + 1, j - 1) - A2(g.msfuy, i, j - 1);
  // w, theta (k_mom_w_fast, k_advect_scalar_fast<true>)
  const float t_mrdy = msftx * g.rdy, t_mrdx = msftx * g.rdx;
  const float w_mr = msftx / msfty;
  const float rmsf = (1.f / msfty);
  const bool one_uy = (msfuy == 1.0f), one_ty = (msfty == 1.0f);
  // ph (k_ph_fast)
  const float muv_n = A2(s.muv, i, j + 1), muu_e = A2(s.muu, i + 1, j);
  const float ya = A2(g.msfvy, i, j + 1), xa = A2(g.msfux, i + 1, j);

  // level 1 of the full-level tendencies: only rk_addtend_dry
  if (st) {
    A3(s.rw_tend, i, 1, j) = 0.f + div_msf(A3(s.rw_tendf, i, 1, j), msfty, one_ty);
    A3(s.ph_tend, i, 1, j) = 0.f + div_msf(A3(s.ph_tendf, i, 1, j), msfty, one_ty);
  }

  // ---- rolling vertical state
  float um1 = 0.f, u0 = A3(s.u, i, 1, j), u1 = A3(s.u, i, 2, j);      // u(k-1), u(k), u(k+1)
  float vm1 = 0.f, v0 = A3(s.v, i, 1, j), v1 = A3(s.v, i, 2, j);
  float qm1 = 0.f, q0 = A3(s.t, i, 1, j), q1 = A3(s.t, i, 2, j);
  float w0 = 0.f, w1 = A3(s.w, i, 1, j), w2 = A3(s.w, i, 2, j);       // w(k-1), w(k), w(k+1)
  float uvf = 0.f, vvf = 0.f, tvf = 0.f, wvf = 0.f;                   // vertical face fluxes at the lower face of level k
  float ups_km1 = 0.f, vps_km1 = 0.f;                                 // p(i-1)+p(i), p(j-1)+p(j) at k-1
  float ud0, vd0;                                                     // dpn(k) of the x and y pressure gradients
  {
    const float p1c = A3(s.p, i, 1, j), p2c = A3(s.p, i, 2, j), p3c = A3(s.p, i, 3, j);
    ud0 = .5f * (g.cf1 * (A3(s.p, i - 1, 1, j) + p1c) + g.cf2 * (A3(s.p, i - 1, 2, j) + p2c) + g.cf3 * (A3(s.p, i - 1, 3, j) + p3c));
    vd0 = .5f * (g.cf1 * (A3(s.p, i, 1, j - 1) + p1c) + g.cf2 * (A3(s.p, i, 2, j - 1) + p2c) + g.cf3 * (A3(s.p, i, 3, j - 1) + p3c));
  }
  // level k-1 values the full-level tendencies read (k_mom_w_fast, k_ph_fast)
  float ru_c_m = 0.f, ru_e_m = 0.f, rv_c_m = 0.f, rv_n_m = 0.f;
  float u_c_m = 0.f, u_e_m = 0.f, v_c_m = 0.f, v_n_m = 0.f;
  float p_m = 0.f, cqw_m = 0.f;
  float wd_k = 0.f;                                                    // rhs_ph's wdwn(k)

  for (int k = 1; k <= kde; k++) {
    __syncthreads();                       // every thread is done with level k-1: its slot takes level k+2
    if (tid == 0 && k + 2 <= kde) issue_level(k + 2);
    const bool half = (k <= ktf), full = (k >= 2), top = (k == kde);
    // global centre-column reads of this level, issued before the wait
    float u2 = 0.f, v2 = 0.f, q2 = 0.f, w3 = 0.f;
    if (k + 2 <= ktf) { u2 = A3(s.u, i, k + 2, j); v2 = A3(s.v, i, k + 2, j); q2 = A3(s.t, i, k + 2, j); }
    if (k + 2 <= kde) w3 = A3(s.w, i, k + 2, j);
    float cqu = 0.f, cqv = 0.f, ru_tf = 0.f, rv_tf = 0.f, t_tf = 0.f, hd = 0.f, cqw = 0.f, rw_tf = 0.f, ph_tf = 0.f;
    if (half) {
      cqu = A3(s.cqu, i, k, j); cqv = A3(s.cqv, i, k, j); ru_tf = A3(s.ru_tendf, i, k, j); rv_tf = A3(s.rv_tendf, i, k, j);
      t_tf = A3(s.t_tendf, i, k, j); hd = A3(s.h_diabatic, i, k, j);
    }
    if (full) { rw_tf = A3(s.rw_tendf, i, k, j); ph_tf = A3(s.ph_tendf, i, k, j); }
    if (full && !top) cqw = A3(s.cqw, i, k, j);
    if (k == 1) mbar_wait(smem_u32(&bars[1]), 0);
    if (!top) mbar_wait(smem_u32(&bars[(k + 1) % TT_STAGES]), (uint32_t)((k / 3) & 1));
    // this thread's view of level k (0) and level k+1 (1, not read at the top): every operand is base + constant
    const float *S0 = tt_smem + (k % TT_STAGES) * TT_STAGE_F + st_off, *S1 = tt_smem + ((k + 1) % TT_STAGES) * TT_STAGE_F + st_off;
    const float *Q0 = tt_smem + (k % TT_STAGES) * TT_STAGE_F + h1_off, *Q1 = tt_smem + ((k + 1) % TT_STAGES) * TT_STAGE_F + h1_off;
#define ST0(f, dx, dy) S0[(f) * TT_STF + (3 + (dy)) * TT_SW + (3 + (dx))]
#define ST1(f, dx, dy) S1[(f) * TT_STF + (3 + (dy)) * TT_SW + (3 + (dx))]
#define H0(f, dx, dy) Q0[(f) * TT_H1F + (1 + (dy)) * TT_HW + (1 + (dx))]
#define H1(f, dx, dy) Q1[(f) * TT_H1F + (1 + (dy)) * TT_HW + (1 + (dx))]
    // ---- operands shared by several tendencies
    const float ru_c = H0(H_RU, 0, 0), ru_w = H0(H_RU, -1, 0), ru_e = H0(H_RU, 1, 0);
    const float rv_c = H0(H_RV, 0, 0), rv_s = H0(H_RV, 0, -1), rv_n = H0(H_RV, 0, 1);
    const float uc = ST0(F_U, 0, 0), ue = ST0(F_U, 1, 0), vc = ST0(F_V, 0, 0), vn = ST0(F_V, 0, 1);
    const float p_c = H0(H_P, 0, 0);
    float rom1_c = 0.f;                                                   // ww(k+1)
    if (!top) rom1_c = H1(H_WW, 0, 0);

    if (half) {
      const float c1h = g.c1h[k], c2h = g.c2h[k];
      const float rdnw = g.rdnw[k];
      const float rom_w = H1(H_WW, -1, 0), rom_s = H1(H_WW, 0, -1);
      const float ph_c = ST0(F_PH, 0, 0), ph1_c = ST1(F_PH, 0, 0);
      float p1_c = 0.f;
      if (k + 1 <= ktf) p1_c = H1(H_P, 0, 0);
      const float alt_c = H0(H_ALT, 0, 0), al_c = H0(H_AL, 0, 0), pb_c = H0(H_PB, 0, 0), php_c = H0(H_PHP, 0, 0);
      const float rw_c = H0(H_RW, 0, 0), rw1_c = H1(H_RW, 0, 0);
      // =====================================================================  ru_tend (k_mom_u_fast)
      {
        const float uy0 = ST0(F_U, 0, -3), uy1 = ST0(F_U, 0, -2), uy2 = ST0(F_U, 0, -1);
        const float uy4 = ST0(F_U, 0, 1), uy5 = ST0(F_U, 0, 2), uy6 = ST0(F_U, 0, 3);
        const float ux0 = ST0(F_U, -3, 0), ux1 = ST0(F_U, -2, 0), ux2 = ST0(F_U, -1, 0);
        const float ux4 = ue, ux5 = ST0(F_U, 2, 0);
        const float rv_c0 = rv_c, rv_w0 = H0(H_RV, -1, 0), rv_c1 = rv_n, rv_w1 = H0(H_RV, -1, 1);
        const float ph_w = ST0(F_PH, -1, 0), ph1_w = ST1(F_PH, -1, 0);
        const float p_w = H0(H_P, -1, 0);
        float p1_w = 0.f;
        if (k + 1 <= ktf) p1_w = H1(H_P, -1, 0);
        const float alt_w = H0(H_ALT, -1, 0), al_w = H0(H_AL, -1, 0), pb_w = H0(H_PB, -1, 0), php_w = H0(H_PHP, -1, 0);
        const float rw_w = H0(H_RW, -1, 0), rw1_w = H1(H_RW, -1, 0);
        const float v_c0 = vc, v_c1 = vn, v_w0 = ST0(F_V, -1, 0), v_w1 = ST0(F_V, -1, 1);
        float t = 0.f;
        {   // y flux divergence (:553-647), faces j and j+1
          const float vel0 = 0.5f * (rv_c0 + rv_w0), vel1 = 0.5f * (rv_c1 + rv_w1);
          const float fq0 = vel0 * flux5(uy0, uy1, uy2, u0, uy4, uy5, vel0, sg);
          const float fq1 = vel1 * flux5(uy1, uy2, u0, uy4, uy5, uy6, vel1, sg);
          t = t - u_mrdy * (fq1 - fq0);
        }
        {   // x flux divergence (:649-745): west face here, east face from the next lane
          const float vel = 0.5f * (ru_c + ru_w);
          const float fq0 = vel * flux5(ux0, ux1, ux2, u0, ux4, ux5, vel, sg);
          const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
          t = t - u_mrdx * (fq1 - fq0);
        }
        float vf1;   // vertical flux divergence (:1336-1498), face k+1
        {
          const int kk = k + 1;
          if (kk == kde) vf1 = 0.f;
          else if (kk == 2 || kk == ktf) vf1 = 0.5f * (rom1_c + rom_w) * (g.fnm[kk] * u1 + g.fnp[kk] * u0);
          else {
            const float vel = 0.5f * (rom_w + rom1_c);
            vf1 = vel * flux3(um1, u0, u1, u2, -vel, sg);
          }
          t = t - rdnw * (vf1 - uvf);
        }
        // horizontal pressure gradient, x (:2320-2400)
        const float ps_k = p_w + p_c, ps_k1 = p1_w + p1_c;
        float d1;
        if (k + 1 == kde) d1 = g.top_lid ? .5f * (g.cfn * ps_k + g.cfn1 * ups_km1) : 0.f;
        else d1 = .5f * (g.fnm[k + 1] * ps_k1 + g.fnp[k + 1] * ps_k);
        {
          float dpx = u_pA * (c1h * muu + c2h) * ((ph1_c - ph1_w + ph_c - ph_w) + (alt_c + alt_w) * (p_c - p_w) + (al_c + al_w) * (pb_c - pb_w));
          dpx = dpx + u_pB * (php_c - php_w) * (rdnw * (d1 - ud0) - .5f * ((c1h * mu_w) + (c1h * mu_c)));
          t = t - cqu * dpx;
        }
        {   // coriolis (:3700-3735) and curvature (:4292-4330)
          const float rvavg = 0.25f * (rv_w1 + rv_c1 + rv_w0 + rv_c0);
          const float rwsum = (rw1_w + rw_w + rw1_c + rw_c);
          t = t + u_cA * (rv_w1 + rv_c1 + rv_w0 + rv_c0) - u_cB * rwsum;
          const float vx_c = 0.5f * (u0 + ux4) * dvx_c * g.rdy - 0.5f * (v_c0 + v_c1) * duy_c * g.rdx;
          const float vx_w = 0.5f * (ux2 + u0) * dvx_w * g.rdy - 0.5f * (v_w0 + v_w1) * duy_w * g.rdx;
          t = t + 0.5f * (vx_c + vx_w) * rvavg - u0 * WG_RERADIUS * 0.25f * rwsum;
        }
        t = t + div_msf(ru_tf, msfuy, one_uy);   // rk_addtend_dry (module_em.F:1037-1043)
        if (st) A3(s.ru_tend, i, k, j) = t;
        uvf = vf1; ud0 = d1; ups_km1 = ps_k;
      }
      // =====================================================================  rv_tend (k_mom_v_fast)
      {
        const float vy0 = ST0(F_V, 0, -3), vy1 = ST0(F_V, 0, -2), vy2 = ST0(F_V, 0, -1);
        const float vy4 = vn, vy5 = ST0(F_V, 0, 2), vy6 = ST0(F_V, 0, 3);
        const float vx0 = ST0(F_V, -3, 0), vx1 = ST0(F_V, -2, 0), vx2 = ST0(F_V, -1, 0);
        const float vx4 = ST0(F_V, 1, 0), vx5 = ST0(F_V, 2, 0);
        const float rv_m = rv_s, rv_p = rv_n;
        const float ru_c0 = ru_c, ru_s0 = H0(H_RU, 0, -1), ru_c1 = ru_e, ru_s1 = H0(H_RU, 1, -1);
        const float ph_s = ST0(F_PH, 0, -1), ph1_s = ST1(F_PH, 0, -1);
        const float p_s = H0(H_P, 0, -1);
        float p1_s = 0.f;
        if (k + 1 <= ktf) p1_s = H1(H_P, 0, -1);
        const float alt_s = H0(H_ALT, 0, -1), al_s = H0(H_AL, 0, -1), pb_s = H0(H_PB, 0, -1), php_s = H0(H_PHP, 0, -1);
        const float rw_s = H0(H_RW, 0, -1), rw1_s = H1(H_RW, 0, -1);
        const float u_c0 = uc, u_c1 = ue, u_s0 = ST0(F_U, 0, -1), u_s1 = ST0(F_U, 1, -1);
        float t = 0.f;
        {   // y flux divergence (:1960-2060), faces j and j+1
          const float vel0 = 0.5f * (rv_c + rv_m), vel1 = 0.5f * (rv_p + rv_c);
          const float fq0 = vel0 * flux5(vy0, vy1, vy2, v0, vy4, vy5, vel0, sg);
          const float fq1 = vel1 * flux5(vy1, vy2, v0, vy4, vy5, vy6, vel1, sg);
          t = t - v_mrdy * (fq1 - fq0);
        }
        {   // x flux divergence (:2062-2160)
          const float vel = 0.5f * (ru_c0 + ru_s0);
          const float fq0 = vel * flux5(vx0, vx1, vx2, v0, vx4, vx5, vel, sg);
          const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
          t = t - v_mrdx * (fq1 - fq0);
        }
        float vf1;   // vertical flux divergence (:2802-2995), face k+1
        {
          const int kk = k + 1;
          if (kk == kde) vf1 = 0.f;
          else if (kk == 2 || kk == ktf) vf1 = 0.5f * (rom1_c + rom_s) * (g.fnm[kk] * v1 + g.fnp[kk] * v0);
          else {
            const float vel = 0.5f * (rom1_c + rom_s);
            vf1 = vel * flux3(vm1, v0, v1, v2, -vel, sg);
          }
          t = t - v_mr * rdnw * (vf1 - vvf);
        }
        // horizontal pressure gradient, y (:2230-2300)
        const float ps_k = p_s + p_c, ps_k1 = p1_s + p1_c;
        float d1;
        if (k + 1 == kde) d1 = g.top_lid ? .5f * (g.cfn * ps_k + g.cfn1 * vps_km1) : 0.f;
        else d1 = .5f * (g.fnm[k + 1] * ps_k1 + g.fnp[k + 1] * ps_k);
        {
          float dpy = v_pA * (c1h * muv + c2h) * ((ph1_c - ph1_s + ph_c - ph_s) + (alt_c + alt_s) * (p_c - p_s) + (al_c + al_s) * (pb_c - pb_s));
          dpy = dpy + v_pB * (php_c - php_s) * (rdnw * (d1 - vd0) - .5f * ((c1h * mu_s) + (c1h * mu_c)));
          t = t - cqv * dpy;
        }
        {   // coriolis (:3745-3765), curvature (:4340-4380)
          const float rwsum = (rw1_s + rw_s + rw1_c + rw_c);
          const float ruavg = 0.25f * (ru_c0 + ru_c1 + ru_s0 + ru_s1);
          t = t - v_cA * (ru_c0 + ru_c1 + ru_s0 + ru_s1) + v_cB * rwsum;
          const float vx_c = 0.5f * (u_c0 + u_c1) * dvx_c * g.rdy - 0.5f * (v0 + vy4) * duy_c * g.rdx;
          const float vx_s = 0.5f * (u_s0 + u_s1) * dvx_s * g.rdy - 0.5f * (vy2 + v0) * duy_s * g.rdx;
          t = t - 0.5f * (vx_c + vx_s) * ruavg - v_mr * v0 * WG_RERADIUS * 0.25f * rwsum;
        }
        t = t + rv_tf * msfvx_inv;
        if (st) A3(s.rv_tend, i, k, j) = t;
        vvf = vf1; vd0 = d1; vps_km1 = ps_k;
      }
      // =====================================================================  t_tend (k_advect_scalar_fast<true>)
      {
        const float y0 = ST0(F_T, 0, -3), y1 = ST0(F_T, 0, -2), y2 = ST0(F_T, 0, -1);
        const float y4 = ST0(F_T, 0, 1), y5 = ST0(F_T, 0, 2), y6 = ST0(F_T, 0, 3);
        const float x0 = ST0(F_T, -3, 0), x1 = ST0(F_T, -2, 0), x2 = ST0(F_T, -1, 0), x4 = ST0(F_T, 1, 0), x5 = ST0(F_T, 2, 0);
        const float ruv = ru_c, rv0 = rv_c, rv1 = rv_n, romv = rom1_c;
        const float fy0 = rv0 * flux5(y0, y1, y2, q0, y4, y5, rv0, sg);
        const float fy1 = rv1 * flux5(y1, y2, q0, y4, y5, y6, rv1, sg);
        const float fx0 = ruv * flux5(x0, x1, x2, q0, x4, x5, ruv, sg);
        const float fx1 = __shfl_down_sync(0xffffffffu, fx0, 1);
        float vf1;   // adv_vface (advect_fast.cu)
        {
          const int kk = k + 1;
          if (kk == kde) vf1 = 0.f;
          else if (kk == 2 || kk == ktf) vf1 = romv * (g.fnm[kk] * q1 + g.fnp[kk] * q0);
          else vf1 = romv * flux3(qm1, q0, q1, q2, -romv, sg);
        }
        float t = 0.f;
        t = t - t_mrdy * (fy1 - fy0);
        t = t - t_mrdx * (fx1 - fx0);
        t = t - rdnw * (vf1 - tvf);
        t = t + div_msf(t_tf, msfty, one_ty) + div_msf((c1h * mut + c2h) * hd, msfty, one_ty);
        if (st) A3(s.t_tend, i, k, j) = t;
        tvf = vf1;
      }
    } else if (st) {   // k == kde: the half-level tendencies are zero there (rk_tendency zeroes the arrays, module_em.F:370-428)
      A3(s.ru_tend, i, kde, j) = 0.f;
      A3(s.rv_tend, i, kde, j) = 0.f;
      A3(s.t_tend, i, kde, j) = 0.f;
    }

    if (k == 1) {
      // vertical face 2 of w (:5996-6031): 2nd order; and rhs_ph's wdwn(1) (:1459-1475)
      const float rom_1 = H0(H_WW, 0, 0);
      wvf = 0.25f * (rom1_c + rom_1) * (w2 + w1);
      const float ph_1 = ST0(F_PH, 0, 0), phb_1 = ST0(F_PHB, 0, 0), ph_2 = ST1(F_PH, 0, 0), phb_2 = ST1(F_PHB, 0, 0);
      if (g.phi_adv_z == 2) wd_k = g.rdnw[1] * (ph_2 - ph_1 + phb_2 - phb_1);
      else wd_k = .5f * (rom1_c + rom_1) * g.rdnw[1] * (ph_2 - ph_1 + phb_2 - phb_1);
    }
    if (full) {
      const float rom_k = H0(H_WW, 0, 0);
      // level k-2 values of the extrapolated top-level velocities (:4847,4856; rhs_ph :1783-1970): read once per column
      float ru_c_mm = 0.f, rv_c_mm = 0.f, rv_n_mm = 0.f, u_c_mm = 0.f, u_e_mm = 0.f, v_c_mm = 0.f, v_n_mm = 0.f;
      if (top) {
        ru_c_mm = A3(s.ru, i, k - 2, j); rv_c_mm = A3(s.rv, i, k - 2, j); rv_n_mm = A3(s.rv, i, k - 2, j + 1);
        u_c_mm = A3(s.u, i, k - 2, j); u_e_mm = A3(s.u, i + 1, k - 2, j); v_c_mm = A3(s.v, i, k - 2, j); v_n_mm = A3(s.v, i, k - 2, j + 1);
      }
      const float ru_ck = top ? 0.f : ru_c, ru_ek = top ? 0.f : ru_e, rv_ck = top ? 0.f : rv_c, rv_nk = top ? 0.f : rv_n;
      const float u_ck = top ? 0.f : uc, u_ek = top ? 0.f : ue, v_ck = top ? 0.f : vc, v_nk = top ? 0.f : vn;
      const float p_k = top ? 0.f : p_c;
      // =====================================================================  rw_tend (k_mom_w_fast)
      {
        const float wy0 = ST0(F_W, 0, -3), wy1 = ST0(F_W, 0, -2), wy2 = ST0(F_W, 0, -1);
        const float wy4 = ST0(F_W, 0, 1), wy5 = ST0(F_W, 0, 2), wy6 = ST0(F_W, 0, 3);
        const float wx0 = ST0(F_W, -3, 0), wx1 = ST0(F_W, -2, 0), wx2 = ST0(F_W, -1, 0);
        const float wx4 = ST0(F_W, 1, 0), wx5 = ST0(F_W, 2, 0);
        const float fa = top ? (2.f - g.fnm[k - 1]) : g.fnm[k];
        const float fb = top ? g.fnp[k - 1] : g.fnp[k];
        float t = 0.f;
        {   // y flux divergence: vel = fzm*rv(k)+fzp*rv(k-1), extrapolated at the top
          const float vel0 = top ? (fa * rv_c_m - fb * rv_c_mm) : (fa * rv_ck + fb * rv_c_m);
          const float vel1 = top ? (fa * rv_n_m - fb * rv_n_mm) : (fa * rv_nk + fb * rv_n_m);
          const float fq0 = vel0 * flux5(wy0, wy1, wy2, w1, wy4, wy5, vel0, sg);
          const float fq1 = vel1 * flux5(wy1, wy2, w1, wy4, wy5, wy6, vel1, sg);
          t = t - t_mrdy * (fq1 - fq0);
        }
        {   // x flux divergence
          const float vel = top ? (fa * ru_c_m - fb * ru_c_mm) : (fa * ru_ck + fb * ru_c_m);
          const float fq0 = vel * flux5(wx0, wx1, wx2, w1, wx4, wx5, vel, sg);
          const float fq1 = __shfl_down_sync(0xffffffffu, fq0, 1);
          t = t - t_mrdx * (fq1 - fq0);
        }
        float vf1 = 0.f;   // vertical (:5996-6031), rdzu = rdn: face k+1 between w(k) and w(k+1)
        if (!top) {
          const int kk = k + 1;
          if (kk == kde) vf1 = 0.25f * (rom1_c + rom_k) * (w2 + w1);
          else {
            const float vel = 0.5f * (rom1_c + rom_k);
            vf1 = vel * flux3(w0, w1, w2, w3, -vel, sg);
          }
          t = t - g.rdn[k] * (vf1 - wvf);
        } else {
          t = t + 2.f * g.rdn[k - 1] * (wvf);
        }
        // pg_buoy_w (:2419-2499); cqw <- 1/(1+cqw) (:2489)
        if (top) {
          const float cq1 = 1.f / (1.f + cqw_m);
          const float cq2 = cqw_m * cq1;
          t = t + rmsf * WG_G * (cq1 * 2.f * g.rdnw[k - 1] * (-p_m) - (g.c1f[k] * mu_c) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
        } else {
          const float cq1 = 1.f / (1.f + cqw);
          const float cq2 = cqw * cq1;
          if (st) A3(s.cqw_out, i, k, j) = cq1;
          t = t + rmsf * WG_G * (cq1 * g.rdn[k] * (p_k - p_m) - (g.c1f[k] * mu_c) - cq2 * (g.c1f[k] * mub + g.c2f[k]));
        }
        if (!top) {
          if (g.w_damping == 1) {   // w_damp (:2635-2695)
            const float mass = (g.c1f[k] * mut + g.c2f[k]);
            const float vert_cfl = fabsf(rom_k / mass * g.rdnw[k] * g.dt);
            if (vert_cfl > g.w_damp_on) t = t - sign1(w1) * 0.3f * (vert_cfl - g.w_crit_cfl) * mass;
          }
          // coriolis (:3839-3844) and curvature (:4459-4463)
          const float fzm = g.fnm[k], fzp = g.fnp[k];
          t = t + e_c * (cosa_c * 0.5f * (fzm * (ru_ck + ru_ek) + fzp * (ru_c_m + ru_e_m)) -
                         w_mr * sina_c * 0.5f * (fzm * (rv_ck + rv_nk) + fzp * (rv_c_m + rv_n_m)));
          t = t + WG_RERADIUS * (0.5f * (fzm * (ru_ck + ru_ek) + fzp * (ru_c_m + ru_e_m)) * 0.5f * (fzm * (u_ck + u_ek) + fzp * (u_c_m + u_e_m)) +
                                 w_mr * 0.5f * (fzm * (rv_ck + rv_nk) + fzp * (rv_c_m + rv_n_m)) * 0.5f * (fzm * (v_ck + v_nk) + fzp * (v_c_m + v_n_m)));
        }
        t = t + div_msf(rw_tf, msfty, one_ty);
        if (st) A3(s.rw_tend, i, k, j) = t;
        wvf = vf1;
      }
      // =====================================================================  ph_tend (k_ph_fast)
      {
        const float c1f = g.c1f[k], c2f = g.c2f[k];
        const float ph_k = ST0(F_PH, 0, 0), phb_k = ST0(F_PHB, 0, 0);
        const float py1 = ST0(F_PH, 0, 1), pym1 = ST0(F_PH, 0, -1), py2 = ST0(F_PH, 0, 2), pym2 = ST0(F_PH, 0, -2);
        const float py3 = ST0(F_PH, 0, 3), pym3 = ST0(F_PH, 0, -3);
        const float by1 = ST0(F_PHB, 0, 1), bym1 = ST0(F_PHB, 0, -1), by2 = ST0(F_PHB, 0, 2), bym2 = ST0(F_PHB, 0, -2);
        const float by3 = ST0(F_PHB, 0, 3), bym3 = ST0(F_PHB, 0, -3);
        const float px1 = ST0(F_PH, 1, 0), pxm1 = ST0(F_PH, -1, 0), px2 = ST0(F_PH, 2, 0), pxm2 = ST0(F_PH, -2, 0);
        const float px3 = ST0(F_PH, 3, 0), pxm3 = ST0(F_PH, -3, 0);
        const float bx1 = ST0(F_PHB, 1, 0), bxm1 = ST0(F_PHB, -1, 0), bx2 = ST0(F_PHB, 2, 0), bxm2 = ST0(F_PHB, -2, 0);
        const float bx3 = ST0(F_PHB, 3, 0), bxm3 = ST0(F_PHB, -3, 0);
        const float wk = w1;
        float t = 0.f;
        float wd_k1 = 0.f;
        if (!top) {   // -omega dphi/deta (:1459-1494)
          const float ph1 = ST1(F_PH, 0, 0), phb1 = ST1(F_PHB, 0, 0);
          if (g.phi_adv_z == 2) {
            wd_k1 = g.rdnw[k] * (ph1 - ph_k + phb1 - phb_k);
            t = t - rom_k * (g.fnm[k] * wd_k1 + g.fnp[k] * wd_k);
          } else {
            wd_k1 = .5f * (rom1_c + rom_k) * g.rdnw[k] * (ph1 - ph_k + phb1 - phb_k);
            t = t - (g.fnm[k] * wd_k1 + g.fnp[k] * wd_k);
          }
        }
        t = t + div_msf((c1f * mut + c2f) * WG_G * wk, msfty, one_ty);   // "gw" term (:1496-1508)
        const float fac = top ? 0.5f : 0.25f;
        {   // y advection, 6th-order centred difference of ph+phb (:1783-1815)
          const float va = top ? (g.cfn * v_n_m + g.cfn1 * v_n_mm) : (v_nk + v_n_m);
          const float vb = top ? (g.cfn * v_c_m + g.cfn1 * v_c_mm) : (v_ck + v_c_m);
          const float ma = (c1f * muv_n + c2f), mb = (c1f * muv + c2f);
          t = t - (fac * g.rdy / msfty) *
                      ((ma * va * ya + mb * vb * msfvy) * (1.f / 60.f) *
                       (45.f * (py1 - pym1) - 9.f * (py2 - pym2) + (py3 - pym3) + 45.f * (by1 - bym1) - 9.f * (by2 - bym2) + (by3 - bym3)));
        }
        {   // x advection (:1940-1970)
          const float ua = top ? (g.cfn * u_e_m + g.cfn1 * u_e_mm) : (u_ek + u_e_m);
          const float ub = top ? (g.cfn * u_c_m + g.cfn1 * u_c_mm) : (u_ck + u_c_m);
          const float ma = (c1f * muu_e + c2f), mb = (c1f * muu + c2f);
          t = t - (fac * g.rdx / msfty) *
                      ((ma * ua * xa + mb * ub * msfux) * (1.f / 60.f) *
                       (45.f * (px1 - pxm1) - 9.f * (px2 - pxm2) + (px3 - pxm3) + 45.f * (bx1 - bxm1) - 9.f * (bx2 - bxm2) + (bx3 - bxm3)));
        }
        t = t + div_msf(ph_tf, msfty, one_ty);
        if (st) A3(s.ph_tend, i, k, j) = t;
        wd_k = wd_k1;
      }
    }
    // ---- roll
    um1 = u0; u0 = u1; u1 = u2;
    vm1 = v0; v0 = v1; v1 = v2;
    qm1 = q0; q0 = q1; q1 = q2;
    w0 = w1; w1 = w2; w2 = w3;
    ru_c_m = ru_c; ru_e_m = ru_e; rv_c_m = rv_c; rv_n_m = rv_n;
    u_c_m = uc; u_e_m = ue; v_c_m = vc; v_n_m = vn;
    p_m = p_c; cqw_m = cqw;
#undef ST0
#undef ST1
#undef H0
#undef H1
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("wrfgpu: cuTensorMapEncodeTiled is not available in this driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// tensor map of one 3-D field: dims (pitch, nk, nj), box (bw, 1, bh); coordinates are (i - ims, k - kms, j - jms)
static CUtensorMap field_map(Solver *s, const Field &f, int bw, int bh) {
  CUtensorMap m;
  const cuuint64_t np = s->pitch();
  cuuint64_t dims[3] = {np, (cuuint64_t)s->g.nk, (cuuint64_t)s->g.nj};
  cuuint64_t strides[2] = {np * sizeof(float), np * (cuuint64_t)s->g.nk * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)bw, 1u, (cuuint32_t)bh};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)f.d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("wrfgpu: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

bool tend_tile_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("WRFGPU_TEND");     // "fast": the five k-marching kernels (A/B measurements)
    on = (e && !strcmp(e, "fast")) ? 0 : 1;
  }
  return on == 1;
}

void launch_tend_tile(Solver *s, const Box &in) {
  if (empty(in)) return;
  if (s->coef_dirty) s->sync_coef();
  if (!s->tile_maps) {
    CUDA_CHECK(cudaFuncSetAttribute(k_tend_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM));   // per device
    auto *m = new TileMaps;
    const char *stn[TT_NST] = {"u_2", "v_2", "w_2", "t_2", "ph_2", "phb"};
    const char *h1n[TT_NH1] = {"ru", "rv", "ww", "rw", "p", "pb", "al", "alt", "php"};
    for (int f = 0; f < TT_NST; f++) m->st[f] = field_map(s, s->F(stn[f]), TT_SW, TT_SH);
    for (int f = 0; f < TT_NH1; f++) m->h1[f] = field_map(s, s->F(h1n[f]), TT_HW, TT_HH);
    s->tile_maps = m;
  }
  TileIn a;
  a.u = s->o("u_2"); a.v = s->o("v_2"); a.w = s->o("w_2"); a.t = s->o("t_2"); a.ww = s->o("ww");
  a.p = s->o("p"); a.ru = s->o("ru"); a.rv = s->o("rv");
  a.cqu = s->o("cqu"); a.cqv = s->o("cqv"); a.cqw = s->o("cqw"); a.h_diabatic = s->o("h_diabatic");
  a.mut = s->o("mut"); a.muu = s->o("muu"); a.muv = s->o("muv"); a.mu = s->o("mu_2"); a.mub = s->o("mub");
  a.ru_tendf = s->o("ru_tendf"); a.rv_tendf = s->o("rv_tendf"); a.rw_tendf = s->o("rw_tendf"); a.ph_tendf = s->o("ph_tendf");
  a.t_tendf = s->o("t_tendf");
  a.ru_tend = s->o("ru_tend"); a.rv_tend = s->o("rv_tend"); a.rw_tend = s->o("rw_tend"); a.ph_tend = s->o("ph_tend");
  a.t_tend = s->o("t_tend"); a.cqw_out = s->o("cqw");
  const int ni = in.i1 - in.i0 + 1, nj = in.j1 - in.j0 + 1;
  dim3 block(TT_TX, TT_TY), grid((unsigned)((ni + TT_OX - 1) / TT_OX), (unsigned)((nj + TT_TY - 1) / TT_TY));
  s->kpre("k_tend_tile");
  k_tend_tile<<<grid, block, TT_SMEM, s->stream>>>(s->g, in, *(const TileMaps *)s->tile_maps, a);
  s->kpost();
  s->launches++;
}

void tend_tile_destroy(Solver *s) {
  delete (TileMaps *)s->tile_maps;
  s->tile_maps = nullptr;
}

}  // namespace wg
