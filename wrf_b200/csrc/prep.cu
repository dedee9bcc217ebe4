// prep.cu -- rk_step_prep (dyn_em/module_em.F:37-186) fused into three kernels:
//   k_prep_mu  : calculate_full(mut) + calc_mu_uv        (2-D)
//   k_prep_3d  : couple_momentum + calc_cq + calc_alt + calc_php, one pass
//   k_calc_ww  : calc_ww_cp column integral (k prefix sum kept in registers)
// The reference makes seven separate sweeps (module_big_step_utilities_em.F:26,329,
// 640,787,910,1227,3587).  Algorithmic bytes: 16+n_moist field accesses per point.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace wg {

// module_big_step_utilities_em.F:3587 (mut) and :26-180 (muu,muv incl. the domain-edge special cases)
__global__ void k_prep_mu(const Geo g, Box b, cfp mu, cfp mub, fp mut, fp muu, fp muv) {
  TID2;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1);
  if (im && jm) A2(mut, i, j) = A2(mub, i, j) + A2(mu, i, j);
  if (jm) {
    float r;
    if (i == g.ids) {
      const int q = g.periodic_x ? g.ids - 1 : g.ids;
      r = 0.5f * (A2(mu, i, j) + A2(mu, q, j) + A2(mub, i, j) + A2(mub, q, j));
    } else if (i == g.ide) {
      const int q = g.periodic_x ? g.ide : g.ide - 1;
      r = 0.5f * (A2(mu, i - 1, j) + A2(mu, q, j) + A2(mub, i - 1, j) + A2(mub, q, j));
    } else {
      r = 0.5f * (A2(mu, i, j) + A2(mu, i - 1, j) + A2(mub, i, j) + A2(mub, i - 1, j));
    }
    A2(muu, i, j) = r;
  }
  if (im) {
    float r;
    if (j == g.jds) {
      const int q = g.periodic_y ? g.jds - 1 : g.jds;
      r = 0.5f * (A2(mu, i, j) + A2(mu, i, q) + A2(mub, i, j) + A2(mub, i, q));
    } else if (j == g.jde) {
      const int q = g.periodic_y ? g.jde : g.jde - 1;
      r = 0.5f * (A2(mu, i, j - 1) + A2(mu, i, q) + A2(mub, i, j - 1) + A2(mub, i, q));
    } else {
      r = 0.5f * (A2(mu, i, j) + A2(mu, i, j - 1) + A2(mub, i, j) + A2(mub, i, j - 1));
    }
    A2(muv, i, j) = r;
  }
}

struct MoistPtrs { const float *m[3]; int n; };

// couple_momentum :329, calc_cq :787, calc_alt :910, calc_php :1227
__global__ void k_prep_3d(const Geo g, Box b, cfp u, cfp v, cfp w, cfp muu, cfp muv, cfp mut, fp ru, fp rv, fp rw,
                          MoistPtrs mp, fp cqu, fp cqv, fp cqw, cfp al, cfp alb, fp alt, cfp ph, cfp phb, fp php) {
  TID3;
  const int ktf = g.kde - 1;
  const bool im = i <= min(g.ite, g.ide - 1), jm = j <= min(g.jte, g.jde - 1), km = k <= ktf;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  if (km && jm) {
    A3(ru, i, k, j) = A3(u, i, k, j) * (c1h * A2(muu, i, j) + c2h) / A2(g.msfuy, i, j);
    float r = 1.f;
    if (mp.n >= 1) {
      float qtot = 0.f;
      for (int s = 0; s < mp.n; s++) qtot = qtot + A3(mp.m[s], i, k, j) + A3(mp.m[s], i - 1, k, j);
      r = 1.f / (1.f + 0.5f * qtot);
    }
    A3(cqu, i, k, j) = r;
  }
  if (km && im) {
    A3(rv, i, k, j) = A3(v, i, k, j) * (c1h * A2(muv, i, j) + c2h) * A2(g.msfvx_inv, i, j);
    float r = 1.f;
    if (mp.n >= 1) {
      float qtot = 0.f;
      for (int s = 0; s < mp.n; s++) qtot = qtot + A3(mp.m[s], i, k, j) + A3(mp.m[s], i, k, j - 1);
      r = 1.f / (1.f + 0.5f * qtot);
    }
    A3(cqv, i, k, j) = r;
  }
  if (im && jm) {
    A3(rw, i, k, j) = A3(w, i, k, j) * (g.c1f[k] * A2(mut, i, j) + g.c2f[k]) / A2(g.msfty, i, j);
    if (km) {
      if (k >= 2) {
        float r = 0.f;
        if (mp.n >= 1) {
          float qtot = 0.f;
          for (int s = 0; s < mp.n; s++) qtot = qtot + A3(mp.m[s], i, k, j) + A3(mp.m[s], i, k - 1, j);
          r = 0.5f * qtot;
        }
        A3(cqw, i, k, j) = r;
      }
      A3(alt, i, k, j) = A3(al, i, k, j) + A3(alb, i, k, j);
      A3(php, i, k, j) = 0.5f * (A3(phb, i, k, j) + A3(phb, i, k + 1, j) + A3(ph, i, k, j) + A3(ph, i, k + 1, j));
    }
  }
}

// the same with 16-byte accesses (four columns per thread, see ld4 / st4 in common.cuh)
__global__ void k_prep_3d_v4(const Geo g, Box b, cfp u, cfp v, cfp w, cfp muu, cfp muv, cfp mut, fp ru, fp rv, fp rw,
                             MoistPtrs mp, fp cqu, fp cqv, fp cqw, cfp al, cfp alb, fp alt, cfp ph, cfp phb, fp php) {
  TID3_V4;
  const int ktf = g.kde - 1;
  const bool jm = j <= min(g.jte, g.jde - 1), km = k <= ktf;
  const float c1h = g.c1h[k], c2h = g.c2h[k];
  float x[4], m[4], ms[4], out[4];
  float mq[3][4];           // the species at (i, k, j), reused by the three cq averages
  if (km && mp.n >= 1)
#pragma unroll
    for (int s = 0; s < 3; s++)
      if (s < mp.n) ld4(&A3(mp.m[s], i, k, j), mq[s]);
  if (km && jm) {
    ld4(&A3(u, i, k, j), x); ld4(&A2(muu, i, j), m); ld4(&A2(g.msfuy, i, j), ms);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = x[q] * (c1h * m[q] + c2h) / ms[q];
    st4(&A3(ru, i, k, j), out, lo, hi);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = 1.f;
    if (mp.n >= 1) {
      float qtot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 3; s++) {
        if (s >= mp.n) continue;
        const float west = A3(mp.m[s], i - 1, k, j);
#pragma unroll
        for (int q = 0; q < 4; q++) qtot[q] = qtot[q] + mq[s][q] + (q == 0 ? west : mq[s][q > 0 ? q - 1 : 0]);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) out[q] = 1.f / (1.f + 0.5f * qtot[q]);
    }
    st4(&A3(cqu, i, k, j), out, lo, hi);
  }
  if (km && him >= lo) {
    ld4(&A3(v, i, k, j), x); ld4(&A2(muv, i, j), m); ld4(&A2(g.msfvx_inv, i, j), ms);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = x[q] * (c1h * m[q] + c2h) * ms[q];
    st4(&A3(rv, i, k, j), out, lo, him);
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = 1.f;
    if (mp.n >= 1) {
      float qtot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 3; s++) {
        if (s >= mp.n) continue;
        float south[4];
        ld4(&A3(mp.m[s], i, k, j - 1), south);
#pragma unroll
        for (int q = 0; q < 4; q++) qtot[q] = qtot[q] + mq[s][q] + south[q];
      }
#pragma unroll
      for (int q = 0; q < 4; q++) out[q] = 1.f / (1.f + 0.5f * qtot[q]);
    }
    st4(&A3(cqv, i, k, j), out, lo, him);
  }
  if (jm && him >= lo) {
    ld4(&A3(w, i, k, j), x); ld4(&A2(mut, i, j), m); ld4(&A2(g.msfty, i, j), ms);
    const float c1f = g.c1f[k], c2f = g.c2f[k];
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = x[q] * (c1f * m[q] + c2f) / ms[q];
    st4(&A3(rw, i, k, j), out, lo, him);
    if (km) {
      if (k >= 2) {
#pragma unroll
        for (int q = 0; q < 4; q++) out[q] = 0.f;
        if (mp.n >= 1) {
          float qtot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int s = 0; s < 3; s++) {
            if (s >= mp.n) continue;
            float below[4];
            ld4(&A3(mp.m[s], i, k - 1, j), below);
#pragma unroll
            for (int q = 0; q < 4; q++) qtot[q] = qtot[q] + mq[s][q] + below[q];
          }
#pragma unroll
          for (int q = 0; q < 4; q++) out[q] = 0.5f * qtot[q];
        }
        st4(&A3(cqw, i, k, j), out, lo, him);
      }
      float a0[4], a1[4], p0[4], p1[4], b0[4], b1[4];
      ld4(&A3(al, i, k, j), a0); ld4(&A3(alb, i, k, j), a1);
#pragma unroll
      for (int q = 0; q < 4; q++) out[q] = a0[q] + a1[q];
      st4(&A3(alt, i, k, j), out, lo, him);
      ld4(&A3(phb, i, k, j), b0); ld4(&A3(phb, i, k + 1, j), b1); ld4(&A3(ph, i, k, j), p0); ld4(&A3(ph, i, k + 1, j), p1);
#pragma unroll
      for (int q = 0; q < 4; q++) out[q] = 0.5f * (b0[q] + b1[q] + p0[q] + p1[q]);
      st4(&A3(php, i, k, j), out, lo, him);
    }
  }
}

// calc_ww_cp, module_big_step_utilities_em.F:640-782.  One thread per column; pass 1
// stores divv(k) in ww(k+1) while summing dmdt, pass 2 turns it into ww in place.
#ifndef WW_MINB
#define WW_MINB 16    // r1w: 6.5 ms per step at 16 resident blocks, 7.5 at 12, 10.6 at 8
#endif
#ifndef WW_UNROLL
#define WW_UNROLL 4
#endif
#define WW_PRAGMA_(x) _Pragma(#x)
#define WW_PRAGMA(x) WW_PRAGMA_(x)
__global__ void __launch_bounds__(128, WW_MINB) k_calc_ww(const Geo g, Box b, cfp u, cfp v, cfp mup, cfp mub, fp ww) {
  TID2;
  const int ktf = g.kde - 1, kte = g.kde;
  if (i > min(g.ite, g.ide - 1)) {   // i = ide column: only the ww(:,1,:) / ww(:,kte,:) reset (:744-749)
    A3(ww, i, 1, j) = 0.f;
    A3(ww, i, kte, j) = 0.f;
    return;
  }
  const float muu0 = 0.5f * (A2(mup, i, j) + A2(mub, i, j) + A2(mup, i - 1, j) + A2(mub, i - 1, j));
  const float muu1 = 0.5f * (A2(mup, i + 1, j) + A2(mub, i + 1, j) + A2(mup, i, j) + A2(mub, i, j));
  const float muv0 = 0.5f * (A2(mup, i, j) + A2(mub, i, j) + A2(mup, i, j - 1) + A2(mub, i, j - 1));
  const float muv1 = 0.5f * (A2(mup, i, j + 1) + A2(mub, i, j + 1) + A2(mup, i, j) + A2(mub, i, j));
  const float msftx = A2(g.msftx, i, j);
  const float my0 = A2(g.msfuy, i, j), my1 = A2(g.msfuy, i + 1, j);
  const float mvi0 = A2(g.msfvx_inv, i, j), mvi1 = A2(g.msfvx_inv, i, j + 1);
  float dmdt = 0.f;
  WW_PRAGMA(unroll WW_UNROLL)
  for (int k = 1; k <= ktf; k++) {
    const float c1h = g.c1h[k], c2h = g.c2h[k];
    const float divv = msftx * g.dnw[k] *
                       (g.rdx * ((c1h * muu1 + c2h) * A3(u, i + 1, k, j) / my1 - (c1h * muu0 + c2h) * A3(u, i, k, j) / my0) +
                        g.rdy * ((c1h * muv1 + c2h) * A3(v, i, k, j + 1) * mvi1 - (c1h * muv0 + c2h) * A3(v, i, k, j) * mvi0));
    dmdt = dmdt + divv;
    A3(ww, i, k + 1, j) = divv;
  }
  float wprev = 0.f;
  A3(ww, i, 1, j) = 0.f;
  WW_PRAGMA(unroll WW_UNROLL)
  for (int k = 2; k <= ktf; k++) {
    const float divv = A3(ww, i, k, j);
    wprev = wprev - g.dnw[k - 1] * g.c1h[k - 1] * dmdt - divv;
    A3(ww, i, k, j) = wprev;
  }
  A3(ww, i, kte, j) = 0.f;
}

// calc_ww_cp with four columns per thread: 16-byte accesses, four independent column sums in flight
__global__ void __launch_bounds__(128, 6) k_calc_ww_v4(const Geo g, Box b, cfp u, cfp v, cfp mup, cfp mub, fp ww) {
  const int gi_ = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (gi_ > b.i1 || j > b.j1) return;
  const int i = g.ims + 4 * gi_;
  const int lo = max(g.its - i, 0), hi = min(g.ite - i, 3), him = min(min(g.ite, g.ide - 1) - i, 3);
  const int ktf = g.kde - 1, kte = g.kde;
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  if (hi > him) {   // the i = ide column of the tile: only the ww(:,1,:) / ww(:,kte,:) reset (:744-749)
    st4(&A3(ww, i, 1, j), zero, max(him + 1, lo), hi);
    st4(&A3(ww, i, kte, j), zero, max(him + 1, lo), hi);
  }
  if (him < lo) return;
  float mp[4], mb[4], mps[4], mbs[4], mpn[4], mbn[4], msftx[4], my[4], mvi0[4], mvi1[4];
  ld4(&A2(mup, i, j), mp); ld4(&A2(mub, i, j), mb); ld4(&A2(mup, i, j - 1), mps); ld4(&A2(mub, i, j - 1), mbs);
  ld4(&A2(mup, i, j + 1), mpn); ld4(&A2(mub, i, j + 1), mbn); ld4(&A2(g.msftx, i, j), msftx); ld4(&A2(g.msfuy, i, j), my);
  ld4(&A2(g.msfvx_inv, i, j), mvi0); ld4(&A2(g.msfvx_inv, i, j + 1), mvi1);
  const float mp_w = A2(mup, i - 1, j), mb_w = A2(mub, i - 1, j), mp_e = A2(mup, i + 4, j), mb_e = A2(mub, i + 4, j), my_e = A2(g.msfuy, i + 4, j);
  float muu0[4], muu1[4], muv0[4], muv1[4], my1[4], dmdt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float pw = q == 0 ? mp_w : mp[q > 0 ? q - 1 : 0], bw = q == 0 ? mb_w : mb[q > 0 ? q - 1 : 0];
    const float pe = q == 3 ? mp_e : mp[q < 3 ? q + 1 : 3], be = q == 3 ? mb_e : mb[q < 3 ? q + 1 : 3];
    muu0[q] = 0.5f * (mp[q] + mb[q] + pw + bw);
    muu1[q] = 0.5f * (pe + be + mp[q] + mb[q]);
    muv0[q] = 0.5f * (mp[q] + mb[q] + mps[q] + mbs[q]);
    muv1[q] = 0.5f * (mpn[q] + mbn[q] + mp[q] + mb[q]);
    my1[q] = q == 3 ? my_e : my[q < 3 ? q + 1 : 3];
  }
  for (int k = 1; k <= ktf; k++) {
    const float c1h = g.c1h[k], c2h = g.c2h[k];
    float ug[4], v0[4], v1[4], divv[4];
    ld4(&A3(u, i, k, j), ug); ld4(&A3(v, i, k, j), v0); ld4(&A3(v, i, k, j + 1), v1);
    const float u_e = A3(u, i + 4, k, j);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float ue = q == 3 ? u_e : ug[q < 3 ? q + 1 : 3];
      divv[q] = msftx[q] * g.dnw[k] *
                (g.rdx * ((c1h * muu1[q] + c2h) * ue / my1[q] - (c1h * muu0[q] + c2h) * ug[q] / my[q]) +
                 g.rdy * ((c1h * muv1[q] + c2h) * v1[q] * mvi1[q] - (c1h * muv0[q] + c2h) * v0[q] * mvi0[q]));
      dmdt[q] = dmdt[q] + divv[q];
    }
    st4(&A3(ww, i, k + 1, j), divv, lo, him);
  }
  float wprev[4] = {0.f, 0.f, 0.f, 0.f};
  st4(&A3(ww, i, 1, j), zero, lo, him);
  for (int k = 2; k <= ktf; k++) {
    float divv[4];
    ld4(&A3(ww, i, k, j), divv);
#pragma unroll
    for (int q = 0; q < 4; q++) wprev[q] = wprev[q] - g.dnw[k - 1] * g.c1h[k - 1] * dmdt[q] - divv[q];
    st4(&A3(ww, i, k, j), wprev, lo, him);
  }
  st4(&A3(ww, i, kte, j), zero, lo, him);
}

void Solver::rk_step_prep(int) {
  const Geo &q = g;
  LAUNCH2(this, k_prep_mu, (Box{q.its, q.ite, 1, 1, q.jts, q.jte}), o("mu_2"), o("mub"), o("mut"), o("muu"), o("muv"));
  MoistPtrs mp;
  mp.n = c.num_moist;
  static const char *mn[] = {"moist_qv", "moist_qc", "moist_qr"};
  for (int s = 0; s < 3; s++) mp.m[s] = s < mp.n ? o(mn[s]) : nullptr;
  static const bool vec4 = [] { const char *e = getenv("WRFGPU_SSP"); return !(e && !strcmp(e, "scalar")); }();
  if (vec4)
    LAUNCH3(this, k_prep_3d_v4, box_v4(q, 1, q.kde), o("u_2"), o("v_2"), o("w_2"), o("muu"), o("muv"),
            o("mut"), o("ru"), o("rv"), o("rw"), mp, o("cqu"), o("cqv"), o("cqw"), o("al"), o("alb"), o("alt"), o("ph_2"),
            o("phb"), o("php"));
  else
    LAUNCH3(this, k_prep_3d, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), o("u_2"), o("v_2"), o("w_2"), o("muu"), o("muv"),
            o("mut"), o("ru"), o("rv"), o("rw"), mp, o("cqu"), o("cqv"), o("cqw"), o("al"), o("alb"), o("alt"), o("ph_2"),
            o("phb"), o("php"));
  if (vec4) {
    Box bx = box_v4(q, 1, 1);
    bx.j1 = std::min(q.jte, q.jde - 1);
    LAUNCH2(this, k_calc_ww_v4, bx, o("u_2"), o("v_2"), o("mu_2"), o("mub"), o("ww"));
  } else
    LAUNCH2(this, k_calc_ww, (Box{q.its, q.ite, 1, 1, q.jts, std::min(q.jte, q.jde - 1)}), o("u_2"), o("v_2"), o("mu_2"),
            o("mub"), o("ww"));
}

// solve_em.F:726-754 (rk_phys_bc_dry_1 = module_bc_em.F:960-1041, then rho, al, ph_2)
void Solver::rk_phys_bc_dry_1() {
  bc({{"ru", 'u'}, {"rv", 'v'}, {"rw", 'w'}, {"ww", 'w'}, {"php", 'w'}, {"alt", 't'}, {"p", 'p'},
      {"muu", 'u'}, {"muv", 'v'}, {"mut", 't'}, {"rho", 'p'}, {"al", 'p'}, {"ph_2", 'w'}});
}

}  // namespace wg
