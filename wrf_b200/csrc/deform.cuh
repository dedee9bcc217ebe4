// deform.cuh -- point functions of the diffusion preamble: compute_diff_metrics
// (dyn_em/module_diffusion_em.F:6882-7129), cal_deform_and_div (:17-1189) and smag2d_km (:1934-2043).
//
// The reference sweeps tile-sized work arrays (hat, hatavg, tmp1) through a dozen loop nests per tensor component.
// Every one of them is a pure function of a few neighbouring points of u, v, w and the metrics, so here the thread
// of a point evaluates them in registers, in the reference's order of operations, and writes each output once:
// no work arrays, one pass over memory.  The functions are __host__ __device__ so that tests/deform_shim.cu can run
// exactly this code on the CPU against the oracle.
//
// The reference finishes defor12 / defor13 / defor23 by copying the first computed row / column onto the domain
// edge it skipped (:826-870, :1143-1187), in the order west, south, east, north.  The net effect of the four copies
// is "take the value of the nearest computed point", which is what the thread of an edge point evaluates (DefR holds
// the computed ranges and which edges are copied).  Elements the reference never writes are not written here either.
#pragma once
#include "common.cuh"
#include "libm_port.h"

namespace wg {

// WG_HD (__host__ __device__ __forceinline__ under nvcc) comes from libm_port.h

struct DefIn {
  cfp u, v, w, rdz, rdzw, zx, zy;
  const float *dn, *dnw, *u_base, *v_base;   // origin shifted 1-D
};
struct DefOut { fp div, defor11, defor22, defor33, defor12, defor13, defor23; };

// loop ranges of cal_deform_and_div on this tile
struct DefR {
  int a_ie, a_je;                              // defor11/22/33, div: its..a_ie, jts..a_je
  int i0_12, i1_12, j0_12, j1_12;              // defor12 computed range
  int cxs12, cys12, cxe12, cye12;              // ... and its edge copies
  int i0_13, i1_13, j0_13, j1_13;              // defor13
  int i0_23, i1_23, j0_23, j1_23;              // defor23
  int cxs, cys, cxe, cye;                      // edge copies of defor13 and defor23 (ranges of the defor23 block)
  int ktf, kte, mix_full_fields;
};

inline DefR def_ranges(const GeoBase &g, int mix_full_fields) {
  DefR r;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  r.kte = g.kde; r.ktf = g.kde - 1; r.mix_full_fields = mix_full_fields;
  r.a_ie = ite < ide - 1 ? ite : ide - 1; r.a_je = jte < jde - 1 ? jte : jde - 1;
  int i_start = its, i_end = ite, j_start = jts, j_end = jte;
  if (g.oxs()) i_start = ids + 1 > its ? ids + 1 : its;
  if (g.oxe()) i_end = ide - 1 < ite ? ide - 1 : ite;
  if (g.oys()) j_start = jds + 1 > jts ? jds + 1 : jts;
  if (g.oye()) j_end = jde - 1 < jte ? jde - 1 : jte;
  if (g.periodic_x) { i_start = its; i_end = ite; }
  r.i0_12 = i_start; r.i1_12 = i_end; r.j0_12 = j_start; r.j1_12 = j_end;
  r.cxs12 = !g.periodic_x && i_start == ids + 1; r.cys12 = !g.periodic_y && j_start == jds + 1;
  r.cxe12 = !g.periodic_x && i_end == ide - 1; r.cye12 = !g.periodic_y && j_end == jde - 1;
  i_start = its; i_end = r.a_ie; j_start = jts; j_end = r.a_je;
  if (g.oxs()) i_start = ids + 1 > its ? ids + 1 : its;
  if (g.oys()) j_start = jds + 1 > jts ? jds + 1 : jts;
  if (g.periodic_x) i_start = its;
  if (g.periodic_y) j_end = jte < jde ? jte : jde;
  r.i0_23 = i_start; r.i1_23 = g.periodic_x ? r.a_ie : i_end; r.j0_23 = j_start; r.j1_23 = j_end;
  r.i0_13 = i_start; r.i1_13 = g.periodic_x ? (ite < ide ? ite : ide) : i_end; r.j0_13 = j_start; r.j1_13 = j_end;
  r.cxs = !g.periodic_x && r.i0_23 == ids + 1; r.cys = !g.periodic_y && r.j0_23 == jds + 1;
  r.cxe = !g.periodic_x && r.i1_23 == ide - 1; r.cye = !g.periodic_y && r.j1_23 == jde - 1;
  return r;
}

// ------------------------------------------------------------------------------------------ compute_diff_metrics
// thread box: i in its..ite, k in 1..kte, j in jts-1..jte
WG_HD void diff_metrics_point(const Geo &g, int i, int k, int j, cfp ph, cfp phb, fp z, fp rdz, fp rdzw, fp zx, fp zy) {
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const int ktf = g.kde - 1;
  const int i_end = min(ite, ide - 1), j_end = min(jte, jde - 1);
  const float rdx = g.rdx, rdy = g.rdy;
  if (i <= i_end && k <= ktf) {   // rows jts-1..jte (:6928-6965)
    const float zw0 = (A3(ph, i, k, j) + A3(phb, i, k, j)) / WG_G, zw1 = (A3(ph, i, k + 1, j) + A3(phb, i, k + 1, j)) / WG_G;
    A3(rdzw, i, k, j) = 1.0f / (zw1 - zw0);
    if (k == 1) A3(rdz, i, 1, j) = 2.f / (zw1 - zw0);
    else {
      const float zwm = (A3(ph, i, k - 1, j) + A3(phb, i, k - 1, j)) / WG_G;
      A3(rdz, i, k, j) = 2.0f / (zw1 - zwm);
    }
  }
  if (j < jts) return;
  // zx (:6972-6990; columns ids and ide :7010-7055)
  if (j <= j_end) {
    if (i >= max(ids + 1, its) && i <= i_end) {
      float v = rdx * (A3(phb, i, k, j) - A3(phb, i - 1, k, j)) / WG_G;
      v = v + rdx * (A3(ph, i, k, j) - A3(ph, i - 1, k, j)) / WG_G;
      A3(zx, i, k, j) = v;
    } else if (((i == ids && its == ids) || (i == ide && ite == ide)) && k <= ktf) {
      float v = 0.0f;
      if (g.periodic_x) {
        v = rdx * (A3(phb, i, k, j) - A3(phb, i - 1, k, j)) / WG_G;
        v = v + rdx * (A3(ph, i, k, j) - A3(ph, i - 1, k, j)) / WG_G;
      }
      A3(zx, i, k, j) = v;
    }
  }
  // zy (:6992-7008; rows jds and jde :7057-7110)
  if (i <= i_end) {
    if (j >= max(jds + 1, jts) && j <= j_end) {
      float v = rdy * (A3(phb, i, k, j) - A3(phb, i, k, j - 1)) / WG_G;
      v = v + rdy * (A3(ph, i, k, j) - A3(ph, i, k, j - 1)) / WG_G;
      A3(zy, i, k, j) = v;
    } else if (((j == jds && jts == jds) || (j == jde && jte == jde)) && k <= ktf) {
      float v = 0.0f;
      if (g.periodic_y) {
        v = rdy * (A3(phb, i, k, j) - A3(phb, i, k, j - 1)) / WG_G;
        v = v + rdy * (A3(ph, i, k, j) - A3(ph, i, k, j - 1)) / WG_G;
      }
      A3(zy, i, k, j) = v;
    }
    if (j <= j_end && k <= ktf) A3(z, i, k, j) = 0.5f * (A3(ph, i, k, j) + A3(phb, i, k, j) + A3(ph, i, k + 1, j) + A3(phb, i, k + 1, j)) / WG_G;
  }
}

// ------------------------------------------------------------------------------------------ cal_deform_and_div
WG_HD float hat_(const Geo &g, cfp vel, cfp msf, int i, int k, int j) { return A3(vel, i, k, j) / A2(msf, i, j); }
// hatavg of the two columns A = (ia,ja), B = (ib,jb) at w-level k (:230-262 and the three blocks like it)
WG_HD float hatavg_(const Geo &g, cfp vel, cfp msf, int ia, int ja, int ib, int jb, int k, int kte, float cft1, float cft2) {
  if (k == 1)
    return 0.5f * (g.cf1 * hat_(g, vel, msf, ia, 1, ja) + g.cf2 * hat_(g, vel, msf, ia, 2, ja) + g.cf3 * hat_(g, vel, msf, ia, 3, ja) +
                   g.cf1 * hat_(g, vel, msf, ib, 1, jb) + g.cf2 * hat_(g, vel, msf, ib, 2, jb) + g.cf3 * hat_(g, vel, msf, ib, 3, jb));
  if (k == kte)
    return 0.5f * (cft1 * (hat_(g, vel, msf, ia, kte - 1, ja) + hat_(g, vel, msf, ib, kte - 1, jb)) +
                   cft2 * (hat_(g, vel, msf, ia, kte - 2, ja) + hat_(g, vel, msf, ib, kte - 2, jb)));
  return 0.5f * (g.fnm[k] * (hat_(g, vel, msf, ia, k, ja) + hat_(g, vel, msf, ib, k, jb)) +
                 g.fnp[k] * (hat_(g, vel, msf, ia, k - 1, ja) + hat_(g, vel, msf, ib, k - 1, jb)));
}

WG_HD float defor12_at(const Geo &g, const DefIn &s, int i, int k, int j, int kte, float cft1, float cft2) {
  cfp zx = s.zx, zy = s.zy, rdzw = s.rdzw;
  const float mm = 0.25f * (A2(g.msfux, i, j - 1) + A2(g.msfux, i, j)) * (A2(g.msfvy, i - 1, j) + A2(g.msfvy, i, j));
  const float tmpzy = 0.25f * (A3(zy, i - 1, k, j) + A3(zy, i, k, j) + A3(zy, i - 1, k + 1, j) + A3(zy, i, k + 1, j));
  float tmp1 = (hatavg_(g, s.u, g.msfux, i, j - 1, i, j, k + 1, kte, cft1, cft2) - hatavg_(g, s.u, g.msfux, i, j - 1, i, j, k, kte, cft1, cft2)) *
               0.25f * tmpzy * (A3(rdzw, i, k, j) + A3(rdzw, i - 1, k, j) + A3(rdzw, i - 1, k, j - 1) + A3(rdzw, i, k, j - 1));
  float d12 = mm * (g.rdy * (hat_(g, s.u, g.msfux, i, k, j) - hat_(g, s.u, g.msfux, i, k, j - 1)) - tmp1);
  const float tmpzx = 0.25f * (A3(zx, i, k, j - 1) + A3(zx, i, k, j) + A3(zx, i, k + 1, j - 1) + A3(zx, i, k + 1, j));
  tmp1 = (hatavg_(g, s.v, g.msfvy, i - 1, j, i, j, k + 1, kte, cft1, cft2) - hatavg_(g, s.v, g.msfvy, i - 1, j, i, j, k, kte, cft1, cft2)) *
         0.25f * tmpzx * (A3(rdzw, i, k, j) + A3(rdzw, i, k, j - 1) + A3(rdzw, i - 1, k, j - 1) + A3(rdzw, i - 1, k, j));
  d12 = d12 + mm * (g.rdx * (hat_(g, s.v, g.msfvy, i, k, j) - hat_(g, s.v, g.msfvy, i - 1, k, j)) - tmp1);
  return d12;
}

// k in 1..kte; 0 at k = 1 and k = kte
WG_HD float defor13_at(const Geo &g, const DefIn &s, int i, int k, int j, int kte, int mix_full_fields) {
  if (k == 1 || k == kte) return 0.0f;
  cfp w = s.w, rdz = s.rdz, u = s.u;
  const float mm = A2(g.msfux, i, j) * A2(g.msfuy, i, j);
  const float h00 = hat_(g, w, g.msfty, i, k, j), h0p = hat_(g, w, g.msfty, i, k + 1, j), hm0 = hat_(g, w, g.msfty, i - 1, k, j),
              hmp = hat_(g, w, g.msfty, i - 1, k + 1, j), h0m = hat_(g, w, g.msfty, i, k - 1, j), hmm = hat_(g, w, g.msfty, i - 1, k - 1, j);
  const float hak = 0.25f * (h00 + h0p + hm0 + hmp), hakm = 0.25f * (h0m + h00 + hmm + hm0);
  float tmp1 = (hak - hakm) * A3(s.zx, i, k, j) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i - 1, k, j));
  float d = mm * (g.rdx * (h00 - hm0) - tmp1);
  if (mix_full_fields) tmp1 = (A3(u, i, k, j) - A3(u, i, k - 1, j)) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i - 1, k, j));
  else tmp1 = (A3(u, i, k, j) - s.u_base[k] - A3(u, i, k - 1, j) + s.u_base[k - 1]) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i - 1, k, j));
  return d + tmp1;
}
WG_HD float defor23_at(const Geo &g, const DefIn &s, int i, int k, int j, int kte, int mix_full_fields) {
  if (k == 1 || k == kte) return 0.0f;
  cfp w = s.w, rdz = s.rdz, v = s.v;
  const float mm = A2(g.msfvx, i, j) * A2(g.msfvy, i, j);
  const float h00 = hat_(g, w, g.msftx, i, k, j), h0p = hat_(g, w, g.msftx, i, k + 1, j), hm0 = hat_(g, w, g.msftx, i, k, j - 1),
              hmp = hat_(g, w, g.msftx, i, k + 1, j - 1), h0m = hat_(g, w, g.msftx, i, k - 1, j), hmm = hat_(g, w, g.msftx, i, k - 1, j - 1);
  const float hak = 0.25f * (h00 + h0p + hm0 + hmp), hakm = 0.25f * (h0m + h00 + hmm + hm0);
  float tmp1 = (hak - hakm) * A3(s.zy, i, k, j) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i, k, j - 1));
  float d = mm * (g.rdy * (h00 - hm0) - tmp1);
  if (mix_full_fields) tmp1 = (A3(v, i, k, j) - A3(v, i, k - 1, j)) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i, k, j - 1));
  else tmp1 = (A3(v, i, k, j) - s.v_base[k] - A3(v, i, k - 1, j) + s.v_base[k - 1]) * 0.5f * (A3(rdz, i, k, j) + A3(rdz, i, k, j - 1));
  return d + tmp1;
}

// thread box: i in its..ite, k in 1..kte, j in jts..jte
WG_HD void deform_point(const Geo &g, const DefR &r, const DefIn &s, const DefOut &o, int i, int k, int j) {
  const int kte = r.kte, ktf = r.ktf;
  const float cft2 = -0.5f * s.dnw[kte - 1] / s.dn[kte - 1];
  const float cft1 = 1.0f - cft2;
  cfp zx = s.zx, zy = s.zy, rdzw = s.rdzw, w = s.w;
  // defor11, defor22, defor33, div (:205-609)
  if (i <= r.a_ie && j <= r.a_je && k <= ktf) {
    const float mm = A2(g.msftx, i, j) * A2(g.msfty, i, j);
    const float tmpzx = 0.25f * (A3(zx, i, k, j) + A3(zx, i + 1, k, j) + A3(zx, i, k + 1, j) + A3(zx, i + 1, k + 1, j));
    float tmp1 = (hatavg_(g, s.u, g.msfuy, i, j, i + 1, j, k + 1, kte, cft1, cft2) - hatavg_(g, s.u, g.msfuy, i, j, i + 1, j, k, kte, cft1, cft2)) *
                 tmpzx * A3(rdzw, i, k, j);
    tmp1 = mm * (g.rdx * (hat_(g, s.u, g.msfuy, i + 1, k, j) - hat_(g, s.u, g.msfuy, i, k, j)) - tmp1);
    A3(o.defor11, i, k, j) = 2.0f * tmp1;
    float div = tmp1;
    const float tmpzy = 0.25f * (A3(zy, i, k, j) + A3(zy, i, k, j + 1) + A3(zy, i, k + 1, j) + A3(zy, i, k + 1, j + 1));
    tmp1 = (hatavg_(g, s.v, g.msfvx, i, j, i, j + 1, k + 1, kte, cft1, cft2) - hatavg_(g, s.v, g.msfvx, i, j, i, j + 1, k, kte, cft1, cft2)) *
           tmpzy * A3(rdzw, i, k, j);
    tmp1 = mm * (g.rdy * (hat_(g, s.v, g.msfvx, i, k, j + 1) - hat_(g, s.v, g.msfvx, i, k, j)) - tmp1);
    A3(o.defor22, i, k, j) = 2.0f * tmp1;
    div = div + tmp1;
    tmp1 = (A3(w, i, k + 1, j) - A3(w, i, k, j)) * A3(rdzw, i, k, j);
    A3(o.defor33, i, k, j) = 2.0f * tmp1;
    div = div + tmp1;
    A3(o.div, i, k, j) = div;
  }
  // defor12 (:611-870)
  if (k <= ktf) {
    int ic = i, jc = j;
    if (r.cxs12 && i == g.ids) ic = g.ids + 1;
    if (r.cxe12 && i == g.ide) ic = g.ide - 1;
    if (r.cys12 && j == g.jds) jc = g.jds + 1;
    if (r.cye12 && j == g.jde) jc = g.jde - 1;
    if (ic >= r.i0_12 && ic <= r.i1_12 && jc >= r.j0_12 && jc <= r.j1_12) A3(o.defor12, i, k, j) = defor12_at(g, s, ic, k, jc, kte, cft1, cft2);
  }
  // defor13, defor23 (:872-1187)
  {
    int ic = i, jc = j;
    if (r.cxs && i == g.ids) ic = g.ids + 1;
    if (r.cxe && i == g.ide) ic = g.ide - 1;
    if (r.cys && j == g.jds) jc = g.jds + 1;
    if (r.cye && j == g.jde) jc = g.jde - 1;
    if (ic >= r.i0_13 && ic <= r.i1_13 && jc >= r.j0_13 && jc <= r.j1_13) A3(o.defor13, i, k, j) = defor13_at(g, s, ic, k, jc, kte, r.mix_full_fields);
    if (ic >= r.i0_23 && ic <= r.i1_23 && jc >= r.j0_23 && jc <= r.j1_23) A3(o.defor23, i, k, j) = defor23_at(g, s, ic, k, jc, kte, r.mix_full_fields);
  }
}

// ------------------------------------------------------------------------------------------ smag2d_km
struct SmagIn { cfp defor11, defor22, defor12, rdzw, zx, zy; float dx, dy, c_s; int diff_opt; };
// thread box: the computed range of smag2d_km (:1974-1989), k in 1..ktf
WG_HD void smag2d_point(const Geo &g, const SmagIn &s, int i, int k, int j, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  const float pr = 1.f / 3.f;   // prandtl, share/module_model_constants.F
  cfp defor11 = s.defor11, defor22 = s.defor22, defor12 = s.defor12, zx = s.zx, zy = s.zy;
  float def2 = 0.25f * ((A3(defor11, i, k, j) - A3(defor22, i, k, j)) * (A3(defor11, i, k, j) - A3(defor22, i, k, j)));
  float tmp = 0.25f * (A3(defor12, i, k, j) + A3(defor12, i, k, j + 1) + A3(defor12, i + 1, k, j) + A3(defor12, i + 1, k, j + 1));
  def2 = def2 + tmp * tmp;
  const float mlen_h = sqrtf(s.dx / A2(g.msftx, i, j) * s.dy / A2(g.msfty, i, j));
  tmp = sqrtf(def2);
  float km = s.c_s * s.c_s * mlen_h * mlen_h * tmp;
  km = fminf(km, 10.f * mlen_h);
  float kh = km / pr;
  if (s.diff_opt == 2) {
    const float dxm = s.dx / A2(g.msftx, i, j), dym = s.dy / A2(g.msfty, i, j);
    const float tmpzx = (0.25f * (fabsf(A3(zx, i, k, j)) + fabsf(A3(zx, i + 1, k, j)) + fabsf(A3(zx, i, k + 1, j)) + fabsf(A3(zx, i + 1, k + 1, j))) *
                         A3(s.rdzw, i, k, j) * dxm);
    const float tmpzy = (0.25f * (fabsf(A3(zy, i, k, j)) + fabsf(A3(zy, i, k, j + 1)) + fabsf(A3(zy, i, k + 1, j)) + fabsf(A3(zy, i, k + 1, j + 1))) *
                         A3(s.rdzw, i, k, j) * dym);
    const float alpha = fmaxf(sqrtf(tmpzx * tmpzx + tmpzy * tmpzy), 1.0f);
    const float def_limit = fmaxf(10.f / mlen_h, 1.e-3f);
    if (tmp > def_limit) km = km / (alpha * alpha);
    else km = km / (alpha);
    kh = km / pr;
  }
  A3(xkmh, i, k, j) = km;
  A3(xkmv, i, k, j) = km;
  A3(xkhh, i, k, j) = kh;
  A3(xkhv, i, k, j) = 0.f;
}

// ------------------------------------------------------------------------------------------ calculate_N2 + smag_km
// km_opt = 3 (module_diffusion_em.F:1485-1712, :1777-1928).  The reference reads th_phy, t_phy, p_phy, p8w, t8w of phy_prep
// (module_big_step_utilities_em.F:4818-4912); they are pure functions of the column's t_2, qv, p, pb, ph, phb, evaluated
// here at the three levels a point needs (one powf and one expf per level, the glibc restatements of libm_port.h).
#define WG_XLV 2.5e6f
#define WG_SVP1 0.6112f
#define WG_SVP2 17.67f
#define WG_SVP3 29.65f
#define WG_SVPT0 273.15f
#define WG_EP_2 (WG_RD / WG_RV)
struct Smag3In {
  cfp t_2, qv, qc, p, pb, ph, phb, rdz, rdzw, defor11, defor22, defor33, defor12, defor13, defor23;   // qv, qc may be null
  const float *dn;
  float dx, dy, dt, c_s, mix_upper_bound;
  int mix_isotropic, use_theta_m;
};
struct Thermo { float theta, t, p, qvs; };
WG_HD Thermo thermo_at(const Geo &g, const Smag3In &s, int i, int k, int j) {
  Thermo r;
  if (s.use_theta_m == 1 && s.qv) r.theta = (A3(s.t_2, i, k, j) + WG_T0) / (1.f + WG_RV / WG_RD * A3(s.qv, i, k, j));
  else r.theta = A3(s.t_2, i, k, j) + WG_T0;
  r.p = A3(s.p, i, k, j) + A3(s.pb, i, k, j);
  const float pi_phy = wg_libm::powf_(r.p / WG_P0, WG_RCP);
  r.t = r.theta * pi_phy;
  const float tc = r.t - WG_SVPT0;
  const float es = 1000.0f * WG_SVP1 * wg_libm::expf_(WG_SVP2 * tc / (r.t - WG_SVP3));
  r.qvs = WG_EP_2 * es / (r.p - es);
  return r;
}
// BN2(i,k,j) for k in 1..ktf-1 (the caller maps ktf to ktf-1, :1707-1711)
WG_HD float bn2_at(const Geo &g, const Smag3In &s, int i, int k, int j) {
  const float qc_cr = 0.00001f;
  cfp rdz = s.rdz, rdzw = s.rdzw;
  auto QV = [&](int kk) -> float { return s.qv ? A3(s.qv, i, kk, j) : 0.f; };
  auto TMP1 = [&](int kk) -> float {   // qv + qc (P_QI is not allocated on this path)
    float v = 0.f;
    if (s.qv) v = v + A3(s.qv, i, kk, j);
    if (s.qc) v = v + A3(s.qc, i, kk, j);
    return v;
  };
  const Thermo c = thermo_at(g, s, i, k, j), up = thermo_at(g, s, i, k + 1, j);
  const float qctmp = s.qc ? A3(s.qc, i, k, j) : 0.0f;
  const bool sat = QV(k) >= c.qvs || qctmp >= qc_cr;
  if (k >= 2) {
    const Thermo dn_ = thermo_at(g, s, i, k - 1, j);
    const float tmpdz = 1.0f / A3(rdz, i, k, j) + 1.0f / A3(rdz, i, k + 1, j);
    if (sat) {
      const float xlvqv = WG_XLV * QV(k);
      const float coefa = (1.0f + xlvqv / WG_RD / c.t) / (1.0f + WG_XLV * xlvqv / WG_CP / WG_RV / c.t / c.t) / c.theta;
      const float thetaep1 = up.theta * (1.0f + WG_XLV * up.qvs / WG_CP / up.t);
      const float thetaem1 = dn_.theta * (1.0f + WG_XLV * dn_.qvs / WG_CP / dn_.t);
      return WG_G * (coefa * (thetaep1 - thetaem1) / tmpdz - (TMP1(k + 1) - TMP1(k - 1)) / tmpdz);
    }
    return WG_G * ((up.theta - dn_.theta) / c.theta / tmpdz + 1.61f * (QV(k + 1) - QV(k - 1)) / tmpdz - (TMP1(k + 1) - TMP1(k - 1)) / tmpdz);
  }
  // k = 1: the surface values (:1663-1703); p8w, t8w at level 1 as phy_prep interpolates them (:4893-4903)
  float tmp1sfc = 0.f;
  if (s.qv) tmp1sfc = tmp1sfc + g.cf1 * A3(s.qv, i, 1, j) + g.cf2 * A3(s.qv, i, 2, j) + g.cf3 * A3(s.qv, i, 3, j);
  if (s.qc) tmp1sfc = tmp1sfc + g.cf1 * A3(s.qc, i, 1, j) + g.cf2 * A3(s.qc, i, 2, j) + g.cf3 * A3(s.qc, i, 3, j);
  const float zw1 = (A3(s.phb, i, 1, j) + A3(s.ph, i, 1, j)) / WG_G, zw2 = (A3(s.phb, i, 2, j) + A3(s.ph, i, 2, j)) / WG_G,
              zw3 = (A3(s.phb, i, 3, j) + A3(s.ph, i, 3, j)) / WG_G;
  const float z0 = zw1, z1 = 0.5f * (zw1 + zw2), z2 = 0.5f * (zw2 + zw3);
  const float w1 = (z0 - z2) / (z1 - z2);
  const float w2 = 1.f - w1;
  const float p8w = w1 * c.p + w2 * up.p, t8w = w1 * c.t + w2 * up.t;
  float tmpdz = 1.0f / A3(rdz, i, k + 1, j) + 0.5f / A3(rdzw, i, k, j);
  const float thetasfc = t8w / wg_libm::powf_(p8w / WG_P0, WG_RD / WG_CP);
  if (sat) {
    const Thermo l3 = thermo_at(g, s, i, 3, j);
    const float qvsfc = g.cf1 * c.qvs + g.cf2 * up.qvs + g.cf3 * l3.qvs;
    const float xlvqv = WG_XLV * QV(k);
    const float coefa = (1.0f + xlvqv / WG_RD / c.t) / (1.0f + WG_XLV * xlvqv / WG_CP / WG_RV / c.t / c.t) / c.theta;
    const float thetaep1 = up.theta * (1.0f + WG_XLV * up.qvs / WG_CP / up.t);
    const float thetaesfc = thetasfc * (1.0f + WG_XLV * qvsfc / WG_CP / t8w);
    return WG_G * (coefa * (thetaep1 - thetaesfc) / tmpdz - (TMP1(k + 1) - tmp1sfc) / tmpdz);
  }
  const float qvsfc = g.cf1 * QV(1) + g.cf2 * QV(2) + g.cf3 * QV(3);
  tmpdz = 1.f / A3(rdzw, i, k, j);
  return WG_G * ((up.theta - c.theta) / c.theta / tmpdz + 1.61f * (QV(k + 1) - qvsfc) / tmpdz - (TMP1(k + 1) - tmp1sfc) / tmpdz);
}
// thread box: the computed range of smag_km (:1816-1831), k in 1..ktf
WG_HD void smag3d_point(const Geo &g, const Smag3In &s, int i, int k, int j, fp bn2, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  const int ktf = g.kde - 1;
  const float pr = 1.f / 3.f;
  cfp defor11 = s.defor11, defor22 = s.defor22, defor33 = s.defor33, defor12 = s.defor12, defor13 = s.defor13, defor23 = s.defor23;
  const float BN2 = bn2_at(g, s, i, k == ktf ? ktf - 1 : k, j);
  A3(bn2, i, k, j) = BN2;
  float def2 = 0.5f * (A3(defor11, i, k, j) * A3(defor11, i, k, j) + A3(defor22, i, k, j) * A3(defor22, i, k, j) + A3(defor33, i, k, j) * A3(defor33, i, k, j));
  float tmp = 0.25f * (A3(defor12, i, k, j) + A3(defor12, i, k, j + 1) + A3(defor12, i + 1, k, j) + A3(defor12, i + 1, k, j + 1));
  def2 = def2 + tmp * tmp;
  tmp = 0.25f * (A3(defor13, i, k + 1, j) + A3(defor13, i, k, j) + A3(defor13, i + 1, k + 1, j) + A3(defor13, i + 1, k, j));
  def2 = def2 + tmp * tmp;
  tmp = 0.25f * (A3(defor23, i, k + 1, j) + A3(defor23, i, k, j) + A3(defor23, i, k + 1, j + 1) + A3(defor23, i, k, j + 1));
  def2 = def2 + tmp * tmp;
  tmp = fmaxf(0.f, def2 - BN2 / pr);
  tmp = wg_libm::powf_(tmp, 0.5f);   // tmp**0.5 (module_diffusion_em.F:1897, :1915) is powf for gfortran, and glibc's powf(x, .5) is not always sqrtf(x)
  const float c_s = s.c_s, mub = s.mix_upper_bound, dt = s.dt, dx = s.dx, dy = s.dy;
  const float rdzw = A3(s.rdzw, i, k, j), mtx = A2(g.msftx, i, j), mty = A2(g.msfty, i, j);
  float km_h, km_v, kh_h, kh_v;
  if (s.mix_isotropic == 0) {
    const float mlen_h = sqrtf(dx / mtx * dy / mty);
    const float mlen_v = 1.f / rdzw;
    km_h = fmaxf(c_s * c_s * mlen_h * mlen_h * tmp, 1.0E-6f * mlen_h * mlen_h);
    km_h = fminf(km_h, mub * mlen_h * mlen_h / dt);
    km_v = fmaxf(c_s * c_s * mlen_v * mlen_v * tmp, 1.0E-6f * mlen_v * mlen_v);
    km_v = fminf(km_v, mub * mlen_v * mlen_v / dt);
    kh_h = km_h / pr;
    kh_h = fminf(kh_h, mub * mlen_h * mlen_h / dt);
    kh_v = km_v / pr;
    kh_v = fminf(kh_v, mub * mlen_v * mlen_v / dt);
  } else {
    const float deltas = wg_libm::powf_(dx / mtx * dy / mty / rdzw, 0.33333333f);
    km_h = fmaxf(c_s * c_s * deltas * deltas * tmp, 1.0E-6f * deltas * deltas);
    km_h = fminf(km_h, mub * dx / mtx * dy / mty / dt);
    km_v = km_h;
    km_v = fminf(km_v, mub / rdzw / rdzw / dt);
    kh_h = km_h / pr;
    kh_h = fminf(kh_h, mub * dx / mtx * dy / mty / dt);
    kh_v = km_v / pr;
    kh_v = fminf(kh_v, mub / rdzw / rdzw / dt);
  }
  A3(xkmh, i, k, j) = km_h; A3(xkmv, i, k, j) = km_v; A3(xkhh, i, k, j) = kh_h; A3(xkhv, i, k, j) = kh_v;
}

// ------------------------------------------------------------------------------------------ tke_km, tke_rhs (km_opt = 2)
// module_diffusion_em.F:2049-2335 (tke_km), :2341-2405 (calc_l_scale), :6099-6228 (tke_rhs = tke_shear :6529 + tke_buoyancy
// :6234 + tke_dissip :6384 + the limiter of the sink)
struct TkeIn {
  Smag3In th;                       // the thermodynamic inputs of calculate_N2 / tke_km and the deformation tensor
  cfp tke, bn2, xkmh, xkmv, xkhv, u, v, mut, rho, hfx, ustm;
  float c_k, cd0, heat_flux;
  int isfflx, diff_opt;
};
WG_HD float theta_at(const Geo &g, const Smag3In &s, int i, int k, int j) {
  if (s.use_theta_m == 1 && s.qv) return (A3(s.t_2, i, k, j) + WG_T0) / (1.f + WG_RV / WG_RD * A3(s.qv, i, k, j));
  return A3(s.t_2, i, k, j) + WG_T0;
}
WG_HD float l_scale_at(const Geo &g, const TkeIn &s, int i, int k, int j, float deltas) {
  float l = deltas;
  const float bn2 = A3(s.bn2, i, k, j);
  if (bn2 > 1.0e-6f) {
    const float tmp = sqrtf(fmaxf(A3(s.tke, i, k, j), 1.0e-6f));
    l = 0.76f * tmp / sqrtf(bn2);
    l = fminf(l, deltas);
    l = fmaxf(l, 0.001f * deltas);
  }
  return l;
}
// thread box: the computed range of tke_km, k in 1..ktf; writes BN2 (calculate_N2) and the four K arrays
WG_HD void tke_km_point(const Geo &g, const TkeIn &s, int i, int k, int j, fp bn2, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  const Smag3In &t = s.th;
  const int ktf = g.kde - 1, kde = g.kde;
  const float prandtl = 1.f / 3.0f;
  cfp rdz = t.rdz, rdzw = t.rdzw;
  const float BN2 = bn2_at(g, t, i, k == ktf ? ktf - 1 : k, j);
  A3(bn2, i, k, j) = BN2;
  float tke_seed = 0.f;
  if (s.isfflx == 0) {
    if (s.diff_opt == 2) { if (s.cd0 < 1.e-10f && s.heat_flux < 1.e-10f) tke_seed = 1.e-06f; }
    else tke_seed = 1.e-06f;
  }
  const float mtx = A2(g.msftx, i, j), mty = A2(g.msfty, i, j), dx = t.dx, dy = t.dy, dt = t.dt, mub = t.mix_upper_bound;
  const float tmp = sqrtf(fmaxf(A3(s.tke, i, k, j), tke_seed));
  float km_h, km_v, kh_h, kh_v;
  if (t.mix_isotropic == 0) {
    float dthrdn;
    if (k == 1) {
      const Thermo c = thermo_at(g, t, i, 1, j), up = thermo_at(g, t, i, 2, j);
      const float zw1 = (A3(t.phb, i, 1, j) + A3(t.ph, i, 1, j)) / WG_G, zw2 = (A3(t.phb, i, 2, j) + A3(t.ph, i, 2, j)) / WG_G,
                  zw3 = (A3(t.phb, i, 3, j) + A3(t.ph, i, 3, j)) / WG_G;
      const float z0 = zw1, z1 = 0.5f * (zw1 + zw2), z2 = 0.5f * (zw2 + zw3);
      const float w1 = (z0 - z2) / (z1 - z2);
      const float w2 = 1.f - w1;
      const float p8w = w1 * c.p + w2 * up.p, t8w = w1 * c.t + w2 * up.t;
      const float tmpdz = 1.0f / A3(rdzw, i, 2, j) + 1.0f / A3(rdzw, i, 1, j);
      const float thetasfc = t8w / wg_libm::powf_(p8w / WG_P0, WG_RD / WG_CP);
      dthrdn = (up.theta - thetasfc) / tmpdz;
    } else if (k == ktf) {
      const Thermo a = thermo_at(g, t, i, kde - 1, j), b = thermo_at(g, t, i, kde - 2, j);
      const float zwe = (A3(t.phb, i, kde, j) + A3(t.ph, i, kde, j)) / WG_G, zwf = (A3(t.phb, i, ktf, j) + A3(t.ph, i, ktf, j)) / WG_G,
                  zwg = (A3(t.phb, i, ktf - 1, j) + A3(t.ph, i, ktf - 1, j)) / WG_G;
      const float z0 = zwe, z1 = 0.5f * (zwf + zwe), z2 = 0.5f * (zwg + zwf);
      const float w1 = (z0 - z2) / (z1 - z2);
      const float w2 = 1.f - w1;
      const float p8w = wg_libm::expf_(w1 * wg_libm::logf_(a.p) + w2 * wg_libm::logf_(b.p)), t8w = w1 * a.t + w2 * b.t;
      const float tmpdz = 1.0f / A3(rdz, i, k, j) + 0.5f / A3(rdzw, i, k, j);
      const float thetatop = t8w / wg_libm::powf_(p8w / WG_P0, WG_RD / WG_CP);
      dthrdn = (thetatop - theta_at(g, t, i, k - 1, j)) / tmpdz;
    } else {
      const float tmpdz = 1.0f / A3(rdz, i, k + 1, j) + 1.0f / A3(rdz, i, k, j);
      dthrdn = (theta_at(g, t, i, k + 1, j) - theta_at(g, t, i, k - 1, j)) / tmpdz;
    }
    const float mlen_h = sqrtf(dx / mtx * dy / mty);
    const float deltas = 1.0f / A3(rdzw, i, k, j);
    float mlen_v = deltas;
    if (dthrdn > 0.f) {
      const float mlen_s = 0.76f * tmp / wg_libm::powf_(fabsf(WG_G / theta_at(g, t, i, k, j) * dthrdn), 0.5f);   // (...)**0.5 (:2212)
      mlen_v = fminf(mlen_v, mlen_s);
    }
    km_h = fmaxf(s.c_k * tmp * mlen_h, 1.0E-6f * mlen_h * mlen_h);
    km_h = fminf(km_h, mub * mlen_h * mlen_h / dt);
    km_v = fmaxf(s.c_k * tmp * mlen_v, 1.0E-6f * deltas * deltas);
    km_v = fminf(km_v, mub * deltas * deltas / dt);
    const float pr_inv_h = 1.f / prandtl;
    const float pr_inv_v = 1.0f + 2.0f * mlen_v / deltas;
    kh_h = km_h * pr_inv_h;
    kh_v = km_v * pr_inv_v;
  } else {
    const float rz = A3(rdzw, i, k, j);
    const float deltas = wg_libm::powf_(dx / mtx * dy / mty / rz, 0.33333333f);
    // calc_l_scale reads BN2 and tke of this very point: use the value just computed (the bn2 field is being written)
    float l = deltas;
    if (BN2 > 1.0e-6f) {
      const float tt = sqrtf(fmaxf(A3(s.tke, i, k, j), 1.0e-6f));
      l = 0.76f * tt / sqrtf(BN2);
      l = fminf(l, deltas);
      l = fmaxf(l, 0.001f * deltas);
    }
    km_h = s.c_k * tmp * l;
    km_h = fminf(mub * dx / mtx * dy / mty / dt, km_h);
    km_v = s.c_k * tmp * l;
    km_v = fminf(mub / rz / rz / dt, km_v);
    const float pr_inv = 1.0f + 2.0f * l / deltas;
    kh_h = fminf(mub * dx / mtx * dy / mty / dt, km_h * pr_inv);
    kh_v = fminf(mub / rz / rz / dt, km_v * pr_inv);
  }
  A3(xkmh, i, k, j) = km_h; A3(xkmv, i, k, j) = km_v; A3(xkhh, i, k, j) = kh_h; A3(xkhv, i, k, j) = kh_v;
}
// thread box: the computed range of tke_rhs (the same as tke_km's), k in 1..ktf
WG_HD void tke_rhs_point(const Geo &g, const TkeIn &s, int i, int k, int j, fp tendency) {
  const Smag3In &t = s.th;
  const int ktf = g.kde - 1;
  cfp defor11 = t.defor11, defor22 = t.defor22, defor33 = t.defor33, defor12 = t.defor12, defor13 = t.defor13, defor23 = t.defor23, u = s.u, v = s.v;
  const float mass = g.c1h[k] * A2(s.mut, i, j) + g.c2h[k];
  const float xkmh = A3(s.xkmh, i, k, j), xkmv = A3(s.xkmv, i, k, j), xkhv = A3(s.xkhv, i, k, j);
  float td = A3(tendency, i, k, j);
  auto sq = [](float a) { return a * a; };
  // tke_shear
  td = td + 0.5f * mass * xkmh * (A3(defor11, i, k, j) * A3(defor11, i, k, j));
  td = td + 0.5f * mass * xkmh * (A3(defor22, i, k, j) * A3(defor22, i, k, j));
  td = td + 0.5f * mass * xkmv * (A3(defor33, i, k, j) * A3(defor33, i, k, j));
  float avg = 0.25f * (sq(A3(defor12, i, k, j)) + sq(A3(defor12, i, k, j + 1)) + sq(A3(defor12, i + 1, k, j)) + sq(A3(defor12, i + 1, k, j + 1)));
  td = td + mass * xkmh * avg;
  auto d13 = [&](int ii, int kk) -> float { return (kk == 1 || kk == ktf + 1) ? 0.0f : A3(defor13, ii, kk, j); };
  auto d23 = [&](int jj, int kk) -> float { return (kk == 1 || kk == ktf + 1) ? 0.0f : A3(defor23, i, kk, jj); };
  avg = 0.25f * (sq(d13(i, k + 1)) + sq(d13(i, k)) + sq(d13(i + 1, k + 1)) + sq(d13(i + 1, k)));
  td = td + mass * xkmv * avg;
  float absU = 0.f, Cd = 0.f;
  if (k == 1) {
    const float a = A3(u, i, k, j) + A3(u, i + 1, k, j), b = A3(v, i, k, j) + A3(v, i, k, j + 1);
    absU = 0.5f * sqrtf(a * a + b * b);
    if (s.isfflx == 0) Cd = s.cd0;
    else { absU = absU + 1.e-10f; Cd = (A2(s.ustm, i, j) * A2(s.ustm, i, j)) / (absU * absU); }
    td = td + mass * ((A3(u, i, k, j) + A3(u, i + 1, k, j)) * 0.5f * Cd * absU * (A3(defor13, i, 2, j) + A3(defor13, i + 1, 2, j)) * 0.5f);
  }
  avg = 0.25f * (sq(d23(j, k + 1)) + sq(d23(j, k)) + sq(d23(j + 1, k + 1)) + sq(d23(j + 1, k)));
  td = td + mass * xkmv * avg;
  if (k == 1) td = td + mass * ((A3(v, i, k, j) + A3(v, i, k, j + 1)) * 0.5f * Cd * absU * (A3(defor23, i, 2, j) + A3(defor23, i, 2, j + 1)) * 0.5f);
  // tke_buoyancy
  const float BN2 = A3(s.bn2, i, k, j);
  if (k >= 2) td = td - mass * xkhv * BN2;
  else {
    float heat_flux;
    if (s.isfflx == 0 || s.isfflx == 2) heat_flux = s.heat_flux;
    else {
      const float cpm = WG_CP * (1.f + 0.8f * (t.qv ? A3(t.qv, i, k, j) : 0.f));
      heat_flux = (A2(s.hfx, i, j) / cpm) / A3(s.rho, i, k, j);
    }
    td = td - mass * ((xkhv * BN2) - (WG_G / theta_at(g, t, i, k, j)) * heat_flux) / 2.f;
  }
  // tke_dissip
  {
    const float ce1 = (s.c_k / 0.10f) * 0.19f;
    const float ce2 = fmaxf(0.0f, 0.93f - ce1);
    const float deltas = wg_libm::powf_(t.dx / A2(g.msftx, i, j) * t.dy / A2(g.msfty, i, j) / A3(t.rdzw, i, k, j), 0.33333333f);
    const float l_scale = l_scale_at(g, s, i, k, j, deltas);
    const float tketmp = fmaxf(A3(s.tke, i, k, j), 1.0e-6f);
    float coefc;
    if (k == 1 || k == ktf) coefc = 3.9f;
    else coefc = ce1 + ce2 * l_scale / deltas;
    td = td - mass * coefc * wg_libm::powf_(tketmp, 1.5f) / l_scale;
  }
  td = fmaxf(td, -mass * fmaxf(0.0f, A3(s.tke, i, k, j)) / t.dt);
  A3(tendency, i, k, j) = td;
}

}  // namespace wg
