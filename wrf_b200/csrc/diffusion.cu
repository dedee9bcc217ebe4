// diffusion.cu -- constant-K diffusion on coordinate surfaces (diff_opt = 1, km_opt = 1):
// horizontal_diffusion (module_big_step_utilities_em.F:2715), horizontal_diffusion_3dmp (:2954),
// vertical_diffusion (:3064), vertical_diffusion_mp (:3192), vertical_diffusion_3dmp (:3286),
// vertical_diffusion_u (:3382), vertical_diffusion_v (:3485), called on the first RK substep from
// rk_tendency (module_em.F:802-878) and rk_scalar_tend (:1378-1419).
//
// One launch per tendency array: the thread of a point adds the horizontal term and then the vertical
// term, in the reference's order, each inside the reference's own loop ranges.  The reference's row
// buffer vflux(i,k) is a pure function of the column, so vflux(k) and vflux(k-1) are evaluated at the
// point (vflux(0) = vflux(1), vflux(ktf) = 0 as in the reference).
//
// The eddy viscosities xkmh / xkhh the reference reads from arrays are the constants khdif and
// khdif/prandtl on every element (isotropic_km, module_diffusion_em.F:1718-1772, then HALO_EM_TKE_E and
// phy_bc): the kernels take the number and keep the averaging expressions as written.
#include "common.cuh"

namespace wg {

enum { DIFF_U = 0, DIFF_V = 1, DIFF_W = 2, DIFF_THETA = 3, DIFF_QV = 4, DIFF_SCALAR = 5 };

struct DiffIn {
  cfp field, mut, mass_v, alt, base_3d;   // mut: horizontal mass; mass_v: muu / muv / mut of the vertical term
  const float *base_1d;                   // u_base / v_base / qv_base (origin shifted), or null
  float xk, kvdif;                        // horizontal K of this field, vertical K of this field
  cfp xkarr;                              // ARR: the horizontal K is this array (xkmh / xkhh of a closure) instead of xk
};

// ARR = false: every element of the reference's xkmhd array holds the constant xk (km_opt = 1), the averaging
// expressions are kept; ARR = true: the array computed by k_smag2d (km_opt = 4), indexed as in the reference
#define XK(i_, k_, j_) (ARR ? A3(s.xkarr, i_, k_, j_) : xk)
template <int KIND, bool ARR>
__global__ void k_diffuse(const Geo g, Box b, DiffIn s, fp tendency) {
  TID3;
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde, kde = g.kde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte, kts = 1;
  const int ktf = kde - 1;
  const float rdx = g.rdx, rdy = g.rdy, xk = s.xk, kvdif = s.kvdif;
  cfp field = s.field, MUT = s.mut, alt = s.alt;
  float t = A3(tendency, i, k, j);
  // ------------------------------------------------------------------ horizontal
  {
    int i_start, i_end, j_start, j_end, k_start = kts;
    if (KIND == DIFF_U) {
      i_start = its; i_end = ite; j_start = jts; j_end = min(jte, jde - 1);
      if (g.oxs()) i_start = max(ids + 1, its);
      if (g.oxe()) i_end = min(ide - 1, ite);
      if (g.oys()) j_start = max(jds + 1, jts);
      if (g.oye()) j_end = min(jde - 2, jte);
      if (g.periodic_x) { i_start = its; i_end = ite; }
    } else if (KIND == DIFF_V) {
      i_start = its; i_end = min(ite, ide - 1); j_start = jts; j_end = jte;
      if (g.oxs()) i_start = max(ids + 1, its);
      if (g.oxe()) i_end = min(ide - 2, ite);
      if (g.oys()) j_start = max(jds + 1, jts);
      if (g.oye()) j_end = min(jde - 1, jte);
      if (g.periodic_x) { i_start = its; i_end = min(ite, ide - 1); }
    } else {
      i_start = its; i_end = min(ite, ide - 1); j_start = jts; j_end = min(jte, jde - 1);
      if (g.oxs()) i_start = max(ids + 1, its);
      if (g.oxe()) i_end = min(ide - 2, ite);
      if (g.oys()) j_start = max(jds + 1, jts);
      if (g.oye()) j_end = min(jde - 2, jte);
      if (g.periodic_x) { i_start = its; i_end = min(ite, ide - 1); }
      if (KIND == DIFF_W) k_start = kts + 1;
    }
    if (i >= i_start && i <= i_end && j >= j_start && j <= j_end && k >= k_start && k <= ktf) {
      const float c1 = KIND == DIFF_W ? g.c1f[k] : g.c1h[k], c2 = KIND == DIFF_W ? g.c2f[k] : g.c2h[k];
      float mrdx, mkrdxm, mkrdxp, mrdy, mkrdym, mkrdyp;
      if (KIND == DIFF_U) {
        mkrdxm = (A2(g.msftx, i - 1, j) / A2(g.msfty, i - 1, j)) * (c1 * A2(MUT, i - 1, j) + c2) * XK(i - 1, k, j) * rdx;
        mkrdxp = (A2(g.msftx, i, j) / A2(g.msfty, i, j)) * (c1 * A2(MUT, i, j) + c2) * XK(i, k, j) * rdx;
        mrdx = A2(g.msfux, i, j) * A2(g.msfuy, i, j) * rdx;
        mkrdym = ((A2(g.msfuy, i, j) + A2(g.msfuy, i, j - 1)) / (A2(g.msfux, i, j) + A2(g.msfux, i, j - 1))) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2) + (c1 * A2(MUT, i - 1, j - 1) + c2) + (c1 * A2(MUT, i - 1, j) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i, k, j - 1) + XK(i - 1, k, j - 1) + XK(i - 1, k, j)) * rdy;
        mkrdyp = ((A2(g.msfuy, i, j) + A2(g.msfuy, i, j + 1)) / (A2(g.msfux, i, j) + A2(g.msfux, i, j + 1))) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j + 1) + c2) + (c1 * A2(MUT, i - 1, j + 1) + c2) + (c1 * A2(MUT, i - 1, j) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i, k, j + 1) + XK(i - 1, k, j + 1) + XK(i - 1, k, j)) * rdy;
        mrdy = A2(g.msfux, i, j) * A2(g.msfuy, i, j) * rdy;
      } else if (KIND == DIFF_V) {
        mkrdxm = ((A2(g.msfvx, i, j) + A2(g.msfvx, i - 1, j)) / (A2(g.msfvy, i, j) + A2(g.msfvy, i - 1, j))) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2) + (c1 * A2(MUT, i - 1, j - 1) + c2) + (c1 * A2(MUT, i - 1, j) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i, k, j - 1) + XK(i - 1, k, j - 1) + XK(i - 1, k, j)) * rdx;
        mkrdxp = ((A2(g.msfvx, i, j) + A2(g.msfvx, i + 1, j)) / (A2(g.msfvy, i, j) + A2(g.msfvy, i + 1, j))) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2) + (c1 * A2(MUT, i + 1, j - 1) + c2) + (c1 * A2(MUT, i + 1, j) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i, k, j - 1) + XK(i + 1, k, j - 1) + XK(i + 1, k, j)) * rdx;
        mrdx = A2(g.msfvx, i, j) * A2(g.msfvy, i, j) * rdx;
        mkrdym = (A2(g.msfty, i, j - 1) / A2(g.msftx, i, j - 1)) * XK(i, k, j - 1) * rdy;
        mkrdyp = (A2(g.msfty, i, j) / A2(g.msftx, i, j)) * XK(i, k, j) * rdy;
        mrdy = A2(g.msfvx, i, j) * A2(g.msfvy, i, j) * rdy;
      } else if (KIND == DIFF_W) {
        mkrdxm = (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i - 1, j) + c2) + (c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i - 1, j) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i - 1, k, j) + XK(i, k - 1, j) + XK(i - 1, k - 1, j)) * rdx;
        mkrdxp = (A2(g.msfux, i + 1, j) / A2(g.msfuy, i + 1, j)) *
                 0.25f * ((c1 * A2(MUT, i + 1, j) + c2) + (c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i + 1, j) + c2) + (c1 * A2(MUT, i, j) + c2)) *
                 0.25f * (XK(i + 1, k, j) + XK(i, k, j) + XK(i + 1, k - 1, j) + XK(i, k - 1, j)) * rdx;
        mrdx = A2(g.msftx, i, j) * A2(g.msfty, i, j) * rdx;
        mkrdym = (A2(g.msfvy, i, j) * A2(g.msfvx_inv, i, j)) *
                 0.25f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2) + (c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2)) *
                 0.25f * (XK(i, k, j) + XK(i, k, j - 1) + XK(i, k - 1, j) + XK(i, k - 1, j - 1)) * rdy;
        mkrdyp = (A2(g.msfvy, i, j + 1) * A2(g.msfvx_inv, i, j + 1)) *
                 0.25f * ((c1 * A2(MUT, i, j + 1) + c2) + (c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j + 1) + c2) + (c1 * A2(MUT, i, j) + c2)) *
                 0.25f * (XK(i, k, j + 1) + XK(i, k, j) + XK(i, k - 1, j + 1) + XK(i, k - 1, j)) * rdy;
        mrdy = A2(g.msftx, i, j) * A2(g.msfty, i, j) * rdy;
      } else {
        mkrdxm = (A2(g.msfux, i, j) / A2(g.msfuy, i, j)) * 0.5f * (XK(i, k, j) + XK(i - 1, k, j)) * 0.5f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i - 1, j) + c2)) * rdx;
        mkrdxp = (A2(g.msfux, i + 1, j) / A2(g.msfuy, i + 1, j)) * 0.5f * (XK(i + 1, k, j) + XK(i, k, j)) * 0.5f * ((c1 * A2(MUT, i + 1, j) + c2) + (c1 * A2(MUT, i, j) + c2)) * rdx;
        mrdx = A2(g.msftx, i, j) * A2(g.msfty, i, j) * rdx;
        mkrdym = (A2(g.msfvy, i, j) * A2(g.msfvx_inv, i, j)) * 0.5f * (XK(i, k, j) + XK(i, k, j - 1)) * 0.5f * ((c1 * A2(MUT, i, j) + c2) + (c1 * A2(MUT, i, j - 1) + c2)) * rdy;
        mkrdyp = (A2(g.msfvy, i, j + 1) * A2(g.msfvx_inv, i, j + 1)) * 0.5f * (XK(i, k, j + 1) + XK(i, k, j)) * 0.5f * ((c1 * A2(MUT, i, j + 1) + c2) + (c1 * A2(MUT, i, j) + c2)) * rdy;
        mrdy = A2(g.msftx, i, j) * A2(g.msfty, i, j) * rdy;
      }
      const float fc = A3(field, i, k, j), fe = A3(field, i + 1, k, j), fw = A3(field, i - 1, k, j), fn = A3(field, i, k, j + 1), fs = A3(field, i, k, j - 1);
      if (KIND == DIFF_THETA) {
        cfp bs = s.base_3d;
        const float bc = A3(bs, i, k, j), be = A3(bs, i + 1, k, j), bw = A3(bs, i - 1, k, j), bn = A3(bs, i, k, j + 1), bso = A3(bs, i, k, j - 1);
        t = t + (mrdx * (mkrdxp * (fe - fc - be + bc) - mkrdxm * (fc - fw - bc + bw)) + mrdy * (mkrdyp * (fn - fc - bn + bc) - mkrdym * (fc - fs - bc + bso)));
      } else {
        t = t + (mrdx * (mkrdxp * (fe - fc) - mkrdxm * (fc - fw)) + mrdy * (mkrdyp * (fn - fc) - mkrdym * (fc - fs)));
      }
    }
  }
  // ------------------------------------------------------------------ vertical (bl_pbl_physics = 0)
  {
    int i_start = its, i_end = min(ite, ide - 1), j_start = jts, j_end = min(jte, jde - 1);
    int k_lo = kts, k_hi = ktf;
    if (KIND == DIFF_U) {
      i_end = ite;
      if (g.oxs()) i_start = max(ids + 1, its);
      if (g.oxe()) i_end = min(ide - 1, ite);
      if (g.periodic_x) { i_start = its; i_end = ite; }
      k_hi = ktf - 1;
    } else if (KIND == DIFF_V) {
      if (g.oys()) j_start = max(jds + 1, jts);
      if (g.oye()) j_end = min(jde - 1, jte);
      k_hi = ktf - 1;
    } else if (KIND == DIFF_W) {
      k_lo = kts + 1;
    }
    if (i >= i_start && i <= i_end && j >= j_start && j <= j_end && k >= k_lo && k <= k_hi) {
      // vflux(kk), kk = kts..ktf-1, of this column; vflux(ktf) = 0; vflux(0) = vflux(1) (not for w)
      auto vflux = [&](int kk) -> float {
        if (kk == ktf) return 0.f;
        if (kk == 0) kk = 1;
        if (KIND == DIFF_W) return (kvdif / A3(alt, i, kk, j)) * g.rdnw[kk] * (A3(field, i, kk + 1, j) - A3(field, i, kk, j));
        if (KIND == DIFF_U)
          return kvdif * g.rdn[kk + 1] / (0.25f * (A3(alt, i, kk, j) + A3(alt, i - 1, kk, j) + A3(alt, i, kk + 1, j) + A3(alt, i - 1, kk + 1, j))) *
                 (A3(field, i, kk + 1, j) - A3(field, i, kk, j) - s.base_1d[kk + 1] + s.base_1d[kk]);
        if (KIND == DIFF_V)
          return kvdif * g.rdn[kk + 1] / (0.25f * (A3(alt, i, kk, j) + A3(alt, i, kk, j - 1) + A3(alt, i, kk + 1, j) + A3(alt, i, kk + 1, j - 1))) *
                 (A3(field, i, kk + 1, j) - A3(field, i, kk, j) - s.base_1d[kk + 1] + s.base_1d[kk]);
        if (KIND == DIFF_THETA)
          return kvdif * g.rdn[kk + 1] / (0.5f * (A3(alt, i, kk, j) + A3(alt, i, kk + 1, j))) *
                 (A3(field, i, kk + 1, j) - A3(field, i, kk, j) - A3(s.base_3d, i, kk + 1, j) + A3(s.base_3d, i, kk, j));
        if (KIND == DIFF_QV)
          return kvdif * g.rdn[kk + 1] / (0.5f * (A3(alt, i, kk, j) + A3(alt, i, kk + 1, j))) *
                 (A3(field, i, kk + 1, j) - A3(field, i, kk, j) - s.base_1d[kk + 1] + s.base_1d[kk]);
        return kvdif * g.rdn[kk + 1] / (0.5f * (A3(alt, i, kk, j) + A3(alt, i, kk + 1, j))) * (A3(field, i, kk + 1, j) - A3(field, i, kk, j));
      };
      const float vk = vflux(k), vkm1 = vflux(k - 1);
      cfp mv = s.mass_v;
      if (KIND == DIFF_W)
        t = t + g.rdn[k] * WG_G * WG_G / (g.c1f[k] * A2(mv, i, j) + g.c2f[k]) / (0.5f * (A3(alt, i, k, j) + A3(alt, i, k - 1, j))) * (vk - vkm1);
      else if (KIND == DIFF_U)
        t = t + WG_G * WG_G * g.rdnw[k] / (g.c1h[k] * A2(mv, i, j) + g.c2h[k]) / (0.5f * (A3(alt, i - 1, k, j) + A3(alt, i, k, j))) * (vk - vkm1);
      else if (KIND == DIFF_V)
        t = t + WG_G * WG_G * g.rdnw[k] / (g.c1h[k] * A2(mv, i, j) + g.c2h[k]) / (0.5f * (A3(alt, i, k, j - 1) + A3(alt, i, k, j))) * (vk - vkm1);
      else
        t = t + WG_G * WG_G / (g.c1h[k] * A2(mv, i, j) + g.c2h[k]) / A3(alt, i, k, j) * g.rdnw[k] * (vk - vkm1);
    }
  }
  A3(tendency, i, k, j) = t;
}

// which = DIFF_*; adds the diffusion of `field` to `tendency` on the whole tile
void launch_diffuse(Solver *s, int which, cfp field, fp tendency, float xk, float kvdif, cfp xkarr) {
  const Geo &q = s->g;
  Box tile{q.its, q.ite, 1, q.kde, q.jts, q.jte};
  DiffIn in;
  in.field = field; in.mut = s->o("mut"); in.mass_v = s->o("mut"); in.alt = s->o("alt"); in.base_3d = nullptr; in.base_1d = nullptr;
  in.xk = xk; in.kvdif = kvdif; in.xkarr = xkarr;
  switch (which) {
    case DIFF_U: in.mass_v = s->o("muu"); in.base_1d = s->o("u_base"); if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_U, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_U, false>), tile, in, tendency); break;
    case DIFF_V: in.mass_v = s->o("muv"); in.base_1d = s->o("v_base"); if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_V, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_V, false>), tile, in, tendency); break;
    case DIFF_W: if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_W, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_W, false>), tile, in, tendency); break;
    case DIFF_THETA: in.base_3d = s->o("t_init"); if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_THETA, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_THETA, false>), tile, in, tendency); break;
    case DIFF_QV:     // rk_scalar_tend passes grid%muts as mut (solve_em.F:2219)
      in.mut = s->o("muts"); in.mass_v = s->o("muts"); in.base_1d = s->o("qv_base");
      if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_QV, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_QV, false>), tile, in, tendency); break;
    default:
      in.mut = s->o("muts"); in.mass_v = s->o("muts");
      if (xkarr) LAUNCH3(s, (k_diffuse<DIFF_SCALAR, true>), tile, in, tendency); else LAUNCH3(s, (k_diffuse<DIFF_SCALAR, false>), tile, in, tendency); break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// sixth_order_diffusion (module_big_step_utilities_em.F:6220-6620): the 6th-order horizontal filter of diff_6th_opt = 1, 2,
// added to the *_tendf arrays on the first RK substep (module_em.F:880-930, 1421-1437).  KIND: DIFF_U, DIFF_V, DIFF_W, else mass.
struct Diff6In {
  cfp field, mut, phb;
  float coef;        // diff_6th_factor * 0.015625 / (2 dt)
  float thresh;      // diff_6th_thresh
  int opt, slopeopt;
  Box r;             // the reference's loop ranges for this variable
};

template <int KIND>
__global__ void k_diff6(const Geo g, Box b, Diff6In s, fp tendency) {
  TID3;
  cfp field = s.field, MUT = s.mut, phb = s.phb;
  const float *c1 = KIND == DIFF_W ? g.c1f : g.c1h, *c2 = KIND == DIFF_W ? g.c2f : g.c2h;
  float dflux_x_p0 = (10.0f * (A3(field, i, k, j) - A3(field, i - 1, k, j)) - 5.0f * (A3(field, i + 1, k, j) - A3(field, i - 2, k, j)) + (A3(field, i + 2, k, j) - A3(field, i - 3, k, j)));
  float dflux_x_p1 = (10.0f * (A3(field, i + 1, k, j) - A3(field, i, k, j)) - 5.0f * (A3(field, i + 2, k, j) - A3(field, i - 1, k, j)) + (A3(field, i + 3, k, j) - A3(field, i - 2, k, j)));
  if (s.opt == 2) {
    if (dflux_x_p0 * (A3(field, i, k, j) - A3(field, i - 1, k, j)) <= 0.0f) dflux_x_p0 = 0.0f;
    if (dflux_x_p1 * (A3(field, i + 1, k, j) - A3(field, i, k, j)) <= 0.0f) dflux_x_p1 = 0.0f;
  }
  float slopedamp_p0 = 1.f, slopedamp_p1 = 1.f;
  if (s.slopeopt >= 1) {
    const float dx = 1.f / g.rdx;
    const float dzthresh = s.thresh * 9.81f * dx;
    float dzmax_p0, dzmax_p1;
    if (KIND == DIFF_U) {
      dzmax_p0 = fmaxf(fabsf(A3(phb, i, k, j) - A3(phb, i - 1, k, j)) * A2(g.msfux, i, j), fabsf(A3(phb, i - 1, k, j) - A3(phb, i - 2, k, j)) * A2(g.msfux, i - 1, j));
      dzmax_p1 = fmaxf(fabsf(A3(phb, i + 1, k, j) - A3(phb, i, k, j)) * A2(g.msfux, i + 1, j), fabsf(A3(phb, i, k, j) - A3(phb, i - 1, k, j)) * A2(g.msfux, i, j));
    } else if (KIND == DIFF_V) {
      dzmax_p0 = fmaxf(fabsf(A3(phb, i, k, j) - A3(phb, i - 1, k, j)) * A2(g.msfux, i, j), fabsf(A3(phb, i, k, j - 1) - A3(phb, i - 1, k, j - 1)) * A2(g.msfux, i, j - 1));
      dzmax_p1 = fmaxf(fabsf(A3(phb, i + 1, k, j) - A3(phb, i, k, j)) * A2(g.msfux, i + 1, j), fabsf(A3(phb, i + 1, k, j - 1) - A3(phb, i, k, j - 1)) * A2(g.msfux, i + 1, j - 1));
    } else {
      dzmax_p0 = fabsf(A3(phb, i, k, j) - A3(phb, i - 1, k, j)) * A2(g.msfux, i, j);
      dzmax_p1 = fabsf(A3(phb, i + 1, k, j) - A3(phb, i, k, j)) * A2(g.msfux, i + 1, j);
    }
    slopedamp_p0 = fmaxf(1.f - (dzmax_p0 / dzthresh), 0.f);
    slopedamp_p1 = fmaxf(1.f - (dzmax_p1 / dzthresh), 0.f);
  }
  float mu_avg_p0, mu_avg_p1, tendency_x, tendency_y;
  if (KIND == DIFF_U) {
    mu_avg_p0 = (c1[k] * A2(MUT, i - 1, j) + c2[k]);
    mu_avg_p1 = (c1[k] * A2(MUT, i, j) + c2[k]);
    tendency_x = s.coef * A2(g.msfux, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
  } else if (KIND == DIFF_V) {
    mu_avg_p0 = 0.25f * ((c1[k] * A2(MUT, i - 1, j - 1) + c2[k]) + (c1[k] * A2(MUT, i, j - 1) + c2[k]) + (c1[k] * A2(MUT, i - 1, j) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]));
    mu_avg_p1 = 0.25f * ((c1[k] * A2(MUT, i, j - 1) + c2[k]) + (c1[k] * A2(MUT, i + 1, j - 1) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]) + (c1[k] * A2(MUT, i + 1, j) + c2[k]));
    tendency_x = s.coef * A2(g.msfvx, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
  } else {
    mu_avg_p0 = 0.5f * ((c1[k] * A2(MUT, i - 1, j) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]));
    mu_avg_p1 = 0.5f * ((c1[k] * A2(MUT, i, j) + c2[k]) + (c1[k] * A2(MUT, i + 1, j) + c2[k]));
    tendency_x = s.coef * A2(g.msftx, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_x_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_x_p0));
  }
  float dflux_y_p0 = (10.0f * (A3(field, i, k, j) - A3(field, i, k, j - 1)) - 5.0f * (A3(field, i, k, j + 1) - A3(field, i, k, j - 2)) + (A3(field, i, k, j + 2) - A3(field, i, k, j - 3)));
  float dflux_y_p1 = (10.0f * (A3(field, i, k, j + 1) - A3(field, i, k, j)) - 5.0f * (A3(field, i, k, j + 2) - A3(field, i, k, j - 1)) + (A3(field, i, k, j + 3) - A3(field, i, k, j - 2)));
  if (s.opt == 2) {
    if (dflux_y_p0 * (A3(field, i, k, j) - A3(field, i, k, j - 1)) <= 0.0f) dflux_y_p0 = 0.0f;
    if (dflux_y_p1 * (A3(field, i, k, j + 1) - A3(field, i, k, j)) <= 0.0f) dflux_y_p1 = 0.0f;
  }
  slopedamp_p0 = 1.f; slopedamp_p1 = 1.f;
  if (s.slopeopt >= 1) {
    const float dy = 1.f / g.rdy;
    const float dzthresh = s.thresh * 9.81f * dy;
    float dzmax_p0, dzmax_p1;
    if (KIND == DIFF_U) {
      dzmax_p0 = fmaxf(fabsf(A3(phb, i, k, j) - A3(phb, i, k, j - 1)) * A2(g.msfvy, i, j), fabsf(A3(phb, i - 1, k, j) - A3(phb, i - 1, k, j - 1)) * A2(g.msfvy, i - 1, j));
      dzmax_p1 = fmaxf(fabsf(A3(phb, i, k, j + 1) - A3(phb, i, k, j)) * A2(g.msfvy, i, j + 1), fabsf(A3(phb, i - 1, k, j + 1) - A3(phb, i - 1, k, j)) * A2(g.msfvy, i - 1, j + 1));
    } else if (KIND == DIFF_V) {
      dzmax_p0 = fmaxf(fabsf(A3(phb, i, k, j) - A3(phb, i, k, j - 1)) * A2(g.msfvy, i, j), fabsf(A3(phb, i, k, j - 1) - A3(phb, i, k, j - 2)) * A2(g.msfvy, i, j - 1));
      dzmax_p1 = fmaxf(fabsf(A3(phb, i, k, j + 1) - A3(phb, i, k, j)) * A2(g.msfvy, i, j + 1), fabsf(A3(phb, i, k, j) - A3(phb, i, k, j - 1)) * A2(g.msfvy, i, j));
    } else {
      dzmax_p0 = fabsf(A3(phb, i, k, j) - A3(phb, i, k, j - 1)) * A2(g.msfvy, i, j);
      dzmax_p1 = fabsf(A3(phb, i, k, j + 1) - A3(phb, i, k, j)) * A2(g.msfvy, i, j + 1);
    }
    slopedamp_p0 = fmaxf(1.f - (dzmax_p0 / dzthresh), 0.f);
    slopedamp_p1 = fmaxf(1.f - (dzmax_p1 / dzthresh), 0.f);
  }
  if (KIND == DIFF_U) {
    mu_avg_p0 = 0.25f * ((c1[k] * A2(MUT, i - 1, j - 1) + c2[k]) + (c1[k] * A2(MUT, i, j - 1) + c2[k]) + (c1[k] * A2(MUT, i - 1, j) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]));
    mu_avg_p1 = 0.25f * ((c1[k] * A2(MUT, i - 1, j) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]) + (c1[k] * A2(MUT, i - 1, j + 1) + c2[k]) + (c1[k] * A2(MUT, i, j + 1) + c2[k]));
    tendency_y = s.coef * A2(g.msfuy, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
  } else if (KIND == DIFF_V) {
    mu_avg_p0 = (c1[k] * A2(MUT, i, j - 1) + c2[k]);
    mu_avg_p1 = (c1[k] * A2(MUT, i, j) + c2[k]);
    tendency_y = s.coef * A2(g.msfvy, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
  } else {
    mu_avg_p0 = 0.5f * ((c1[k] * A2(MUT, i, j - 1) + c2[k]) + (c1[k] * A2(MUT, i, j) + c2[k]));
    mu_avg_p1 = 0.5f * ((c1[k] * A2(MUT, i, j) + c2[k]) + (c1[k] * A2(MUT, i, j + 1) + c2[k]));
    tendency_y = s.coef * A2(g.msfty, i, j) * ((slopedamp_p1 * mu_avg_p1 * dflux_y_p1) - (slopedamp_p0 * mu_avg_p0 * dflux_y_p0));
  }
  A3(tendency, i, k, j) = A3(tendency, i, k, j) + tendency_x + tendency_y;
}

// which = DIFF_U / DIFF_V / DIFF_W / anything else (mass points); adds the filter of `field` to `tendency`
void launch_diff6(Solver *s, int which, cfp field, fp tendency, cfp mut, float dt) {
  const Geo &q = s->g;
  const wrfgpu_config &c = s->c;
  const int ktf = q.kde - 1;
  Box r{q.its, 0, 1, ktf, q.jts, 0};
  if (which == DIFF_U) {
    r.i1 = q.ite; r.j1 = std::min(q.jde - 1, q.jte);
    if (c.open_xs) r.i0 = std::max(q.its, q.ids + 3);
    if (c.open_ys) r.j0 = std::max(q.jts, q.jds + 3);
    if (c.open_xe) r.i1 = std::min(q.ite, q.ide - 3);
    if (c.open_ye) r.j1 = std::min(q.jte, q.jde - 4);
    if (q.spec) {   // :6354
      r.i0 = std::max(q.its, q.ids + 3); r.i1 = std::min(q.ite, q.ide - 3);
      r.j0 = std::max(q.jts, q.jds + 3); r.j1 = std::min(q.jte, q.jde - 4);
    }
  } else if (which == DIFF_V) {
    r.i1 = std::min(q.ide - 1, q.ite); r.j1 = q.jte;
    if (c.open_xs) r.i0 = std::max(q.its, q.ids + 3);
    if (c.open_ys) r.j0 = std::max(q.jts, q.jds + 3);
    if (c.open_xe) r.i1 = std::min(q.ite, q.ide - 4);
    if (c.open_ye) r.j1 = std::min(q.jte, q.jde - 3);
    if (c.open_xs || q.spec) {   // sic (:6381): "open_xs .or. specified" restricts all four sides
      r.i0 = std::max(q.its, q.ids + 3); r.i1 = std::min(q.ite, q.ide - 4);
      r.j0 = std::max(q.jts, q.jds + 3); r.j1 = std::min(q.jte, q.jde - 3);
    }
  } else {
    r.i1 = std::min(q.ide - 1, q.ite); r.j1 = std::min(q.jde - 1, q.jte);
    if (c.open_xs) r.i0 = std::max(q.its, q.ids + 3);
    if (c.open_ys) r.j0 = std::max(q.jts, q.jds + 3);
    if (c.open_xe) r.i1 = std::min(q.ite, q.ide - 4);
    if (c.open_ye) r.j1 = std::min(q.jte, q.jde - 4);
    if (q.spec) {   // :6408, :6435
      r.i0 = std::max(q.its, q.ids + 3); r.i1 = std::min(q.ide - 4, q.ite);
      r.j0 = std::max(q.jts, q.jds + 3); r.j1 = std::min(q.jde - 4, q.jte);
    }
    if (which == DIFF_W) r.k0 = 2;
  }
  Diff6In in;
  in.field = field; in.mut = mut; in.phb = s->o("phb");
  in.coef = c.diff_6th_factor * 0.015625f / (2.0f * dt);
  in.thresh = c.diff_6th_thresh; in.opt = c.diff_6th_opt; in.slopeopt = c.diff_6th_slopeopt;
  in.r = r;
  if (which == DIFF_U) LAUNCH3(s, k_diff6<DIFF_U>, r, in, tendency);
  else if (which == DIFF_V) LAUNCH3(s, k_diff6<DIFF_V>, r, in, tendency);
  else if (which == DIFF_W) LAUNCH3(s, k_diff6<DIFF_W>, r, in, tendency);
  else LAUNCH3(s, k_diff6<DIFF_SCALAR>, r, in, tendency);
}

}  // namespace wg
