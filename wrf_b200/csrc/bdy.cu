// bdy.cu -- the specified / nested lateral boundary zone: share/module_bc.F:1122-2456 (relax_bdytend,
// spec_bdytend, spec_bdyupdate, spec_bdy_final, zero_grad_bdy, flow_dep_bdy) and dyn_em/module_bc_em.F:17-700
// (spec_bdyupdate_ph, relax_bdy_dry, relax_bdy_scalar, spec_bdy_dry, spec_bdy_scalar), called from solve_em.F:937-1007,
// :1344-1365, :1459-1490, :1582-1620, :2261-2355, :2446-2456, :4540-4722.
//
// Every routine of the reference walks four strips: Y-start and Y-end (they own the corners: the i range shrinks by the
// distance b_dist from the edge on both sides unless periodic_x), then X-start and X-end (j strictly inside the Y strips
// of the same distance; none when periodic_x), over the rows b_lo <= b_dist < b_hi of a field whose last index is
// ide / jde (u / v) or ide-1 / jde-1.  Here a point decides for itself which strip it is in (zone_b); a phase is one launch
// per side over the rows that side can hold, with every field of the phase in the same kernel.  The relaxation
// is a 5-point operator on (boundary value - mass-weighted field): the mass weighting the reference stores in a scratch
// array (mass_weight, module_bc_em.F:1716) is evaluated at the point of use.
#include "common.cuh"

namespace wg {

namespace {
enum Side { YS = 0, YE = 1, XS = 2, XE = 3 };

struct Ext { int ibe, jbe, itf, jtf; };
__host__ __device__ inline Ext ext_of(const Geo &g, char var) {
  Ext e{g.ide - 1, g.jde - 1, min(g.ite, g.ide - 1), min(g.jte, g.jde - 1)};
  if (var == 'u') { e.ibe = g.ide; e.itf = min(g.ite, g.ide); }
  if (var == 'v') { e.jbe = g.jde; e.jtf = min(g.jte, g.jde); }
  return e;
}

// b_dist of (i, j) if it lies in the rows [b_lo, b_hi) of strip `side` of a field with extents e, else -1
__device__ __forceinline__ int zone_b(const Geo &g, const Ext &e, int side, int b_lo, int b_hi, int i, int j) {
  if (i < g.its || i > e.itf || j < g.jts || j > e.jtf) return -1;
  const int ibs = g.ids, jbs = g.jds;
  const int b = side == YS ? j - jbs : (side == YE ? e.jbe - j : (side == XS ? i - ibs : e.ibe - i));
  if (b < b_lo || b >= b_hi) return -1;
  if (side < 2) {
    const int bl = g.periodic_x ? 0 : b;
    if (i < ibs + bl || i > e.ibe - bl) return -1;
  } else {
    if (g.periodic_x) return -1;
    if (j < jbs + b + 1 || j > e.jbe - b - 1) return -1;
  }
  return b;
}

// boundary arrays of one field: x sides (jms:jme, kds:kde, spec_bdy_width), y sides (ims:ime, kds:kde, spec_bdy_width),
// contiguous with the host's extents; b is the zero-based row (b_dist)
__device__ __forceinline__ long bat(const Geo &g, const BdyArr &B, int side, int along, int k, int b) {
  return side < 2 ? (long)(along - g.ims) + (long)(k - 1) * g.hni + (long)b * g.hni * B.nk
                  : (long)(along - g.jms) + (long)(k - 1) * g.hnj + (long)b * g.hnj * B.nk;
}
__device__ __forceinline__ float bvalue(const Geo &g, const BdyArr &B, int side, int along, int k, int b, float dtbc) {
  const long n = bat(g, B, side, along, k, b);
  return B.v[side][n] + dtbc * B.t[side][n];
}

// relax_bdytend_core (module_bc.F:1221-1427) at one point: a = fcx * fls0, c = gcx * (fls1 + fls2 + fls3 + fls4 - 4 fls0);
// the caller forms (field_tend + a) - c
template <class Fld>
__device__ __forceinline__ void relax_terms(const Geo &g, const BdyArr &B, const Ext &e, int side, int b, int i, int k, int j, float dtbc,
                                            const float *fcx, const float *gcx, Fld field, float &a, float &c) {
  float fls0, fls1, fls2, fls3, fls4;
  if (side < 2) {
    const int im1 = max(i - 1, g.ids), ip1 = min(i + 1, e.ibe);
    const int jo = side == YS ? j - 1 : j + 1, ji = side == YS ? j + 1 : j - 1;
    fls0 = bvalue(g, B, side, i, k, b, dtbc) - field(i, k, j);
    fls1 = bvalue(g, B, side, im1, k, b, dtbc) - field(im1, k, j);
    fls2 = bvalue(g, B, side, ip1, k, b, dtbc) - field(ip1, k, j);
    fls3 = bvalue(g, B, side, i, k, b - 1, dtbc) - field(i, k, jo);
    fls4 = bvalue(g, B, side, i, k, b + 1, dtbc) - field(i, k, ji);
  } else {
    const int io = side == XS ? i - 1 : i + 1, ii = side == XS ? i + 1 : i - 1;
    fls0 = bvalue(g, B, side, j, k, b, dtbc) - field(i, k, j);
    fls1 = bvalue(g, B, side, j - 1, k, b, dtbc) - field(i, k, j - 1);
    fls2 = bvalue(g, B, side, j + 1, k, b, dtbc) - field(i, k, j + 1);
    fls3 = bvalue(g, B, side, j, k, b - 1, dtbc) - field(io, k, j);
    fls4 = bvalue(g, B, side, j, k, b + 1, dtbc) - field(ii, k, j);
  }
  a = fcx[b] * fls0;
  c = gcx[b] * (fls1 + fls2 + fls3 + fls4 - 4.f * fls0);
}

struct RelaxDry {
  cfp ru, rv, ph, t, w, mu, mut;
  fp u_save, v_save, ph_save, t_save, w_save, mu_tend;
  fp ru_tendf, rv_tendf, ph_tendf, t_tendf, rw_tendf;
  BdyArr bu, bv, bph, bt, bw, bmu;
  const float *fcx, *gcx;
  float dtbc;
  int nested, store_save;
};

// relax_bdy_dry (module_bc_em.F:161-346) into the *_save arrays rk_tendency has zeroed, and the first-substep part of
// rk_addtend_dry that folds them into *_tendf (module_em.F:1037-1074).  store_save = 0 (inside a whole step): the *_save
// values themselves are not kept, small_step_prep overwrites them before anything reads them.
__global__ void k_relax_dry(const Geo g, Box b, int side, RelaxDry r) {
  TID3;
  const int sz = g.spec_zone, rz = g.relax_zone, ktf = g.kde - 1;
  float a, c;
  if (k <= ktf) {
    const Ext e = ext_of(g, 'u');
    const int bd = zone_b(g, e, side, sz, rz, i, j);
    if (bd >= 0) {
      relax_terms(g, r.bu, e, side, bd, i, k, j, r.dtbc, r.fcx, r.gcx, [&](int ii, int kk, int jj) { return A3(r.ru, ii, kk, jj); }, a, c);
      const float sv = 0.f + a - c;
      if (r.store_save) A3(r.u_save, i, k, j) = sv;
      A3(r.ru_tendf, i, k, j) = A3(r.ru_tendf, i, k, j) + sv * A2(g.msfuy, i, j);
    }
  }
  if (k <= ktf) {
    const Ext e = ext_of(g, 'v');
    const int bd = zone_b(g, e, side, sz, rz, i, j);
    if (bd >= 0) {
      relax_terms(g, r.bv, e, side, bd, i, k, j, r.dtbc, r.fcx, r.gcx, [&](int ii, int kk, int jj) { return A3(r.rv, ii, kk, jj); }, a, c);
      const float sv = 0.f + a - c;
      if (r.store_save) A3(r.v_save, i, k, j) = sv;
      A3(r.rv_tendf, i, k, j) = A3(r.rv_tendf, i, k, j) + sv * A2(g.msfvx, i, j);
    }
  }
  const Ext e = ext_of(g, 't');
  const int bd = zone_b(g, e, side, sz, rz, i, j);
  if (bd < 0) return;
  {   // ph: mass_weight with c1f, c2f; 'h' (all kde levels)
    relax_terms(g, r.bph, e, side, bd, i, k, j, r.dtbc, r.fcx, r.gcx,
                [&](int ii, int kk, int jj) { return A3(r.ph, ii, kk, jj) * (g.c1f[kk] * A2(r.mut, ii, jj) + g.c2f[kk]); }, a, c);
    const float sv = 0.f + a - c;
    if (r.store_save) A3(r.ph_save, i, k, j) = sv;
    A3(r.ph_tendf, i, k, j) = A3(r.ph_tendf, i, k, j) + sv;
  }
  if (k <= ktf) {
    relax_terms(g, r.bt, e, side, bd, i, k, j, r.dtbc, r.fcx, r.gcx,
                [&](int ii, int kk, int jj) { return A3(r.t, ii, kk, jj) * (g.c1h[kk] * A2(r.mut, ii, jj) + g.c2h[kk]); }, a, c);
    const float sv = 0.f + a - c;
    if (r.store_save) A3(r.t_save, i, k, j) = sv;
    A3(r.t_tendf, i, k, j) = A3(r.t_tendf, i, k, j) + sv;
  }
  if (k == 1) {   // mu relaxes straight into mu_tend (zeroed by rk_tendency)
    relax_terms(g, r.bmu, e, side, bd, i, 1, j, r.dtbc, r.fcx, r.gcx, [&](int ii, int, int jj) { return A2(r.mu, ii, jj); }, a, c);
    A2(r.mu_tend, i, j) = A2(r.mu_tend, i, j) + a - c;
  }
  if (r.nested) {
    relax_terms(g, r.bw, e, side, bd, i, k, j, r.dtbc, r.fcx, r.gcx,
                [&](int ii, int kk, int jj) { return A3(r.w, ii, kk, jj) * (g.c1f[kk] * A2(r.mut, ii, jj) + g.c2f[kk]); }, a, c);
    const float sv = 0.f + a - c;
    if (r.store_save) A3(r.w_save, i, k, j) = sv;
    A3(r.rw_tendf, i, k, j) = A3(r.rw_tendf, i, k, j) + sv * A2(g.msfty, i, j);
  }
}

struct SpecDry {
  fp ru_tend, rv_tend, ph_tend, t_tend, rw_tend, mu_tend;
  BdyArr bu, bv, bph, bt, bw, bmu;
  int nested;
};
// spec_bdy_dry (module_bc_em.F:413-530): the tendencies of the specified rows are the boundary file's
__global__ void k_spec_dry(const Geo g, Box b, int side, SpecDry r) {
  TID3;
  const int sz = g.spec_zone, ktf = g.kde - 1;
  if (k <= ktf) {
    const Ext e = ext_of(g, 'u');
    const int bd = zone_b(g, e, side, 0, sz, i, j);
    if (bd >= 0) A3(r.ru_tend, i, k, j) = r.bu.t[side][bat(g, r.bu, side, side < 2 ? i : j, k, bd)];
  }
  if (k <= ktf) {
    const Ext e = ext_of(g, 'v');
    const int bd = zone_b(g, e, side, 0, sz, i, j);
    if (bd >= 0) A3(r.rv_tend, i, k, j) = r.bv.t[side][bat(g, r.bv, side, side < 2 ? i : j, k, bd)];
  }
  const Ext e = ext_of(g, 't');
  const int bd = zone_b(g, e, side, 0, sz, i, j);
  if (bd < 0) return;
  const int al = side < 2 ? i : j;
  A3(r.ph_tend, i, k, j) = r.bph.t[side][bat(g, r.bph, side, al, k, bd)];
  if (k <= ktf) A3(r.t_tend, i, k, j) = r.bt.t[side][bat(g, r.bt, side, al, k, bd)];
  if (k == 1) A2(r.mu_tend, i, j) = r.bmu.t[side][bat(g, r.bmu, side, al, 1, bd)];
  if (r.nested) A3(r.rw_tend, i, k, j) = r.bw.t[side][bat(g, r.bw, side, al, k, bd)];
}

// spec_bdyupdate of u, v after advance_uv (solve_em.F:1344-1365)
__global__ void k_spec_update_uv(const Geo g, Box b, int side, float dt, fp u, cfp ru_tend, fp v, cfp rv_tend) {
  TID3;
  const int sz = g.spec_zone;
  if (zone_b(g, ext_of(g, 'u'), side, 0, sz, i, j) >= 0) A3(u, i, k, j) = A3(u, i, k, j) + dt * A3(ru_tend, i, k, j);
  if (zone_b(g, ext_of(g, 'v'), side, 0, sz, i, j) >= 0) A3(v, i, k, j) = A3(v, i, k, j) + dt * A3(rv_tend, i, k, j);
}
// spec_bdyupdate of t, mu, muts after advance_mu_t (solve_em.F:1459-1490)
__global__ void k_spec_update_tmu(const Geo g, Box b, int side, float dt, fp t, cfp t_tend, fp mu, fp muts, cfp mu_tend) {
  TID3;
  if (zone_b(g, ext_of(g, 't'), side, 0, g.spec_zone, i, j) < 0) return;
  A3(t, i, k, j) = A3(t, i, k, j) + dt * A3(t_tend, i, k, j);
  if (k == 1) {
    A2(mu, i, j) = A2(mu, i, j) + dt * A2(mu_tend, i, j);
    A2(muts, i, j) = A2(muts, i, j) + dt * A2(mu_tend, i, j);
  }
}
// spec_bdyupdate_ph (module_bc_em.F:17-157), then zero_grad_bdy of w (specified) or spec_bdyupdate of w (nested),
// after sumflux (solve_em.F:1582-1620)
__global__ void k_spec_update_phw(const Geo g, Box b, int side, float dt, int specified, cfp ph_save, fp ph, cfp ph_tend, cfp mu_tend, cfp muts,
                                  fp w, cfp rw_tend) {
  TID3;
  const int sz = g.spec_zone;
  const Ext e = ext_of(g, 't');
  if (zone_b(g, e, side, 0, sz, i, j) < 0) return;
  const float mts = A2(muts, i, j);
  const float mu_old = mts - dt * A2(mu_tend, i, j);
  const float c1 = g.c1f[k], c2 = g.c2f[k];
  A3(ph, i, k, j) = A3(ph, i, k, j) * (c1 * mu_old + c2) / (c1 * mts + c2) + dt * A3(ph_tend, i, k, j) / (c1 * mts + c2) +
                    A3(ph_save, i, k, j) * ((c1 * mu_old + c2) / (c1 * mts + c2) - 1.f);
  if (specified) {   // zero_grad_bdy, module_bc.F:2219-2332: the value of the nearest point inside the specified rows
    if (side < 2) {
      int i_inner = min(max(i, g.ids + sz), e.ibe - sz);
      if (g.periodic_x) i_inner = i;
      A3(w, i, k, j) = A3(w, i_inner, k, side == YS ? g.jds + sz : e.jbe - sz);
    } else {
      const int j_inner = min(max(j, g.jds + sz), e.jbe - sz);
      A3(w, i, k, j) = A3(w, side == XS ? g.ids + sz : e.ibe - sz, k, j_inner);
    }
  } else {
    A3(w, i, k, j) = A3(w, i, k, j) + dt * A3(rw_tend, i, k, j);
  }
}

// relax_bdy_scalar + spec_bdy_scalar (module_bc_em.F:348-410, :658-701; variable 'q': kde - 1 levels)
__global__ void k_bdy_scalar(const Geo g, Box b, int side, fp tend, cfp scalar, cfp mut, BdyArr B, const float *fcx, const float *gcx, float dtbc) {
  TID3;
  const Ext e = ext_of(g, 't');
  const int bd = zone_b(g, e, side, 0, g.relax_zone, i, j);
  if (bd < 0) return;
  if (bd >= g.spec_zone) {
    float a, c;
    relax_terms(g, B, e, side, bd, i, k, j, dtbc, fcx, gcx,
                [&](int ii, int kk, int jj) { return A3(scalar, ii, kk, jj) * (g.c1h[kk] * A2(mut, ii, jj) + g.c2h[kk]); }, a, c);
    A3(tend, i, k, j) = A3(tend, i, k, j) + a - c;
  } else {
    A3(tend, i, k, j) = B.t[side][bat(g, B, side, side < 2 ? i : j, k, bd)];
  }
}

// flow_dep_bdy (module_bc.F:2335-2456): zero gradient on outflow, zero on inflow
__global__ void k_flow_dep(const Geo g, Box b, int side, fp field, cfp u, cfp v) {
  TID3;
  const int sz = g.spec_zone;
  const Ext e = ext_of(g, 't');
  if (zone_b(g, e, side, 0, sz, i, j) < 0) return;
  float val = 0.f;
  if (side < 2) {
    int i_inner = min(max(i, g.ids + sz), e.ibe - sz);
    if (g.periodic_x) i_inner = i;
    const bool outflow = side == YS ? A3(v, i, k, j) < 0.f : A3(v, i, k, j + 1) > 0.f;
    if (outflow) val = A3(field, i_inner, k, side == YS ? g.jds + sz : e.jbe - sz);
  } else {
    const int j_inner = min(max(j, g.jds + sz), e.jbe - sz);
    const bool outflow = side == XS ? A3(u, i, k, j) < 0.f : A3(u, i + 1, k, j) > 0.f;
    if (outflow) val = A3(field, side == XS ? g.ids + sz : e.ibe - sz, k, j_inner);
  }
  A3(field, i, k, j) = val;
}

struct FinalIn {
  fp u, v, w, t, ph, mu;
  cfp muus, muvs, muts;
  BdyArr bu, bv, bw, bt, bph, bmu;
  int nested;
  float dtbc;
};
// spec_bdy_final (module_bc.F:2066-2216) of the dry fields: the specified rows are reset to the boundary file's value
__global__ void k_spec_final_dry(const Geo g, Box b, int side, FinalIn r) {
  TID3;
  const int sz = g.spec_zone, ktf = g.kde - 1;
  if (k <= ktf) {
    const Ext e = ext_of(g, 'u');
    const int bd = zone_b(g, e, side, 0, sz, i, j);
    if (bd >= 0) A3(r.u, i, k, j) = A2(g.msfuy, i, j) * bvalue(g, r.bu, side, side < 2 ? i : j, k, bd, r.dtbc) / (g.c1h[k] * A2(r.muus, i, j) + g.c2h[k]);
  }
  if (k <= ktf) {
    const Ext e = ext_of(g, 'v');
    const int bd = zone_b(g, e, side, 0, sz, i, j);
    if (bd >= 0) A3(r.v, i, k, j) = A2(g.msfvx, i, j) * bvalue(g, r.bv, side, side < 2 ? i : j, k, bd, r.dtbc) / (g.c1h[k] * A2(r.muvs, i, j) + g.c2h[k]);
  }
  const Ext e = ext_of(g, 't');
  const int bd = zone_b(g, e, side, 0, sz, i, j);
  if (bd < 0) return;
  const int al = side < 2 ? i : j;
  const float mts = A2(r.muts, i, j);
  if (r.nested) A3(r.w, i, k, j) = A2(g.msfty, i, j) * bvalue(g, r.bw, side, al, k, bd, r.dtbc) / (g.c1f[k] * mts + g.c2f[k]);
  if (k <= ktf) A3(r.t, i, k, j) = 1.f * bvalue(g, r.bt, side, al, k, bd, r.dtbc) / (g.c1h[k] * mts + g.c2h[k]);
  A3(r.ph, i, k, j) = 1.f * bvalue(g, r.bph, side, al, k, bd, r.dtbc) / (g.c1f[k] * mts + g.c2f[k]);
  if (k == 1) A2(r.mu, i, j) = 1.f * bvalue(g, r.bmu, side, al, 1, bd, r.dtbc) / 1.f;
}
__global__ void k_spec_final_scalar(const Geo g, Box b, int side, fp q, cfp muts, BdyArr B, float dtbc) {
  TID3;
  const int bd = zone_b(g, ext_of(g, 't'), side, 0, g.spec_zone, i, j);
  if (bd < 0) return;
  A3(q, i, k, j) = 1.f * bvalue(g, B, side, side < 2 ? i : j, k, bd, dtbc) / (g.c1h[k] * A2(muts, i, j) + g.c2h[k]);
}

// the rows strip `side` can hold for b_dist in [b_lo, b_hi), for any staggering, cut to the tile
Box strip_box(const Geo &q, int side, int b_lo, int b_hi, int k0, int k1) {
  Box bx{q.its, q.ite, k0, k1, q.jts, q.jte};
  if (side == YS) { bx.j0 = std::max(bx.j0, q.jds + b_lo); bx.j1 = std::min(bx.j1, q.jds + b_hi - 1); }
  if (side == YE) { bx.j0 = std::max(bx.j0, q.jde - b_hi); bx.j1 = std::min(bx.j1, q.jde - b_lo); }
  if (side == XS) { bx.i0 = std::max(bx.i0, q.ids + b_lo); bx.i1 = std::min(bx.i1, q.ids + b_hi - 1); }
  if (side == XE) { bx.i0 = std::max(bx.i0, q.ide - b_hi); bx.i1 = std::min(bx.i1, q.ide - b_lo); }
  return bx;
}
}  // namespace

#define FOR_SIDES(S, b_lo, b_hi, k0, k1, kern, ...)                                                      \
  for (int side_ = 0; side_ < ((S)->g.periodic_x ? 2 : 4); side_++)                                     \
    LAUNCH3(S, kern, strip_box((S)->g, side_, b_lo, b_hi, k0, k1), side_, __VA_ARGS__)

void Solver::add_raw(const std::string &n, size_t count) {
  Field f;
  f.kind = 1;   // contiguous: copied like the 1-D coefficient arrays
  f.n = f.nd = count;
  CUDA_CHECK(cudaMalloc(&f.d, count * sizeof(float)));
  CUDA_CHECK(cudaMemset(f.d, 0, count * sizeof(float)));
  f.o = f.d;
  f.raw = true;
  fields[n] = f;
  names.push_back(n);
}

void Solver::add_bdy(const std::string &f, int levels) {
  const size_t W = c.spec_bdy_width, ni = d.ime - d.ims + 1, nj = d.jme - d.jms + 1;
  for (const char *t : {"_b", "_bt"}) {
    add_raw(f + t + "xs", nj * levels * W); add_raw(f + t + "xe", nj * levels * W);
    add_raw(f + t + "ys", ni * levels * W); add_raw(f + t + "ye", ni * levels * W);
  }
}

BdyArr Solver::bdy_arr(const std::string &f) {
  static const char *sfx[4] = {"ys", "ye", "xs", "xe"};
  BdyArr B;
  for (int q = 0; q < 4; q++) { B.v[q] = F(f + "_b" + sfx[q]).d; B.t[q] = F(f + "_bt" + sfx[q]).d; }
  B.nk = f == "mu" ? 1 : d.kde - d.kds + 1;
  return B;
}

void Solver::relax_bdy_dry(bool store_save) {
  RelaxDry r;
  r.ru = o("ru"); r.rv = o("rv"); r.ph = o("ph_2"); r.t = o("t_2"); r.w = o("w_2"); r.mu = o("mu_2"); r.mut = o("mut");
  r.u_save = o("u_save"); r.v_save = o("v_save"); r.ph_save = o("ph_save"); r.t_save = o("t_save"); r.w_save = o("w_save"); r.mu_tend = o("mu_tend");
  r.ru_tendf = o("ru_tendf"); r.rv_tendf = o("rv_tendf"); r.ph_tendf = o("ph_tendf"); r.t_tendf = o("t_tendf"); r.rw_tendf = o("rw_tendf");
  r.bu = bdy_arr("u"); r.bv = bdy_arr("v"); r.bph = bdy_arr("ph"); r.bt = bdy_arr("t"); r.bw = bdy_arr("w"); r.bmu = bdy_arr("mu");
  r.fcx = F("fcx").d; r.gcx = F("gcx").d; r.dtbc = dtbc; r.nested = c.nested; r.store_save = store_save ? 1 : 0;
  FOR_SIDES(this, c.spec_zone, c.relax_zone, 1, g.kde, k_relax_dry, r);
}

void Solver::spec_bdy_dry() {
  SpecDry r;
  r.ru_tend = o("ru_tend"); r.rv_tend = o("rv_tend"); r.ph_tend = o("ph_tend"); r.t_tend = o("t_tend"); r.rw_tend = o("rw_tend"); r.mu_tend = o("mu_tend");
  r.bu = bdy_arr("u"); r.bv = bdy_arr("v"); r.bph = bdy_arr("ph"); r.bt = bdy_arr("t"); r.bw = bdy_arr("w"); r.bmu = bdy_arr("mu");
  r.nested = c.nested;
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde, k_spec_dry, r);
}

void Solver::spec_update_uv(float dts_rk) {
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde - 1, k_spec_update_uv, dts_rk, o("u_2"), (cfp)o("ru_tend"), o("v_2"), (cfp)o("rv_tend"));
}
void Solver::spec_update_tmu(float dts_rk) {
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde - 1, k_spec_update_tmu, dts_rk, o("t_2"), (cfp)o("t_tend"), o("mu_2"), o("muts"), (cfp)o("mu_tend"));
}
void Solver::spec_update_phw(float dts_rk) {
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde, k_spec_update_phw, dts_rk, (int)c.specified, (cfp)o("ph_save"), o("ph_2"), (cfp)o("ph_tend"),
            (cfp)o("mu_tend"), (cfp)o("muts"), o("w_2"), (cfp)o("rw_tend"));
}

bool Solver::has_bdy(const std::string &f) const { return fields.count(f + "_bxs") != 0; }

void Solver::bdy_scalar(const Species &sp) {
  FOR_SIDES(this, 0, c.relax_zone, 1, g.kde - 1, k_bdy_scalar, o(sp.t), (cfp)o(sp.f), (cfp)o("mut"), bdy_arr(sp.f), (const float *)F("fcx").d,
            (const float *)F("gcx").d, dtbc);
}
void Solver::flow_dep_bdy(const std::string &f) {
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde - 1, k_flow_dep, o(f), (cfp)o("ru_m"), (cfp)o("rv_m"));
}

// solve_em.F:4540-4722
void Solver::spec_final() {
  FinalIn r;
  r.u = o("u_2"); r.v = o("v_2"); r.w = o("w_2"); r.t = o("t_2"); r.ph = o("ph_2"); r.mu = o("mu_2");
  r.muus = o("muus"); r.muvs = o("muvs"); r.muts = o("muts");
  r.bu = bdy_arr("u"); r.bv = bdy_arr("v"); r.bw = bdy_arr("w"); r.bt = bdy_arr("t"); r.bph = bdy_arr("ph"); r.bmu = bdy_arr("mu");
  r.nested = c.nested; r.dtbc = dtbc;
  FOR_SIDES(this, 0, c.spec_zone, 1, g.kde, k_spec_final_dry, r);
  for (auto &sp : species) {
    const bool moist = sp.f.rfind("moist_", 0) == 0;
    if (!has_bdy(sp.f) || !(moist || c.nested)) continue;   // scalars only when nested (:4702)
    FOR_SIDES(this, 0, c.spec_zone, 1, g.kde - 1, k_spec_final_scalar, o(sp.f), (cfp)o("muts"), bdy_arr(sp.f), dtbc);
  }
}

}  // namespace wg
