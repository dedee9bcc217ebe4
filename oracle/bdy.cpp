// bdy.cpp -- oracle restatement of the specified / nested lateral boundary routines:
// share/module_bc.F:1122-2456 (relax_bdytend, spec_bdytend, spec_bdyupdate, spec_bdy_final, zero_grad_bdy, flow_dep_bdy)
// and dyn_em/module_bc_em.F:17-700 (spec_bdyupdate_ph, relax_bdy_dry, relax_bdy_scalar, spec_bdy_dry, spec_bdy_scalar).
// Test infrastructure only; see oracle.h header.
//
// Every routine of share/module_bc.F walks the same four strips -- Y-start, Y-end (they own the corners: the i range
// shrinks by b_dist on each side unless periodic_x), then X-start, X-end (j strictly inside the Y strips at the same
// distance; skipped when periodic_x) -- over the rows b_lo <= b_dist < b_hi.  `zone` is that walk, written once; the
// bodies below are the reference's loop bodies.
#include "oracle.h"
namespace orc {

namespace {
enum Side { YS = 0, YE = 1, XS = 2, XE = 3 };

struct Ext { int ibs, ibe, itf, jbs, jbe, jtf, ktf; };
// the index set-up every routine starts with (e.g. module_bc.F:1283-1296); ktf_h: the value 'm' / 'h' (and 'w') take
Ext extents(const Cx &x, char variable, int kts, int kte, int ktf_h) {
  (void)kts;
  Ext e;
  e.ibs = x.ids; e.ibe = x.ide - 1; e.itf = std::min(x.ite, x.ide - 1);
  e.jbs = x.jds; e.jbe = x.jde - 1; e.jtf = std::min(x.jte, x.jde - 1);
  e.ktf = x.kde - 1;
  if (variable == 'u') { e.ibe = x.ide; e.itf = std::min(x.ite, x.ide); }
  if (variable == 'v') { e.jbe = x.jde; e.jtf = std::min(x.jte, x.jde); }
  if (variable == 'm' || variable == 'h' || variable == 'w') e.ktf = ktf_h;
  (void)kte;
  return e;
}

template <class Fn> void zone(const Cx &x, const Ext &e, int b_lo, int b_hi, int kts, Fn body) {
  const int its = x.its, jts = x.jts;
  const bool periodic_x = x.c.periodic_x;
  for (int j = std::max(jts, e.jbs + b_lo); j <= std::min(e.jtf, e.jbs + b_hi - 1); j++) {   // Y-start boundary
    const int b_dist = j - e.jbs;
    const int b_limit = periodic_x ? 0 : b_dist;
    for (int k = kts; k <= e.ktf; k++)
      for (int i = std::max(its, b_limit + e.ibs); i <= std::min(e.itf, e.ibe - b_limit); i++) body(YS, b_dist, i, k, j);
  }
  for (int j = std::max(jts, e.jbe - b_hi + 1); j <= std::min(e.jtf, e.jbe - b_lo); j++) {   // Y-end boundary
    const int b_dist = e.jbe - j;
    const int b_limit = periodic_x ? 0 : b_dist;
    for (int k = kts; k <= e.ktf; k++)
      for (int i = std::max(its, b_limit + e.ibs); i <= std::min(e.itf, e.ibe - b_limit); i++) body(YE, b_dist, i, k, j);
  }
  if (periodic_x) return;
  for (int i = std::max(its, e.ibs + b_lo); i <= std::min(e.itf, e.ibs + b_hi - 1); i++) {   // X-start boundary
    const int b_dist = i - e.ibs;
    for (int k = kts; k <= e.ktf; k++)
      for (int j = std::max(jts, b_dist + e.jbs + 1); j <= std::min(e.jtf, e.jbe - b_dist - 1); j++) body(XS, b_dist, i, k, j);
  }
  for (int i = std::max(its, e.ibe - b_hi + 1); i <= std::min(e.itf, e.ibe - b_lo); i++) {   // X-end boundary
    const int b_dist = e.ibe - i;
    for (int k = kts; k <= e.ktf; k++)
      for (int j = std::max(jts, b_dist + e.jbs + 1); j <= std::min(e.jtf, e.jbe - b_dist - 1); j++) body(XE, b_dist, i, k, j);
  }
}
}  // namespace

// field_bdy_?s(along, k, b) + dtbc * field_bdy_tend_?s(along, k, b)
static inline float bdy_value(const Bdy &B, int side, int along, int k, int b, float dtbc) {
  return B.val(side, along, k, b) + dtbc * B.tend(side, along, k, b);
}

// module_bc.F:1221-1427 relax_bdytend_core; `field(i,k,j)` is the (mass-weighted) field, any callable
template <class Fld, class Tnd>
static void relax_bdytend(const Cx &x, Fld field, Tnd field_tend, const Bdy &B, char variable, int kts, int kte, float dtbc, A1x fcx, A1x gcx) {
  const Ext e = extents(x, variable, kts, kte, kte);
  const int spec_zone = x.c.spec_zone, relax_zone = x.c.relax_zone;
  zone(x, e, spec_zone, relax_zone, kts, [&](int side, int b_dist, int i, int k, int j) {
    float fls0, fls1, fls2, fls3, fls4;
    if (side == YS || side == YE) {
      const int im1 = std::max(i - 1, e.ibs), ip1 = std::min(i + 1, e.ibe);
      const int jo = side == YS ? j - 1 : j + 1, ji = side == YS ? j + 1 : j - 1;   // one row out, one row in
      fls0 = bdy_value(B, side, i, k, b_dist + 1, dtbc) - field(i, k, j);
      fls1 = bdy_value(B, side, im1, k, b_dist + 1, dtbc) - field(im1, k, j);
      fls2 = bdy_value(B, side, ip1, k, b_dist + 1, dtbc) - field(ip1, k, j);
      fls3 = bdy_value(B, side, i, k, b_dist, dtbc) - field(i, k, jo);
      fls4 = bdy_value(B, side, i, k, b_dist + 2, dtbc) - field(i, k, ji);
    } else {
      const int io = side == XS ? i - 1 : i + 1, ii = side == XS ? i + 1 : i - 1;
      fls0 = bdy_value(B, side, j, k, b_dist + 1, dtbc) - field(i, k, j);
      fls1 = bdy_value(B, side, j - 1, k, b_dist + 1, dtbc) - field(i, k, j - 1);
      fls2 = bdy_value(B, side, j + 1, k, b_dist + 1, dtbc) - field(i, k, j + 1);
      fls3 = bdy_value(B, side, j, k, b_dist, dtbc) - field(io, k, j);
      fls4 = bdy_value(B, side, j, k, b_dist + 2, dtbc) - field(ii, k, j);
    }
    field_tend(i, k, j) = field_tend(i, k, j) + fcx(b_dist + 1) * fls0 - gcx(b_dist + 1) * (fls1 + fls2 + fls3 + fls4 - 4.f * fls0);
  });
}

// module_bc.F:1430-1549
template <class Tnd> static void spec_bdytend(const Cx &x, Tnd field_tend, const Bdy &B, char variable, int kts, int kte) {
  const Ext e = extents(x, variable, kts, kte, kte);
  zone(x, e, 0, x.c.spec_zone, kts, [&](int side, int b_dist, int i, int k, int j) {
    field_tend(i, k, j) = B.tend(side, (side == YS || side == YE) ? i : j, k, b_dist + 1);
  });
}

// module_bc.F:1955-2064
template <class Fld, class Tnd> static void spec_bdyupdate_t(const Cx &x, Fld field, Tnd field_tend, float dt, char variable, int kts, int kte) {
  const Ext e = extents(x, variable, kts, kte, kte);
  zone(x, e, 0, x.c.spec_zone, kts, [&](int, int, int i, int k, int j) { field(i, k, j) = field(i, k, j) + dt * field_tend(i, k, j); });
}
void spec_bdyupdate(const Cx &x, A3 field, A3 field_tend, float dt, char variable) {
  spec_bdyupdate_t(x, field, field_tend, dt, variable, x.kts, x.kte);
}
void spec_bdyupdate(const Cx &x, A2 field, A2 field_tend, float dt) {
  auto f = [&](int i, int, int j) -> float & { return field(i, j); };
  auto t = [&](int i, int, int j) -> float & { return field_tend(i, j); };
  spec_bdyupdate_t(x, f, t, dt, 'm', 1, 1);
}

// module_bc_em.F:17-157
void spec_bdyupdate_ph(const Cx &x, A3 ph_save, A3 field, A3 field_tend, A2 mu_tend, A2 muts, A1 c1, A1 c2, float dt) {
  const Ext e = extents(x, 'h', x.kts, x.kte, x.kte);
  zone(x, e, 0, x.c.spec_zone, x.kts, [&](int, int, int i, int k, int j) {
    const float mu_old = muts(i, j) - dt * mu_tend(i, j);
    field(i, k, j) = field(i, k, j) * (c1(k) * mu_old + c2(k)) / (c1(k) * muts(i, j) + c2(k)) + dt * field_tend(i, k, j) / (c1(k) * muts(i, j) + c2(k)) +
                     ph_save(i, k, j) * ((c1(k) * mu_old + c2(k)) / (c1(k) * muts(i, j) + c2(k)) - 1.f);
  });
}

// module_bc.F:2219-2332 (variable 'w': ktf = kde)
void zero_grad_bdy(const Cx &x, A3 field, char variable) {
  Ext e = extents(x, variable == 'w' ? 't' : variable, x.kts, x.kte, x.kte);
  if (variable == 'w') e.ktf = x.kde;
  const int sz = x.c.spec_zone;
  zone(x, e, 0, sz, x.kts, [&](int side, int, int i, int k, int j) {
    if (side == YS || side == YE) {
      int i_inner = std::max(i, e.ibs + sz);
      i_inner = std::min(i_inner, e.ibe - sz);
      if (x.c.periodic_x) i_inner = i;
      field(i, k, j) = field(i_inner, k, side == YS ? e.jbs + sz : e.jbe - sz);
    } else {
      int j_inner = std::max(j, e.jbs + sz);
      j_inner = std::min(j_inner, e.jbe - sz);
      field(i, k, j) = field(side == XS ? e.ibs + sz : e.ibe - sz, k, j_inner);
    }
  });
}

// module_bc.F:2335-2456: zero gradient on outflow, zero value on inflow (unstaggered field; u, v only for their sign)
void flow_dep_bdy(const Cx &x, A3 field, A3 u, A3 v) {
  const Ext e = extents(x, 't', x.kts, x.kte, x.kte);
  const int sz = x.c.spec_zone;
  zone(x, e, 0, sz, x.kts, [&](int side, int, int i, int k, int j) {
    if (side == YS || side == YE) {
      int i_inner = std::max(i, e.ibs + sz);
      i_inner = std::min(i_inner, e.ibe - sz);
      if (x.c.periodic_x) i_inner = i;
      const bool outflow = side == YS ? v(i, k, j) < 0.f : v(i, k, j + 1) > 0.f;
      field(i, k, j) = outflow ? field(i_inner, k, side == YS ? e.jbs + sz : e.jbe - sz) : 0.f;
    } else {
      int j_inner = std::max(j, e.jbs + sz);
      j_inner = std::min(j_inner, e.jbe - sz);
      const bool outflow = side == XS ? u(i, k, j) < 0.f : u(i + 1, k, j) > 0.f;
      field(i, k, j) = outflow ? field(side == XS ? e.ibs + sz : e.ibe - sz, k, j_inner) : 0.f;
    }
  });
}

// module_bc.F:2066-2216: force the specified rows to the boundary-file value (coupling u, v, w: mu/msf, others: mu)
template <class Fld> static void spec_bdy_final_t(const Cx &x, Fld field, A2 mu, A1 c1, A1 c2, A2 msf, const Bdy &B, char variable, int kts, int ktf_h, float dtbc) {
  const Ext e = extents(x, variable, kts, ktf_h, ktf_h);
  const bool msfcouple = variable == 'u' || variable == 'v' || variable == 'w';
  const bool mucouple = variable != 'm';
  zone(x, e, 0, x.c.spec_zone, kts, [&](int side, int b_dist, int i, int k, int j) {
    const float bfield = bdy_value(B, side, (side == YS || side == YE) ? i : j, k, b_dist + 1, dtbc);
    float xmsf = 1.f, xmu = 1.f;
    if (msfcouple) xmsf = msf(i, j);
    if (mucouple) xmu = c1(k) * mu(i, j) + c2(k);
    field(i, k, j) = xmsf * bfield / xmu;
  });
}
void spec_bdy_final(const Cx &x, A3 field, A2 mu, A1 c1, A1 c2, A2 msf, const Bdy &B, char variable, float dtbc) {
  spec_bdy_final_t(x, field, mu, c1, c2, msf, B, variable, x.kts, x.kde, dtbc);   // 'h', 'w': ktf = kde; others kde - 1
}
void spec_bdy_final(const Cx &x, A2 field, A2 mu, A1 c1, A1 c2, A2 msf, const Bdy &B, float dtbc) {
  auto f = [&](int i, int, int j) -> float & { return field(i, j); };
  spec_bdy_final_t(x, f, mu, c1, c2, msf, B, 'm', 1, 1, dtbc);   // called with kds = kde = 1 (solve_em.F:4616-4627)
}

// module_bc_em.F:161-346.  The tendencies go to the *_save arrays rk_tendency has just zeroed (solve_em.F:939-941);
// rk_addtend_dry folds them into *_tendf on the first RK substep.
void relax_bdy_dry(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 ph_tendf, A3 t_tendf, A3 rw_tendf, A2 mu_tend, A3 ru, A3 rv, A3 ph, A3 t, A3 w,
                   A2 mu, A2 mut, const Bdy &bu, const Bdy &bv, const Bdy &bph, const Bdy &bt, const Bdy &bw, const Bdy &bmu, float dtbc,
                   A1x fcx, A1x gcx) {
  const int kts = x.kts, kte = x.kte;
  relax_bdytend(x, ru, ru_tendf, bu, 'u', kts, kte, dtbc, fcx, gcx);
  relax_bdytend(x, rv, rv_tendf, bv, 'v', kts, kte, dtbc, fcx, gcx);
  // rfield = mass_weight(ph, mut, c1f, c2f) one point beyond the tile; relax_bdytend_tile, 'h' (ktf = kte)
  auto rph = [&](int i, int k, int j) { return ph(i, k, j) * (x.c1f(k) * mut(i, j) + x.c2f(k)); };
  relax_bdytend(x, rph, ph_tendf, bph, 'h', kts, kte, dtbc, fcx, gcx);
  auto rt = [&](int i, int k, int j) { return t(i, k, j) * (x.c1h(k) * mut(i, j) + x.c2h(k)); };
  relax_bdytend(x, rt, t_tendf, bt, 't', kts, kte, dtbc, fcx, gcx);
  auto m = [&](int i, int, int j) -> float & { return mu(i, j); };
  auto mt = [&](int i, int, int j) -> float & { return mu_tend(i, j); };
  relax_bdytend(x, m, mt, bmu, 'm', 1, 1, dtbc, fcx, gcx);
  if (x.c.nested) {
    auto rw = [&](int i, int k, int j) { return w(i, k, j) * (x.c1f(k) * mut(i, j) + x.c2f(k)); };
    relax_bdytend(x, rw, rw_tendf, bw, 'h', kts, kte, dtbc, fcx, gcx);
  }
}

// module_bc_em.F:348-410 (variable 'q': ktf = kde - 1)
void relax_bdy_scalar(const Cx &x, A3 scalar_tend, A3 scalar, A2 mu, const Bdy &B, float dtbc, A1x fcx, A1x gcx) {
  auto rs = [&](int i, int k, int j) { return scalar(i, k, j) * (x.c1h(k) * mu(i, j) + x.c2h(k)); };
  relax_bdytend(x, rs, scalar_tend, B, 'q', x.kts, x.kte, dtbc, fcx, gcx);
}

// module_bc_em.F:413-530
void spec_bdy_dry(const Cx &x, A3 ru_tend, A3 rv_tend, A3 ph_tend, A3 t_tend, A3 rw_tend, A2 mu_tend, const Bdy &bu, const Bdy &bv,
                  const Bdy &bph, const Bdy &bt, const Bdy &bw, const Bdy &bmu) {
  const int kts = x.kts, kte = x.kte;
  spec_bdytend(x, ru_tend, bu, 'u', kts, kte);
  spec_bdytend(x, rv_tend, bv, 'v', kts, kte);
  spec_bdytend(x, ph_tend, bph, 'h', kts, kte);
  spec_bdytend(x, t_tend, bt, 't', kts, kte);
  auto mt = [&](int i, int, int j) -> float & { return mu_tend(i, j); };
  spec_bdytend(x, mt, bmu, 'm', 1, 1);
  if (x.c.nested) spec_bdytend(x, rw_tend, bw, 'h', kts, kte);
}

// module_bc_em.F:658-701
void spec_bdy_scalar(const Cx &x, A3 scalar_tend, const Bdy &B) { spec_bdytend(x, scalar_tend, B, 'q', x.kts, x.kte); }

}  // namespace orc
