// advect_mono.cpp -- advect_scalar_mono (dyn_em/module_advect_em.F:9495-10560), the monotonic flux limiter
// (Smolarkiewicz MWR 1989 with first-order fluxes added first) used on the last RK substep when
// moist_adv_opt / scalar_adv_opt = 2.  h_sca_adv_order = 5, v_sca_adv_order = 3, tenddec = .false.
// TEST INFRASTRUCTURE (see oracle.h).  Same structure as advect_scalar_pd (advect.cpp): high-order minus upwind
// fluxes on the tile + 1 ring, a limiter pass, the flux divergence; what differs is the upwind flux
// (vel * flux_upwind with cr = vel, :9602-9603), the bounds qmin/qmax gathered from the upwind neighbours of
// every computed face, and the two-sided scale_in / scale_out factors.
#include "oracle.h"

namespace orc {

void advect_scalar_mono(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom, A3 romI,
                        A2 mut, A2 mub, A2 mu_old, float dt) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int ktf = std::min(kte, kde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &fzm = x.fnm, &fzp = x.fnp, &rdzw = x.rdnw;
  const A2 &msftx = x.msftx, &msfty = x.msfty;
  const float rdx = x.rdx, rdy = x.rdy;
  const float eps = 1.e-20f;
  // statement functions :9584-9603 (no sign(time_step) factor in this routine)
  auto flux4 = [](float q_im2, float q_im1, float q_i, float q_ip1) {
    return (7.f / 12.f) * (q_i + q_im1) - (1.f / 12.f) * (q_ip1 + q_im2);
  };
  auto flux3 = [&](float q_im2, float q_im1, float q_i, float q_ip1, float ua) {
    return flux4(q_im2, q_im1, q_i, q_ip1) + sign1(ua) * (1.f / 12.f) * ((q_ip1 - q_im2) - 3.f * (q_i - q_im1));
  };
  auto flux6 = [](float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2) {
    return (37.f / 60.f) * (q_i + q_im1) - (2.f / 15.f) * (q_ip1 + q_im2) + (1.f / 60.f) * (q_ip2 + q_im3);
  };
  auto flux5 = [&](float q_im3, float q_im2, float q_im1, float q_i, float q_ip1, float q_ip2, float ua) {
    return flux6(q_im3, q_im2, q_im1, q_i, q_ip1, q_ip2) -
           sign1(ua) * (1.f / 60.f) * ((q_ip2 - q_im3) - 5.f * (q_ip1 - q_im2) + 10.f * (q_i - q_im1));
  };
  auto flux_upwind = [](float q_im1, float q_i, float cr) {
    return 0.5f * (1.f + sign1(cr)) * q_im1 + 0.5f * (1.f - sign1(cr)) * q_i;
  };
  // (its-2:ite+2, kts:kte, jts-2:jte+2)
  const int ni = ite - its + 5, nk = kte - kts + 1, nj = jte - jts + 5;
  const size_t n = (size_t)ni * nk * nj;
  std::vector<float> b_fqx(n, 0.f), b_fqy(n, 0.f), b_fqz(n, 0.f), b_fqxl(n, 0.f), b_fqyl(n, 0.f), b_fqzl(n, 0.f);
  std::vector<float> b_qmin(n), b_qmax(n), b_sin(n, 1.f), b_sout(n, 1.f);
  auto idx = [&](int i, int k, int j) { return (size_t)(i - its + 2) + (size_t)(k - kts) * ni + (size_t)(j - jts + 2) * ni * nk; };
#define L3(name, buf) auto name = [&](int i, int k, int j) -> float & { return buf[idx(i, k, j)]; }
  L3(fqx, b_fqx); L3(fqy, b_fqy); L3(fqz, b_fqz); L3(fqxl, b_fqxl); L3(fqyl, b_fqyl); L3(fqzl, b_fqzl);
  L3(qmin, b_qmin); L3(qmax, b_qmax); L3(scale_in, b_sin); L3(scale_out, b_sout);
#undef L3
  for (int j = jts - 2; j <= jte + 2; j++)
    for (int k = kts; k <= kte; k++)
      for (int i = its - 2; i <= ite + 2; i++) { qmin(i, k, j) = field_old(i, k, j); qmax(i, k, j) = field_old(i, k, j); }
  // the face between `lo` (upwind when cr > 0) and `hi` cells
  auto bounds = [&](float cr, int il, int kl, int jl, int ih, int kh, int jh) {
    if (cr > 0.f) {
      qmax(ih, kh, jh) = std::max(qmax(ih, kh, jh), field_old(il, kl, jl));
      qmin(ih, kh, jh) = std::min(qmin(ih, kh, jh), field_old(il, kl, jl));
    } else {
      qmax(il, kl, jl) = std::max(qmax(il, kl, jl), field_old(ih, kh, jh));
      qmin(il, kl, jl) = std::min(qmin(il, kl, jl), field_old(ih, kh, jh));
    }
  };
  bool degrade_xs = true, degrade_xe = true, degrade_ys = true, degrade_ye = true;
  if (x.c.periodic_x || x.c.symmetric_xs || (its > ids + 3)) degrade_xs = false;
  if (x.c.periodic_x || x.c.symmetric_xe || (ite < ide - 4)) degrade_xe = false;
  if (x.c.periodic_y || x.c.symmetric_ys || (jts > jds + 3)) degrade_ys = false;
  if (x.c.periodic_y || x.c.symmetric_ye || (jte < jde - 4)) degrade_ye = false;

  // y fluxes (:9663-9828)
  int i_start = its - 1, i_end = std::min(ite, ide - 1) + 1, j_start = jts - 1, j_end = std::min(jte, jde - 1) + 1;
  int j_start_f = j_start, j_end_f = j_end + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) { j_start = std::max(jts - 1, jds + 1); j_start_f = jds + 3; }
  if (degrade_ye) { j_end = std::min(jte + 1, jde - 2); j_end_f = jde - 3; }
  for (int j = j_start; j <= j_end + 1; j++) {
    int mode;  // 5, 3, 2 or 0 (none)
    if (j >= j_start_f && j <= j_end_f) mode = 5;
    else if (j == jds + 1) mode = 2;
    else if (j == jds + 2) mode = 3;
    else if (j == jde - 1) mode = 2;
    else if (j == jde - 2) mode = 3;
    else mode = 0;
    if (!mode) continue;
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float vel = rv(i, k, j);
        const float cr = vel;
        fqyl(i, k, j) = vel * flux_upwind(field_old(i, k, j - 1), field_old(i, k, j), cr);
        if (mode == 5)
          fqy(i, k, j) = vel * flux5(field(i, k, j - 3), field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), field(i, k, j + 2), vel);
        else if (mode == 2)
          fqy(i, k, j) = 0.5f * rv(i, k, j) * (field(i, k, j) + field(i, k, j - 1));
        else
          fqy(i, k, j) = vel * flux3(field(i, k, j - 2), field(i, k, j - 1), field(i, k, j), field(i, k, j + 1), vel);
        fqy(i, k, j) = fqy(i, k, j) - fqyl(i, k, j);
        bounds(cr, i, k, j - 1, i, k, j);
      }
  }
  // x fluxes (:9830-10012)
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1;
  int i_start_f = i_start, i_end_f = i_end + 1;
  j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  if (degrade_xs) { i_start = std::max(ids + 1, its - 1); i_start_f = ids + 3; }
  if (degrade_xe) { i_end = std::min(ide - 2, ite + 1); i_end_f = ide - 3; }
  for (int j = j_start; j <= j_end; j++) {
    auto doface = [&](int i, int k, int mode) {
      const float vel = ru(i, k, j);
      const float cr = vel;
      fqxl(i, k, j) = vel * flux_upwind(field_old(i - 1, k, j), field_old(i, k, j), cr);
      if (mode == 5)
        fqx(i, k, j) = vel * flux5(field(i - 3, k, j), field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), field(i + 2, k, j), vel);
      else if (mode == 2)
        fqx(i, k, j) = 0.5f * (ru(i, k, j)) * (field(i, k, j) + field(i - 1, k, j));
      else
        fqx(i, k, j) = vel * flux3(field(i - 2, k, j), field(i - 1, k, j), field(i, k, j), field(i + 1, k, j), vel);
      fqx(i, k, j) = fqx(i, k, j) - fqxl(i, k, j);
      bounds(cr, i - 1, k, j, i, k, j);
    };
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start_f; i <= i_end_f; i++) doface(i, k, 5);
    if (degrade_xs)
      for (int i = i_start; i <= i_start_f - 1; i++) {
        if (i == ids + 1) for (int k = kts; k <= ktf; k++) doface(i, k, 2);
        if (i == ids + 2) for (int k = kts; k <= ktf; k++) doface(i, k, 3);
      }
    if (degrade_xe)
      for (int i = i_end_f + 1; i <= i_end + 1; i++) {
        if (i == ide - 1) for (int k = kts; k <= ktf; k++) doface(i, k, 2);
        if (i == ide - 2) for (int k = kts; k <= ktf; k++) doface(i, k, 3);
      }
  }

  // open boundary terms (:10019-10083)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::min(0.5f * (ru(its, k, j) + ru(its + 1, k, j)), 0.f);
        tendency(its, k, j) = tendency(its, k, j) -
                              rdx * (ub * (field_old(its + 1, k, j) - field_old(its, k, j)) + field(its, k, j) * (ru(its + 1, k, j) - ru(its, k, j)));
      }
  if (x.c.open_xe && ite == ide)
    for (int j = j_start; j <= j_end; j++)
      for (int k = kts; k <= ktf; k++) {
        float ub = std::max(0.5f * (ru(ite - 1, k, j) + ru(ite, k, j)), 0.f);
        tendency(i_end, k, j) = tendency(i_end, k, j) -
                                rdx * (ub * (field_old(i_end, k, j) - field_old(i_end - 1, k, j)) + field(i_end, k, j) * (ru(ite, k, j) - ru(ite - 1, k, j)));
      }
  if (x.c.open_ys && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::min(0.5f * (rv(i, k, jts) + rv(i, k, jts + 1)), 0.f);
        tendency(i, k, jts) = tendency(i, k, jts) -
                              rdy * (vb * (field_old(i, k, jts + 1) - field_old(i, k, jts)) + field(i, k, jts) * (rv(i, k, jts + 1) - rv(i, k, jts)));
      }
  if (x.c.open_ye && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        float vb = std::max(0.5f * (rv(i, k, jte - 1) + rv(i, k, jte)), 0.f);
        tendency(i, k, j_end) = tendency(i, k, j_end) -
                                rdy * (vb * (field_old(i, k, j_end) - field_old(i, k, j_end - 1)) + field(i, k, j_end) * (rv(i, k, jte) - rv(i, k, jte - 1)));
      }

  // vertical fluxes, v_sca_adv_order = 3 (:10085-10151)
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1; j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  for (int j = j_start; j <= j_end; j++) {
    for (int i = i_start; i <= i_end; i++) {
      fqz(i, 1, j) = 0.f; fqzl(i, 1, j) = 0.f; fqz(i, kde, j) = 0.f; fqzl(i, kde, j) = 0.f;
    }
    auto zface = [&](int i, int k, bool second) {
      const float vel = rom(i, k, j);
      const float cr = -vel;
      fqzl(i, k, j) = vel * flux_upwind(field_old(i, k - 1, j), field_old(i, k, j), cr);
      if (second) fqz(i, k, j) = rom(i, k, j) * (fzm(k) * field(i, k, j) + fzp(k) * field(i, k - 1, j));
      else fqz(i, k, j) = vel * flux3(field(i, k - 2, j), field(i, k - 1, j), field(i, k, j), field(i, k + 1, j), -vel);
      fqz(i, k, j) = fqz(i, k, j) - fqzl(i, k, j);
      bounds(cr, i, k - 1, j, i, k, j);
    };
    for (int k = kts + 2; k <= ktf - 1; k++)
      for (int i = i_start; i <= i_end; i++) zface(i, k, false);
    for (int i = i_start; i <= i_end; i++) {
      zface(i, kts + 1, true);
      zface(i, ktf, true);
    }
  }

  // monotonic limiter (:10322-10420)
  i_start = its - 1; i_end = std::min(ite, ide - 1) + 1; j_start = jts - 1; j_end = std::min(jte, jde - 1) + 1;
  if (degrade_xs) i_start = std::max(its - 1, ids);
  if (degrade_xe) i_end = std::min(ite + 1, ide - 1);
  if (degrade_ys) j_start = std::max(jts - 1, jds);
  if (degrade_ye) j_end = std::min(jte + 1, jde - 1);
  if (x.c.specified || x.c.nested) {
    if (degrade_xs) i_start = std::max(its - 1, ids + 1);
    if (degrade_xe) i_end = std::min(ite + 1, ide - 2);
    if (degrade_ys) j_start = std::max(jts - 1, jds + 1);
    if (degrade_ye) j_end = std::min(jte + 1, jde - 2);
  }
  if (x.c.open_xs && degrade_xs) i_start = std::max(its - 1, ids + 1);
  if (x.c.open_xe && degrade_xe) i_end = std::min(ite + 1, ide - 2);
  if (x.c.open_ys && degrade_ys) j_start = std::max(jts - 1, jds + 1);
  if (x.c.open_ye && degrade_ye) j_end = std::min(jte + 1, jde - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        const float ieva_corr = (c1(k) * mut(i, j) + c2(k)) + dt * msfty(i, j) * rdzw(k) * (romI(i, k + 1, j) - romI(i, k, j));
        const float ph_upwind = ((c1(k) * mub(i, j) + c2(k)) + (c1(k) * mu_old(i, j))) * field_old(i, k, j) -
                                dt * (msftx(i, j) * msfty(i, j) * (rdx * (fqxl(i + 1, k, j) - fqxl(i, k, j)) + rdy * (fqyl(i, k, j + 1) - fqyl(i, k, j))) +
                                      msfty(i, j) * rdzw(k) * (fqzl(i, k + 1, j) - fqzl(i, k, j)));
        const float flux_in = -dt * ((msftx(i, j) * msfty(i, j)) * (rdx * (std::min(0.f, fqx(i + 1, k, j)) - std::max(0.f, fqx(i, k, j))) +
                                                                    rdy * (std::min(0.f, fqy(i, k, j + 1)) - std::max(0.f, fqy(i, k, j)))) +
                                     msfty(i, j) * rdzw(k) * (std::max(0.f, fqz(i, k + 1, j)) - std::min(0.f, fqz(i, k, j))));
        const float ph_hi = ieva_corr * qmax(i, k, j) - ph_upwind;
        if (flux_in > ph_hi) scale_in(i, k, j) = std::max(0.f, ph_hi / (flux_in + eps));
        const float flux_out = dt * ((msftx(i, j) * msfty(i, j)) * (rdx * (std::max(0.f, fqx(i + 1, k, j)) - std::min(0.f, fqx(i, k, j))) +
                                                                    rdy * (std::max(0.f, fqy(i, k, j + 1)) - std::min(0.f, fqy(i, k, j)))) +
                                     msfty(i, j) * rdzw(k) * (std::min(0.f, fqz(i, k + 1, j)) - std::max(0.f, fqz(i, k, j))));
        const float ph_low = ph_upwind - ieva_corr * qmin(i, k, j);
        if (flux_out > ph_low) scale_out(i, k, j) = std::max(0.f, ph_low / (flux_out + eps));
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end + 1; i++) {
        if (fqx(i, k, j) > 0.f) fqx(i, k, j) = std::min(scale_in(i, k, j), scale_out(i - 1, k, j)) * fqx(i, k, j);
        else fqx(i, k, j) = std::min(scale_out(i, k, j), scale_in(i - 1, k, j)) * fqx(i, k, j);
      }
  for (int j = j_start; j <= j_end + 1; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        if (fqy(i, k, j) > 0.f) fqy(i, k, j) = std::min(scale_in(i, k, j), scale_out(i, k, j - 1)) * fqy(i, k, j);
        else fqy(i, k, j) = std::min(scale_out(i, k, j), scale_in(i, k, j - 1)) * fqy(i, k, j);
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++) {
        if (fqz(i, k, j) < 0.f) fqz(i, k, j) = std::min(scale_in(i, k, j), scale_out(i, k - 1, j)) * fqz(i, k, j);
        else fqz(i, k, j) = std::min(scale_out(i, k, j), scale_in(i, k - 1, j)) * fqz(i, k, j);
      }

  // flux divergence (:10424-10556)
  i_start = its; i_end = std::min(ite, ide - 1); j_start = jts; j_end = std::min(jte, jde - 1);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - rdzw(k) * (fqz(i, k + 1, j) - fqz(i, k, j) + fqzl(i, k + 1, j) - fqzl(i, k, j));
  if (degrade_xs) i_start = std::max(its, ids + 1);
  if (degrade_xe) i_end = std::min(ite, ide - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - msftx(i, j) * (rdx * (fqx(i + 1, k, j) - fqx(i, k, j) + fqxl(i + 1, k, j) - fqxl(i, k, j)));
  i_start = its; i_end = std::min(ite, ide - 1);
  if (degrade_ys) j_start = std::max(jts, jds + 1);
  if (degrade_ye) j_end = std::min(jte, jde - 2);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        tendency(i, k, j) = tendency(i, k, j) - msftx(i, j) * (rdy * (fqy(i, k, j + 1) - fqy(i, k, j) + fqyl(i, k, j + 1) - fqyl(i, k, j)));
}

}  // namespace orc
