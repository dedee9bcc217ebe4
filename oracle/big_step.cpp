// big_step.cpp -- oracle restatement of the dyn_em/module_big_step_utilities_em.F
// routines on the solve_em path (+ rk_addtend_dry, ww_split copy branch,
// set_w_surface).  Test infrastructure only; see oracle.h header.
#include "oracle.h"
namespace orc {

// module_big_step_utilities_em.F:3587 calculate_full (2-D use: mut = mub + mu)
void calculate_full(const Cx &x, A2 rfield, A2 rfieldb, A2 rfieldp) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int i = x.its; i <= itf; i++) rfield(i, j) = rfieldb(i, j) + rfieldp(i, j);
}

// module_big_step_utilities_em.F:26-180
void calc_mu_uv(const Cx &x, A2 mu, A2 mub, A2 muu, A2 muv) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  int itf = ite, jtf = std::min(jte, jde - 1);
  auto mid_u = [&](int i0, int i1) {
    for (int j = jts; j <= jtf; j++)
      for (int i = i0; i <= i1; i++) muu(i, j) = 0.5f * (mu(i, j) + mu(i - 1, j) + mub(i, j) + mub(i - 1, j));
  };
  auto west = [&]() {
    int i = its, im = its;
    if (x.c.periodic_x) im = its - 1;
    for (int j = jts; j <= jtf; j++) muu(i, j) = 0.5f * (mu(i, j) + mu(im, j) + mub(i, j) + mub(im, j));
  };
  auto east = [&]() {
    int i = ite, im = ite - 1;
    if (x.c.periodic_x) im = ite;
    for (int j = jts; j <= jtf; j++) muu(i, j) = 0.5f * (mu(i - 1, j) + mu(im, j) + mub(i - 1, j) + mub(im, j));
  };
  if (its != ids && ite != ide) mid_u(its, itf);
  else if (its == ids && ite != ide) { mid_u(its + 1, itf); west(); }
  else if (its != ids && ite == ide) { mid_u(its, itf - 1); east(); }
  else { mid_u(its + 1, itf - 1); west(); east(); }

  itf = std::min(ite, ide - 1); jtf = jte;
  auto mid_v = [&](int j0, int j1) {
    for (int j = j0; j <= j1; j++)
      for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j) + mu(i, j - 1) + mub(i, j) + mub(i, j - 1));
  };
  auto south = [&]() {
    int j = jts, jm = jts;
    if (x.c.periodic_y) jm = jts - 1;
    for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j) + mu(i, jm) + mub(i, j) + mub(i, jm));
  };
  auto north = [&]() {
    int j = jte, jm = jte - 1;
    if (x.c.periodic_y) jm = jte;
    for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j - 1) + mu(i, jm) + mub(i, j - 1) + mub(i, jm));
  };
  if (jts != jds && jte != jde) mid_v(jts, jtf);
  else if (jts == jds && jte != jde) { mid_v(jts + 1, jtf); south(); }
  else if (jts != jds && jte == jde) { mid_v(jts, jtf - 1); north(); }
  else { mid_v(jts + 1, jtf - 1); south(); north(); }
}

// module_big_step_utilities_em.F:184-321
void calc_mu_uv_1(const Cx &x, A2 mu, A2 muu, A2 muv) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  int itf = ite, jtf = std::min(jte, jde - 1);
  auto mid_u = [&](int i0, int i1) {
    for (int j = jts; j <= jtf; j++)
      for (int i = i0; i <= i1; i++) muu(i, j) = 0.5f * (mu(i, j) + mu(i - 1, j));
  };
  auto west = [&]() {
    int i = its, im = its;
    if (x.c.periodic_x) im = its - 1;
    for (int j = jts; j <= jtf; j++) muu(i, j) = 0.5f * (mu(i, j) + mu(im, j));
  };
  auto east = [&]() {
    int i = ite, im = ite - 1;
    if (x.c.periodic_x) im = ite;
    for (int j = jts; j <= jtf; j++) muu(i, j) = 0.5f * (mu(i - 1, j) + mu(im, j));
  };
  if (its != ids && ite != ide) mid_u(its, itf);
  else if (its == ids && ite != ide) { mid_u(its + 1, itf); west(); }
  else if (its != ids && ite == ide) { mid_u(its, itf - 1); east(); }
  else { mid_u(its + 1, itf - 1); west(); east(); }
  itf = std::min(ite, ide - 1); jtf = jte;
  auto mid_v = [&](int j0, int j1) {
    for (int j = j0; j <= j1; j++)
      for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j) + mu(i, j - 1));
  };
  auto south = [&]() {
    int j = jts, jm = jts;
    if (x.c.periodic_y) jm = jts - 1;
    for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j) + mu(i, jm));
  };
  auto north = [&]() {
    int j = jte, jm = jte - 1;
    if (x.c.periodic_y) jm = jte;
    for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mu(i, j - 1) + mu(i, jm));
  };
  if (jts != jds && jte != jde) mid_v(jts, jtf);
  else if (jts == jds && jte != jde) { mid_v(jts + 1, jtf); south(); }
  else if (jts != jds && jte == jde) { mid_v(jts, jtf - 1); north(); }
  else { mid_v(jts + 1, jtf - 1); south(); north(); }
}

// module_big_step_utilities_em.F:329-399
void couple_momentum(const Cx &x, A2 muu, A3 ru, A3 u, A2 msfu, A2 muv, A3 rv, A3 v, A2 msfv,
                     A2 msfv_inv, A2 mut, A3 rw, A3 w, A2 msft) {
  (void)msfv;
  const int ktf = std::min(x.kte, x.kde - 1);
  int itf = x.ite, jtf = std::min(x.jte, x.jde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++) ru(i, k, j) = u(i, k, j) * (x.c1h(k) * muu(i, j) + x.c2h(k)) / msfu(i, j);
  itf = std::min(x.ite, x.ide - 1); jtf = x.jte;
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++) rv(i, k, j) = v(i, k, j) * (x.c1h(k) * muv(i, j) + x.c2h(k)) * msfv_inv(i, j);
  itf = std::min(x.ite, x.ide - 1); jtf = std::min(x.jte, x.jde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= x.kte; k++)
      for (int i = x.its; i <= itf; i++) rw(i, k, j) = w(i, k, j) * (x.c1f(k) * mut(i, j) + x.c2f(k)) / msft(i, j);
}

// module_big_step_utilities_em.F:640-782
void calc_ww_cp(const Cx &x, A3 u, A3 v, A2 mup, A2 mub, A3 ww) {
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const int jtf = std::min(jte, x.jde - 1), ktf = std::min(kte, x.kde - 1), itf = std::min(ite, x.ide - 1);
  const A1 &c1h = x.c1h, &c2h = x.c2h, &dnw = x.dnw;
  const int ni = ite - its + 3, nj = jte - jts + 3;
  std::vector<float> muu_(ni * nj), muv_(ni * nj), divv_(ni * (kte + 1)), dmdt_(ni);
  auto muu = [&](int i, int j) -> float & { return muu_[(i - its + 1) + (j - jts + 1) * ni]; };
  auto muv = [&](int i, int j) -> float & { return muv_[(i - its + 1) + (j - jts + 1) * ni]; };
  auto divv = [&](int i, int k) -> float & { return divv_[(i - its + 1) + k * ni]; };
  auto dmdt = [&](int i) -> float & { return dmdt_[i - its + 1]; };
  for (int j = jts; j <= jtf; j++)
    for (int i = its; i <= std::min(ite + 1, x.ide); i++)
      muu(i, j) = 0.5f * (mup(i, j) + mub(i, j) + mup(i - 1, j) + mub(i - 1, j));
  for (int j = jts; j <= std::min(jte + 1, x.jde); j++)
    for (int i = its; i <= itf; i++) muv(i, j) = 0.5f * (mup(i, j) + mub(i, j) + mup(i, j - 1) + mub(i, j - 1));
  for (int j = jts; j <= jtf; j++) {
    for (int i = its; i <= ite; i++) {
      dmdt(i) = 0.f;
      ww(i, 1, j) = 0.f;
      ww(i, kte, j) = 0.f;
    }
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= itf; i++) {
        divv(i, k) = x.msftx(i, j) * dnw(k) *
                     (x.rdx * ((c1h(k) * muu(i + 1, j) + c2h(k)) * u(i + 1, k, j) / x.msfuy(i + 1, j) -
                               (c1h(k) * muu(i, j) + c2h(k)) * u(i, k, j) / x.msfuy(i, j)) +
                      x.rdy * ((c1h(k) * muv(i, j + 1) + c2h(k)) * v(i, k, j + 1) * x.msfvx_inv(i, j + 1) -
                               (c1h(k) * muv(i, j) + c2h(k)) * v(i, k, j) * x.msfvx_inv(i, j)));
        dmdt(i) = dmdt(i) + divv(i, k);
      }
    for (int k = 2; k <= ktf; k++)
      for (int i = its; i <= itf; i++) ww(i, k, j) = ww(i, k - 1, j) - dnw(k - 1) * c1h(k - 1) * dmdt(i) - divv(i, k - 1);
  }
}

// module_big_step_utilities_em.F:787-906
void calc_cq(const Cx &x, std::vector<A3> &moist, A3 cqu, A3 cqv, A3 cqw) {
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts;
  const int ktf = std::min(x.kte, x.kde - 1);
  const int n = (int)moist.size();
  std::vector<float> q_(ite - its + 2);
  auto qtot = [&](int i) -> float & { return q_[i - its]; };
  if (n >= 1) {
    int itf = ite, jtf = std::min(jte, x.jde - 1);
    for (int j = jts; j <= jtf; j++)
      for (int k = kts; k <= ktf; k++) {
        std::fill(q_.begin(), q_.end(), 0.f);
        for (int s = 0; s < n; s++)
          for (int i = its; i <= itf; i++) qtot(i) = qtot(i) + moist[s](i, k, j) + moist[s](i - 1, k, j);
        for (int i = its; i <= itf; i++) cqu(i, k, j) = 1.f / (1.f + 0.5f * qtot(i));
      }
    itf = std::min(ite, x.ide - 1); jtf = jte;
    for (int j = jts; j <= jtf; j++)
      for (int k = kts; k <= ktf; k++) {
        std::fill(q_.begin(), q_.end(), 0.f);
        for (int s = 0; s < n; s++)
          for (int i = its; i <= itf; i++) qtot(i) = qtot(i) + moist[s](i, k, j) + moist[s](i, k, j - 1);
        for (int i = its; i <= itf; i++) cqv(i, k, j) = 1.f / (1.f + 0.5f * qtot(i));
      }
    itf = std::min(ite, x.ide - 1); jtf = std::min(jte, x.jde - 1);
    for (int j = jts; j <= jtf; j++)
      for (int k = kts + 1; k <= ktf; k++) {
        std::fill(q_.begin(), q_.end(), 0.f);
        for (int s = 0; s < n; s++)
          for (int i = its; i <= itf; i++) qtot(i) = qtot(i) + moist[s](i, k, j) + moist[s](i, k - 1, j);
        for (int i = its; i <= itf; i++) cqw(i, k, j) = 0.5f * qtot(i);
      }
  } else {
    int itf = ite, jtf = std::min(jte, x.jde - 1);
    for (int j = jts; j <= jtf; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = its; i <= itf; i++) cqu(i, k, j) = 1.f;
    itf = std::min(ite, x.ide - 1); jtf = jte;
    for (int j = jts; j <= jtf; j++)
      for (int k = kts; k <= ktf; k++)
        for (int i = its; i <= itf; i++) cqv(i, k, j) = 1.f;
    itf = std::min(ite, x.ide - 1); jtf = std::min(jte, x.jde - 1);
    for (int j = jts; j <= jtf; j++)
      for (int k = kts + 1; k <= ktf; k++)
        for (int i = its; i <= itf; i++) cqw(i, k, j) = 0.f;
  }
}

// module_big_step_utilities_em.F:910-949
void calc_alt(const Cx &x, A3 alt, A3 al, A3 alb) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1), ktf = std::min(x.kte, x.kde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++) alt(i, k, j) = al(i, k, j) + alb(i, k, j);
}

// module_big_step_utilities_em.F:1227-1266
void calc_php(const Cx &x, A3 php, A3 ph, A3 phb) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1), ktf = std::min(x.kte, x.kde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++)
        php(i, k, j) = 0.5f * (phb(i, k, j) + phb(i, k + 1, j) + ph(i, k, j) + ph(i, k + 1, j));
}

// module_big_step_utilities_em.F:953-1223 (non-hydrostatic branch); vspow =
// frame/libmassv.F elementwise x**p
void calc_p_rho_phi(const Cx &x, std::vector<A3> &moist, A3 al, A3 alb, A2 mu, A2 muts, A3 ph,
                    A3 phb, A3 p, A3 pb, A3 t) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1), ktf = std::min(x.kte, x.kde - 1);
  const A1 &c1 = x.c1h, &c2 = x.c2h, &c3f = x.c3f, &c4f = x.c4f, &c3h = x.c3h, &c4h = x.c4h, &rdnw = x.rdnw;
  const float ptop = x.c.p_top;
  const int n_moist = (int)moist.size();
  if (x.c.hypsometric_opt == 1) {
    for (int j = x.jts; j <= jtf; j++)
      for (int k = x.kts; k <= ktf; k++)
        for (int i = x.its; i <= itf; i++)
          al(i, k, j) = -1.f / (c1(k) * muts(i, j) + c2(k)) *
                        (alb(i, k, j) * (c1(k) * mu(i, j)) + rdnw(k) * (ph(i, k + 1, j) - ph(i, k, j)));
  } else {
    for (int j = x.jts; j <= jtf; j++)
      for (int k = x.kts; k <= ktf; k++)
        for (int i = x.its; i <= itf; i++) {
          float pfu = c3f(k + 1) * muts(i, j) + c4f(k + 1) + ptop;
          float pfd = c3f(k) * muts(i, j) + c4f(k) + ptop;
          float phm = c3h(k) * muts(i, j) + c4h(k) + ptop;
          al(i, k, j) = (ph(i, k + 1, j) - ph(i, k, j) + phb(i, k + 1, j) - phb(i, k, j)) / phm / std::log(pfd / pfu) - alb(i, k, j);
        }
  }
  if (n_moist >= 1) {
    for (int j = x.jts; j <= jtf; j++)
      for (int k = x.kts; k <= ktf; k++)
        for (int i = x.its; i <= itf; i++) {
          float temp;
          if (x.c.use_theta_m == 1) {
            temp = (r_d * (t0 + t(i, k, j))) / (p0 * (al(i, k, j) + alb(i, k, j)));
          } else {
            float qvf = 1.f + rvovrd * moist[0](i, k, j);
            temp = (r_d * (t0 + t(i, k, j)) * qvf) / (p0 * (al(i, k, j) + alb(i, k, j)));
          }
          float pw = std::pow(temp, cpovcv);
          p(i, k, j) = pw * p0 - pb(i, k, j);
        }
  } else {
    for (int j = x.jts; j <= jtf; j++)
      for (int k = x.kts; k <= ktf; k++)
        for (int i = x.its; i <= itf; i++)
          p(i, k, j) = p0 * std::pow((r_d * (t0 + t(i, k, j))) / (p0 * (al(i, k, j) + alb(i, k, j))), cpovcv) - pb(i, k, j);
  }
}

// module_big_step_utilities_em.F:4573-4646
void zero_tend(const Cx &x, A3 tendency) {
  for (int j = x.jts; j <= x.jte; j++)
    for (int k = x.kts; k <= x.kte; k++)
      for (int i = x.its; i <= x.ite; i++) tendency(i, k, j) = 0.f;
}
void zero_tend2d(const Cx &x, A2 tendency) {
  for (int j = x.jts; j <= x.jte; j++)
    for (int i = x.its; i <= x.ite; i++) tendency(i, j) = 0.f;
}

// module_ieva_em.F:49-303, zadvect_implicit = 0 branch (:294-303)
void ww_split(const Cx &x, A3 wwE, A3 wwI, A3 ww) {
  const int i_start = x.its - 1, i_end = std::min(x.ite, x.ide - 1) + 1;
  const int j_start = x.jts - 1, j_end = std::min(x.jte, x.jde - 1) + 1;
  for (int j = j_start; j <= j_end; j++)
    for (int k = x.kts; k <= x.kte; k++)
      for (int i = i_start; i <= i_end; i++) {
        wwE(i, k, j) = ww(i, k, j);
        wwI(i, k, j) = 0.0f;
      }
}

// module_big_step_utilities_em.F:1365-2181, advective_order 5/6 branch (:1768-2063)
// + open lateral boundary terms (:2065-2179)
void rhs_ph(const Cx &x, A3 ph_tend, A3 u, A3 v, A3 ww, A3 ph, A3 ph_old, A3 phb, A3 w, A2 mut,
            A2 muuf, A2 muvf) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kte = x.kte;
  const A1 &c1f = x.c1f, &c2f = x.c2f, &fnm = x.fnm, &fnp = x.fnp, &rdnw = x.rdnw;
  const float rdx = x.rdx, rdy = x.rdy, cfn = x.cfn, cfn1 = x.cfn1;
  const A2 &msfty = x.msfty, &msftx = x.msftx, &msfvy = x.msfvy, &msfux = x.msfux;
  int itf = std::min(ite, ide - 1), jtf = std::min(jte, jde - 1);
  const int ni = ite - its + 2;
  std::vector<float> wdwn_(ni * (kte + 2));
  auto wdwn = [&](int i, int k) -> float & { return wdwn_[(i - its) + k * ni]; };

  for (int j = jts; j <= jtf; j++) {
    if (x.c.phi_adv_z == 2) {
      for (int k = 2; k <= kte; k++)
        for (int i = its; i <= itf; i++)
          wdwn(i, k) = rdnw(k - 1) * (ph(i, k, j) - ph(i, k - 1, j) + phb(i, k, j) - phb(i, k - 1, j));
      for (int k = 2; k <= kte - 1; k++)
        for (int i = its; i <= itf; i++)
          ph_tend(i, k, j) = ph_tend(i, k, j) - ww(i, k, j) * (fnm(k) * wdwn(i, k + 1) + fnp(k) * wdwn(i, k));
    } else {
      for (int k = 2; k <= kte; k++)
        for (int i = its; i <= itf; i++)
          wdwn(i, k) = .5f * (ww(i, k, j) + ww(i, k - 1, j)) * rdnw(k - 1) *
                       (ph(i, k, j) - ph(i, k - 1, j) + phb(i, k, j) - phb(i, k - 1, j));
      for (int k = 2; k <= kte - 1; k++)
        for (int i = its; i <= itf; i++)
          ph_tend(i, k, j) = ph_tend(i, k, j) - (fnm(k) * wdwn(i, k + 1) + fnp(k) * wdwn(i, k));
    }
  }
  // non_hydrostatic: "gw" term
  for (int j = jts; j <= jtf; j++) {
    for (int i = its; i <= itf; i++) ph_tend(i, kde, j) = 0.f;
    for (int k = 2; k <= kte; k++)
      for (int i = its; i <= itf; i++)
        ph_tend(i, k, j) = ph_tend(i, k, j) + (c1f(k) * mut(i, j) + c2f(k)) * g * w(i, k, j) / msfty(i, j);
  }

  // --- 6th-order branch: y
  auto vavg = [&](int i, int k, int j) { return v(i, k, j) + v(i, k - 1, j); };
  auto vtop = [&](int i, int k, int j) { return cfn * v(i, k - 1, j) + cfn1 * v(i, k - 2, j); };
  auto uavg = [&](int i, int k, int j) { return u(i, k, j) + u(i, k - 1, j); };
  auto utop = [&](int i, int k, int j) { return cfn * u(i, k - 1, j) + cfn1 * u(i, k - 2, j); };
  auto y6 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float va = top ? vtop(i, k, j + 1) : vavg(i, k, j + 1);
    float vb = top ? vtop(i, k, j) : vavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdy / msfty(i, j)) *
            (((c1f(k) * muvf(i, j + 1) + c2f(k)) * va * msfvy(i, j + 1) + (c1f(k) * muvf(i, j) + c2f(k)) * vb * msfvy(i, j)) *
             (1.f / 60.f) *
             (45.f * (ph(i, k, j + 1) - ph(i, k, j - 1)) - 9.f * (ph(i, k, j + 2) - ph(i, k, j - 2)) +
              (ph(i, k, j + 3) - ph(i, k, j - 3)) + 45.f * (phb(i, k, j + 1) - phb(i, k, j - 1)) -
              9.f * (phb(i, k, j + 2) - phb(i, k, j - 2)) + (phb(i, k, j + 3) - phb(i, k, j - 3))));
  };
  auto y4 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float va = top ? vtop(i, k, j + 1) : vavg(i, k, j + 1);
    float vb = top ? vtop(i, k, j) : vavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdy / msfty(i, j)) *
            (((c1f(k) * muvf(i, j + 1) + c2f(k)) * va * msfvy(i, j + 1) + (c1f(k) * muvf(i, j) + c2f(k)) * vb * msfvy(i, j)) *
             (1.f / 12.f) *
             (8.f * (ph(i, k, j + 1) - ph(i, k, j - 1)) - (ph(i, k, j + 2) - ph(i, k, j - 2)) +
              8.f * (phb(i, k, j + 1) - phb(i, k, j - 1)) - (phb(i, k, j + 2) - phb(i, k, j - 2))));
  };
  auto y2 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float va = top ? vtop(i, k, j + 1) : vavg(i, k, j + 1);
    float vb = top ? vtop(i, k, j) : vavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdy / msfty(i, j)) *
            ((c1f(k) * muvf(i, j + 1) + c2f(k)) * va * msfvy(i, j + 1) *
                 (phb(i, k, j + 1) - phb(i, k, j) + ph(i, k, j + 1) - ph(i, k, j)) +
             (c1f(k) * muvf(i, j) + c2f(k)) * vb * msfvy(i, j) *
                 (phb(i, k, j) - phb(i, k, j - 1) + ph(i, k, j) - ph(i, k, j - 1)));
  };
  int i_start = its, j_start = jts;
  itf = std::min(ite, ide - 1); jtf = std::min(jte, jde - 1);
  if (x.oys()) j_start = std::max(jts, jds + 3);
  if (x.oye()) jtf = std::min(jtf, jde - 4);
  for (int j = j_start; j <= jtf; j++) {
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) y6(i, k, j, false);
    for (int i = i_start; i <= itf; i++) y6(i, kte, j, true);
  }
  if (x.oys() && jts <= jds + 2 && jds + 2 <= jte) {
    int j = jds + 2;
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) y4(i, k, j, false);
    for (int i = i_start; i <= itf; i++) y4(i, kte, j, true);
  }
  if (x.oye() && jts <= jde - 3 && jde - 3 <= jte) {
    int j = jde - 3;
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) y4(i, k, j, false);
    for (int i = i_start; i <= itf; i++) y4(i, kte, j, true);
  }
  if (x.oys() && jts <= jds + 1 && jds + 1 <= jte) {
    int j = jds + 1;
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) y2(i, k, j, false);
    for (int i = i_start; i <= itf; i++) y2(i, kte, j, true);
  }
  if (x.oye() && jts <= jde - 2 && jde - 2 <= jte) {
    int j = jde - 2;
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) y2(i, k, j, false);
    for (int i = i_start; i <= itf; i++) y2(i, kte, j, true);
  }

  // --- x
  auto x6 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float ua = top ? utop(i + 1, k, j) : uavg(i + 1, k, j);
    float ub = top ? utop(i, k, j) : uavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdx / msfty(i, j)) *
            (((c1f(k) * muuf(i + 1, j) + c2f(k)) * ua * msfux(i + 1, j) + (c1f(k) * muuf(i, j) + c2f(k)) * ub * msfux(i, j)) *
             (1.f / 60.f) *
             (45.f * (ph(i + 1, k, j) - ph(i - 1, k, j)) - 9.f * (ph(i + 2, k, j) - ph(i - 2, k, j)) +
              (ph(i + 3, k, j) - ph(i - 3, k, j)) + 45.f * (phb(i + 1, k, j) - phb(i - 1, k, j)) -
              9.f * (phb(i + 2, k, j) - phb(i - 2, k, j)) + (phb(i + 3, k, j) - phb(i - 3, k, j))));
  };
  auto x4 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float ua = top ? utop(i + 1, k, j) : uavg(i + 1, k, j);
    float ub = top ? utop(i, k, j) : uavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdx / msfty(i, j)) *
            (((c1f(k) * muuf(i + 1, j) + c2f(k)) * ua * msfux(i + 1, j) + (c1f(k) * muuf(i, j) + c2f(k)) * ub * msfux(i, j)) *
             (1.f / 12.f) *
             (8.f * (ph(i + 1, k, j) - ph(i - 1, k, j)) - (ph(i + 2, k, j) - ph(i - 2, k, j)) +
              8.f * (phb(i + 1, k, j) - phb(i - 1, k, j)) - (phb(i + 2, k, j) - phb(i - 2, k, j))));
  };
  auto x2 = [&](int i, int k, int j, bool top) {
    float fac = top ? 0.5f : 0.25f;
    float ua = top ? utop(i + 1, k, j) : uavg(i + 1, k, j);
    float ub = top ? utop(i, k, j) : uavg(i, k, j);
    ph_tend(i, k, j) = ph_tend(i, k, j) -
        (fac * rdx / msfty(i, j)) *
            ((c1f(k) * muuf(i + 1, j) + c2f(k)) * ua * msfux(i + 1, j) *
                 (phb(i + 1, k, j) - phb(i, k, j) + ph(i + 1, k, j) - ph(i, k, j)) +
             (c1f(k) * muuf(i, j) + c2f(k)) * ub * msfux(i, j) *
                 (phb(i, k, j) - phb(i - 1, k, j) + ph(i, k, j) - ph(i - 1, k, j)));
  };
  i_start = its; j_start = jts;
  itf = std::min(ite, ide - 1); jtf = std::min(jte, jde - 1);
  if (x.oxs()) i_start = std::max(its, ids + 3);
  if (x.oxe()) itf = std::min(itf, ide - 4);
  for (int j = j_start; j <= jtf; j++) {
    for (int k = 2; k <= kte - 1; k++)
      for (int i = i_start; i <= itf; i++) x6(i, k, j, false);
    for (int i = i_start; i <= itf; i++) x6(i, kte, j, true);
  }
  if (x.c.open_xs && its <= ids + 2 && ids + 2 <= ite) {
    int i = ids + 2;
    for (int j = j_start; j <= jtf; j++) {
      for (int k = 2; k <= kte - 1; k++) x4(i, k, j, false);
      x4(i, kte, j, true);
    }
  }
  if (x.c.open_xe && its <= ide - 3 && ide - 3 <= ite) {
    int i = ide - 3;
    for (int j = j_start; j <= jtf; j++) {
      for (int k = 2; k <= kte - 1; k++) x4(i, k, j, false);
      x4(i, kte, j, true);
    }
  }
  if (x.oxs() && its <= ids + 1 && ids + 1 <= ite) {
    int i = ids + 1;
    for (int j = j_start; j <= jtf; j++)
      for (int k = 2; k <= kte - 1; k++) x2(i, k, j, false);
    for (int j = j_start; j <= jtf; j++) x2(i, kte, j, true);
  }
  if (x.oxe() && its <= ide - 2 && ide - 2 <= ite) {
    int i = ide - 2;
    for (int j = j_start; j <= jtf; j++)
      for (int k = 2; k <= kte - 1; k++) x2(i, k, j, false);
    for (int j = j_start; j <= jtf; j++) x2(i, kte, j, true);
  }

  // lateral open boundary conditions (:2065-2179)
  itf = std::min(ite, ide - 1);
  if (x.c.open_ys && jts == jds) {
    int j = jts;
    for (int k = 2; k <= kde; k++) {
      int kz = std::min(k, kde - 1);
      for (int i = its; i <= itf; i++) {
        float vb = .5f * (fnm(kz) * (v(i, kz, j + 1) + v(i, kz, j)) + fnp(kz) * (v(i, kz - 1, j + 1) + v(i, kz - 1, j)));
        float vl = std::min(vb, 0.f);
        ph_tend(i, k, j) = ph_tend(i, k, j) - rdy * (c1f(k) * mut(i, j) + c2f(k)) * (+vl * (ph_old(i, k, j + 1) - ph_old(i, k, j)));
      }
    }
  }
  if (x.c.open_ye && jte == jde) {
    int j = jte - 1;
    for (int k = 2; k <= kde; k++) {
      int kz = std::min(k, kde - 1);
      for (int i = its; i <= itf; i++) {
        float vb = .5f * (fnm(kz) * (v(i, kz, j + 1) + v(i, kz, j)) + fnp(kz) * (v(i, kz - 1, j + 1) + v(i, kz - 1, j)));
        float vr = std::max(vb, 0.f);
        ph_tend(i, k, j) = ph_tend(i, k, j) - rdy * (c1f(k) * mut(i, j) + c2f(k)) * (+vr * (ph_old(i, k, j) - ph_old(i, k, j - 1)));
      }
    }
  }
  jtf = std::min(jte, jde - 1);
  if (x.c.open_xs && its == ids) {
    int i = its;
    for (int j = jts; j <= jtf; j++) {
      for (int k = 2; k <= kde; k++) {   // k=2..kde-1 and k=kde both use kz = k (:2130-2146)
        int kz = k;
        float ub = .5f * (fnm(kz) * (u(i + 1, kz, j) + u(i, kz, j)) + fnp(kz) * (u(i + 1, kz - 1, j) + u(i, kz - 1, j)));
        float ul = std::min(ub, 0.f);
        ph_tend(i, k, j) = ph_tend(i, k, j) - (msftx(i, j) / msfty(i, j)) * rdx * (c1f(k) * mut(i, j) + c2f(k)) *
                                                  (+ul * (ph_old(i + 1, k, j) - ph_old(i, k, j)));
      }
    }
  }
  if (x.c.open_xe && ite == ide) {
    int i = ite - 1;
    for (int j = jts; j <= jtf; j++) {
      for (int k = 2; k <= kde; k++) {
        int kz = (k == kde) ? k - 1 : k;   // :2157, :2166
        float ub = .5f * (fnm(kz) * (u(i + 1, kz, j) + u(i, kz, j)) + fnp(kz) * (u(i + 1, kz - 1, j) + u(i, kz - 1, j)));
        float ur = std::max(ub, 0.f);
        ph_tend(i, k, j) = ph_tend(i, k, j) - (msftx(i, j) / msfty(i, j)) * rdx * (c1f(k) * mut(i, j) + c2f(k)) *
                                                  (+ur * (ph_old(i, k, j) - ph_old(i - 1, k, j)));
      }
    }
  }
}

// module_big_step_utilities_em.F:2183-2415 (non-hydrostatic branch)
void horizontal_pressure_gradient(const Cx &x, A3 ru_tend, A3 rv_tend, A3 ph, A3 alt, A3 p, A3 pb,
                                  A3 al, A3 php, A3 cqu, A3 cqv, A2 muu, A2 muv, A2 mu) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  const A1 &c1h = x.c1h, &c2h = x.c2h, &fnm = x.fnm, &fnp = x.fnp, &rdnw = x.rdnw;
  const A1 &c1 = x.c1h;
  const float rdx = x.rdx, rdy = x.rdy, cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3, cfn = x.cfn, cfn1 = x.cfn1;
  const bool top_lid = x.c.top_lid;
  const int ni = ite - its + 2;
  std::vector<float> dpn_(ni * (kde + 1));
  auto dpn = [&](int i, int k) -> float & { return dpn_[(i - its) + k * ni]; };
  int itf = std::min(ite, ide - 1), jtf = jte, ktf = std::min(x.kte, kde - 1);
  int i_start = its, j_start = jts;
  if (x.oys() && jts == jds) j_start = jts + 1;
  if (x.oye() && jte == jde) jtf = jtf - 1;
  for (int j = j_start; j <= jtf; j++) {
    int k = 1;
    for (int i = i_start; i <= itf; i++) {
      dpn(i, k) = .5f * (cf1 * (p(i, k, j - 1) + p(i, k, j)) + cf2 * (p(i, k + 1, j - 1) + p(i, k + 1, j)) +
                         cf3 * (p(i, k + 2, j - 1) + p(i, k + 2, j)));
      dpn(i, kde) = 0.f;
    }
    if (top_lid)
      for (int i = i_start; i <= itf; i++)
        dpn(i, kde) = .5f * (cfn * (p(i, kde - 1, j - 1) + p(i, kde - 1, j)) + cfn1 * (p(i, kde - 2, j - 1) + p(i, kde - 2, j)));
    for (k = 2; k <= ktf; k++)
      for (int i = i_start; i <= itf; i++)
        dpn(i, k) = .5f * (fnm(k) * (p(i, k, j - 1) + p(i, k, j)) + fnp(k) * (p(i, k - 1, j - 1) + p(i, k - 1, j)));
    for (k = 1; k <= ktf; k++)
      for (int i = i_start; i <= itf; i++) {
        float dpy = (x.msfvy(i, j) / x.msfvx(i, j)) * .5f * rdy * (c1h(k) * muv(i, j) + c2h(k)) *
                    ((ph(i, k + 1, j) - ph(i, k + 1, j - 1) + ph(i, k, j) - ph(i, k, j - 1)) +
                     (alt(i, k, j) + alt(i, k, j - 1)) * (p(i, k, j) - p(i, k, j - 1)) +
                     (al(i, k, j) + al(i, k, j - 1)) * (pb(i, k, j) - pb(i, k, j - 1)));
        dpy = dpy + (x.msfvy(i, j) / x.msfvx(i, j)) * rdy * (php(i, k, j) - php(i, k, j - 1)) *
                        (rdnw(k) * (dpn(i, k + 1) - dpn(i, k)) - .5f * ((c1(k) * mu(i, j - 1)) + (c1(k) * mu(i, j))));
        rv_tend(i, k, j) = rv_tend(i, k, j) - cqv(i, k, j) * dpy;
      }
  }
  itf = ite; jtf = std::min(jte, jde - 1);
  i_start = its; j_start = jts;
  if (x.oxs() && its == ids) i_start = its + 1;
  if (x.oxe() && ite == ide) itf = itf - 1;
  if (x.c.periodic_x) { i_start = its; itf = ite; }
  for (int j = j_start; j <= jtf; j++) {
    int k = 1;
    for (int i = i_start; i <= itf; i++) {
      dpn(i, k) = .5f * (cf1 * (p(i - 1, k, j) + p(i, k, j)) + cf2 * (p(i - 1, k + 1, j) + p(i, k + 1, j)) +
                         cf3 * (p(i - 1, k + 2, j) + p(i, k + 2, j)));
      dpn(i, kde) = 0.f;
    }
    if (top_lid)
      for (int i = i_start; i <= itf; i++)
        dpn(i, kde) = .5f * (cfn * (p(i - 1, kde - 1, j) + p(i, kde - 1, j)) + cfn1 * (p(i - 1, kde - 2, j) + p(i, kde - 2, j)));
    for (k = 2; k <= ktf; k++)
      for (int i = i_start; i <= itf; i++)
        dpn(i, k) = .5f * (fnm(k) * (p(i - 1, k, j) + p(i, k, j)) + fnp(k) * (p(i - 1, k - 1, j) + p(i, k - 1, j)));
    for (k = 1; k <= ktf; k++)
      for (int i = i_start; i <= itf; i++) {
        float dpx = (x.msfux(i, j) / x.msfuy(i, j)) * .5f * rdx * (c1h(k) * muu(i, j) + c2h(k)) *
                    ((ph(i, k + 1, j) - ph(i - 1, k + 1, j) + ph(i, k, j) - ph(i - 1, k, j)) +
                     (alt(i, k, j) + alt(i - 1, k, j)) * (p(i, k, j) - p(i - 1, k, j)) +
                     (al(i, k, j) + al(i - 1, k, j)) * (pb(i, k, j) - pb(i - 1, k, j)));
        dpx = dpx + (x.msfux(i, j) / x.msfuy(i, j)) * rdx * (php(i, k, j) - php(i - 1, k, j)) *
                        (rdnw(k) * (dpn(i, k + 1) - dpn(i, k)) - .5f * ((c1(k) * mu(i - 1, j)) + (c1(k) * mu(i, j))));
        ru_tend(i, k, j) = ru_tend(i, k, j) - cqu(i, k, j) * dpx;
      }
  }
}

// module_big_step_utilities_em.F:2419-2499
void pg_buoy_w(const Cx &x, A3 rw_tend, A3 p, A3 cqw, A2 muf, A2 mubf) {
  const int kde = x.kde;
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1);
  const A1 &c1f = x.c1f, &c2f = x.c2f, &rdnw = x.rdnw, &rdn = x.rdn;
  for (int j = x.jts; j <= jtf; j++) {
    int k = kde;
    for (int i = x.its; i <= itf; i++) {
      float cq1 = 1.f / (1.f + cqw(i, k - 1, j));
      float cq2 = cqw(i, k - 1, j) * cq1;
      rw_tend(i, k, j) = rw_tend(i, k, j) + (1.f / x.msfty(i, j)) * g *
                                                (cq1 * 2.f * rdnw(k - 1) * (-p(i, k - 1, j)) - (c1f(k) * muf(i, j)) -
                                                 cq2 * (c1f(k) * mubf(i, j) + c2f(k)));
    }
    for (k = 2; k <= kde - 1; k++)
      for (int i = x.its; i <= itf; i++) {
        float cq1 = 1.f / (1.f + cqw(i, k, j));
        float cq2 = cqw(i, k, j) * cq1;
        cqw(i, k, j) = cq1;
        rw_tend(i, k, j) = rw_tend(i, k, j) + (1.f / x.msfty(i, j)) * g *
                                                  (cq1 * rdn(k) * (p(i, k, j) - p(i, k - 1, j)) - (c1f(k) * muf(i, j)) -
                                                   cq2 * (c1f(k) * mubf(i, j) + c2f(k)));
      }
  }
}

// module_big_step_utilities_em.F:2503-2711 (not polar)
void w_damp(const Cx &x, A3 rw_tend, float &max_vert_cfl, float &max_horiz_cfl, A3 u, A3 v, A3 ww,
            A3 w, A2 mut) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1);
  const A1 &c1f = x.c1f, &c2f = x.c2f, &rdnw = x.rdnw;
  max_vert_cfl = 0.f;
  max_horiz_cfl = 0.f;
  const float w_crit_cfl = x.c.w_crit_cfl;
  const float w_damp_on = x.c.zadvect_implicit > 0 ? w_crit_cfl : w_beta;   // :2602-2608
  const float dt = x.dt;
  for (int j = x.jts; j <= jtf; j++)
    for (int k = 2; k <= x.kde - 1; k++)
      for (int i = x.its; i <= itf; i++) {
        float vert_cfl = std::fabs(ww(i, k, j) / (c1f(k) * mut(i, j) + c2f(k)) * rdnw(k) * dt);
        float msfuxt = x.msfux(i, j);
        if (vert_cfl > max_vert_cfl) max_vert_cfl = vert_cfl;
        float horiz_cfl = std::max(std::fabs(u(i, k, j) * x.rdx * msfuxt * dt), std::fabs(v(i, k, j) * x.rdy * x.msfvy(i, j) * dt));
        if (horiz_cfl > max_horiz_cfl) max_horiz_cfl = horiz_cfl;
        if (vert_cfl > w_damp_on && x.c.w_damping == 1)
          rw_tend(i, k, j) = rw_tend(i, k, j) - sign1(w(i, k, j)) * w_alpha * (vert_cfl - w_crit_cfl) * (c1f(k) * mut(i, j) + c2f(k));
      }
}

// module_big_step_utilities_em.F:3640-3850
void coriolis(const Cx &x, A3 ru, A3 rv, A3 rw, A3 ru_tend, A3 rv_tend, A3 rw_tend) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts;
  const int ktf = std::min(x.kte, x.kde - 1);
  const A2 &f = x.f, &e = x.e, &sina = x.sina, &cosa = x.cosa;
  const A1 &fzm = x.fnm, &fzp = x.fnp;
  int i_start = its, i_end = ite;
  if (x.oxs()) i_start = std::max(ids + 1, its);
  if (x.oxe()) i_end = std::min(ide - 1, ite);
  if (x.c.periodic_x) { i_start = its; i_end = ite; }
  for (int j = jts; j <= std::min(jte, jde - 1); j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        ru_tend(i, k, j) = ru_tend(i, k, j) +
                           (x.msfux(i, j) / x.msfuy(i, j)) * 0.5f * (f(i, j) + f(i - 1, j)) * 0.25f *
                               (rv(i - 1, k, j + 1) + rv(i, k, j + 1) + rv(i - 1, k, j) + rv(i, k, j)) -
                           0.5f * (e(i, j) + e(i - 1, j)) * 0.5f * (cosa(i, j) + cosa(i - 1, j)) * 0.25f *
                               (rw(i - 1, k + 1, j) + rw(i - 1, k, j) + rw(i, k + 1, j) + rw(i, k, j));
  int j_start = jts, j_end = jte;
  if (x.oys()) j_start = std::max(jds + 1, jts);
  if (x.oye()) j_end = std::min(jde - 1, jte);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= std::min(ide - 1, ite); i++)
        rv_tend(i, k, j) = rv_tend(i, k, j) -
                           (x.msfvy(i, j) / x.msfvx(i, j)) * 0.5f * (f(i, j) + f(i, j - 1)) * 0.25f *
                               (ru(i, k, j) + ru(i + 1, k, j) + ru(i, k, j - 1) + ru(i + 1, k, j - 1)) +
                           (x.msfvy(i, j) / x.msfvx(i, j)) * 0.5f * (e(i, j) + e(i, j - 1)) * 0.5f * (sina(i, j) + sina(i, j - 1)) * 0.25f *
                               (rw(i, k + 1, j - 1) + rw(i, k, j - 1) + rw(i, k + 1, j) + rw(i, k, j));
  for (int j = jts; j <= std::min(jte, jde - 1); j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = its; i <= std::min(ite, ide - 1); i++)
        rw_tend(i, k, j) = rw_tend(i, k, j) +
                           e(i, j) * (cosa(i, j) * 0.5f * (fzm(k) * (ru(i, k, j) + ru(i + 1, k, j)) + fzp(k) * (ru(i, k - 1, j) + ru(i + 1, k - 1, j))) -
                                      (x.msftx(i, j) / x.msfty(i, j)) * sina(i, j) * 0.5f *
                                          (fzm(k) * (rv(i, k, j) + rv(i, k, j + 1)) + fzp(k) * (rv(i, k - 1, j) + rv(i, k - 1, j + 1))));
}

// module_big_step_utilities_em.F:3854-4173
void perturbation_coriolis(const Cx &x, A3 ru_in, A3 rv_in, A3 rw, A3 ru_tend, A3 rv_tend,
                           A3 rw_tend, A2 muu, A2 muv, A3 phb, A3 ph) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts;
  const int ktf = std::min(x.kte, x.kde - 1);
  const A2 &f = x.f, &e = x.e, &sina = x.sina, &cosa = x.cosa;
  const A1 &fzm = x.fnm, &fzp = x.fnp, &c1h = x.c1h, &c2h = x.c2h, &z_base = x.z_base, &u_base = x.u_base, &v_base = x.v_base;
  // local ru, rv (its-1:ite+1, kts:kte, jts-1:jte+1)
  const int ni = ite - its + 3, nk = x.kte - kts + 1, nj = jte - jts + 3;
  std::vector<float> ru_((size_t)ni * nk * nj, 0.f), rv_((size_t)ni * nk * nj, 0.f);
  auto ru = [&](int i, int k, int j) -> float & { return ru_[(i - its + 1) + (size_t)(k - kts) * ni + (size_t)(j - jts + 1) * ni * nk]; };
  auto rv = [&](int i, int k, int j) -> float & { return rv_[(i - its + 1) + (size_t)(k - kts) * ni + (size_t)(j - jts + 1) * ni * nk]; };
  int i_start = its, i_end = ite;
  if (x.oxs()) i_start = std::max(ids + 1, its);
  if (x.oxe()) i_end = std::min(ide - 1, ite);
  if (x.c.periodic_x) { i_start = its; i_end = ite; }

  auto zv = [&](int i, int k, int j) {
    return 0.25f * (phb(i, k, j) + phb(i, k + 1, j) + phb(i, k, j - 1) + phb(i, k + 1, j - 1) + ph(i, k, j) + ph(i, k + 1, j) +
                    ph(i, k, j - 1) + ph(i, k + 1, j - 1)) / g;
  };
  auto zu = [&](int i, int k, int j) {
    return 0.25f * (phb(i, k, j) + phb(i, k + 1, j) + phb(i - 1, k, j) + phb(i - 1, k + 1, j) + ph(i, k, j) + ph(i, k + 1, j) +
                    ph(i - 1, k, j) + ph(i - 1, k + 1, j)) / g;
  };
  for (int j = jts; j <= std::min(jte, jde - 1) + 1; j++)
    for (int k = kts + 1; k <= ktf - 1; k++)
      for (int i = i_start - 1; i <= i_end; i++) {
        float z = zv(i, k, j);
        float wkp1 = std::min(1.f, std::max(0.f, z - z_base(k)) / (z_base(k + 1) - z_base(k)));
        float wkm1 = std::min(1.f, std::max(0.f, z_base(k) - z) / (z_base(k) - z_base(k - 1)));
        float wk = 1.f - wkp1 - wkm1;
        rv(i, k, j) = rv_in(i, k, j) - (c1h(k) * muv(i, j) + c2h(k)) * (wkm1 * v_base(k - 1) + wk * v_base(k) + wkp1 * v_base(k + 1));
      }
  for (int j = jts; j <= std::min(jte, jde - 1) + 1; j++)
    for (int i = i_start - 1; i <= i_end; i++) {
      int k = kts;
      float z = zv(i, k, j);
      float wkp1 = std::min(1.f, std::max(0.f, z - z_base(k)) / (z_base(k + 1) - z_base(k)));
      float wk = 1.f - wkp1;
      rv(i, k, j) = rv_in(i, k, j) - (c1h(k) * muv(i, j) + c2h(k)) * (+wk * v_base(k) + wkp1 * v_base(k + 1));
      k = ktf;
      z = zv(i, k, j);
      float wkm1 = std::min(1.f, std::max(0.f, z_base(k) - z) / (z_base(k) - z_base(k - 1)));
      wk = 1.f - wkm1;
      rv(i, k, j) = rv_in(i, k, j) - (c1h(k) * muv(i, j) + c2h(k)) * (wkm1 * v_base(k - 1) + wk * v_base(k));
    }
  for (int j = jts; j <= std::min(jte, jde - 1); j++) {
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        ru_tend(i, k, j) = ru_tend(i, k, j) +
                           (x.msfux(i, j) / x.msfuy(i, j)) * 0.5f * (f(i, j) + f(i - 1, j)) * 0.25f *
                               (rv(i - 1, k, j + 1) + rv(i, k, j + 1) + rv(i - 1, k, j) + rv(i, k, j)) -
                           0.5f * (e(i, j) + e(i - 1, j)) * 0.5f * (cosa(i, j) + cosa(i - 1, j)) * 0.25f *
                               (rw(i - 1, k + 1, j) + rw(i - 1, k, j) + rw(i, k + 1, j) + rw(i, k, j));
    if (x.c.open_xs && its == ids)
      for (int k = kts; k <= ktf; k++)
        ru_tend(its, k, j) = ru_tend(its, k, j) +
                             (x.msfux(its, j) / x.msfuy(its, j)) * 0.5f * (f(its, j) + f(its, j)) * 0.25f *
                                 (rv(its, k, j + 1) + rv(its, k, j + 1) + rv(its, k, j) + rv(its, k, j)) -
                             0.5f * (e(its, j) + e(its, j)) * 0.5f * (cosa(its, j) + cosa(its, j)) * 0.25f *
                                 (rw(its, k + 1, j) + rw(its, k, j) + rw(its, k + 1, j) + rw(its, k, j));
    if (x.c.open_xe && ite == ide)
      for (int k = kts; k <= ktf; k++)
        ru_tend(ite, k, j) = ru_tend(ite, k, j) +
                             (x.msfux(ite, j) / x.msfuy(ite, j)) * 0.5f * (f(ite - 1, j) + f(ite - 1, j)) * 0.25f *
                                 (rv(ite - 1, k, j + 1) + rv(ite - 1, k, j + 1) + rv(ite - 1, k, j) + rv(ite - 1, k, j)) -
                             0.5f * (e(ite - 1, j) + e(ite - 1, j)) * 0.5f * (cosa(ite - 1, j) + cosa(ite - 1, j)) * 0.25f *
                                 (rw(ite - 1, k + 1, j) + rw(ite - 1, k, j) + rw(ite - 1, k + 1, j) + rw(ite - 1, k, j));
  }
  int j_start = jts, j_end = jte;
  if (x.oys()) j_start = std::max(jds + 1, jts);
  if (x.oye()) j_end = std::min(jde - 1, jte);
  for (int j = j_start - 1; j <= j_end; j++)
    for (int k = kts + 1; k <= ktf - 1; k++)
      for (int i = its; i <= std::min(ite, ide - 1) + 1; i++) {
        float z = zu(i, k, j);
        float wkp1 = std::min(1.f, std::max(0.f, z - z_base(k)) / (z_base(k + 1) - z_base(k)));
        float wkm1 = std::min(1.f, std::max(0.f, z_base(k) - z) / (z_base(k) - z_base(k - 1)));
        float wk = 1.f - wkp1 - wkm1;
        ru(i, k, j) = ru_in(i, k, j) - (c1h(k) * muu(i, j) + c2h(k)) * (wkm1 * u_base(k - 1) + wk * u_base(k) + wkp1 * u_base(k + 1));
      }
  for (int j = j_start - 1; j <= j_end; j++)
    for (int i = its; i <= std::min(ite, ide - 1) + 1; i++) {
      int k = kts;
      float z = zu(i, k, j);
      float wkp1 = std::min(1.f, std::max(0.f, z - z_base(k)) / (z_base(k + 1) - z_base(k)));
      float wk = 1.f - wkp1;
      ru(i, k, j) = ru_in(i, k, j) - (c1h(k) * muu(i, j) + c2h(k)) * (+wk * u_base(k) + wkp1 * u_base(k + 1));
      k = ktf;
      z = zu(i, k, j);
      float wkm1 = std::min(1.f, std::max(0.f, z_base(k) - z) / (z_base(k) - z_base(k - 1)));
      wk = 1.f - wkm1;
      ru(i, k, j) = ru_in(i, k, j) - (c1h(k) * muu(i, j) + c2h(k)) * (wkm1 * u_base(k - 1) + wk * u_base(k));
    }
  if (x.c.open_ys && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= std::min(ide - 1, ite); i++)
        rv_tend(i, k, jts) = rv_tend(i, k, jts) -
                             (x.msfvy(i, jts) / x.msfvx(i, jts)) * 0.5f * (f(i, jts) + f(i, jts)) * 0.25f *
                                 (ru(i, k, jts) + ru(i + 1, k, jts) + ru(i, k, jts) + ru(i + 1, k, jts)) +
                             (x.msfvy(i, jts) / x.msfvx(i, jts)) * 0.5f * (e(i, jts) + e(i, jts)) * 0.5f * (sina(i, jts) + sina(i, jts)) * 0.25f *
                                 (rw(i, k + 1, jts) + rw(i, k, jts) + rw(i, k + 1, jts) + rw(i, k, jts));
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= std::min(ide - 1, ite); i++)
        rv_tend(i, k, j) = rv_tend(i, k, j) -
                           (x.msfvy(i, j) / x.msfvx(i, j)) * 0.5f * (f(i, j) + f(i, j - 1)) * 0.25f *
                               (ru(i, k, j) + ru(i + 1, k, j) + ru(i, k, j - 1) + ru(i + 1, k, j - 1)) +
                           (x.msfvy(i, j) / x.msfvx(i, j)) * 0.5f * (e(i, j) + e(i, j - 1)) * 0.5f * (sina(i, j) + sina(i, j - 1)) * 0.25f *
                               (rw(i, k + 1, j - 1) + rw(i, k, j - 1) + rw(i, k + 1, j) + rw(i, k, j));
  if (x.c.open_ye && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= std::min(ide - 1, ite); i++)
        rv_tend(i, k, jte) = rv_tend(i, k, jte) -
                             (x.msfvy(i, jte) / x.msfvx(i, jte)) * 0.5f * (f(i, jte - 1) + f(i, jte - 1)) * 0.25f *
                                 (ru(i, k, jte - 1) + ru(i + 1, k, jte - 1) + ru(i, k, jte - 1) + ru(i + 1, k, jte - 1)) +
                             (x.msfvy(i, jte) / x.msfvx(i, jte)) * 0.5f * (e(i, jte - 1) + e(i, jte - 1)) * 0.5f *
                                 (sina(i, jte - 1) + sina(i, jte - 1)) * 0.25f *
                                 (rw(i, k + 1, jte - 1) + rw(i, k, jte - 1) + rw(i, k + 1, jte - 1) + rw(i, k, jte - 1));
  for (int j = jts; j <= std::min(jte, jde - 1); j++)
    for (int k = kts + 1; k <= ktf; k++)
      for (int i = its; i <= std::min(ite, ide - 1); i++)
        rw_tend(i, k, j) = rw_tend(i, k, j) +
                           e(i, j) * (cosa(i, j) * 0.5f * (fzm(k) * (ru(i, k, j) + ru(i + 1, k, j)) + fzp(k) * (ru(i, k - 1, j) + ru(i + 1, k - 1, j))) -
                                      (x.msftx(i, j) / x.msfty(i, j)) * sina(i, j) * 0.5f *
                                          (fzm(k) * (rv(i, k, j) + rv(i, k, j + 1)) + fzp(k) * (rv(i, k - 1, j) + rv(i, k - 1, j + 1))));
}

// module_big_step_utilities_em.F:4175-4469 (map_proj /= 6, not polar)
void curvature(const Cx &x, A3 ru, A3 rv, A3 rw, A3 u, A3 v, A3 w, A3 ru_tend, A3 rv_tend,
               A3 rw_tend) {
  (void)w;
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts;
  const int ktf = std::min(x.kte, x.kde - 1);
  const A1 &fzm = x.fnm, &fzp = x.fnp;
  const float rdx = x.rdx, rdy = x.rdy;
  const int ni = ite - its + 3, nk = x.kte - kts + 1, nj = jte - jts + 3;
  std::vector<float> vx_((size_t)ni * nk * nj, 0.f);
  auto vxgm = [&](int i, int k, int j) -> float & { return vx_[(i - its + 1) + (size_t)(k - kts) * ni + (size_t)(j - jts + 1) * ni * nk]; };
  int i_start = its - 1, i_end = ite, j_start = jts - 1, j_end = jte;
  if (x.oxs() && its == ids) i_start = its;
  if (x.oxe() && ite == ide) i_end = ite - 1;
  if (x.oys() && jts == jds) j_start = jts;
  if (x.oye() && jte == jde) j_end = jte - 1;
  if (x.c.periodic_x) { i_start = its - 1; i_end = ite; }
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        vxgm(i, k, j) = 0.5f * (u(i, k, j) + u(i + 1, k, j)) * (x.msfvx(i, j + 1) - x.msfvx(i, j)) * rdy -
                        0.5f * (v(i, k, j) + v(i, k, j + 1)) * (x.msfuy(i + 1, j) - x.msfuy(i, j)) * rdx;
  if ((x.c.open_xs || (x.spec && !x.c.periodic_x)) && its == ids)
    for (int j = jts; j <= jte - 1; j++)
      for (int k = kts; k <= ktf; k++) vxgm(its - 1, k, j) = vxgm(its, k, j);
  if ((x.c.open_xe || (x.spec && !x.c.periodic_x)) && ite == ide)
    for (int j = jts; j <= jte - 1; j++)
      for (int k = kts; k <= ktf; k++) vxgm(ite, k, j) = vxgm(ite - 1, k, j);
  if (x.oys() && jts == jds)
    for (int k = kts; k <= ktf; k++)
      for (int i = its - 1; i <= ite; i++) vxgm(i, k, jts - 1) = vxgm(i, k, jts);
  if (x.oye() && jte == jde)
    for (int k = kts; k <= ktf; k++)
      for (int i = its - 1; i <= ite; i++) vxgm(i, k, jte) = vxgm(i, k, jte - 1);
  i_start = its;
  if (x.oxs()) i_start = std::max(ids + 1, its);
  if (x.oxe()) i_end = std::min(ide - 1, ite);
  if (x.c.periodic_x) { i_start = its; i_end = ite; }
  for (int j = jts; j <= std::min(jde - 1, jte); j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = i_start; i <= i_end; i++)
        ru_tend(i, k, j) = ru_tend(i, k, j) +
                           0.5f * (vxgm(i, k, j) + vxgm(i - 1, k, j)) * 0.25f *
                               (rv(i - 1, k, j + 1) + rv(i, k, j + 1) + rv(i - 1, k, j) + rv(i, k, j)) -
                           u(i, k, j) * reradius * 0.25f * (rw(i - 1, k + 1, j) + rw(i - 1, k, j) + rw(i, k + 1, j) + rw(i, k, j));
  j_start = jts;
  if (x.oys()) j_start = std::max(jds + 1, jts);
  if (x.oye()) j_end = std::min(jde - 1, jte);
  for (int j = j_start; j <= j_end; j++)
    for (int k = kts; k <= ktf; k++)
      for (int i = its; i <= std::min(ite, ide - 1); i++)
        rv_tend(i, k, j) = rv_tend(i, k, j) -
                           0.5f * (vxgm(i, k, j) + vxgm(i, k, j - 1)) * 0.25f *
                               (ru(i, k, j) + ru(i + 1, k, j) + ru(i, k, j - 1) + ru(i + 1, k, j - 1)) -
                           (x.msfvy(i, j) / x.msfvx(i, j)) * v(i, k, j) * reradius * 0.25f *
                               (rw(i, k + 1, j - 1) + rw(i, k, j - 1) + rw(i, k + 1, j) + rw(i, k, j));
  for (int j = jts; j <= std::min(jte, jde - 1); j++)
    for (int k = std::max(2, kts); k <= ktf; k++)
      for (int i = its; i <= std::min(ite, ide - 1); i++)
        rw_tend(i, k, j) = rw_tend(i, k, j) +
                           reradius * (0.5f * (fzm(k) * (ru(i, k, j) + ru(i + 1, k, j)) + fzp(k) * (ru(i, k - 1, j) + ru(i + 1, k - 1, j))) * 0.5f *
                                           (fzm(k) * (u(i, k, j) + u(i + 1, k, j)) + fzp(k) * (u(i, k - 1, j) + u(i + 1, k - 1, j))) +
                                       (x.msftx(i, j) / x.msfty(i, j)) * 0.5f *
                                           (fzm(k) * (rv(i, k, j) + rv(i, k, j + 1)) + fzp(k) * (rv(i, k - 1, j) + rv(i, k - 1, j + 1))) * 0.5f *
                                           (fzm(k) * (v(i, k, j) + v(i, k, j + 1)) + fzp(k) * (v(i, k - 1, j) + v(i, k - 1, j + 1))));
}

// module_big_step_utilities_em.F:5836-6110
void rk_rayleigh_damp(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 rw_tendf, A3 t_tendf, A3 u, A3 v,
                      A3 w, A3 t, A2 mut, A2 muu, A2 muv, A3 ph, A3 phb) {
  const int kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const A1 &c1h = x.c1h, &c2h = x.c2h, &c1f = x.c1f, &c2f = x.c2f;
  const A1 &u_base = x.u_base, &v_base = x.v_base, &t_base = x.t_base, &z_base = x.z_base;
  const float dampcoef = x.c.dampcoef, zdamp = x.c.zdamp;
  const float pii = 2.0f * std::asin(1.0f);
  const int ktf = std::min(kte, kde - 1);
  std::vector<float> z00_(kte + 2), b00_(kte + 2);
  auto z00 = [&](int k) -> float & { return z00_[k]; };
  auto b00 = [&](int k) -> float & { return b00_[k]; };
  auto interp = [&](const A1 &base, int k1) {
    for (int k = k1; k <= ktf; k++) {
      int k2 = ktf;
      while (z_base(k2) > z00(k)) k2 = k2 - 1;
      if (k2 + 1 > ktf)
        b00(k) = base(k2) + (base(k2) - base(k2 - 1)) * (z00(k) - z_base(k2)) / (z_base(k2) - z_base(k2 - 1));
      else
        b00(k) = base(k2) + (base(k2 + 1) - base(k2)) * (z00(k) - z_base(k2)) / (z_base(k2 + 1) - z_base(k2));
    }
  };
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int i = its; i <= std::min(ite, x.ide); i++) {
      float ztop = 0.5f * (phb(i, kde, j) + phb(i - 1, kde, j) + ph(i, kde, j) + ph(i - 1, kde, j)) / g;
      int k1 = ktf;
      float z = ztop;
      while (z >= (ztop - zdamp)) {
        z = 0.25f * (phb(i, k1, j) + phb(i, k1 + 1, j) + phb(i - 1, k1, j) + phb(i - 1, k1 + 1, j) + ph(i, k1, j) + ph(i, k1 + 1, j) +
                     ph(i - 1, k1, j) + ph(i - 1, k1 + 1, j)) / g;
        z00(k1) = z;
        k1 = k1 - 1;
      }
      k1 = k1 + 2;
      interp(u_base, k1);
      for (int k = k1; k <= ktf; k++) {
        float dcoef = 1.0f - std::min(1.0f, (ztop - z00(k)) / zdamp);
        dcoef = std::sin(0.5f * pii * dcoef) * std::sin(0.5f * pii * dcoef);
        ru_tendf(i, k, j) = ru_tendf(i, k, j) - (c1h(k) * muu(i, j) + c2h(k)) * (dcoef * dampcoef) * (u(i, k, j) - b00(k));
      }
    }
  for (int j = jts; j <= std::min(jte, x.jde); j++)
    for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
      float ztop = 0.5f * (phb(i, kde, j) + phb(i, kde, j - 1) + ph(i, kde, j) + ph(i, kde, j - 1)) / g;
      int k1 = ktf;
      float z = ztop;
      while (z >= (ztop - zdamp)) {
        z = 0.25f * (phb(i, k1, j) + phb(i, k1 + 1, j) + phb(i, k1, j - 1) + phb(i, k1 + 1, j - 1) + ph(i, k1, j) + ph(i, k1 + 1, j) +
                     ph(i, k1, j - 1) + ph(i, k1 + 1, j - 1)) / g;
        z00(k1) = z;
        k1 = k1 - 1;
      }
      k1 = k1 + 2;
      interp(v_base, k1);
      for (int k = k1; k <= ktf; k++) {
        float dcoef = 1.0f - std::min(1.0f, (ztop - z00(k)) / zdamp);
        dcoef = std::sin(0.5f * pii * dcoef) * std::sin(0.5f * pii * dcoef);
        rv_tendf(i, k, j) = rv_tendf(i, k, j) - (c1h(k) * muv(i, j) + c2h(k)) * (dcoef * dampcoef) * (v(i, k, j) - b00(k));
      }
    }
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
      float ztop = (phb(i, kde, j) + ph(i, kde, j)) / g;
      for (int k = kts; k <= std::min(kte, kde); k++) {
        float z = (phb(i, k, j) + ph(i, k, j)) / g;
        if (z >= (ztop - zdamp)) {
          float dcoef = 1.0f - std::min(1.0f, (ztop - z) / zdamp);
          dcoef = std::sin(0.5f * pii * dcoef) * std::sin(0.5f * pii * dcoef);
          rw_tendf(i, k, j) = rw_tendf(i, k, j) - (c1f(k) * mut(i, j) + c2f(k)) * (dcoef * dampcoef) * w(i, k, j);
        }
      }
    }
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
      float ztop = (phb(i, kde, j) + ph(i, kde, j)) / g;
      int k1 = ktf;
      float z = ztop;
      while (z >= (ztop - zdamp)) {
        z = 0.5f * (phb(i, k1, j) + phb(i, k1 + 1, j) + ph(i, k1, j) + ph(i, k1 + 1, j)) / g;
        z00(k1) = z;
        k1 = k1 - 1;
      }
      k1 = k1 + 2;
      interp(t_base, k1);
      for (int k = k1; k <= ktf; k++) {
        float dcoef = 1.0f - std::min(1.0f, (ztop - z00(k)) / zdamp);
        dcoef = std::sin(0.5f * pii * dcoef) * std::sin(0.5f * pii * dcoef);
        t_tendf(i, k, j) = t_tendf(i, k, j) - (c1h(k) * mut(i, j) + c2h(k)) * (dcoef * dampcoef) * (t(i, k, j) - b00(k));
      }
    }
}

// dyn_em/module_em.F:959-1092
void rk_addtend_dry(const Cx &x, A3 ru_tend, A3 rv_tend, A3 rw_tend, A3 ph_tend, A3 t_tend,
                    A3 ru_tendf, A3 rv_tendf, A3 rw_tendf, A3 ph_tendf, A3 t_tendf, A3 u_save,
                    A3 v_save, A3 w_save, A3 ph_save, A3 t_save, A2 mu_tend, A2 mu_tendf,
                    int rk_step, A3 h_diabatic, A2 mut) {
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const A1 &c1 = x.c1h, &c2 = x.c2h;
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int k = kts; k <= kte - 1; k++)
      for (int i = its; i <= ite; i++) {
        if (rk_step == 1) ru_tendf(i, k, j) = ru_tendf(i, k, j) + u_save(i, k, j) * x.msfuy(i, j);
        ru_tend(i, k, j) = ru_tend(i, k, j) + ru_tendf(i, k, j) / x.msfuy(i, j);
      }
  for (int j = jts; j <= jte; j++)
    for (int k = kts; k <= kte - 1; k++)
      for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
        if (rk_step == 1) rv_tendf(i, k, j) = rv_tendf(i, k, j) + v_save(i, k, j) * x.msfvx(i, j);
        rv_tend(i, k, j) = rv_tend(i, k, j) + rv_tendf(i, k, j) * x.msfvx_inv(i, j);
      }
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
        if (rk_step == 1) rw_tendf(i, k, j) = rw_tendf(i, k, j) + w_save(i, k, j) * x.msfty(i, j);
        rw_tend(i, k, j) = rw_tend(i, k, j) + rw_tendf(i, k, j) / x.msfty(i, j);
        if (rk_step == 1) ph_tendf(i, k, j) = ph_tendf(i, k, j) + ph_save(i, k, j);
        ph_tend(i, k, j) = ph_tend(i, k, j) + ph_tendf(i, k, j) / x.msfty(i, j);
      }
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int k = kts; k <= kte - 1; k++)
      for (int i = its; i <= std::min(ite, x.ide - 1); i++) {
        if (rk_step == 1) t_tendf(i, k, j) = t_tendf(i, k, j) + t_save(i, k, j);
        t_tend(i, k, j) = t_tend(i, k, j) + t_tendf(i, k, j) / x.msfty(i, j) +
                          (c1(k) * mut(i, j) + c2(k)) * h_diabatic(i, k, j) / x.msfty(i, j);
      }
  for (int j = jts; j <= std::min(jte, x.jde - 1); j++)
    for (int i = its; i <= std::min(ite, x.ide - 1); i++) mu_tend(i, j) = mu_tend(i, j) + mu_tendf(i, j);
}

// dyn_em/module_bc_em.F:1196-1270 (fill_w_flag = .false.)
void set_w_surface(const Cx &x, A3 w, A3 u, A3 v) {
  const A2 &ht = x.ht;
  const float cf1 = x.cf1, cf2 = x.cf2, cf3 = x.cf3;
  int jm1_limit = x.jds, jp1_limit = x.jde - 1, im1_limit = x.ids, ip1_limit = x.ide - 1;
  if (x.c.periodic_x) { im1_limit = x.ids - 1; ip1_limit = x.ide; }
  if (x.c.periodic_y) { jm1_limit = x.jds - 1; jp1_limit = x.jde; }
  for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++) {
    int jm1 = std::max(j - 1, jm1_limit), jp1 = std::min(j + 1, jp1_limit);
    for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) {
      int im1 = std::max(i - 1, im1_limit), ip1 = std::min(i + 1, ip1_limit);
      w(i, 1, j) = x.msfty(i, j) * .5f * x.rdy *
                       ((ht(i, jp1) - ht(i, j)) * (cf1 * v(i, 1, j + 1) + cf2 * v(i, 2, j + 1) + cf3 * v(i, 3, j + 1)) +
                        (ht(i, j) - ht(i, jm1)) * (cf1 * v(i, 1, j) + cf2 * v(i, 2, j) + cf3 * v(i, 3, j))) +
                   x.msftx(i, j) * .5f * x.rdx *
                       ((ht(ip1, j) - ht(i, j)) * (cf1 * u(i + 1, 1, j) + cf2 * u(i + 1, 2, j) + cf3 * u(i + 1, 3, j)) +
                        (ht(i, j) - ht(im1, j)) * (cf1 * u(i, 1, j) + cf2 * u(i, 2, j) + cf3 * u(i, 3, j)));
    }
  }
}

}  // namespace orc
