// solve_em.cpp -- oracle restatement of the sequencing in dyn_em/solve_em.F
// (:578-4848) and dyn_em/module_em.F (rk_step_prep :37, rk_tendency :190,
// rk_scalar_tend :1096, rk_update_scalar :1587, rk_update_scalar_pd :1803), plus
// the C entry points the tests bind with ctypes.
// Test infrastructure only; see oracle.h header.
#include "oracle.h"
#include <cstdio>
namespace orc {

static const char *k3d[] = {
    "u_1", "u_2", "v_1", "v_2", "w_1", "w_2", "t_1", "t_2", "ph_1", "ph_2",
    "phb", "pb", "alb", "p", "al", "alt", "php", "ru", "rv", "rw", "ww",
    "ru_m", "rv_m", "ww_m", "ru_tend", "rv_tend", "u_save", "v_save", "t_save", "h_diabatic",
    "ru_tendf", "rv_tendf", "rw_tend", "rw_tendf", "w_save", "ww1", "wwE", "wwI",
    "ph_tend", "ph_tendf", "ph_save", "t_tend", "t_tendf", "t_2save", "advect_tend",
    "alpha", "a", "gamma", "c2a", "cqu", "cqv", "cqw", "pm1", "t_init",
    "rho", "th_phy", "pi_phy", "p_phy", "z", "z_at_w", "dz8w", "qv_diabatic", "qc_diabatic", nullptr};
static const char *k2d[] = {"mu_1", "mu_2", "mub", "mudf", "muu", "muus", "muv", "muvs", "mut", "muts",
                            "muave", "mu_save", "mu_tend", "mu_tendf",
                            "mut_old", "mut_new", "muu_old", "muu_new", "muv_old", "muv_new",   // IEVA locals of rk_tendency (module_em.F:343-347)
                            "msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy",
                            "f", "e", "sina", "cosa", "ht", "rainnc", "rainncv", nullptr};
static const char *k1d[] = {"znu", "znw", "fnm", "fnp", "rdnw", "rdn", "dnw", "dn",
                            "c1h", "c2h", "c1f", "c2f", "c3h", "c4h", "c3f", "c4f",
                            "u_base", "v_base", "t_base", "qv_base", "z_base", nullptr};
static const char *moist_names[] = {"qv", "qc", "qr"};

Oracle::Oracle(const wrfgpu_dims &d_, const wrfgpu_config &c_) : d(d_), c(c_) {
  n1 = d.kme - d.kms + 1;
  n2 = (long)(d.ime - d.ims + 1) * (d.jme - d.jms + 1);
  n3 = n2 * n1;
  for (int i = 0; k3d[i]; i++) add(k3d[i], 3);
  for (int i = 0; k2d[i]; i++) add(k2d[i], 2);
  for (int i = 0; k1d[i]; i++) add(k1d[i], 1);
  for (int s = 0; s < c.num_moist; s++) {
    add(std::string("moist_") + moist_names[s], 3);
    add(std::string("moist_old_") + moist_names[s], 3);
    add(std::string("moist_tend_") + moist_names[s], 3);
  }
  for (int s = 1; s <= nsc(); s++) {
    add("scalar_" + std::to_string(s), 3);
    add("scalar_old_" + std::to_string(s), 3);
    add("scalar_tend_" + std::to_string(s), 3);
  }
  if (c.diff_opt != 0 && (c.km_opt >= 2 || (c.diff_opt == 2 && !diff2_is_identity(c)))) {   // metrics, deformation, K arrays (Registry.EM_COMMON state)
    for (const char *n : {"ustm", "hfx", "qfx", "p8w_sfc", "t8w_sfc"}) add(n, 2);
    for (const char *n : {"bn2", "t_phy"}) add(n, 3);
    for (const char *n : {"p8w_top", "t8w_top"}) add(n, 2);
    // (the prognostic tke rides as the last species of the scalar family, added by the nsc() loop above: tke_2 / tke_1 / tke_tend)
    for (const char *n : {"zx", "zy", "rdz", "rdzw", "div", "defor11", "defor22", "defor33", "defor12", "defor13", "defor23",
                          "xkmh", "xkmv", "xkhh", "xkhv"}) add(n, 3);
  }
  if (spec()) {   // boundary arrays of the specified / nested zone (Registry.EM_COMMON 'b' entries), fcx / gcx (lbc_fcx_gcx)
    if (c.spec_bdy_width < 1 || c.spec_zone < 0 || c.relax_zone < c.spec_zone || c.relax_zone + 1 > c.spec_bdy_width)
      throw std::runtime_error("oracle: spec_bdy_width / spec_zone / relax_zone");
    const size_t W = c.spec_bdy_width, ni = d.ime - d.ims + 1, nj = d.jme - d.jms + 1, nk = d.kde - d.kds + 1;
    add_raw("fcx", W); add_raw("gcx", W);
    auto add_bdy = [&](const std::string &f, size_t levels) {
      for (const char *t : {"_b", "_bt"}) {
        add_raw(f + t + "xs", nj * levels * W); add_raw(f + t + "xe", nj * levels * W);
        add_raw(f + t + "ys", ni * levels * W); add_raw(f + t + "ye", ni * levels * W);
      }
    };
    for (const char *f : {"u", "v", "w", "ph", "t"}) add_bdy(f, nk);
    add_bdy("mu", 1);
    for (int s = 0; s < c.num_moist; s++)
      if (s == 0 || c.nested || c.have_bcs_moist) add_bdy(std::string("moist_") + moist_names[s], nk);
    if (c.nested || c.have_bcs_scalar)
      for (int s = 1; s <= c.num_scalar; s++) add_bdy("scalar_" + std::to_string(s), nk);
  }
  // harmless defaults so that an un-set map factor does not divide by zero
  for (const char *m : {"msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy"})
    std::fill(data[m].begin(), data[m].end(), 1.f);
  if (c.time_step_sound == 0) {
    float spacing = std::min(c.dx, c.dy);
    num_sound_steps = std::max(2 * ((int)(300.f * c.dt / spacing - 0.01f) + 1), 4);
  } else {
    num_sound_steps = c.time_step_sound;
  }
  dts = c.dt / (float)num_sound_steps;
  refresh_cx();
}

void Oracle::add(const std::string &n, int k) {
  names.push_back(n);
  kind[n] = k;
  data[n].assign(k == 3 ? n3 : (k == 2 ? n2 : n1), 0.f);
}
void Oracle::add_raw(const std::string &n, size_t count) {
  names.push_back(n);
  kind[n] = 4;
  data[n].assign(count, 0.f);
}
Bdy Oracle::bdy(const std::string &f) {
  Bdy B;
  static const char *sfx[4] = {"ys", "ye", "xs", "xe"};
  for (int q = 0; q < 4; q++) {
    B.v[q] = data.at(f + "_b" + sfx[q]).data();
    B.t[q] = data.at(f + "_bt" + sfx[q]).data();
  }
  B.ims = d.ims; B.jms = d.jms; B.ni = d.ime - d.ims + 1; B.nj = d.jme - d.jms + 1;
  B.nk = f == "mu" ? 1 : d.kde - d.kds + 1;
  return B;
}
A3 Oracle::a3(const std::string &n) {
  auto it = data.find(n);
  if (it == data.end() || kind[n] != 3) throw std::runtime_error("oracle: no 3-D field " + n);
  A3 a; a.p = it->second.data(); a.ims = d.ims; a.kms = d.kms; a.jms = d.jms;
  a.sk = d.ime - d.ims + 1; a.sj = a.sk * (d.kme - d.kms + 1);
  return a;
}
A2 Oracle::a2(const std::string &n) {
  auto it = data.find(n);
  if (it == data.end() || kind[n] != 2) throw std::runtime_error("oracle: no 2-D field " + n);
  A2 a; a.p = it->second.data(); a.ims = d.ims; a.jms = d.jms; a.sj = d.ime - d.ims + 1;
  return a;
}
A1 Oracle::a1(const std::string &n) {
  auto it = data.find(n);
  if (it == data.end() || kind[n] != 1) throw std::runtime_error("oracle: no 1-D field " + n);
  A1 a; a.p = it->second.data(); a.kms = d.kms;
  return a;
}

void Oracle::refresh_cx() {
  x.ids = d.ids; x.ide = d.ide; x.jds = d.jds; x.jde = d.jde; x.kds = d.kds; x.kde = d.kde;
  x.ims = d.ims; x.ime = d.ime; x.jms = d.jms; x.jme = d.jme; x.kms = d.kms; x.kme = d.kme;
  x.ips = d.ips; x.ipe = d.ipe; x.jps = d.jps; x.jpe = d.jpe;
  x.its = d.ips; x.ite = d.ipe; x.jts = d.jps; x.jte = d.jpe; x.kts = d.kds; x.kte = d.kde;
  x.c = c;
  x.rdx = 1.f / c.dx; x.rdy = 1.f / c.dy; x.dt = c.dt;
  x.cf1 = c.cf1; x.cf2 = c.cf2; x.cf3 = c.cf3; x.cfn = c.cfn; x.cfn1 = c.cfn1;
  x.time_step = c.time_step;
  x.spec = c.specified || c.nested;
  x.znu = a1("znu"); x.znw = a1("znw"); x.fnm = a1("fnm"); x.fnp = a1("fnp");
  x.rdnw = a1("rdnw"); x.rdn = a1("rdn"); x.dnw = a1("dnw"); x.dn = a1("dn");
  x.c1h = a1("c1h"); x.c2h = a1("c2h"); x.c1f = a1("c1f"); x.c2f = a1("c2f");
  x.c3h = a1("c3h"); x.c4h = a1("c4h"); x.c3f = a1("c3f"); x.c4f = a1("c4f");
  x.u_base = a1("u_base"); x.v_base = a1("v_base"); x.t_base = a1("t_base");
  x.qv_base = a1("qv_base"); x.z_base = a1("z_base");
  x.msftx = a2("msftx"); x.msfty = a2("msfty"); x.msfux = a2("msfux"); x.msfuy = a2("msfuy");
  x.msfvx = a2("msfvx"); x.msfvx_inv = a2("msfvx_inv"); x.msfvy = a2("msfvy");
  x.f = a2("f"); x.e = a2("e"); x.sina = a2("sina"); x.cosa = a2("cosa"); x.ht = a2("ht");
}

void Oracle::bc(const std::string &name, char var) {
  if (kind.at(name) == 3) set_physical_bc3d(x, a3(name), var);
  else set_physical_bc2d(x, a2(name), var);
}

// solve_em.F:585-615
void Oracle::rk_params(int rk_step, float &dt_rk, float &dts_rk, int &nsmall) const {
  const int rk_order = c.rk_ord;
  if (rk_order == 2) {
    if (rk_step == 1) { dt_rk = 0.5f * c.dt; dts_rk = dts; nsmall = num_sound_steps / 2; }
    else { dt_rk = c.dt; dts_rk = dts; nsmall = num_sound_steps; }
  } else {
    if (rk_step == 1) { dt_rk = c.dt / 3.f; dts_rk = dt_rk; nsmall = 1; }
    else if (rk_step == 2) { dt_rk = 0.5f * c.dt; dts_rk = dts; nsmall = num_sound_steps / 2; }
    else { dt_rk = c.dt; dts_rk = dts; nsmall = num_sound_steps; }
  }
}

static std::vector<A3> moist_list(Oracle &o, const char *prefix) {
  std::vector<A3> v;
  for (int s = 0; s < o.c.num_moist; s++) v.push_back(o.a3(std::string(prefix) + moist_names[s]));
  return v;
}

// module_em.F:37-186
void Oracle::rk_step_prep(int) {
  par([&](const Cx &xt) { calculate_full(xt, a2("mut"), a2("mub"), a2("mu_2")); });
  par([&](const Cx &xt) { calc_mu_uv(xt, a2("mu_2"), a2("mub"), a2("muu"), a2("muv")); });
  par([&](const Cx &xt) { couple_momentum(xt, a2("muu"), a3("ru"), a3("u_2"), x.msfuy, a2("muv"), a3("rv"), a3("v_2"), x.msfvx, x.msfvx_inv,
                  a2("mut"), a3("rw"), a3("w_2"), x.msfty); });
  par([&](const Cx &xt) { calc_ww_cp(xt, a3("u_2"), a3("v_2"), a2("mu_2"), a2("mub"), a3("ww")); });
  auto m = moist_list(*this, "moist_");
  par([&](const Cx &xt) { calc_cq(xt, m, a3("cqu"), a3("cqv"), a3("cqw")); });
  par([&](const Cx &xt) { calc_alt(xt, a3("alt"), a3("al"), a3("alb")); });
  par([&](const Cx &xt) { calc_php(xt, a3("php"), a3("ph_2"), a3("phb")); });
}

// solve_em.F:726-754 (+ rk_phys_bc_dry_1, module_bc_em.F:960-1041)
void Oracle::rk_phys_bc_dry_1() {
  bc("ru", 'u'); bc("rv", 'v'); bc("rw", 'w'); bc("ww", 'w'); bc("php", 'w'); bc("alt", 't'); bc("p", 'p');
  bc("muu", 'u'); bc("muv", 'v'); bc("mut", 't');
  bc("rho", 'p'); bc("al", 'p'); bc("ph_2", 'w');
}

// module_em.F:1920 init_zero_tendency (called from first_rk_step_part1)
void Oracle::init_zero_tendency() {
  for (const char *n : {"ru_tendf", "rv_tendf", "rw_tendf", "ph_tendf", "t_tendf"}) zero_tend(x, a3(n));
  par([&](const Cx &xt) { zero_tend2d(xt, a2("mu_tendf")); });
  for (int s = 0; s < c.num_moist; s++) zero_tend(x, a3(std::string("moist_tend_") + moist_names[s]));
  for (int s = 1; s <= nsc(); s++) zero_tend(x, a3("scalar_tend_" + std::to_string(s)));
}

// first_rk_step_part2 (dyn_em/module_first_rk_step_part2.F:427-680, :758-760): metrics of the physical-space
// diffusion, deformation and divergence, the eddy viscosities and their boundary fills.  Run for the closures that
// compute K (km_opt = 4 here); with km_opt = 1 every K element is a constant the diffusion routines take directly.
void Oracle::diffusion_prep() {
  if (c.diff_opt == 0) return;
  if (c.diff_opt == 1 && c.km_opt == 1) return;   // every K element is a constant the diffusion routines of rk_tendency take directly
  if (diff2_is_identity(c)) return;               // diff_opt = 2 with all coefficients zero adds exact zeros (shipped em_hill2d_x)
  if (c.km_opt == 2 && c.diff_opt != 2) throw std::runtime_error("oracle: km_opt 2 (TKE) needs diff_opt 2 (tke_rhs is only called there)");
  A3 ph = a3("ph_2"), phb = a3("phb"), z = a3("z"), rdz = a3("rdz"), rdzw = a3("rdzw"), zx = a3("zx"), zy = a3("zy");
  A3 xkmh = a3("xkmh"), xkmv = a3("xkmv"), xkhh = a3("xkhh"), xkhv = a3("xkhv"), rho = a3("rho");
  if (c.diff_opt == 2) {   // phy_prep (first_rk_step_part1): the density the stresses are built with
    if (c.num_moist >= 1) { A3 qv = a3("moist_qv"); par([&](const Cx &xt) { phy_prep_rho(xt, rho, a3("alt"), &qv); }); }
    else par([&](const Cx &xt) { phy_prep_rho(xt, rho, a3("alt"), nullptr); });
  }
  par([&](const Cx &xt) { compute_diff_metrics(xt, ph, phb, z, rdz, rdzw, zx, zy); });
  halo("HALO_EM_TKE_C");   // + PERIOD_BDY_EM_A1, :446-447
  if (c.diff_opt == 2) { halo("HALO_EM_TKE_C_SFC"); bc("ustm", 't'); }   // ustm of HALO_EM_TKE_C, set_physical_bc2d :492
  bc("rdzw", 'w'); bc("rdz", 'w'); bc("z", 'w'); bc("zx", 'e'); bc("zy", 'f');
  par([&](const Cx &xt) {
    cal_deform_and_div(xt, a3("u_2"), a3("v_2"), a3("w_2"), a3("div"), a3("defor11"), a3("defor22"), a3("defor33"), a3("defor12"),
                       a3("defor13"), a3("defor23"), rdz, rdzw, zx, zy);
  });
  halo("HALO_EM_TKE_D");   // :571
  if (c.km_opt == 1) {     // isotropic_km, module_diffusion_em.F:1718-1772
    const float prandtl = 1.f / 3.0f;
    const float khdif3 = c.khdif / prandtl, kvdif3 = c.kvdif / prandtl;
    for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
      for (int k = x.kts; k <= x.kte; k++)
        for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) {
          xkmh(i, k, j) = c.khdif; xkmv(i, k, j) = c.kvdif; xkhh(i, k, j) = khdif3; xkhv(i, k, j) = kvdif3;
        }
  } else if (c.km_opt == 3 || c.km_opt == 2) {   // calculate_N2 (on phy_prep's th_phy, t_phy, p_phy, p8w, t8w), then smag_km or tke_km
    A3 qv, qc;
    if (c.num_moist >= 1) qv = a3("moist_qv");
    if (c.num_moist >= 2) qc = a3("moist_qc");
    const A3 *pqv = c.num_moist >= 1 ? &qv : nullptr, *pqc = c.num_moist >= 2 ? &qc : nullptr;
    par([&](const Cx &xt) {
      phy_prep_thermo(xt, a3("t_2"), pqv, a3("p"), a3("pb"), ph, phb, a3("th_phy"), a3("p_phy"), a3("t_phy"), a2("p8w_sfc"), a2("t8w_sfc"));
    });
    par([&](const Cx &xt) { calculate_N2(xt, a3("bn2"), pqv, pqc, a3("th_phy"), a3("t_phy"), a3("p_phy"), a2("p8w_sfc"), a2("t8w_sfc"), rdz, rdzw); });
    if (c.km_opt == 3) {
      par([&](const Cx &xt) {
        smag_km(xt, xkmh, xkmv, xkhh, xkhv, a3("bn2"), a3("defor11"), a3("defor22"), a3("defor33"), a3("defor12"), a3("defor13"), a3("defor23"), rdzw);
      });
    } else {
      par([&](const Cx &xt) { phy_prep_top(xt, ph, phb, a3("p_phy"), a3("t_phy"), a2("p8w_top"), a2("t8w_top")); });
      par([&](const Cx &xt) {
        tke_km(xt, xkmh, xkmv, xkhh, xkhv, a3("bn2"), a3("scalar_" + std::to_string(tke_index())), a2("p8w_sfc"), a2("t8w_sfc"), a2("p8w_top"),
               a2("t8w_top"), a3("th_phy"), rdz, rdzw);
      });
    }
  } else {
    par([&](const Cx &xt) { smag2d_km(xt, xkmh, xkmv, xkhh, xkhv, a3("defor11"), a3("defor22"), a3("defor12"), rdzw, zx, zy); });
  }
  halo("HALO_EM_TKE_E");   // + PERIOD_BDY_EM_PHY_BC, HALO_EM_PHYS_DIFFUSION, :644-650, :759
  bc("xkmh", 't'); bc("xkhh", 't');   // phy_bc, module_diffusion_em.F:5901
  if (c.diff_opt != 2) return;
  bc("xkmv", 't'); bc("xkhv", 't'); bc("div", 't'); bc("defor11", 't'); bc("defor22", 't'); bc("defor33", 't');
  bc("defor12", 'd'); bc("defor13", 'e'); bc("defor23", 'f'); bc("rho", 't');
  // vertical diffusion first and then horizontal (module_first_rk_step_part2.F:1019-1122; bl_pbl_physics = 0)
  A3 u = a3("u_2"), v = a3("v_2"), t = a3("t_2");
  A3 ru_tendf = a3("ru_tendf"), rv_tendf = a3("rv_tendf"), rw_tendf = a3("rw_tendf"), t_tendf = a3("t_tendf");
  const std::string tk = tke_index() ? std::to_string(tke_index()) : std::string();
  if (tke_index()) {   // tke_rhs (:904-940), before the diffusion
    A3 qv; if (c.num_moist >= 1) qv = a3("moist_qv");
    par([&](const Cx &xt) {
      tke_rhs(xt, a3("scalar_tend_" + tk), a3("bn2"), a3("defor11"), a3("defor22"), a3("defor33"), a3("defor12"), a3("defor13"), a3("defor23"), u, v,
              a3("scalar_" + tk), a2("mut"), a3("th_phy"), xkmh, xkmv, xkhv, rdzw, a2("hfx"), c.num_moist >= 1 ? &qv : nullptr, a2("ustm"), rho);
    });
  }
  par([&](const Cx &xt) { vertical_diffusion_u_2(xt, ru_tendf, a3("defor13"), xkmv, rho); });
  par([&](const Cx &xt) { vertical_diffusion_v_2(xt, rv_tendf, a3("defor23"), xkmv, rho); });
  par([&](const Cx &xt) { vertical_diffusion_w_2(xt, rw_tendf, a3("defor33"), xkmh, rho); });
  par([&](const Cx &xt) { surface_drag(xt, ru_tendf, rv_tendf, u, v, rho, a2("ustm")); });
  const A1 *tb = c.mix_full_fields ? nullptr : &x.t_base, *qb = c.mix_full_fields ? nullptr : &x.qv_base;
  par([&](const Cx &xt) { vertical_diffusion_s(xt, t_tendf, t, tb, xkhv, rdz, rho); });
  {
    A3 qv; if (c.num_moist >= 1) qv = a3("moist_qv");
    par([&](const Cx &xt) { surface_heat_flux(xt, t_tendf, t, c.num_moist >= 1 ? &qv : nullptr, rho, a2("hfx"), a2("qfx")); });
  }
  for (int sp = 0; sp < c.num_moist; sp++) {
    A3 q = a3(std::string("moist_") + moist_names[sp]), qt = a3(std::string("moist_tend_") + moist_names[sp]);
    par([&](const Cx &xt) { vertical_diffusion_s(xt, qt, q, sp == 0 ? qb : nullptr, xkhv, rdz, rho); });
    if (sp == 0) par([&](const Cx &xt) { surface_moisture_flux(xt, qt, a2("qfx")); });
  }
  for (int sp = 1; sp <= nsc(); sp++) {
    const bool is_tke = sp == tke_index();   // tke: K = xkmv, doing_tke (:4302-4308)
    par([&](const Cx &xt) { vertical_diffusion_s(xt, a3("scalar_tend_" + std::to_string(sp)), a3("scalar_" + std::to_string(sp)), nullptr, is_tke ? xkmv : xkhv, rdz, rho, is_tke); });
  }
  par([&](const Cx &xt) { horizontal_diffusion_u_2(xt, ru_tendf, a3("defor11"), a3("defor12"), xkmh, zx, zy, rdzw, rho); });
  par([&](const Cx &xt) { horizontal_diffusion_v_2(xt, rv_tendf, a3("defor12"), a3("defor22"), xkmh, zx, zy, rdzw, rho); });
  par([&](const Cx &xt) { horizontal_diffusion_w_2(xt, rw_tendf, a3("defor13"), a3("defor23"), xkmv, zx, zy, rdz, rho); });
  par([&](const Cx &xt) { horizontal_diffusion_s(xt, t_tendf, t, xkhh, zx, zy, rdzw, rho); });
  for (int sp = 0; sp < c.num_moist; sp++)
    par([&](const Cx &xt) { horizontal_diffusion_s(xt, a3(std::string("moist_tend_") + moist_names[sp]), a3(std::string("moist_") + moist_names[sp]), xkhh, zx, zy, rdzw, rho); });
  for (int sp = 1; sp <= nsc(); sp++) {
    const bool is_tke = sp == tke_index();   // tke: K = xkmh, doing_tke (:2968-2982)
    par([&](const Cx &xt) { horizontal_diffusion_s(xt, a3("scalar_tend_" + std::to_string(sp)), a3("scalar_" + std::to_string(sp)), is_tke ? xkmh : xkhh, zx, zy, rdzw, rho, is_tke); });
  }
}

// module_em.F:190-955 rk_tendency followed by rk_addtend_dry (solve_em.F:881-1007)
void Oracle::rk_tendency(int rk_step) {
  A3 ru_tend = a3("ru_tend"), rv_tend = a3("rv_tend"), rw_tend = a3("rw_tend"), ph_tend = a3("ph_tend"), t_tend = a3("t_tend");
  A3 u = a3("u_2"), v = a3("v_2"), w = a3("w_2"), t = a3("t_2"), ph = a3("ph_2");
  A3 ru = a3("ru"), rv = a3("rv"), rw = a3("rw"), ww = a3("ww"), wwE = a3("wwE"), wwI = a3("wwI");
  A2 mut = a2("mut"), muu = a2("muu"), muv = a2("muv"), mu = a2("mu_2"), mub = a2("mub");
  for (A3 z : {ru_tend, rv_tend, rw_tend, t_tend, ph_tend, a3("u_save"), a3("v_save"), a3("w_save"), a3("ph_save"), a3("t_save")})
    par([&](const Cx &xt) { zero_tend(xt, z); });
  par([&](const Cx &xt) { zero_tend2d(xt, a2("mu_tend")); });
  par([&](const Cx &xt) { zero_tend2d(xt, a2("mu_save")); });
  // IEVA (zadvect_implicit > 0) on the last RK substep, CHK_IEVA module_ieva_em.F:13-43
  const bool ieva = c.zadvect_implicit > 0 && rk_step == c.rk_ord;
  const float dt_step = c.dt / (float)(c.rk_ord - rk_step + 1);
  A2 mut_old = a2("mut_old"), mut_new = a2("mut_new");
  if (ieva) {   // module_em.F:438-480
    par([&](const Cx &xt) { ww_split_ieva(xt, wwE, wwI, u, v, ww, mut, c.dt); });
    par([&](const Cx &xt) { calc_mut_old(xt, mut_old, mub, a2("mu_1")); });
    par([&](const Cx &xt) { calc_mut_new(xt, u, v, mut_old, muu, muv, mut_new, c.dt); });
    par([&](const Cx &xt) { calc_mu_uv_1(xt, mut_old, a2("muu_old"), a2("muv_old")); });
    par([&](const Cx &xt) { calc_mu_uv_1(xt, mut_new, a2("muu_new"), a2("muv_new")); });
  } else {
    par([&](const Cx &xt) { ww_split(xt, wwE, wwI, ww); });
  }
  // momentum_adv_opt = 3 (WENO_MOM, Registry.EM_COMMON:3273): advect_weno_u/v/w on the last RK substep (module_em.F:491-600)
  const bool weno_m = c.momentum_adv_opt == 3 && rk_step == c.rk_ord;
  par([&](const Cx &xt) { advect_u(xt, u, u, ru_tend, ru, rv, wwE, mut, weno_m); });
  if (ieva) par([&](const Cx &xt) { advect_u_implicit(xt, a3("u_1"), ru_tend, wwI, a2("muu_old"), a2("muu_new"), dt_step); });
  par([&](const Cx &xt) { advect_v(xt, v, v, rv_tend, ru, rv, wwE, mut, weno_m); });
  if (ieva) par([&](const Cx &xt) { advect_v_implicit(xt, a3("v_1"), rv_tend, wwI, a2("muv_old"), a2("muv_new"), dt_step); });
  par([&](const Cx &xt) { advect_w(xt, w, w, rw_tend, ru, rv, wwE, mut, weno_m); });
  // theta goes with the scalars' WENO option on the last RK substep (module_em.F:603-621)
  const bool weno_t = (c.scalar_adv_opt == 3 || c.scalar_adv_opt == 4 || c.moist_adv_opt == 3 || c.moist_adv_opt == 4) && rk_step == c.rk_ord;
  if (weno_t) par([&](const Cx &xt) { advect_scalar_weno(xt, t, t, t_tend, ru, rv, wwE); });
  else par([&](const Cx &xt) { advect_scalar(xt, t, t, t_tend, ru, rv, wwE, mut); });
  if (ieva) par([&](const Cx &xt) { advect_s_implicit(xt, a3("t_1"), t_tend, wwI, mut_old, mut_new, dt_step); });
  par([&](const Cx &xt) { rhs_ph(xt, ph_tend, u, v, wwE, ph, ph, a3("phb"), w, mut, muu, muv); });
  if (ieva) {   // module_em.F:682-715
    par([&](const Cx &xt) { advect_ph_implicit(xt, a3("ph_1"), ph_tend, a3("phb"), wwI, mut, dt_step); });
    par([&](const Cx &xt) { advect_w_implicit(xt, a3("w_1"), rw_tend, ru_tend, rv_tend, wwI, ph, a3("ph_1"), ph_tend, mut_old, mut, mut_new, dt_step); });
  }
  par([&](const Cx &xt) { horizontal_pressure_gradient(xt, ru_tend, rv_tend, ph, a3("alt"), a3("p"), a3("pb"), a3("al"), a3("php"), a3("cqu"), a3("cqv"), muu, muv, mu); });
  par([&](const Cx &xt) { pg_buoy_w(xt, rw_tend, a3("p"), a3("cqw"), mu, mub); });
  float mv, mh;
  w_damp(x, rw_tend, mv, mh, u, v, ww, w, mut);
  if (c.use_adaptive_time_step) {   // solve_em.F:914-922
    if (mh > max_horiz_cfl) max_horiz_cfl = mh;
    if (mv > max_vert_cfl) max_vert_cfl = mv;
  }
  if (c.pert_coriolis) par([&](const Cx &xt) { perturbation_coriolis(xt, ru, rv, rw, ru_tend, rv_tend, rw_tend, muu, muv, a3("phb"), ph); });
  else par([&](const Cx &xt) { coriolis(xt, ru, rv, rw, ru_tend, rv_tend, rw_tend); });
  par([&](const Cx &xt) { curvature(xt, ru, rv, rw, u, v, w, ru_tend, rv_tend, rw_tend); });
  if (rk_step == 1 && c.diff_opt == 1) {   // forward_step: diff_opt1, module_em.F:802-878 (bl_pbl_physics = 0)
    // xkmhd = grid%xkmh = khdif and xkhh = khdif/prandtl everywhere (isotropic_km, module_diffusion_em.F:1718-1772)
    const float prandtl = 1.f / 3.0f;
    const float kvdif = c.kvdif;
    const XK xkmh = c.km_opt == 1 ? XK(c.khdif) : XK(a3("xkmh")), xkhh = c.km_opt == 1 ? XK(c.khdif / prandtl) : XK(a3("xkhh"));
    A3 alt = a3("alt");
    par([&](const Cx &xt) { horizontal_diffusion(xt, 'u', u, a3("ru_tendf"), mut, xt.c1h, xt.c2h, xkmh); });
    par([&](const Cx &xt) { horizontal_diffusion(xt, 'v', v, a3("rv_tendf"), mut, xt.c1h, xt.c2h, xkmh); });
    par([&](const Cx &xt) { horizontal_diffusion(xt, 'w', w, a3("rw_tendf"), mut, xt.c1f, xt.c2f, xkmh); });
    par([&](const Cx &xt) { horizontal_diffusion_3dmp(xt, t, a3("t_tendf"), mut, xt.c1h, xt.c2h, a3("t_init"), xkhh); });
    par([&](const Cx &xt) { vertical_diffusion_u(xt, u, a3("ru_tendf"), xt.u_base, xt.c1h, xt.c2h, alt, muu, kvdif); });
    par([&](const Cx &xt) { vertical_diffusion_v(xt, v, a3("rv_tendf"), xt.v_base, xt.c1h, xt.c2h, alt, muv, kvdif); });
    par([&](const Cx &xt) { vertical_diffusion(xt, 'w', w, a3("rw_tendf"), xt.c1f, xt.c2f, alt, mut, kvdif); });
    const float kvdq = 3.f * kvdif;
    par([&](const Cx &xt) { vertical_diffusion_3dmp(xt, t, a3("t_tendf"), a3("t_init"), xt.c1h, xt.c2h, alt, mut, kvdq); });
  }
  if (rk_step == 1 && c.diff_6th_opt != 0) {   // module_em.F:880-930
    A3 phb = a3("phb");
    par([&](const Cx &xt) { sixth_order_diffusion(xt, 'u', u, a3("ru_tendf"), mut, c.dt, xt.c1h, xt.c2h, phb); });
    par([&](const Cx &xt) { sixth_order_diffusion(xt, 'v', v, a3("rv_tendf"), mut, c.dt, xt.c1h, xt.c2h, phb); });
    par([&](const Cx &xt) { sixth_order_diffusion(xt, 'w', w, a3("rw_tendf"), mut, c.dt, xt.c1f, xt.c2f, phb); });
    par([&](const Cx &xt) { sixth_order_diffusion(xt, 'm', t, a3("t_tendf"), mut, c.dt, xt.c1h, xt.c2h, phb); });
  }
  if (rk_step == 1 && c.damp_opt == 2)
    par([&](const Cx &xt) { rk_rayleigh_damp(xt, a3("ru_tendf"), a3("rv_tendf"), a3("rw_tendf"), a3("t_tendf"), u, v, w, t, mut, muu, muv, ph, a3("phb")); });
  const A1x fcx{spec() ? data.at("fcx").data() : nullptr}, gcx{spec() ? data.at("gcx").data() : nullptr};
  if (spec() && rk_step == 1) {   // solve_em.F:937-968: the relaxation tendencies land in the *_save arrays zeroed above
    const Bdy bu = bdy("u"), bv = bdy("v"), bph = bdy("ph"), bt = bdy("t"), bw = bdy("w"), bmu = bdy("mu");
    par([&](const Cx &xt) {
      relax_bdy_dry(xt, a3("u_save"), a3("v_save"), a3("ph_save"), a3("t_save"), a3("w_save"), a2("mu_tend"), ru, rv, ph, t, w, mu, mut, bu, bv, bph,
                    bt, bw, bmu, dtbc, fcx, gcx);
    });
  }
  par([&](const Cx &xt) { rk_addtend_dry(xt, ru_tend, rv_tend, rw_tend, ph_tend, t_tend, a3("ru_tendf"), a3("rv_tendf"), a3("rw_tendf"), a3("ph_tendf"),
                 a3("t_tendf"), a3("u_save"), a3("v_save"), a3("w_save"), a3("ph_save"), a3("t_save"), a2("mu_tend"), a2("mu_tendf"),
                 rk_step, a3("h_diabatic"), mut); });
  if (spec()) {   // solve_em.F:982-1007
    const Bdy bu = bdy("u"), bv = bdy("v"), bph = bdy("ph"), bt = bdy("t"), bw = bdy("w"), bmu = bdy("mu");
    par([&](const Cx &xt) { spec_bdy_dry(xt, ru_tend, rv_tend, ph_tend, t_tend, rw_tend, a2("mu_tend"), bu, bv, bph, bt, bw, bmu); });
  }
}

// solve_em.F:1089-1263
void Oracle::small_step_prep(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { orc::small_step_prep(xt, a3("u_1"), a3("u_2"), a3("v_1"), a3("v_2"), a3("w_1"), a3("w_2"), a3("t_1"), a3("t_2"), a3("ph_1"), a3("ph_2"),
                       a2("mub"), a2("mu_1"), a2("mu_2"), a2("muu"), a2("muus"), a2("muv"), a2("muvs"), a2("mut"), a2("muts"), a2("mudf"),
                       a3("u_save"), a3("v_save"), a3("w_save"), a3("t_save"), a3("ph_save"), a2("mu_save"), a3("ww"), a3("ww1"),
                       a3("c2a"), a3("pb"), a3("p"), a3("alt"), rk_step); });
  par([&](const Cx &xt) { orc::calc_p_rho(xt, a3("al"), a3("p"), a3("ph_2"), a3("alt"), a3("t_2"), a3("t_save"), a3("c2a"), a3("pm1"), a2("mu_2"), a2("muts"),
                  c.smdiv, 0); });
  par([&](const Cx &xt) { calc_coef_w(xt, a3("a"), a3("alpha"), a3("gamma"), a2("mut"), a3("cqw"), a3("c2a"), dts_rk, c.epssm); });
  halo("HALO_EM_B");   // + PERIOD_BDY_EM_B, solve_em.F:1176-1177
  bc("ru_tend", 'u'); bc("rv_tend", 'v'); bc("ph_2", 'w'); bc("al", 'p'); bc("p", 'p'); bc("t_1", 'p'); bc("t_save", 't');
  bc("mu_1", 't'); bc("mu_2", 't'); bc("mudf", 't');
}

void Oracle::advance_uv(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { orc::advance_uv(xt, a3("u_2"), a3("ru_tend"), a3("v_2"), a3("rv_tend"), a3("p"), a3("pb"), a3("ph_2"), a3("php"), a3("alt"), a3("al"),
                  a2("mu_2"), a2("muu"), a3("cqu"), a2("muv"), a3("cqv"), a2("mudf"), dts_rk); });
  if (spec()) {   // solve_em.F:1344-1365
    par([&](const Cx &xt) { spec_bdyupdate(xt, a3("u_2"), a3("ru_tend"), dts_rk, 'u'); });
    par([&](const Cx &xt) { spec_bdyupdate(xt, a3("v_2"), a3("rv_tend"), dts_rk, 'v'); });
  }
}
void Oracle::advance_mu_t(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { orc::advance_mu_t(xt, a3("ww"), a3("ww1"), a3("u_2"), a3("u_save"), a3("v_2"), a3("v_save"), a2("mu_2"), a2("mut"), a2("muave"),
                    a2("muts"), a2("muu"), a2("muv"), a2("mudf"), a3("t_2"), a3("t_save"), a3("t_2save"), a3("t_tend"), a2("mu_tend"),
                    dts_rk, c.epssm); });
  if (spec()) {   // solve_em.F:1459-1490
    par([&](const Cx &xt) { spec_bdyupdate(xt, a3("t_2"), a3("t_tend"), dts_rk, 't'); });
    par([&](const Cx &xt) { spec_bdyupdate(xt, a2("mu_2"), a2("mu_tend"), dts_rk); });
    par([&](const Cx &xt) { spec_bdyupdate(xt, a2("muts"), a2("mu_tend"), dts_rk); });
  }
}
void Oracle::advance_w(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { orc::advance_w(xt, a3("w_2"), a3("rw_tend"), a3("ww"), a3("w_save"), a3("u_2"), a3("v_2"), a2("mu_2"), a2("mut"), a2("muave"), a2("muts"),
                 a3("t_2save"), a3("t_2"), a3("t_save"), a3("ph_2"), a3("ph_save"), a3("phb"), a3("ph_tend"), a3("c2a"), a3("cqw"),
                 a3("alt"), a3("a"), a3("alpha"), a3("gamma"), dts_rk, c.epssm); });
}
void Oracle::sumflux(int rk_step, int iteration) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { orc::sumflux(xt, a3("u_2"), a3("v_2"), a3("ww"), a3("u_save"), a3("v_save"), a3("ww1"), a2("muu"), a2("muv"), a3("ru_m"), a3("rv_m"),
               a3("ww_m"), iteration, ns); });
  if (spec()) {   // solve_em.F:1582-1620 (non_hydrostatic)
    par([&](const Cx &xt) { spec_bdyupdate_ph(xt, a3("ph_save"), a3("ph_2"), a3("ph_tend"), a2("mu_tend"), a2("muts"), xt.c1f, xt.c2f, dts_rk); });
    if (c.specified) par([&](const Cx &xt) { zero_grad_bdy(xt, a3("w_2"), 'w'); });
    else par([&](const Cx &xt) { spec_bdyupdate(xt, a3("w_2"), a3("rw_tend"), dts_rk, 'h'); });
  }
}
void Oracle::calc_p_rho(int iteration) {
  par([&](const Cx &xt) { orc::calc_p_rho(xt, a3("al"), a3("p"), a3("ph_2"), a3("alt"), a3("t_2"), a3("t_save"), a3("c2a"), a3("pm1"), a2("mu_2"), a2("muts"),
                  c.smdiv, iteration); });
  halo("HALO_EM_C2");  // + PERIOD_BDY_EM_B3, solve_em.F:1659-1660
  bc("ph_2", 'w'); bc("al", 'p'); bc("p", 'p'); bc("muts", 't'); bc("mu_2", 't'); bc("mudf", 't');
}
// solve_em.F:1732-1800
void Oracle::small_step_finish(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  par([&](const Cx &xt) { calc_mu_uv_1(xt, a2("muts"), a2("muus"), a2("muvs")); });
  par([&](const Cx &xt) { orc::small_step_finish(xt, a3("u_2"), a3("u_1"), a3("v_2"), a3("v_1"), a3("w_2"), a3("w_1"), a3("t_2"), a3("t_1"), a3("ph_2"), a3("ph_1"),
                         a3("ww"), a3("ww1"), a2("mu_2"), a2("mu_1"), a2("mut"), a2("muts"), a2("muu"), a2("muus"), a2("muv"), a2("muvs"),
                         a3("u_save"), a3("v_save"), a3("w_save"), a3("t_save"), a3("ph_save"), a2("mu_save"), a3("h_diabatic"), ns,
                         dts_rk, rk_step, c.rk_ord); });
  if (rk_step == c.rk_ord) {
    bc("ru_m", 'u'); bc("rv_m", 'v'); bc("ww_m", 'w'); bc("mut", 't'); bc("muts", 't');
  }
}

// module_em.F:1803-1916
static void rk_update_scalar_pd(const Cx &x, A3 scalar, A3 sc_tend, A2 mu_old, A2 mu_new, A2 mu_base, float dt) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1), ktf = std::min(x.kte, x.kde - 1);
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++) {
        float tendency = 0.f;
        tendency = tendency + sc_tend(i, k, j);
        sc_tend(i, k, j) = 0.f;
        float muold = mu_old(i, j) + mu_base(i, j), munew = mu_new(i, j) + mu_base(i, j);
        scalar(i, k, j) = ((x.c1h(k) * muold + x.c2h(k)) * scalar(i, k, j) + dt * tendency) / (x.c1h(k) * munew + x.c2h(k));
      }
}
// module_em.F:1587-1799
static void rk_update_scalar(const Cx &x, A3 scalar_1, A3 scalar_2, A3 sc_tend, A3 advect_tend, A2 mu_old, A2 mu_new, A2 mu_base,
                             int rk_step, float dt) {
  const int itf = std::min(x.ite, x.ide - 1), jtf = std::min(x.jte, x.jde - 1), ktf = std::min(x.kte, x.kde - 1);
  // the advective tendency is not used in the specified zone (:1671-1678); sc_tend is (the *_spc ranges)
  int i_start = x.its, i_end = itf, j_start = x.jts, j_end = jtf;
  if (x.spec) {
    const int sz = x.c.spec_zone;
    if (!x.c.periodic_x) { i_start = std::max(x.its, x.ids + sz); i_end = std::min(x.ite, x.ide - sz - 1); }
    j_start = std::max(x.jts, x.jds + sz); j_end = std::min(x.jte, x.jde - sz - 1);
  }
  for (int j = x.jts; j <= jtf; j++)
    for (int k = x.kts; k <= ktf; k++)
      for (int i = x.its; i <= itf; i++) {
        float tendency = 0.f;
        if (i >= i_start && i <= i_end && j >= j_start && j <= j_end) tendency = advect_tend(i, k, j) * x.msfty(i, j);
        tendency = tendency + sc_tend(i, k, j);
        float muold = mu_old(i, j) + mu_base(i, j), munew = mu_new(i, j) + mu_base(i, j);
        if (rk_step == 1) scalar_1(i, k, j) = scalar_2(i, k, j);
        scalar_2(i, k, j) = ((x.c1h(k) * muold + x.c2h(k)) * scalar_1(i, k, j) + dt * tendency) / (x.c1h(k) * munew + x.c2h(k));
      }
}

// solve_em.F:1850-2955 (moist + generic scalar loops)
void Oracle::rk_scalar_step(int rk_step) {
  float dt_rk, dts_rk; int ns;
  rk_params(rk_step, dt_rk, dts_rk, ns);
  const int rk_order = c.rk_ord;
  struct Sp { std::string f, o, t; int adv_opt; bool moist; };
  std::vector<Sp> sp;
  for (int s = 0; s < c.num_moist; s++)
    sp.push_back({std::string("moist_") + moist_names[s], std::string("moist_old_") + moist_names[s],
                  std::string("moist_tend_") + moist_names[s], c.moist_adv_opt, true});
  for (int s = 1; s <= nsc(); s++)
    sp.push_back({"scalar_" + std::to_string(s), "scalar_old_" + std::to_string(s), "scalar_tend_" + std::to_string(s),
                  s == tke_index() ? c.tke_adv_opt : c.scalar_adv_opt, false});
  if (rk_step == rk_order)
    for (auto &s : sp)
      if (s.adv_opt == 1 || s.adv_opt == 2 || s.adv_opt == 4) par([&](const Cx &xt) { rk_update_scalar_pd(xt, a3(s.o), a3(s.t), a2("mu_1"), a2("mu_1"), a2("mub"), dt_rk); });
  if (rk_step == rk_order) {
    bool any = false;
    for (auto &s : sp) any |= (s.adv_opt == 1 || s.adv_opt == 2 || s.adv_opt == 4);   // /= ORIGINAL and /= WENO_SCALAR, solve_em.F:1850-1977
    if (any) halo("HALO_EM_MOIST_OLD_E_7");   // + PERIOD_BDY_EM_MOIST_OLD (and the SCALAR_OLD twins), solve_em.F:1878-2142
    for (auto &s : sp)
      if (s.adv_opt == 1 || s.adv_opt == 2 || s.adv_opt == 4) bc(s.o, 'p');
  }
  halo("HALO_EM_D");   // + PERIOD_EM_DA, solve_em.F:2179-2181
  if (c.zadvect_implicit > 0 && rk_step == rk_order && !sp.empty()) {
    // WW_SPLIT below reads u_1, v_1 one point beyond the tile.  The Registry exchanges the _1 fields only at the end of a
    // step (HALO_EM_D3_5 / E_5), i.e. before small_step_prep refreshes their interior, so under DM the reference reads
    // one-step-old halo values here and its IEVA results depend on the decomposition.  Deliberate deviation: exchange
    // them (and refresh the physical boundary zone) so that N ranks equal 1 rank; open-boundary cases are unaffected.
    halo("HALO_EM_IEVA");
    bc("u_1", 'U'); bc("v_1", 'V');
  }
  for (auto &s : sp) {
    // rk_scalar_tend, module_em.F:1096-1442
    const bool ieva = c.zadvect_implicit > 0 && rk_step == rk_order;   // module_em.F:1216-1243
    if (ieva) {
      par([&](const Cx &xt) { ww_split_ieva(xt, a3("wwE"), a3("wwI"), a3("u_1"), a3("v_1"), a3("ww_m"), a2("muts"), c.dt); });
      par([&](const Cx &xt) { calc_mut_old(xt, a2("mut_old"), a2("mub"), a2("mu_1")); });
    } else {
      par([&](const Cx &xt) { ww_split(xt, a3("wwE"), a3("wwI"), a3("ww_m")); });
    }
    par([&](const Cx &xt) { zero_tend(xt, a3("advect_tend")); });
    if (rk_step == rk_order && (s.adv_opt == 1 || s.adv_opt == 4))   // POSITIVEDEF :1266-1280, WENOPD_SCALAR :1314-1327
      par([&](const Cx &xt) { advect_scalar_pd(xt, a3(s.f), a3(s.o), a3("advect_tend"), a3("ru_m"), a3("rv_m"), a3("wwE"), a2("muts"), a2("mub"), a2("mu_1"), dt_rk, s.adv_opt == 4); });
    else if (rk_step == rk_order && s.adv_opt == 3)   // WENO_SCALAR, module_em.F:1300-1312
      par([&](const Cx &xt) { advect_scalar_weno(xt, a3(s.f), a3(s.f), a3("advect_tend"), a3("ru_m"), a3("rv_m"), a3("wwE")); });
    else if (rk_step == rk_order && s.adv_opt == 2)   // MONOTONIC, module_em.F:1282-1298
      par([&](const Cx &xt) { advect_scalar_mono(xt, a3(s.f), a3(s.o), a3("advect_tend"), a3("ru_m"), a3("rv_m"), a3("wwE"), a3("wwI"), a2("muts"), a2("mub"), a2("mu_1"), dt_rk); });
    else
      par([&](const Cx &xt) { advect_scalar(xt, a3(s.f), a3(s.f), a3("advect_tend"), a3("ru_m"), a3("rv_m"), a3("wwE"), a2("muts")); });
    if (ieva)   // module_em.F:1346-1364: mut_new = mut (= muts)
      par([&](const Cx &xt) { advect_s_implicit(xt, a3(s.o), a3("advect_tend"), a3("wwI"), a2("mut_old"), a2("muts"), dt_rk); });
    if (rk_step == 1 && c.diff_opt == 1) {   // rk_step_1: diff_opt1, module_em.F:1378-1419 (xkmhd = grid%xkhh, solve_em.F:2230)
      const float prandtl = 1.f / 3.0f;
      const float kvdq = c.kvdif / prandtl;
      const XK xkhh = c.km_opt == 1 ? XK(c.khdif / prandtl) : XK(a3("xkhh"));
      par([&](const Cx &xt) { horizontal_diffusion(xt, 'm', a3(s.f), a3(s.t), a2("muts"), xt.c1h, xt.c2h, xkhh); });
      if (s.f == "moist_qv")
        par([&](const Cx &xt) { vertical_diffusion_mp(xt, a3(s.f), a3(s.t), xt.qv_base, xt.c1h, xt.c2h, a3("alt"), a2("muts"), kvdq); });
      else
        par([&](const Cx &xt) { vertical_diffusion(xt, 'm', a3(s.f), a3(s.t), xt.c1h, xt.c2h, a3("alt"), a2("muts"), kvdq); });
    }
    if (rk_step == 1 && c.diff_6th_opt != 0 && !(s.moist ? c.moist_mix6_off : c.scalar_mix6_off))   // module_em.F:1421-1437 (dt_step = dt_rk)
      par([&](const Cx &xt) { sixth_order_diffusion(xt, 'm', a3(s.f), a3(s.t), a2("muts"), dt_rk, xt.c1h, xt.c2h, a3("phb")); });
    const bool is_tke = tke_index() && s.f == "scalar_" + std::to_string(tke_index());
    if (spec() && rk_step == 1 && has_bdy(s.f)) {   // solve_em.F:2261-2296 (moist), :2808-2843 (scalar): qv, or every species with boundary arrays
      const Bdy B = bdy(s.f);
      const A1x fcx{data.at("fcx").data()}, gcx{data.at("gcx").data()};
      par([&](const Cx &xt) { relax_bdy_scalar(xt, a3(s.t), a3(s.f), a2("mut"), B, dtbc, fcx, gcx); });
      par([&](const Cx &xt) { spec_bdy_scalar(xt, a3(s.t), B); });
    }
    par([&](const Cx &xt) { rk_update_scalar(xt, a3(s.o), a3(s.f), a3(s.t), a3("advect_tend"), a2("mu_1"), a2("mu_2"), a2("mub"), rk_step, dt_rk); });
    if (is_tke)   // bound the tke (solve_em.F:2437-2445)
      par([&](const Cx &xt) { bound_tke(xt, a3(s.f), c.tke_upper_bound); });
    // zero gradient on outflow / zero on inflow: moist species without boundary arrays (solve_em.F:2342-2355, specified only), tke (:2446-2456)
    if ((s.moist && c.specified && !c.have_bcs_moist && s.f != "moist_qv") || (is_tke && spec()))
      par([&](const Cx &xt) { flow_dep_bdy(xt, a3(s.f), a3("ru_m"), a3("rv_m")); });
  }
}

void Oracle::calc_p_rho_phi() {
  auto m = moist_list(*this, "moist_");
  par([&](const Cx &xt) { orc::calc_p_rho_phi(xt, m, a3("al"), a3("alb"), a2("mu_2"), a2("muts"), a3("ph_2"), a3("phb"), a3("p"), a3("pb"), a3("t_2")); });
}

// solve_em.F:3245-3343 (rk_phys_bc_dry_2, module_bc_em.F:1045-1104, + species)
void Oracle::rk_phys_bc_dry_2() {
  halo("HALO_EM_D2_5");        // + PERIOD_BDY_EM_D, solve_em.F:3222-3236
  halo("HALO_EM_MOIST_E_7");   // HALO_EM_MOIST_E_5/7, SCALAR_E_5/7, solve_em.F:3359-3447
  bc("u_2", 'U'); bc("v_2", 'V'); bc("w_2", 'w'); bc("t_2", 'p'); bc("ph_2", 'w'); bc("mu_2", 't');
  for (int s = 0; s < c.num_moist; s++) bc(std::string("moist_") + moist_names[s], 'p');
  for (int s = 1; s <= nsc(); s++) bc("scalar_" + std::to_string(s), 'p');
}

// moist_physics_prep_em / kessler / moist_physics_finish_em, solve_em.F:3629-4115
void Oracle::microphysics() {
  if (c.mp_physics == 0) return;
  // solve_em.F:3623-3645: the specified zone is left out (sz rows; all of x when periodic_x)
  const int sz = spec() ? c.spec_zone : 0;
  const int i_start = c.periodic_x ? std::max(x.its, x.ids) : std::max(x.its, x.ids + sz);
  const int i_end = c.periodic_x ? std::min(x.ite, x.ide - 1) : std::min(x.ite, x.ide - 1 - sz);
  const int j_start = std::max(x.jts, x.jds + sz), j_end = std::min(x.jte, x.jde - 1 - sz);
  const int k_start = x.kts, kte = x.kte, k_end = std::min(x.kte, x.kde - 1);
  const float dt = c.dt;
  A3 t_new = a3("t_2"), rho = a3("rho"), al = a3("al"), alb = a3("alb"), p = a3("p"), pb = a3("pb"), ph = a3("ph_2"), phb = a3("phb");
  A3 th_phy = a3("th_phy"), pii = a3("pi_phy"), pf = a3("p_phy"), z = a3("z"), z_at_w = a3("z_at_w"), dz8w = a3("dz8w");
  A3 h_diabatic = a3("h_diabatic"), qv = a3("moist_qv"), qc = a3("moist_qc"), qr = a3("moist_qr");
  A3 qv_diabatic = a3("qv_diabatic"), qc_diabatic = a3("qc_diabatic");
  // module_big_step_utilities_em.F:5394-5589 (p8w is not consumed by Kessler and is omitted)
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= kte; k++)
      for (int i = i_start; i <= i_end; i++) z_at_w(i, k, j) = (ph(i, k, j) + phb(i, k, j)) / g;
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= kte - 1; k++)
      for (int i = i_start; i <= i_end; i++) dz8w(i, k, j) = z_at_w(i, k + 1, j) - z_at_w(i, k, j);
  for (int j = j_start; j <= j_end; j++)
    for (int i = i_start; i <= i_end; i++) dz8w(i, kte, j) = 0.f;
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        if (c.use_theta_m == 1) th_phy(i, k, j) = (t_new(i, k, j) + t0) / (1.f + (r_v / r_d) * qv(i, k, j));
        else th_phy(i, k, j) = t_new(i, k, j) + t0;
      }
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        h_diabatic(i, k, j) = th_phy(i, k, j);
        qv_diabatic(i, k, j) = qv(i, k, j);
        qc_diabatic(i, k, j) = qc(i, k, j);
        rho(i, k, j) = 1.f / (al(i, k, j) + alb(i, k, j));
        pii(i, k, j) = std::pow((p(i, k, j) + pb(i, k, j)) / p0, rcp);
        z(i, k, j) = 0.5f * (z_at_w(i, k, j) + z_at_w(i, k + 1, j));
        pf(i, k, j) = p(i, k, j) + pb(i, k, j);
      }
  // microphysics_driver -> kessler with kte = kde-1 (solve_em.F: KTE=min(k_end,kde-1))
  Cx xk = x;
  xk.its = i_start; xk.ite = i_end; xk.jts = j_start; xk.jte = j_end; xk.kte = k_end;
  { Cx keep = x; x = xk; par([&](const Cx &xt) { kessler(xt, th_phy, qv, qc, qr, rho, pii, dt, z, dz8w, a2("rainnc"), a2("rainncv")); }); x = keep; }
  // module_big_step_utilities_em.F:5593-5784 (no_mp_heating = 0, mp_tend_lim = 10.)
  const float mp_tend_lim = 10.f;
  for (int j = j_start; j <= j_end; j++)
    for (int k = k_start; k <= k_end; k++)
      for (int i = i_start; i <= i_end; i++) {
        float mpten = th_phy(i, k, j) - h_diabatic(i, k, j);
        float qvten = qv(i, k, j) - qv_diabatic(i, k, j);
        float qcten = qc(i, k, j) - qc_diabatic(i, k, j);
        mpten = std::min(mp_tend_lim * dt, mpten);
        mpten = std::max(-mp_tend_lim * dt, mpten);
        if (c.use_theta_m == 1) {
          t_new(i, k, j) = h_diabatic(i, k, j) * (1.f + (r_v / r_d) * qv_diabatic(i, k, j)) + mpten * (1.f + (r_v / r_d) * qv(i, k, j)) +
                           (r_v / r_d) * qvten * th_phy(i, k, j) - t0;
          h_diabatic(i, k, j) = (mpten * (1.f + (r_v / r_d) * qv(i, k, j)) + (r_v / r_d) * qvten * th_phy(i, k, j)) / dt;
        } else {
          t_new(i, k, j) = t_new(i, k, j) + mpten;
          h_diabatic(i, k, j) = mpten / dt;
        }
        qv_diabatic(i, k, j) = qvten / dt;
        qc_diabatic(i, k, j) = qcten / dt;
      }
  halo("HALO_EM_THETAM");   // + PERIOD_EM_THETAM, solve_em.F:4109-4110
  bc("h_diabatic", 'p');
}

// solve_em.F:4175-4759: second calc_p_rho_phi, set_phys_bc_dry_2 + species BCs, set_w_surface
void Oracle::end_of_step() {
  calc_p_rho_phi();
  halo("HALO_EM_D3_5");        // + PERIOD_BDY_EM_D3, solve_em.F:4437-4449
  halo("HALO_EM_MOIST_E_7");
  bc("u_1", 'U'); bc("u_2", 'U'); bc("v_1", 'V'); bc("v_2", 'V'); bc("w_1", 'w'); bc("w_2", 'w');
  bc("t_1", 'p'); bc("t_2", 'p'); bc("ph_1", 'w'); bc("ph_2", 'w'); bc("mu_1", 't'); bc("mu_2", 't');
  for (int s = 0; s < c.num_moist; s++) bc(std::string("moist_") + moist_names[s], 'p');
  for (int s = 1; s <= nsc(); s++) bc("scalar_" + std::to_string(s), 'p');
  if (spec()) spec_final();
  halo("HALO_EM_C");           // + PERIOD_BDY_EM_E, solve_em.F:4730-4731
  par([&](const Cx &xt) { set_w_surface(xt, a3("w_2"), a3("u_2"), a3("v_2")); });
  halo("HALO_EM_E_5");         // solve_em.F:4781-4846
}

// solve_em.F:4540-4722: force the specified rows to the boundary-file values (no drift from summing tendencies)
void Oracle::spec_final() {
  par([&](const Cx &xt) { spec_bdy_final(xt, a3("u_2"), a2("muus"), xt.c1h, xt.c2h, xt.msfuy, bdy("u"), 'u', dtbc); });
  par([&](const Cx &xt) { spec_bdy_final(xt, a3("v_2"), a2("muvs"), xt.c1h, xt.c2h, xt.msfvx, bdy("v"), 'v', dtbc); });
  if (c.nested) par([&](const Cx &xt) { spec_bdy_final(xt, a3("w_2"), a2("muts"), xt.c1f, xt.c2f, xt.msfty, bdy("w"), 'w', dtbc); });
  par([&](const Cx &xt) { spec_bdy_final(xt, a3("t_2"), a2("muts"), xt.c1h, xt.c2h, xt.msfty, bdy("t"), 't', dtbc); });
  par([&](const Cx &xt) { spec_bdy_final(xt, a3("ph_2"), a2("muts"), xt.c1f, xt.c2f, xt.msfty, bdy("ph"), 'h', dtbc); });
  par([&](const Cx &xt) { spec_bdy_final(xt, a2("mu_2"), a2("muts"), xt.c1h, xt.c2h, xt.msfty, bdy("mu"), dtbc); });
  for (int s = 0; s < c.num_moist; s++) {
    const std::string f = std::string("moist_") + moist_names[s];
    if (has_bdy(f)) par([&](const Cx &xt) { spec_bdy_final(xt, a3(f), a2("muts"), xt.c1h, xt.c2h, xt.msfty, bdy(f), 't', dtbc); });
  }
  if (c.nested)
    for (int s = 1; s <= c.num_scalar; s++) {
      const std::string f = "scalar_" + std::to_string(s);
      par([&](const Cx &xt) { spec_bdy_final(xt, a3(f), a2("muts"), xt.c1h, xt.c2h, xt.msfty, bdy(f), 't', dtbc); });
    }
}

void Oracle::step() {
  const int rk_order = c.rk_ord;
  dtbc = dtbc + c.dt;   // solve_em.F:377
  // solve_em.F:520-550 initialize_moist_old + bc
  if (c.num_moist >= 1) {
    A3 mo = a3("moist_old_qv"), m = a3("moist_qv");
    for (int j = x.jts; j <= std::min(x.jte, x.jde - 1); j++)
      for (int k = x.kts; k <= x.kte - 1; k++)
        for (int i = x.its; i <= std::min(x.ite, x.ide - 1); i++) mo(i, k, j) = m(i, k, j);
    halo("HALO_EM_MOIST_OLD_E_7");   // + PERIOD_BDY_EM_MOIST_OLD, solve_em.F:535-536
    bc("moist_old_qv", 'p');
  }
  for (int rk_step = 1; rk_step <= rk_order; rk_step++) {
    float dt_rk, dts_rk; int ns;
    rk_params(rk_step, dt_rk, dts_rk, ns);
    rk_step_prep(rk_step);
    halo("HALO_EM_A");   // + PERIOD_BDY_EM_A, solve_em.F:708,715
    rk_phys_bc_dry_1();
    if (rk_step == 1) init_zero_tendency();
    if (rk_step == 1) diffusion_prep();
    rk_tendency(rk_step);
    small_step_prep(rk_step);
    for (int it = 1; it <= ns; it++) {
      advance_uv(rk_step);
      halo("HALO_EM_C");   // solve_em.F:1382
      advance_mu_t(rk_step);
      advance_w(rk_step);
      sumflux(rk_step, it);
      calc_p_rho(it);
    }
    small_step_finish(rk_step);
    rk_scalar_step(rk_step);
    calc_p_rho_phi();
    if (rk_step < rk_order) rk_phys_bc_dry_2();
  }
  microphysics();
  end_of_step();
}

}  // namespace orc

// ---------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------
using orc::Oracle;
static thread_local std::string g_err;
#define OGUARD(body)                                  \
  try { body; return 0; }                             \
  catch (const std::exception &e) { g_err = e.what(); return 1; }

extern "C" {
const char *oracle_last_error() { return g_err.c_str(); }
int oracle_create(const wrfgpu_dims *d, const wrfgpu_config *c, Oracle **out) { OGUARD(*out = new Oracle(*d, *c)) }
int oracle_destroy(Oracle *o) { delete o; return 0; }
size_t oracle_field_size(Oracle *o, const char *name) {
  auto it = o->data.find(name);
  return it == o->data.end() ? 0 : it->second.size();
}
int oracle_num_fields(Oracle *o) { return (int)o->names.size(); }
const char *oracle_field_name(Oracle *o, int i) { return o->names[i].c_str(); }
int oracle_set_field(Oracle *o, const char *name, const float *h, size_t n) {
  OGUARD({
    auto &v = o->data.at(name);
    if (v.size() != n) throw std::runtime_error(std::string("oracle_set_field: size mismatch for ") + name);
    std::memcpy(v.data(), h, n * sizeof(float));
  })
}
int oracle_get_field(Oracle *o, const char *name, float *h, size_t n) {
  OGUARD({
    auto &v = o->data.at(name);
    if (v.size() != n) throw std::runtime_error(std::string("oracle_get_field: size mismatch for ") + name);
    std::memcpy(h, v.data(), n * sizeof(float));
  })
}
int oracle_set_halo_callback(Oracle *o, void (*cb)(void *, const char *), void *ctx) { o->halo_cb = cb; o->halo_ctx = ctx; return 0; }
float *oracle_field_ptr(Oracle *o, const char *name) {
  auto it = o->data.find(name);
  return it == o->data.end() ? nullptr : it->second.data();
}
int oracle_max_cfl(Oracle *o, float *mv, float *mh, int reset) {
  *mv = o->max_vert_cfl; *mh = o->max_horiz_cfl;
  if (reset) o->max_vert_cfl = o->max_horiz_cfl = 0.f;
  return 0;
}
// the physical constants of oracle.h by their names in share/module_model_constants.F (tests/test_reference_tables.py)
int oracle_constant(const char *name, float *value) {
  static const std::map<std::string, float> k = {
      {"g", orc::g}, {"r_d", orc::r_d}, {"cp", orc::cp}, {"r_v", orc::r_v}, {"cv", orc::cv}, {"cvpm", orc::cvpm}, {"rcp", orc::rcp},
      {"p1000mb", orc::p1000mb}, {"t0", orc::t0}, {"p0", orc::p0}, {"cpovcv", orc::cpovcv}, {"rvovrd", orc::rvovrd},
      {"reradius", orc::reradius}, {"cb", orc::cb}, {"xlv", orc::xlv}, {"rhowater", orc::rhowater}, {"svp1", orc::svp1},
      {"svp2", orc::svp2}, {"svp3", orc::svp3}, {"svpt0", orc::svpt0}, {"ep_2", orc::ep_2}};
  auto it = k.find(name);
  if (it == k.end()) return 1;
  *value = it->second;
  return 0;
}
int oracle_set_dtbc(Oracle *o, float dtbc) { o->dtbc = dtbc; return 0; }
float oracle_get_dtbc(Oracle *o) { return o->dtbc; }
int oracle_set_threads(Oracle *o, int n) { o->nthreads = n < 1 ? 1 : n; return o->nthreads; }
int oracle_step(Oracle *o, int n) { OGUARD(for (int i = 0; i < n; i++) o->step()) }
int oracle_rk_step_prep(Oracle *o, int rk) { OGUARD(o->rk_step_prep(rk)) }
int oracle_rk_phys_bc_dry_1(Oracle *o) { OGUARD(o->rk_phys_bc_dry_1()) }
int oracle_init_zero_tendency(Oracle *o) { OGUARD(o->init_zero_tendency()) }
int oracle_rk_tendency(Oracle *o, int rk) { OGUARD(o->rk_tendency(rk)) }
int oracle_diffusion_prep(Oracle *o) { OGUARD(o->diffusion_prep()) }
int oracle_small_step_prep(Oracle *o, int rk) { OGUARD(o->small_step_prep(rk)) }
int oracle_advance_uv(Oracle *o, int rk) { OGUARD(o->advance_uv(rk)) }
int oracle_advance_mu_t(Oracle *o, int rk) { OGUARD(o->advance_mu_t(rk)) }
int oracle_advance_w(Oracle *o, int rk) { OGUARD(o->advance_w(rk)) }
int oracle_sumflux(Oracle *o, int rk, int it) { OGUARD(o->sumflux(rk, it)) }
int oracle_calc_p_rho(Oracle *o, int it) { OGUARD(o->calc_p_rho(it)) }
int oracle_small_step_finish(Oracle *o, int rk) { OGUARD(o->small_step_finish(rk)) }
int oracle_rk_scalar_step(Oracle *o, int rk) { OGUARD(o->rk_scalar_step(rk)) }
int oracle_calc_p_rho_phi(Oracle *o) { OGUARD(o->calc_p_rho_phi()) }
int oracle_rk_phys_bc_dry_2(Oracle *o) { OGUARD(o->rk_phys_bc_dry_2()) }
int oracle_microphysics(Oracle *o) { OGUARD(o->microphysics()) }
int oracle_end_of_step(Oracle *o) { OGUARD(o->end_of_step()) }
int oracle_set_physical_bc(Oracle *o, const char *name, char var) { OGUARD(o->bc(name, var)) }
}

// start_domain: the subset of dyn_em/start_em.F (start_domain_em) the hot path
// needs before step 1 -- physical-BC fills of the base state, map factors and
// both time levels (start_em.F:1421-1433 and the set_physical_bc calls before it).
extern "C" int oracle_start_domain(orc::Oracle *o) {
  OGUARD({
    o->halo("HALO_EM_INIT");   // HALO_EM_INIT_1..5, dyn_em/start_em.F
    for (const char *n : {"phb"}) o->bc(n, 'w');
    for (const char *n : {"pb", "alb", "p", "al", "alt", "t_init"}) o->bc(n, 't');   // start_em.F:1723
    for (const char *n : {"mub", "ht", "msftx", "msfty", "f", "e", "sina", "cosa"}) o->bc(n, 't');
    for (const char *n : {"msfux", "msfuy"}) o->bc(n, 'u');
    for (const char *n : {"msfvx", "msfvx_inv", "msfvy"}) o->bc(n, 'v');
    o->bc("u_1", 'U'); o->bc("u_2", 'U'); o->bc("v_1", 'V'); o->bc("v_2", 'V'); o->bc("w_1", 'w'); o->bc("w_2", 'w');
    o->bc("t_1", 'p'); o->bc("t_2", 'p'); o->bc("ph_1", 'w'); o->bc("ph_2", 'w'); o->bc("mu_1", 't'); o->bc("mu_2", 't');
    for (auto &n : o->names)
      if ((n.rfind("moist_", 0) == 0 || n.rfind("scalar_", 0) == 0) && o->kind[n] == 3) o->bc(n, 'p');
  })
}
