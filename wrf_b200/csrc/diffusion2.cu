// diffusion2.cu -- diffusion in physical space (diff_opt = 2): one launch per tendency array, each thread
// adding the vertical term, the surface flux and the horizontal term of its point in the reference's order
// (first_rk_step_part2, module_first_rk_step_part2.F:1019-1122).  Point functions and their derivation: diffusion2.cuh.
#include "diffusion2.cuh"

namespace wg {

__global__ void k_phy_rho(const Geo g, Box b, cfp alt, cfp qv, fp rho) {
  TID3;
  rho_point(g, i, k, j, alt, qv, rho);
}
__global__ void k_isotropic_km(const Geo g, Box b, float khdif, float kvdif, float khdif3, float kvdif3, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  TID3;
  A3(xkmh, i, k, j) = khdif; A3(xkmv, i, k, j) = kvdif; A3(xkhh, i, k, j) = khdif3; A3(xkhv, i, k, j) = kvdif3;
}
__global__ void k_diff2_u(const Geo g, Box b, D2In s, R4 r, fp tendency) {
  TID3;
  diff2_u_point(g, s, r, i, k, j, tendency);
}
__global__ void k_diff2_v(const Geo g, Box b, D2In s, R4 r, fp tendency) {
  TID3;
  diff2_v_point(g, s, r, i, k, j, tendency);
}
__global__ void k_diff2_w(const Geo g, Box b, D2In s, fp tendency) {
  TID3;
  diff2_w_point(g, s, i, k, j, tendency);
}
__global__ void k_diff2_s(const Geo g, Box b, D2In s, R4 r, int which, cfp var, const float *base, cfp xkhv, cfp xkhh, fp tendency, int doing_tke) {
  TID3;
  diff2_s_point(g, s, r, which, var, base, xkhv, xkhh, i, k, j, tendency, doing_tke);
}

// one thread per column: every H1 / H2 face flux once per level (diff2_s_column)
__global__ void __launch_bounds__(128) k_diff2_s_march(const Geo g, Box b, D2In s, R4 r, int which, cfp var, const float *base, cfp xkhv, cfp xkhh,
                                                         fp tendency, int doing_tke) {
  TID2;
  diff2_s_column(g, s, r, which, var, base, xkhv, xkhh, i, j, tendency, doing_tke);
}

bool diff2_is_identity(const wrfgpu_config &c) {
  return c.diff_opt == 2 && c.km_opt == 1 && c.khdif == 0.f && c.kvdif == 0.f && c.isfflx == 0 && c.tke_drag_coefficient == 0.f &&
         c.tke_heat_flux == 0.f;
}

// phy_prep's rho before the metrics (first_rk_step_part1)
void Solver::diffusion2_rho() {
  const Geo &q = g;
  LAUNCH3(this, k_phy_rho, (Box{q.its, std::min(q.ite, q.ide - 1), 1, q.kde - 1, q.jts, std::min(q.jte, q.jde - 1)}), o("alt"),
          c.num_moist >= 1 ? o("moist_qv") : (const float *)nullptr, o("rho"));
}
// isotropic_km (module_diffusion_em.F:1718-1772)
void Solver::diffusion2_isotropic_km() {
  const Geo &q = g;
  const float prandtl = 1.f / 3.0f;
  LAUNCH3(this, k_isotropic_km, (Box{q.its, std::min(q.ite, q.ide - 1), 1, q.kde, q.jts, std::min(q.jte, q.jde - 1)}), c.khdif, c.kvdif,
          c.khdif / prandtl, c.kvdif / prandtl, o("xkmh"), o("xkmv"), o("xkhh"), o("xkhv"));
}
// the rest of phy_bc, then vertical_diffusion_2 and horizontal_diffusion_2
void Solver::diffusion2_tendencies() {
  const Geo &q = g;
  bc({{"xkmv", 't'}, {"xkhv", 't'}, {"div", 't'}, {"defor11", 't'}, {"defor22", 't'}, {"defor33", 't'}, {"defor12", 'd'}, {"defor13", 'e'},
      {"defor23", 'f'}, {"rho", 't'}});
  D2In s;
  s.defor11 = o("defor11"); s.defor22 = o("defor22"); s.defor33 = o("defor33"); s.defor12 = o("defor12"); s.defor13 = o("defor13");
  s.defor23 = o("defor23"); s.xkmh = o("xkmh"); s.xkmv = o("xkmv"); s.xkhh = o("xkhh"); s.xkhv = o("xkhv"); s.zx = o("zx"); s.zy = o("zy");
  s.rdz = o("rdz"); s.rdzw = o("rdzw"); s.rho = o("rho"); s.dn = o("dn");
  s.u = o("u_2"); s.v = o("v_2"); s.t = o("t_2"); s.qv = c.num_moist >= 1 ? o("moist_qv") : nullptr;
  s.ustm = o("ustm"); s.hfx_in = o("hfx"); s.qfx = o("qfx"); s.hfx_out = o("hfx");
  s.isfflx = c.isfflx; s.use_theta_m = c.use_theta_m; s.cd0 = c.tke_drag_coefficient; s.heat_flux = c.tke_heat_flux;
  const R4 ru = d2_range(q, 'u'), rv = d2_range(q, 'v'), rm = d2_range(q, 'm');
  const Box tile{q.its, q.ite, 1, q.kde - 1, q.jts, q.jte};
  if (tke_index()) tke_rhs();   // before the diffusion (module_first_rk_step_part2.F:904-940)
  LAUNCH3(this, k_diff2_u, tile, s, ru, o("ru_tendf"));
  LAUNCH3(this, k_diff2_v, tile, s, rv, o("rv_tendf"));
  LAUNCH3(this, k_diff2_w, (Box{rm.i0, rm.i1, 2, q.kde - 1, rm.j0, rm.j1}), s, o("rw_tendf"));
  const float *tb = c.mix_full_fields ? nullptr : o("t_base"), *qb = c.mix_full_fields ? nullptr : o("qv_base");
  const Box cols{q.its, q.ite, 1, 1, q.jts, q.jte};
  LAUNCH2(this, k_diff2_s_march, cols, s, rm, 0, o("t_2"), tb, o("xkhv"), o("xkhh"), o("t_tendf"), 0);
  const std::string tke_name = tke_index() ? "scalar_" + std::to_string(tke_index()) : std::string("-");
  for (auto &sp : species) {
    const bool is_tke = sp.f == tke_name;   // tke mixes with K_m and doing_tke (:4302-4308, :2968-2982)
    LAUNCH2(this, k_diff2_s_march, cols, s, rm, sp.f == "moist_qv" ? 1 : 2, o(sp.f), sp.f == "moist_qv" ? qb : (const float *)nullptr,
            is_tke ? o("xkmv") : o("xkhv"), is_tke ? o("xkmh") : o("xkhh"), o(sp.t), is_tke ? 1 : 0);
  }
}

}  // namespace wg
