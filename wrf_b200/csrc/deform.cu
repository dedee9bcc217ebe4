// deform.cu -- the diffusion preamble of the first RK substep for the closures that compute K
// (first_rk_step_part2, dyn_em/module_first_rk_step_part2.F:427-680): compute_diff_metrics, cal_deform_and_div,
// the eddy viscosities (smag2d_km, km_opt = 4), their exchanges (HALO_EM_TKE_C / _D / _E) and boundary fills.
// Three launches per time step; the point functions and their derivation are in deform.cuh.
#include "deform.cuh"

namespace wg {

__global__ void k_diff_metrics(const Geo g, Box b, cfp ph, cfp phb, fp z, fp rdz, fp rdzw, fp zx, fp zy) {
  TID3;
  diff_metrics_point(g, i, k, j, ph, phb, z, rdz, rdzw, zx, zy);
}

__global__ void k_deform(const Geo g, Box b, DefR r, DefIn s, DefOut o) {
  TID3;
  deform_point(g, r, s, o, i, k, j);
}

__global__ void k_smag2d(const Geo g, Box b, SmagIn s, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  TID3;
  smag2d_point(g, s, i, k, j, xkmh, xkmv, xkhh, xkhv);
}

__global__ void k_smag3d(const Geo g, Box b, Smag3In s, fp bn2, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  TID3;
  smag3d_point(g, s, i, k, j, bn2, xkmh, xkmv, xkhh, xkhv);
}

__global__ void k_tke_km(const Geo g, Box b, TkeIn s, fp bn2, fp xkmh, fp xkmv, fp xkhh, fp xkhv) {
  TID3;
  tke_km_point(g, s, i, k, j, bn2, xkmh, xkmv, xkhh, xkhv);
}
__global__ void k_tke_rhs(const Geo g, Box b, TkeIn s, fp tendency) {
  TID3;
  tke_rhs_point(g, s, i, k, j, tendency);
}

// inputs of tke_km / tke_rhs
static TkeIn tke_in(Solver *S) {
  const wrfgpu_config &c = S->c;
  TkeIn t;
  Smag3In &s3 = t.th;
  s3.t_2 = S->o("t_2"); s3.qv = c.num_moist >= 1 ? S->o("moist_qv") : nullptr; s3.qc = c.num_moist >= 2 ? S->o("moist_qc") : nullptr;
  s3.p = S->o("p"); s3.pb = S->o("pb"); s3.ph = S->o("ph_2"); s3.phb = S->o("phb"); s3.rdz = S->o("rdz"); s3.rdzw = S->o("rdzw");
  s3.defor11 = S->o("defor11"); s3.defor22 = S->o("defor22"); s3.defor33 = S->o("defor33"); s3.defor12 = S->o("defor12");
  s3.defor13 = S->o("defor13"); s3.defor23 = S->o("defor23"); s3.dn = S->o("dn");
  s3.dx = c.dx; s3.dy = c.dy; s3.dt = c.dt; s3.c_s = c.c_s; s3.mix_upper_bound = c.mix_upper_bound; s3.mix_isotropic = c.mix_isotropic;
  s3.use_theta_m = c.use_theta_m;
  t.tke = S->o("tke_2"); t.bn2 = S->o("bn2"); t.xkmh = S->o("xkmh"); t.xkmv = S->o("xkmv"); t.xkhv = S->o("xkhv"); t.u = S->o("u_2");
  t.v = S->o("v_2"); t.mut = S->o("mut"); t.rho = S->o("rho"); t.hfx = S->o("hfx"); t.ustm = S->o("ustm");
  t.c_k = c.c_k; t.cd0 = c.tke_drag_coefficient; t.heat_flux = c.tke_heat_flux; t.isfflx = c.isfflx; t.diff_opt = c.diff_opt;
  return t;
}
// tke_rhs (module_first_rk_step_part2.F:904-940) on tke_km's range
void Solver::tke_rhs() {
  const Geo &q = g;
  int i_start = q.its, i_end = std::min(q.ite, q.ide - 1), j_start = q.jts, j_end = std::min(q.jte, q.jde - 1);
  if (q.oxs()) i_start = std::max(q.ids + 1, q.its);
  if (q.oxe()) i_end = std::min(q.ide - 2, q.ite);
  if (q.oys()) j_start = std::max(q.jds + 1, q.jts);
  if (q.oye()) j_end = std::min(q.jde - 2, q.jte);
  if (q.periodic_x) { i_start = q.its; i_end = std::min(q.ite, q.ide - 1); }
  LAUNCH3(this, k_tke_rhs, (Box{i_start, i_end, 1, q.kde - 1, j_start, j_end}), tke_in(this), o("tke_tend"));
}

void Solver::diffusion_prep() {
  if (c.diff_opt == 0) return;
  if (c.diff_opt == 1 && c.km_opt == 1) return;   // K is a constant the diffusion kernels of rk_tendency take by value
  if (diff2_is_identity(c)) return;               // diff_opt = 2 with all coefficients zero adds exact zeros (shipped em_hill2d_x)
  const Geo &q = g;
  if (c.diff_opt == 2) diffusion2_rho();
  // compute_diff_metrics: rdz, rdzw on rows jts-1..jte, the rest on the tile
  LAUNCH3(this, k_diff_metrics, (Box{q.its, q.ite, 1, q.kde, q.jts - 1, q.jte}), o("ph_2"), o("phb"), o("z"), o("rdz"), o("rdzw"), o("zx"), o("zy"));
  halo("HALO_EM_TKE_C");   // + PERIOD_BDY_EM_A1 (:446-447)
  if (c.diff_opt == 2) { halo("HALO_EM_TKE_C_SFC"); bc({{"ustm", 't'}}); }   // ustm of HALO_EM_TKE_C, set_physical_bc2d :492
  bc({{"rdzw", 'w'}, {"rdz", 'w'}, {"z", 'w'}, {"zx", 'e'}, {"zy", 'f'}});
  // cal_deform_and_div
  DefIn in;
  in.u = o("u_2"); in.v = o("v_2"); in.w = o("w_2"); in.rdz = o("rdz"); in.rdzw = o("rdzw"); in.zx = o("zx"); in.zy = o("zy");
  in.dn = o("dn"); in.dnw = o("dnw"); in.u_base = o("u_base"); in.v_base = o("v_base");
  DefOut out;
  out.div = o("div"); out.defor11 = o("defor11"); out.defor22 = o("defor22"); out.defor33 = o("defor33");
  out.defor12 = o("defor12"); out.defor13 = o("defor13"); out.defor23 = o("defor23");
  const DefR r = def_ranges(q, c.mix_full_fields);
  LAUNCH3(this, k_deform, (Box{q.its, q.ite, 1, q.kde, q.jts, q.jte}), r, in, out);
  halo("HALO_EM_TKE_D");   // :571
  if (c.km_opt == 1) diffusion2_isotropic_km();
  else {   // calculate_km_kh on the range its closures share (module_diffusion_em.F:1816-1831, :1974-1989, :2078-2093)
    int i_start = q.its, i_end = std::min(q.ite, q.ide - 1), j_start = q.jts, j_end = std::min(q.jte, q.jde - 1);
    if (q.oxs()) i_start = std::max(q.ids + 1, q.its);
    if (q.oxe()) i_end = std::min(q.ide - 2, q.ite);
    if (q.oys()) j_start = std::max(q.jds + 1, q.jts);
    if (q.oye()) j_end = std::min(q.jde - 2, q.jte);
    if (q.periodic_x) { i_start = q.its; i_end = std::min(q.ite, q.ide - 1); }
    if (c.km_opt == 2) {   // calculate_N2 + tke_km in one launch
      LAUNCH3(this, k_tke_km, (Box{i_start, i_end, 1, q.kde - 1, j_start, j_end}), tke_in(this), o("bn2"), o("xkmh"), o("xkmv"), o("xkhh"), o("xkhv"));
    } else if (c.km_opt == 3) {   // calculate_N2 + smag_km in one launch
      Smag3In s3;
      s3.t_2 = o("t_2"); s3.qv = c.num_moist >= 1 ? o("moist_qv") : nullptr; s3.qc = c.num_moist >= 2 ? o("moist_qc") : nullptr;
      s3.p = o("p"); s3.pb = o("pb"); s3.ph = o("ph_2"); s3.phb = o("phb"); s3.rdz = o("rdz"); s3.rdzw = o("rdzw");
      s3.defor11 = o("defor11"); s3.defor22 = o("defor22"); s3.defor33 = o("defor33"); s3.defor12 = o("defor12"); s3.defor13 = o("defor13");
      s3.defor23 = o("defor23"); s3.dn = o("dn");
      s3.dx = c.dx; s3.dy = c.dy; s3.dt = c.dt; s3.c_s = c.c_s; s3.mix_upper_bound = c.mix_upper_bound; s3.mix_isotropic = c.mix_isotropic;
      s3.use_theta_m = c.use_theta_m;
      LAUNCH3(this, k_smag3d, (Box{i_start, i_end, 1, q.kde - 1, j_start, j_end}), s3, o("bn2"), o("xkmh"), o("xkmv"), o("xkhh"), o("xkhv"));
    } else {
      SmagIn sm;
      sm.defor11 = o("defor11"); sm.defor22 = o("defor22"); sm.defor12 = o("defor12"); sm.rdzw = o("rdzw"); sm.zx = o("zx"); sm.zy = o("zy");
      sm.dx = c.dx; sm.dy = c.dy; sm.c_s = c.c_s; sm.diff_opt = c.diff_opt;
      LAUNCH3(this, k_smag2d, (Box{i_start, i_end, 1, q.kde - 1, j_start, j_end}), sm, o("xkmh"), o("xkmv"), o("xkhh"), o("xkhv"));
    }
  }
  halo("HALO_EM_TKE_E");   // + PERIOD_BDY_EM_PHY_BC, HALO_EM_PHYS_DIFFUSION (:644-650, :759)
  bc({{"xkmh", 't'}, {"xkhh", 't'}});   // phy_bc (module_diffusion_em.F:5901), diff_opt = 1 part
  if (c.diff_opt == 2) diffusion2_tendencies();
}

}  // namespace wg
