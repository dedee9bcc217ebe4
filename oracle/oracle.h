// oracle.h -- CPU restatement of the WRF-ARW dry dynamical-core step.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it, and only as the checker / CPU baseline.
//
// PARITY PIN: the reference ships no golden outputs for this path
// and no Fortran compiler exists in the build container, so this restatement
// could not be diffed against wrf.exe.  It is pinned instead to the reference's
// own routines EXECUTED FROM THEIR SOURCE (tests/fortran_subset.py translates
// them with real*4 semantics; tests/test_reference_source.py compares every phase
// of the step bit for bit over 49 configurations; the resulting digests travel as
// tests/golden/reference_source.json).  It follows the reference routine by
// routine (file:line cited on each function), in real*4, Fortran evaluation
// order, compiled with -ffp-contract=off like the GNU stanza of
// arch/configure.defaults:774-814 (no FMA).
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include <thread>
#include "../include/wrfgpu.h"

namespace orc {

// share/module_model_constants.F:17-89
constexpr float g = 9.81f;
constexpr float r_d = 287.f;
constexpr float cp = 7.f * r_d / 2.f;
constexpr float r_v = 461.6f;
constexpr float cv = cp - r_d;
constexpr float cvpm = -cv / cp;
constexpr float rcp = r_d / cp;
constexpr float p1000mb = 100000.f;
constexpr float t0 = 300.f;
constexpr float p0 = p1000mb;
constexpr float cpovcv = cp / (cp - r_d);
constexpr float rvovrd = r_v / r_d;
constexpr float reradius = 1.f / 6370.0e03f;
constexpr float cb = 25.f;
constexpr float xlv = 2.5e6f;
constexpr float rhowater = 1000.f;
constexpr float svp1 = 0.6112f, svp2 = 17.67f, svp3 = 29.65f, svpt0 = 273.15f;
constexpr float ep_2 = r_d / r_v;
constexpr float w_alpha = 0.3f, w_beta = 1.0f;

// Fortran-style array views: (ims:ime,kms:kme,jms:jme), i fastest.
struct A3 {
  float *p = nullptr; int ims = 0, kms = 0, jms = 0; long sk = 0, sj = 0;
  inline float &operator()(int i, int k, int j) const {
    return p[(long)(i - ims) + (long)(k - kms) * sk + (long)(j - jms) * sj];
  }
};
struct A2 {
  float *p = nullptr; int ims = 0, jms = 0; long sj = 0;
  inline float &operator()(int i, int j) const { return p[(long)(i - ims) + (long)(j - jms) * sj]; }
};
struct A1 {
  float *p = nullptr; int kms = 0;
  inline float &operator()(int k) const { return p[k - kms]; }
};

// Everything a model-layer routine receives besides its arrays: the three
// index triplets, config_flags and the grid scalars.
struct Cx {
  int ids, ide, jds, jde, kds, kde;
  int ims, ime, jms, jme, kms, kme;
  int ips, ipe, jps, jpe;
  int its, ite, jts, jte, kts, kte;
  wrfgpu_config c;
  float rdx, rdy, dt;
  float cf1, cf2, cf3, cfn, cfn1;
  int time_step;
  // specified / nested lateral boundaries: `specified = config_flags%specified .or. config_flags%nested`,
  // the local most routines of the reference build; o?? = "open_?? .or. specified" as nearly every loop-range test reads
  bool spec = false;
  inline bool oxs() const { return c.open_xs || spec; }
  inline bool oxe() const { return c.open_xe || spec; }
  inline bool oys() const { return c.open_ys || spec; }
  inline bool oye() const { return c.open_ye || spec; }
  // 1-D coefficient arrays (Registry.EM_COMMON:349-354, registry.hyb_coord)
  A1 znu, znw, fnm, fnp, rdnw, rdn, dnw, dn;
  A1 c1h, c2h, c1f, c2f, c3h, c4h, c3f, c4f;
  A1 u_base, v_base, t_base, qv_base, z_base;
  // 2-D map factors etc.
  A2 msftx, msfty, msfux, msfuy, msfvx, msfvx_inv, msfvy, f, e, sina, cosa, ht;
};

// eddy viscosity argument of the horizontal diffusion routines: the array xkmhd(i,k,j) of the reference, or the
// constant every element of it holds when km_opt = 1 (isotropic_km)
struct XK {
  float c = 0.f; A3 a;
  XK(float c_) : c(c_) {}
  XK(A3 a_) : a(a_) {}
  inline float operator()(int i, int k, int j) const { return a.p ? a(i, k, j) : c; }
};

// one field's boundary arrays (Registry.EM_COMMON 'b' entries: <f>_bxs/_bxe/_bys/_bye and the tendencies <f>_btxs ...):
// x sides are (jms:jme, kds:kde, spec_bdy_width), y sides (ims:ime, kds:kde, spec_bdy_width); mu has one level.
// side: 0 = Y-start, 1 = Y-end, 2 = X-start, 3 = X-end
struct Bdy {
  const float *v[4] = {nullptr, nullptr, nullptr, nullptr}, *t[4] = {nullptr, nullptr, nullptr, nullptr};
  int ims = 0, jms = 0, ni = 0, nj = 0, nk = 0;
  inline long at(int side, int along, int k, int b) const {
    return side < 2 ? (long)(along - ims) + (long)(k - 1) * ni + (long)(b - 1) * ni * nk : (long)(along - jms) + (long)(k - 1) * nj + (long)(b - 1) * nj * nk;
  }
  inline float val(int side, int along, int k, int b) const { return v[side][at(side, along, k, b)]; }
  inline float tend(int side, int along, int k, int b) const { return t[side][at(side, along, k, b)]; }
};
struct A1x {   // fcx(1:spec_bdy_width), gcx
  const float *p = nullptr;
  inline float operator()(int b) const { return p[b - 1]; }
};

inline float sign1(float a) { return std::copysign(1.0f, a); }  // SIGN(1.,a)

struct Oracle {
  wrfgpu_dims d;
  wrfgpu_config c;
  std::map<std::string, std::vector<float>> data;
  std::map<std::string, int> kind;   // 3, 2, 1
  std::vector<std::string> names;
  Cx x;
  int num_sound_steps = 0;
  float dts = 0.f;
  long n3, n2, n1;
  // multi-rank runs: called at every HALO_EM_* / PERIOD_* include site of solve_em.F with the Registry
  // name of the exchange; the test harness moves the strips between the ranks' oracles (RSL_LITE's role)
  void (*halo_cb)(void *ctx, const char *name) = nullptr;
  void *halo_ctx = nullptr;
  void halo(const char *name) { if (halo_cb) halo_cb(halo_ctx, name); }
  // km_opt = 2: the prognostic tke is carried as one more species of the scalar family (scalar_N = tke_2, scalar_old_N = tke_1,
  // scalar_tend_N = tke_tend with N = num_scalar + 1), advected with tke_adv_opt
  int tke_index() const { return (c.km_opt == 2 && c.diff_opt == 2) ? c.num_scalar + 1 : 0; }
  int nsc() const { return c.num_scalar + (tke_index() ? 1 : 0); }
  // grid%max_vert_cfl, grid%max_horiz_cfl (solve_em.F:914-930): running maxima of w_damp's Courant numbers
  float max_vert_cfl = 0.f, max_horiz_cfl = 0.f;
  // specified / nested lateral boundaries: grid%dtbc (time since the boundary arrays were valid; solve_em.F:377) and the
  // boundary arrays of `field` ("u", "v", "w", "ph", "t", "mu", "moist_qv", "scalar_1", ...)
  float dtbc = 0.f;
  bool spec() const { return c.specified || c.nested; }
  bool has_bdy(const std::string &field) const { return data.count(field + "_bxs") != 0; }
  Bdy bdy(const std::string &field);
  void add_raw(const std::string &n, size_t count);
  void spec_final();                  // solve_em.F:4540-4722
  int nthreads = 1;   // j-tiles run on std::threads like the reference's OpenMP tiles (solve_em.F:652-674)
  template <class Fn> void par(Fn fn);

  Oracle(const wrfgpu_dims &d, const wrfgpu_config &c);
  void add(const std::string &n, int kind);
  A3 a3(const std::string &n);
  A2 a2(const std::string &n);
  A1 a1(const std::string &n);
  void refresh_cx();

  // solve_em pieces (oracle/solve_em.cpp)
  void rk_step_prep(int rk_step);
  void rk_phys_bc_dry_1();
  void init_zero_tendency();
  void diffusion_prep();              // first_rk_step_part2: metrics, deformation, K (km_opt = 4)
  void rk_tendency(int rk_step);      // rk_tendency + rk_addtend_dry
  void small_step_prep(int rk_step);  // + calc_p_rho(0) + calc_coef_w + bc fills
  void advance_uv(int rk_step);
  void advance_mu_t(int rk_step);
  void advance_w(int rk_step);
  void sumflux(int rk_step, int iteration);
  void calc_p_rho(int iteration);     // + bc fills
  void small_step_finish(int rk_step);
  void rk_scalar_step(int rk_step);
  void calc_p_rho_phi();
  void rk_phys_bc_dry_2();
  void microphysics();
  void end_of_step();
  void step();
  void bc(const std::string &name, char var);
  void rk_params(int rk_step, float &dt_rk, float &dts_rk, int &nsmall) const;
};

template <class Fn> void Oracle::par(Fn fn) {
  const int nj = x.jte - x.jts + 1;
  const int T = std::max(1, std::min(nthreads, nj / 2));
  if (T <= 1) { fn(x); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) {
    Cx xt = x;
    xt.jts = x.jts + (int)((long)t * nj / T);
    xt.jte = x.jts + (int)((long)(t + 1) * nj / T) - 1;
    th.emplace_back([fn, xt]() { fn(xt); });
  }
  for (auto &q : th) q.join();
}

// ---- model-layer routines (signatures follow the Fortran dummy-arg roles) ----
// big_step.cpp
void calculate_full(const Cx &x, A2 rfield, A2 rfieldb, A2 rfieldp);
void calc_mu_uv(const Cx &x, A2 mu, A2 mub, A2 muu, A2 muv);
void calc_mu_uv_1(const Cx &x, A2 mu, A2 muu, A2 muv);
void couple_momentum(const Cx &x, A2 muu, A3 ru, A3 u, A2 msfu, A2 muv, A3 rv, A3 v, A2 msfv,
                     A2 msfv_inv, A2 mut, A3 rw, A3 w, A2 msft);
void calc_ww_cp(const Cx &x, A3 u, A3 v, A2 mup, A2 mub, A3 ww);
void calc_cq(const Cx &x, std::vector<A3> &moist, A3 cqu, A3 cqv, A3 cqw);
void calc_alt(const Cx &x, A3 alt, A3 al, A3 alb);
void calc_php(const Cx &x, A3 php, A3 ph, A3 phb);
void calc_p_rho_phi(const Cx &x, std::vector<A3> &moist, A3 al, A3 alb, A2 mu, A2 muts, A3 ph,
                    A3 phb, A3 p, A3 pb, A3 t);
void zero_tend(const Cx &x, A3 tendency);
void zero_tend2d(const Cx &x, A2 tendency);
void rhs_ph(const Cx &x, A3 ph_tend, A3 u, A3 v, A3 ww, A3 ph, A3 ph_old, A3 phb, A3 w, A2 mut,
            A2 muuf, A2 muvf);
void horizontal_pressure_gradient(const Cx &x, A3 ru_tend, A3 rv_tend, A3 ph, A3 alt, A3 p, A3 pb,
                                  A3 al, A3 php, A3 cqu, A3 cqv, A2 muu, A2 muv, A2 mu);
void pg_buoy_w(const Cx &x, A3 rw_tend, A3 p, A3 cqw, A2 muf, A2 mubf);
void w_damp(const Cx &x, A3 rw_tend, float &max_vert_cfl, float &max_horiz_cfl, A3 u, A3 v, A3 ww,
            A3 w, A2 mut);
void coriolis(const Cx &x, A3 ru, A3 rv, A3 rw, A3 ru_tend, A3 rv_tend, A3 rw_tend);
void perturbation_coriolis(const Cx &x, A3 ru_in, A3 rv_in, A3 rw, A3 ru_tend, A3 rv_tend,
                           A3 rw_tend, A2 muu, A2 muv, A3 phb, A3 ph);
void curvature(const Cx &x, A3 ru, A3 rv, A3 rw, A3 u, A3 v, A3 w, A3 ru_tend, A3 rv_tend,
               A3 rw_tend);
void rk_rayleigh_damp(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 rw_tendf, A3 t_tendf, A3 u, A3 v,
                      A3 w, A3 t, A2 mut, A2 muu, A2 muv, A3 ph, A3 phb);
void rk_addtend_dry(const Cx &x, A3 ru_tend, A3 rv_tend, A3 rw_tend, A3 ph_tend, A3 t_tend,
                    A3 ru_tendf, A3 rv_tendf, A3 rw_tendf, A3 ph_tendf, A3 t_tendf, A3 u_save,
                    A3 v_save, A3 w_save, A3 ph_save, A3 t_save, A2 mu_tend, A2 mu_tendf,
                    int rk_step, A3 h_diabatic, A2 mut);
void ww_split(const Cx &x, A3 wwE, A3 wwI, A3 ww);
void set_w_surface(const Cx &x, A3 w, A3 u, A3 v);
// diffusion.cpp (diff_opt = 1, km_opt = 1: constant K on coordinate surfaces)
void horizontal_diffusion(const Cx &x, char name, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, XK xk);
void horizontal_diffusion_3dmp(const Cx &x, A3 field, A3 tendency, A2 MUT, A1 c1, A1 c2, A3 base_3d, XK xk);
void vertical_diffusion(const Cx &x, char name, A3 field, A3 tendency, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif);
void vertical_diffusion_mp(const Cx &x, A3 field, A3 tendency, A1 base, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif);
void sixth_order_diffusion(const Cx &x, char name, A3 field, A3 tendency, A2 MUT, float dt, A1 c1, A1 c2, A3 phb);
void vertical_diffusion_3dmp(const Cx &x, A3 field, A3 tendency, A3 base_3d, A1 c1, A1 c2, A3 alt, A2 MUT, float kvdif);
void vertical_diffusion_u(const Cx &x, A3 field, A3 tendency, A1 u_base, A1 c1h, A1 c2h, A3 alt, A2 muu, float kvdif);
void vertical_diffusion_v(const Cx &x, A3 field, A3 tendency, A1 v_base, A1 c1h, A1 c2h, A3 alt, A2 muv, float kvdif);
// deform.cpp (first_rk_step_part2's diffusion preamble: metrics, deformation, 2-D Smagorinsky K)
void compute_diff_metrics(const Cx &x, A3 ph, A3 phb, A3 z, A3 rdz, A3 rdzw, A3 zx, A3 zy);
void cal_deform_and_div(const Cx &x, A3 u, A3 v, A3 w, A3 div, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13,
                        A3 defor23, A3 rdz, A3 rdzw, A3 zx, A3 zy);
void phy_prep_thermo(const Cx &x, A3 t, const A3 *qv, A3 p, A3 pb, A3 ph, A3 phb, A3 th_phy, A3 p_phy, A3 t_phy, A2 p8w_sfc, A2 t8w_sfc);
void calculate_N2(const Cx &x, A3 BN2, const A3 *qv, const A3 *qc, A3 theta, A3 t, A3 p, A2 p8w_sfc, A2 t8w_sfc, A3 rdz, A3 rdzw);
void smag_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 BN2, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13, A3 defor23,
             A3 rdzw);
void smag2d_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 defor11, A3 defor22, A3 defor12, A3 rdzw, A3 zx, A3 zy);
// diffusion2.cpp (diff_opt = 2: diffusion in physical space, called from first_rk_step_part2)
void phy_prep_rho(const Cx &x, A3 rho, A3 alt, const A3 *qv);
void horizontal_diffusion_u_2(const Cx &x, A3 tendency, A3 defor11, A3 defor12, A3 xkmh, A3 zx, A3 zy, A3 rdzw, A3 rho);
void horizontal_diffusion_v_2(const Cx &x, A3 tendency, A3 defor12, A3 defor22, A3 xkmh, A3 zx, A3 zy, A3 rdzw, A3 rho);
void horizontal_diffusion_w_2(const Cx &x, A3 tendency, A3 defor13, A3 defor23, A3 xkmv, A3 zx, A3 zy, A3 rdz, A3 rho);
void horizontal_diffusion_s(const Cx &x, A3 tendency, A3 var, A3 xkhh, A3 zx, A3 zy, A3 rdzw, A3 rho, bool doing_tke = false);
void vertical_diffusion_u_2(const Cx &x, A3 tendency, A3 defor13, A3 xkmv, A3 rho);
void vertical_diffusion_v_2(const Cx &x, A3 tendency, A3 defor23, A3 xkmv, A3 rho);
void vertical_diffusion_w_2(const Cx &x, A3 tendency, A3 defor33, A3 xkmh, A3 rho);
void vertical_diffusion_s(const Cx &x, A3 tendency, A3 var_in, const A1 *base, A3 xkhv, A3 rdz, A3 rho, bool doing_tke = false);
void surface_drag(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 u_2, A3 v_2, A3 rho, A2 ust);
void surface_heat_flux(const Cx &x, A3 rt_tendf, A3 t_2, const A3 *qv, A3 rho, A2 hfx, A2 qfx);
void surface_moisture_flux(const Cx &x, A3 qv_tendf, A2 qfx);
// diff_opt = 2 with constant K = 0 and no surface fluxes adds exact zeros to zero-initialised tendencies
inline bool diff2_is_identity(const wrfgpu_config &c) {
  return c.diff_opt == 2 && c.km_opt == 1 && c.khdif == 0.f && c.kvdif == 0.f && c.isfflx == 0 && c.tke_drag_coefficient == 0.f && c.tke_heat_flux == 0.f;
}
// tke.cpp (km_opt = 2)
void phy_prep_top(const Cx &x, A3 ph, A3 phb, A3 p_phy, A3 t_phy, A2 p8w_top, A2 t8w_top);
void tke_km(const Cx &x, A3 xkmh, A3 xkmv, A3 xkhh, A3 xkhv, A3 BN2, A3 tke, A2 p8w_sfc, A2 t8w_sfc, A2 p8w_top, A2 t8w_top, A3 theta, A3 rdz,
            A3 rdzw);
void tke_rhs(const Cx &x, A3 tendency, A3 BN2, A3 defor11, A3 defor22, A3 defor33, A3 defor12, A3 defor13, A3 defor23, A3 u, A3 v, A3 tke, A2 mu,
             A3 theta, A3 xkmh, A3 xkmv, A3 xkhv, A3 rdzw, A2 hfx, const A3 *qv, A2 ust, A3 rho);
void bound_tke(const Cx &x, A3 tke, float tke_upper_bound);
// advect.cpp
void advect_u(const Cx &x, A3 u, A3 u_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno = false);
void advect_v(const Cx &x, A3 v, A3 v_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno = false);
void advect_w(const Cx &x, A3 w, A3 w_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut, bool weno = false);
void advect_scalar(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom, A2 mut);
void advect_scalar_pd(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom,
                      A2 mut, A2 mub, A2 mu_old, float dt, bool weno = false);
void advect_scalar_weno(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom);
// ieva.cpp (module_ieva_em.F)
void ww_split_ieva(const Cx &x, A3 wwE, A3 wwI, A3 u, A3 v, A3 ww, A2 mut, float dt);
void calc_mut_old(const Cx &x, A2 mut_old, A2 mub, A2 mu_old);
void calc_mut_new(const Cx &x, A3 u, A3 v, A2 mut_old, A2 muu, A2 muv, A2 mut_new, float dt);
void advect_ph_implicit(const Cx &x, A3 pho, A3 tendency, A3 phb, A3 wwI, A2 mut, float dt_rk);
void advect_s_implicit(const Cx &x, A3 s_old, A3 tendency, A3 rom, A2 mut_old, A2 mut_new, float dt_rk);
void advect_u_implicit(const Cx &x, A3 u_old, A3 tendency, A3 rom, A2 muu_old, A2 muu_new, float dt_rk);
void advect_v_implicit(const Cx &x, A3 v_old, A3 tendency, A3 rom, A2 muv_old, A2 muv_new, float dt_rk);
void advect_w_implicit(const Cx &x, A3 w_old, A3 tendency, A3 utend, A3 vtend, A3 rom, A3 ph_new, A3 ph_old, A3 ph_tend,
                       A2 mut_old, A2 mut, A2 mut_new, float dt_rk);
// advect_mono.cpp
void advect_scalar_mono(const Cx &x, A3 field, A3 field_old, A3 tendency, A3 ru, A3 rv, A3 rom, A3 romI,
                        A2 mut, A2 mub, A2 mu_old, float dt);
// bdy.cpp (specified / nested lateral boundaries: share/module_bc.F:1122-2456, dyn_em/module_bc_em.F:17-700)
void relax_bdy_dry(const Cx &x, A3 ru_tendf, A3 rv_tendf, A3 ph_tendf, A3 t_tendf, A3 rw_tendf, A2 mu_tend, A3 ru, A3 rv, A3 ph, A3 t, A3 w,
                   A2 mu, A2 mut, const Bdy &bu, const Bdy &bv, const Bdy &bph, const Bdy &bt, const Bdy &bw, const Bdy &bmu, float dtbc,
                   A1x fcx, A1x gcx);
void relax_bdy_scalar(const Cx &x, A3 scalar_tend, A3 scalar, A2 mu, const Bdy &B, float dtbc, A1x fcx, A1x gcx);
void spec_bdy_dry(const Cx &x, A3 ru_tend, A3 rv_tend, A3 ph_tend, A3 t_tend, A3 rw_tend, A2 mu_tend, const Bdy &bu, const Bdy &bv,
                  const Bdy &bph, const Bdy &bt, const Bdy &bw, const Bdy &bmu);
void spec_bdy_scalar(const Cx &x, A3 scalar_tend, const Bdy &B);
void spec_bdyupdate(const Cx &x, A3 field, A3 field_tend, float dt, char variable);
void spec_bdyupdate(const Cx &x, A2 field, A2 field_tend, float dt);
void spec_bdyupdate_ph(const Cx &x, A3 ph_save, A3 field, A3 field_tend, A2 mu_tend, A2 muts, A1 c1, A1 c2, float dt);
void zero_grad_bdy(const Cx &x, A3 field, char variable);
void flow_dep_bdy(const Cx &x, A3 field, A3 u, A3 v);
void spec_bdy_final(const Cx &x, A3 field, A2 mu, A1 c1, A1 c2, A2 msf, const Bdy &B, char variable, float dtbc);
void spec_bdy_final(const Cx &x, A2 field, A2 mu, A1 c1, A1 c2, A2 msf, const Bdy &B, float dtbc);
// small_step.cpp
void small_step_prep(const Cx &x, A3 u_1, A3 u_2, A3 v_1, A3 v_2, A3 w_1, A3 w_2, A3 t_1, A3 t_2,
                     A3 ph_1, A3 ph_2, A2 mub, A2 mu_1, A2 mu_2, A2 muu, A2 muus, A2 muv, A2 muvs,
                     A2 mut, A2 muts, A2 mudf, A3 u_save, A3 v_save, A3 w_save, A3 t_save,
                     A3 ph_save, A2 mu_save, A3 ww, A3 ww_save, A3 c2a, A3 pb, A3 p, A3 alt,
                     int rk_step);
void small_step_finish(const Cx &x, A3 u_2, A3 u_1, A3 v_2, A3 v_1, A3 w_2, A3 w_1, A3 t_2, A3 t_1,
                       A3 ph_2, A3 ph_1, A3 ww, A3 ww1, A2 mu_2, A2 mu_1, A2 mut, A2 muts, A2 muu,
                       A2 muus, A2 muv, A2 muvs, A3 u_save, A3 v_save, A3 w_save, A3 t_save,
                       A3 ph_save, A2 mu_save, A3 h_diabatic, int number_of_small_timesteps,
                       float dts, int rk_step, int rk_order);
void calc_p_rho(const Cx &x, A3 al, A3 p, A3 ph, A3 alt, A3 t_2, A3 t_1, A3 c2a, A3 pm1, A2 mu,
                A2 mut, float smdiv, int step);
void calc_coef_w(const Cx &x, A3 a, A3 alpha, A3 gamma, A2 mut, A3 cqw, A3 c2a, float dts,
                 float epssm);
void advance_uv(const Cx &x, A3 u, A3 ru_tend, A3 v, A3 rv_tend, A3 p, A3 pb, A3 ph, A3 php,
                A3 alt, A3 al, A2 mu, A2 muu, A3 cqu, A2 muv, A3 cqv, A2 mudf, float dts);
void advance_mu_t(const Cx &x, A3 ww, A3 ww_1, A3 u, A3 u_1, A3 v, A3 v_1, A2 mu, A2 mut,
                  A2 muave, A2 muts, A2 muu, A2 muv, A2 mudf, A3 t, A3 t_1, A3 t_ave, A3 ft,
                  A2 mu_tend, float dts, float epssm);
void advance_w(const Cx &x, A3 w, A3 rw_tend, A3 ww, A3 w_save, A3 u, A3 v, A2 mu1, A2 mut,
               A2 muave, A2 muts, A3 t_2ave, A3 t_2, A3 t_1, A3 ph, A3 ph_1, A3 phb, A3 ph_tend,
               A3 c2a, A3 cqw, A3 alt, A3 a, A3 alpha, A3 gamma, float dts, float epssm);
void sumflux(const Cx &x, A3 ru, A3 rv, A3 ww, A3 u_lin, A3 v_lin, A3 ww_lin, A2 muu, A2 muv,
             A3 ru_m, A3 rv_m, A3 ww_m, int iteration, int number_of_small_timesteps);
// bc.cpp
void set_physical_bc3d(const Cx &x, A3 dat, char variable_in);
void set_physical_bc2d(const Cx &x, A2 dat, char variable_in);
// kessler.cpp
void kessler(const Cx &x, A3 t, A3 qv, A3 qc, A3 qr, A3 rho, A3 pii, float dt_in, A3 z, A3 dz8w,
             A2 rainnc, A2 rainncv);

}  // namespace orc
