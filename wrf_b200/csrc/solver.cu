// solver.cu -- device state, solve_em sequencing and the C-ABI of include/wrfgpu.h.
// Sequencing follows dyn_em/solve_em.F:578-4848 (dry path, no nest/specified/polar).
#include "common.cuh"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace wg {

static const char *k3d[] = {
    "u_1", "u_2", "v_1", "v_2", "w_1", "w_2", "t_1", "t_2", "ph_1", "ph_2",
    "phb", "pb", "alb", "p", "al", "alt", "php", "ru", "rv", "rw", "ww",
    "ru_m", "rv_m", "ww_m", "ru_tend", "rv_tend", "u_save", "v_save", "t_save", "h_diabatic",
    "ru_tendf", "rv_tendf", "rw_tend", "rw_tendf", "w_save", "ww1",
    "ph_tend", "ph_tendf", "ph_save", "t_tend", "t_tendf", "t_2save", "advect_tend",
    "alpha", "a", "gamma", "c2a", "cqu", "cqv", "cqw", "pm1", "rho", "pd_scale", "pd_scale2", "t_init", nullptr};
static const char *k3d_mp[] = {"th_phy", "pi_phy", "z", "dz8w", "qv_diabatic", "qc_diabatic", nullptr};
static const char *k2d[] = {"mu_1", "mu_2", "mub", "mudf", "muu", "muus", "muv", "muvs", "mut", "muts",
                            "muave", "mu_save", "mu_tend", "mu_tendf",
                            "msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy",
                            "f", "e", "sina", "cosa", "ht", "rainnc", "rainncv", nullptr};
static const char *k1d[] = {"znu", "znw", "fnm", "fnp", "rdnw", "rdn", "dnw", "dn",
                            "c1h", "c2h", "c1f", "c2f", "c3h", "c4h", "c3f", "c4f",
                            "u_base", "v_base", "t_base", "qv_base", "z_base", nullptr};
static const char *moist_names[] = {"qv", "qc", "qr"};

static void check_config(const wrfgpu_config &c) {
  auto bad = [](const std::string &m) { throw std::runtime_error("wrfgpu: unsupported configuration: " + m); };
  if (c.polar) bad("polar boundary conditions / filters are not on the device path");
  if (c.specified && c.nested) bad("specified and nested are exclusive");
  if (c.specified || c.nested) {
    if (c.periodic_x || c.periodic_y || c.open_xs || c.open_xe || c.open_ys || c.open_ye || c.symmetric_xs || c.symmetric_xe ||
        c.symmetric_ys || c.symmetric_ye)
      bad("specified / nested boundaries on all four sides only (the periodic_x channel of the reference is not on the device path)");
    if (c.spec_bdy_width < 2 || c.spec_zone < 1 || c.relax_zone <= c.spec_zone || c.relax_zone + 1 > c.spec_bdy_width)
      bad("need 1 <= spec_zone < relax_zone < spec_bdy_width");
    if (c.momentum_adv_opt != 1 || c.moist_adv_opt > 1 || c.scalar_adv_opt > 1 || (c.km_opt == 2 && c.tke_adv_opt > 1))
      bad("specified / nested boundaries with monotonic or WENO advection");
    if (c.zadvect_implicit > 0) bad("specified / nested boundaries with IEVA");
  }
  if (c.rk_ord != 3 && c.rk_ord != 2) bad("rk_ord must be 2 or 3");
  if ((c.h_mom_adv_order != 5 && c.h_mom_adv_order != 6) || (c.h_sca_adv_order != 5 && c.h_sca_adv_order != 6)) bad("h_*_adv_order must be 5 or 6");
  if ((c.v_mom_adv_order != 3 && c.v_mom_adv_order != 4) || (c.v_sca_adv_order != 3 && c.v_sca_adv_order != 4)) bad("v_*_adv_order must be 3 or 4");
  // advect_scalar_mono has no other orders: the reference stops with "h_order not known" (module_advect_em.F), also for the tke
  if ((c.moist_adv_opt == 2 || c.scalar_adv_opt == 2 || (c.km_opt == 2 && c.tke_adv_opt == 2)) && (c.h_sca_adv_order != 5 || c.v_sca_adv_order != 3))
    bad("monotonic advection (*_adv_opt = 2) needs h_sca_adv_order = 5 and v_sca_adv_order = 3");
  if (c.moist_adv_opt < 0 || c.moist_adv_opt > 4) bad("moist_adv_opt must be 0 (original), 1 (positive-definite), 2 (monotonic), 3 (WENO) or 4 (WENO, positive-definite)");
  if (c.scalar_adv_opt < 0 || c.scalar_adv_opt > 4) bad("scalar_adv_opt must be 0 ... 4");
  if (c.momentum_adv_opt != 1 && c.momentum_adv_opt != 3) bad("momentum_adv_opt must be 1 (original) or 3 (WENO)");
  if (!c.non_hydrostatic) bad("hydrostatic option");
  if (c.damp_opt != 0 && c.damp_opt != 2 && c.damp_opt != 3) bad("damp_opt must be 0, 2 or 3");
  if (c.zadvect_implicit < 0) bad("zadvect_implicit must be 0 or positive (IEVA on the last RK substep)");
  // diffusion: on coordinate surfaces (diff_opt 1) and in physical space (diff_opt 2), each with constant K
  // (km_opt 1), the 2-D or 3-D Smagorinsky closure (km_opt 4, 3; deform.cu) and, with diff_opt 2, the 1.5-order TKE
  // closure (km_opt 2) -- every combination the shipped ideal namelists select.  What check_config still refuses is
  // listed in DESIGN.md section 2.
  if (c.diff_6th_opt < 0 || c.diff_6th_opt > 2) bad("diff_6th_opt must be 0, 1 or 2");
  if (c.diff_opt != 0) {
    if (c.km_opt < 1 || c.km_opt > 4) bad("km_opt must be 1 (constant K), 2 (TKE), 3 (3-D Smagorinsky) or 4 (2-D Smagorinsky)");
    if (c.km_opt == 2 && c.diff_opt != 2) bad("km_opt = 2 (TKE) needs diff_opt = 2: tke_rhs is only called there (module_first_rk_step_part2.F:904)");
    if (c.km_opt == 2 && (c.tke_adv_opt < 0 || c.tke_adv_opt > 4)) bad("tke_adv_opt must be 0 ... 4");
    if (c.km_opt == 2 && c.diff_6th_opt != 0) bad("the 6th-order filter on tke (tke_mix6_off) is not on the device path");
    if (c.km_opt == 2 && !(c.c_k > 0.f)) bad("c_k must be positive with km_opt = 2");
    if ((c.km_opt == 2 || c.km_opt == 3) && !(c.mix_upper_bound > 0.f)) bad("mix_upper_bound must be positive with km_opt = 2, 3");
    if ((c.km_opt == 2 || c.km_opt == 3) && c.mix_isotropic != 0 && c.mix_isotropic != 1) bad("mix_isotropic must be 0 or 1");
    if (c.km_opt >= 3 && !(c.c_s > 0.f)) bad("c_s must be positive with km_opt = 3, 4");
    if (c.num_scalar > 7 && c.km_opt == 2) bad("at most 7 scalars with km_opt = 2 (tke is carried as one more)");
    if (c.isfflx < 0 || c.isfflx > 2) bad("isfflx must be 0, 1 or 2");
    if (c.diff_opt != 1 && c.diff_opt != 2) bad("diff_opt must be 0, 1 or 2");
    if (c.khdif < 0.f || c.kvdif < 0.f) bad("khdif, kvdif must not be negative");
  } else if (c.km_opt != 0 && c.km_opt != 1) bad("km_opt must be 0 or 1");
  if (c.mp_physics != 0 && c.mp_physics != 1) bad("mp_physics must be 0 or 1 (Kessler)");
  if (c.mp_physics == 1 && c.num_moist != 3) bad("Kessler needs num_moist = 3 (qv,qc,qr)");
  if (c.hypsometric_opt != 1 && c.hypsometric_opt != 2) bad("hypsometric_opt must be 1 or 2");
  if (c.map_proj == 6) bad("map_proj = 6 (global lat-lon)");
  if (c.num_moist < 0 || c.num_moist > 3 || c.num_scalar < 0 || c.num_scalar > 8) bad("species counts");
  if (c.time_step_sound < 0 || (c.time_step_sound % 2) != 0) bad("time_step_sound must be even (solve_em.F:607-610)");
  if ((c.periodic_x && (c.open_xs || c.open_xe || c.symmetric_xs || c.symmetric_xe)) ||
      (c.periodic_y && (c.open_ys || c.open_ye || c.symmetric_ys || c.symmetric_ye)))
    bad("conflicting lateral boundary flags");
  if (c.dx <= 0.f || c.dy <= 0.f || c.dt == 0.f) bad("dx, dy, dt");
}

Solver::Solver(const wrfgpu_dims &d_, const wrfgpu_config &c_, int dev) : d(d_), c(c_), device(dev) {
  check_config(c);
  if (d.kms != 1 || d.kds != 1) throw std::runtime_error("wrfgpu: kds/kms must be 1");
  if (d.kde - d.kds < 4) throw std::runtime_error("wrfgpu: need at least 4 vertical levels");
  if (d.kme > WG_NK - 2) throw std::runtime_error("wrfgpu: more than " + std::to_string(WG_NK - 2) + " vertical levels are not supported");
  if (d.ips - d.ims < 4 || d.ime - d.ipe < 4 || d.jps - d.jms < 4 || d.jme - d.jpe < 4)
    throw std::runtime_error("wrfgpu: memory halo must be at least 4 points");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw std::runtime_error("wrfgpu: no CUDA device available (there is no CPU fallback)");
  CUDA_CHECK(cudaSetDevice(device));
  CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (int i = 0; k3d[i]; i++) add(k3d[i], 3);
  if (c.mp_physics) for (int i = 0; k3d_mp[i]; i++) add(k3d_mp[i], 3);
  for (int i = 0; k2d[i]; i++) add(k2d[i], 2);
  for (int i = 0; k1d[i]; i++) add(k1d[i], 1);
  if (c.diff_opt != 0 && (c.km_opt >= 2 || (c.diff_opt == 2 && !diff2_is_identity(c)))) {   // metrics, deformation, K arrays (deform.cu)
    if (!c.mp_physics) add("z", 3);
    if (c.km_opt == 3 || c.km_opt == 2) add("bn2", 3);
    for (const char *n : {"ustm", "hfx", "qfx"}) add(n, 2);   // surface fluxes of a host-side surface scheme (isfflx 1, 2); zero otherwise
    for (const char *n : {"zx", "zy", "rdz", "rdzw", "div", "defor11", "defor22", "defor33", "defor12", "defor13", "defor23",
                          "xkmh", "xkmv", "xkhh", "xkhv"}) add(n, 3);
  }
  if (c.zadvect_implicit > 0) {   // IEVA: wwE / wwI and the locals of rk_tendency (module_em.F:343-347)
    add("wwE", 3); add("wwI", 3);
    for (const char *n : {"mut_old", "mut_new", "muu_old", "muu_new", "muv_old", "muv_new"}) add(n, 2);
  }
  for (int s = 0; s < c.num_moist; s++) {
    Species sp{std::string("moist_") + moist_names[s], std::string("moist_old_") + moist_names[s],
               std::string("moist_tend_") + moist_names[s], c.moist_adv_opt};
    add(sp.f, 3); add(sp.o, 3); add(sp.t, 3);
    species.push_back(sp);
  }
  for (int s = 1; s <= c.num_scalar + (tke_index() ? 1 : 0); s++) {
    Species sp{"scalar_" + std::to_string(s), "scalar_old_" + std::to_string(s), "scalar_tend_" + std::to_string(s),
               s == tke_index() ? c.tke_adv_opt : c.scalar_adv_opt};
    add(sp.f, 3); add(sp.o, 3); add(sp.t, 3);
    species.push_back(sp);
  }
  if (spec()) {   // boundary arrays (Registry.EM_COMMON 'b' entries) and the relaxation weights of lbc_fcx_gcx
    if (d.ide - d.ids < 2 * c.spec_bdy_width + 2 || d.jde - d.jds < 2 * c.spec_bdy_width + 2)
      throw std::runtime_error("wrfgpu: the domain is too small for its boundary zone");
    const int nk = d.kde - d.kds + 1;
    add_raw("fcx", c.spec_bdy_width); add_raw("gcx", c.spec_bdy_width);
    for (const char *f : {"u", "v", "w", "ph", "t"}) add_bdy(f, nk);
    add_bdy("mu", 1);
    for (auto &sp : species) {
      const bool moist = sp.f.rfind("moist_", 0) == 0, is_tke = tke_index() && sp.f == "scalar_" + std::to_string(tke_index());
      if (moist ? (sp.f == "moist_qv" || c.nested || c.have_bcs_moist) : (!is_tke && (c.nested || c.have_bcs_scalar))) add_bdy(sp.f, nk);
    }
  }
  CUDA_CHECK(cudaMalloc(&d_cfl, 2 * sizeof(float)));
  CUDA_CHECK(cudaMemset(d_cfl, 0, 2 * sizeof(float)));
  // map factors default to 1 so an un-set factor cannot divide by zero
  for (const char *m : {"msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy"}) {
    Field &f = F(m);
    std::vector<float> ones(f.nd, 1.f);
    CUDA_CHECK(cudaMemcpy(f.d, ones.data(), f.nd * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (c.time_step_sound == 0) {   // solve_em.F:446-462
    float spacing = std::min(c.dx, c.dy);
    num_sound_steps = std::max(2 * ((int)(300.f * c.dt / spacing - 0.01f) + 1), 4);
  } else {
    num_sound_steps = c.time_step_sound;
  }
  dts = c.dt / (float)num_sound_steps;
  if (const char *e = getenv("WRFGPU_BLK3")) {
    unsigned x = 0, y = 0, z = 0;
    if (sscanf(e, "%ux%ux%u", &x, &y, &z) != 3 || x * y * z == 0 || x * y * z > 1024 || (x % 32) != 0) throw std::runtime_error("wrfgpu: bad WRFGPU_BLK3");
    blk3 = dim3(x, y, z);
  }
  if (const char *e = getenv("WRFGPU_BLK2")) {
    unsigned x = 0, y = 0;
    if (sscanf(e, "%ux%u", &x, &y) != 2 || x * y != 128 || (x % 32) != 0) throw std::runtime_error("wrfgpu: bad WRFGPU_BLK2 (must hold 128 threads)");
    blk2 = dim3(x, y, 1);
  }
  rank = c.mytask;
  nranks = std::max(1, c.ntasks_x * c.ntasks_y);
  if (const char *e = getenv("WRFGPU_OVERLAP")) overlap = e[0] == '1';
  build_geo();
}

Solver::~Solver() {
  cudaSetDevice(device);
  for (auto &kv : fields) cudaFree(kv.second.d);
  halo_destroy(this);
  snapshot_destroy(this);
  tend_tile_destroy(this);
  cudaFree(d_cfl);
  if (stream && own_stream) cudaStreamDestroy(stream);
}

void Solver::flush_timing() {
  if (pending.empty()) return;
  cudaStreamSynchronize(stream);
  for (auto &p : pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    KStat &k = kstat[p.name];
    k.n++; k.ms += ms;
    cudaEventDestroy(p.a); cudaEventDestroy(p.b);
  }
  pending.clear();
}

// Device rows are padded to a multiple of 32 floats (Solver::pitch): every (k,j) row starts on a 128-byte line and the
// strides are multiples of 16 bytes, which the TMA tensor maps of the tile-staged kernels need (tend_tile.cu).  The host
// arrays keep the caller's extents, so uploads / downloads are pitched copies (xfer below).
void Solver::add(const std::string &n, int kind) {
  Field f;
  f.kind = kind;
  const size_t ni = d.ime - d.ims + 1, nk = d.kme - d.kms + 1, nj = d.jme - d.jms + 1;
  const size_t np = pitch();
  f.n = kind == 3 ? ni * nk * nj : (kind == 2 ? ni * nj : nk);
  f.nd = kind == 3 ? np * nk * nj : (kind == 2 ? np * nj : nk);
  CUDA_CHECK(cudaMalloc(&f.d, f.nd * sizeof(float)));
  CUDA_CHECK(cudaMemset(f.d, 0, f.nd * sizeof(float)));
  long off = 0;
  if (kind == 3) off = (long)d.ims + (long)d.kms * (long)np + (long)d.jms * (long)(np * nk);
  else if (kind == 2) off = (long)d.ims + (long)d.jms * (long)np;
  else off = d.kms;
  f.o = f.d - off;
  fields[n] = f;
  names.push_back(n);
}

// host (contiguous rows of ni floats) <-> device (rows of pitch() floats); `host` may also be device memory (bench.py
// builds its synthetic state on the GPU)
void Solver::copy_in(Field &f, const float *host, cudaMemcpyKind kind) {
  if (f.kind == 1) { CUDA_CHECK(cudaMemcpyAsync(f.d, host, f.n * sizeof(float), kind, stream)); return; }
  const size_t ni = d.ime - d.ims + 1;
  CUDA_CHECK(cudaMemcpy2DAsync(f.d, pitch() * sizeof(float), host, ni * sizeof(float), ni * sizeof(float), f.n / ni, kind, stream));
}
void Solver::copy_out(const Field &f, float *host, cudaMemcpyKind kind, cudaStream_t st) {
  if (f.kind == 1) { CUDA_CHECK(cudaMemcpyAsync(host, f.d, f.n * sizeof(float), kind, st)); return; }
  const size_t ni = d.ime - d.ims + 1;
  CUDA_CHECK(cudaMemcpy2DAsync(host, ni * sizeof(float), f.d, pitch() * sizeof(float), ni * sizeof(float), f.n / ni, kind, st));
}

Field &Solver::F(const std::string &n0) {
  std::string n = n0;
  if (tke_index()) {   // Registry names of the tke species
    const std::string N = std::to_string(tke_index());
    if (n0 == "tke_2") n = "scalar_" + N; else if (n0 == "tke_1") n = "scalar_old_" + N; else if (n0 == "tke_tend") n = "scalar_tend_" + N;
  }
  auto it = fields.find(n);
  if (it == fields.end()) throw std::runtime_error("wrfgpu: unknown field '" + n + "'");
  return it->second;
}

void Solver::build_geo() {
  Geo &q = g;
  q.ids = d.ids; q.ide = d.ide; q.jds = d.jds; q.jde = d.jde; q.kds = d.kds; q.kde = d.kde;
  q.ims = d.ims; q.ime = d.ime; q.jms = d.jms; q.jme = d.jme; q.kms = d.kms; q.kme = d.kme;
  q.ips = d.ips; q.ipe = d.ipe; q.jps = d.jps; q.jpe = d.jpe;
  q.its = d.ips; q.ite = d.ipe; q.jts = d.jps; q.jte = d.jpe;
  q.hni = d.ime - d.ims + 1; q.hnj = d.jme - d.jms + 1;
  q.ni = (int)pitch(); q.nk = d.kme - d.kms + 1; q.nj = d.jme - d.jms + 1;   // ni = the device row pitch (>= ime-ims+1)
  q.sk = q.ni; q.sj = (long)q.ni * q.nk;
  if ((long)q.ni * q.nk * q.nj >= (1L << 31)) throw std::runtime_error("wrfgpu: a patch of more than 2^31 points per field is not supported");
  q.isk = q.ni; q.isj = q.ni * q.nk;
  q.periodic_x = c.periodic_x; q.periodic_y = c.periodic_y;
  q.sym_xs = c.symmetric_xs; q.sym_xe = c.symmetric_xe; q.sym_ys = c.symmetric_ys; q.sym_ye = c.symmetric_ye;
  q.open_xs = c.open_xs; q.open_xe = c.open_xe; q.open_ys = c.open_ys; q.open_ye = c.open_ye;
  q.spec = (c.specified || c.nested) ? 1 : 0; q.spec_zone = c.spec_zone; q.relax_zone = c.relax_zone;
  q.top_lid = c.top_lid; q.phi_adv_z = c.phi_adv_z; q.damp_opt = c.damp_opt; q.w_damping = c.w_damping;
  q.pert_coriolis = c.pert_coriolis; q.use_theta_m = c.use_theta_m; q.hypsometric_opt = c.hypsometric_opt;
  q.rk_ord = c.rk_ord; q.num_moist = c.num_moist;
  q.rdx = 1.f / c.dx; q.rdy = 1.f / c.dy; q.dt = c.dt;
  q.cf1 = c.cf1; q.cf2 = c.cf2; q.cf3 = c.cf3; q.cfn = c.cfn; q.cfn1 = c.cfn1;
  q.epssm = c.epssm; q.smdiv = c.smdiv; q.emdiv = c.emdiv; q.dampcoef = c.dampcoef; q.zdamp = c.zdamp;
  q.w_crit_cfl = c.w_crit_cfl; q.p_top = c.p_top;
  q.w_damp_on = c.zadvect_implicit > 0 ? c.w_crit_cfl : 1.0f;   // module_big_step_utilities_em.F:2602-2608 (w_beta = 1)
  q.sgn_ts = c.time_step >= 0 ? 1.f : -1.f;
  q.sg_hm = c.h_mom_adv_order == 6 ? 0.f : q.sgn_ts; q.sg_vm = c.v_mom_adv_order == 4 ? 0.f : q.sgn_ts;
  q.sg_hs = c.h_sca_adv_order == 6 ? 0.f : q.sgn_ts; q.sg_vs = c.v_sca_adv_order == 4 ? 0.f : q.sgn_ts;
  coef_dirty = true;   // fnm ... c4f are copied from the device fields before the next launch (sync_coef)
  q.u_base = o("u_base"); q.v_base = o("v_base"); q.t_base = o("t_base"); q.z_base = o("z_base");
  q.msftx = o("msftx"); q.msfty = o("msfty"); q.msfux = o("msfux"); q.msfuy = o("msfuy"); q.msfvx = o("msfvx");
  q.msfvx_inv = o("msfvx_inv"); q.msfvy = o("msfvy"); q.f = o("f"); q.e = o("e"); q.sina = o("sina");
  q.cosa = o("cosa"); q.ht = o("ht");
}

// The 1-D coefficient arrays live in Geo by value (kernel parameter block); refresh them from the
// device fields the caller uploaded.  Runs once after the uploads, never inside the time loop.
void Solver::sync_coef() {
  struct { const char *n; float *dst; } tab[] = {
      {"fnm", g.fnm}, {"fnp", g.fnp}, {"rdnw", g.rdnw}, {"rdn", g.rdn}, {"dnw", g.dnw}, {"c1h", g.c1h}, {"c2h", g.c2h},
      {"c1f", g.c1f}, {"c2f", g.c2f}, {"c3h", g.c3h}, {"c4h", g.c4h}, {"c3f", g.c3f}, {"c4f", g.c4f}};
  CUDA_CHECK(cudaStreamSynchronize(stream));
  for (auto &t : tab) {
    Field &f = F(t.n);
    for (int k = 0; k < WG_NK; k++) t.dst[k] = 0.f;
    CUDA_CHECK(cudaMemcpy(t.dst + d.kms, f.d, f.n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  coef_dirty = false;
}

// solve_em.F:585-615
void Solver::rk_params(int rk_step, float &dt_rk, float &dts_rk, int &nsmall) const {
  if (c.rk_ord == 2) {
    if (rk_step == 1) { dt_rk = 0.5f * c.dt; dts_rk = dts; nsmall = num_sound_steps / 2; }
    else { dt_rk = c.dt; dts_rk = dts; nsmall = num_sound_steps; }
  } else {
    if (rk_step == 1) { dt_rk = c.dt / 3.f; dts_rk = dt_rk; nsmall = 1; }
    else if (rk_step == 2) { dt_rk = 0.5f * c.dt; dts_rk = dts; nsmall = num_sound_steps / 2; }
    else { dt_rk = c.dt; dts_rk = dts; nsmall = num_sound_steps; }
  }
}

__global__ void k_copy_box(const Geo g, Box b, const float *__restrict__ src, float *__restrict__ dst) {
  TID3;
  A3(dst, i, k, j) = A3(src, i, k, j);
}

__global__ void k_copy_box_v4(const Geo g, Box b, const float *__restrict__ src, float *__restrict__ dst) {   // mass columns, 16-byte accesses
  TID3_V4;
  (void)hi;
  if (him < lo) return;
  float v[4];
  ld4(&A3(src, i, k, j), v);
  st4(&A3(dst, i, k, j), v, lo, him);
}

struct PhaseTimer {
  Solver *s; const char *name; cudaEvent_t a, b;
  PhaseTimer(Solver *s_, const char *n) : s(s_), name(n) {
    if (s->timing) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s->stream); }
  }
  ~PhaseTimer() {
    if (s->timing) {
      cudaEventRecord(b, s->stream); cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b); s->phase_ms[name] += ms;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
  }
};
#define PHASE(n) PhaseTimer pt_(this, n)

// one model step, dyn_em/solve_em.F:520-4848
void Solver::step() {
  const int rk_order = c.rk_ord;
  dtbc = dtbc + c.dt;   // solve_em.F:377
  if (c.num_moist >= 1) {   // initialize_moist_old + bc (solve_em.F:520-550)
    Box bx = box_v4(g, 1, g.kde - 1);
    bx.j1 = std::min(g.jte, g.jde - 1);
    LAUNCH3(this, k_copy_box_v4, bx, (const float *)o("moist_qv"), o("moist_old_qv"));
    halo("HALO_EM_MOIST_OLD_E_7");
    bc({{"moist_old_qv", 'p'}});
  }
  for (int rk_step = 1; rk_step <= rk_order; rk_step++) {
    float dt_rk, dts_rk; int ns;
    rk_params(rk_step, dt_rk, dts_rk, ns);
    { PHASE("rk_step_prep"); rk_step_prep(rk_step); }
    halo("HALO_EM_A");
    { PHASE("bc"); rk_phys_bc_dry_1(); }
    if (rk_step == 1) { PHASE("rk_tendency"); init_zero_tendency(); }
    if (rk_step == 1) { PHASE("diffusion_prep"); diffusion_prep(); }
    // the zeroing of u_save ... t_save (module_em.F:370-428) is folded into small_step_prep's sweep: nothing reads them in between
    { PHASE("rk_tendency"); rk_tendency(rk_step, false); }
    { PHASE("small_step_prep"); small_step_prep(rk_step, true); }
    for (int it = 1; it <= ns; it++) {
      if (spec()) {   // the specified rows are advanced between the kernels (solve_em.F:1344-1620): no fusion across them
        { PHASE("advance_uv"); advance_uv(rk_step); halo("HALO_EM_C"); }
        { PHASE("advance_mu_t"); advance_mu_t(rk_step); }
        { PHASE("advance_w"); advance_w(rk_step); sumflux(rk_step, it); calc_p_rho(it); }
        continue;
      }
      { PHASE("advance_uv"); advance_uv_fused(rk_step, it); }   // + HALO_EM_C, overlapped with the kernel's interior
      { PHASE("advance_mu_t"); advance_mu_t(rk_step, true); }
      { PHASE("advance_w"); advance_w_fused(rk_step, it); }
    }
    { PHASE("small_step_finish"); small_step_finish(rk_step); }
    { PHASE("rk_scalar"); rk_scalar_step(rk_step); }
    { PHASE("calc_p_rho_phi"); calc_p_rho_phi(); }
    if (rk_step < rk_order) { PHASE("bc"); rk_phys_bc_dry_2(); }
  }
  if (c.mp_physics) { PHASE("microphysics"); microphysics(); }
  { PHASE("end_of_step"); end_of_step(); }
}

}  // namespace wg

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using wg::Solver;
thread_local std::string g_wrfgpu_err;
#define g_err g_wrfgpu_err

#define GUARD(body)                                                \
  try { cudaSetDevice(s->s.device); body; return 0; }              \
  catch (const std::exception &e) { g_err = e.what(); return 1; }

extern "C" {

const char *wrfgpu_last_error(void) { return g_err.c_str(); }
int wrfgpu_version(void) { return 200; }
int wrfgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int wrfgpu_create(const wrfgpu_dims *d, const wrfgpu_config *c, int device, wrfgpu_solver **out) {
  try { *out = new wrfgpu_solver(*d, *c, device); return 0; }
  catch (const std::exception &e) { g_err = e.what(); *out = nullptr; return 1; }
}
int wrfgpu_destroy(wrfgpu_solver *s) { delete s; return 0; }

size_t wrfgpu_field_size(const wrfgpu_solver *s, const char *name) {
  auto it = s->s.fields.find(name);
  return it == s->s.fields.end() ? 0 : it->second.n;
}
int wrfgpu_num_fields(const wrfgpu_solver *s) { return (int)s->s.names.size(); }
const char *wrfgpu_field_name(const wrfgpu_solver *s, int idx) { return s->s.names[idx].c_str(); }
void *wrfgpu_field_devptr(wrfgpu_solver *s, const char *name) {
  auto it = s->s.fields.find(name);
  return it == s->s.fields.end() ? nullptr : (void *)it->second.d;
}

static void xfer(Solver &S, const char *name, const float *in, float *out, size_t n, bool async) {
  wg::Field &f = S.F(name);
  if (n != f.n) throw std::runtime_error(std::string("wrfgpu: size mismatch for field '") + name + "'");
  cudaPointerAttributes at;
  const void *hp = in ? (const void *)in : (const void *)out;
  bool dev = (cudaPointerGetAttributes(&at, hp) == cudaSuccess && at.type == cudaMemoryTypeDevice);
  cudaGetLastError();
  cudaMemcpyKind kind = in ? (dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                           : (dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost);
  if (in && f.kind == 1 && !f.raw) S.coef_dirty = true;
  if (in) S.copy_in(f, in, kind);
  else S.copy_out(f, out, kind, S.stream);
  if (!async) CUDA_CHECK(cudaStreamSynchronize(S.stream));
}
int wrfgpu_set_field(wrfgpu_solver *s, const char *name, const float *host, size_t n) { GUARD(xfer(s->s, name, host, nullptr, n, false)) }
int wrfgpu_get_field(wrfgpu_solver *s, const char *name, float *host, size_t n) { GUARD(xfer(s->s, name, nullptr, host, n, false)) }
int wrfgpu_upload(wrfgpu_solver *s, int nf, const char *const *names, const float *const *host) {
  GUARD({
    for (int i = 0; i < nf; i++) xfer(s->s, names[i], host[i], nullptr, s->s.F(names[i]).n, true);
    CUDA_CHECK(cudaStreamSynchronize(s->s.stream));
  })
}
int wrfgpu_download(wrfgpu_solver *s, int nf, const char *const *names, float *const *host) {
  GUARD({
    for (int i = 0; i < nf; i++) xfer(s->s, names[i], nullptr, host[i], s->s.F(names[i]).n, true);
    CUDA_CHECK(cudaStreamSynchronize(s->s.stream));
  })
}
int wrfgpu_step(wrfgpu_solver *s, int nsteps) {
  GUARD({
    for (int i = 0; i < nsteps; i++) s->s.step();
    CUDA_CHECK(cudaGetLastError());
  })
}
int wrfgpu_solve_em(wrfgpu_solver *s, int n_in, const char *const *in_names, const float *const *in_host,
                    int n_out, const char *const *out_names, float *const *out_host) {
  GUARD({
    for (int i = 0; i < n_in; i++) xfer(s->s, in_names[i], in_host[i], nullptr, s->s.F(in_names[i]).n, true);
    s->s.step();
    for (int i = 0; i < n_out; i++) xfer(s->s, out_names[i], nullptr, out_host[i], s->s.F(out_names[i]).n, true);
    CUDA_CHECK(cudaStreamSynchronize(s->s.stream));
    CUDA_CHECK(cudaGetLastError());
  })
}
int wrfgpu_sync(wrfgpu_solver *s) { GUARD({ CUDA_CHECK(cudaStreamSynchronize(s->s.stream)); CUDA_CHECK(cudaGetLastError()); }) }
long long wrfgpu_launch_count(const wrfgpu_solver *s) { return s->s.launches; }
int wrfgpu_max_cfl(wrfgpu_solver *s, float *max_vert_cfl, float *max_horiz_cfl, int reset) {
  GUARD({
    if (!s->s.c.use_adaptive_time_step) throw std::runtime_error("wrfgpu_max_cfl: the CFL maxima are only computed with use_adaptive_time_step = 1");
    float h[2];
    wg::halo_max2(&s->s, s->s.d_cfl, h);
    *max_vert_cfl = h[0]; *max_horiz_cfl = h[1];
    if (reset) CUDA_CHECK(cudaMemsetAsync(s->s.d_cfl, 0, 2 * sizeof(float), s->s.stream));
  })
}
int wrfgpu_set_dtbc(wrfgpu_solver *s, float dtbc) { GUARD(s->s.dtbc = dtbc) }
int wrfgpu_get_dtbc(wrfgpu_solver *s, float *dtbc) { GUARD(*dtbc = s->s.dtbc) }
int wrfgpu_set_stream(wrfgpu_solver *s, void *cuda_stream) {
  GUARD({
    CUDA_CHECK(cudaStreamSynchronize(s->s.stream));
    if (s->s.own_stream) cudaStreamDestroy(s->s.stream);
    s->s.stream = (cudaStream_t)cuda_stream;
    s->s.own_stream = false;
  })
}
/* per-kernel statistics as text lines "name count total_ms\n"; returns bytes needed */
size_t wrfgpu_kernel_report(wrfgpu_solver *s, char *buf, size_t cap) {
  s->s.flush_timing();
  std::string out;
  for (auto &kv : s->s.kstat) out += kv.first + " " + std::to_string(kv.second.n) + " " + std::to_string(kv.second.ms) + "\n";
  if (buf && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}
double wrfgpu_phase_ms(wrfgpu_solver *s, const char *name) {
  if (!name) { s->s.flush_timing(); s->s.phase_ms.clear(); s->s.kstat.clear(); s->s.timing = true; return 0; }
  if (!strcmp(name, "__off__")) { s->s.flush_timing(); s->s.timing = false; return 0; }
  auto it = s->s.phase_ms.find(name);
  return it == s->s.phase_ms.end() ? 0.0 : it->second;
}

int wrfgpu_start_domain(wrfgpu_solver *s) { GUARD(s->s.start_domain()) }
int wrfgpu_rk_step_prep(wrfgpu_solver *s, int rk) { GUARD(s->s.rk_step_prep(rk)) }
int wrfgpu_rk_phys_bc_dry_1(wrfgpu_solver *s) { GUARD(s->s.rk_phys_bc_dry_1()) }
int wrfgpu_init_zero_tendency(wrfgpu_solver *s) { GUARD(s->s.init_zero_tendency()) }
int wrfgpu_diffusion_prep(wrfgpu_solver *s) { GUARD(s->s.diffusion_prep()) }
int wrfgpu_rk_tendency(wrfgpu_solver *s, int rk) { GUARD(s->s.rk_tendency(rk)) }
int wrfgpu_small_step_prep(wrfgpu_solver *s, int rk) { GUARD(s->s.small_step_prep(rk)) }
int wrfgpu_advance_uv(wrfgpu_solver *s, int rk) { GUARD(s->s.advance_uv(rk)) }
int wrfgpu_advance_mu_t(wrfgpu_solver *s, int rk) { GUARD(s->s.advance_mu_t(rk)) }
int wrfgpu_advance_w(wrfgpu_solver *s, int rk) { GUARD(s->s.advance_w(rk)) }
int wrfgpu_sumflux(wrfgpu_solver *s, int rk, int it) { GUARD(s->s.sumflux(rk, it)) }
int wrfgpu_calc_p_rho(wrfgpu_solver *s, int it) { GUARD(s->s.calc_p_rho(it)) }
int wrfgpu_small_step_finish(wrfgpu_solver *s, int rk) { GUARD(s->s.small_step_finish(rk)) }
int wrfgpu_rk_scalar_step(wrfgpu_solver *s, int rk) { GUARD(s->s.rk_scalar_step(rk)) }
int wrfgpu_calc_p_rho_phi(wrfgpu_solver *s) { GUARD(s->s.calc_p_rho_phi()) }
int wrfgpu_rk_phys_bc_dry_2(wrfgpu_solver *s) { GUARD(s->s.rk_phys_bc_dry_2()) }
int wrfgpu_microphysics(wrfgpu_solver *s) { GUARD(s->s.microphysics()) }
int wrfgpu_end_of_step(wrfgpu_solver *s) { GUARD(s->s.end_of_step()) }
int wrfgpu_set_physical_bc(wrfgpu_solver *s, const char *name, char var) {
  GUARD(s->s.bc({{std::string(name), var}}))
}
int wrfgpu_halo(wrfgpu_solver *s, const char *halo_name) { GUARD(s->s.halo(halo_name)) }

}  // extern "C"
