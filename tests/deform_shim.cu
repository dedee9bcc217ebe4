// Host build of wrf_b200/csrc/deform.cuh for tests/test_deform_host.py: runs the product's __host__ __device__ point
// functions of the diffusion preamble (metrics, deformation, 2-D Smagorinsky K) in plain loops on the CPU, so that the
// exact code the kernels of deform.cu execute can be compared with the oracle bit for bit without a GPU.
// Test infrastructure; compiled with nvcc (host code only is run) and -ffp-contract=off.
#include "../wrf_b200/csrc/diffusion2.cuh"
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

using namespace wg;

// stage: 0 rho, 1 metrics, 2 deformation, 3 smag2d K, 7 calculate_N2 + smag_km, 4 isotropic K, 5 diff_opt = 2 tendencies of u, v, w, theta,
// 6 diff_opt = 2 tendency of the species names[nn-2] (field) / names[nn-1] (tendency) -- which = 1 for qv, 3 for tke, else 2;
// 8 calculate_N2 + tke_km, 9 tke_rhs
extern "C" int deform_shim_run(int stage, const wrfgpu_dims *dp, const wrfgpu_config *cp, int nn, const char *const *names, float *const *ptrs,
                               int which) {
  const wrfgpu_dims &d = *dp;
  const wrfgpu_config &c = *cp;
  std::map<std::string, float *> tab;
  for (int n = 0; n < nn; n++) tab[names[n]] = ptrs[n];
  Geo *gp = new Geo();   // ~7 KB
  Geo &g = *gp;
  memset(gp, 0, sizeof(Geo));
  g.ids = d.ids; g.ide = d.ide; g.jds = d.jds; g.jde = d.jde; g.kds = d.kds; g.kde = d.kde;
  g.ims = d.ims; g.ime = d.ime; g.jms = d.jms; g.jme = d.jme; g.kms = d.kms; g.kme = d.kme;
  g.ips = d.ips; g.ipe = d.ipe; g.jps = d.jps; g.jpe = d.jpe;
  g.its = d.ips; g.ite = d.ipe; g.jts = d.jps; g.jte = d.jpe;
  g.ni = d.ime - d.ims + 1; g.nk = d.kme - d.kms + 1; g.nj = d.jme - d.jms + 1;
  g.sk = g.ni; g.sj = (long)g.ni * g.nk; g.isk = g.ni; g.isj = g.ni * g.nk;
  g.periodic_x = c.periodic_x; g.periodic_y = c.periodic_y;
  g.open_xs = c.open_xs; g.open_xe = c.open_xe; g.open_ys = c.open_ys; g.open_ye = c.open_ye;
  g.rdx = 1.f / c.dx; g.rdy = 1.f / c.dy; g.cf1 = c.cf1; g.cf2 = c.cf2; g.cf3 = c.cf3;
  const long o3 = d.ims + (long)d.kms * g.ni + (long)d.jms * g.ni * g.nk, o2 = d.ims + (long)d.jms * g.ni, o1 = d.kms;
  auto p3 = [&](const char *n) -> float * { auto it = tab.find(n); return it == tab.end() ? nullptr : it->second - o3; };
  auto p2 = [&](const char *n) -> float * { return tab.at(n) - o2; };
  auto p1 = [&](const char *n) -> float * { return tab.at(n) - o1; };
  g.msftx = p2("msftx"); g.msfty = p2("msfty"); g.msfux = p2("msfux"); g.msfuy = p2("msfuy"); g.msfvx = p2("msfvx"); g.msfvy = p2("msfvy");
  for (int k = d.kms; k <= d.kme; k++) { g.fnm[k] = p1("fnm")[k]; g.fnp[k] = p1("fnp")[k]; g.dnw[k] = p1("dnw")[k]; }
  const int ie = std::min(g.ite, g.ide - 1), je = std::min(g.jte, g.jde - 1);
  if (stage == 0) {
    for (int j = g.jts; j <= je; j++)
      for (int k = 1; k <= g.kde - 1; k++)
        for (int i = g.its; i <= ie; i++) rho_point(g, i, k, j, p3("alt"), p3("moist_qv"), p3("rho"));
  } else if (stage == 1) {
    for (int j = g.jts - 1; j <= g.jte; j++)
      for (int k = 1; k <= g.kde; k++)
        for (int i = g.its; i <= g.ite; i++) diff_metrics_point(g, i, k, j, p3("ph_2"), p3("phb"), p3("z"), p3("rdz"), p3("rdzw"), p3("zx"), p3("zy"));
  } else if (stage == 2) {
    DefIn in;
    in.u = p3("u_2"); in.v = p3("v_2"); in.w = p3("w_2"); in.rdz = p3("rdz"); in.rdzw = p3("rdzw"); in.zx = p3("zx"); in.zy = p3("zy");
    in.dn = p1("dn"); in.dnw = p1("dnw"); in.u_base = p1("u_base"); in.v_base = p1("v_base");
    DefOut out;
    out.div = p3("div"); out.defor11 = p3("defor11"); out.defor22 = p3("defor22"); out.defor33 = p3("defor33"); out.defor12 = p3("defor12");
    out.defor13 = p3("defor13"); out.defor23 = p3("defor23");
    const DefR r = def_ranges(g, c.mix_full_fields);
    for (int j = g.jts; j <= g.jte; j++)
      for (int k = 1; k <= g.kde; k++)
        for (int i = g.its; i <= g.ite; i++) deform_point(g, r, in, out, i, k, j);
  } else if (stage == 3) {
    int i_start = g.its, i_end = ie, j_start = g.jts, j_end = je;   // as Solver::diffusion_prep
    if (g.open_xs) i_start = std::max(g.ids + 1, g.its);
    if (g.open_xe) i_end = std::min(g.ide - 2, g.ite);
    if (g.open_ys) j_start = std::max(g.jds + 1, g.jts);
    if (g.open_ye) j_end = std::min(g.jde - 2, g.jte);
    if (g.periodic_x) { i_start = g.its; i_end = ie; }
    SmagIn sm;
    sm.defor11 = p3("defor11"); sm.defor22 = p3("defor22"); sm.defor12 = p3("defor12"); sm.rdzw = p3("rdzw"); sm.zx = p3("zx"); sm.zy = p3("zy");
    sm.dx = c.dx; sm.dy = c.dy; sm.c_s = c.c_s; sm.diff_opt = c.diff_opt;
    for (int j = j_start; j <= j_end; j++)
      for (int k = 1; k <= g.kde - 1; k++)
        for (int i = i_start; i <= i_end; i++) smag2d_point(g, sm, i, k, j, p3("xkmh"), p3("xkmv"), p3("xkhh"), p3("xkhv"));
  } else if (stage == 7) {
    int i_start = g.its, i_end = ie, j_start = g.jts, j_end = je;   // as Solver::diffusion_prep
    if (g.open_xs) i_start = std::max(g.ids + 1, g.its);
    if (g.open_xe) i_end = std::min(g.ide - 2, g.ite);
    if (g.open_ys) j_start = std::max(g.jds + 1, g.jts);
    if (g.open_ye) j_end = std::min(g.jde - 2, g.jte);
    if (g.periodic_x) { i_start = g.its; i_end = ie; }
    Smag3In s3;
    s3.t_2 = p3("t_2"); s3.qv = p3("moist_qv"); s3.qc = p3("moist_qc"); s3.p = p3("p"); s3.pb = p3("pb"); s3.ph = p3("ph_2"); s3.phb = p3("phb");
    s3.rdz = p3("rdz"); s3.rdzw = p3("rdzw"); s3.defor11 = p3("defor11"); s3.defor22 = p3("defor22"); s3.defor33 = p3("defor33");
    s3.defor12 = p3("defor12"); s3.defor13 = p3("defor13"); s3.defor23 = p3("defor23"); s3.dn = p1("dn");
    s3.dx = c.dx; s3.dy = c.dy; s3.dt = c.dt; s3.c_s = c.c_s; s3.mix_upper_bound = c.mix_upper_bound; s3.mix_isotropic = c.mix_isotropic;
    s3.use_theta_m = c.use_theta_m;
    for (int j = j_start; j <= j_end; j++)
      for (int k = 1; k <= g.kde - 1; k++)
        for (int i = i_start; i <= i_end; i++) smag3d_point(g, s3, i, k, j, p3("bn2"), p3("xkmh"), p3("xkmv"), p3("xkhh"), p3("xkhv"));
  } else if (stage == 8 || stage == 9) {   // tke_km / tke_rhs, inputs as deform.cu's tke_in()
    int i_start = g.its, i_end = ie, j_start = g.jts, j_end = je;
    if (g.open_xs) i_start = std::max(g.ids + 1, g.its);
    if (g.open_xe) i_end = std::min(g.ide - 2, g.ite);
    if (g.open_ys) j_start = std::max(g.jds + 1, g.jts);
    if (g.open_ye) j_end = std::min(g.jde - 2, g.jte);
    if (g.periodic_x) { i_start = g.its; i_end = ie; }
    for (int k = d.kms; k <= d.kme; k++) { g.c1h[k] = p1("c1h")[k]; g.c2h[k] = p1("c2h")[k]; }
    TkeIn t;
    Smag3In &s3 = t.th;
    s3.t_2 = p3("t_2"); s3.qv = p3("moist_qv"); s3.qc = p3("moist_qc"); s3.p = p3("p"); s3.pb = p3("pb"); s3.ph = p3("ph_2"); s3.phb = p3("phb");
    s3.rdz = p3("rdz"); s3.rdzw = p3("rdzw"); s3.defor11 = p3("defor11"); s3.defor22 = p3("defor22"); s3.defor33 = p3("defor33");
    s3.defor12 = p3("defor12"); s3.defor13 = p3("defor13"); s3.defor23 = p3("defor23"); s3.dn = p1("dn");
    s3.dx = c.dx; s3.dy = c.dy; s3.dt = c.dt; s3.c_s = c.c_s; s3.mix_upper_bound = c.mix_upper_bound; s3.mix_isotropic = c.mix_isotropic;
    s3.use_theta_m = c.use_theta_m;
    t.tke = p3("tke_2"); t.bn2 = p3("bn2"); t.xkmh = p3("xkmh"); t.xkmv = p3("xkmv"); t.xkhv = p3("xkhv"); t.u = p3("u_2"); t.v = p3("v_2");
    t.mut = p2("mut"); t.rho = p3("rho"); t.hfx = p2("hfx"); t.ustm = p2("ustm");
    t.c_k = c.c_k; t.cd0 = c.tke_drag_coefficient; t.heat_flux = c.tke_heat_flux; t.isfflx = c.isfflx; t.diff_opt = c.diff_opt;
    for (int j = j_start; j <= j_end; j++)
      for (int k = 1; k <= g.kde - 1; k++)
        for (int i = i_start; i <= i_end; i++) {
          if (stage == 8) tke_km_point(g, t, i, k, j, p3("bn2"), p3("xkmh"), p3("xkmv"), p3("xkhh"), p3("xkhv"));
          else tke_rhs_point(g, t, i, k, j, p3("tke_tend"));
        }
  } else if (stage == 4) {
    const float prandtl = 1.f / 3.0f;   // as Solver::diffusion2_isotropic_km
    const float khdif3 = c.khdif / prandtl, kvdif3 = c.kvdif / prandtl;
    float *xkmh = p3("xkmh"), *xkmv = p3("xkmv"), *xkhh = p3("xkhh"), *xkhv = p3("xkhv");
    for (int j = g.jts; j <= je; j++)
      for (int k = 1; k <= g.kde; k++)
        for (int i = g.its; i <= ie; i++) { A3(xkmh, i, k, j) = c.khdif; A3(xkmv, i, k, j) = c.kvdif; A3(xkhh, i, k, j) = khdif3; A3(xkhv, i, k, j) = kvdif3; }
  } else {
    D2In s;   // as Solver::diffusion2_tendencies
    s.defor11 = p3("defor11"); s.defor22 = p3("defor22"); s.defor33 = p3("defor33"); s.defor12 = p3("defor12"); s.defor13 = p3("defor13");
    s.defor23 = p3("defor23"); s.xkmh = p3("xkmh"); s.xkmv = p3("xkmv"); s.xkhh = p3("xkhh"); s.xkhv = p3("xkhv"); s.zx = p3("zx"); s.zy = p3("zy");
    s.rdz = p3("rdz"); s.rdzw = p3("rdzw"); s.rho = p3("rho"); s.dn = p1("dn");
    s.u = p3("u_2"); s.v = p3("v_2"); s.t = p3("t_2"); s.qv = p3("moist_qv");
    s.ustm = p2("ustm"); s.hfx_in = p2("hfx"); s.qfx = p2("qfx"); s.hfx_out = p2("hfx");
    s.isfflx = c.isfflx; s.use_theta_m = c.use_theta_m; s.cd0 = c.tke_drag_coefficient; s.heat_flux = c.tke_heat_flux;
    const R4 ru = d2_range(g, 'u'), rv = d2_range(g, 'v'), rm = d2_range(g, 'm');
    const float *tb = c.mix_full_fields ? nullptr : p1("t_base"), *qb = c.mix_full_fields ? nullptr : p1("qv_base");
    if (stage == 5) {
      for (int j = g.jts; j <= g.jte; j++)
        for (int k = 1; k <= g.kde - 1; k++)
          for (int i = g.its; i <= g.ite; i++) {
            diff2_u_point(g, s, ru, i, k, j, p3("ru_tendf"));
            diff2_v_point(g, s, rv, i, k, j, p3("rv_tendf"));
          }
      for (int j = rm.j0; j <= rm.j1; j++)
        for (int k = 2; k <= g.kde - 1; k++)
          for (int i = rm.i0; i <= rm.i1; i++) diff2_w_point(g, s, i, k, j, p3("rw_tendf"));
      for (int j = g.jts; j <= g.jte; j++)
        for (int k = 1; k <= g.kde - 1; k++)
          for (int i = g.its; i <= g.ite; i++) diff2_s_point(g, s, rm, 0, p3("t_2"), tb, s.xkhv, s.xkhh, i, k, j, p3("t_tendf"));
    } else {
      // the species go through the k-marching form the device launches (diff2_s_column); theta above through the point form,
      // so both are pinned against the oracle here
      float *var = p3(names[nn - 2]), *tend = p3(names[nn - 1]);
      for (int j = g.jts; j <= g.jte; j++)
        for (int i = g.its; i <= g.ite; i++)
          diff2_s_column(g, s, rm, which == 3 ? 2 : which, var, which == 1 ? qb : nullptr, which == 3 ? s.xkmv : s.xkhv, which == 3 ? s.xkmh : s.xkhh, i, j,
                         tend, which == 3 ? 1 : 0);
    }
  }
  delete gp;
  return 0;
}
