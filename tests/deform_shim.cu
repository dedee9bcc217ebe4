// Host build of wrf_b200/csrc/deform.cuh for tests/test_deform_host.py: runs the product's __host__ __device__ point
// functions of the diffusion preamble (metrics, deformation, 2-D Smagorinsky K) in plain loops on the CPU, so that the
// exact code the kernels of deform.cu execute can be compared with the oracle bit for bit without a GPU.
// Test infrastructure; compiled with nvcc (host code only is run) and -ffp-contract=off.
#include "../wrf_b200/csrc/deform.cuh"
#include <algorithm>

using namespace wg;

enum { PH, PHB, U, V, W, Z, RDZ, RDZW, ZX, ZY, DIV, D11, D22, D33, D12, D13, D23, XKMH, XKMV, XKHH, XKHV,
       MSFTX, MSFTY, MSFUX, MSFUY, MSFVX, MSFVY, FNM, FNP, DN, DNW, UBASE, VBASE, NPTR };

extern "C" int deform_shim_run(int stage, const wrfgpu_dims *dp, const wrfgpu_config *cp, float *const *ptr) {
  const wrfgpu_dims &d = *dp;
  const wrfgpu_config &c = *cp;
  Geo *gp = new Geo();   // ~7 KB
  Geo &g = *gp;
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
  auto p3 = [&](int n) { return ptr[n] - o3; };
  auto p2 = [&](int n) { return ptr[n] - o2; };
  auto p1 = [&](int n) { return ptr[n] - o1; };
  g.msftx = p2(MSFTX); g.msfty = p2(MSFTY); g.msfux = p2(MSFUX); g.msfuy = p2(MSFUY); g.msfvx = p2(MSFVX); g.msfvy = p2(MSFVY);
  for (int k = d.kms; k <= d.kme; k++) { g.fnm[k] = p1(FNM)[k]; g.fnp[k] = p1(FNP)[k]; }
  if (stage == 1) {
    for (int j = g.jts - 1; j <= g.jte; j++)
      for (int k = 1; k <= g.kde; k++)
        for (int i = g.its; i <= g.ite; i++) diff_metrics_point(g, i, k, j, p3(PH), p3(PHB), p3(Z), p3(RDZ), p3(RDZW), p3(ZX), p3(ZY));
  } else if (stage == 2) {
    DefIn in;
    in.u = p3(U); in.v = p3(V); in.w = p3(W); in.rdz = p3(RDZ); in.rdzw = p3(RDZW); in.zx = p3(ZX); in.zy = p3(ZY);
    in.dn = p1(DN); in.dnw = p1(DNW); in.u_base = p1(UBASE); in.v_base = p1(VBASE);
    DefOut out;
    out.div = p3(DIV); out.defor11 = p3(D11); out.defor22 = p3(D22); out.defor33 = p3(D33); out.defor12 = p3(D12); out.defor13 = p3(D13);
    out.defor23 = p3(D23);
    const DefR r = def_ranges(g, c.mix_full_fields);
    for (int j = g.jts; j <= g.jte; j++)
      for (int k = 1; k <= g.kde; k++)
        for (int i = g.its; i <= g.ite; i++) deform_point(g, r, in, out, i, k, j);
  } else {
    int i_start = g.its, i_end = std::min(g.ite, g.ide - 1), j_start = g.jts, j_end = std::min(g.jte, g.jde - 1);   // as Solver::diffusion_prep
    if (g.open_xs) i_start = std::max(g.ids + 1, g.its);
    if (g.open_xe) i_end = std::min(g.ide - 2, g.ite);
    if (g.open_ys) j_start = std::max(g.jds + 1, g.jts);
    if (g.open_ye) j_end = std::min(g.jde - 2, g.jte);
    if (g.periodic_x) { i_start = g.its; i_end = std::min(g.ite, g.ide - 1); }
    SmagIn sm;
    sm.defor11 = p3(D11); sm.defor22 = p3(D22); sm.defor12 = p3(D12); sm.rdzw = p3(RDZW); sm.zx = p3(ZX); sm.zy = p3(ZY);
    sm.dx = c.dx; sm.dy = c.dy; sm.c_s = c.c_s; sm.diff_opt = c.diff_opt;
    for (int j = j_start; j <= j_end; j++)
      for (int k = 1; k <= g.kde - 1; k++)
        for (int i = i_start; i <= i_end; i++) smag2d_point(g, sm, i, k, j, p3(XKMH), p3(XKMV), p3(XKHH), p3(XKHV));
  }
  delete gp;
  return 0;
}
