// common.cuh -- shared declarations of the B200-native solve_em path.
//
// Data layout in HBM: every field keeps WRF's (i,k,j) order with i fastest and
// the caller's memory halo (ims:ime, kms:kme, jms:jme), so an upload/download is
// one contiguous cudaMemcpy of the Fortran array and a warp that spans 32
// consecutive i reads/writes one 128-byte line per field.  Kernel pointers are
// "origin shifted" (base - offset of (0,0,0)), so A3(f,i,k,j) is
// f[i + k*sk + j*sj] with the Fortran indices used as they are.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/wrfgpu.h"

namespace wg {

// share/module_model_constants.F:17-89
#define WG_G 9.81f
#define WG_RD 287.f
#define WG_CP (7.f * 287.f / 2.f)
#define WG_RV 461.6f
#define WG_T0 300.f
#define WG_P0 100000.f
#define WG_CPOVCV (WG_CP / (WG_CP - WG_RD))
#define WG_RCP (WG_RD / WG_CP)
#define WG_RVOVRD (WG_RV / WG_RD)
#define WG_RERADIUS (1.f / 6370.0e03f)
#define WG_CB 25.f

struct Box { int i0, i1, k0, k1, j0, j1; };   // inclusive index box of a launch

// levels a Geo can carry: index 0..WG_NK-1 holds Fortran k = 0..WG_NK-1 (kme <= WG_NK-2, checked at create)
#define WG_NK 128
// GeoBase: extents, strides, flags and scalars (what the strip kernels of bc.cu / halo.cu need);
// Geo adds the per-level coefficient arrays.
struct GeoBase {
  int ids, ide, jds, jde, kds, kde;
  int ims, ime, jms, jme, kms, kme;
  int ips, ipe, jps, jpe;
  int its, ite, jts, jte;          // tile == patch (numtiles = 1)
  int ni, nk, nj;
  int hni, hnj;                    // row / column lengths of the caller's arrays (ime-ims+1, jme-jms+1): the boundary arrays keep them
  long sk, sj;                     // strides of k and j in a 3-D field
  int isk, isj;                    // the same as 32-bit ints: element indices fit (checked at create)
  int periodic_x, periodic_y, sym_xs, sym_xe, sym_ys, sym_ye, open_xs, open_xe, open_ys, open_ye;
  // specified / nested lateral boundaries: spec = config_flags%specified .or. config_flags%nested; nearly
  // every loop-range test of the reference reads "open_?? .or. specified" (o??() below), the radiative blocks open_?? alone
  int spec, spec_zone, relax_zone;
  __host__ __device__ bool oxs() const { return open_xs || spec; }
  __host__ __device__ bool oxe() const { return open_xe || spec; }
  __host__ __device__ bool oys() const { return open_ys || spec; }
  __host__ __device__ bool oye() const { return open_ye || spec; }
  int top_lid, phi_adv_z, damp_opt, w_damping, pert_coriolis, use_theta_m, hypsometric_opt, rk_ord;
  int num_moist;
  float rdx, rdy, dt, cf1, cf2, cf3, cfn, cfn1;
  float epssm, smdiv, emdiv, dampcoef, zdamp, w_crit_cfl, p_top;
  float w_damp_on;                 // vertical Courant number where w_damp starts: w_crit_cfl with IEVA, w_beta = 1 without (:2602-2608)
  float sgn_ts;                    // sign(1,time_step)
  // factor of the upwind term of the momentum / scalar fluxes, horizontal / vertical: sgn_ts for the odd orders (5, 3),
  // 0 for the even ones (h_*_adv_order = 6, v_*_adv_order = 4), which are the same operators without that term
  float sg_hm, sg_vm, sg_hs, sg_vs;
  // base-state sounding (Rayleigh damping, perturbation Coriolis only): origin-shifted global pointers
  const float *u_base, *v_base, *t_base, *z_base;
  // 2-D fields, origin shifted: A2(f,i,j)
  const float *msftx, *msfty, *msfux, *msfuy, *msfvx, *msfvx_inv, *msfvy, *f, *e, *sina, *cosa, *ht;
};
struct Coef {
  // 1-D coefficients the kernels index with the (warp-uniform) Fortran k.  They travel BY VALUE in
  // the kernel parameter block, i.e. in the constant bank: g.c1h[k] is one LDC with a uniform
  // address.  As global loads they shared scoreboards with the k-loops' prefetch loads, so the
  // first use of c1h[k] waited for the whole prefetch batch (ncu r1c: 32 % of k_advance_uv's
  // stall samples on `FMUL c1h*muu`), which serialised what the prefetch was meant to overlap.
  float fnm[WG_NK], fnp[WG_NK], rdnw[WG_NK], rdn[WG_NK], dnw[WG_NK];
  float c1h[WG_NK], c2h[WG_NK], c1f[WG_NK], c2f[WG_NK], c3h[WG_NK], c4h[WG_NK], c3f[WG_NK], c4f[WG_NK];
};
#define WG_NCOEF 13
struct Geo : GeoBase, Coef {};

// 32-bit element indices: one IMAD per term and one widening add per access instead of 64-bit
// multiplies (ncu: the stencil kernels were issue-bound on integer address arithmetic)
#define A3(f, i, k, j) (f)[(i) + (k) * g.isk + (j) * g.isj]
#define A2(f, i, j) (f)[(i) + (j) * g.ni]

typedef float *__restrict__ fp;
typedef const float *__restrict__ cfp;

__device__ __forceinline__ float sign1(float a) { return copysignf(1.0f, a); }

// In-place updates: on sm_100 a store that is issued while a load of the SAME address is still
// in flight (its value not yet consumed) stalls the SM's L1 pipeline until the load returns --
// measured with ncu on k_calc_p_rho: 7 % DRAM utilisation, every warp in long-scoreboard stall,
// 5.3x slower than the same kernel with the store ordered after the load (profiles/).  The
// compiler only orders such a pair in program order, not by completion, so a store whose VALUE
// does not depend on the earlier load (pm1 <- unsmoothed p, sc_tend <- 0) must be given a data
// dependence on it.  after_load(v, x) returns exactly v (bitwise, for every x including NaN):
// fminf(|x|, 0) is +0 and OR-ing +0 into v's bits changes nothing.
__device__ __forceinline__ float after_load(float v, float loaded) {
  return __int_as_float(__float_as_int(v) | __float_as_int(fminf(fabsf(loaded), 0.f)));
}

// k_advance_w reads ~20 coefficients per level; with the warps of an SM at different levels the
// constant cache thrashes (measured: 7.6 -> 9.3 ms per launch when the coefficients moved from L1
// to the constant bank).  It stages the levels 0..kde+1 of the 13 arrays in shared memory once per
// block (uniform-address LDS = broadcast, short scoreboard).  Must be called by every thread of the
// block before any early return.
#ifdef __CUDACC__
__device__ __forceinline__ void coef_to_smem(const Geo &g, Coef &sc) {
  const int nk = g.kde + 2, nt = blockDim.x * blockDim.y * blockDim.z;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const float *src = static_cast<const Coef &>(g).fnm;
  float *dst = sc.fnm;
  for (int n = tid; n < WG_NCOEF * nk; n += nt) {
    const int a = n / nk, k = n - a * nk;
    dst[a * WG_NK + k] = src[a * WG_NK + k];
  }
  __syncthreads();
}
#endif
// per kernel, whichever measured faster (profiles/README.md, r1e, r1u): shared memory for k_advance_w and
// k_advance_mu_t (many coefficient reads per level), the constant bank for k_advance_uv (few)
#define COEF_SMEM __shared__ Coef sc_; coef_to_smem(g, sc_); const Coef &cf = sc_
#define COEF_CONST const Coef &cf = g

#define CUDA_CHECK(x)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess)                                                                \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                 \
  } while (0)

struct Field {
  float *d = nullptr;   // device allocation
  float *o = nullptr;   // origin-shifted pointer
  int kind = 3;         // 3, 2, 1
  size_t n = 0;         // floats of the caller's (host) array: (ime-ims+1) per row
  size_t nd = 0;        // floats of the device allocation: Solver::pitch() per row
  bool raw = false;     // a contiguous array that is not a model field (boundary arrays, fcx / gcx): kind 1 copies, no coefficient refresh
};

// the eight boundary arrays of one field (bdy.cu), sides in the order Y-start, Y-end, X-start, X-end
struct BdyArr { const float *v[4], *t[4]; int nk; };

struct Species {
  std::string f, o, t;  // field, old, tend names
  int adv_opt;
};

struct Solver {
  wrfgpu_dims d;
  wrfgpu_config c;
  Geo g;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::map<std::string, Field> fields;
  std::vector<std::string> names;
  std::vector<Species> species;
  int num_sound_steps = 0;
  float dts = 0.f;
  // km_opt = 2: the prognostic tke rides as one more species of the scalar family (scalar_N = tke_2, scalar_old_N = tke_1,
  // scalar_tend_N = tke_tend, N = num_scalar + 1; the Registry names are aliases), advected with tke_adv_opt
  int tke_index() const { return (c.km_opt == 2 && c.diff_opt == 2) ? c.num_scalar + 1 : 0; }
  long long launches = 0;
  float *d_cfl = nullptr;   // 2 floats: max_vert_cfl, max_horiz_cfl
  bool mt_smem_set = false; // acoustic.cu: the dynamic shared memory size of k_advance_mu_t_st has been set on this device
  // specified / nested lateral boundaries (bdy.cu): grid%dtbc, the time since the boundary arrays were valid (solve_em.F:377;
  // reset by the host when it uploads the next boundary time, wrfgpu_set_dtbc)
  float dtbc = 0.f;
  bool spec() const { return c.specified || c.nested; }
  void add_raw(const std::string &n, size_t count);
  void add_bdy(const std::string &field, int levels);
  bool has_bdy(const std::string &field) const;
  BdyArr bdy_arr(const std::string &field);
  void relax_bdy_dry(bool store_save);
  void spec_bdy_dry();
  void spec_update_uv(float dts_rk);
  void spec_update_tmu(float dts_rk);
  void spec_update_phw(float dts_rk);
  void bdy_scalar(const Species &sp);
  void flow_dep_bdy(const std::string &field);
  void spec_final();
  Box spec_mass_box() const;
  // halo exchange state (HaloState, halo.cu)
  void *comm = nullptr;
  void *snapshot = nullptr;   // history.cu: staging / pinned buffers of wrfgpu_snapshot_*
  void *tile_maps = nullptr;  // tend_tile.cu: the TMA tensor maps of the tile-staged tendency kernel
  int rank = 0, nranks = 1;
  // phase / kernel timing (cudaEvents on `stream`, only when switched on)
  bool timing = false;
  std::map<std::string, double> phase_ms;
  struct KStat { long n = 0; double ms = 0; };
  std::map<std::string, KStat> kstat;
  struct Pending { const char *name; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  bool own_stream = true;
  // thread-block shapes: blk3 = (i, k, j) of the point kernels, blk2 = (i, j) of the column kernels;
  // WRFGPU_BLK3 / WRFGPU_BLK2 ("32x1x4", "32x4") override for tuning (measured r1d: blocks spanning
  // several j rows are 2-6 % slower than 128 x 1).  blk2 must hold 128 threads (__launch_bounds__).
  dim3 blk3{128, 1, 1}, blk2{128, 1, 1};
  bool coef_dirty = true;   // a 1-D field was uploaded since the Geo's coefficient copies were made
  void sync_coef();         // device 1-D fields -> g.fnm ... g.c4f (kernel-parameter copies)
  inline void kpre(const char *name) {
    if (!timing) return;
    Pending p; p.name = name;
    cudaEventCreate(&p.a); cudaEventCreate(&p.b);
    cudaEventRecord(p.a, stream);
    pending.push_back(p);
  }
  inline void kpost() {
    if (!timing) return;
    cudaEventRecord(pending.back().b, stream);
  }
  void flush_timing();

  Solver(const wrfgpu_dims &d, const wrfgpu_config &c, int device);
  ~Solver();
  void add(const std::string &n, int kind);
  size_t pitch() const { return ((size_t)(d.ime - d.ims + 1) + 31) / 32 * 32; }   // device row length in floats
  void copy_in(Field &f, const float *host, cudaMemcpyKind kind);
  void copy_out(const Field &f, float *host, cudaMemcpyKind kind, cudaStream_t st);
  Field &F(const std::string &n);
  float *o(const std::string &n) { return F(n).o; }   // origin-shifted
  void build_geo();
  void rk_params(int rk_step, float &dt_rk, float &dts_rk, int &nsmall) const;

  // phases (solve_em.cu)
  void rk_step_prep(int rk_step);
  void rk_phys_bc_dry_1();
  void init_zero_tendency();
  void diffusion_prep();   // deform.cu: first_rk_step_part2's metrics, deformation, K (km_opt = 4), diff_opt = 2 tendencies
  void diffusion2_rho();   // diffusion2.cu
  void diffusion2_isotropic_km();
  void diffusion2_tendencies();
  void tke_rhs();          // deform.cu
  void rk_tendency(int rk_step, bool zero_save = true);
  void rk_tendency_main(int rk_step, bool zero_save);
  // IEVA (ieva.cu): WW_SPLIT + mass estimates; one implicit column solve (0 scalar, 1 u, 2 v, 3 ph, 4 w)
  void ieva_split(bool dry, const float *u, const float *v, const float *ww, const float *mut, bool active = true);
  void ieva_implicit(int which, float *tendency, const float *q_old, const float *mu_old, const float *mu_new, float dt_rk);   // zero_save = false: small_step_prep(rk_step, true) writes those zeros
  void small_step_prep(int rk_step, bool zero_rest = false);
  void advance_uv(int rk_step);
  void advance_uv_fused(int rk_step, int iteration);
  void sumflux_w(int rk_step, int iteration);
  void advance_mu_t(int rk_step, bool final_ave = false);   // final_ave: see k_advance_mu_t<true>
  void advance_w(int rk_step);
  void advance_w_fused(int rk_step, int iteration);
  void sumflux(int rk_step, int iteration);
  void calc_p_rho(int iteration);
  void small_step_finish(int rk_step);
  void rk_scalar_step(int rk_step);
  void calc_p_rho_phi();
  void rk_phys_bc_dry_2();
  void microphysics();
  void end_of_step();
  void start_domain();
  void step();
  // bc.cu
  struct BcItem { std::string name; char var; };
  void bc(const std::vector<BcItem> &items);
  // halo.cu
  void halo(const std::string &name);
  // Overlap of an exchange with the kernel that produces its data (north star: "overlapped with interior-tile compute"):
  // the producer runs first on the RING of the patch (the points within 3 of a side that has a neighbour: everything a send
  // strip holds), then the exchange goes to a second stream while the producer does the interior on the main stream; the
  // main stream joins before the first consumer of the halo.  `launch(box)` must give the same values for any partition
  // of the box (the column kernels of the acoustic loop do: they read no neighbour of the fields they update).
  // The ring strips and the exchange run on high-priority streams (their few blocks are scheduled ahead of the interior's
  // pending blocks), one stream per strip so that the four latency-bound column marches run side by side.
  cudaStream_t hstream = nullptr, rstream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ring[4] = {nullptr, nullptr, nullptr, nullptr};
  // Measured on 4 and 8 B200 (profiles/README.md, r2v-r2x): with the peer-window transport an acoustic exchange costs
  // 55-90 us, less than the four latency-bound ring launches cost the producer, so the split loses 0.6-1 % of the step.
  // It is therefore OFF unless WRFGPU_OVERLAP=1; tests/test_multi_gpu.py runs it for the bits.
  bool overlap = false;
  void overlap_streams();
  int ring_split(const Box &full, int w, Box ring[4], Box &inner);
  template <class F> void overlapped(const Box &full, const char *halo_name, F launch) {
    Box ring[4], inner;
    const int n = (nranks > 1 && overlap) ? ring_split(full, 3, ring, inner) : 0;
    if (n == 0) { launch(full); halo(halo_name); return; }
    overlap_streams();
    CUDA_CHECK(cudaEventRecord(ev_fork, stream));
    cudaStream_t main_stream = stream;
    for (int m = 0; m < n; m++) {      // ring first, on the priority streams
      CUDA_CHECK(cudaStreamWaitEvent(rstream[m], ev_fork, 0));
      stream = rstream[m];
      try { launch(ring[m]); } catch (...) { stream = main_stream; throw; }
      stream = main_stream;
      CUDA_CHECK(cudaEventRecord(ev_ring[m], rstream[m]));
      CUDA_CHECK(cudaStreamWaitEvent(hstream, ev_ring[m], 0));
    }
    if (inner.i0 <= inner.i1 && inner.j0 <= inner.j1) launch(inner);      // the bulk, on the main stream
    stream = hstream;                  // the exchange, behind the ring
    try { halo(halo_name); } catch (...) { stream = main_stream; throw; }
    stream = main_stream;
    CUDA_CHECK(cudaEventRecord(ev_join, hstream));
    CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0));
  }
};

// launch helpers -------------------------------------------------------------
inline dim3 grid3(const Box &b, const dim3 &t) {   // thread block t = (i, k, j)
  return dim3((unsigned)((b.i1 - b.i0 + t.x) / t.x), (unsigned)((b.k1 - b.k0 + t.y) / t.y), (unsigned)((b.j1 - b.j0 + t.z) / t.z));
}
// the tile [its, ite] x [k0, k1] x [jts, jte] as groups of four columns counted from ims (the *_v4 kernels)
inline Box box_v4(const Geo &q, int k0, int k1) { return Box{(q.its - q.ims) / 4, (q.ite - q.ims) / 4, k0, k1, q.jts, q.jte}; }
inline bool empty(const Box &b) { return b.i1 < b.i0 || b.k1 < b.k0 || b.j1 < b.j0; }

#define WG_BX 128
// 3-D launch: kernel(g, box, args...) with one thread per (i,k,j)
#define LAUNCH3(S, kern, box, ...)                                                   \
  do {                                                                               \
    Box b_ = (box);                                                                  \
    if (!empty(b_)) {                                                                \
      if ((S)->coef_dirty) (S)->sync_coef();                                         \
      (S)->kpre(#kern);                                                              \
      kern<<<grid3(b_, (S)->blk3), (S)->blk3, 0, (S)->stream>>>((S)->g, b_, __VA_ARGS__); \
      (S)->kpost();                                                                  \
      (S)->launches++;                                                               \
    }                                                                                \
  } while (0)
// column launch: one thread per (i,j)
#define LAUNCH2(S, kern, box, ...)                                                   \
  do {                                                                               \
    Box b_ = (box);                                                                  \
    if (!empty(b_)) {                                                                \
      /* strips a few columns wide (the ring of an overlapped exchange): 4 x 32 threads instead of 128 x 1 */ \
      const dim3 t_ = (b_.i1 - b_.i0 < 4) ? dim3(4, 32, 1) : (S)->blk2;             \
      dim3 gr_((unsigned)((b_.i1 - b_.i0 + t_.x) / t_.x), (unsigned)((b_.j1 - b_.j0 + t_.y) / t_.y)); \
      if ((S)->coef_dirty) (S)->sync_coef();                                         \
      (S)->kpre(#kern);                                                              \
      kern<<<gr_, t_, 0, (S)->stream>>>((S)->g, b_, __VA_ARGS__);                    \
      (S)->kpost();                                                                  \
      (S)->launches++;                                                               \
    }                                                                                \
  } while (0)

// 16-byte accesses of the streaming kernels (k_ssp_3d_v4, k_ssf_v4, k_prep_3d_v4, ...): a thread owns a 16-byte aligned
// group of four columns of the padded device row; p points at lane 0
#ifdef __CUDACC__
__device__ __forceinline__ void ld4(const float *p, float v[4]) {
  const float4 q = *reinterpret_cast<const float4 *>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
// lanes lo..hi (0..3, inclusive) of the group are to be written
__device__ __forceinline__ void st4(float *p, const float v[4], int lo, int hi) {
  if (lo == 0 && hi == 3) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else {
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (q >= lo && q <= hi) p[q] = v[q];
  }
}
#endif
// column kernels: thread -> (group of four columns, j)
#define TID2_V4                                               \
  const int gi_ = b.i0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y; \
  if (gi_ > b.i1 || j > b.j1) return;                         \
  const int i = g.ims + 4 * gi_;                              \
  const int lo = max(g.its - i, 0), him = min(min(g.ite, g.ide - 1) - i, 3);
// thread -> (group of four columns, k, j) of a box whose i0 / i1 count groups from the start of the memory row
#define TID3_V4                                               \
  const int gi_ = b.i0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int k = b.k0 + blockIdx.y * blockDim.y + threadIdx.y; \
  const int j = b.j0 + blockIdx.z * blockDim.z + threadIdx.z; \
  if (gi_ > b.i1 || k > b.k1 || j > b.j1) return;             \
  const int i = g.ims + 4 * gi_;                              \
  const int lo = max(g.its - i, 0), hi = min(g.ite - i, 3), him = min(min(g.ite, g.ide - 1) - i, 3);
#define TID3                                              \
  const int i = b.i0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int k = b.k0 + blockIdx.y * blockDim.y + threadIdx.y; \
  const int j = b.j0 + blockIdx.z * blockDim.z + threadIdx.z; \
  if (i > b.i1 || k > b.k1 || j > b.j1) return;
#define TID2                                              \
  const int i = b.i0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y; \
  if (i > b.i1 || j > b.j1) return;

void halo_destroy(Solver *s);   // halo.cu
void halo_max2(Solver *s, float *d2, float out[2]);   // halo.cu: max over the ranks of the mesh of two non-negative device floats
void snapshot_destroy(Solver *s);   // history.cu
// advect_fast.cu
Box interior_box(const Geo &q);
int edge_strips(const Box &tile, const Box &in, Box out[4]);
void launch_advect_scalar_fast(Solver *s, bool theta, const Box &in, cfp q, cfp ru, cfp rv, cfp rom, cfp mut, cfp t_tendf,
                               cfp h_diabatic, fp tend);

// pd_fast.cu: pass 1 = limiter factor on bx -> out, pass 2 = limited flux divergence on bx -> out
void launch_pd_fast(Solver *s, int pass, const Box &bx, cfp q, cfp q_old, float dt, fp out);
// diffusion.cu: adds the constant-K diffusion of `field` to `tendency` (which = 0 u, 1 v, 2 w, 3 theta, 4 qv, 5 other scalar)
// sixth_order_diffusion of `field` added to `tendency` (which as above; anything >= 3 = mass points)
void launch_diff6(Solver *s, int which, const float *field, float *tendency, const float *mut, float dt);
void launch_diffuse(Solver *s, int which, cfp field, fp tendency, float xk, float kvdif, cfp xkarr = nullptr);
// diffusion2.cu: diff_opt = 2 with constant K = 0 and no surface fluxes adds exact zeros
bool diff2_is_identity(const wrfgpu_config &c);
// tend_tile.cu: all five dry tendencies of the interior box in one TMA-staged kernel
bool tend_tile_enabled();
void launch_tend_tile(Solver *s, const Box &in);
void tend_tile_destroy(Solver *s);
// mom_fast.cu: which = 0 ru_tend, 1 rv_tend, 2 rw_tend on the interior box
void launch_mom_fast(Solver *s, int which, const Box &in);

}  // namespace wg

// the opaque handle of include/wrfgpu.h
struct wrfgpu_solver {
  wg::Solver s;
  wrfgpu_solver(const wrfgpu_dims &d, const wrfgpu_config &c, int dev) : s(d, c, dev) {}
};
extern thread_local std::string g_wrfgpu_err;
