// halo.cu -- the Registry-generated HALO_EM_* / PERIOD_* exchanges of the solve_em path
// (Registry/Registry.EM_COMMON:3418-3551, inc/HALO_EM_*.inc as emitted by tools/gen_comms.c:296-672)
// and their RSL_LITE mechanics (external/RSL_LITE/c_code.c:306-1076, period.c, module_dm.F:2761)
// restated for one-process-per-GPU execution.
//
// What is kept from the reference, because N-GPU == 1-GPU bit for bit depends on it:
//   * patch ownership (decomp.cu) and which fields travel in which exchange, with which width;
//   * Y exchange first over i in [ips-w, ipe+w] clipped to the domain, then X over j in
//     [jps-w, jpe+w], so corner points arrive in two hops (c_code.c:460-477, :587);
//   * a periodic dimension that is split over ranks is closed by an exchange between the first and
//     the last rank of the mesh row/column (period.c); one that lives on a single rank is a local
//     copy done by set_physical_bc (share/module_bc.F:750) -- bc.cu here.
// What is different: strips are packed by ONE kernel per phase for all fields and both directions
// straight into device buffers and move GPU-to-GPU -- ncclSend/ncclRecv over NVLink when ranks are
// processes (bench.py under torchrun, wrf.exe under mpirun), or a peer-memory read of the
// neighbour's send buffer when the ranks are solvers of one process (wrfgpu_comm_init_local).
// A HALO_EM_x and the PERIOD_* that follows it in solve_em.F are one exchange here: the periodic
// neighbour is just another peer, with the wider PERIOD width (e.g. HALO_EM_A 1, PERIOD_BDY_EM_A 2).
#include "common.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>

namespace wg {

// ---------------------------------------------------------------------------------------------
// exchange table
struct HaloDef { const char *name; int width; const char *f3; const char *f2; };
// '@moist' -> moist_q*, '@moist_old' -> moist_old_q*, '@scalar' -> scalar_n, '@scalar_old' -> scalar_old_n
static const HaloDef kHalos[] = {
    // start of step: initialize_moist_old (solve_em.F:520-550); HALO_EM_MOIST_OLD_E_7 :3480 + PERIOD :3530, width 4
    {"HALO_EM_MOIST_OLD_E_7", 4, "@moist_old,@scalar_old", ""},
    // after rk_step_prep (:708,:715): HALO_EM_A :3425 (1) + PERIOD_BDY_EM_A :3538 (2)
    {"HALO_EM_A", 2, "ru,rv,rw,ww,php,alt,al,p,rho", "muu,muv,mut"},
    // after small_step_prep (:1176-1177): HALO_EM_B :3447 (1) + PERIOD_BDY_EM_B :3542 (2, adds ru_tend,rv_tend)
    {"HALO_EM_B", 2, "ph_2,al,p,t_1,t_save,u_save,v_save,php,alt,pb,ru_tend,rv_tend", "mu_1,mu_2,mudf"},
    // after advance_uv (:1382): HALO_EM_C :3449 (1); the periodic twin is PERIOD_BDY_EM_E :3536 (2)
    {"HALO_EM_C", 2, "u_2,v_2", ""},
    // after calc_p_rho (:1659-1660): HALO_EM_C2 :3450 (1) + PERIOD_BDY_EM_B3 :3543 (2)
    {"HALO_EM_C2", 2, "ph_2,al,p", "mu_2,muts,mudf"},
    // before scalar advection (:2179-2181): HALO_EM_D :3451 (2) + PERIOD_EM_DA :3548 (2)
    {"HALO_EM_D", 2, "ru_m,rv_m,ww_m", "mut,muts"},
    // end of RK substeps 1,2 (:3222-3236): HALO_EM_D2_5 :3454 (3) + PERIOD_BDY_EM_D :3546 (3); the species go
    // with HALO_EM_MOIST_E_7 / SCALAR_E_7 (:3359-3447, width 4 because advect_scalar_pd needs a tile+1 ring)
    {"HALO_EM_D2_5", 3, "u_2,v_2,w_2,t_2,ph_2,al", "mu_2"},
    {"HALO_EM_MOIST_E_7", 4, "@moist,@scalar", ""},
    // after microphysics (:4109-4110): HALO_EM_THETAM :3476 (3) + PERIOD_EM_THETAM :3551
    {"HALO_EM_THETAM", 3, "h_diabatic", ""},
    // end of step (:4437-4449): HALO_EM_D3_5 :3456 (3) + PERIOD_BDY_EM_D3 :3547, species :3521 (4)
    {"HALO_EM_D3_5", 3, "u_1,u_2,v_1,v_2,w_1,w_2,t_1,t_2,ph_1,ph_2", "mu_1,mu_2"},
    // IEVA only, before rk_scalar_tend's WW_SPLIT on the last RK substep (module_em.F:1222): that call reads u_1, v_1 one
    // point beyond the tile, but the Registry exchanges the _1 fields only at the end of a step (HALO_EM_D3_5 / E_5), before
    // small_step_prep refreshes their interior -- under DM the reference reads one-step-old halo values there and its IEVA
    // results depend on the decomposition.  This exchange (not in the Registry) makes N ranks equal 1 rank.
    {"HALO_EM_IEVA", 2, "u_1,v_1", ""},
    // diffusion preamble of the first RK substep (first_rk_step_part2, module_first_rk_step_part2.F), closures that compute K:
    // metrics (:446-447): HALO_EM_TKE_C :3438 (1) + PERIOD_BDY_EM_A1 :3539 (3); its u_2, v_2 are current since HALO_EM_E_5
    {"HALO_EM_TKE_C", 3, "z,zx,zy,rdz,rdzw", ""},
    // the 2-D part of HALO_EM_TKE_C (ustm; diff_opt = 2 only: the surface drag averages it to the u and v points)
    {"HALO_EM_TKE_C_SFC", 1, "", "ustm"},
    // deformation (:571): HALO_EM_TKE_D :3439 (1)
    {"HALO_EM_TKE_D", 1, "defor11,defor22,defor33,defor12,defor13,defor23,div", ""},
    // eddy viscosities (:644-650, :759): HALO_EM_TKE_E :3440 (1) + PERIOD_BDY_EM_PHY_BC :3540 (2) + HALO_EM_PHYS_DIFFUSION :3432 (1)
    {"HALO_EM_TKE_E", 2, "xkmv,xkmh,xkhv,xkhh,div,defor11,defor22,defor33,defor12,defor13,defor23,rho", ""},
    // very end (:4781-4846): HALO_EM_E_5 :3458 (3) after set_w_surface -- NOT followed by a PERIOD_* closure or a set_physical_bc
    // in the reference, so the periodic ghost zone keeps the w_2(k=1) of before set_w_surface (what the boundary fill of a
    // single rank leaves there as well): negative width = no closure.  cal_deform_and_div is the one reader of that value.
    {"HALO_EM_E_5", -3, "u_1,u_2,v_1,v_2,w_1,w_2,t_1,t_2,ph_1,ph_2", "mu_1,mu_2"},
    // start_domain_em (dyn_em/start_em.F): HALO_EM_INIT_1..5 carry the base state, map factors and both time levels
    {"HALO_EM_INIT", 4,
     "phb,pb,alb,p,al,alt,t_init,u_1,u_2,v_1,v_2,w_1,w_2,t_1,t_2,ph_1,ph_2,@moist,@moist_old,@scalar,@scalar_old",
     "mub,ht,msftx,msfty,f,e,sina,cosa,msfux,msfuy,msfvx,msfvx_inv,msfvy,mu_1,mu_2"},
};

// Registry width of the HALO_* part where it is narrower than the PERIOD_* part it is merged with (Registry.EM_COMMON:3425,
// :3447-3450: 4-point stencils = width 1): a neighbour inside the domain gets the HALO width, the neighbour across a split
// periodic edge the PERIOD width.  Halves the bytes of the four most frequent exchanges.
static int regular_width(const std::string &name, int width) {
  for (const char *n : {"HALO_EM_A", "HALO_EM_B", "HALO_EM_C", "HALO_EM_C2"})
    if (name == n) return 1;
  return width;
}

static void split(const std::string &s, std::vector<std::string> &out) {
  size_t a = 0;
  while (a < s.size()) {
    size_t b = s.find(',', a);
    if (b == std::string::npos) b = s.size();
    if (b > a) out.push_back(s.substr(a, b - a));
    a = b + 1;
  }
}

// expands the '@' tokens for the species counts of `c`
static bool halo_fields(const wrfgpu_config &c, const std::string &name, int &width, std::vector<std::string> &f3,
                        std::vector<std::string> &f2) {
  static const char *mn[] = {"qv", "qc", "qr"};
  const int nsc = c.num_scalar + ((c.km_opt == 2 && c.diff_opt == 2) ? 1 : 0);   // + the tke species (Solver::tke_index)
  for (const HaloDef &h : kHalos) {
    if (name != h.name) continue;
    width = h.width;
    std::vector<std::string> raw;
    split(h.f3, raw);
    for (auto &t : raw) {
      if (t == "@moist") for (int s = 0; s < c.num_moist; s++) f3.push_back(std::string("moist_") + mn[s]);
      else if (t == "@moist_old") for (int s = 0; s < c.num_moist; s++) f3.push_back(std::string("moist_old_") + mn[s]);
      else if (t == "@scalar") for (int s = 1; s <= nsc; s++) f3.push_back("scalar_" + std::to_string(s));
      else if (t == "@scalar_old") for (int s = 1; s <= nsc; s++) f3.push_back("scalar_old_" + std::to_string(s));
      else f3.push_back(t);
    }
    split(h.f2, f2);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// strip geometry.  dir: 0 = -y, 1 = +y (Y phase), 2 = -x, 3 = +x (X phase)
struct Strip { int peer; int is, ie, js, je; };
// A negative width means "HALO only": an exchange the reference does not follow with a PERIOD_* closure, so the ghost zone
// beyond a periodic domain edge keeps what the last closure / set_physical_bc put there (see HALO_EM_E_5 in the table).
// hw: the width towards a neighbour inside the domain (0 = the same as w, the width across a periodic edge)
static bool halo_box(const wrfgpu_dims &d, const wrfgpu_config &c, int w, int dir, bool send, Strip &s, int hw = 0) {
  const int ntx = std::max(1, c.ntasks_x), nty = std::max(1, c.ntasks_y);
  const int mx = c.mytask % ntx, my = c.mytask / ntx;
  const bool closure = w > 0;
  if (w < 0) w = -w;
  if (hw <= 0) hw = w;
  bool wrap = false;
  int px = mx, py = my;
  switch (dir) {
    case 0: if (my > 0) py = my - 1; else if (c.periodic_y && nty > 1 && closure) { py = nty - 1; wrap = true; } else return false; break;
    case 1: if (my < nty - 1) py = my + 1; else if (c.periodic_y && nty > 1 && closure) { py = 0; wrap = true; } else return false; break;
    case 2: if (mx > 0) px = mx - 1; else if (c.periodic_x && ntx > 1 && closure) { px = ntx - 1; wrap = true; } else return false; break;
    default: if (mx < ntx - 1) px = mx + 1; else if (c.periodic_x && ntx > 1 && closure) { px = 0; wrap = true; } else return false; break;
  }
  s.peer = py * ntx + px;
  if (!wrap) w = hw;
  // extent along the exchanged dimension: [lo, hi] of the patch, domain [ds, de]
  const bool isy = dir < 2;
  const int ps = isy ? d.jps : d.ips, pe = isy ? d.jpe : d.ipe, ds = isy ? d.jds : d.ids, de = isy ? d.jde : d.ide;
  int a, b;
  const bool low = (dir == 0 || dir == 2);
  if (!wrap) {
    if (low) { if (send) { a = ps; b = ps + w - 1; } else { a = ps - w; b = ps - 1; } }
    else { if (send) { a = pe - w + 1; b = pe; } else { a = pe + 1; b = pe + w; } }
  } else {
    // periodic closure: dat(de+n) = dat(ds+n), n = 0..w  and  dat(ds-n) = dat(de-n), n = 1..w
    // (share/module_bc.F:750-790 for the one-patch case; period.c for the split one).  The high side
    // carries one point more because the point at `de` itself (the duplicate of `ds`) is replaced too.
    if (low) { if (send) { a = ds; b = ds + w; } else { a = ds - w; b = ds - 1; } }
    else { if (send) { a = de - w; b = de - 1; } else { a = de; b = de + w; } }
  }
  if (isy) {
    s.js = a; s.je = b;
    s.is = std::max(d.ips - w, d.ids); s.ie = std::min(d.ipe + w, d.ide);   // c_code.c:460-477
    // a periodic x dimension that lives on this rank: the ghost columns were filled by set_physical_bc
    // one row into the y halo only; they are refreshed by the next bc fill, not exchanged
  } else {
    s.is = a; s.ie = b;
    // rows include the y halo just received, so corners arrive in two hops (c_code.c:587 JMAX/JMIN clip to
    // the domain; with a split periodic y the ghost rows beyond the domain came from the wrap and go along)
    const int ext = (c.periodic_y && nty > 1 && closure) ? w : 0;
    s.js = std::max(d.jps - w, d.jds - ext); s.je = std::min(d.jpe + w, d.jde + ext);
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// device pack / unpack: one launch moves every field of an exchange for up to two strips
#define HALO_MAX 40
struct HaloArgs {
  float *p[HALO_MAX];       // origin-shifted field pointers
  long off[HALO_MAX];       // element offset of the field inside a strip buffer, per unit (ni*nj) area
  char is3d[HALO_MAX];
  int n, nk;
  // two strips (one per direction of the phase)
  int is[2], ie[2], js[2], je[2];
  float *buf[2];
  int active[2];
};

// One thread per strip element, flattened over (row = (j,k), i) so that the 3-column strips of the
// x phase fill their warps too (a block per (j,k) row left 125 of 128 threads idle there: 20 ms of
// pack + unpack per step on 8 GPUs, r1h).  The buffer layout is unchanged: element e of a field's
// strip is ((jj * nk + kk) * ni + ii).
template <bool UNPACK>
__global__ void k_halo_copy(const GeoBase g, HaloArgs a) {
  const int f = blockIdx.z % a.n, sd = blockIdx.z / a.n;
  if (!a.active[sd]) return;
  const int ni = a.ie[sd] - a.is[sd] + 1, nj = a.je[sd] - a.js[sd] + 1;
  const bool is3d = a.is3d[f];
  const int nk = is3d ? a.nk : 1;
  const int total = ni * nj * nk;
  float *buf = a.buf[sd] + a.off[f] * (long)ni * nj;
  float *fld = a.p[f];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int row = e / ni, ii = e - row * ni;
    const int jj = row / nk, kk = row - jj * nk;
    const int i = a.is[sd] + ii, j = a.js[sd] + jj, k = 1 + kk;
    const long x = is3d ? ((long)i + (long)k * g.sk + (long)j * g.sj) : ((long)i + (long)j * g.ni);
    if (UNPACK) fld[x] = buf[e];
    else buf[e] = fld[x];
  }
}

// A periodic dimension that lives on ONE rank is closed locally by set_physical_bc, but only on the
// rows [jts-1, jte+1] (share/module_bc.F:752).  When the other dimension is split, the deeper halo
// rows a rank has just received (and the ghost rows of a periodic wrap) would keep stale ghost
// columns -- the corner points the positive-definite limiter's tile+1 ring reads.  period.c lets
// "the extent in other dimension go into boundary region" for the same purpose; here the received
// rows/columns get the same periodic copy set_physical_bc applies (dat(ds-n) = dat(de-n), n = 1..4;
// dat(de+n) = dat(ds+n), n = 0..4), so that N ranks reproduce one rank bit for bit.
// along_x: close x on rows [a0,a1]; else close y on columns [a0,a1]
__global__ void k_halo_close(const GeoBase g, HaloArgs a, int along_x, int a0, int a1) {
  const int f = blockIdx.z;
  const bool is3d = a.is3d[f];
  const int nk = is3d ? a.nk : 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;      // index over the (a0..a1) x k lines
  const int nl = a1 - a0 + 1;
  if (t >= nl * nk) return;
  const int l = a0 + t / nk, k = 1 + t % nk;
  float *fld = a.p[f];
  const int ds = along_x ? g.ids : g.jds, de = along_x ? g.ide : g.jde;
  auto at = [&](int m) -> float & {
    const int i = along_x ? m : l, j = along_x ? l : m;
    return is3d ? fld[(long)i + (long)k * g.sk + (long)j * g.sj] : fld[(long)i + (long)j * g.ni];
  };
  for (int n = 1; n <= 4; n++) at(ds - n) = at(de - n);
  for (int n = 0; n <= 4; n++) at(de + n) = at(ds + n);
}

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time so that the library also loads where no NCCL exists (CPU test box)
struct Id128 { char b[128]; };   // ncclUniqueId: a 128-byte struct passed BY VALUE
struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static Nccl &nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *env = getenv("WRFGPU_NCCL_LIB");
    for (const char *name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
      n.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) break;
    }
    if (!n.h) return;
    n.GetUniqueId = (int (*)(void *))dlsym(n.h, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(n.h, "ncclCommInitRank");
    n.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(n.h, "ncclSend");
    n.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(n.h, "ncclRecv");
    n.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(n.h, "ncclAllReduce");
    n.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(n.h, "ncclAllGather");
    n.GroupStart = (int (*)())dlsym(n.h, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.h, "ncclGroupEnd");
    n.CommDestroy = (int (*)(void *))dlsym(n.h, "ncclCommDestroy");
    n.GetErrorString = (const char *(*)(int))dlsym(n.h, "ncclGetErrorString");
  });
  return n;
}
static void nccl_check(int rc, const char *what) {
  if (rc != 0) {
    Nccl &n = nccl();
    throw std::runtime_error(std::string("wrfgpu: NCCL ") + what + " failed: " + (n.GetErrorString ? n.GetErrorString(rc) : "?"));
  }
}

// ---------------------------------------------------------------------------------------------
// ranks that are solvers of ONE process (one host thread each): a host barrier orders the
// exchange, the data moves by reading the neighbour's send buffer from the unpack kernel
struct LocalGroup {
  std::vector<Solver *> ranks;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  long generation = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const long gen = generation;
    if (++waiting == (int)ranks.size()) { waiting = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != gen; });
  }
};

struct HaloState {
  void *comm = nullptr;               // ncclComm_t
  LocalGroup *local = nullptr;
  bool owns_local = false;
  float *send[4] = {nullptr, nullptr, nullptr, nullptr};
  float *recv[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;                     // floats per buffer
  cudaEvent_t packed = nullptr, consumed = nullptr;
  long long exchanges = 0, bytes_sent = 0;
  float max2[2] = {0.f, 0.f};         // halo_max2's host copy (local groups)
  // direct transport (see "peer windows" below)
  bool direct = false;
  char *win = nullptr;                // this rank's window: flags + 4 directions x 2 parities of receive buffers
  size_t win_cap = 0;                 // floats per receive buffer
  std::vector<char *> peer_win;       // the windows of the other ranks as this rank addresses them (nullptr: not opened)
  std::vector<bool> peer_ipc;         // opened with cudaIpcOpenMemHandle (to be closed)
  unsigned seq[4] = {0, 0, 0, 0};     // exchanges done per direction; the n-th uses receive buffer n & 1
  unsigned *d_count = nullptr;        // 4 block counters of the put / get kernels
};

// ---------------------------------------------------------------------------------------------
// Direct transport over peer windows: no library call and no host synchronisation on the exchange path.
// Every rank owns a window in its HBM: per direction (0 = -y, 1 = +y, 2 = -x, 3 = +x) a `ready` and an `ack` counter and two
// receive buffers.  Neighbours address each other's windows directly: CUDA IPC handles between processes (one per GPU:
// bench.py under torchrun, wrf.exe under mpirun; the handles travel once through ncclAllGather), plain pointers between the
// solvers of one process.  The n-th exchange of a rank towards direction q:
//   k_halo_put  waits until ack[q] >= n - 2 (the neighbour has emptied the buffer of exchange n - 2, which had the same
//               parity), packs the strip straight INTO THE NEIGHBOUR'S receive buffer (q^1, n & 1) -- stores over NVLink --
//               and, when the last block is through, publishes ready[q^1] = n in the neighbour's window (release, system scope);
//   k_halo_get  waits for its own ready[q] >= n, unpacks its receive buffer (q, n & 1) and, last block, writes
//               ack[q^1] = n into the neighbour's window.
// Both kernels run on the solver's stream like any other kernel of the step, so an exchange is two launches per phase and
// costs NVLink latency instead of two NCCL groups (8 GPUs, r1: ~156 ncclSend/Recv kernels and their rendezvous per step).
// Every spin is bounded and ends in a trap: a lost neighbour becomes a CUDA error, never a hung GPU.
#define WIN_HDR 256
__host__ __device__ inline unsigned *win_ready(char *w, int dir) { return reinterpret_cast<unsigned *>(w) + dir; }
__host__ __device__ inline unsigned *win_ack(char *w, int dir) { return reinterpret_cast<unsigned *>(w + 64) + dir; }
__host__ __device__ inline float *win_buf(char *w, int dir, unsigned parity, size_t cap) {
  return reinterpret_cast<float *>(w + WIN_HDR) + ((size_t)dir * 2 + parity) * cap;
}

struct LinkArgs {                 // per strip (two per phase)
  unsigned *wait_flag[2];         // local flag to wait on (put: ack[q]; get: ready[q]); nullptr = no wait
  unsigned wait_val[2];
  unsigned *post_flag[2];         // remote flag the last block writes (put: neighbour's ready[q^1]; get: neighbour's ack[q^1])
  unsigned post_val[2];
  unsigned *count[2];             // local block counter (zero between launches)
  unsigned nblocks[2];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The waits of a phase as a kernel of their own (two threads): used between solvers that SHARE a GPU, where hundreds of
// spinning blocks inside the copy kernels could hold the registers a neighbour's compute kernel (256 threads x 128
// registers, 91 KB of shared memory) needs to run and reach its own put -- a deadlock by starvation.  One warp cannot.
__global__ void k_halo_wait(LinkArgs l) {
  const int sd = threadIdx.x;
  if (sd < 2 && l.wait_flag[sd]) {
    int spin = 0;
    while ((int)(ld_acquire_sys(l.wait_flag[sd]) - l.wait_val[sd]) < 0) {
      if (++spin > (1 << 26)) __trap();
      __nanosleep(64);
    }
  }
}

// PUT: fields -> a.buf (the neighbour's receive buffer); GET: a.buf (my receive buffer) -> fields.  MODE 0 = put, 1 = get,
// 2 = both in one launch (processes: blockIdx.y = 0 packs with `a`, `l`; blockIdx.y = 1 unpacks with `a2`, `l2`; the put
// blocks have the lower block indices, so they are dispatched first, and nothing in a launch waits for the same launch).
struct HaloPair { HaloArgs a; LinkArgs l; };
#ifndef HU
#define HU 8
#endif
#ifndef HALO_TPB
#define HALO_TPB 512
#endif
template <int MODE>
__global__ void k_halo_direct(const GeoBase g, const HaloPair pp, const HaloPair pg) {
  const bool GET = MODE == 1 || (MODE == 2 && blockIdx.y == 1);
  const HaloArgs &a = GET ? pg.a : pp.a;
  const LinkArgs &l = GET ? pg.l : pp.l;
  const int f = blockIdx.z % a.n, sd = blockIdx.z / a.n;
  if (!a.active[sd]) return;
  if (threadIdx.x == 0 && l.wait_flag[sd]) {
    int spin = 0;
    while ((int)(ld_acquire_sys(l.wait_flag[sd]) - l.wait_val[sd]) < 0) {
      if (++spin > (1 << 26)) __trap();      // seconds: the neighbour is gone
      __nanosleep(64);
    }
  }
  __syncthreads();
  const int ni = a.ie[sd] - a.is[sd] + 1, nj = a.je[sd] - a.js[sd] + 1;
  const bool is3d = a.is3d[f];
  const int nk = is3d ? a.nk : 1;
  const int total = ni * nj * nk;
  float *buf = a.buf[sd] + a.off[f] * (long)ni * nj;
  float *fld = a.p[f];
  // HU elements per thread and trip: the loads of a trip are independent, so their latencies overlap
  const int stride = gridDim.x * blockDim.x;
  for (int e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += HU * stride) {
    long x[HU];
    float v[HU];
#pragma unroll
    for (int u = 0; u < HU; u++) {
      const int e = e0 + u * stride;
      const int ee = e < total ? e : e0;
      const int row = ee / ni, ii = ee - row * ni;
      const int jj = row / nk, kk = row - jj * nk;
      const int i = a.is[sd] + ii, j = a.js[sd] + jj, k = 1 + kk;
      x[u] = is3d ? ((long)i + (long)k * g.sk + (long)j * g.sj) : ((long)i + (long)j * g.ni);
      v[u] = GET ? __ldcg(buf + ee) : fld[x[u]];       // a receive buffer is written by the neighbour's GPU: never through L1
    }
#pragma unroll
    for (int u = 0; u < HU; u++) {
      const int e = e0 + u * stride;
      if (e < total) { if (GET) fld[x[u]] = v[u]; else buf[e] = v[u]; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                  // cumulative: the block's stores are ordered before the counter / flag
    if (atomicAdd(l.count[sd], 1u) == l.nblocks[sd] - 1) {
      *l.count[sd] = 0;
      __threadfence_system();
      st_release_sys(l.post_flag[sd], l.post_val[sd]);
    }
  }
}

static HaloState &hs(Solver *s) {
  if (!s->comm) s->comm = new HaloState();
  return *(HaloState *)s->comm;
}

static void ensure_buffers(Solver *s, size_t nfloats) {
  HaloState &h = hs(s);
  if (nfloats <= h.cap) return;
  CUDA_CHECK(cudaStreamSynchronize(s->stream));
  const size_t cap = nfloats + nfloats / 4;
  for (int d = 0; d < 4; d++) {
    if (h.send[d]) cudaFree(h.send[d]);
    if (h.recv[d]) cudaFree(h.recv[d]);
    CUDA_CHECK(cudaMalloc(&h.send[d], cap * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&h.recv[d], cap * sizeof(float)));
  }
  h.cap = cap;
}

void halo_destroy(Solver *s) {
  if (!s->comm) return;
  HaloState &h = hs(s);
  for (int d = 0; d < 4; d++) { cudaFree(h.send[d]); cudaFree(h.recv[d]); }
  for (size_t r = 0; r < h.peer_win.size(); r++)
    if (h.peer_win[r] && h.peer_ipc[r]) cudaIpcCloseMemHandle(h.peer_win[r]);
  if (s->hstream) {
    cudaStreamDestroy(s->hstream); cudaEventDestroy(s->ev_fork); cudaEventDestroy(s->ev_join); s->hstream = nullptr;
    for (int m = 0; m < 4; m++) { cudaStreamDestroy(s->rstream[m]); cudaEventDestroy(s->ev_ring[m]); }
  }
  if (h.win) cudaFree(h.win);
  if (h.d_count) cudaFree(h.d_count);
  if (h.packed) cudaEventDestroy(h.packed);
  if (h.consumed) cudaEventDestroy(h.consumed);
  if (h.comm && nccl().CommDestroy) nccl().CommDestroy(h.comm);
  delete &h;
  s->comm = nullptr;
}

static void launch_copy(Solver *s, bool unpack, HaloArgs &a) {
  long total = 0;
  for (int sd = 0; sd < 2; sd++)
    if (a.active[sd]) total = std::max(total, (long)(a.ie[sd] - a.is[sd] + 1) * (a.je[sd] - a.js[sd] + 1) * a.nk);
  if (!total) return;
  dim3 grid((unsigned)std::min(2048L, (total + 255) / 256), 1, (unsigned)(2 * a.n));
  s->kpre(unpack ? "k_halo_unpack" : "k_halo_pack");
  if (unpack) k_halo_copy<true><<<grid, 256, 0, s->stream>>>(s->g, a);
  else k_halo_copy<false><<<grid, 256, 0, s->stream>>>(s->g, a);
  s->kpost();
  s->launches++;
}

// one phase (Y: dirs 0,1; X: dirs 2,3) of one exchange
static void exchange_phase(Solver *s, int width, int hwidth, const std::vector<Field *> &flds, int d0) {
  HaloState &h = hs(s);
  HaloArgs snd{}, rcv{};
  snd.n = rcv.n = (int)flds.size();
  snd.nk = rcv.nk = s->g.nk;
  long unit = 0;
  for (size_t f = 0; f < flds.size(); f++) {
    snd.p[f] = rcv.p[f] = flds[f]->o;
    snd.is3d[f] = rcv.is3d[f] = flds[f]->kind == 3;
    snd.off[f] = rcv.off[f] = unit;
    unit += flds[f]->kind == 3 ? s->g.nk : 1;
  }
  Strip ss[2], rs[2];
  bool has[2];
  size_t nsend[2] = {0, 0}, nrecv[2] = {0, 0}, need = 0;
  for (int q = 0; q < 2; q++) {
    has[q] = halo_box(s->d, s->c, width, d0 + q, true, ss[q], hwidth);
    if (has[q]) {
      halo_box(s->d, s->c, width, d0 + q, false, rs[q], hwidth);
      nsend[q] = (size_t)unit * (ss[q].ie - ss[q].is + 1) * (ss[q].je - ss[q].js + 1);
      nrecv[q] = (size_t)unit * (rs[q].ie - rs[q].is + 1) * (rs[q].je - rs[q].js + 1);
      need = std::max(need, std::max(nsend[q], nrecv[q]));
    }
  }
  if (!has[0] && !has[1]) {
    if (h.local) { h.local->barrier(); if (!h.direct) h.local->barrier(); }   // keep the collective call pattern aligned
    return;
  }
  if (h.direct) {   // peer windows: pack into the neighbour's receive buffer, unpack mine (see k_halo_direct)
    if (need > h.win_cap) throw std::runtime_error("wrfgpu: halo strip larger than the peer window (halo_capacity)");
    LinkArgs put{}, get{};
    for (int q = 0; q < 2; q++) {
      snd.active[q] = rcv.active[q] = has[q];
      if (!has[q]) continue;
      const int dir = d0 + q;
      const unsigned n = ++h.seq[dir];
      char *pw = h.peer_win[ss[q].peer];
      if (!pw) throw std::runtime_error("wrfgpu: no peer window for rank " + std::to_string(ss[q].peer));
      snd.is[q] = ss[q].is; snd.ie[q] = ss[q].ie; snd.js[q] = ss[q].js; snd.je[q] = ss[q].je;
      snd.buf[q] = win_buf(pw, dir ^ 1, n & 1u, h.win_cap);
      rcv.is[q] = rs[q].is; rcv.ie[q] = rs[q].ie; rcv.js[q] = rs[q].js; rcv.je[q] = rs[q].je;
      rcv.buf[q] = win_buf(h.win, dir, n & 1u, h.win_cap);
      put.wait_flag[q] = n >= 3 ? win_ack(h.win, dir) : nullptr; put.wait_val[q] = n - 2;
      put.post_flag[q] = win_ready(pw, dir ^ 1); put.post_val[q] = n;
      put.count[q] = h.d_count + q;
      get.wait_flag[q] = win_ready(h.win, dir); get.wait_val[q] = n;
      get.post_flag[q] = win_ack(pw, dir ^ 1); get.post_val[q] = n;
      get.count[q] = h.d_count + 2 + q;
    }
    long total = 0;
    for (int q = 0; q < 2; q++)
      if (has[q]) total = std::max(total, std::max((long)(snd.ie[q] - snd.is[q] + 1) * (snd.je[q] - snd.js[q] + 1),
                                                    (long)(rcv.ie[q] - rcv.is[q] + 1) * (rcv.je[q] - rcv.js[q] + 1)) * snd.nk);
    // about two blocks per SM in all: every block ends with a system-scope fence and an atomic, and strips are small
    const long per_fs = std::max(1L, 296L / (2 * snd.n));
    const unsigned gx = (unsigned)std::min(per_fs, (total + HU * HALO_TPB - 1) / (HU * HALO_TPB));
    for (int q = 0; q < 2; q++) put.nblocks[q] = get.nblocks[q] = gx * (unsigned)snd.n;
    HaloPair pp{snd, put}, pg{rcv, get};
    if (!h.local) {   // processes: pack and unpack in one launch
      s->kpre("k_halo_exchange");
      k_halo_direct<2><<<dim3(gx, 2, (unsigned)(2 * snd.n)), HALO_TPB, 0, s->stream>>>(s->g, pp, pg);
      s->kpost();
      s->launches++;
    }
    for (int pass = 0; pass < 2 && h.local; pass++) {
      HaloPair &pr = pass ? pg : pp;
      LinkArgs &l = pr.l;
      dim3 grid(gx, 1, (unsigned)(2 * snd.n));
      if (l.wait_flag[0] || l.wait_flag[1]) {   // solvers sharing a GPU: wait in a kernel of one warp
        s->kpre("k_halo_wait");
        k_halo_wait<<<1, 32, 0, s->stream>>>(l);
        s->kpost();
        s->launches++;
        l.wait_flag[0] = l.wait_flag[1] = nullptr;
      }
      s->kpre(pass ? "k_halo_get" : "k_halo_put");
      if (pass) k_halo_direct<1><<<grid, HALO_TPB, 0, s->stream>>>(s->g, pp, pr);
      else k_halo_direct<0><<<grid, HALO_TPB, 0, s->stream>>>(s->g, pr, pg);
      s->kpost();
      s->launches++;
      // Solvers of ONE process share a CUDA context: an implicitly synchronising call of one host thread (cudaMalloc, a
      // first-launch module load) would wait for a spinning get kernel of another whose put has not been launched yet.
      // A host barrier between the two launches rules that out: when a get kernel starts, every put it can wait for is
      // already in a stream.  Processes (one context per GPU) need no such thing.
      if (pass == 0 && h.local) h.local->barrier();
    }
    h.exchanges++;
    h.bytes_sent += (long long)((nsend[0] + nsend[1]) * sizeof(float));
    const bool isy_d = d0 == 0;
    const bool close_d = isy_d ? (s->c.periodic_x && s->c.ntasks_x == 1) : (s->c.periodic_y && s->c.ntasks_y == 1);
    if (close_d)
      for (int q = 0; q < 2; q++) {
        if (!has[q]) continue;
        const int a0 = isy_d ? rs[q].js : rs[q].is, a1 = isy_d ? rs[q].je : rs[q].ie;
        const int lines = (a1 - a0 + 1) * s->g.nk;
        s->kpre("k_halo_close");
        k_halo_close<<<dim3((unsigned)((lines + 127) / 128), 1, (unsigned)rcv.n), 128, 0, s->stream>>>(s->g, rcv, isy_d ? 1 : 0, a0, a1);
        s->kpost();
        s->launches++;
      }
    return;
  }
  ensure_buffers(s, need);
  for (int q = 0; q < 2; q++) {
    snd.active[q] = rcv.active[q] = has[q];
    if (!has[q]) continue;
    snd.is[q] = ss[q].is; snd.ie[q] = ss[q].ie; snd.js[q] = ss[q].js; snd.je[q] = ss[q].je; snd.buf[q] = h.send[d0 + q];
    rcv.is[q] = rs[q].is; rcv.ie[q] = rs[q].ie; rcv.js[q] = rs[q].js; rcv.je[q] = rs[q].je; rcv.buf[q] = h.recv[d0 + q];
  }
  launch_copy(s, false, snd);
  if (h.local) {
    CUDA_CHECK(cudaEventRecord(h.packed, s->stream));
    h.local->barrier();
    for (int q = 0; q < 2; q++) {
      if (!has[q]) continue;
      Solver *peer = h.local->ranks[rs[q].peer];
      HaloState &ph = hs(peer);
      CUDA_CHECK(cudaStreamWaitEvent(s->stream, ph.packed, 0));
      // my recv from direction d is the peer's send towards the opposite direction; read it in place
      rcv.buf[q] = ph.send[(d0 + q) ^ 1];
    }
    launch_copy(s, true, rcv);
    CUDA_CHECK(cudaEventRecord(h.consumed, s->stream));
    h.local->barrier();
    for (int q = 0; q < 2; q++)
      if (has[q]) CUDA_CHECK(cudaStreamWaitEvent(s->stream, hs(h.local->ranks[ss[q].peer]).consumed, 0));
  } else {
    Nccl &n = nccl();
    if (!h.comm) throw std::runtime_error("wrfgpu: halo exchange on a multi-rank mesh before wrfgpu_comm_init");
    nccl_check(n.GroupStart(), "GroupStart");
    // sends in the order (low, high), receives in the order (high, low): when both neighbours are the
    // same rank (2 ranks, periodic) the n-th send to a peer then meets that peer's n-th receive
    for (int q = 0; q < 2; q++)
      if (has[q]) nccl_check(n.Send(h.send[d0 + q], nsend[q], 7 /*ncclFloat*/, ss[q].peer, h.comm, s->stream), "Send");
    for (int q = 1; q >= 0; q--)
      if (has[q]) nccl_check(n.Recv(h.recv[d0 + q], nrecv[q], 7, rs[q].peer, h.comm, s->stream), "Recv");
    nccl_check(n.GroupEnd(), "GroupEnd");
    launch_copy(s, true, rcv);
  }
  h.exchanges++;
  h.bytes_sent += (long long)((nsend[0] + nsend[1]) * sizeof(float));
  // close the other dimension on what was just received if it is periodic and lives on this rank
  const bool isy = d0 == 0;
  const bool close_other = isy ? (s->c.periodic_x && s->c.ntasks_x == 1) : (s->c.periodic_y && s->c.ntasks_y == 1);
  if (close_other)
    for (int q = 0; q < 2; q++) {
      if (!has[q]) continue;
      const int a0 = isy ? rs[q].js : rs[q].is, a1 = isy ? rs[q].je : rs[q].ie;
      const int lines = (a1 - a0 + 1) * s->g.nk;
      s->kpre("k_halo_close");
      k_halo_close<<<dim3((unsigned)((lines + 127) / 128), 1, (unsigned)rcv.n), 128, 0, s->stream>>>(s->g, rcv, isy ? 1 : 0, a0, a1);
      s->kpost();
      s->launches++;
    }
}

// floats of the largest strip any exchange of this solver packs (all fields of a batch, the wider of the two widths,
// the longer side of the patch plus corners)
static size_t halo_capacity(Solver *s) {
  size_t cap = 0;
  // the same number on every rank (the neighbours index each other's windows with it): the largest patch of the mesh
  const int ntx = std::max(1, s->c.ntasks_x), nty = std::max(1, s->c.ntasks_y);
  const size_t side = (size_t)std::max((s->d.ide - s->d.ids + ntx) / ntx, (s->d.jde - s->d.jds + nty) / nty) + 2;
  for (const HaloDef &hd : kHalos) {
    int width = 0;
    std::vector<std::string> f3, f2;
    halo_fields(s->c, hd.name, width, f3, f2);
    const size_t w = (size_t)std::abs(width);
    size_t unit = 0, nf = 0;
    for (auto &n : f3) if (s->fields.count(n)) { unit += s->g.nk; nf++; }
    for (auto &n : f2) if (s->fields.count(n)) { unit += 1; nf++; }
    (void)nf;
    cap = std::max(cap, unit * (w + 1) * (side + 2 * w + 2));
  }
  return cap + 64;
}

static void direct_alloc(Solver *s) {
  HaloState &h = hs(s);
  h.win_cap = (halo_capacity(s) + 31) / 32 * 32;
  const size_t bytes = WIN_HDR + 8 * h.win_cap * sizeof(float);
  CUDA_CHECK(cudaMalloc(&h.win, bytes));
  CUDA_CHECK(cudaMemset(h.win, 0, WIN_HDR));
  CUDA_CHECK(cudaMalloc(&h.d_count, 4 * sizeof(unsigned)));
  CUDA_CHECK(cudaMemset(h.d_count, 0, 4 * sizeof(unsigned)));
  h.peer_win.assign(s->nranks, nullptr);
  h.peer_ipc.assign(s->nranks, false);
  CUDA_CHECK(cudaDeviceSynchronize());
}

// which ranks this rank ever exchanges with
static std::vector<int> halo_peers(Solver *s) {
  std::vector<int> out;
  for (int w : {4, -3})          // with and without the periodic closure
    for (int dir = 0; dir < 4; dir++) {
      Strip st;
      if (halo_box(s->d, s->c, w, dir, true, st) && std::find(out.begin(), out.end(), st.peer) == out.end()) out.push_back(st.peer);
    }
  return out;
}

// ranks are processes: exchange the windows' IPC handles through NCCL and map the neighbours' windows
static void direct_setup_ipc(Solver *s) {
  HaloState &h = hs(s);
  Nccl &n = nccl();
  if (!n.AllGather) throw std::runtime_error("wrfgpu: ncclAllGather not found");
  direct_alloc(s);
  struct Rec { cudaIpcMemHandle_t handle; unsigned long long cap; char pad[56]; };
  static_assert(sizeof(Rec) == 128, "Rec");
  Rec mine{};
  CUDA_CHECK(cudaIpcGetMemHandle(&mine.handle, h.win));
  mine.cap = h.win_cap;
  Rec *d_all = nullptr;
  CUDA_CHECK(cudaMalloc(&d_all, sizeof(Rec) * s->nranks));
  CUDA_CHECK(cudaMemcpy(d_all + s->rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  nccl_check(n.AllGather(d_all + s->rank, d_all, sizeof(Rec), 0 /*ncclChar*/, h.comm, s->stream), "AllGather");
  CUDA_CHECK(cudaStreamSynchronize(s->stream));
  std::vector<Rec> all(s->nranks);
  CUDA_CHECK(cudaMemcpy(all.data(), d_all, sizeof(Rec) * s->nranks, cudaMemcpyDeviceToHost));
  cudaFree(d_all);
  for (int r : halo_peers(s)) {
    if (r == s->rank) { h.peer_win[r] = h.win; continue; }
    if (all[r].cap != h.win_cap) throw std::runtime_error("wrfgpu: peer windows of different sizes");
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw std::runtime_error(std::string("wrfgpu: cudaIpcOpenMemHandle failed (") + cudaGetErrorString(e) + ")");
    }
    h.peer_win[r] = (char *)p;
    h.peer_ipc[r] = true;
  }
  h.direct = true;
}

// wrf_dm_maxval's role for grid%max_vert_cfl / max_horiz_cfl (dyn_em/adapt_timestep_em.F): ncclAllReduce(max) between
// processes; solvers of one process compare their host copies between two barriers.
void halo_max2(Solver *s, float *d2, float out[2]) {
  if (s->nranks > 1) {
    HaloState &h = hs(s);
    if (h.local) {
      CUDA_CHECK(cudaMemcpyAsync(h.max2, d2, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
      CUDA_CHECK(cudaStreamSynchronize(s->stream));
      h.local->barrier();
      out[0] = out[1] = 0.f;
      for (Solver *r : h.local->ranks) { out[0] = std::max(out[0], hs(r).max2[0]); out[1] = std::max(out[1], hs(r).max2[1]); }
      h.local->barrier();
      CUDA_CHECK(cudaMemcpyAsync(d2, out, 2 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
      CUDA_CHECK(cudaStreamSynchronize(s->stream));
      return;
    }
    Nccl &n = nccl();
    if (!h.comm || !n.AllReduce) throw std::runtime_error("wrfgpu: CFL maxima on a multi-rank mesh before wrfgpu_comm_init");
    nccl_check(n.AllReduce(d2, d2, 2, 7 /*ncclFloat*/, 2 /*ncclMax*/, h.comm, s->stream), "AllReduce");
  }
  CUDA_CHECK(cudaMemcpyAsync(out, d2, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  CUDA_CHECK(cudaStreamSynchronize(s->stream));
}

// the strips of `full` within w of a patch side across which this rank has a neighbour (mesh or periodic wrap), and the rest
int Solver::ring_split(const Box &full, int w, Box ring[4], Box &inner) {
  Strip st;
  bool link[4];
  for (int dir = 0; dir < 4; dir++) link[dir] = halo_box(d, c, 4, dir, true, st);
  inner = full;
  int n = 0;
  if (link[0]) { ring[n++] = Box{full.i0, full.i1, full.k0, full.k1, full.j0, std::min(full.j1, full.j0 + w - 1)}; inner.j0 = full.j0 + w; }
  if (link[1] && inner.j0 <= full.j1) {
    const int j0 = std::max(inner.j0, full.j1 - w + 1);
    ring[n++] = Box{full.i0, full.i1, full.k0, full.k1, j0, full.j1}; inner.j1 = j0 - 1;
  }
  if (inner.j0 > inner.j1) return n;
  if (link[2]) { ring[n++] = Box{full.i0, std::min(full.i1, full.i0 + w - 1), full.k0, full.k1, inner.j0, inner.j1}; inner.i0 = full.i0 + w; }
  if (link[3] && inner.i0 <= full.i1) {
    const int i0 = std::max(inner.i0, full.i1 - w + 1);
    ring[n++] = Box{i0, full.i1, full.k0, full.k1, inner.j0, inner.j1}; inner.i1 = i0 - 1;
  }
  return n;
}

void Solver::overlap_streams() {
  if (hstream) return;
  int lo = 0, hi = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // hi = the greatest priority (numerically lowest)
  CUDA_CHECK(cudaStreamCreateWithPriority(&hstream, cudaStreamNonBlocking, hi));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  for (int m = 0; m < 4; m++) {
    CUDA_CHECK(cudaStreamCreateWithPriority(&rstream[m], cudaStreamNonBlocking, hi));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_ring[m], cudaEventDisableTiming));
  }
}

void Solver::halo(const std::string &name) {
  if (nranks == 1) return;   // a single task: every HALO_*.inc is a no-op, periodic closure is set_physical_bc's job
  int width = 0;
  std::vector<std::string> f3, f2;
  if (!halo_fields(c, name, width, f3, f2)) throw std::runtime_error("wrfgpu: unknown halo '" + name + "'");
  std::vector<Field *> all;
  for (auto &n : f3) if (fields.count(n)) all.push_back(&F(n));
  for (auto &n : f2) if (fields.count(n)) all.push_back(&F(n));
  for (size_t pos = 0; pos < all.size(); pos += HALO_MAX) {
    std::vector<Field *> batch(all.begin() + pos, all.begin() + std::min(all.size(), pos + HALO_MAX));
    const int hw = regular_width(name, width);
    exchange_phase(this, width, hw, batch, 0);   // Y
    exchange_phase(this, width, hw, batch, 2);   // X (j range includes the rows just received)
  }
}

}  // namespace wg

// ---------------------------------------------------------------------------------------------
using wg::Solver;

extern "C" {

int wrfgpu_halo_spec(const wrfgpu_config *c, const char *halo_name, int *width, char *fields, size_t cap) {
  std::vector<std::string> f3, f2;
  int w = 0;
  if (!wg::halo_fields(*c, halo_name, w, f3, f2)) return 1;
  std::string out;
  for (auto &n : f3) out += (out.empty() ? "" : ",") + n;
  for (auto &n : f2) out += (out.empty() ? "" : ",") + n;
  if (width) *width = w;
  if (fields && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(fields, out.data(), n); fields[n] = 0; }
  return 0;
}

int wrfgpu_halo_box(const wrfgpu_dims *d, const wrfgpu_config *c, int width, int dir, int send, int *peer, int *is, int *ie,
                    int *js, int *je) {
  return wrfgpu_halo_box2(d, c, width, 0, dir, send, peer, is, ie, js, je);
}

int wrfgpu_halo_box2(const wrfgpu_dims *d, const wrfgpu_config *c, int width, int regular_width, int dir, int send, int *peer,
                     int *is, int *ie, int *js, int *je) {
  wg::Strip s;
  if (!wg::halo_box(*d, *c, width, dir, send != 0, s, regular_width)) return 0;
  *peer = s.peer; *is = s.is; *ie = s.ie; *js = s.js; *je = s.je;
  return 1;
}

int wrfgpu_halo_regular_width(const char *halo_name, int width) { return wg::regular_width(halo_name, width); }

int wrfgpu_comm_unique_id(void *id128) {
  wg::Nccl &n = wg::nccl();
  if (!n.GetUniqueId) { g_wrfgpu_err = "wrfgpu: libnccl.so.2 could not be loaded (set WRFGPU_NCCL_LIB)"; return 1; }
  return n.GetUniqueId(id128) == 0 ? 0 : 1;
}

int wrfgpu_comm_init(wrfgpu_solver *s, const void *id128, int rank, int nranks) {
  try {
    Solver &S = s->s;
    if (nranks != S.nranks || rank != S.rank) throw std::runtime_error("wrfgpu_comm_init: rank/nranks do not match the solver's mesh");
    wg::Nccl &n = wg::nccl();
    if (!n.CommInitRank) throw std::runtime_error("wrfgpu: libnccl.so.2 could not be loaded (set WRFGPU_NCCL_LIB)");
    CUDA_CHECK(cudaSetDevice(S.device));
    wg::HaloState &h = wg::hs(&S);
    wg::Id128 id;
    memcpy(id.b, id128, 128);
    wg::nccl_check(n.CommInitRank(&h.comm, nranks, id, rank), "CommInitRank");
    // halos travel through peer windows (CUDA IPC over NVLink); WRFGPU_HALO=nccl keeps ncclSend/ncclRecv (A/B runs)
    const char *mode = getenv("WRFGPU_HALO");
    if (!(mode && !strcmp(mode, "nccl"))) {
      try { wg::direct_setup_ipc(&S); }
      catch (const std::exception &e) {
        if (mode && !strcmp(mode, "direct")) throw;
        fprintf(stderr, "wrfgpu: rank %d: peer windows unavailable (%s); halos go through ncclSend/ncclRecv\n", rank, e.what());
        wg::hs(&S).direct = false;
      }
    }
    return 0;
  } catch (const std::exception &e) { g_wrfgpu_err = e.what(); return 1; }
}

int wrfgpu_comm_init_local(wrfgpu_solver *const *all, int n) {
  try {
    auto *grp = new wg::LocalGroup();
    bool distinct = n > 1;    // every rank on its own GPU: the peer-window transport (their kernels run concurrently for sure)
    for (int r = 0; r < n; r++) {
      Solver &S = all[r]->s;
      if (S.nranks != n || S.rank != r) throw std::runtime_error("wrfgpu_comm_init_local: solvers must be given in rank order");
      grp->ranks.push_back(&S);
      for (int q = 0; q < r; q++) distinct = distinct && all[q]->s.device != S.device;
    }
    const char *mode = getenv("WRFGPU_HALO_DIRECT");   // 1: also between solvers that share a GPU (tests; needs concurrent kernels)
    const bool direct = (mode && mode[0] == '1') || (distinct && !(mode && mode[0] == '0'));
    for (int r = 0; r < n; r++) {
      Solver &S = all[r]->s;
      CUDA_CHECK(cudaSetDevice(S.device));
      wg::HaloState &h = wg::hs(&S);
      h.local = grp;
      h.owns_local = r == 0;
      CUDA_CHECK(cudaEventCreateWithFlags(&h.packed, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&h.consumed, cudaEventDisableTiming));
      for (int q = 0; q < n; q++) {
        const int pd = all[q]->s.device;
        if (pd != S.device) {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, S.device, pd);
          if (!can) throw std::runtime_error("wrfgpu_comm_init_local: no peer access between the solvers' devices");
          cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
          cudaGetLastError();
        }
      }
    }
    if (direct) {
      for (int r = 0; r < n; r++) { CUDA_CHECK(cudaSetDevice(all[r]->s.device)); wg::direct_alloc(&all[r]->s); }
      for (int r = 0; r < n; r++) {
        wg::HaloState &h = wg::hs(&all[r]->s);
        for (int q = 0; q < n; q++) {
          if (wg::hs(&all[q]->s).win_cap != h.win_cap) throw std::runtime_error("wrfgpu_comm_init_local: peer windows of different sizes");
          h.peer_win[q] = wg::hs(&all[q]->s).win;
        }
        h.direct = true;
      }
    }
    return 0;
  } catch (const std::exception &e) { g_wrfgpu_err = e.what(); return 1; }
}

long long wrfgpu_halo_bytes(const wrfgpu_solver *s) {
  return s->s.comm ? ((wg::HaloState *)s->s.comm)->bytes_sent : 0;
}

}  // extern "C"
