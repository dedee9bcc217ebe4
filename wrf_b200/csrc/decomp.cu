// decomp.cu -- host-side restatement of the RSL_LITE decomposition rules the GPU build
// must reproduce exactly (a patch boundary in the wrong place would break the
// "N-GPU == 1-GPU bit for bit" property): mpaspect/compute_mesh
// (external/RSL_LITE/module_dm.F:98-175), task_for_point (task_for_point.c:6-160),
// patch + memory dims (module_dm.F:924-1003, :1174-1201) and the single-hop strip
// bounds of rsl_comm_iter (:2761) / RSL_LITE_PACK (c_code.c:405-700).  No device code.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace wg {

// task_for_point.c: owner of index i along one dimension of n = ide-ids+1 points on np tasks;
// the remainder is distributed to both ends of the dimension
static int task_for_point_1d(int i, int ds, int de, int np) {
  i -= 1; ds -= 1; de -= 1;
  const int dim = de - ds + 1;
  if (np > dim) np = dim;
  if (np < 1) np = 1;
  i = i >= ds ? i : ds;
  i = i <= de ? i : de;
  const int rem = dim % np;
  const int a = (rem / 2) * ((dim / np) + 1);
  const int b = a + (np - rem) * (dim / np);
  int P;
  if (i - ds < a) P = (i - ds) / ((dim / np) + 1);
  else if (i - ds < b) P = (a / ((dim / np) + 1)) + (i - a - ds) / ((b - a) / (np - rem));
  else P = (a / ((dim / np) + 1)) + (b - a - ds) / ((b - a) / (np - rem)) + (i - b - ds) / ((dim / np) + 1);
  return P;
}

}  // namespace wg

extern "C" {

int wrfgpu_compute_mesh(int ntasks, int nproc_x, int nproc_y, int *ntasks_x, int *ntasks_y) {
  if (ntasks < 1) return 1;
  int nx = nproc_x, ny = nproc_y;
  if (nx > 0 || ny > 0) {
    if (nx > 0 && ny == -1) ny = ntasks / nx;
    else if (nx == -1 && ny > 0) nx = ntasks / ny;
    if (nx * ny != ntasks) return 1;
  } else {   // MPASPECT(P, MINM, MINN, 1, 1)
    int mini = 2 * ntasks;
    nx = 1; ny = ntasks;
    for (int m = 1; m <= ntasks; m++)
      if (ntasks % m == 0) {
        const int n = ntasks / m;
        if (std::abs(m - n) < mini) { mini = std::abs(m - n); nx = m; ny = n; }
      }
  }
  *ntasks_x = nx; *ntasks_y = ny;
  return 0;
}

int wrfgpu_patch_dims(int ids, int ide, int jds, int jde, int kds, int kde, int ntasks_x, int ntasks_y, int mytask,
                      wrfgpu_dims *out) {
  if (ntasks_x < 1 || ntasks_y < 1 || mytask < 0 || mytask >= ntasks_x * ntasks_y) return 1;
  const int mx = mytask % ntasks_x, my = mytask / ntasks_x;
  const int max_halo_width = 6, bdx = 4, bdy = 4;   // module_dm.F:19, share/module_bc.F:33
  int ips = -1, ipe = -1, jps = -1, jpe = -1;
  for (int i = ids; i <= ide; i++)
    if (wg::task_for_point_1d(i, ids, ide, ntasks_x) == mx) { ipe = i; if (ips == -1) ips = i; }
  for (int j = jds; j <= jde; j++)
    if (wg::task_for_point_1d(j, jds, jde, ntasks_y) == my) { jpe = j; if (jps == -1) jps = j; }
  if (ips == -1 || jps == -1) return 2;   // more tasks than points
  out->ids = ids; out->ide = ide; out->jds = jds; out->jde = jde; out->kds = kds; out->kde = kde;
  out->ips = ips; out->ipe = ipe; out->jps = jps; out->jpe = jpe; out->kps = kds; out->kpe = kde;
  out->ims = std::max(ips - max_halo_width, ids - bdx) - 1;
  out->ime = std::min(ipe + max_halo_width, ide + bdx) + 1;
  out->jms = std::max(jps - max_halo_width, jds - bdy) - 1;
  out->jme = std::min(jpe + max_halo_width, jde + bdy) + 1;
  out->kms = kds; out->kme = kde;
  return 0;
}

}  // extern "C"
