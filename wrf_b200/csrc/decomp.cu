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

// returns 1 and the strip bounds when this rank exchanges in direction `dir` (0=-y 1=+y 2=-x 3=+x), else 0.
// `periodic` neighbours (mesh wrap, external/RSL_LITE/period.c) are included when the BC flag is set
// and the dimension is split over more than one task.
int wrfgpu_halo_strip(const wrfgpu_dims *d, const wrfgpu_config *c, int width, int xstag, int ystag, int dir, int send,
                      int *is, int *ie, int *js, int *je) {
  (void)xstag; (void)ystag;   // "not used" in RSL_LITE_PACK as well (c_code.c:408)
  const int ntx = c->ntasks_x, nty = c->ntasks_y;
  const int mx = c->mytask % ntx, my = c->mytask / ntx;
  const int w = width;
  bool have;
  switch (dir) {
    case 0: have = my > 0 || (c->periodic_y && nty > 1); break;
    case 1: have = my < nty - 1 || (c->periodic_y && nty > 1); break;
    case 2: have = mx > 0 || (c->periodic_x && ntx > 1); break;
    default: have = mx < ntx - 1 || (c->periodic_x && ntx > 1); break;
  }
  if (!have) return 0;
  if (dir < 2) {   // y exchange: i range clipped to the domain (IMAX/IMIN, c_code.c:460-461)
    *is = std::max(d->ips - w, d->ids); *ie = std::min(d->ipe + w, d->ide);
    if (dir == 0) { if (send) { *js = d->jps; *je = d->jps + w - 1; } else { *js = d->jps - w; *je = d->jps - 1; } }
    else { if (send) { *js = d->jpe - w + 1; *je = d->jpe; } else { *js = d->jpe + 1; *je = d->jpe + w; } }
  } else {         // x exchange: j range includes the y halo just received, so corners propagate
    *js = std::max(d->jps - w, d->jds); *je = std::min(d->jpe + w, d->jde);
    if (dir == 2) { if (send) { *is = d->ips; *ie = d->ips + w - 1; } else { *is = d->ips - w; *ie = d->ips - 1; } }
    else { if (send) { *is = d->ipe - w + 1; *ie = d->ipe; } else { *is = d->ipe + 1; *ie = d->ipe + w; } }
  }
  return 1;
}

}  // extern "C"
