// bc.cu -- physical boundary fills, share/module_bc.F set_physical_bc2d (:202-648)
// and set_physical_bc3d (:651-1113): periodic copy (only if the whole dimension is
// on this patch), symmetric mirror, open zero-gradient copy.  bdyzone = 4 (:33).
//
// The reference calls these ~40 times per RK substep, one field at a time; each
// call touches only a few ghost rows, i.e. it is pure launch overhead on a GPU.
// Here one launch handles a BATCH of fields: the x pass (ghost columns) and the y
// pass (ghost rows over the full i extent, so corners are taken from x-filled
// data exactly as in the reference's two-stage fill) are two kernels per batch.
#include "common.cuh"

namespace wg {

#define BC_MAX 24
struct BcBatch {
  float *p[BC_MAX];   // origin-shifted
  char var[BC_MAX];
  char is3d[BC_MAX];
  int n;
};

#define BDY 4

struct BcInfo {
  int istag, jstag, k_end;
  bool open_bc_copy;
  char variable;
};

__device__ __forceinline__ BcInfo bc_info(const GeoBase &g, char v_in, bool is3d) {
  BcInfo q;
  char v = v_in;
  if (v >= 'A' && v <= 'Z') v = (char)(v - 'A' + 'a');
  q.variable = v;
  q.istag = -1; q.jstag = -1;
  const int kte = g.kde;
  q.k_end = max(1, min(g.kde - 1, kte));
  if (v == 'u' || v == 'x') q.istag = 0;
  if (v == 'v' || v == 'y') q.jstag = 0;
  if (v == 'd') { q.istag = 0; q.jstag = 0; }
  if (v == 'e') { q.istag = 0; q.k_end = min(g.kde, kte); }
  if (v == 'f') { q.jstag = 0; q.k_end = min(g.kde, kte); }
  if (v == 'w') q.k_end = min(g.kde, kte);
  if (is3d)
    q.open_bc_copy = (v == 'u' || v == 'v' || v == 'w' || v == 't' || v == 'd' || v == 'e' || v == 'x' || v == 'y' ||
                      v == 'f' || v == 'r' || v == 'p');
  else
    q.open_bc_copy = (v == 'u' || v == 'v' || v == 'w' || v == 't' || v == 'x' || v == 'y' || v == 'r' || v == 'p');
  if (!is3d) q.k_end = 1;
  return q;
}

// x pass: one thread per (j, k, field); loops over the few ghost columns
__global__ void k_bc_x(const GeoBase g, BcBatch bb) {
  const int j = g.jms + blockIdx.x * blockDim.x + threadIdx.x;
  const int k = 1 + blockIdx.y;
  const int f = blockIdx.z;
  if (j > g.jme) return;
  const bool is3d = bb.is3d[f];
  const BcInfo q = bc_info(g, bb.var[f], is3d);
  if (k > q.k_end) return;
  float *dat = bb.p[f];
  const long sj = is3d ? g.sj : (long)g.ni;
  const long kk = is3d ? (long)k * g.sk : 0;
#define D(i) dat[(long)(i) + kk + (long)j * sj]
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  const int j0 = max(jds, jts - 1), j1 = min(jte + 1, jde + q.jstag);
  const char v = q.variable;
  if (g.periodic_x) {
    if (ids == g.ips && ide == g.ipe && j >= j0 && j <= j1) {
      if (its == ids)
        for (int i = 0; i >= -(BDY - 1); i--) D(ids + i - 1) = D(ide + i - 1);
      if (ite == ide)
        for (int i = -q.istag; i <= BDY; i++) D(ide + i + q.istag) = D(ids + i + q.istag);
    }
  } else {
    if (g.sym_xs && its == ids && j >= j0 && j <= j1) {
      if (is3d) {
        if (q.istag == -1) { for (int i = 1; i <= BDY; i++) D(ids - i) = D(ids + i - 1); }
        else if (v == 'u') { for (int i = 1; i <= BDY; i++) D(ids - i) = -D(ids + i); }
        else { for (int i = 1; i <= BDY; i++) D(ids - i) = D(ids + i); }
      } else {
        if (v != 'u' && v != 'x') { for (int i = 1; i <= BDY; i++) D(ids - i) = D(ids + i - 1); }
        else if (v == 'u') { for (int i = 0; i <= BDY - 1; i++) D(ids - i) = -D(ids + i); }
        else { for (int i = 0; i <= BDY - 1; i++) D(ids - i) = D(ids + i); }
      }
    }
    if (g.sym_xe && ite == ide && j >= j0 && j <= j1) {
      if (is3d) {
        if (q.istag == -1) { for (int i = 1; i <= BDY; i++) D(ide + i - 1) = D(ide - i); }
        else if (v == 'u') { for (int i = 1; i <= BDY; i++) D(ide + i) = -D(ide - i); }
        else { for (int i = 1; i <= BDY; i++) D(ide + i) = D(ide - i); }
      } else {
        if (v != 'u' && v != 'x') { for (int i = 1; i <= BDY; i++) D(ide + i - 1) = D(ide - i); }
        else if (v == 'u') { for (int i = 0; i <= BDY - 1; i++) D(ide + i) = -D(ide - i); }
        else { for (int i = 0; i <= BDY - 1; i++) D(ide + i) = D(ide - i); }
      }
    }
    if (g.oxs() && its == ids && q.open_bc_copy) {
      const int a0 = is3d ? jts - BDY : j0, a1 = is3d ? min(jte, jde + q.jstag) + BDY : j1;
      if (j >= a0 && j <= a1) {
        const float s = D(ids);
        D(ids - 1) = s; D(ids - 2) = s; D(ids - 3) = s;
      }
    }
    if (g.oxe() && ite == ide && q.open_bc_copy) {
      if (v != 'u' && v != 'x') {
        const int a0 = is3d ? jts - BDY : j0, a1 = is3d ? min(jte, jde + q.jstag) + BDY : j1;
        if (j >= a0 && j <= a1) {
          const float s = D(ide - 1);
          D(ide) = s; D(ide + 1) = s; D(ide + 2) = s;
        }
      } else {
        const int a0 = is3d ? j0 - BDY : j0, a1 = is3d ? j1 + BDY : j1;
        if (j >= a0 && j <= a1) {
          const float s = D(ide);
          D(ide + 1) = s; D(ide + 2) = s; D(ide + 3) = s;
        }
      }
    }
  }
#undef D
}

// y pass: one thread per (i, k, field); loops over the few ghost rows
__global__ void k_bc_y(const GeoBase g, BcBatch bb) {
  const int i = g.ims + blockIdx.x * blockDim.x + threadIdx.x;
  const int k = 1 + blockIdx.y;
  const int f = blockIdx.z;
  if (i > g.ime) return;
  const bool is3d = bb.is3d[f];
  const BcInfo q = bc_info(g, bb.var[f], is3d);
  if (k > q.k_end) return;
  float *dat = bb.p[f];
  const long sj = is3d ? g.sj : (long)g.ni;
  const long kk = is3d ? (long)k * g.sk : 0;
#define D(j) dat[(long)i + kk + (long)(j) * sj]
  const int ids = g.ids, ide = g.ide, jds = g.jds, jde = g.jde;
  const int its = g.its, ite = g.ite, jts = g.jts, jte = g.jte;
  int i_start = max(ids, its - 1), i_end = min(ite + 1, ide + q.istag);
  if (its == ids) i_start = g.ims;
  if (ite == ide) i_end = g.ime;
  if (i < i_start || i > i_end) return;
  const char v = q.variable;
  if (g.periodic_y) {
    if (jds == g.jps && jde == g.jpe) {
      if (jts == jds)
        for (int j = 0; j >= -(BDY - 1); j--) D(jds + j - 1) = D(jde + j - 1);
      if (jte == jde)
        for (int j = -q.jstag; j <= BDY; j++) D(jde + j + q.jstag) = D(jds + j + q.jstag);
    }
  } else {
    if (g.sym_ys && jts == jds) {
      const bool unst = is3d ? (q.jstag == -1) : (v != 'v' && v != 'y');
      if (unst) { for (int j = 1; j <= BDY; j++) D(jds - j) = D(jds + j - 1); }
      else if (v == 'v') { for (int j = 1; j <= BDY; j++) D(jds - j) = -D(jds + j); }
      else { for (int j = 1; j <= BDY; j++) D(jds - j) = D(jds + j); }
    }
    if (g.sym_ye && jte == jde) {
      const bool unst = is3d ? (q.jstag == -1) : (v != 'v' && v != 'y');
      if (unst) { for (int j = 1; j <= BDY; j++) D(jde + j - 1) = D(jde - j); }
      else if (v == 'v') { for (int j = 1; j <= BDY; j++) D(jde + j) = -D(jde - j); }
      else { for (int j = 1; j <= BDY; j++) D(jde + j) = D(jde - j); }
    }
    if (g.oys() && jts == jds && q.open_bc_copy) {
      const float s = D(jds);
      D(jds - 1) = s; D(jds - 2) = s; D(jds - 3) = s;
    }
    if (g.oye() && jte == jde && q.open_bc_copy) {
      if (v != 'v' && v != 'y') {
        const float s = D(jde - 1);
        D(jde) = s; D(jde + 1) = s; D(jde + 2) = s;
      } else {
        const float s = D(jde);
        D(jde + 1) = s; D(jde + 2) = s; D(jde + 3) = s;
      }
    }
  }
#undef D
}

void Solver::bc(const std::vector<BcItem> &items) {
  size_t pos = 0;
  while (pos < items.size()) {
    BcBatch bb;
    bb.n = 0;
    bool any3 = false;
    while (pos < items.size() && bb.n < BC_MAX) {
      Field &f = F(items[pos].name);
      if (f.kind == 1) throw std::runtime_error("wrfgpu: bc on a 1-D field");
      bb.p[bb.n] = f.o;
      bb.var[bb.n] = items[pos].var;
      bb.is3d[bb.n] = f.kind == 3;
      any3 |= f.kind == 3;
      bb.n++; pos++;
    }
    const int nk = any3 ? g.kde : 1;
    dim3 gx((unsigned)((g.nj + 63) / 64), (unsigned)nk, (unsigned)bb.n);
    kpre("k_bc_x");
    k_bc_x<<<gx, 64, 0, stream>>>(g, bb);
    kpost();
    dim3 gy((unsigned)((g.ni + 127) / 128), (unsigned)nk, (unsigned)bb.n);
    kpre("k_bc_y");
    k_bc_y<<<gy, 128, 0, stream>>>(g, bb);
    kpost();
    launches += 2;
  }
}

}  // namespace wg
