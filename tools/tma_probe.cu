// tma_probe.cu -- standalone check of TMA tiled loads (debug tool for wrf_b200/csrc/tend_tile.cu)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK, bool GLOBALMAP>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int c0, int c1, int c2, float *out, int *spins, int nel, int bytes, int fence_kind) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap *map = GLOBALMAP ? gmap : &pmap;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
    if (fence_kind == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else if (fence_kind == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    else { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(sm)),
                   "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                   "l"(map), "r"(c0), "r"(c2), "r"(s32(&bar)) : "memory");
  }
  int n = 0; uint32_t ok = 0;
  for (; n < (1 << 20) && !ok; n++)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
  if (threadIdx.x == 0) spins[0] = ok ? n : -n;
  __syncthreads();
  for (int e = threadIdx.x; e < nel; e += blockDim.x) out[e] = sm[e];
}
typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int fk = argc > 2 ? atoi(argv[2]) : 0;
  const int np = 2016, nk = 80, nj = 2010;
  std::vector<float> h((size_t)np * nk * nj);
  for (size_t x = 0; x < h.size(); x++) h[x] = (float)(x % 1000003);
  float *d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  float *out; int *spins; CK(cudaMalloc(&out, 8192 * 4)); CK(cudaMalloc(&spins, 16));
  CUtensorMap *gmap; CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
  struct V { int rank, bw, bk, bh; bool gm; } vs[] = {{2, 32, 1, 8, false}, {2, 40, 1, 14, false}, {3, 32, 1, 8, false}, {3, 32, 2, 8, false}, {3, 40, 1, 14, false}, {3, 40, 1, 14, true}, {3, 64, 1, 8, false}};
  int vi = -1;
  for (auto &v : vs) {
    vi++;
    if (only >= 0 && vi != only) continue;
    CUtensorMap m;
    CUresult r;
    if (v.rank == 3) {
      cuuint64_t dims[3] = {(cuuint64_t)np, (cuuint64_t)nk, (cuuint64_t)nj};
      cuuint64_t str[2] = {(cuuint64_t)np * 4, (cuuint64_t)np * nk * 4};
      cuuint32_t box[3] = {(cuuint32_t)v.bw, (cuuint32_t)v.bk, (cuuint32_t)v.bh}, es[3] = {1, 1, 1};
      r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)np, (cuuint64_t)nk * nj};
      cuuint64_t str[1] = {(cuuint64_t)np * 4};
      cuuint32_t box[2] = {(cuuint32_t)v.bw, (cuuint32_t)v.bh}, es[2] = {1, 1};
      r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    CK(cudaMemcpy(gmap, &m, sizeof(m), cudaMemcpyHostToDevice));
    const int nel = v.bw * v.bk * v.bh;
    const size_t smem = 8192 * 4;
    if (v.rank == 3 && v.gm) { CK(cudaFuncSetAttribute(probe<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<3, true><<<1, 128, smem>>>(m, gmap, 2, 1, 2, out, spins, nel, nel * 4, fk); }
    else if (v.rank == 3) { CK(cudaFuncSetAttribute(probe<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<3, false><<<1, 128, smem>>>(m, gmap, 2, 1, 2, out, spins, nel, nel * 4, fk); }
    else { CK(cudaFuncSetAttribute(probe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<2, false><<<1, 128, smem>>>(m, gmap, 2, 1, 2, out, spins, nel, nel * 4, fk); }
    cudaError_t e = cudaDeviceSynchronize();
    int hs = 0; std::vector<float> ho(nel);
    if (e == cudaSuccess) { cudaMemcpy(&hs, spins, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), out, nel * 4, cudaMemcpyDeviceToHost); }
    printf("fence %d rank %d box %dx%dx%d map in %s: encode %d, %s, spins %d, first %g %g\n", fk, v.rank, v.bw, v.bk, v.bh, v.gm ? "global" : "param", (int)r, cudaGetErrorString(e), hs, ho[0], ho[1]);
    if (e != cudaSuccess) { cudaDeviceReset(); return 2; }
  }
  return 0;
}
