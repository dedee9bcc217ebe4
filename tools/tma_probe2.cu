// the CUDA programming guide's TMA example with one knob at a time turned towards tend_tile.cu's usage (debug tool)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// style 0: libcu++; 1: raw PTX barrier + copy, static smem; 2: raw PTX, dynamic smem
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float *out, int style) {
  __shared__ alignas(128) float smem_buffer[8][32];
  extern __shared__ __align__(128) float dyn[];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  __shared__ __align__(8) uint64_t rbar;
  float *dst = style == 2 ? dyn : &smem_buffer[0][0];
  if (style == 0) {
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
      cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
      token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
  } else {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&rbar)), "r"(1) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&rbar)), "r"(1024) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)),
                   "l"(&tensor_map), "r"(x), "r"(y), "r"(s32(&rbar)) : "memory");
    }
    uint32_t ok = 0;
    for (int n = 0; n < (1 << 20) && !ok; n++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&rbar)), "r"(0) : "memory");
    __syncthreads();
  }
  for (int e = threadIdx.x; e < 256; e += blockDim.x) out[e] = dst[e];
}
typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 1024, H = argc > 2 ? atoi(argv[2]) : 256, x = argc > 3 ? atoi(argv[3]) : 32, y = argc > 4 ? atoi(argv[4]) : 8;
  const int style = argc > 5 ? atoi(argv[5]) : 0;
  std::vector<float> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 1000003);
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1024); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t str[1] = {(cuuint64_t)W * 4}; cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
  CUresult r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  kernel<<<1, 128, style == 2 ? 2048 : 0>>>(m, x, y, out, style);
  cudaError_t e = cudaDeviceSynchronize();
  float ho[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
  printf("W %d H %d x %d y %d style %d: encode %d, %s, first %g %g (want %g %g)\n", W, H, x, y, style, (int)r, cudaGetErrorString(e), ho[0], ho[1],
         h[(size_t)y * W + x], h[(size_t)y * W + x + 1]);
  return 0;
}
