// flux.cuh -- the reference's flux operators (statement functions of dyn_em/module_advect_em.F:202-213),
// shared by the gather kernels (tend.cu) and the k-marching interior kernels (advect_fast.cu).
#pragma once
#include "common.cuh"
#include "exact_div.h"

namespace wg {

// module_advect_em.F:202-213
__device__ __forceinline__ float flux4(float qm2, float qm1, float q0, float qp1) {
  return WG_DIV12(7.f * (q0 + qm1) - (qp1 + qm2));
}
__device__ __forceinline__ float flux3(float qm2, float qm1, float q0, float qp1, float ua, float sg) {
  return flux4(qm2, qm1, q0, qp1) + WG_DIV12(sg * sign1(ua) * ((qp1 - qm2) - 3.f * (q0 - qm1)));
}
__device__ __forceinline__ float flux6(float qm3, float qm2, float qm1, float q0, float qp1, float qp2) {
  return WG_DIV60(37.f * (q0 + qm1) - 8.f * (qp1 + qm2) + (qp2 + qm3));
}
__device__ __forceinline__ float flux5(float qm3, float qm2, float qm1, float q0, float qp1, float qp2, float ua, float sg) {
  return flux6(qm3, qm2, qm1, q0, qp1, qp2) - WG_DIV60(sg * sign1(ua) * ((qp2 - qm3) - 5.f * (qp1 - qm2) + 10.f * (q0 - qm1)));
}

}  // namespace wg
