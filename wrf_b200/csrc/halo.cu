// halo.cu -- RSL_LITE halo exchanges (external/RSL_LITE) on NCCL.  Filled in by the
// multi-GPU milestone; on a 1x1 mesh every HALO_EM_* is a no-op exactly as the
// generated HALO_*.inc are in a single-task run.
#include "common.cuh"
namespace wg {
void Solver::halo(const std::string &name) {
  if (nranks == 1) return;
  throw std::runtime_error("wrfgpu: halo '" + name + "' requested on a multi-rank mesh but the communicator is not initialised");
}
}  // namespace wg
extern "C" {
int wrfgpu_comm_unique_id(void *) { return 1; }
int wrfgpu_comm_init(wrfgpu_solver *, const void *, int, int) { return 1; }
}
