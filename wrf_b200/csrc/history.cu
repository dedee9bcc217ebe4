// history.cu -- history / restart snapshots that do not stall the time loop.
//
// The reference's contract is "the host grid% arrays are current when med_before_solve_io runs"
// (frame/module_integrate.F:406, share/mediation_integrate.F: med_before_solve_io -> med_hist_out / med_restart_out at
// the history and restart alarms).  With the state resident in HBM that would mean a synchronous download of every
// 'h' / 'r' field at each alarm.  Here an alarm costs one device-to-device copy of the listed fields into staging
// buffers, enqueued on the compute stream (HBM bandwidth: ~2 x bytes / 6.5 TB/s), after which the time loop goes on;
// the device-to-host copies into pinned memory run on a second stream behind an event and overlap the next steps.
// The I/O layer calls wrfgpu_snapshot_wait before it touches the buffers (wrfgpu_snapshot_host).
#include "common.cuh"
#include <algorithm>

namespace wg {

struct SnapField { std::string name; float *stage = nullptr; float *host = nullptr; size_t n = 0; };
struct Snapshot {
  cudaStream_t copy = nullptr;
  cudaEvent_t staged = nullptr, done = nullptr;
  std::map<std::string, SnapField> buf;     // allocated once per field, reused by every snapshot
  std::vector<std::string> current;         // fields of the snapshot in flight / last finished
  bool in_flight = false;
};

static Snapshot *snap_of(Solver *s) {
  if (!s->snapshot) {
    auto *p = new Snapshot;
    CUDA_CHECK(cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&p->staged, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
    s->snapshot = p;
  }
  return static_cast<Snapshot *>(s->snapshot);
}

void snapshot_wait(Solver *s) {
  if (!s->snapshot) return;
  Snapshot *p = static_cast<Snapshot *>(s->snapshot);
  if (p->in_flight) { CUDA_CHECK(cudaEventSynchronize(p->done)); p->in_flight = false; }
}

void snapshot_begin(Solver *s, int nf, const char *const *names) {
  Snapshot *p = snap_of(s);
  snapshot_wait(s);                          // the pinned buffers are about to be overwritten
  p->current.clear();
  for (int i = 0; i < nf; i++) {
    Field &f = s->F(names[i]);
    SnapField &b = p->buf[names[i]];
    if (!b.stage) {
      b.name = names[i]; b.n = f.n;
      CUDA_CHECK(cudaMalloc(&b.stage, f.n * sizeof(float)));
      CUDA_CHECK(cudaMallocHost(&b.host, f.n * sizeof(float)));
    }
    s->copy_out(f, b.stage, cudaMemcpyDeviceToDevice, s->stream);   // staging holds the caller's (unpadded) layout
    p->current.push_back(names[i]);
  }
  CUDA_CHECK(cudaEventRecord(p->staged, s->stream));
  CUDA_CHECK(cudaStreamWaitEvent(p->copy, p->staged, 0));
  for (auto &n : p->current) {
    SnapField &b = p->buf[n];
    CUDA_CHECK(cudaMemcpyAsync(b.host, b.stage, b.n * sizeof(float), cudaMemcpyDeviceToHost, p->copy));
  }
  CUDA_CHECK(cudaEventRecord(p->done, p->copy));
  p->in_flight = true;
}

int snapshot_busy(Solver *s) {
  if (!s->snapshot) return 0;
  Snapshot *p = static_cast<Snapshot *>(s->snapshot);
  if (!p->in_flight) return 0;
  cudaError_t e = cudaEventQuery(p->done);
  if (e == cudaSuccess) return 0;
  if (e != cudaErrorNotReady) CUDA_CHECK(e);
  cudaGetLastError();
  return 1;
}

const float *snapshot_host(Solver *s, const char *name, size_t *n) {
  if (!s->snapshot) return nullptr;
  Snapshot *p = static_cast<Snapshot *>(s->snapshot);
  auto it = p->buf.find(name);
  if (it == p->buf.end() || std::find(p->current.begin(), p->current.end(), name) == p->current.end()) return nullptr;
  if (n) *n = it->second.n;
  return it->second.host;
}

void snapshot_destroy(Solver *s) {
  if (!s->snapshot) return;
  Snapshot *p = static_cast<Snapshot *>(s->snapshot);
  if (p->in_flight) cudaEventSynchronize(p->done);
  for (auto &kv : p->buf) { cudaFree(kv.second.stage); cudaFreeHost(kv.second.host); }
  cudaEventDestroy(p->staged); cudaEventDestroy(p->done); cudaStreamDestroy(p->copy);
  delete p;
  s->snapshot = nullptr;
}

}  // namespace wg

#define GUARD(body)                                                 \
  try { cudaSetDevice(s->s.device); body; return 0; }              \
  catch (const std::exception &e) { g_wrfgpu_err = e.what(); return 1; }

extern "C" {
int wrfgpu_snapshot_begin(wrfgpu_solver *s, int nfields, const char *const *names) { GUARD(wg::snapshot_begin(&s->s, nfields, names)) }
int wrfgpu_snapshot_wait(wrfgpu_solver *s) { GUARD(wg::snapshot_wait(&s->s)) }
int wrfgpu_snapshot_busy(wrfgpu_solver *s) {
  try { cudaSetDevice(s->s.device); return wg::snapshot_busy(&s->s); }
  catch (const std::exception &e) { g_wrfgpu_err = e.what(); return -1; }
}
const float *wrfgpu_snapshot_host(wrfgpu_solver *s, const char *name, size_t *n) {
  try { return wg::snapshot_host(&s->s, name, n); }
  catch (const std::exception &e) { g_wrfgpu_err = e.what(); return nullptr; }
}
}
