"""Host-side mirror of the reference interface for the solve_em path.

`Solver` binds the C-ABI of include/wrfgpu.h (libwrfgpu.so, built in-tree by
wrf_b200/csrc/Makefile) with ctypes.  It is the Python stand-in for the Fortran
shim wrf_b200/fortran/solve_em_gpu.F90: same entry points, same field names
(Registry names without `grid%`), same error behaviour (a non-zero status becomes
an exception, the way the shim turns it into wrf_error_fatal).  There is no CPU
fallback: creating a Solver without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .ideal import Case, Config, Dims

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PHASES = ("rk_step_prep", "rk_phys_bc_dry_1", "init_zero_tendency", "diffusion_prep", "rk_tendency", "small_step_prep",
          "advance_uv", "advance_mu_t", "advance_w", "sumflux", "calc_p_rho", "small_step_finish",
          "rk_scalar_step", "calc_p_rho_phi", "rk_phys_bc_dry_2", "microphysics", "end_of_step", "start_domain")

# every symbol include/wrfgpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "wrfgpu_create", "wrfgpu_destroy", "wrfgpu_last_error", "wrfgpu_version", "wrfgpu_device_count", "wrfgpu_field_size", "wrfgpu_set_field",
    "wrfgpu_get_field", "wrfgpu_upload", "wrfgpu_download", "wrfgpu_num_fields", "wrfgpu_field_name", "wrfgpu_field_devptr",
    "wrfgpu_step", "wrfgpu_solve_em", "wrfgpu_sync", "wrfgpu_set_dtbc", "wrfgpu_get_dtbc", "wrfgpu_set_stream", "wrfgpu_kernel_report", "wrfgpu_launch_count", "wrfgpu_max_cfl", "wrfgpu_phase_ms", "wrfgpu_set_physical_bc",
    "wrfgpu_compute_mesh", "wrfgpu_patch_dims", "wrfgpu_comm_unique_id", "wrfgpu_comm_init", "wrfgpu_comm_init_local", "wrfgpu_halo",
    "wrfgpu_halo_spec", "wrfgpu_halo_box", "wrfgpu_halo_box2", "wrfgpu_halo_regular_width", "wrfgpu_halo_bytes",
    "wrfgpu_snapshot_begin", "wrfgpu_snapshot_wait", "wrfgpu_snapshot_busy", "wrfgpu_snapshot_host",
) + tuple("wrfgpu_" + p for p in PHASES)


def library_path() -> str:
    # WRFGPU_LIB: an alternative build of the same library (tuning experiments)
    return os.environ.get("WRFGPU_LIB") or os.path.join(_HERE, "libwrfgpu.so")


def load():
    """Load libwrfgpu.so; raises (loudly) when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        so = library_path()
        if not os.path.exists(so):
            raise RuntimeError(f"{so} is missing: build it with `make -C wrf_b200/csrc` (or __graft_entry__.build()); "
                               "there is no CPU fallback for the solve_em path")
        lib = ctypes.CDLL(so)
        lib.wrfgpu_last_error.restype = ctypes.c_char_p
        lib.wrfgpu_field_size.restype = ctypes.c_size_t
        lib.wrfgpu_field_name.restype = ctypes.c_char_p
        lib.wrfgpu_field_devptr.restype = ctypes.c_void_p
        lib.wrfgpu_launch_count.restype = ctypes.c_longlong
        lib.wrfgpu_phase_ms.restype = ctypes.c_double
        lib.wrfgpu_kernel_report.restype = ctypes.c_size_t
        lib.wrfgpu_snapshot_host.restype = ctypes.c_void_p
        _LIB = lib
    return _LIB


class WrfGpuError(RuntimeError):
    pass


class Solver:
    """One domain on one GPU (one MPI rank of wrf.exe)."""

    def __init__(self, case: Case, device: int = 0, upload: bool = True):
        self.lib = load()
        self.case = case
        self.h = ctypes.c_void_p()
        rc = self.lib.wrfgpu_create(ctypes.byref(case.dims), ctypes.byref(case.config), ctypes.c_int(device), ctypes.byref(self.h))
        if rc != 0:
            self.h = None
            raise WrfGpuError(self.lib.wrfgpu_last_error().decode())
        if upload:
            for k, a in case.fields.items():
                self.set(k, a)

    def _chk(self, rc):
        if rc != 0:
            raise WrfGpuError(self.lib.wrfgpu_last_error().decode())

    # ---- state ----
    def names(self):
        return [self.lib.wrfgpu_field_name(self.h, i).decode() for i in range(self.lib.wrfgpu_num_fields(self.h))]

    def size(self, name) -> int:
        return int(self.lib.wrfgpu_field_size(self.h, name.encode()))

    def set(self, name, a):
        """a: numpy array (host) or an object with data_ptr()/numel() (torch tensor, host or device)."""
        if hasattr(a, "data_ptr"):
            ptr, n = ctypes.c_void_p(a.data_ptr()), a.numel()
        else:
            a = np.ascontiguousarray(a, dtype=np.float32)
            ptr, n = a.ctypes.data_as(ctypes.c_void_p), a.size
        self._chk(self.lib.wrfgpu_set_field(self.h, name.encode(), ptr, ctypes.c_size_t(n)))

    def get(self, name, out=None):
        n = self.size(name)
        if n == 0:
            raise KeyError(name)
        if out is not None and hasattr(out, "data_ptr"):
            self._chk(self.lib.wrfgpu_get_field(self.h, name.encode(), ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n)))
            return out
        a = np.empty(n, np.float32) if out is None else out
        self._chk(self.lib.wrfgpu_get_field(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n)))
        d = self.case.dims
        ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
        if n == ni * nk * nj:
            return a.reshape(nj, nk, ni)
        if n == ni * nj:
            return a.reshape(nj, ni)
        return a

    def devptr(self, name) -> int:
        return int(self.lib.wrfgpu_field_devptr(self.h, name.encode()) or 0)

    # ---- hot path ----
    def step(self, n=1):
        self._chk(self.lib.wrfgpu_step(self.h, ctypes.c_int(n)))

    def sync(self):
        self._chk(self.lib.wrfgpu_sync(self.h))

    def set_dtbc(self, dtbc: float):
        """grid%dtbc: seconds since the uploaded boundary arrays were valid (0 right after a boundary time was read)"""
        self._chk(self.lib.wrfgpu_set_dtbc(self.h, ctypes.c_float(dtbc)))

    def dtbc(self) -> float:
        v = ctypes.c_float()
        self._chk(self.lib.wrfgpu_get_dtbc(self.h, ctypes.byref(v)))
        return v.value

    def solve_em(self, in_names, in_ptrs, out_names, out_ptrs):
        """Strict drop-in call: upload `in`, one solve_em, download `out` (host pointers as ints)."""
        def tab(names, ptrs):
            n = len(names)
            return n, (ctypes.c_char_p * n)(*[s.encode() for s in names]), (ctypes.c_void_p * n)(*ptrs)
        ni, nn, pn = tab(in_names, in_ptrs)
        no, on, po = tab(out_names, out_ptrs)
        self._chk(self.lib.wrfgpu_solve_em(self.h, ni, nn, pn, no, on, po))

    def call(self, phase, *args):
        self._chk(getattr(self.lib, "wrfgpu_" + phase)(self.h, *[ctypes.c_int(int(x)) for x in args]))

    def start_domain(self):
        self.call("start_domain")

    def bc(self, name, var):
        self._chk(self.lib.wrfgpu_set_physical_bc(self.h, name.encode(), ctypes.c_char(var.encode())))

    def halo(self, name):
        self._chk(self.lib.wrfgpu_halo(self.h, name.encode()))

    # ---- history / restart snapshots ----
    def snapshot_begin(self, names):
        """Stage `names` for output: returns as soon as the device-to-device copies are enqueued; the time loop may go on."""
        n = len(names)
        self._chk(self.lib.wrfgpu_snapshot_begin(self.h, ctypes.c_int(n), (ctypes.c_char_p * n)(*[s.encode() for s in names])))
        self._snap_names = list(names)

    def snapshot_busy(self) -> bool:
        r = int(self.lib.wrfgpu_snapshot_busy(self.h))
        if r < 0:
            raise WrfGpuError(self.lib.wrfgpu_last_error().decode())
        return bool(r)

    def snapshot_wait(self):
        """Block until the snapshot has landed in pinned host memory; returns {name: array shaped like get(name)} (copies)."""
        self._chk(self.lib.wrfgpu_snapshot_wait(self.h))
        out = {}
        for name in getattr(self, "_snap_names", []):
            n = ctypes.c_size_t(0)
            p = self.lib.wrfgpu_snapshot_host(self.h, name.encode(), ctypes.byref(n))
            if not p:
                raise WrfGpuError(f"no snapshot of '{name}'")
            a = np.ctypeslib.as_array((ctypes.c_float * n.value).from_address(p)).copy()
            d = self.case.dims
            ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
            out[name] = a.reshape(nj, nk, ni) if n.value == ni * nk * nj else (a.reshape(nj, ni) if n.value == ni * nj else a)
        return out

    # ---- ranks ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        """128-byte ncclUniqueId; make it on rank 0 and broadcast it with the host's own communicator"""
        lib = load()
        _preload_nccl()
        buf = ctypes.create_string_buffer(128)
        if lib.wrfgpu_comm_unique_id(buf) != 0:
            raise WrfGpuError(lib.wrfgpu_last_error().decode() or "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """ranks are processes, one per GPU: halos travel by NCCL over NVLink"""
        _preload_nccl()
        self._chk(self.lib.wrfgpu_comm_init(self.h, ctypes.c_char_p(unique_id), ctypes.c_int(rank), ctypes.c_int(nranks)))

    def halo_bytes(self) -> int:
        self.lib.wrfgpu_halo_bytes.restype = ctypes.c_longlong
        return int(self.lib.wrfgpu_halo_bytes(self.h))

    def max_cfl(self, reset=False):
        """grid%max_vert_cfl, grid%max_horiz_cfl (solve_em.F:914-930) over all ranks; needs use_adaptive_time_step = 1"""
        mv, mh = ctypes.c_float(), ctypes.c_float()
        self._chk(self.lib.wrfgpu_max_cfl(self.h, ctypes.byref(mv), ctypes.byref(mh), ctypes.c_int(int(reset))))
        return mv.value, mh.value

    def launch_count(self) -> int:
        return int(self.lib.wrfgpu_launch_count(self.h))

    def phase_timing(self, on=True):
        self.lib.wrfgpu_phase_ms(self.h, None if on else b"__off__")

    def set_stream(self, cuda_stream: int):
        self._chk(self.lib.wrfgpu_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def kernel_report(self) -> dict:
        """{kernel: (launches, total_ms)} collected while phase_timing is on."""
        n = self.lib.wrfgpu_kernel_report(self.h, None, ctypes.c_size_t(0))
        buf = ctypes.create_string_buffer(n + 16)
        self.lib.wrfgpu_kernel_report(self.h, buf, ctypes.c_size_t(n + 16))
        out = {}
        for line in buf.value.decode().splitlines():
            k, c, ms = line.split()
            out[k] = (int(c), float(ms))
        return out

    def phase_ms(self, name) -> float:
        return float(self.lib.wrfgpu_phase_ms(self.h, name.encode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.wrfgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _preload_nccl():
    """make the NCCL that ships with torch visible to dlopen("libnccl.so.2") in the C library"""
    if os.environ.get("WRFGPU_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            so = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(so):
                ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
                os.environ["WRFGPU_NCCL_LIB"] = so
    except Exception:
        pass


def comm_init_local(solvers):
    """ranks are solvers of this process (rank order): neighbours read each other's send buffers"""
    lib = load()
    arr = (ctypes.c_void_p * len(solvers))(*[s.h for s in solvers])
    if lib.wrfgpu_comm_init_local(arr, ctypes.c_int(len(solvers))) != 0:
        raise WrfGpuError(lib.wrfgpu_last_error().decode())


def run_ranks(solvers, fn):
    """call fn(solver) for every rank concurrently (one host thread per rank, as wrf.exe has one
    process per rank); exceptions are re-raised"""
    import threading
    errs = []

    def work(s):
        try:
            fn(s)
        except BaseException as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(s,)) for s in solvers]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
