"""The diffusion preamble (compute_diff_metrics, cal_deform_and_div, smag2d_km -- SURVEY.md 8f-1) on the device is written
as __host__ __device__ point functions (wrf_b200/csrc/deform.cuh).  Here the same header is compiled for the host
(tests/deform_shim.cu, nvcc, no GPU needed), run in plain loops and compared with the oracle's restatement of the reference
routines bit for bit, on every lateral boundary type -- so the GPU parity tests of these kernels only have to confirm that
the device evaluates the same IEEE operations."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.parity import small_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
IN3 = ("ph_2", "phb", "u_2", "v_2", "w_2", "t_2", "alt", "p", "pb")
OUT3 = ("z", "rdz", "rdzw", "zx", "zy", "div", "defor11", "defor22", "defor33", "defor12", "defor13", "defor23", "xkmh", "xkmv", "xkhh", "xkhv")
D2 = ("bn2", "rho", "ru_tendf", "rv_tendf", "rw_tendf", "t_tendf", "ustm", "hfx", "qfx")
REST = ("mut", "c1h", "c2h", "msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvy", "fnm", "fnp", "dn", "dnw", "u_base", "v_base", "t_base", "qv_base")


@pytest.fixture(scope="module")
def shim():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "deform_shim.so")
    src = os.path.join(ROOT, "tests", "deform_shim.cu")
    deps = [src] + [os.path.join(ROOT, "wrf_b200", "csrc", f) for f in ("deform.cuh", "diffusion2.cuh", "common.cuh")] + [os.path.join(ROOT, "include", "wrfgpu.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", so, src])
    L = ctypes.CDLL(so)
    L.deform_shim_run.restype = ctypes.c_int
    return L


def run_shim(L, case, arrays, stage, last=(), which=0):
    names = [n for n in arrays if n not in last] + list(last)
    cn = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
    ptrs = (ctypes.c_void_p * len(names))(*[arrays[n].ctypes.data for n in names])
    assert L.deform_shim_run(ctypes.c_int(stage), ctypes.byref(case.dims), ctypes.byref(case.config), ctypes.c_int(len(names)), cn, ptrs,
                             ctypes.c_int(which)) == 0


@pytest.mark.parametrize("name", ["smag2d", "smag2d_periodic", "smag2d_hill", "smag2d_b_wave", "diff2", "diff2_periodic", "diff2_hill", "diff2_b_wave",
                                  "smag3d", "smag3d_periodic", "smag3d_hill", "tke", "tke_periodic", "tke_hill"])
def test_device_point_functions_equal_the_oracle(shim, name):
    case = small_cases()[name]
    c = case.config
    o = Oracle(case)
    o.start_domain()
    o.step(2)   # the state (and the stale contents of the work fields) of a running model
    for ph, args in (("rk_step_prep", (1,)), ("rk_phys_bc_dry_1", ()), ("init_zero_tendency", ())):
        o.call(ph, *args)
    species = [(n, n.replace("moist_", "moist_tend_").replace("scalar_", "scalar_tend_")) for n in o.names()
               if n.startswith(("moist_q", "scalar_")) and "old" not in n and "tend" not in n]
    have = set(o.names())
    names = [n for n in IN3 + OUT3 + D2 + REST if n in have] + [n for pair in species for n in pair]
    arrays = {n: np.ascontiguousarray(o.get(n)).copy() for n in names}
    tke = None
    if c.km_opt == 2:   # the tke species is the last one of the scalar family; the shim takes it under its Registry names
        tke = "scalar_%d" % (c.num_scalar + 1)
        arrays["tke_2"] = arrays[tke]
        arrays["tke_tend"] = arrays[tke.replace("scalar_", "scalar_tend_")]
    applier = Oracle(case)   # set_physical_bc of the oracle, applied to the shim's arrays between the stages

    def bc(n, var):
        applier.set(n, arrays[n])
        applier.bc(n, var)
        arrays[n][...] = applier.get(n)

    if c.diff_opt == 2:
        run_shim(shim, case, arrays, 0)
    run_shim(shim, case, arrays, 1)
    for n, var in (("rdzw", "w"), ("rdz", "w"), ("z", "w"), ("zx", "e"), ("zy", "f")):
        bc(n, var)
    run_shim(shim, case, arrays, 2)
    run_shim(shim, case, arrays, {1: 4, 2: 8, 3: 7, 4: 3}[c.km_opt])
    bc("xkmh", "t"); bc("xkhh", "t")
    out = list(OUT3) + (["bn2"] if c.km_opt in (2, 3) else [])
    if c.diff_opt == 2:
        for n, var in (("xkmv", "t"), ("xkhv", "t"), ("div", "t"), ("defor11", "t"), ("defor22", "t"), ("defor33", "t"), ("defor12", "d"),
                       ("defor13", "e"), ("defor23", "f"), ("rho", "t")):
            bc(n, var)
        if tke:
            run_shim(shim, case, arrays, 9)
        run_shim(shim, case, arrays, 5)
        for f, t in species:
            run_shim(shim, case, arrays, 6, last=(f, t), which=1 if f == "moist_qv" else (3 if f == tke else 2))
        out += ["rho", "ru_tendf", "rv_tendf", "rw_tendf", "t_tendf", "hfx"] + [t for _, t in species]
    o.call("diffusion_prep")
    bad = []
    for n in out:
        ref = o.get(n)
        a, b = arrays[n].view(np.uint32), ref.view(np.uint32)
        if not np.array_equal(a, b):
            w = np.argwhere(a != b)
            bad.append((n, len(w), w[:3].tolist(), float(np.abs(arrays[n] - ref).max())))
    assert not bad, bad
    assert float(o.get("xkmh").max()) > 0.0
    if tke:
        tt = o.get(tke.replace("scalar_", "scalar_tend_"))
        assert float(tt.max()) > 0.0 > float(tt.min())        # production and dissipation both act
    if name == "smag3d":   # both branches of calculate_N2 and both limits of smag_km are exercised
        lo = float(o.get("xkmh")[o.get("xkmh") > 0].min())
        assert float(o.get("xkmh").max()) > 1.5 * lo and float(o.get("bn2").min()) < 0.0 < float(o.get("bn2").max())
    if c.diff_opt == 2:
        assert all(float(np.abs(o.get(n)).max()) > 0.0 for n in ("ru_tendf", "rw_tendf", "t_tendf"))
