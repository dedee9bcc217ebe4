"""The C-ABI boundary: libwrfgpu.so loads and exports every symbol include/wrfgpu.h declares;
the host-only entry points (decomposition rules) work without a GPU; creating a solver
without a GPU fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

import wrf_b200
from wrf_b200 import lib as wlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "wrfgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wrfgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = wrf_b200.load()
    syms = header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python mirror binds nothing the header does not declare
    assert set(wlib.ABI_SYMBOLS) <= set(syms), set(wlib.ABI_SYMBOLS) - set(syms)


def test_version():
    assert wrf_b200.load().wrfgpu_version() >= 100


def test_struct_layout_matches_header():
    """ctypes mirrors of wrfgpu_dims / wrfgpu_config have the header's field order and size."""
    text = open(os.path.join(ROOT, "include", "wrfgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for cname, cls in (("wrfgpu_dims", wrf_b200.Dims), ("wrfgpu_config", wrf_b200.Config)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, rest = decl.split(None, 1)
            assert ty in ("int", "float"), decl
            for n in rest.split(","):
                names.append((n.strip(), ty))
        mine = [(n, "int" if t is ctypes.c_int else "float") for n, t in cls._fields_]
        assert mine == names
        assert ctypes.sizeof(cls) == 4 * len(names)


def test_no_cpu_fallback(have_gpu):
    """Without a CUDA device wrfgpu_create must return an error -- never a CPU path."""
    if have_gpu:
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU runner")
    case = wrf_b200.make_case("quarter_ss", nx=12, ny=12, nz=9)
    with pytest.raises(wrf_b200.WrfGpuError) as e:
        wrf_b200.Solver(case)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """Nothing under wrf_b200/ (the product) may reference oracle/ (test infrastructure)."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "wrf_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".F90", "Makefile")):
                t = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"liboracle|oracle/|import oracle|from oracle|oracle_lib", t):
                    bad.append(fn)
    assert not bad, bad


def _build_driver():
    import subprocess
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "abi_driver")
    libdir = os.path.join(ROOT, "wrf_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "abi_driver.c"), "-L" + libdir, "-lwrfgpu",
                           "-Wl,-rpath," + libdir])
    return exe


def _write_state(case, path):
    import numpy as np
    with open(path, "wb") as f:
        f.write(bytes(case.dims)); f.write(bytes(case.config))
        f.write(np.int32(len(case.fields)).tobytes())
        for k, a in case.fields.items():
            a = np.ascontiguousarray(a, dtype=np.float32)
            f.write(k.encode().ljust(24, b"\0")[:24]); f.write(np.int64(a.size).tobytes()); f.write(a.tobytes())


def _read_state(path):
    import numpy as np
    out = {}
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(4), np.int32)[0])
        for _ in range(n):
            name = f.read(24).split(b"\0")[0].decode()
            cnt = int(np.frombuffer(f.read(8), np.int64)[0])
            out[name] = np.frombuffer(f.read(4 * cnt), np.float32).copy()
    return out


def test_c_driver_links_against_the_abi_and_fails_loudly_without_a_gpu(have_gpu, tmp_path):
    """tests/abi_driver.c (the Fortran shim's call sequence in C) compiles against include/wrfgpu.h and links with
    libwrfgpu.so; on a box without a GPU it ends like the shim: a fatal message, no CPU path"""
    import subprocess
    exe = _build_driver()
    if have_gpu:
        pytest.skip("a GPU is present; the run itself is test_c_driver_replays_the_shim_call_sequence")
    case = wrf_b200.make_case("quarter_ss", nx=12, ny=12, nz=9)
    st = str(tmp_path / "state.bin")
    _write_state(case, st)
    r = subprocess.run([exe, st, "1", str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["tke", "specified"])
def test_c_driver_replays_the_shim_call_sequence(tmp_path, kind):
    """create -> upload (pointer table) -> start_domain -> [set_dtbc ->] step x n -> download through plain C equals the
    ctypes mirror; with specified lateral boundaries the boundary arrays travel in the same pointer table"""
    import subprocess
    import numpy as np
    exe = _build_driver()
    if kind == "tke":
        case = wrf_b200.make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=2, mix_full_fields=1)
    else:
        from tests import parity
        case = parity.small_cases()["specified"]
    st, out = str(tmp_path / "state.bin"), str(tmp_path / "out.bin")
    _write_state(case, st)
    r = subprocess.run([exe, st, "3", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read_state(out)
    s = wrf_b200.Solver(case)
    s.start_domain(); s.step(3)
    assert {"u_2", "w_2", "t_2", "ph_2", "mu_2", "moist_qv", "moist_qr", "rainnc"} | ({"scalar_1"} if kind == "tke" else set()) <= set(got)
    for n, a in got.items():
        assert np.array_equal(a, s.get(n).ravel()), n
    s.close()


def test_fortran_binding_lists_the_header_members():
    """the BIND(C) types of wrf_b200/fortran/module_wrfgpu.F90 name the members of wrfgpu_dims / wrfgpu_config in the
    header's order and with the header's types (the shim cannot be compiled here, so at least its layout is pinned)"""
    src = open(os.path.join(ROOT, "wrf_b200", "fortran", "module_wrfgpu.F90")).read()
    for cname, cls in (("wrfgpu_dims", wrf_b200.Dims), ("wrfgpu_config", wrf_b200.Config)):
        body = re.search(r"TYPE, BIND\(C\) :: %s(.*?)END TYPE %s" % (cname, cname), src, flags=re.S).group(1)
        names = []
        for ln in body.splitlines():
            ln = ln.split("!")[0].strip()
            m = re.match(r"(INTEGER\(c_int\)|REAL\(c_float\))\s*::\s*(.*)", ln)
            if m:
                names += [(x.strip(), "int" if "INTEGER" in m.group(1) else "float") for x in m.group(2).split(",")]
        mine = [(n, "int" if t is ctypes.c_int else "float") for n, t in cls._fields_]
        assert names == mine, cname
    shim = open(os.path.join(ROOT, "wrf_b200", "fortran", "solve_em_gpu.F90")).read()
    for member, _ in wrf_b200.Config._fields_:
        assert re.search(r"\bc%%%s\b" % member, shim), f"solve_em_gpu.F90 never sets c%{member}"
    assert "Is_alarm_tstep(grid%domain_clock, grid%alarms(HISTORY_ALARM))" in shim and "restart_flag_alarm_rings" not in shim
