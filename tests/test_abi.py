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
