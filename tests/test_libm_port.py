"""wrf_b200/csrc/libm_port.h restates glibc's powf/logf/expf/sinf so that the device
evaluates the reference's transcendentals bit-identically.  Here the same header is compiled
for the host and compared with the system libm (the one the oracle, like the reference's
gfortran build, calls) over tens of millions of arguments, including every argument class the
path produces: pressures ratios near 1, tiny/zero rain mixing ratios, Exner values, the
saturation-vapour-pressure exponent range and the damping-profile angles in [0, pi/2]."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def shim():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libm_port_shim.so")
    src = os.path.join(ROOT, "tests", "libm_port_shim.cpp")
    hdr = os.path.join(ROOT, "wrf_b200", "csrc", "libm_port.h")
    hdr2 = os.path.join(ROOT, "wrf_b200", "csrc", "exact_div.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    L = ctypes.CDLL(so)
    L.port_scan.restype = ctypes.c_longlong
    L.port_scan.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float,
                            ctypes.POINTER(ctypes.c_uint32)]
    L.div_scan.restype = ctypes.c_longlong
    L.div_scan.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
    for f in ("port_powf", "port_logf", "port_expf", "port_sinf"):
        getattr(L, f).restype = ctypes.c_float
        getattr(L, f).argtypes = [ctypes.c_float] * (2 if f == "port_powf" else 1)
    return L


def bits(x):
    return np.float32(x).view(np.uint32)


def glibc_new_enough():
    try:
        v = os.confstr("CS_GNU_LIBC_VERSION").split()[1].split(".")
        return (int(v[0]), int(v[1])) >= (2, 28)
    except Exception:
        return False


pytestmark = pytest.mark.skipif(not glibc_new_enough(), reason="host libm is not the glibc >= 2.28 family the port restates")

# exponents that occur on the path: cp/cv (calc_p_rho_phi), R/cp (pi_phy), Kessler's 0.1364, .875, 1004/287, .2046, .525
EXPONENTS = [7.0 / 5.0, 287.0 / 1004.5, 0.1364, 0.875, 1004.0 / 287.0, 0.2046, 0.525]


@pytest.mark.parametrize("which,name", [(0, "logf"), (1, "expf"), (2, "sinf")])
def test_univariate_all_float_classes(shim, which, name):
    """every 211th float bit pattern over the whole positive AND negative range (subnormals, inf, NaN included)"""
    fb = ctypes.c_uint32(0)
    bad = shim.port_scan(which, 0, 0xFFFFFFFF, 211, 0.0, ctypes.byref(fb))
    assert bad == 0, f"{name}: {bad} mismatches, first at bits {fb.value:#x}"


@pytest.mark.parametrize("y", EXPONENTS)
def test_powf_positive_range(shim, y):
    fb = ctypes.c_uint32(0)
    bad = shim.port_scan(3, 0, 0x7F800001, 499, y, ctypes.byref(fb))
    assert bad == 0, f"powf(x,{y}): {bad} mismatches, first at bits {fb.value:#x}"


def test_dense_scan_of_path_ranges(shim):
    """stride-1 scans of the ranges the hot path actually evaluates"""
    fb = ctypes.c_uint32(0)
    for a, b, which, y in [
        (bits(0.9), bits(1.1), 0, 0.0),            # log(pfd/pfu)
        (bits(1.0), bits(1.5), 0, 0.0),
        (bits(0.2), bits(1.3), 3, 7.0 / 5.0),      # (Rd*theta/(p0*alt))**gamma : one octave + neighbours
        (bits(1e-9), bits(2e-9), 3, 0.1364),       # rain terminal velocity
        (bits(1e-4), bits(2e-4), 3, 0.525),
        (bits(0.3), bits(1.05), 3, 287.0 / 1004.5),
        (bits(0.0), bits(1.5707964), 2, 0.0),      # 0.5*pi*dcoef, dcoef in [0,1]: all floats in [0, pi/2]
    ]:
        step = max(1, (int(b) - int(a)) // 4_000_000)
        bad = shim.port_scan(which, int(a), int(b), step, y, ctypes.byref(fb))
        assert bad == 0, (which, y, hex(fb.value))
    # saturation-vapour-pressure exponent svp2*(T-svpt0)/(T-svp3): about [-40, 20]
    assert shim.port_scan(1, int(bits(1e-3)), int(bits(40.0)), 97, 0.0, ctypes.byref(fb)) == 0
    assert shim.port_scan(1, int(bits(-1e-3)), int(bits(-40.0)), 97, 0.0, ctypes.byref(fb)) == 0


def test_special_values(shim):
    assert shim.port_powf(0.0, 0.1364) == 0.0            # qrr = 0 -> vt = 0 (module_mp_kessler.F:101)
    assert shim.port_powf(0.0, 0.875) == 0.0
    assert shim.port_powf(1.0, 1.4) == 1.0
    assert shim.port_logf(1.0) == 0.0
    assert shim.port_expf(0.0) == 1.0
    assert shim.port_sinf(0.0) == 0.0
    assert np.isinf(shim.port_expf(1000.0)) and shim.port_expf(-1000.0) == 0.0
    assert np.isnan(shim.port_logf(-1.0)) and np.isinf(shim.port_logf(0.0))


@pytest.mark.parametrize("d", [60, 12, 981])
def test_exact_division_by_constant(shim, d):
    """csrc/exact_div.h: x/60, x/12, x/g in three FMA-class instructions, bit-identical to IEEE
    division (exhaustively verified offline over all 2^32 inputs; here every 61st bit pattern of
    both signs plus the whole subnormal / tiny range where the guard switches to plain division)"""
    fb = ctypes.c_uint32(0)
    assert shim.div_scan(d, 0, 0xFFFFFFFF, 61, ctypes.byref(fb)) == 0, hex(fb.value)
    assert shim.div_scan(d, 0, 0x0E000000, 7, ctypes.byref(fb)) == 0, hex(fb.value)
    assert shim.div_scan(d, 0x80000000, 0x8E000000, 7, ctypes.byref(fb)) == 0, hex(fb.value)
