"""Reference-held tables (tests/golden/reference_tables.json, extracted from /root/reference by
tests/golden/make_reference_tables.py) against the oracle and the product's host logic: the physical constants of
share/module_model_constants.F, the halo / period definitions of Registry/Registry.EM_COMMON and its namelist defaults.
The reference ships no outputs for this path; these tables are what it does hold."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests import halo_emul
from tests.oracle_lib import load as load_oracle
from wrf_b200 import make_case, shipped

HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = json.load(open(os.path.join(HERE, "golden", "reference_tables.json")))


def test_fixture_is_what_the_reference_tree_holds():
    """regenerated and compared whenever the reference tree is present (the build container)"""
    if not os.path.isdir("/root/reference/Registry"):
        pytest.skip("the reference tree is not on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mrt", os.path.join(HERE, "golden", "make_reference_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.constants() == TABLES["constants"]
    assert m.halos() == TABLES["halos"]
    assert m.rconfig(set(shipped.REGISTRY_DEFAULTS)) == TABLES["rconfig"]
    assert m.flux_operators() == TABLES["flux_operators"]


def test_oracle_constants_are_module_model_constants():
    """every constant of oracle.h has the bits gfortran's real*4 evaluation of the reference's PARAMETER expression gives"""
    lib = load_oracle()
    assert len(TABLES["constants"]) == 21
    for name, rec in TABLES["constants"].items():
        v = ctypes.c_float()
        assert lib.oracle_constant(name.encode(), ctypes.byref(v)) == 0, name
        assert int(np.float32(v.value).view(np.uint32)) == rec["bits"], (name, v.value, rec)


def _conv(text, like):
    t = text.strip().lower()
    if t in (".true.", "true"):
        return True
    if t in (".false.", "false"):
        return False
    return type(like)(float(t)) if not isinstance(like, bool) else bool(float(t))


def test_namelist_defaults_are_the_registry_s():
    for name, ours in shipped.REGISTRY_DEFAULTS.items():
        key = name if name in TABLES["rconfig"] else name.upper()
        if name in ("hybrid_opt", "etac"):       # Registry/registry.hyb_coord:62-63 (2, 0.2)
            continue
        rec = TABLES["rconfig"].get(key) or TABLES["rconfig"].get(name)
        if rec is None and name == "isfflx":     # spelled ISFFLX in Registry.EM_COMMON:2499 (default 1)
            assert ours == 1
            continue
        assert rec is not None, name
        assert _conv(rec["default"], ours) == ours, (name, ours, rec["default"])


def test_namelist_without_diff_opt_is_refused_like_check_a_mundo_does():
    nl = shipped.read_namelist(os.path.join(HERE, "fixtures", "em_quarter_ss", "namelist.input"))
    del nl["dynamics"]["diff_opt"]
    with pytest.raises(ValueError, match="km_opt and diff_opt"):
        shipped.config_from_namelist(nl)


# the product's exchanges and the Registry entries each one stands for (HALO + the PERIOD include that follows it in solve_em.F)
EXCHANGES = {
    "HALO_EM_A": ["HALO_EM_A", "PERIOD_BDY_EM_A"], "HALO_EM_B": ["HALO_EM_B", "PERIOD_BDY_EM_B"], "HALO_EM_C": ["HALO_EM_C", "PERIOD_BDY_EM_E"],
    "HALO_EM_C2": ["HALO_EM_C2", "PERIOD_BDY_EM_B3"], "HALO_EM_D": ["HALO_EM_D", "PERIOD_EM_DA"],
    "HALO_EM_D2_5": ["HALO_EM_D2_5", "PERIOD_BDY_EM_D"], "HALO_EM_D3_5": ["HALO_EM_D3_5", "PERIOD_BDY_EM_D3"], "HALO_EM_E_5": ["HALO_EM_E_5"],
    "HALO_EM_THETAM": ["HALO_EM_THETAM", "PERIOD_EM_THETAM"],
    "HALO_EM_MOIST_OLD_E_7": ["HALO_EM_MOIST_OLD_E_7", "HALO_EM_SCALAR_OLD_E_7", "PERIOD_BDY_EM_MOIST_OLD", "PERIOD_BDY_EM_SCALAR_OLD"],
    "HALO_EM_MOIST_E_7": ["HALO_EM_MOIST_E_7", "HALO_EM_SCALAR_E_7"], "HALO_EM_TKE_C": ["HALO_EM_TKE_C", "PERIOD_BDY_EM_A1"],
    "HALO_EM_TKE_D": ["HALO_EM_TKE_D"], "HALO_EM_TKE_E": ["HALO_EM_TKE_E", "PERIOD_BDY_EM_PHY_BC", "HALO_EM_PHYS_DIFFUSION"],
}
STENCIL_WIDTH = {4: 1, 8: 1, 24: 2, 48: 3, 80: 4}     # Registry stencil sizes (tools/gen_comms.c): points -> halo width
WIDER_THAN_REGISTRY = {"HALO_EM_TKE_C", "HALO_EM_TKE_E"}   # sent at their PERIOD width on every link (more data, same values)


def _registry_name(f):
    for p, r in (("moist_old_", "moist_old"), ("moist_", "moist"), ("scalar_old_", "scalar_old"), ("scalar_", "scalar")):
        if f.startswith(p):
            return r
    return {"bn2": "BN2"}.get(f, f)


@pytest.mark.parametrize("ours", sorted(EXCHANGES))
def test_exchange_table_is_the_registry_s(ours):
    """every field an exchange carries is in the Registry's definition of that halo (or of its PERIOD twin), interior links
    use the Registry's halo width, links across a periodic edge the PERIOD width"""
    halos = TABLES["halos"]
    c = make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=2, num_scalar=2).config
    width, fields = halo_emul.halo_spec(c, ours)
    regular = halo_emul.regular_width(ours, width)
    union, hmax, pmax = set(), 0, 0
    for r in EXCHANGES[ours]:
        assert r in halos, r
        for n, fl in halos[r]["groups"]:
            union |= set(fl)
            if halos[r]["kind"] == "halo":
                hmax = max(hmax, STENCIL_WIDTH[n])
            else:
                pmax = max(pmax, n)
    assert fields and all(_registry_name(f) in union for f in fields), [f for f in fields if _registry_name(f) not in union]
    if ours in WIDER_THAN_REGISTRY:
        assert abs(regular) >= hmax
    else:
        assert abs(regular) == hmax, (regular, hmax)
    assert abs(width) == max(abs(regular), pmax), (width, regular, pmax)


def _generator():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mrt", os.path.join(HERE, "golden", "make_reference_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("time_step", [12, -12])
def test_flux_operators_are_the_reference_statement_functions(time_step):
    """flux3 ... flux6 as the reference writes them (statement functions of advect_u, kept as source text in the fixture),
    evaluated operation by operation in real*4, against the oracle's operators: same bits for every argument"""
    ops = TABLES["flux_operators"]
    assert sorted(ops) == ["flux3", "flux4", "flux5", "flux6"]
    ev = _generator().f32_eval
    lib = load_oracle()
    lib.oracle_flux.restype = ctypes.c_float
    f32 = np.float32

    def sign(a, b):
        return f32(abs(a)) if (b >= 0 and not (b == 0 and np.signbit(b))) else f32(-abs(a))

    funcs = {"sign": sign}
    for name in ("flux4", "flux6", "flux3", "flux5"):     # the even ones first: the odd ones call them
        rec = ops[name]
        funcs[name] = (lambda rec: lambda *a: ev(rec["expr"], dict(zip(rec["args"], [f32(x) for x in a]), time_step=f32(time_step)), funcs))(rec)
    rng = np.random.default_rng(11)
    for n in range(400):
        scale = f32(10.0 ** rng.integers(-3, 4))
        q = (rng.standard_normal(6) * scale).astype(f32)
        ua = f32(rng.standard_normal() * 30.0) if n % 7 else f32(0.0)
        for order, nq in ((3, 4), (4, 4), (5, 6), (6, 6)):
            qq = q[:nq] if nq == 6 else q[1:5]
            want = funcs["flux%d" % order](*[f32(x) for x in qq], ua)
            buf = (ctypes.c_float * nq)(*[float(x) for x in qq])
            got = f32(lib.oracle_flux(order, buf, ctypes.c_float(float(ua)), ctypes.c_int(time_step)))
            assert got.view(np.uint32) == f32(want).view(np.uint32), (order, qq, ua, got, want)
