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
