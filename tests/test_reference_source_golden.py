"""The oracle against what the reference's own source computed (tests/test_reference_source.py), from a committed fixture.

tests/golden/reference_source.json holds, for every phase / Runge-Kutta stage / configuration that module compares, a digest of every
field as the REFERENCE's Fortran routines left it when they were executed in the build container (through tests/fortran_subset.py, on
the state an Oracle had reached).  This module needs no reference tree: it brings an Oracle to the same point, runs the oracle's phase
and requires the same digests -- so the pin travels with the repository (the GPU box has no /root/reference), and a change of the
oracle, the compiler or libm that moves a single bit of any of these fields fails here.

Regenerate (build container only):  REFSRC_WRITE=1 python -m pytest tests/test_reference_source.py"""
import json
import os

import numpy as np
import pytest

from tests.refcases import at_phase, build, digest, oracle_feeder, own

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_source.json")
RECORDS = json.load(open(GOLDEN))["records"] if os.path.exists(GOLDEN) else []


def test_fixture_is_there():
    assert len(RECORDS) >= 200, "tests/golden/reference_source.json is missing or short: regenerate it in the build container"
    phases = {r["phase"] for r in RECORDS}
    assert {"rk_step_prep", "rk_tendency", "small_step_prep", "advance_uv", "advance_mu_t", "advance_w", "sumflux", "calc_p_rho",
            "small_step_finish", "rk_scalar_step", "calc_p_rho_phi", "microphysics", "diffusion_prep", "end_of_step", "rk_phys_bc_dry_1",
            "rk_phys_bc_dry_2"} <= phases


@pytest.mark.parametrize("n", range(len(RECORDS)))
def test_oracle_reproduces_reference_digest(n):
    r = RECORDS[n]
    case = build(r["case"])
    o, args = at_phase(case, r["phase"], r["rk"], r["occurrence"])
    o.call(r["phase"], *args)
    bad = []
    for k, want in r["fields"].items():
        a = np.asarray(o.get(k))
        if digest(a if r["whole"] else own(case, k, a)) != want:
            bad.append(k)
    assert not bad, f"{r['case']} {r['phase']} rk {r['rk']}: the oracle no longer equals the reference source in {bad}"


FEEDER_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "feeder_columns.json")
FEEDER = json.load(open(FEEDER_GOLDEN)) if os.path.exists(FEEDER_GOLDEN) else {}


@pytest.mark.parametrize("adv_opt", [0, 1, 2, 3, 4])
def test_feeder_output_files(adv_opt):
    """PROGRAM feeder (dyn_em/module_advect_em.F:10567, the reference's own kernel harness): what it writes to 000200.txt -- NINT(field*100)
    along x after 200 trips of its loop -- for each of its five scalar_adv_opt settings, obtained by running the reference's kernels from
    source inside the program's loop (tests/feeder_ref.py; fixture tests/golden/feeder_columns.json).  The oracle's rk_scalar_step, driven
    the same way, must write the same file; also half way (loop 100)."""
    assert str(adv_opt) in FEEDER, "tests/golden/feeder_columns.json is missing: REFSRC_FEEDER=1 python -m pytest tests/test_reference_source.py -k feeder_harness"
    want = {int(k): v for k, v in FEEDER[str(adv_opt)].items()}
    cols, row = oracle_feeder(adv_opt, max(want), tuple(want))
    assert cols == want
    hat = np.array(want[200])
    assert hat.max() >= 95
    if adv_opt in (1, 2, 4):
        assert hat.min() >= 0            # PD / monotonic / WENO-PD stay non-negative
    if adv_opt == 0:
        assert hat.min() < 0             # the unlimited 5th-order scheme undershoots


INIT_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "shipped_init.json")
INIT = json.load(open(INIT_GOLDEN)) if os.path.exists(INIT_GOLDEN) else {}


@pytest.mark.parametrize("name", ["em_quarter_ss", "em_hill2d_x", "em_les", "em_b_wave"])
def test_shipped_initial_state_digests(name):
    """wrf_b200/shipped.py's initial state for the four shipped run directories against digests of what the REFERENCE's init_domain_rk
    (module_initialize_ideal.F) + the first trip of start_domain_em produced when executed from source in the build container
    (tests/test_reference_init.py wrote tests/golden/shipped_init.json): 53 fields and 6 scalars per case, bit for bit."""
    from wrf_b200 import shipped
    assert name in INIT, "tests/golden/shipped_init.json is missing: REFSRC_WRITE=1 python -m pytest tests/test_reference_init.py -k initial_state"
    case = shipped.init_domain(os.path.join(os.path.dirname(__file__), "fixtures", name))
    d = case.dims
    sj, si = slice(1 - d.jms, d.jde - d.jms + 1), slice(1 - d.ims, d.ide - d.ims + 1)
    bad = []
    for k, want in INIT[name].items():
        if isinstance(want, float):
            if np.float32(getattr(case.config, k)) != np.float32(want):
                bad.append(k)
            continue
        a = case.fields[k]
        if digest(a[sj, ..., si] if a.ndim >= 2 else a) != want:
            bad.append(k)
    assert not bad, f"{name}: {bad}"
    assert len(INIT[name]) >= 55
