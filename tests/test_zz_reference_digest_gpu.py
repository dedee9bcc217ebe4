"""The CUDA path against what the REFERENCE'S OWN SOURCE computed.

tests/golden/reference_source.json holds a digest of every field the reference's Fortran routines produced when they were executed in
the build container on the state of an Oracle just before a phase (tests/test_reference_source.py wrote it; the reference tree is not
needed here).  For every configuration of that fixture the device solver is stepped through one solve_em step next to the oracle;
before each recorded phase it is handed the oracle's state, runs the phase through the C-ABI, and every field it holds must have the
reference's digest -- device == reference source, bit for bit, without the oracle in between as the yardstick (the oracle only supplies
the input state, and decides NaN == NaN for never-written scratch points whose payload bits need not agree).

(The file sorts last on purpose: it widens the parity evidence and must not shadow the established suites under `pytest -x`.)"""
import json
import os

import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from tests.refcases import CASES, build, digest, own
from wrf_b200 import Solver

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_source.json")
RECORDS = json.load(open(GOLDEN))["records"] if os.path.exists(GOLDEN) else []
NAMES = sorted({r["case"] for r in RECORDS})


def test_fixture_covers_the_catalogue():
    assert len(NAMES) >= 30 and set(NAMES) <= set(CASES)


@pytest.mark.parametrize("name", NAMES)
def test_device_reproduces_reference_digests(name):
    case = build(name)
    recs = [dict(r, seen=0) for r in RECORDS if r["case"] == name]
    o, s = Oracle(case), Solver(case)
    try:
        o.start_domain(); s.start_domain()
        if case.config.specified or case.config.nested:
            o.set_dtbc(case.config.dt); s.set_dtbc(case.config.dt)
        common = set(parity.common_fields(o, s))
        checked, bad = 0, []
        for ph, args in parity.phase_sequence(case):
            todo = []
            for r in recs:                       # the occurrence counting of refcases.at_phase
                if r["phase"] == ph and (r["rk"] is None or (args and args[0] == r["rk"])):
                    if r["seen"] == r["occurrence"]:
                        todo.append(r)
                    r["seen"] += 1
            if todo:
                parity.sync_from_oracle(o, s)    # the state the reference routines were run on
            o.call(ph, *args)
            s.call(ph, *args)
            for r in todo:
                for k, want in r["fields"].items():
                    if k not in common or (ph, k) in parity.NOT_MATERIALISED:
                        continue
                    a, b = np.asarray(s.get(k)), np.asarray(o.get(k))
                    da = digest(a if r["whole"] else own(case, k, a))
                    if da != want:
                        db = digest(b if r["whole"] else own(case, k, b))
                        if not (db == want and np.array_equal(a, b, equal_nan=True)):
                            bad.append((ph, r["rk"], k))
                    checked += 1
            if todo:
                rep = parity.diff_report(o, s, sorted(common))
                parity.sync_from_oracle(o, s, [x[0] for x in rep])
        assert not bad, f"{name}: device fields that differ from the reference source's: {bad[:8]}"
        assert checked >= 5 * len(recs), f"{name}: only {checked} fields compared over {len(recs)} phases"
        assert all(r["seen"] > r["occurrence"] for r in recs), "a recorded phase was never reached"
    finally:
        s.close()


FEEDER_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "feeder_columns.json")
FEEDER = json.load(open(FEEDER_GOLDEN)) if os.path.exists(FEEDER_GOLDEN) else {}


@pytest.mark.parametrize("adv_opt", [0, 1, 2, 3, 4])
def test_device_writes_the_feeder_output_files(adv_opt):
    """PROGRAM feeder, the reference's own kernel harness (dyn_em/module_advect_em.F:10567): the device's scalar step, driven as the program
    drives its kernels, must write the files the reference's kernels wrote (tests/golden/feeder_columns.json, made by tests/feeder_ref.py)."""
    want = {int(k): v for k, v in FEEDER[str(adv_opt)].items()}
    case = parity.tophat_case(adv_opt)
    s = Solver(case)
    try:
        s.start_domain()
        d = case.dims
        for n in range(1, max(want) + 1):
            for rk in (1, 2, 3):
                s.call("rk_scalar_step", rk)
            s.bc("scalar_1", "p")
            if n in want:
                row = np.asarray(s.get("scalar_1"))[2 - d.jms, 0, 1 - d.ims:91 - d.ims + 1]
                assert [int(x) for x in np.rint((row * np.float32(100)).astype(np.float32))] == want[n], f"loop {n}"
    finally:
        s.close()
