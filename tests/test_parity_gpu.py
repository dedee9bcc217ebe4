"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import Solver

pytestmark = pytest.mark.gpu

CASES = None


def cases():
    global CASES
    if CASES is None:
        CASES = parity.small_cases()
    return CASES


@pytest.mark.parametrize("name", ["periodic", "hill2d_x", "quarter_ss", "quarter_ss_nopd", "b_wave"])
def test_phase_parity(name):
    """Every phase of one solve_em step, fed with the oracle's state, reproduces the oracle:
    bit-exact where the arithmetic has no transcendental, <= 1e-6 relative L2 where it has."""
    fails = parity.run_phase_parity(cases()[name])
    assert not fails, f"{len(fails)} field mismatches, first: {fails[:5]}"


@pytest.mark.parametrize("name", ["periodic", "hill2d_x", "quarter_ss", "b_wave"])
def test_full_step_parity(name):
    """north star: prognostic fields within 1e-5 relative L2 of the reference after one step."""
    case = cases()[name]
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    o.step(1); s.step(1)
    prog = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "al"] + [n for n in s.names() if n.startswith("moist_q") or n.startswith("scalar_") and "old" not in n and "tend" not in n]
    for n in prog:
        r = parity.rel_l2_field(o, n, s.get(n), o.get(n))
        assert r <= parity.STEP_TOL_REL_L2, (n, r)


def test_drift_10_steps():
    """bounded drift: 10 steps of the moist supercell stay within 1e-3 relative L2."""
    case = cases()["quarter_ss"]
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    o.step(10); s.step(10)
    for n in ["u_2", "v_2", "t_2", "ph_2", "mu_2", "moist_qv"]:
        r = parity.rel_l2_field(o, n, s.get(n), o.get(n))
        assert np.isfinite(r) and r <= 1.0e-3, (n, r)
