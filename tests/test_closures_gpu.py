"""GPU parity of the 3-D Smagorinsky closure (km_opt = 3: calculate_N2 + smag_km in k_smag3d, wrf_b200/csrc/deform.cuh) and of
the TKE closure (km_opt = 2: k_tke_km, k_tke_rhs, k_bound_tke, the doing_tke flag of k_diff2_s) over several steps and
over N ranks; the phase-by-phase and one-step cases are in test_parity_gpu.py (ALL).

The kernel bodies are also pinned bit for bit against the oracle on the CPU (tests/test_deform_host.py runs the same
__host__ __device__ point functions, including the restated glibc powf / expf the device uses)."""
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import Solver

pytestmark = pytest.mark.gpu
NAMES = ["smag3d", "smag3d_periodic", "smag3d_hill", "tke", "tke_periodic", "tke_hill"]


@pytest.mark.parametrize("name", NAMES)
def test_five_steps_bit_exact(name):
    case = parity.small_cases()[name]
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    o.step(5); s.step(5)
    rep = parity.diff_report(o, s)
    assert not rep, rep[:10]


@pytest.mark.parametrize("mesh", [(1, 2), (2, 2)])
@pytest.mark.parametrize("name", ["smag3d_periodic", "tke", "tke_periodic"])
def test_n_rank_gpu_equals_single_rank(name, mesh):
    """the reference's BIT4BIT criterion on the device for the two closures (the tke species travels with the scalar halos)"""
    from tests.test_multi_gpu import _run
    bad, sent = _run(parity.small_cases()[name], *mesh, nsteps=2)
    assert not bad, bad[:8]
    assert all(b > 0 for b in sent)
