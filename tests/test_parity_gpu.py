"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle.

Bar: BIT equality of every field both sides hold (prognostic state, diagnostics and the
i1 scratch arrays), phase by phase and over whole solve_em steps."""
import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import Solver
from wrf_b200.decomp import bdy_side as parity_bdy_side

pytestmark = pytest.mark.gpu

CASES = None
ALL = ["hill2d_x", "quarter_ss", "quarter_ss_nopd", "b_wave", "les", "periodic", "b_wave_lid", "quarter_ss_opts", "rk2_sym_x",
       "mapfac", "mapfac_open", "diffusion", "diffusion_periodic", "b_wave_shipped",
       "mono", "mono_periodic", "weno", "weno_periodic", "wenopd", "wenopd_periodic",
       "weno_mom", "weno_mom_periodic", "weno_mom_b_wave", "ieva", "ieva_wdamp", "ieva_periodic", "ieva_weno_hill",
       "diff6", "diff6_periodic", "diff6_hill", "tall", "tiny", "rk2_everything", "backward",
       "order64", "order63_periodic", "order_mixed", "smag2d", "smag2d_periodic", "smag2d_hill", "smag2d_b_wave",
       "diff2", "diff2_periodic", "diff2_hill", "diff2_b_wave",
       "smag3d", "smag3d_periodic", "smag3d_hill", "tke", "tke_periodic", "tke_hill",
       "shipped_hill2d_x", "shipped_quarter_ss", "shipped_b_wave", "shipped_les",
       "specified", "specified_real", "nested", "specified_bcs_moist", "specified_tke"]


def cases():
    global CASES
    if CASES is None:
        CASES = parity.small_cases()
    return CASES


def _pair(name):
    case = cases()[name]
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    return o, s


@pytest.mark.parametrize("name", ALL)
def test_phase_parity(name):
    """Every phase of one solve_em step, fed with the oracle's state,
    reproduces the oracle bit for bit."""
    fails = parity.run_phase_parity(cases()[name])
    assert not fails, f"{len(fails)} field mismatches, first: {fails[:5]}"


@pytest.mark.parametrize("name", ALL)
def test_full_step_bit_exact(name):
    """One whole solve_em step: every common field identical to the oracle's."""
    o, s = _pair(name)
    o.step(1); s.step(1)
    rep = parity.diff_report(o, s)
    assert not rep, rep[:5]
    # and the north-star statement of the same fact; diff_report lets NaN equal NaN (scratch arrays keep untouched
    # allocation values on both sides), so the prognostic state is also required to be finite on every domain point
    d = cases()[name].dims
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "al"]:
        a = s.get(n)
        assert parity.rel_l2_field(o, n, a, o.get(n)) <= parity.STEP_TOL_REL_L2
        assert np.isfinite(a[d.jds - d.jms:d.jde - d.jms, ..., d.ids - d.ims:d.ide - d.ims]).all(), n
    s.close()


@pytest.mark.parametrize("name", ["quarter_ss", "hill2d_x", "b_wave", "periodic"])
def test_drift_20_steps_bit_exact(name):
    """bounded drift: 20 steps stay identical to the oracle (drift = 0)."""
    o, s = _pair(name)
    o.step(20); s.step(20)
    rep = parity.diff_report(o, s)
    assert not rep, rep[:5]
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2"]:
        assert np.isfinite(s.get(n)).all()
    s.close()


@pytest.mark.parametrize("name", ["quarter_ss", "periodic"])
def test_drift_100_steps(name):
    """north star: "bounded drift reported over 100 steps".  The drift is zero: after 100 steps (20 min of the
    supercell case, Kessler rain reaching the ground) every prognostic field still equals the oracle's bit for bit."""
    o, s = _pair(name)
    o.step(100); s.step(100)
    worst = 0.0
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2"] + [m for m in s.names() if m.startswith("moist_q")]:
        a, b = s.get(n), o.get(n)
        assert np.isfinite(a).all(), n
        worst = max(worst, parity.rel_l2_field(o, n, a, b))
        assert np.array_equal(a, b), (n, worst)
    assert np.abs(s.get("w_2")).max() > 0.5          # and the flow is far from trivial by then
    s.close()


@pytest.mark.parametrize("name", ["shipped_hill2d_x", "shipped_quarter_ss", "shipped_b_wave", "shipped_les"])
def test_shipped_configurations_ten_steps_bit_exact(name):
    """BASELINE configs C1-C4 as shipped (namelist.input, input_sounding / input_jet, shipped dimensions and closures): ten
    steps on the device equal the oracle bit for bit in every field"""
    o, s = _pair(name)
    o.set_threads(8)
    o.step(10); s.step(10)
    rep = parity.diff_report(o, s)
    assert not rep, rep[:5]
    assert np.isfinite(s.get("w_2")).all() and np.abs(s.get("w_2")).max() > 0.0
    s.close()


@pytest.mark.parametrize("name", ["specified", "specified_real", "nested", "specified_bcs_moist", "specified_tke"])
def test_boundary_zone_twelve_steps_bit_exact(name):
    """f3: specified / nested lateral boundaries.  Twelve steps with a new boundary time read after the sixth (the host
    uploads fresh boundary arrays and resets dtbc, as share/input_wrf.F does): every field equals the oracle bit for bit,
    the specified rows hold exactly the boundary file's value at the end of each step, and grid%dtbc follows."""
    case = cases()[name]
    o, s = _pair(name)
    d, c = case.dims, case.config
    for n in range(12):
        if n == 6:     # next boundary time: the old values advanced by six steps of their tendencies, new tendencies
            for f in [k for k in case.fields if "_bt" in k and parity_bdy_side(k)]:
                val = f.replace("_bt", "_b")
                new = (case.fields[val] + np.float32(6 * c.dt) * case.fields[f]).astype(np.float32)
                o.set(val, new); s.set(val, new)
                o.set(f, -0.5 * case.fields[f]); s.set(f, -0.5 * case.fields[f])
            o.set_dtbc(0.0); s.set_dtbc(0.0)
        o.step(1); s.step(1)
        assert s.dtbc() == o.dtbc() and abs(s.dtbc() - ((n % 6) + 1) * c.dt) < 1e-3
    rep = parity.diff_report(o, s)
    assert not rep, rep[:5]
    # spec_bdy_final: mu_2 on the outermost row is the boundary value + dtbc * tendency
    mu, bx, bt = s.get("mu_2"), s.get("mu_bxs").reshape(c.spec_bdy_width, 1, -1), s.get("mu_btxs").reshape(c.spec_bdy_width, 1, -1)
    j0, j1 = d.jds + 1 - d.jms, d.jde - 2 - d.jms
    want = (bx[0, 0, j0:j1 + 1] + np.float32(s.dtbc()) * bt[0, 0, j0:j1 + 1]).astype(np.float32)
    assert np.array_equal(mu[j0:j1 + 1, d.ids - d.ims], want)
    assert np.isfinite(s.get("w_2")).all() and np.abs(s.get("w_2")).max() > 0.0
    s.close()


def test_alternative_kernel_forms_are_still_exact():
    """WRFGPU_MT=column selects k_advance_mu_t (operands prefetched into registers) instead of k_advance_mu_t_st (operands
    staged by the bulk-copy engine), WRFGPU_SSP=scalar the one-point-per-thread streaming kernels instead of the *_v4
    ones (four columns per thread, 16-byte accesses); the switches are read once per process, so the check runs in a
    child: phase parity and one whole step of a few cases"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WRFGPU_MT="column", WRFGPU_SSP="scalar", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_parity_gpu.py"), "-x", "-q", "-k",
                        "(phase_parity or full_step) and (quarter_ss or periodic or specified_real) and not shipped"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1000:]


@pytest.mark.parametrize("name", ["ieva", "ieva_periodic"])
def test_ieva_three_steps_bit_exact(name):
    """IEVA with a non-zero implicit part (wwI) on every step: three steps stay identical to the oracle, including
    the split mass fluxes and the mass estimates of the last substep."""
    o, s = _pair(name)
    o.step(3); s.step(3)
    assert np.abs(o.get("wwI")).max() > 0.0
    rep = parity.diff_report(o, s)
    assert not rep, rep[:5]
    s.close()


def test_w_damp_cfl_maxima_equal_the_oracle():
    """a7: max_vert_cfl / max_horiz_cfl (w_damp's reduction, consumed by the adaptive time step, solve_em.F:914-930) equal
    the oracle's bit for bit after every step, and the reset works"""
    case = cases()["ieva_wdamp"]
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    for n in range(3):
        o.step(1); s.step(1)
        a, b = s.max_cfl(reset=(n == 1)), o.max_cfl(reset=(n == 1))
        assert a == b, (n, a, b)
        if n < 2:      # the initial updraft exceeds w_crit_cfl = 2 (the damping has tamed it by the third step, after the reset)
            assert a[0] > 2.0
        assert 0.0 < a[1] < 1.0                    # horizontal Courant numbers are ordinary
    s.close()
    with pytest.raises(Exception):
        Solver(cases()["quarter_ss"]).max_cfl()


def test_solve_em_host_round_trip():
    """the strict drop-in call (host arrays in, one solve_em, host arrays out) equals
    stepping the resident state."""
    case = cases()["quarter_ss"]
    a, b = Solver(case), Solver(case)
    a.start_domain(); b.start_domain()
    prog = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "moist_qv", "moist_qc", "moist_qr"]
    host = {n: np.ascontiguousarray(a.get(n)).copy() for n in prog}
    ptrs = [host[n].ctypes.data for n in prog]
    a.step(1)
    b.solve_em(prog, ptrs, prog, ptrs)
    for n in prog:
        assert np.array_equal(host[n].ravel(), a.get(n).ravel()), n
    a.close(); b.close()


def test_shipped_hill2d_x_diffusion_settings_are_the_identity():
    """em_hill2d_x ships diff_opt = 2, km_opt = 1, khdif = kvdif = 0: every diffusion tendency is a product with a
    zero eddy viscosity, so the step must equal the diff_opt = 0 step bit for bit (and is accepted at create)."""
    from wrf_b200 import make_case
    a = Solver(make_case("hill2d_x", nx=62, ny=3, nz=21))
    b = Solver(make_case("hill2d_x", nx=62, ny=3, nz=21, diff_opt=2, km_opt=1, khdif=0.0, kvdif=0.0))
    for s in (a, b):
        s.start_domain(); s.step(2)
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "moist_qv"]:
        assert np.array_equal(a.get(n), b.get(n)), n
    a.close(); b.close()


REJECTED = [
    ("polar", dict(polar=1)),                                   # polar filters: out of scope
    ("momentum_adv_opt=2", dict(momentum_adv_opt=2)),           # no such option in the Registry
    ("km_opt=2 needs diff_opt=2", dict(diff_opt=1, km_opt=2)),  # tke_rhs is only called from the diff_opt=2 branch
    ("odd time_step_sound", dict(time_step_sound=5)),           # solve_em.F:607-610
    ("hydrostatic", dict(non_hydrostatic=0)),
    ("damp_opt=1", dict(damp_opt=1)),
    ("global lat-lon", dict(map_proj=6)),
    ("h_sca_adv_order=4", dict(h_sca_adv_order=4)),
    ("mp_physics=2", dict(mp_physics=2)),
    ("conflicting bc", dict(periodic_x=1, open_xs=1)),
    ("monotonic tke with h_sca_adv_order=6", dict(diff_opt=2, km_opt=2, tke_adv_opt=2, h_sca_adv_order=6)),   # wrf_error_fatal in the reference
    # specified / nested boundaries are built for all four sides and the non-WENO, non-monotonic, non-IEVA options
    ("specified with open sides", dict(specified=1, spec_bdy_width=5, spec_zone=1, relax_zone=4)),
    ("specified and nested", dict(specified=1, nested=1, open_xs=0, open_xe=0, open_ys=0, open_ye=0, spec_bdy_width=5, spec_zone=1, relax_zone=4)),
    ("specified with WENO", dict(specified=1, open_xs=0, open_xe=0, open_ys=0, open_ye=0, spec_bdy_width=5, spec_zone=1, relax_zone=4, moist_adv_opt=3)),
    ("specified with IEVA", dict(specified=1, open_xs=0, open_xe=0, open_ys=0, open_ye=0, spec_bdy_width=5, spec_zone=1, relax_zone=4, zadvect_implicit=1)),
    ("relax_zone outside spec_bdy_width", dict(specified=1, open_xs=0, open_xe=0, open_ys=0, open_ye=0, spec_bdy_width=4, spec_zone=1, relax_zone=4)),
]


@pytest.mark.parametrize("why,overrides", REJECTED, ids=[r[0] for r in REJECTED])
def test_unsupported_namelist_fails_loudly(why, overrides):
    """no CPU fallback and no silent approximation: a namelist value the device path does not implement is an error at
    wrfgpu_create (the Fortran shim turns it into wrf_error_fatal)"""
    from wrf_b200 import WrfGpuError, make_case
    case = make_case("quarter_ss", nx=26, ny=24, nz=21)
    for k, v in overrides.items():
        assert hasattr(case.config, k), k
        setattr(case.config, k, v)
    with pytest.raises(WrfGpuError):
        Solver(case)
