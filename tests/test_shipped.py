"""The reference-held inputs of the path: the shipped test/em_*/namelist.input, input_sounding and input_jet (copied unchanged
to tests/fixtures/) are read the way ideal.exe / wrf.exe read them (wrf_b200/shipped.py) and give the initial
state the four BASELINE configurations C1-C4 start from, at the shipped dimensions and with the shipped settings.

What is pinned here on the CPU: the readers against values one can read off the files, the initial state against the
properties module_initialize_ideal.F gives it by construction, and the oracle on these states against the documented
behaviour of the cases (doc/README.test_cases).  The GPU == oracle comparison on the same states is in test_parity_gpu.py."""
import os

import numpy as np
import pytest

from tests.oracle_lib import Oracle
from wrf_b200 import shipped

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")
CASES = {"em_hill2d_x": (202, 3, 41), "em_quarter_ss": (42, 42, 41), "em_b_wave": (41, 81, 65), "em_les": (40, 40, 40)}
_CACHE = {}


def case(name):
    if name not in _CACHE:
        _CACHE[name] = shipped.init_domain(os.path.join(FIX, name))
    return _CACHE[name]


def test_fixtures_are_the_reference_files():
    """byte-identical copies of the shipped files (checked whenever the reference tree is present, i.e. in the build container)"""
    ref = "/root/reference/test"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this box")
    for name in CASES:
        for fn in os.listdir(os.path.join(FIX, name)):
            assert open(os.path.join(FIX, name, fn), "rb").read() == open(os.path.join(ref, name, fn), "rb").read(), (name, fn)


def test_namelist_reader_and_registry_defaults():
    nl = shipped.read_namelist(os.path.join(FIX, "em_quarter_ss", "namelist.input"))
    assert nl["domains"]["e_we"] == [42, 43, 43] and nl["domains"]["time_step"] == [12]
    assert nl["dynamics"]["mix_full_fields"] == [True, True, True] and nl["bdy_control"]["open_xs"] == [True]
    c, g = shipped.config_from_namelist(nl)
    assert (g["e_we"], g["e_sn"], g["e_vert"], g["ztop"], g["ideal_case"], g["hybrid_opt"]) == (42, 42, 41, 20000.0, 2, 0)
    assert (c.diff_opt, c.km_opt, c.damp_opt, c.mp_physics, c.num_moist, c.time_step_sound, c.moist_adv_opt, c.tke_adv_opt) == (2, 2, 2, 1, 3, 6, 1, 1)
    assert abs(c.dampcoef - 0.003) < 1e-9 and c.zdamp == 5000.0 and c.khdif == 500.0 and c.dt == 12.0
    assert (c.open_xs, c.open_xe, c.open_ys, c.open_ye, c.periodic_x) == (1, 1, 1, 1, 0)
    assert c.isfflx == 1 and c.use_theta_m == 1 and c.hypsometric_opt == 2 and abs(c.c_k - 0.15) < 1e-7      # Registry defaults
    c, g = shipped.config_from_namelist(shipped.read_namelist(os.path.join(FIX, "em_les", "namelist.input")))
    assert (c.isfflx, c.pert_coriolis, c.mix_isotropic, c.periodic_x, c.periodic_y, g["ideal_case"]) == (2, 1, 1, 1, 1, 9)
    assert abs(c.c_k - 0.10) < 1e-7 and abs(c.c_s - 0.18) < 1e-7 and abs(c.tke_heat_flux - 0.24) < 1e-7 and c.dt == 1.0
    c, g = shipped.config_from_namelist(shipped.read_namelist(os.path.join(FIX, "em_b_wave", "namelist.input")))
    assert (c.periodic_x, c.symmetric_ys, c.symmetric_ye, c.time_step_sound, c.diff_opt, c.km_opt, g["ideal_case"]) == (1, 1, 1, 4, 1, 1, 7)
    c, g = shipped.config_from_namelist(shipped.read_namelist(os.path.join(FIX, "em_hill2d_x", "namelist.input")))
    assert (c.open_xs, c.periodic_y, c.damp_opt, g["hybrid_opt"], g["ztop"], c.dt) == (1, 1, 3, 2, 30000.0, 20.0)


def test_sounding_and_jet_readers():
    s = shipped.read_sounding(os.path.join(FIX, "em_quarter_ss", "input_sounding"))
    assert (s["ps"], s["ts"], s["qvs"]) == (1000.0, 300.0, 14.0) and s["h"][0] == 10.0 and abs(s["u"][0] + 8.0998) < 1e-4
    assert np.all(np.diff(s["h"]) > 0) and len(s["h"]) > 100
    s = shipped.read_sounding(os.path.join(FIX, "em_hill2d_x", "input_sounding"))
    assert s["ts"] == 288.0 and s["h"][0] == 0.0 and np.all(s["u"] == 10.0) and np.all(s["qv"] == 0.0)
    jet = shipped.read_input_jet(os.path.join(FIX, "em_b_wave", "input_jet"))
    assert jet["u"].shape == (64, 80) and jet["z"][0, 0] == 125.0 and jet["z"][63, 0] == 125.0 + 250.0 * 63
    k, j = np.unravel_index(np.argmax(jet["u"]), jet["u"].shape)
    assert 55.0 < jet["u"].max() < 75.0 and 30 <= j <= 50 and 8000.0 < jet["z"][k, 0] < 12000.0      # a mid-latitude jet at the tropopause
    assert 250.0 < jet["th"][0].min() and np.all(np.diff(jet["th"], axis=0) > 0)                      # statically stable
    d = shipped.get_sounding(s, dry=True)
    assert abs(d["p"][0] - 1.0e5) < 1.0 and np.all(np.diff(d["p"]) < 0) and np.allclose(d["p"], d["pd"])


@pytest.mark.parametrize("name", list(CASES))
def test_initial_state_has_the_shipped_dimensions_and_is_balanced(name):
    c = case(name)
    d = c.dims
    assert (d.ide, d.jde, d.kde) == CASES[name]
    f = c.fields
    own = (slice(d.jps - d.jms, d.jpe - d.jms), slice(None), slice(d.ips - d.ims, d.ipe - d.ims))
    for n in ("u_2", "v_2", "t_2", "ph_2", "p", "al", "alt", "pb", "phb", "alb"):
        assert np.isfinite(f[n]).all(), n
    assert (f["alt"][own][:, :-1] > 0).all() and (f["pb"][own][:, :-1] > 0).all()
    # the state ideal.exe writes is in hydrostatic balance with respect to the model's own equations: a step leaves w small.
    # (Below the top quarter of the levels: ideal.exe integrates ph linearly in eta whatever hypsometric_opt says, start_em
    # then diagnoses al with the LOG form of the default hypsometric_opt = 2, and over the thick top layers of em_hill2d_x
    # (p_top = 66 Pa) the two differ by several per cent -- wrf.exe starts from that mismatch too; damp_opt = 3 absorbs it.)
    o = Oracle(c)
    o.start_domain(); o.step(1)
    w = np.abs(o.get("w_2")[own][:, :3 * d.kde // 4]).max()
    bound = {"em_hill2d_x": 0.5, "em_quarter_ss": 1.0, "em_b_wave": 0.05, "em_les": 0.05}[name]
    assert w < bound, w


def test_hill2d_x_has_no_y_variation_and_zero_v():
    """doc/README.test_cases:43-47 (2-D cases): 'v velocity should be zero and there should be no variation in y'"""
    c = case("em_hill2d_x")
    o = Oracle(c)
    o.start_domain(); o.step(5)
    d = c.dims
    j0 = d.jps - d.jms
    for n in ("u_2", "w_2", "t_2", "ph_2"):
        a = o.get(n)
        assert np.array_equal(a[j0], a[j0 + 1]), n
    assert np.abs(o.get("v_2")[j0:j0 + 2]).max() == 0.0
    assert abs(float(c.fields["ht"].max()) - 100.0) < 1e-3                         # the 100 m bell at icm = ide/2
    assert int(np.argmax(c.fields["ht"][j0])) + d.ims == d.ide // 2


def test_b_wave_jet_is_in_the_initial_state_and_the_channel_is_zonally_uniform_but_for_the_bubble():
    c = case("em_b_wave")
    d = c.dims
    u = c.fields["u_2"][d.jps - d.jms:d.jpe - d.jms, :-1, d.ips - d.ims:d.ipe - d.ims + 1]
    assert 55.0 < u.max() < 75.0
    assert np.array_equal(u[..., 0], u[..., -1]) and np.array_equal(u[..., 0], u[..., 7])          # no x variation in the jet
    t = c.fields["t_2"][d.jps - d.jms:d.jpe - d.jms, :-1, d.ips - d.ims:d.ipe - d.ims]
    pert = t - t[..., :1]
    assert 0.5 < np.abs(pert).max() <= 2.0001                                                       # the +-1 K wave-number-1 bubble
    assert float(c.fields["f"].max()) == np.float32(1.0e-4)


def test_quarter_ss_bubble_and_shear_from_the_shipped_sounding():
    c = case("em_quarter_ss")
    d = c.dims
    o = Oracle(c)
    qv = c.fields["moist_qv"]
    assert abs(float(qv.max()) - 0.014) < 1e-6
    dry = shipped.get_sounding(shipped.read_sounding(os.path.join(FIX, "em_quarter_ss", "input_sounding")), dry=True)
    assert abs(c.config.p_top - float(np.interp(20000.0, dry["zk"], dry["p"]))) < 1.0
    # the 3 K bubble of radius 10 km x 1.5 km at the domain centre, z = 1.5 km
    t = c.fields["t_2"]
    jc, ic = (d.jde - 1) // 2 - d.jms, (d.ide - 1) // 2 - d.ims
    col = t[jc, :-1, ic] - t[d.jps - d.jms, :-1, d.ips - d.ims]
    assert 2.5 < col.max() < 3.6 and np.argmax(col) in range(2, 8)
    o.start_domain(); o.step(10)
    assert np.abs(o.get("w_2")).max() > 1.0                                                          # the bubble rises


def test_compare_wrfout_on_two_of_our_own_files(tmp_path):
    """tools/compare_wrfout.py (the check against a real wrf.exe wrfout, for the day a Fortran box exists) on two files
    written by tools/run_shipped_case.py: identical runs compare equal; a run one step further exceeds the 1e-5 bar"""
    import io
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(FIX), "..", "tools"))
    import compare_wrfout
    import run_shipped_case
    a, b, c = (str(tmp_path / n) for n in ("wrfout_a", "wrfout_b", "wrfout_c"))
    run_shipped_case.run(os.path.join(FIX, "em_hill2d_x"), 2, a, backend="oracle")
    run_shipped_case.run(os.path.join(FIX, "em_hill2d_x"), 2, b, backend="oracle")
    run_shipped_case.run(os.path.join(FIX, "em_hill2d_x"), 3, c, backend="oracle")
    buf = io.StringIO()
    res = compare_wrfout.compare(a, b, out=buf)
    assert {"U", "W", "T", "PH", "MU", "P"} <= set(res) and all(r == 0.0 for r, _ in res.values()), buf.getvalue()
    res = compare_wrfout.compare(a, c, out=buf)
    assert res["W"][0] > 1.0e-5 and res["U"][0] > 0.0
    # the file has the reference's dimension names and staggering
    from scipy.io import netcdf_file
    nc = netcdf_file(a, "r", mmap=False)
    assert nc.variables["U"].shape == (2, 40, 2, 202) and nc.variables["W"].shape == (2, 41, 2, 201)
    assert nc.variables["U"].dimensions == ("Time", "bottom_top", "south_north", "west_east_stag")
    nc.close()
