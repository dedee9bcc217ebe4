"""History output: the wrfout writer on CPU, the non-blocking device snapshot on the GPU."""
import os

import numpy as np
import pytest

from tests import parity
from wrf_b200 import decomp, make_case, wrfout


def _read(path):
    from scipy.io import netcdf_file
    return netcdf_file(path, "r", mmap=False)


def test_wrfout_layout_follows_the_reference(tmp_path):
    """share/output_wrf.F + external/io_netcdf: dimension names and order, upper-cased datanames, stagger / MemoryOrder /
    FieldType attributes, 19-character Times, patch interior only, staggered variables one longer in their direction."""
    case = make_case("quarter_ss", nx=26, ny=24, nz=21)
    d = case.dims
    names = wrfout.stream_fields(list(case.fields), "h")
    assert names[:4] == ["u_2", "v_2", "w_2", "ph_2"] and "moist_qv" in names and "u_1" not in names
    assert "u_1" in wrfout.stream_fields(list(case.fields), "r")
    w = wrfout.WrfOut(str(tmp_path / "wrfout_d01_0001-01-01_00:00:00"), d, case.config)
    flds = {n: case.fields[n] for n in names}
    w.write(wrfout.time_string("0001-01-01_00:00:00", 0.0), flds)
    w.write(wrfout.time_string("0001-01-01_00:00:00", 25 * 12.0), flds)
    w.close()
    nc = _read(w.path)
    nx, ny, nz = d.ide - 1, d.jde - 1, d.kde - 1
    dims = nc.dimensions
    assert (dims["west_east"], dims["south_north"], dims["bottom_top"]) == (nx, ny, nz)
    assert (dims["west_east_stag"], dims["south_north_stag"], dims["bottom_top_stag"], dims["DateStrLen"]) == (nx + 1, ny + 1, nz + 1, 19)
    assert nc.variables["U"].dimensions == ("Time", "bottom_top", "south_north", "west_east_stag")
    assert nc.variables["V"].dimensions == ("Time", "bottom_top", "south_north_stag", "west_east")
    assert nc.variables["PH"].dimensions == ("Time", "bottom_top_stag", "south_north", "west_east")
    assert nc.variables["MU"].dimensions == ("Time", "south_north", "west_east")
    assert nc.variables["ZNW"].dimensions == ("Time", "bottom_top_stag")
    u = nc.variables["U"]
    assert u.stagger == b"X" and u.MemoryOrder == b"XYZ" and int(u.FieldType) == 104 and u.units == b"m s-1"
    assert nc.variables["QVAPOR"].description == b"Water vapor mixing ratio"
    times = [b"".join(t).decode() for t in nc.variables["Times"][:]]
    assert times == ["0001-01-01_00:00:00", "0001-01-01_00:05:00"]
    sj, si = slice(d.jds - d.jms, d.jde - d.jms), slice(d.ids - d.ims, d.ide - d.ims)
    assert np.array_equal(nc.variables["THM"][1], case.fields["t_2"][sj, :nz, si].transpose(1, 0, 2))
    assert np.array_equal(nc.variables["U"][0], case.fields["u_2"][sj, :nz, slice(si.start, si.stop + 1)].transpose(1, 0, 2))
    assert np.array_equal(nc.variables["W"][0], case.fields["w_2"][sj, :nz + 1, si].transpose(1, 0, 2))
    assert np.array_equal(nc.variables["MUB"][0], case.fields["mub"][sj, si])
    # T = theta - t0 from THM = theta_m - t0 and qv (use_theta_m = 1)
    th = (case.fields["t_2"].astype(np.float64) + 300.0) / (1.0 + 461.6 / 287.0 * case.fields["moist_qv"]) - 300.0
    assert np.abs(nc.variables["T"][0] - th[sj, :nz, si].transpose(1, 0, 2)).max() < 1e-4
    assert int(getattr(nc, "WEST-EAST_GRID_DIMENSION")) == d.ide and int(nc.MP_PHYSICS) == 1 and abs(float(nc.DT) - 12.0) < 1e-6
    nc.close()


def test_split_tiles_stitch_back_to_the_single_file(tmp_path):
    """io_form 102: every rank writes its patch; the PATCH_START/END attributes place it in the domain."""
    case = make_case("quarter_ss", nx=29, ny=27, nz=15)
    d = case.dims
    names = ["u_2", "v_2", "w_2", "t_2", "mu_2", "moist_qv"]
    one = wrfout.WrfOut(str(tmp_path / "one"), d, case.config)
    one.write("0001-01-01_00:00:00", {n: case.fields[n] for n in names})
    one.close()
    ref = _read(one.path)
    for pc in decomp.split_case(case, 3, 2):
        w = wrfout.WrfOut(str(tmp_path / "tile"), pc.dims, pc.config, split=True)
        assert w.path.endswith(f"_{pc.config.mytask:04d}")
        w.write("0001-01-01_00:00:00", {n: pc.fields[n] for n in names})
        w.close()
    got = {v: np.full(ref.variables[v].shape, np.nan, np.float32) for v in ("U", "V", "W", "THM", "MU", "QVAPOR")}
    for r in range(6):
        nc = _read(str(tmp_path / "tile") + f"_{r:04d}")
        for v, a in got.items():
            stag = nc.variables[v].stagger.decode()
            i0 = int(getattr(nc, "WEST-EAST_PATCH_START_STAG" if stag == "X" else "WEST-EAST_PATCH_START_UNSTAG"))
            i1 = int(getattr(nc, "WEST-EAST_PATCH_END_STAG" if stag == "X" else "WEST-EAST_PATCH_END_UNSTAG"))
            j0 = int(getattr(nc, "SOUTH-NORTH_PATCH_START_STAG" if stag == "Y" else "SOUTH-NORTH_PATCH_START_UNSTAG"))
            j1 = int(getattr(nc, "SOUTH-NORTH_PATCH_END_STAG" if stag == "Y" else "SOUTH-NORTH_PATCH_END_UNSTAG"))
            a[..., j0 - 1:j1, i0 - 1:i1] = nc.variables[v][:]
        nc.close()
    for v, a in got.items():
        assert np.array_equal(a, ref.variables[v][:]), v      # every point written exactly once with the right value
    ref.close()


@pytest.mark.gpu
def test_snapshot_does_not_stall_and_is_consistent():
    """wrfgpu_snapshot_begin returns before the data are on the host; stepping on meanwhile does not disturb the
    snapshot: it holds the state of the moment of the alarm, bit for bit."""
    from wrf_b200 import Solver
    case = parity.small_cases()["quarter_ss"]
    s = Solver(case)
    s.start_domain(); s.step(2)
    names = wrfout.stream_fields(s.names(), "h")
    want = {n: s.get(n).copy() for n in names}
    s.snapshot_begin(names)
    s.step(3)                                  # the time loop goes on while the copies are in flight
    snap = s.snapshot_wait()
    assert not s.snapshot_busy()
    for n in names:
        assert np.array_equal(snap[n], want[n]), n
    assert not np.array_equal(s.get("t_2"), want["t_2"])     # and the model did move on
    # a second snapshot reuses the buffers
    s.snapshot_begin(["t_2", "mu_2"])
    snap2 = s.snapshot_wait()
    assert set(snap2) == {"t_2", "mu_2"} and np.array_equal(snap2["t_2"], s.get("t_2"))
    s.close()


@pytest.mark.gpu
def test_history_stream_writes_wrfout_from_the_device(tmp_path):
    """med_before_solve_io's history alarm without host arrays: 10 steps, history every 5 -> 3 records (0, 5, 10),
    each equal to what a blocking download at that step returns."""
    from wrf_b200 import Solver
    case = parity.small_cases()["quarter_ss"]
    a, b = Solver(case), Solver(case)
    a.start_domain(); b.start_domain()
    h = wrfout.HistoryStream(a, str(tmp_path / "wrfout_d01_0001-01-01_00:00:00"), interval=5)
    h.after_step(0)
    ref = {0: b.get("t_2").copy()}
    for it in range(1, 11):
        a.step(1); h.after_step(it)
        b.step(1)
        if it % 5 == 0:
            ref[it] = b.get("t_2").copy()
    h.close()
    nc = _read(h.out.path)
    d = case.dims
    sj, si, nz = slice(d.jds - d.jms, d.jde - d.jms), slice(d.ids - d.ims, d.ide - d.ims), d.kde - 1
    assert nc.variables["THM"].shape[0] == 3
    assert [b"".join(t).decode() for t in nc.variables["Times"][:]] == ["0001-01-01_00:00:00", "0001-01-01_00:01:00", "0001-01-01_00:02:00"]
    for r, it in enumerate((0, 5, 10)):
        assert np.array_equal(nc.variables["THM"][r], ref[it][sj, :nz, si].transpose(1, 0, 2)), it
    assert "QRAIN" in nc.variables and "RAINNC" in nc.variables
    nc.close()
    a.close(); b.close()


# ---------------------------------------------------------------------------------------------------------------
# restart: the reference's regression suite asks a restarted run to reproduce the continuous one bit for bit
# (tools/regtest.csh, "restart" comparisons with BIT4BIT).  Same criterion here, through a wrfrst file.
PROG = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "u_1", "v_1", "w_1", "t_1", "ph_1", "mu_1", "p", "al"]


def _restart_round_trip(make, case, tmp_path, n1=3, n2=3):
    import dataclasses
    cont = make(case)
    cont.start_domain(); cont.step(n1 + n2)
    first = make(case)
    first.start_domain(); first.step(n1)
    names = wrfout.stream_fields(first.names(), "r")
    w = wrfout.WrfOut(str(tmp_path / "wrfrst_d01"), case.dims, case.config)
    w.write(wrfout.time_string("0001-01-01_00:00:00", n1 * case.config.dt), {n: first.get(n) for n in names})
    w.close()
    fields = wrfout.read_fields(w.path, case.dims)
    second = make(dataclasses.replace(case, fields=fields))
    second.start_domain(); second.step(n2)
    prog = PROG + [n for n in cont.names() if n.startswith(("moist_q", "scalar_")) and "old" not in n and "tend" not in n]
    prog += [n for n in ("rainnc", "h_diabatic") if n in cont.names()]
    bad = [n for n in prog if not np.array_equal(_owned(case, cont.get(n)), _owned(case, second.get(n)))]
    for m in (cont, first, second):
        if hasattr(m, "close"):
            m.close()
    return bad


def _owned(case, a):
    d = case.dims
    if a.ndim == 3:
        return a[d.jds - d.jms:d.jde - d.jms + 1, :, d.ids - d.ims:d.ide - d.ims + 1]
    if a.ndim == 2:
        return a[d.jds - d.jms:d.jde - d.jms + 1, d.ids - d.ims:d.ide - d.ims + 1]
    return a


@pytest.mark.parametrize("name", ["quarter_ss", "periodic", "b_wave", "hill2d_x", "diffusion", "ieva_periodic"])
def test_restart_reproduces_the_continuous_run_oracle(name, tmp_path):
    from tests.oracle_lib import Oracle
    bad = _restart_round_trip(Oracle, parity.small_cases()[name], tmp_path)
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quarter_ss", "periodic", "ieva_periodic"])
def test_restart_reproduces_the_continuous_run_gpu(name, tmp_path):
    from wrf_b200 import Solver
    bad = _restart_round_trip(Solver, parity.small_cases()[name], tmp_path)
    assert not bad, bad
