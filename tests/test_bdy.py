"""Specified / nested lateral boundaries on the CPU: the host helpers that build the boundary arrays
(wrf_b200/bdy.py) and the oracle's restatement of share/module_bc.F:1122-2456 / dyn_em/module_bc_em.F:17-700, pinned by
the properties the reference documents for them."""
import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import bdy, make_case


def _case(**kw):
    return make_case("quarter_ss", nx=30, ny=28, nz=19, specified=1, **parity.SPEC, **kw)


def test_lbc_fcx_gcx_is_the_start_em_ramp():
    """module_bc_em.F:1297-1341: a linear ramp from 0.1/dt (fcx) and 1/(50 dt) (gcx) at the first relaxation row to zero at
    the last, times exp(-(row) spec_exp) when specified; nothing in the specified rows"""
    c = _case().config
    fcx, gcx = bdy.lbc_fcx_gcx(c, spec_exp=0.0)
    assert fcx[0] == 0 and gcx[0] == 0
    np.testing.assert_allclose(fcx[1:], [0.1 / c.dt, 0.1 / c.dt * 2 / 3, 0.1 / c.dt / 3, 0.0], rtol=1e-6)
    np.testing.assert_allclose(gcx[1:], [1 / c.dt / 50, 1 / c.dt / 50 * 2 / 3, 1 / c.dt / 50 / 3, 0.0], rtol=1e-6)
    f2, _ = bdy.lbc_fcx_gcx(c, spec_exp=0.33)
    np.testing.assert_allclose(f2[1:4], fcx[1:4] * np.exp(-0.33 * np.arange(3)), rtol=1e-6)


def test_stuff_bdy_layout():
    """share/module_bc.F:2694-2891: row b of the x-start array is column ids+b-1, of the x-end array ibe-b+1 (ibe = ide for
    u); the y arrays likewise; Fortran (along, k, b) = C (b, k, along)"""
    case = _case()
    d = case.dims
    u = case.fields["u_2"]
    a = bdy.stuff_bdy(d, u, "U", 5)
    j, k = d.jds + 3 - d.jms, 2
    assert a["xs"][1, k, j] == u[j, k, d.ids + 1 - d.ims]
    assert a["xe"][0, k, j] == u[j, k, d.ide - d.ims]
    i = d.ids + 4 - d.ims
    assert a["ys"][2, k, i] == u[d.jds + 2 - d.jms, k, i]
    assert a["ye"][1, k, i] == u[d.jde - 2 - d.jms, k, i]


def test_a_state_that_equals_its_boundary_file_is_not_relaxed():
    """relax_bdytend: every term is (boundary value - field); with the boundary arrays taken from the state itself and no
    tendencies the relaxation adds exactly nothing -- up to the rounding of the mass coupling the boundary file stores"""
    case = bdy.add_boundary_arrays(_case())
    o = Oracle(case)
    o.start_domain()
    for ph, args in parity.phase_sequence(case)[:5]:     # up to and including rk_tendency(1)
        o.call(ph, *args)
    d = case.dims
    for n, scale in (("u_save", 1.0), ("v_save", 1.0), ("t_save", 1.0), ("ph_save", 1.0)):
        a = np.abs(o.get(n)[d.jds - d.jms:d.jde - d.jms, :-1, d.ids - d.ims:d.ide - d.ims])
        assert a.max() < 2e-3 * scale, (n, a.max())     # coupled fields are O(1e4..1e6): one ulp of them times fcx


def test_relaxation_pulls_towards_the_boundary_file_only_in_the_relaxation_rows():
    """a warm anomaly over the whole domain against an unchanged boundary file: t_save (the relaxation tendency) is negative
    in rows 1..3 from each side, strongest in row 1, and zero in the specified row and in the interior"""
    case = bdy.add_boundary_arrays(_case())
    case.fields["t_2"] = (case.fields["t_2"] + 2.0).astype(np.float32)
    case.fields["t_1"] = case.fields["t_2"].copy()
    o = Oracle(case)
    o.start_domain()
    for ph, args in parity.phase_sequence(case)[:5]:
        o.call(ph, *args)
    d = case.dims
    ts = o.get("t_save")
    jm, k = (d.jds + d.jde) // 2 - d.jms, 3
    row = ts[jm, k, d.ids - d.ims:d.ids - d.ims + 8]
    assert row[0] == 0 and (row[1:4] < 0).all() and row[1] < row[2] < row[3] and (row[4:] == 0).all(), row
    col = ts[d.jde - 1 - d.jms - np.arange(8), k, (d.ids + d.ide) // 2 - d.ims]
    assert col[0] == 0 and (col[1:4] < 0).all() and (col[4:] == 0).all(), col


@pytest.mark.parametrize("name", ["specified", "nested"])
def test_specified_rows_follow_the_boundary_file(name):
    """spec_bdy_final: at the end of every step the specified rows hold value + dtbc * tendency of the boundary file,
    uncoupled with the new column mass; dtbc grows by dt per step (solve_em.F:377)"""
    case = parity.small_cases()[name]
    d, c = case.dims, case.config
    o = Oracle(case)
    o.start_domain()
    o.step(3)
    assert abs(o.dtbc() - 3 * c.dt) < 1e-4
    f32 = np.float32
    j = slice(d.jds + 1 - d.jms, d.jde - 1 - d.jms)
    mu = o.get("mu_2")
    want = case.fields["mu_bxe"][0, 0, j] + f32(o.dtbc()) * case.fields["mu_btxe"][0, 0, j]
    assert np.array_equal(mu[j, d.ide - 1 - d.ims], want.astype(f32))
    t, muts, c1h, c2h = o.get("t_2"), o.get("muts"), o.get("c1h"), o.get("c2h")
    i = slice(d.ids - d.ims, d.ide - d.ims)
    for k in (0, 5):
        bt = case.fields["t_bys"][0, k, i] + f32(o.dtbc()) * case.fields["t_btys"][0, k, i]
        assert np.array_equal(t[d.jds - d.jms, k, i], (bt / (c1h[k] * muts[d.jds - d.jms, i] + c2h[k])).astype(f32))


def test_inflow_species_without_boundary_arrays_are_zeroed():
    """flow_dep_bdy: qc, qr of a specified run without have_bcs_moist are zero on inflow and take the value one row in on
    outflow; qv has boundary arrays and is left to spec_bdy_final"""
    case = parity.small_cases()["specified"]
    d = case.dims
    o = Oracle(case)
    qc = o.get("moist_qc"); qc[...] = 1.0e-4
    o.set("moist_qc", qc); o.set("moist_old_qc", qc)
    o.start_domain()
    o.step(1)
    qc, ru = o.get("moist_qc"), o.get("ru_m")
    j = slice(d.jds + 1 - d.jms, d.jde - 2 - d.jms + 1)
    west = qc[j, 3, d.ids - d.ims]
    inflow = ru[j, 3, d.ids - d.ims] >= 0
    assert inflow.any() and (west[inflow] == 0).all()
    assert np.array_equal(west[~inflow], qc[j, 3, d.ids + 1 - d.ims][~inflow])
