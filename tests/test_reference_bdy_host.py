"""wrf_b200/bdy.py (the host-side helpers of the specified / nested boundary zone) against the reference's routines executed from source:
lbc_fcx_gcx (dyn_em/module_bc_em.F:1297-1341, the relaxation weights start_em computes) and stuff_bdy_new (share/module_bc.F:2694-2891, how
real.exe / med_force_domain lay a field out in the boundary arrays), for every stagger.  Needs /root/reference."""
import numpy as np
import pytest

from tests import refrun
from tests.refcases import build
from wrf_b200 import bdy

pytestmark = pytest.mark.skipif(not refrun.have_reference(), reason="needs the reference tree (/root/reference)")
f32 = np.float32


@pytest.mark.parametrize("kind", ["specified", "nested"])
@pytest.mark.parametrize("width,sz,rz,spec_exp,dt", [(5, 1, 4, 0.0, 12.0), (5, 1, 4, 0.33, 12.0), (10, 1, 9, 0.33, 40.0), (8, 2, 6, 0.0, 6.0)])
def test_lbc_fcx_gcx(kind, width, sz, rz, spec_exp, dt):
    R = refrun.library(1).routine("lbc_fcx_gcx")
    fcx, gcx = np.zeros(width, f32), np.zeros(width, f32)
    R(fcx=fcx, gcx=gcx, spec_bdy_width=width, spec_zone=sz, relax_zone=rz, dt=f32(dt), spec_exp=f32(spec_exp), specified=kind == "specified",
      nested=kind == "nested")

    class C:
        pass
    c = C()
    c.spec_bdy_width, c.spec_zone, c.relax_zone, c.dt, c.specified, c.nested = width, sz, rz, dt, kind == "specified", kind == "nested"
    mf, mg = bdy.lbc_fcx_gcx(c, spec_exp)
    assert np.array_equal(fcx.view(np.uint32), mf.view(np.uint32)) and np.array_equal(gcx.view(np.uint32), mg.view(np.uint32))
    assert fcx[sz] > 0 and fcx[:sz].max() == 0


@pytest.mark.parametrize("stagger", ["U", "V", "W", "M", "T"])
def test_stuff_bdy(stagger):
    case = build("specified")
    d, W = case.dims, case.config.spec_bdy_width
    R = refrun.library(1).routine("stuff_bdy_new")
    ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
    rng = np.random.default_rng(7)
    if stagger == "M":                          # 2-D fields go through the same routine with one level (kds = kde = 1)
        a = rng.standard_normal((nj, ni)).astype(f32)
        data, kde = a.T[:, None, :].copy(), 1
        kms, kme = 1, 1
    else:
        a = rng.standard_normal((nj, nk, ni)).astype(f32)
        data, kde = np.ascontiguousarray(a.transpose(2, 1, 0)), d.kde
        kms, kme = d.kms, d.kme
    xs, xe = np.zeros((nj, kde, W), f32), np.zeros((nj, kde, W), f32)
    ys, ye = np.zeros((ni, kde, W), f32), np.zeros((ni, kde, W), f32)
    # (the translator folds the whole source to lower case, character literals included: 'W' in the routine reads 'w')
    R(data3d=data, space_bdy_xs=xs, space_bdy_xe=xe, space_bdy_ys=ys, space_bdy_ye=ye, char_stagger=stagger.lower(), spec_bdy_width=W, ids=d.ids, ide=d.ide,
      jds=d.jds, jde=d.jde, kds=1, kde=kde, ims=d.ims, ime=d.ime, jms=d.jms, jme=d.jme, kms=kms, kme=kme, its=d.ids, ite=d.ide, jts=d.jds,
      jte=d.jde, kts=1, kte=kde)
    mine = bdy.stuff_bdy(d, a, stagger, W)
    nk_cmp = 1 if stagger == "M" else (d.kde if stagger == "W" else d.kde)      # the boundary arrays carry kds:kde levels for every stagger
    for side, ref in (("xs", xs), ("xe", xe), ("ys", ys), ("ye", ye)):
        got = mine[side]                                                          # (W, levels, along)
        assert got.shape == (W, nk_cmp, ref.shape[0])
        assert np.array_equal(np.ascontiguousarray(ref.transpose(2, 1, 0)).view(np.uint32), got.view(np.uint32)), (stagger, side)
    assert np.count_nonzero(xs) > 0 and np.count_nonzero(ye) > 0
