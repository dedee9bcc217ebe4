"""Host side of the specified / nested lateral boundary zone.

In wrf.exe these pieces run on the host and stay there: start_em computes the relaxation weights (lbc_fcx_gcx,
dyn_em/module_bc_em.F:1297-1341), real.exe / med_force_domain fill the boundary arrays grid%u_bxs ... grid%mu_btye
(stuff_bdy_new / stuff_bdytend_new, share/module_bc.F:2694-3103, on mass-coupled fields: couple, dyn_em/module_big_step_
utilities_em.F) and share/input_wrf.F re-reads them at every boundary time, resetting grid%dtbc.  The device only
consumes the arrays.  What is here is what the tests, the shim-less Python driver and a user who builds an idealised
specified-boundary case need to produce those inputs in the reference's layout:

    x sides  <f>_bxs, <f>_bxe (and the tendencies <f>_btxs, <f>_btxe):  (spec_bdy_width, kde-kds+1, jms:jme)  in C order
    y sides  <f>_bys, <f>_bye (<f>_btys, <f>_btye):                     (spec_bdy_width, kde-kds+1, ims:ime)
    mu_b*:   one level

(the transposes of Fortran's (jms:jme, kds:kde, spec_bdy_width)), row b = 1 the outermost.
"""
from __future__ import annotations

import numpy as np

from .ideal import expf

DRY = ("u", "v", "w", "ph", "t", "mu")
SIDES = ("xs", "xe", "ys", "ye")


def lbc_fcx_gcx(c, spec_exp=0.0):
    """fcx, gcx of module_bc_em.F:1297-1341 in real*4, for the `specified` (exponential sponge weight) or `nested` form"""
    W, sz, rz = c.spec_bdy_width, c.spec_zone, c.relax_zone
    f32 = np.float32
    fcx, gcx = np.zeros(W, f32), np.zeros(W, f32)
    if not (c.specified or c.nested):
        return fcx, gcx
    dt = f32(c.dt)
    for loop in range(sz + 1, min(sz + rz, W) + 1):
        fc = f32(f32(f32(0.1) / dt) * f32(sz + rz - loop)) / f32(rz - 1)
        gc = f32(f32(f32(f32(1.0) / dt) / f32(50.0)) * f32(sz + rz - loop)) / f32(rz - 1)
        if c.specified:
            w = expf(f32(f32(-(loop - (sz + 1))) * f32(spec_exp)))      # EXP of gfortran's real*4 = the host libm's expf
            fc, gc = f32(fc * w), f32(gc * w)
        fcx[loop - 1], gcx[loop - 1] = fc, gc
    return fcx, gcx


def stuff_bdy(d, a, stagger, W):
    """The four boundary arrays of one field (share/module_bc.F:2694-2891): a is (nj, nk, ni) or (nj, ni) in memory
    extents; returns {"xs","xe","ys","ye"} in the layout of the module docstring, zero where the reference leaves
    the array untouched"""
    a3 = a[:, None, :] if a.ndim == 2 else a
    nj, nkm, ni = a3.shape
    nk = 1 if a.ndim == 2 else d.kde - d.kds + 1
    ibe = d.ide if stagger == "U" else d.ide - 1
    jbe = d.jde if stagger == "V" else d.jde - 1
    out = {s: np.zeros((W, nk, nj if s[0] == "x" else ni), np.float32) for s in SIDES}
    j0, j1 = d.jds - d.jms, jbe - d.jms          # rows of the field that exist on this stagger
    i0, i1 = d.ids - d.ims, ibe - d.ims
    nc = nk if (a.ndim == 2 or stagger == "W") else nk - 1      # half-level fields: k = kds .. kde-1, the top row of the arrays stays untouched
    for b in range(1, W + 1):
        out["xs"][b - 1, :nc, j0:j1 + 1] = a3[j0:j1 + 1, :nc, d.ids + b - 1 - d.ims].T
        out["xe"][b - 1, :nc, j0:j1 + 1] = a3[j0:j1 + 1, :nc, ibe - (b - 1) - d.ims].T
        out["ys"][b - 1, :nc, i0:i1 + 1] = a3[d.jds + b - 1 - d.jms, :nc, i0:i1 + 1]
        out["ye"][b - 1, :nc, i0:i1 + 1] = a3[jbe - (b - 1) - d.jms, :nc, i0:i1 + 1]
    return out


def coupled(case, name):
    """the mass-coupled form the boundary file holds: u, v, w: mu/msf, other 3-D fields: mu, mu itself uncoupled
    (the comment of spec_bdy_final, share/module_bc.F:2081-2083); float64 in, float32 out"""
    f, d = case.fields, case.dims
    mut = (f["mub"] + f["mu_2"]).astype(np.float64)
    one = np.ones_like(mut)
    c1h, c2h, c1f, c2f = (f[n].astype(np.float64)[None, :, None] for n in ("c1h", "c2h", "c1f", "c2f"))
    if name == "mu":
        return f["mu_2"].astype(np.float32), "M2"
    if name == "u":
        muu = mut.copy()
        muu[:, 1:] = 0.5 * (mut[:, 1:] + mut[:, :-1])
        return ((c1h * muu[:, None, :] + c2h) * f["u_2"] / f.get("msfuy", one)[:, None, :]).astype(np.float32), "U"
    if name == "v":
        muv = mut.copy()
        muv[1:, :] = 0.5 * (mut[1:, :] + mut[:-1, :])
        return ((c1h * muv[:, None, :] + c2h) * f["v_2"] / f.get("msfvx", one)[:, None, :]).astype(np.float32), "V"
    if name == "w":
        return ((c1f * mut[:, None, :] + c2f) * f["w_2"] / f.get("msfty", one)[:, None, :]).astype(np.float32), "W"
    if name == "ph":
        return ((c1f * mut[:, None, :] + c2f) * f["ph_2"]).astype(np.float32), "W"
    src = "t_2" if name == "t" else name
    q = f[src] if src in f else np.zeros_like(f["t_2"])   # a species the case leaves at zero
    return ((c1h * mut[:, None, :] + c2h) * q).astype(np.float32), "T"


def bdy_species(c):
    """the species that carry boundary arrays (solve_em.F:2261-2263, :2808-2810): qv always, every moist species when
    nested or have_bcs_moist, the scalars when nested or have_bcs_scalar"""
    moist = ["moist_" + q for q in ("qv", "qc", "qr")[:c.num_moist]]
    out = [m for n, m in enumerate(moist) if n == 0 or c.nested or c.have_bcs_moist]
    if c.nested or c.have_bcs_scalar:
        out += ["scalar_%d" % s for s in range(1, c.num_scalar + 1)]
    return out


def add_boundary_arrays(case, tendency=None, spec_exp=0.0):
    """Fill case.fields with fcx, gcx and the boundary arrays of every field that has them, taken from the case's own
    initial state as real.exe does for the first boundary time.  `tendency(name, values) -> array` gives the boundary
    tendencies (per second, same shape); zero (constant_bc) if None."""
    c, d = case.config, case.dims
    case.fields["fcx"], case.fields["gcx"] = lbc_fcx_gcx(c, spec_exp)
    for name in list(DRY) + bdy_species(c):
        a, stag = coupled(case, name)
        val = stuff_bdy(d, a, stag[0], c.spec_bdy_width)
        for s in SIDES:
            case.fields[f"{name}_b{s}"] = val[s]
            case.fields[f"{name}_bt{s}"] = (np.zeros_like(val[s]) if tendency is None else np.asarray(tendency(f"{name}_{s}", val[s]), np.float32))
    return case
