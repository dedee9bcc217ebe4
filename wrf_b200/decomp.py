"""RSL_LITE patch decomposition on the host side (external/RSL_LITE/module_dm.F:98-175, :924-1003;
task_for_point.c): the rules themselves live behind the C-ABI (wrfgpu_compute_mesh,
wrfgpu_patch_dims -- csrc/decomp.cu); this module applies them to Case objects so that tests and
the benchmark can cut a global ideal-case state into the per-rank states wrf.exe would hold."""
from __future__ import annotations

import copy
import ctypes

import numpy as np

from .ideal import Case, Config, Dims
from .lib import load


def compute_mesh(ntasks: int, nproc_x: int = -1, nproc_y: int = -1):
    lib = load()
    nx, ny = ctypes.c_int(0), ctypes.c_int(0)
    if lib.wrfgpu_compute_mesh(ntasks, nproc_x, nproc_y, ctypes.byref(nx), ctypes.byref(ny)) != 0:
        raise ValueError(f"no {nproc_x}x{nproc_y} mesh for {ntasks} tasks")
    return nx.value, ny.value


def patch_dims(d: Dims, ntasks_x: int, ntasks_y: int, mytask: int) -> Dims:
    lib = load()
    out = Dims()
    rc = lib.wrfgpu_patch_dims(d.ids, d.ide, d.jds, d.jde, d.kds, d.kde, ntasks_x, ntasks_y, mytask, ctypes.byref(out))
    if rc != 0:
        raise ValueError("patch_dims failed (more tasks than points?)")
    return out


def patch_config(c: Config, ntasks_x: int, ntasks_y: int, mytask: int) -> Config:
    pc = copy.copy(c)
    pc = Config.from_buffer_copy(bytes(c))
    pc.ntasks_x, pc.ntasks_y, pc.mytask = ntasks_x, ntasks_y, mytask
    return pc


def window(gd: Dims, pd: Dims):
    """slices of the patch memory window inside the global memory arrays: (j-slice, i-slice)"""
    return (slice(pd.jms - gd.jms, pd.jme - gd.jms + 1), slice(pd.ims - gd.ims, pd.ime - gd.ims + 1))


def bdy_side(name: str) -> str:
    """'x' / 'y' for a boundary array name (<f>_bxs, <f>_btye, ...), '' for every other field"""
    for t in ("_bx", "_btx", "_by", "_bty"):
        if name[-len(t) - 1:-1] == t and name[-1] in "se":
            return t[-1]
    return ""


def split_case(case: Case, ntasks_x: int, ntasks_y: int, poison: bool = False):
    """global Case -> list of per-rank Cases (rank = my*ntasks_x + mx, module_dm.F:135-175).
    poison=True overwrites everything a rank does not own with NaN, so that a test fails unless the
    halo exchanges and boundary fills really deliver every value the rank later reads."""
    gd = case.dims
    out = []
    for r in range(ntasks_x * ntasks_y):
        pd = patch_dims(gd, ntasks_x, ntasks_y, r)
        pc = patch_config(case.config, ntasks_x, ntasks_y, r)
        sj, si = window(gd, pd)
        fl = {}
        for k, a in case.fields.items():
            side = bdy_side(k)
            if side:     # boundary arrays (spec_bdy_width, levels, jms:jme | ims:ime): cut along the side (wrf_b200/bdy.py)
                fl[k] = np.ascontiguousarray(a[..., sj if side == "x" else si]).copy()
                continue
            if a.ndim == 3:
                b = np.ascontiguousarray(a[sj, :, si]).copy()
            elif a.ndim == 2:
                b = np.ascontiguousarray(a[sj, si]).copy()
            else:
                b = a.copy()
            if poison and a.ndim >= 2:
                own_j = np.zeros(b.shape[0], bool)
                own_i = np.zeros(b.shape[-1], bool)
                own_j[pd.jps - pd.jms:pd.jpe - pd.jms + 1] = True
                own_i[pd.ips - pd.ims:pd.ipe - pd.ims + 1] = True
                mask = own_j[:, None] & own_i[None, :]
                if a.ndim == 3:
                    b[~np.broadcast_to(mask[:, None, :], b.shape)] = np.nan
                else:
                    b[~mask] = np.nan
            fl[k] = b
        out.append(Case(name=f"{case.name}[{r}/{ntasks_x}x{ntasks_y}]", dims=pd, config=pc, fields=fl))
    return out


XSTAG = {"u_1", "u_2", "ru", "ru_m", "ru_tend", "ru_tendf", "u_save", "cqu", "muu", "muus", "msfux", "msfuy"}
YSTAG = {"v_1", "v_2", "rv", "rv_m", "rv_tend", "rv_tendf", "v_save", "cqv", "muv", "muvs", "msfvx", "msfvx_inv", "msfvy"}


def owned(pd: Dims, a: np.ndarray, gd: Dims | None = None, name: str = ""):
    """view of the grid points the rank owns in a patch array (with gd: in a global array):
    [jps..jpe] x [ips..ipe] cut back to the field's own extent -- the last column i = ide exists only
    for X-staggered fields and the last row j = jde only for Y-staggered ones (Registry stagger column)"""
    d = gd or pd
    ie = pd.ipe if name in XSTAG else min(pd.ipe, pd.ide - 1)
    je = pd.jpe if name in YSTAG else min(pd.jpe, pd.jde - 1)
    sj = slice(pd.jps - d.jms, je - d.jms + 1)
    si = slice(pd.ips - d.ims, ie - d.ims + 1)
    return a[sj, ..., si] if a.ndim == 3 else a[sj, si]
