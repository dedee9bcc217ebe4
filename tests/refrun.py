"""Runs one routine of the reference (through tests/fortran_subset.py) on the state an Oracle object holds, bound exactly
as dyn_em/solve_em.F binds it: the CALL statement of the routine is read from solve_em.F and every actual argument
is resolved to the oracle's field of that name (grid%u_2 -> "u_2", ww1 -> "ww1", ...), to a dimension, a namelist value
or a step scalar.  Test infrastructure; needs the reference tree (/root/reference), i.e. the build container."""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

from wrf_b200 import decomp
from wrf_b200.ideal import Config
from tests.fortran_subset import Library, _cpp, _strip_comment, module_parameters

REF = "/root/reference"
f32 = np.float32
_CONSTANTS = None


def have_reference():
    return os.path.isdir(os.path.join(REF, "dyn_em"))


def call_statement(routine, occurrence=0, path="dyn_em/solve_em.F", defines=None):
    """the occurrence-th `CALL routine (...)` of a reference file as (positional actuals, keyword actuals)"""
    lines = [_strip_comment(ln).rstrip() for ln in _cpp(open(os.path.join(REF, path), errors="replace").read().split("\n"), defines or {"DM_PARALLEL": 1})]
    cur, found = "", []
    for ln in lines:
        t = ln.strip()
        if not t or t.startswith("!$"):
            continue
        if cur:
            cur += " " + (t[1:] if t.startswith("&") else t)
        else:
            cur = t
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        m = re.match(r"call\s+%s\s*\((.*)\)\s*$" % re.escape(routine), cur, re.I)
        if m:
            found.append(m.group(1))
        cur = ""
    if isinstance(occurrence, str):          # the call whose argument list mentions this text (the scalar loop's: "scalar_tend(")
        text = next(t for t in found if occurrence in t.lower().replace(" ", ""))
    else:
        text = found[occurrence]
    args, depth, piece = [], 0, ""
    for ch in text:
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            args.append(piece.strip())
            piece = ""
        else:
            piece += ch
    args.append(piece.strip())
    pos, kw = [], {}
    for a in args:
        m = re.match(r"([A-Za-z_]\w*)\s*=\s*(.+)$", a)
        if m and not re.match(r"[A-Za-z_]\w*\s*==", a):
            kw[m.group(1).lower()] = m.group(2).strip()
        else:
            pos.append(a)
    return pos, kw


CONFIG_FLAGS = ("periodic_x", "periodic_y", "symmetric_xs", "symmetric_xe", "symmetric_ys", "symmetric_ye", "open_xs", "open_xe", "open_ys",
                "open_ye", "specified", "nested", "polar", "non_hydrostatic", "top_lid", "pert_coriolis", "mix_full_fields")
CONFIG_INT = ("damp_opt", "w_damping", "diff_opt", "km_opt", "h_mom_adv_order", "v_mom_adv_order", "h_sca_adv_order", "v_sca_adv_order",
              "time_step", "rk_ord", "spec_zone", "relax_zone", "spec_bdy_width", "use_theta_m", "hypsometric_opt", "phi_adv_z", "map_proj",
              "zadvect_implicit", "momentum_adv_opt", "diff_6th_opt", "diff_6th_slopeopt", "mix_isotropic", "isfflx")
CONFIG_REAL = ("dampcoef", "zdamp", "w_crit_cfl", "khdif", "kvdif", "epssm", "smdiv", "emdiv", "diff_6th_factor", "diff_6th_thresh", "c_s", "c_k",
               "mix_upper_bound", "tke_upper_bound", "tke_drag_coefficient", "tke_heat_flux", "dx", "dy", "dt", "p_top")


SOURCES = ("dyn_em/module_em.F", "dyn_em/module_advect_em.F", "dyn_em/module_big_step_utilities_em.F", "dyn_em/module_small_step_em.F",
           "dyn_em/module_bc_em.F", "dyn_em/module_ieva_em.F", "dyn_em/module_diffusion_em.F", "share/module_bc.F",
           "phys/module_mp_kessler.F", "dyn_em/module_initialize_ideal.F", "dyn_em/module_init_utilities.F")
_LIBRARY = None


def library(time_step):
    """every SUBROUTINE of the path's reference files, translated on first use (tests/fortran_subset.Library)"""
    global _CONSTANTS, _LIBRARY
    if _LIBRARY is None:
        _CONSTANTS = module_parameters(os.path.join(REF, "share", "module_model_constants.F"))
        _CONSTANTS["param_first_scalar"] = (2, "i")          # tools/gen_mod_state_descr.c:69
        # module_state_description: one INTEGER PARAMETER per Registry `package` (original_mom = 1, weno_mom = 3, positivedef = 1, ...)
        for ln in open(os.path.join(REF, "Registry", "Registry.EM_COMMON"), errors="replace"):
            m = re.match(r"package\s+(\w+)\s+\w+==(-?\d+)\b", ln)
            if m:
                _CONSTANTS.setdefault(m.group(1).lower(), (int(m.group(2)), "i"))
        for f in SOURCES:          # module-level PARAMETERs of the files themselves (alpha_min of module_ieva_em, bdyzone of module_bc)
            for k, v in module_parameters(os.path.join(REF, f), _CONSTANTS, header_only=True).items():
                _CONSTANTS.setdefault(k, v)
        comp = dict([(k, "l") for k in CONFIG_FLAGS] + [(k, "i") for k in CONFIG_INT] + [(k, "r") for k in CONFIG_REAL])
        for k, ct in Config._fields_:            # every namelist entry the path's configuration carries
            comp.setdefault(k, "r" if ct is ctypes.c_float else "i")
        for ln in open(os.path.join(REF, "Registry", "Registry.EM_COMMON"), errors="replace"):     # and every other rconfig entry, by its type
            f = ln.split()
            if len(f) >= 3 and f[0] == "rconfig" and f[1] in ("real", "integer", "logical"):
                comp.setdefault(f[2].lower(), f[1][0])
        types = {"config_flags": comp}

        def fatal(*a, **k):
            raise RuntimeError("wrf_error_fatal called by the reference routine")

        hooks = {"nl_get_time_step": lambda _id, _ts: {1: _LIBRARY.time_step}, "wrf_error_fatal": fatal, "wrf_debug": lambda *a, **k: {},
                 "wrf_message": lambda *a, **k: {}, "get_current_time_string": lambda *a, **k: {}, "get_current_grid_name": lambda *a, **k: {}}
        _LIBRARY = Library([os.path.join(REF, f) for f in SOURCES], defines={"EM_CORE": 1}, types=types, constants=_CONSTANTS, hooks=hooks)
        _LIBRARY.globals.update({q: 1 for q in ("p_qv", "p_qc", "p_qr", "p_qi", "p_qs", "p_qg", "p_qh", "p_qni", "p_qnc")})
        # the NBA stress arrays of sfs_opt > 0 (not on this path): their species indices are 1 = unused
        _LIBRARY.globals.update({q: 1 for q in ("p_m11", "p_m12", "p_m13", "p_m22", "p_m23", "p_m33", "p_r12", "p_r13", "p_r23", "p_smnsmn")})
    _LIBRARY.time_step = int(time_step)
    return _LIBRARY


def set_species(c):
    """module_state_description's species indices for this configuration (set_scalar_indices_from_config): active species count up
    from PARAM_FIRST_SCALAR = 2 (tools/gen_mod_state_descr.c:69), inactive ones are 1"""
    g = library(c.time_step).globals
    for n, q in enumerate(("p_qv", "p_qc", "p_qr")):
        g[q] = n + 2 if n < c.num_moist else 1
    for q in ("p_qi", "p_qs", "p_qg", "p_qh", "p_qni", "p_qnc"):
        g[q] = 1


_REGISTRY = None


def registry_default(key):
    """the Registry default of a namelist entry the path's configuration does not carry (mp_tend_lim = 10., no_mp_heating = 0,
    rad_nudge = 0, ...): the value config_flags holds in a run whose namelist does not mention it"""
    global _REGISTRY
    if _REGISTRY is None:
        _REGISTRY = {}
        for ln in open(os.path.join(REF, "Registry", "Registry.EM_COMMON"), errors="replace"):
            f = ln.split()
            if len(f) >= 6 and f[0] == "rconfig":
                kind, name, val = f[1], f[2].lower(), f[5]
                try:
                    if kind == "real":
                        _REGISTRY[name] = f32(val.rstrip("fF"))
                    elif kind == "integer":      # -1 = "must be chosen" (the physics selectors): off, as the namelists of this path set them
                        _REGISTRY[name] = max(int(val), 0)
                    elif kind == "logical":
                        _REGISTRY[name] = val.lower().strip(".") in ("true", "t")
                except ValueError:
                    pass
    return _REGISTRY.get(key, 0)


class _Flags(dict):
    """config_flags: namelist entries the path does not carry have their Registry defaults"""

    def __missing__(self, key):
        return registry_default(key)


def _fortran_order(x):
    return np.ascontiguousarray(x.transpose(2, 1, 0) if x.ndim == 3 else (x.T if x.ndim == 2 else x))


class RefState:
    """The oracle's fields as the reference routines see them: Fortran index order (i,k,j), fetched from the oracle on first
    use and then owned by the reference side, so that a sequence of reference calls works on its own results."""

    def __init__(self, oracle):
        self.oracle = oracle
        self.have = set(oracle.names()) | {"moist", "moist_old", "moist_tend", "scalar", "scalar_old", "scalar_tend"}
        self.fields = {}
        self.absent = set()            # array actuals the oracle has no field for (bound to zeros)
        self.returns = []              # (routine, {scalar dummy: value after the call}) of every call made on this state

    def groups(self):
        """the 4-D species arrays of the Registry (moist, scalar, their _old and _tend forms) -> the oracle's per-species fields;
        Fortran index 1 is the unused slot below PARAM_FIRST_SCALAR"""
        c = self.oracle.case.config
        out = {}
        for fam, mid in (("moist", ""), ("moist_old", "old_"), ("moist_tend", "tend_")):
            out[fam] = ["moist_" + mid + q for q in ("qv", "qc", "qr")[:c.num_moist]]
        for fam, mid in (("scalar", ""), ("scalar_old", "old_"), ("scalar_tend", "tend_")):
            out[fam] = ["scalar_" + mid + str(n) for n in range(1, c.num_scalar + 1)]
        return out

    def array(self, key):
        if key not in self.fields:
            for fam, members in self.groups().items():
                if key == fam or key in members:
                    d = self.oracle.case.dims
                    a4 = np.zeros((d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1, len(members) + 1), np.float32)
                    for n, m in enumerate(members):
                        a4[..., n + 1] = _fortran_order(np.array(self.oracle.get(m), dtype=np.float32))
                        self.fields[m] = a4[..., n + 1]
                    self.fields[fam] = a4
                    self.absent.add(fam)          # compared through its members
                    return self.fields[key]
            x = np.array(self.oracle.get(key), dtype=np.float32)
            side = decomp.bdy_side(key)
            if side:        # a boundary array: (spec_bdy_width, levels, along) in C order = Fortran's (jms:jme | ims:ime, kds:kde | 1, width)
                d, W = self.oracle.case.dims, self.oracle.case.config.spec_bdy_width
                along = (d.jme - d.jms + 1) if side == "x" else (d.ime - d.ims + 1)
                x = x.reshape(W, x.size // (W * along), along)
            self.fields[key] = _fortran_order(x)
        return self.fields[key]

    def result(self, key):
        """a field after the reference calls, in the oracle's (j,k,i) order"""
        return _fortran_order(self.fields[key])


class RefCall:
    def __init__(self, routine, occurrence=0, call_path="dyn_em/solve_em.F", alias=None, actuals=None):
        """actuals: {dummy: actual text} for a routine solve_em reaches through a driver this path does not translate
        (first_rk_step_part2's calculate_km_kh -> isotropic_km); dummies it leaves out are bound by their own names"""
        self.R = library(1).routine(routine)
        if actuals is None:
            self.pos, self.kw = call_statement(routine, occurrence, call_path)
        else:
            self.pos, self.kw = [], {d: actuals.get(d, d) for d in self.R.dummies}
        self.alias = alias or {}

    def bind(self, state, scalars):
        """-> kwargs for the routine; arrays are the RefState's (shared between calls)"""
        oracle = state.oracle
        case = oracle.case
        d, c = case.dims, case.config
        R = self.R
        actual = dict(zip(R.dummies, self.pos))
        actual.update(self.kw)
        args = {}
        cfg = _Flags({k: bool(getattr(c, k)) for k in CONFIG_FLAGS if hasattr(c, k)})
        cfg.update({k: int(getattr(c, k)) for k in CONFIG_INT if hasattr(c, k)})
        cfg.update({k: f32(getattr(c, k)) for k in CONFIG_REAL if hasattr(c, k)})
        for k, ct in Config._fields_:
            cfg.setdefault(k, f32(getattr(c, k)) if ct is ctypes.c_float else int(getattr(c, k)))
        dims = {n: getattr(d, n) for n in ("ids", "ide", "jds", "jde", "kds", "kde", "ims", "ime", "jms", "jme", "kms", "kme")}
        dims.update(ips=d.ips, ipe=d.ipe, jps=d.jps, jpe=d.jpe, kps=d.kds, kpe=d.kde, its=d.ips, ite=d.ipe, jts=d.jps, jte=d.jpe,
                    kts=d.kds, kte=d.kde)
        tile = {"grid%i_start(ij)": d.ips, "grid%i_end(ij)": d.ipe, "grid%j_start(ij)": d.jps, "grid%j_end(ij)": d.jpe, "k_start": d.kds,
                "k_end": d.kde}
        tile["grid%warm_rain"] = c.mp_physics == 1          # mp_init, phys/module_physics_init.F:4463-4481 (Kessler)
        tile.update({"num_nba_mij": 1, "num_nba_rij": 1, "n_nba_mij": 1, "n_nba_rij": 1})
        tile.update({"num_moist": c.num_moist + 1, "num_scalar": c.num_scalar + 1, "n_moist": c.num_moist + 1})
        tile.update({"num_3d_m": c.num_moist + 1, "num_3d_s": c.num_scalar + 1, "grid%u_frame": f32(0), "grid%v_frame": f32(0),
                     "max_vert_cfl_tmp(ij)": f32(0), "max_horiz_cfl_tmp(ij)": f32(0), "grid%id": 1, "grid%itimestep": 1})
        for dummy in R.dummies:
            if dummy not in actual:
                if dummy in R.optional:
                    continue
                raise KeyError(f"{R.name}: no actual argument for {dummy}")
            if not isinstance(actual[dummy], str):       # the value itself
                args[dummy] = actual[dummy]
                continue
            a = actual[dummy].strip().lower().replace(" ", "")
            a = self.alias.get(a, a)
            if not isinstance(a, str):               # an alias may give the value itself
                args[dummy] = a
                continue
            key = a[5:] if a.startswith("grid%") else a
            if dummy in R.dims:
                if key not in state.have:        # an array of options the path does not have (physics tendencies, ...): zeros
                    ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
                    rank = len(R.dims[dummy])
                    state.absent.add(key)
                    if key not in state.fields:
                        state.fields[key] = np.zeros({4: (ni, nk, nj, 1), 3: (ni, nk, nj), 2: (ni, nj), 1: (nk,)}[rank], np.float32)
                    args[dummy] = state.fields[key]
                    continue
                args[dummy] = state.array(key)
            elif R.kind.get(dummy) == "t":
                args[dummy] = cfg
            elif a in scalars:
                args[dummy] = scalars[a]
            elif a in tile:
                args[dummy] = tile[a]
            elif key in dims:
                args[dummy] = dims[key]
            elif key in scalars:
                args[dummy] = scalars[key]
            elif a.startswith("config_flags%"):
                args[dummy] = cfg[a.split("%", 1)[1]]
            elif key in ("rdx", "rdy"):
                args[dummy] = f32(f32(1.0) / f32(c.dx if key == "rdx" else c.dy))
            elif key in cfg:
                args[dummy] = cfg[key]
            elif hasattr(c, key):
                v = getattr(c, key)
                args[dummy] = f32(v) if R.kind.get(dummy) == "r" else int(v)
            elif key in _CONSTANTS:                  # a module PARAMETER passed as an actual (t0, g, ...)
                args[dummy] = _CONSTANTS[key][0]
            elif re.fullmatch(r"[-+]?\d+", a):
                args[dummy] = int(a)
            elif re.fullmatch(r"[-+]?\d*\.\d*", a):
                args[dummy] = f32(a)
            elif a in (".true.", ".false."):
                args[dummy] = a == ".true."
            elif re.fullmatch(r"'[^']*'", a):
                args[dummy] = a[1:-1]
            else:
                raise KeyError(f"{R.name}: cannot resolve {dummy} <- {a}")
        return args

    def run(self, state, **scalars):
        set_species(state.oracle.case.config)
        return self.R(**self.bind(state, scalars))


def interior(case, a, stag=""):
    """the domain points of an oracle-ordered array (what the routines own; halo rows are the boundary fills' business)"""
    d = case.dims
    ie = d.ide if "x" in stag else d.ide - 1
    je = d.jde if "y" in stag else d.jde - 1
    sj, si = slice(d.jds - d.jms, je - d.jms + 1), slice(d.ids - d.ims, ie - d.ims + 1)
    return a[sj, ..., si] if a.ndim >= 2 else a
