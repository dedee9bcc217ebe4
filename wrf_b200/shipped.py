"""The shipped ideal cases as ideal.exe + start_em would hand them to solve_em: namelist.input and input_sounding /
input_jet are read as the reference reads them, and the initial state follows dyn_em/module_initialize_ideal.F
(init_domain_rk :70-1738, get_sounding :1767, read_sounding :1927, get_sounding_b_wave :1965, calc_jet_sounding :2122,
read_input_jet :2200) and the first-trip diagnosis of start_domain_em (dyn_em/start_em.F:780-845: al and p re-derived from
ph with the model's own hypsometric option), statement for statement in real*4.

Difference that cannot be avoided without the reference's compiler, documented where it occurs:
  * em_les seeds its perturbation with RANDOM_NUMBER (module_initialize_ideal.F:1385): the stream is gfortran's; a
    deterministic hash of (i, j) takes its place (same amplitude, same four levels);
  (x**y, EXP, LOG, COS go to the host libm's float routines -- powf, expf, logf, cosf -- exactly as gfortran's real*4 intrinsics do;
  numpy's own float32 versions differ from them in the last bit now and then.  get_sounding / get_sounding_b_wave /
  calc_jet_sounding below equal the reference's routines bit for bit on the shipped inputs: tests/test_reference_init.py.)

Host language note: this is the role of ideal.exe, which is outside the hot path; Python is enough here."""
from __future__ import annotations

import math
import os
import re
import struct

import numpy as np

from . import ideal
from .ideal import F, Case, cosf, cvpm, expf, g, logf, p1000mb, powf, r_d, r_v, rcp, rvovrd, t0

def interp_0(v_in, z_in, z_out):
    """interp_0 (dyn_em/module_init_utilities.F:5-91), operand for operand: the table is NOT reversed when z decreases with the index
    (pressure tables), the bracketing pair is the one the reference's downward scan finds first (the higher index pair on a tie), the
    weight w2 belongs to the LOWER index, w1 = 1 - w2, result w1*v(kp) + w2*v(kp-1); the two extrapolation branches are the same
    expression with kp = nz_in or 2 (their numerator and denominator are both negated, the quotient is bit-identical).  Vectorised over z_out."""
    v_in, z_in = np.asarray(v_in, dtype=F), np.asarray(z_in, dtype=F)
    h = np.asarray(z_out, dtype=F)
    n = len(z_in)
    if z_in[n - 1] > z_in[0]:
        kp = np.searchsorted(z_in, h, side="right")            # nodes <= h: the pair (kp-1, kp) in 0-based indices
    else:
        kp = np.searchsorted(-z_in, -h, side="right")          # nodes >= h
    kp = np.clip(kp, 1, n - 1)
    w2 = ((h - z_in[kp]) / (z_in[kp - 1] - z_in[kp])).astype(F)
    w1 = (F(1.0) - w2).astype(F)
    out = ((w1 * v_in[kp]).astype(F) + (w2 * v_in[kp - 1]).astype(F)).astype(F)
    return F(out) if np.ndim(out) == 0 else out


HILL2D_X, QUARTER_SS, B_WAVE, LES = 1, 2, 7, 9      # Registry ideal_case values (share/module_model_constants / namelist &ideal)


# ----------------------------------------------------------------------------------------------------------------------
# namelist.input
def read_namelist(path):
    """Fortran namelist -> {group: {name: [values]}}; values are int / float / bool / str.  Handles what the shipped
    namelist.input files use: `name = v1, v2, ...` entries, .true./.false., quoted strings, `!` comments, `/` terminators."""
    groups, cur = {}, None
    text = open(path).read()
    for raw in text.splitlines():
        line = raw.split("!")[0].strip()
        if not line:
            continue
        if line.startswith("&"):
            cur = groups.setdefault(line[1:].strip().lower(), {})
            continue
        if line.startswith("/"):
            cur = None
            continue
        if cur is None or "=" not in line:
            continue
        name, rhs = line.split("=", 1)
        vals = []
        for tok in re.findall(r"'[^']*'|\"[^\"]*\"|[^,\s]+", rhs):
            t = tok.strip()
            if not t or t == "/":
                continue
            tl = t.lower()
            if tl in (".true.", ".t.", "t"):
                vals.append(True)
            elif tl in (".false.", ".f.", "f"):
                vals.append(False)
            elif t[0] in "'\"":
                vals.append(t[1:-1])
            else:
                try:
                    vals.append(int(t))
                except ValueError:
                    vals.append(float(t.replace("d", "e").replace("D", "e")))
        cur[name.strip().lower()] = vals
    return groups


# Registry.EM_COMMON rconfig defaults of the entries the path reads (namelist.input only lists what it changes)
# Registry/Registry.EM_COMMON rconfig defaults (+ registry.hyb_coord for hybrid_opt, etac), pinned against the reference by
# tests/test_reference_tables.py.  -1 = "has to be set in namelist.input" (diff_opt, km_opt: share/module_check_a_mundo.F:320-325).
REGISTRY_DEFAULTS = dict(
    rk_ord=3, time_step_sound=0, h_mom_adv_order=5, v_mom_adv_order=3, h_sca_adv_order=5, v_sca_adv_order=3,
    moist_adv_opt=1, scalar_adv_opt=1, tke_adv_opt=1, momentum_adv_opt=1, non_hydrostatic=True, top_lid=False, damp_opt=3,
    zdamp=5000.0, dampcoef=0.2, w_damping=0, pert_coriolis=False, use_theta_m=1, hypsometric_opt=2, phi_adv_z=1, zadvect_implicit=0,
    w_crit_cfl=1.0, diff_opt=-1, km_opt=-1, diff_6th_opt=0, diff_6th_factor=0.12, diff_6th_thresh=0.10, diff_6th_slopeopt=0,
    moist_mix6_off=False, scalar_mix6_off=False, khdif=0.0, kvdif=0.0, c_s=0.25, c_k=0.15, mix_full_fields=False, mix_isotropic=0,
    mix_upper_bound=0.1, tke_upper_bound=1000.0, tke_drag_coefficient=0.0, tke_heat_flux=0.0, epssm=0.1, smdiv=0.1, emdiv=0.01,
    hybrid_opt=2, etac=0.2, mp_physics=-1, isfflx=1, use_adaptive_time_step=False, specified=False, nested=False, polar=False,
    periodic_x=False, periodic_y=False, symmetric_xs=False, symmetric_xe=False, symmetric_ys=False, symmetric_ye=False,
    open_xs=False, open_xe=False, open_ys=False, open_ye=False)


def namelist_value(nl, name, dom=1):
    """the value of `name` for domain `dom` (entries dimensioned max_domains repeat per domain), Registry default if absent"""
    for grp in nl.values():
        if name in grp:
            v = grp[name]
            return v[dom - 1] if len(v) >= dom else v[0]
    if name in REGISTRY_DEFAULTS:
        return REGISTRY_DEFAULTS[name]
    raise KeyError(name)


def config_from_namelist(nl, dom=1):
    """namelist.input -> (wrfgpu_config, dict of grid parameters) with Registry defaults for what the file leaves out"""
    v = lambda n: namelist_value(nl, n, dom)   # noqa: E731
    c = ideal.default_config()
    for n in ("rk_ord", "time_step_sound", "h_mom_adv_order", "v_mom_adv_order", "h_sca_adv_order", "v_sca_adv_order", "moist_adv_opt",
              "scalar_adv_opt", "tke_adv_opt", "momentum_adv_opt", "damp_opt", "w_damping", "use_theta_m", "hypsometric_opt", "phi_adv_z",
              "zadvect_implicit", "diff_opt", "km_opt", "diff_6th_opt", "diff_6th_slopeopt", "mix_isotropic", "mp_physics", "isfflx"):
        setattr(c, n, int(v(n)))
    if c.diff_opt == -1 or c.km_opt == -1:     # share/module_check_a_mundo.F:320-325
        raise ValueError("Both km_opt and diff_opt need to be set in the namelist.input file.")
    if c.mp_physics == -1:                     # no physics_suite fills it in for the idealised cases
        raise ValueError("mp_physics needs to be set in the namelist.input file.")
    for n in ("non_hydrostatic", "top_lid", "pert_coriolis", "mix_full_fields", "moist_mix6_off", "scalar_mix6_off", "use_adaptive_time_step",
              "periodic_x", "periodic_y", "symmetric_xs", "symmetric_xe", "symmetric_ys", "symmetric_ye", "open_xs", "open_xe", "open_ys",
              "open_ye", "specified", "nested", "polar"):
        setattr(c, n, 1 if v(n) else 0)
    for n in ("zdamp", "dampcoef", "w_crit_cfl", "diff_6th_factor", "diff_6th_thresh", "khdif", "kvdif", "c_s", "c_k", "mix_upper_bound",
              "tke_upper_bound", "tke_drag_coefficient", "tke_heat_flux", "epssm", "smdiv", "emdiv"):
        setattr(c, n, float(v(n)))
    ts = int(v("time_step"))
    num, den = int(namelist_value(nl, "time_step_fract_num") if "time_step_fract_num" in nl.get("domains", {}) else 0), \
        int(namelist_value(nl, "time_step_fract_den") if "time_step_fract_den" in nl.get("domains", {}) else 1)
    c.time_step = ts
    c.dt = float(ts) + (float(num) / float(den) if den else 0.0)       # grid%dt (share/set_timekeeping.F)
    c.dx, c.dy = float(v("dx")), float(v("dy"))
    c.num_moist = 3 if c.mp_physics == 1 else 1                         # Kessler: qv, qc, qr; passiveqv otherwise
    grid = dict(e_we=int(v("e_we")), e_sn=int(v("e_sn")), e_vert=int(v("e_vert")), ztop=float(v("ztop")), hybrid_opt=int(v("hybrid_opt")),
                etac=float(v("etac")), ideal_case=int(namelist_value(nl, "ideal_case")))
    return c, grid


# ----------------------------------------------------------------------------------------------------------------------
# input_sounding / input_jet
def read_sounding(path):
    """read_sounding (module_initialize_ideal.F:1927-1963): first line p_surf (mb), theta_surf (K), qv_surf (g/kg); then
    height (m), theta (K), qv (g/kg), u, v (m/s) per level until the end of the file"""
    rows = [ln.split() for ln in open(path) if ln.strip()]
    ps, ts, qvs = (float(x) for x in rows[0][:3])
    lev = np.array([[float(x) for x in r[:5]] for r in rows[1:]], dtype=np.float64)
    return dict(ps=ps, ts=ts, qvs=qvs, h=lev[:, 0].astype(F), th=lev[:, 1].astype(F), qv=lev[:, 2].astype(F), u=lev[:, 3].astype(F),
                v=lev[:, 4].astype(F))


def read_input_jet(path, nz=64, ny=80):
    """read_input_jet (:2200-2248): Fortran unformatted, big-endian (the GNU stanza builds with -fconvert=big-endian,
    arch/configure.defaults:788): record 1 = ny, nz; then u, theta, rho as (ny, nz) arrays; z = 125 + 250 (k-1)"""
    b = open(path, "rb").read()
    pos = 0

    def record():
        nonlocal pos
        n = struct.unpack(">i", b[pos:pos + 4])[0]
        data = b[pos + 4:pos + 4 + n]
        assert struct.unpack(">i", b[pos + 4 + n:pos + 8 + n])[0] == n
        pos += 8 + n
        return data
    ny_in, nz_in = struct.unpack(">ii", record())
    if (ny_in, nz_in) != (ny, nz):
        raise ValueError(f"error in input jet dimensions: {ny_in} x {nz_in}")
    out = {}
    for name in ("u", "th", "rho"):
        a = np.frombuffer(record(), dtype=">f4").astype(F).reshape(nz, ny)   # field_in(ny,nz) in Fortran order
        out[name] = a                                                        # [k, j] like u_jet(k,j)
    out["z"] = np.broadcast_to((125.0 + 250.0 * np.arange(nz, dtype=np.float64)).astype(F)[:, None], (nz, ny)).copy()
    return out


def jet_sounding(jet, jp, dry):
    """calc_jet_sounding (:2122-2196) for column jp (1-based): the snd dict of ideal.get_sounding with p_surf in Pa, qv in g/g"""
    k = jp - 1
    h, th, rho, u = jet["z"][:, k].copy(), jet["th"][:, k].copy(), jet["rho"][:, k].copy(), jet["u"][:, k].copy()
    nz = len(h)
    qv = np.zeros(nz, F)
    if dry:
        rel_hum = np.zeros(nz, F)
    else:
        rel_hum = np.where(h > F(8000.0), F(0.1), F(1.0) - F(0.90) * powf(h / F(8000.0), F(1.25))).astype(F)
        rel_hum = np.minimum(F(0.7), rel_hum)
    cpovcv = F(ideal.cp / (ideal.cp - r_d))
    p = (p1000mb * powf(r_d * rho * th / p1000mb, cpovcv)).astype(F)
    if not dry:
        svp1, svp2, svp3, svpt0, ep_2 = F(0.6112), F(17.67), F(29.65), F(273.15), F(r_d / r_v)
        for kk in range(nz):
            tmppi = powf(p[kk] / p1000mb, rcp)
            temperature = tmppi * th[kk]
            if temperature > svpt0:
                es = F(1000.0) * svp1 * expf(svp2 * (temperature - svpt0) / (temperature - svp3))
            else:
                es = F(1000.0) * svp1 * expf(F(21.8745584) * (temperature - F(273.16)) / (temperature - F(7.66)))
            qvs = ep_2 * es / (p[kk] - es)
            qv[kk] = rel_hum[kk] * qvs
            th[kk] = th[kk] / (F(1.0) + F(0.61) * qv[kk])
    return dict(p_surf_pa=F(F(1.5) * p[0] - F(0.5) * p[1]), ts=F(F(1.5) * th[0] - F(0.5) * th[1]), h=h.astype(F), th=th.astype(F),
                qv_gg=qv.astype(F), u=u.astype(F), v=np.zeros(nz, F))


def get_sounding(snd, dry):
    """get_sounding (:1767-1925) / get_sounding_b_wave (:1965-2120): hydrostatic moist pressure from the surface up (ten
    iterations per level), dry pressure from the top down"""
    h, th = snd["h"].astype(F), snd["th"].astype(F)
    if "qv_gg" in snd:                                 # the jet's moisture is already g/g and its surface pressure Pa
        qv = (np.zeros_like(h) if dry else snd["qv_gg"]).astype(F)
        p_surf = F(snd["p_surf_pa"])
    else:
        qv = (np.zeros_like(h) if dry else (snd["qv"].astype(F) * F(0.001))).astype(F)
        p_surf = F(F(100.0) * F(snd["ps"]))
    nl = len(h)
    th_surf = F(snd["ts"])
    qvf = F(1.0) + rvovrd * qv[0]
    rho_surf = F(1.0) / ((r_d / p1000mb) * th_surf * qvf * powf(p_surf / p1000mb, cvpm))
    rho, pm = np.zeros(nl, F), np.zeros(nl, F)
    qvf1 = F(1.0) + qv[0]
    rho[0] = rho_surf
    dz = h[0]
    for _ in range(10):
        pm[0] = p_surf - F(0.5) * dz * (rho_surf + rho[0]) * g * qvf1
        rho[0] = F(1.0) / ((r_d / p1000mb) * th[0] * qvf * powf(pm[0] / p1000mb, cvpm))
    for k in range(1, nl):
        rho[k] = rho[k - 1]
        dz = h[k] - h[k - 1]
        qvf1 = F(0.5) * (F(2.0) + (qv[k - 1] + qv[k]))
        qvf = F(1.0) + rvovrd * qv[k]
        for _ in range(10):
            pm[k] = pm[k - 1] - F(0.5) * dz * (rho[k] + rho[k - 1]) * g * qvf1
            if pm[k] <= 0:
                raise ValueError("Integrated pressure has gone negative - too cold for chosen height")
            rho[k] = F(1.0) / ((r_d / p1000mb) * th[k] * qvf * powf(pm[k] / p1000mb, cvpm))
    pd = np.zeros(nl, F)
    pd[nl - 1] = pm[nl - 1]
    for k in range(nl - 2, -1, -1):
        dz = h[k + 1] - h[k]
        pd[k] = pd[k + 1] + F(0.5) * dz * (rho[k] + rho[k + 1]) * g
    return dict(zk=h, p=pm, pd=pd, theta=th, rho=rho, u=snd["u"].astype(F), v=snd["v"].astype(F), qv=qv, th_surf=th_surf, p_surf=p_surf)


# ----------------------------------------------------------------------------------------------------------------------
def _hash01(i, j, seed=20260101):
    """deterministic stand-in for RANDOM_NUMBER: uniform [0,1) from the global (i, j)"""
    m = 0xFFFFFFFFFFFFFFFF
    x = (int(i) * 0x9E3779B97F4A7C15 + int(j) * 0xC2B2AE3D27D4EB4F + int(seed)) & m
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & m
    x ^= x >> 33
    return float(x >> 40) / float(1 << 24)


def init_domain(rundir, dom=1) -> Case:
    """ideal.exe + the first trip of start_domain_em for the run directory `rundir` (namelist.input + input_sounding or
    input_jet, as shipped under test/em_*)."""
    nl = read_namelist(os.path.join(rundir, "namelist.input"))
    c, gp = config_from_namelist(nl, dom)
    case_id = gp["ideal_case"]
    if case_id not in (HILL2D_X, QUARTER_SS, B_WAVE, LES):
        raise ValueError(f"ideal_case {case_id} is not one of the BASELINE configurations (1 hill2d_x, 2 quarter_ss, 7 b_wave, 9 les)")
    nx, ny, nz = gp["e_we"], gp["e_sn"], gp["e_vert"]
    ide, jde, kde = nx, ny, nz
    d = ideal.memory_dims(nx, ny, nz)
    stretch = case_id != LES
    v, sc = ideal.vertical_coords(nz, gp["ztop"], stretch, exp=expf)
    for k_, val in sc.items():
        setattr(c, k_, float(val))
    pi = F(2.0) * np.arcsin(F(1.0), dtype=F)

    jet = None
    if case_id == B_WAVE:
        jet = read_input_jet(os.path.join(rundir, "input_jet"))
        dry_s = get_sounding(jet_sounding(jet, 80 // 2, True), dry=True)
    else:
        snd = read_sounding(os.path.join(rundir, "input_sounding"))
        dry_s = get_sounding(snd, dry=True)
    p_top = F(interp_0(dry_s["p"], dry_s["zk"], F(gp["ztop"])))
    c.p_top = float(p_top)
    c3f = None
    if gp["hybrid_opt"] == 2:       # :795-809, x**n as gfortran expands it (__powisf2: x*x, x*(x*x), (x*x)*(x*x))
        e, znw = F(gp["etac"]), v["znw"]
        sq = lambda x: (x * x).astype(F) if isinstance(x, np.ndarray) else F(x * x)
        cube = lambda x: (x * sq(x)).astype(F) if isinstance(x, np.ndarray) else F(x * sq(x))
        B1 = F(F(F(2.0) * sq(e)) * F(F(1.0) - e))
        B2 = F(-e * F(F(F(4.0) - F(F(3.0) * e)) - cube(e)))
        B3 = F(F(2.0) * F(F(1.0) - cube(e)))
        B4 = F(-F(F(1.0) - sq(e)))
        B5 = sq(sq(F(F(1.0) - e)))
        c3f = ((((B1 + (B2 * znw).astype(F)).astype(F) + (B3 * sq(znw)).astype(F)).astype(F) + (B4 * cube(znw)).astype(F)).astype(F) / B5).astype(F)
        c3f[znw < e] = 0
        c3f[0], c3f[-1] = 1, 0
    hy = ideal.hybrid_coefs(v, float(p_top), gp["hybrid_opt"], gp["etac"], c3f=c3f)
    c3h, c4h, c1h, c2h, c1f, c2f, c3f, c4f = (hy[n] for n in ("c3h", "c4h", "c1h", "c2h", "c1f", "c2f", "c3f", "c4f"))
    dnw, rdnw, rdn = v["dnw"], v["rdnw"], v["rdn"]

    nj, nk, ni = d.jme - d.jms + 1, nz, d.ime - d.ims + 1
    I0, J0 = -d.ims, -d.jms                      # array index of Fortran index 0
    fl = {n: np.zeros((nj, nk, ni), F) for n in ("u", "v", "t", "ph", "phb", "pb", "alb", "p", "al", "alt", "t_init", "qv")}
    mub, mu, ht = (np.zeros((nj, ni), F) for _ in range(3))
    its, ite, jts, jte = 1, ide, 1, jde
    # terrain (:876-960)
    if case_id == HILL2D_X:
        hm, xa, icm = F(100.0), F(5.0), ide // 2
        for i in range(its, ite + 1):
            ht[:, I0 + i] = hm / (F(1.0) + (F(i - icm) / xa) ** 2)
    # base state (:962-985), every column of the patch (flat terrain: one column, copied)
    def base_column(jj, ii):
        p_surf = F(interp_0(dry_s["p"], dry_s["zk"], F(g * ht[jj, ii] / g)))
        mub[jj, ii] = p_surf - p_top
        pl = (c3h[:kde - 1] * (p_surf - p_top) + c4h[:kde - 1] + p_top).astype(F)
        fl["pb"][jj, :kde - 1, ii] = pl
        ti = (interp_0(dry_s["theta"], dry_s["p"], pl) - t0).astype(F)
        fl["t_init"][jj, :kde - 1, ii] = ti
        fl["alb"][jj, :kde - 1, ii] = ((r_d / p1000mb) * (ti + t0) * powf(pl / p1000mb, cvpm)).astype(F)
        phb = fl["phb"][jj, :, ii]
        phb[0] = g * ht[jj, ii]
        for kk in range(1, kde):
            phb[kk] = phb[kk - 1] - dnw[kk - 1] * (c1h[kk - 1] * mub[jj, ii] + c2h[kk - 1]) * fl["alb"][jj, kk - 1, ii]
    if case_id == HILL2D_X:
        for i in range(its, ite + 1):            # the terrain does not vary in y
            base_column(J0 + jts, I0 + i)
        for n in ("pb", "t_init", "alb", "phb"):
            fl[n][J0 + jts:J0 + jte + 1] = fl[n][J0 + jts][None]
        mub[J0 + jts:J0 + jte + 1] = mub[J0 + jts][None]
    else:
        base_column(J0 + jts, I0 + its)
        for n in ("pb", "t_init", "alb", "phb"):
            fl[n][J0 + jts:J0 + jte + 1, :, I0 + its:I0 + ite + 1] = fl[n][J0 + jts, :, I0 + its][None, :, None]
        mub[J0 + jts:J0 + jte + 1, I0 + its:I0 + ite + 1] = mub[J0 + jts, I0 + its]
    p_surf_last = F(mub[J0 + jte, I0 + ite] + p_top)       # the scalar p_surf the b_wave u interpolation inherits (:1085-1091)
    # full state (:994-1094)
    moist_s = None if case_id == B_WAVE else get_sounding(snd, dry=False)
    u_cols = {}
    for j in range(jts, min(jde - 1, jte) + 1):
        if case_id == B_WAVE:
            moist_s = get_sounding(jet_sounding(jet, j, c.mp_physics == 0), dry=(c.mp_physics == 0))
        col_cache = None
        for i in range(its, min(ide - 1, ite) + 1):
            jj, ii = J0 + j, I0 + i
            if col_cache is not None and case_id != HILL2D_X:   # flat terrain: the column only depends on j
                for n in ("qv", "t", "p", "alt", "al", "ph"):
                    fl[n][jj, :, ii] = col_cache[n]
                mu[jj, ii] = col_cache["mu"]
                if case_id == B_WAVE:
                    fl["u"][jj, :, ii] = col_cache["u"]
                continue
            pd_surf = F(interp_0(moist_s["pd"], moist_s["zk"], F(fl["phb"][jj, 0, ii] / g)))
            mu1 = F(pd_surf - p_top - mub[jj, ii])
            mu[jj, ii] = mu1
            pl = (c3h[:kde - 1] * (pd_surf - p_top) + c4h[:kde - 1] + p_top).astype(F)
            qv = interp_0(moist_s["qv"], moist_s["pd"], pl).astype(F)
            t1 = (interp_0(moist_s["theta"], moist_s["pd"], pl) - t0).astype(F)
            p = np.zeros(kde, F)
            pbc, albc = fl["pb"][jj, :, ii], fl["alb"][jj, :, ii]
            alt, al = np.zeros(kde, F), np.zeros(kde, F)
            kk = kde - 2
            k = kk + 1
            qvf1 = F(0.5) * (qv[kk] + qv[kk])
            qvf2 = F(1.0) / (F(1.0) + qvf1)
            qvf1 = qvf1 * qvf2
            p[kk] = -F(0.5) * ((c1f[k] * mu1) + qvf1 * (c1f[k] * mub[jj, ii] + c2f[k])) / rdnw[kk] / qvf2
            qvf = F(1.0) + rvovrd * qv[kk]
            alt[kk] = (r_d / p1000mb) * (t1[kk] + t0) * qvf * powf((p[kk] + pbc[kk]) / p1000mb, cvpm)
            al[kk] = alt[kk] - albc[kk]
            for kk in range(kde - 3, -1, -1):
                k = kk + 1
                qvf1 = F(0.5) * (qv[kk] + qv[kk + 1])
                qvf2 = F(1.0) / (F(1.0) + qvf1)
                qvf1 = qvf1 * qvf2
                p[kk] = p[kk + 1] - ((c1f[k] * mu1) + qvf1 * (c1f[k] * mub[jj, ii] + c2f[k])) / qvf2 / rdn[kk + 1]
                qvf = F(1.0) + rvovrd * qv[kk]
                alt[kk] = (r_d / p1000mb) * (t1[kk] + t0) * qvf * powf((p[kk] + pbc[kk]) / p1000mb, cvpm)
                al[kk] = alt[kk] - albc[kk]
            ph = np.zeros(kde, F)
            for kk in range(1, kde):
                k = kk - 1
                ph[kk] = ph[kk - 1] - dnw[kk - 1] * (((c1h[k] * mub[jj, ii] + c2h[k]) + (c1h[k] * mu1)) * al[kk - 1] + (c1h[k] * mu1) * albc[kk - 1])
            fl["qv"][jj, :kde - 1, ii] = qv
            fl["t"][jj, :kde - 1, ii] = t1
            fl["p"][jj, :, ii], fl["alt"][jj, :, ii], fl["al"][jj, :, ii], fl["ph"][jj, :, ii] = p, alt, al, ph
            col_cache = {n: fl[n][jj, :, ii].copy() for n in ("qv", "t", "p", "alt", "al", "ph")}
            col_cache["mu"] = mu1
            if case_id == B_WAVE:
                plu = (c3h[:kde] * (p_surf_last - p_top) + c4h[:kde] + p_top).astype(F)
                fl["u"][jj, :, ii] = interp_0(moist_s["u"], moist_s["p"], plu)
                col_cache["u"] = fl["u"][jj, :, ii].copy()

    # perturbation (:1098-1436) and hydrostatic rebalance of the columns it touched
    def perturb(j, i, dth):
        jj, ii = J0 + j, I0 + i
        hit = False
        for k in range(kde - 1):
            if dth[k] != 0.0:
                fl["t"][jj, k, ii] = fl["t"][jj, k, ii] + dth[k]
                qvf = F(1.0) + rvovrd * fl["qv"][jj, k, ii]
                fl["alt"][jj, k, ii] = (r_d / p1000mb) * (fl["t"][jj, k, ii] + t0) * qvf * powf((fl["p"][jj, k, ii] + fl["pb"][jj, k, ii]) / p1000mb, cvpm)
                fl["al"][jj, k, ii] = fl["alt"][jj, k, ii] - fl["alb"][jj, k, ii]
                hit = True
        if hit or case_id in (QUARTER_SS, B_WAVE, LES):      # the reference re-integrates every column of the loop
            ph = fl["ph"][jj, :, ii]
            for k in range(1, kde):
                ph[k] = ph[k - 1] - dnw[k - 1] * (((c1h[k - 1] * mub[jj, ii] + c2h[k - 1]) + (c1h[k - 1] * mu[jj, ii])) * fl["al"][jj, k - 1, ii] +
                                                  (c1h[k - 1] * mu[jj, ii]) * fl["alb"][jj, k - 1, ii])

    if case_id in (QUARTER_SS, B_WAVE, LES):
        nxc, nyc = ((ide - 1) // 2, (jde - 1) // 2) if case_id != B_WAVE else ((ide - 1) // 4, (jde - 1) // 2)
        delt = F(3.0)
        piov2 = F(2.0) * np.arctan(F(1.0), dtype=F)
        for j in range(jts, min(jde - 1, jte) + 1):
            for i in range(its, min(ide - 1, ite) + 1):
                jj, ii = J0 + j, I0 + i
                dth = np.zeros(kde - 1, F)
                zmid = (F(0.5) * (fl["ph"][jj, :-1, ii] + fl["ph"][jj, 1:, ii] + fl["phb"][jj, :-1, ii] + fl["phb"][jj, 1:, ii]) / g).astype(F)
                if case_id == QUARTER_SS:
                    yrad = F(c.dy) * F(j - nyc) / F(10000.0)
                    xrad = F(c.dx) * F(i - nxc) / F(10000.0)
                    zrad = (zmid - F(1500.0)) / F(1500.0)
                    rad = np.sqrt(xrad * xrad + yrad * yrad + zrad * zrad, dtype=F)
                    dth = np.where(rad <= F(1.0), delt * cosf(F(0.5) * pi * rad) ** 2, F(0.0)).astype(F)
                elif case_id == B_WAVE:
                    yrad = F(c.dy) * F(j - jde // 2 - 1) / F(2000000.0)
                    xrad = F(i - 1) / F(ide - 1)
                    zrad = (zmid - F(8000.0)) / F(8000.0)
                    rad = np.sqrt(yrad * yrad + zrad * zrad, dtype=F)
                    tp = F(1.0) * cosf(rad * piov2) * cosf(rad * piov2) * cosf(xrad * F(2.0) * pi + pi)
                    dth = np.where(rad <= F(1.0), tp, F(0.0)).astype(F)
                else:      # les: 0.1 (rand - 0.5) on the lowest four levels (RANDOM_NUMBER replaced, see the module docstring)
                    dth[:4] = F(0.1) * F(_hash01(i, j) - 0.5)
                perturb(j, i, dth)

    # winds (:1498-1555); w = 0
    if case_id == B_WAVE:
        fl["u"][J0 + jts:J0 + min(jde - 1, jte) + 1, :, I0 + ite] = fl["u"][J0 + jts:J0 + min(jde - 1, jte) + 1, :, I0 + its]
    else:
        wind_s = moist_s
        for j in range(jts, jte + 1):
            for i in range(its, min(ide - 1, ite) + 1):
                jj, ii = J0 + j, I0 + i
                if j == 1:
                    z_at_v = fl["phb"][jj, 0, ii] / g
                elif j == jde:
                    z_at_v = fl["phb"][jj - 1, 0, ii] / g
                else:
                    z_at_v = F(0.5) * (fl["phb"][jj, 0, ii] + fl["phb"][jj - 1, 0, ii]) / g
                ps = F(interp_0(wind_s["p"], wind_s["zk"], F(z_at_v)))
                pl = (c3h[:kde] * (ps - p_top) + c4h[:kde] + p_top).astype(F)
                fl["v"][jj, :, ii] = interp_0(wind_s["v"], wind_s["p"], pl)
                if case_id != HILL2D_X and i > its:       # flat: same profile everywhere
                    fl["v"][jj, :, I0 + its:I0 + min(ide - 1, ite) + 1] = fl["v"][jj, :, ii][:, None]
                    break
        for j in range(jts, min(jde - 1, jte) + 1):
            for i in range(its, ite + 1):
                jj, ii = J0 + j, I0 + i
                if i == 1:
                    z_at_u = fl["phb"][jj, 0, ii] / g
                elif i == ide:
                    z_at_u = fl["phb"][jj, 0, ii - 1] / g
                else:
                    z_at_u = F(0.5) * (fl["phb"][jj, 0, ii] + fl["phb"][jj, 0, ii - 1]) / g
                ps = F(interp_0(wind_s["p"], wind_s["zk"], F(z_at_u)))
                pl = (c3h[:kde] * (ps - p_top) + c4h[:kde] + p_top).astype(F)
                fl["u"][jj, :, ii] = interp_0(wind_s["u"], wind_s["p"], pl)
    # base profiles from column (1,1) (:1579-1591)
    b1 = {n: np.zeros(nz, F) for n in ("u_base", "v_base", "t_base", "qv_base", "z_base")}
    j1, i1 = J0 + 1, I0 + 1
    b1["t_base"][:-1] = fl["t"][j1, :-1, i1]
    b1["qv_base"][:-1] = fl["qv"][j1, :-1, i1]
    b1["u_base"][:-1] = fl["u"][j1, :-1, i1]
    b1["v_base"][:-1] = fl["v"][j1, :-1, i1]
    b1["z_base"][:-1] = F(0.5) * (fl["phb"][j1, :-1, i1] + fl["phb"][j1, 1:, i1] + fl["ph"][j1, :-1, i1] + fl["ph"][j1, 1:, i1]) / g
    # moist theta (:1726-1734)
    if c.use_theta_m == 1:
        sl = (slice(J0 + jts, J0 + min(jde - 1, jte) + 1), slice(0, kde - 1), slice(I0 + its, I0 + min(ide - 1, ite) + 1))
        fl["t"][sl] = (fl["t"][sl] + t0) * (F(1.0) + (r_v / r_d) * fl["qv"][sl]) - t0
    # first trip of start_domain_em (start_em.F:780-845): al from ph with the model's hypsometric option, p from the EOS
    sj, si = slice(J0 + jts, J0 + min(jde - 1, jte) + 1), slice(I0 + its, I0 + min(ide - 1, ite) + 1)
    MUB, MU = mub[sj, None, si], mu[sj, None, si]
    kk = slice(0, kde - 1)
    C1H, C2H, RDNW = c1h[None, kk, None], c2h[None, kk, None], rdnw[None, kk, None]
    ph, phb, alb = fl["ph"][sj, :, si], fl["phb"][sj, :, si], fl["alb"][sj, kk, si]
    if c.hypsometric_opt == 1:
        al = (F(-1.0) / ((C1H * MUB + C2H) + (C1H * MU)) * (alb * (C1H * MU) + RDNW * (ph[:, 1:, :] - ph[:, :-1, :]))).astype(F)
    else:
        pfu = c3f[None, 1:kde, None] * (MUB + MU) + c4f[None, 1:kde, None] + p_top
        pfd = c3f[None, 0:kde - 1, None] * (MUB + MU) + c4f[None, 0:kde - 1, None] + p_top
        phm = c3h[None, kk, None] * (MUB + MU) + c4h[None, kk, None] + p_top
        al = ((ph[:, 1:, :] - ph[:, :-1, :] + phb[:, 1:, :] - phb[:, :-1, :]) / phm / logf(pfd / pfu) - alb).astype(F)
    qvf = F(1.0) if c.use_theta_m == 1 else (F(1.0) + rvovrd * fl["qv"][sj, kk, si])
    cpovcv = F(ideal.cp / (ideal.cp - r_d))
    fl["al"][sj, kk, si] = al
    fl["p"][sj, kk, si] = (p1000mb * powf((r_d * (t0 + fl["t"][sj, kk, si]) * qvf) / (p1000mb * (al + alb)), cpovcv) - fl["pb"][sj, kk, si]).astype(F)
    fl["alt"][sj, kk, si] = al + alb

    # ---- the fields of the C-ABI
    out = {}
    for n, a in (("u", fl["u"]), ("v", fl["v"]), ("t", fl["t"]), ("ph", fl["ph"])):
        out[n + "_1"], out[n + "_2"] = a.copy(), a.copy()
    out["w_1"], out["w_2"] = np.zeros((nj, nk, ni), F), np.zeros((nj, nk, ni), F)
    out["mu_1"], out["mu_2"], out["mub"], out["ht"] = mu.copy(), mu.copy(), mub, ht
    for n in ("phb", "pb", "alb", "p", "al", "alt", "t_init"):
        out[n] = fl[n]
    out["moist_qv"] = fl["qv"]
    ones = np.ones((nj, ni), F)
    for n in ("msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy", "cosa"):
        out[n] = ones.copy()
    out["f"] = np.full((nj, ni), 1.0e-4 if case_id in (B_WAVE, LES) else 0.0, F)
    out["e"], out["sina"] = np.zeros((nj, ni), F), np.zeros((nj, ni), F)
    for n in ("znu", "znw", "fnm", "fnp", "rdnw", "rdn", "dnw", "dn"):
        out[n] = v[n]
    for n in ("c1h", "c2h", "c1f", "c2f", "c3h", "c4h", "c3f", "c4f"):
        out[n] = hy[n]
    out.update(b1)
    if c.diff_opt == 2 and c.isfflx != 0:
        # ustm, hfx, qfx of the surface-layer scheme (sf_sfclay_physics: host-side physics, outside the hot path; uploaded per
        # step by the caller).  Without one they keep their allocation-time zeros, as in wrf.exe with sf_sfclay_physics = 0.
        for n in ("ustm", "hfx", "qfx"):
            out[n] = np.zeros((nj, ni), F)
    name = {HILL2D_X: "em_hill2d_x", QUARTER_SS: "em_quarter_ss", B_WAVE: "em_b_wave", LES: "em_les"}[case_id]
    return Case(name=name + " (shipped)", dims=d, config=c, fields={k: np.ascontiguousarray(a, dtype=F) for k, a in out.items()})
