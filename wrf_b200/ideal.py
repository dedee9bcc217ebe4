"""Idealised initial states for the solve_em hot path (host-side test/bench input).

This plays the role of ideal.exe for the parity tests and the benchmark: it
follows the recipe of dyn_em/module_initialize_ideal.F (levels :617-732, hybrid
coefficients :791-877, base state :962-985, moist column balance :1000-1085,
quarter_ss bubble :1089-1133, winds :1500-1555, theta_m conversion :1726) with
analytic soundings, and returns plain numpy arrays in WRF memory order
(j, k, i) == Fortran (i,k,j).  It is NOT on the device hot path and is not the
oracle; both the CUDA path and the oracle consume exactly the same arrays.

Soundings are generated analytically instead of read from test/em_*/input_sounding
(those files live only in the read-only reference tree):
  * quarter_ss / synthetic : Weisman & Klemp (1982) thermodynamic profile with the
    quarter-circle hodograph of the em_quarter_ss case (qv capped at 14 g/kg);
  * hill2d_x               : N = 0.01 1/s, U = 10 m/s, dry;
  * b_wave                 : analytic mid-latitude jet (not the input_jet table);
  * les                    : neutral boundary layer capped by an inversion.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from dataclasses import dataclass, field

import numpy as np

F = np.float32
g = F(9.81)
r_d = F(287.0)
cp = F(7.0 * 287.0 / 2.0)
r_v = F(461.6)
cv = F(cp - r_d)
cvpm = F(-cv / cp)
rcp = F(r_d / cp)
p1000mb = F(100000.0)
t0 = F(300.0)
rvovrd = F(r_v / r_d)


_LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")


def _libm(name, nargs):
    """a real*4 intrinsic as gfortran evaluates it: the host libm's float routine, element by element"""
    fn = getattr(_LIBM, name)
    fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float] * nargs
    uf = np.frompyfunc(lambda *a: fn(*[float(x) for x in a]), nargs, 1)

    def call(*a):
        r = uf(*[np.asarray(x, dtype=F) for x in a])
        return F(r) if np.ndim(r) == 0 else np.asarray(r, dtype=F)
    return call


powf, expf, logf, cosf = _libm("powf", 2), _libm("expf", 1), _libm("logf", 1), _libm("cosf", 1)


class Dims(ctypes.Structure):
    """wrfgpu_dims (include/wrfgpu.h)."""
    _fields_ = [(n, ctypes.c_int) for n in (
        "ids", "ide", "jds", "jde", "kds", "kde",
        "ims", "ime", "jms", "jme", "kms", "kme",
        "ips", "ipe", "jps", "jpe", "kps", "kpe")]


class Config(ctypes.Structure):
    """wrfgpu_config (include/wrfgpu.h) -- field order must match the header."""
    _fields_ = (
        [(n, ctypes.c_int) for n in (
            "periodic_x", "periodic_y",
            "symmetric_xs", "symmetric_xe", "symmetric_ys", "symmetric_ye",
            "open_xs", "open_xe", "open_ys", "open_ye",
            "specified", "nested", "polar",
            "rk_ord", "time_step_sound",
            "h_mom_adv_order", "v_mom_adv_order", "h_sca_adv_order", "v_sca_adv_order",
            "moist_adv_opt", "scalar_adv_opt", "momentum_adv_opt",
            "non_hydrostatic", "top_lid", "damp_opt", "w_damping", "pert_coriolis",
            "use_theta_m", "hypsometric_opt", "phi_adv_z", "zadvect_implicit",
            "diff_opt", "km_opt", "diff_6th_opt", "mp_physics", "time_step", "map_proj")]
        + [(n, ctypes.c_float) for n in (
            "dt", "dx", "dy", "epssm", "smdiv", "emdiv", "dampcoef", "zdamp", "w_crit_cfl",
            "p_top", "cf1", "cf2", "cf3", "cfn", "cfn1", "khdif", "kvdif")]
        + [(n, ctypes.c_int) for n in ("num_moist", "num_scalar", "ntasks_x", "ntasks_y", "mytask")]
        + [(n, ctypes.c_float) for n in ("diff_6th_factor", "diff_6th_thresh")]
        + [(n, ctypes.c_int) for n in ("diff_6th_slopeopt", "moist_mix6_off", "scalar_mix6_off")]
        + [(n, ctypes.c_float) for n in ("c_s", "c_k")]
        + [("mix_full_fields", ctypes.c_int), ("isfflx", ctypes.c_int)]
        + [(n, ctypes.c_float) for n in ("tke_drag_coefficient", "tke_heat_flux")]
        + [("mix_isotropic", ctypes.c_int), ("mix_upper_bound", ctypes.c_float), ("tke_adv_opt", ctypes.c_int), ("tke_upper_bound", ctypes.c_float),
           ("use_adaptive_time_step", ctypes.c_int)]
        + [(n, ctypes.c_int) for n in ("spec_bdy_width", "spec_zone", "relax_zone", "have_bcs_moist", "have_bcs_scalar")])


def default_config(**kw) -> Config:
    """Registry defaults that matter on the path (Registry.EM_COMMON rconfig)."""
    c = Config()
    c.rk_ord = 3
    c.h_mom_adv_order = c.h_sca_adv_order = 5
    c.v_mom_adv_order = c.v_sca_adv_order = 3
    c.moist_adv_opt = c.scalar_adv_opt = 1
    c.momentum_adv_opt = 1
    c.non_hydrostatic = 1
    c.damp_opt = 0
    c.use_theta_m = 1
    c.hypsometric_opt = 2
    c.phi_adv_z = 1
    c.time_step = 1
    c.epssm, c.smdiv, c.emdiv = 0.1, 0.1, 0.01
    c.dampcoef, c.zdamp = 0.2, 5000.0
    c.w_crit_cfl = 1.0      # Registry.EM_COMMON:2842
    c.diff_6th_factor, c.diff_6th_thresh = 0.12, 0.10
    c.c_s, c.c_k = 0.25, 0.15
    c.mix_upper_bound = 0.1
    c.tke_adv_opt, c.tke_upper_bound = 1, 1000.0
    c.ntasks_x = c.ntasks_y = 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@dataclass
class Case:
    name: str
    dims: Dims
    config: Config
    fields: dict = field(default_factory=dict)   # name -> np.float32 array (j,k,i) / (j,i) / (k,)

    @property
    def shape3(self):
        d = self.dims
        return (d.jme - d.jms + 1, d.kme - d.kms + 1, d.ime - d.ims + 1)

    @property
    def mass_points(self):
        d = self.dims
        return (d.ide - d.ids) * (d.jde - d.jds) * (d.kde - d.kds)


# ----------------------------------------------------------------------------
def interp_0(v_in, z_in, z_out):
    """dyn_em/module_init_utilities.F:5-91: linear interpolation with linear extrapolation;
    vectorised over z_out."""
    v_in = np.asarray(v_in, dtype=F)
    z_in = np.asarray(z_in, dtype=F)
    z_out = np.asarray(z_out, dtype=F)
    if z_in[-1] < z_in[0]:
        v_in, z_in = v_in[::-1], z_in[::-1]
    kp = np.clip(np.searchsorted(z_in, z_out, side="left"), 1, len(z_in) - 1)
    w2 = ((z_out - z_in[kp]) / (z_in[kp - 1] - z_in[kp])).astype(F)
    w1 = (F(1.0) - w2).astype(F)
    return (w1 * v_in[kp] + w2 * v_in[kp - 1]).astype(F)


def wk_sounding(zmax=30000.0, dz=250.0, qv0=14.0e-3, u_scale=1.0):
    """Weisman-Klemp (1982) analytic sounding + quarter-circle shear."""
    h = np.concatenate([[10.0, 35.0, 100.0, 200.0], np.arange(375.0, zmax + dz, dz)]).astype(np.float64)
    ztr, thtr, ttr, th0 = 12000.0, 343.0, 213.0, 300.0
    th = np.where(h <= ztr, th0 + (thtr - th0) * (h / ztr) ** 1.25,
                  thtr * np.exp(9.81 / (1004.0 * ttr) * (h - ztr)))
    rh = np.where(h <= ztr, 1.0 - 0.75 * (h / ztr) ** 1.25, 0.25)
    # qv from rh needs p,T: iterate once on a dry-hydrostatic guess
    p = 1.0e5 * (1.0 - 9.81 * h / (1004.0 * 300.0)) ** (1004.0 / 287.0)
    for _ in range(3):
        T = th * (p / 1.0e5) ** (287.0 / 1004.0)
        with np.errstate(over="ignore"):   # the top of the 30 km table is far below 29.65 K-offset validity; qv is capped below
            es = 611.2 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
        qv = np.minimum(qv0, rh * 0.622 * es / np.maximum(p - es, 1.0))
        thv = th * (1.0 + 0.61 * qv)
        pi = np.empty_like(h)
        pi[0] = 1.0 - 9.81 * h[0] / (1004.0 * thv[0])
        for k in range(1, len(h)):
            pi[k] = pi[k - 1] - 9.81 * (h[k] - h[k - 1]) / (1004.0 * 0.5 * (thv[k] + thv[k - 1]))
        p = 1.0e5 * pi ** (1004.0 / 287.0)
    # quarter-circle hodograph over the lowest 2 km, then westerly shear to 6 km
    us = 12.0 * u_scale
    ang = np.clip(h / 2000.0, 0, 1) * (math.pi / 2)
    u = np.where(h <= 2000.0, us * (1 - np.cos(ang)) - 8.0, us - 8.0 + np.clip((h - 2000.0) / 4000.0, 0, 1) * 19.0)
    v = np.where(h <= 2000.0, us * np.sin(ang) - 5.0, us - 5.0)
    return dict(ps=1000.0, ts=300.0, qvs=qv0 * 1e3, h=h.astype(F), th=th.astype(F), qv=(qv * 1e3).astype(F),
                u=u.astype(F), v=v.astype(F))


def const_n_sounding(zmax=36000.0, dz=60.0, N=0.01, U=10.0):
    h = np.arange(dz, zmax + dz, dz, dtype=np.float64)
    th = 300.0 * np.exp(N * N * h / 9.81)
    return dict(ps=1000.0, ts=300.0, qvs=0.0, h=h.astype(F), th=th.astype(F), qv=np.zeros_like(h, dtype=F),
                u=np.full_like(h, U, dtype=F), v=np.zeros_like(h, dtype=F))


def les_sounding(zmax=3000.0, dz=50.0):
    h = np.arange(dz * 0.5, zmax + dz, dz, dtype=np.float64)
    th = np.where(h < 1000.0, 300.0, 300.0 + 8.0 * np.clip((h - 1000.0) / 150.0, 0, 1) + 0.003 * np.maximum(h - 1150.0, 0))
    return dict(ps=1000.0, ts=300.0, qvs=0.0, h=h.astype(F), th=th.astype(F), qv=np.zeros_like(h, dtype=F),
                u=np.full_like(h, 10.0, dtype=F), v=np.zeros_like(h, dtype=F))


def get_sounding(snd, dry):
    """module_initialize_ideal.F:1767-1925."""
    h, th = snd["h"].astype(F), snd["th"].astype(F)
    qv = (np.zeros_like(h) if dry else snd["qv"].astype(F) * F(0.001)).astype(F)
    nl = len(h)
    p_surf = F(100.0 * snd["ps"])
    th_surf = F(snd["ts"])
    qvf = F(1.0) + rvovrd * qv[0]
    rho_surf = F(1.0) / ((r_d / p1000mb) * th_surf * qvf * ((p_surf / p1000mb) ** cvpm))
    rho = np.zeros(nl, F)
    pm = np.zeros(nl, F)
    qvf1 = F(1.0) + qv[0]
    rho[0] = rho_surf
    dz = h[0]
    for _ in range(10):
        pm[0] = p_surf - F(0.5) * dz * (rho_surf + rho[0]) * g * qvf1
        rho[0] = F(1.0) / ((r_d / p1000mb) * th[0] * qvf * ((pm[0] / p1000mb) ** cvpm))
    for k in range(1, nl):
        rho[k] = rho[k - 1]
        dz = h[k] - h[k - 1]
        qvf1 = F(0.5) * (F(2.0) + (qv[k - 1] + qv[k]))
        qvf = F(1.0) + rvovrd * qv[k]
        for _ in range(10):
            pm[k] = pm[k - 1] - F(0.5) * dz * (rho[k] + rho[k - 1]) * g * qvf1
            rho[k] = F(1.0) / ((r_d / p1000mb) * th[k] * qvf * ((pm[k] / p1000mb) ** cvpm))
    pd = np.zeros(nl, F)
    pd[nl - 1] = pm[nl - 1]
    for k in range(nl - 2, -1, -1):
        dz = h[k + 1] - h[k]
        pd[k] = pd[k + 1] + F(0.5) * dz * (rho[k] + rho[k + 1]) * g
    return dict(zk=h, p=pm, pd=pd, theta=th, rho=rho, u=snd["u"].astype(F), v=snd["v"].astype(F), qv=qv)


def memory_dims(nx_stag, ny_stag, nz_stag, halo=5):
    """1-task dims: ids=1..ide=e_we; memory = domain +- (bdyzone+1) as
    external/RSL_LITE/module_dm.F:1174-1201 gives for a single patch."""
    d = Dims()
    d.ids, d.ide, d.jds, d.jde, d.kds, d.kde = 1, nx_stag, 1, ny_stag, 1, nz_stag
    d.ims, d.ime, d.jms, d.jme, d.kms, d.kme = 1 - halo, nx_stag + halo, 1 - halo, ny_stag + halo, 1, nz_stag
    d.ips, d.ipe, d.jps, d.jpe, d.kps, d.kpe = 1, nx_stag, 1, ny_stag, 1, nz_stag
    return d


def vertical_coords(kde, ztop, stretch=True, exp=np.exp):
    """module_initialize_ideal.F:617-732.  exp: the real*4 EXP to use (shipped.py passes the host libm's expf, what ideal.exe calls)"""
    z_scale = F(F(8000.0) / F(ztop))
    k = np.arange(kde, dtype=F)
    if stretch:
        znw = ((exp(-(k) / F(kde - 1) / z_scale) - exp(F(-1.0) / z_scale)) / (F(1.0) - exp(F(-1.0) / z_scale))).astype(F)
    else:
        znw = (F(1.0) - k / F(kde - 1)).astype(F)
    znw[0], znw[-1] = F(1.0), F(0.0)
    o = {n: np.zeros(kde, F) for n in ("znu", "dnw", "rdnw", "dn", "rdn", "fnp", "fnm")}
    o["znw"] = znw
    o["dnw"][:-1] = znw[1:] - znw[:-1]
    o["rdnw"][:-1] = F(1.0) / o["dnw"][:-1]
    o["znu"][:-1] = F(0.5) * (znw[1:] + znw[:-1])
    dnw, dn = o["dnw"], o["dn"]
    dn[1:-1] = F(0.5) * (dnw[1:-1] + dnw[:-2])
    o["rdn"][1:-1] = F(1.0) / dn[1:-1]
    o["fnp"][1:-1] = F(0.5) * dnw[1:-1] / dn[1:-1]
    o["fnm"][1:-1] = F(0.5) * dnw[:-2] / dn[1:-1]
    cof1 = (F(2.0) * dn[1] + dn[2]) / (dn[1] + dn[2]) * dnw[0] / dn[1]
    cof2 = dn[1] / (dn[1] + dn[2]) * dnw[0] / dn[2]
    sc = dict(cf1=F(o["fnp"][1] + cof1), cf2=F(o["fnm"][1] - cof1 - cof2), cf3=F(cof2),
              cfn=F((F(0.5) * dnw[kde - 2] + dn[kde - 2]) / dn[kde - 2]), cfn1=F(F(-0.5) * dnw[kde - 2] / dn[kde - 2]))
    return o, sc


def hybrid_coefs(v, p_top, hybrid_opt, etac=0.2, c3f=None):
    """module_initialize_ideal.F:791-877.  c3f: the full-level B(eta) computed by the caller (shipped.py evaluates the cubic with
    gfortran's multiply sequences for x**2, x**3, x**4)"""
    znw, znu = v["znw"], v["znu"]
    kde = len(znw)
    if c3f is not None:
        c3f = np.asarray(c3f, dtype=F)
    elif hybrid_opt in (0, 1):
        c3f = znw.copy()
    else:
        e = F(etac)
        B1 = F(2.0) * e ** 2 * (F(1.0) - e)
        B2 = -e * (F(4.0) - F(3.0) * e - e ** 3)
        B3 = F(2.0) * (F(1.0) - e ** 3)
        B4 = -(F(1.0) - e ** 2)
        B5 = (F(1.0) - e) ** 4
        c3f = ((B1 + B2 * znw + B3 * znw ** 2 + B4 * znw ** 3) / B5).astype(F)
        c3f[znw < e] = 0
        c3f[0], c3f[-1] = 1, 0
    pdiff = F(p1000mb - F(p_top))
    o = {n: np.zeros(kde, F) for n in ("c3h", "c4h", "c1f", "c2f", "c1h", "c2h")}
    o["c3f"] = c3f.astype(F)
    o["c4f"] = ((znw - c3f) * pdiff).astype(F)
    o["c3h"][:-1] = (c3f[1:] + c3f[:-1]) * F(0.5)
    o["c4h"][:-1] = (znu[:-1] - o["c3h"][:-1]) * pdiff
    o["c1f"][1:-1] = (o["c3h"][1:-1] - o["c3h"][:-2]) / (znu[1:-1] - znu[:-2])
    o["c1f"][0] = 1
    o["c1f"][-1] = 1 if hybrid_opt in (0, 1) else 0
    o["c2f"][:] = (F(1.0) - o["c1f"]) * pdiff
    o["c1h"][:-1] = (c3f[1:] - c3f[:-1]) / (znw[1:] - znw[:-1])
    o["c2h"][:-1] = (F(1.0) - o["c1h"][:-1]) * pdiff
    return o


def _columns(ht, v, hy, p_top, dry_s, moist_s, use_moist, hypsometric_opt):
    """Base state + balanced full state for an array of terrain heights `ht` (any shape).
    Returns dict of arrays with a trailing/embedded k axis at position -2 ... shaped (..., k)."""
    kde = len(v["znw"])
    nk = kde
    shp = ht.shape
    p_surf = interp_0(dry_s["p"], dry_s["zk"], ht)
    mub = (p_surf - F(p_top)).astype(F)
    c3h, c4h, c1h, c2h, c1f, c2f = hy["c3h"], hy["c4h"], hy["c1h"], hy["c2h"], hy["c1f"], hy["c2f"]
    dnw, rdnw, rdn = v["dnw"], v["rdnw"], v["rdn"]
    pb = np.zeros(shp + (nk,), F)
    t_init = np.zeros_like(pb)
    alb = np.zeros_like(pb)
    phb = np.zeros_like(pb)
    for k in range(kde - 1):
        pl = (c3h[k] * (p_surf - F(p_top)) + c4h[k] + F(p_top)).astype(F)
        pb[..., k] = pl
        t_init[..., k] = interp_0(dry_s["theta"], dry_s["p"], pl) - t0
        alb[..., k] = (r_d / p1000mb) * (t_init[..., k] + t0) * (pl / p1000mb) ** cvpm
    phb[..., 0] = g * ht
    for kk in range(1, kde):
        k = kk - 1
        phb[..., kk] = phb[..., kk - 1] - dnw[kk - 1] * (c1h[k] * mub + c2h[k]) * alb[..., kk - 1]
    s = moist_s if use_moist else dry_s
    pd_surf = interp_0(s["pd"], s["zk"], ht)
    mu = (pd_surf - F(p_top) - mub).astype(F)
    qv = np.zeros_like(pb)
    t = np.zeros_like(pb)
    for k in range(kde - 1):
        pl = (c3h[k] * (pd_surf - F(p_top)) + c4h[k] + F(p_top)).astype(F)
        qv[..., k] = interp_0(s["qv"], s["pd"], pl)
        t[..., k] = interp_0(s["theta"], s["pd"], pl) - t0
    return dict(mub=mub, pb=pb, t_init=t_init, alb=alb, phb=phb, mu=mu, qv=qv, t=t, p_surf=p_surf)


def balance(col, v, hy, p_top, hypsometric_opt):
    """Hydrostatic p (from the top down, :1027-1050), alt/al from the EOS, and ph
    consistent with the model's own diagnosis of al (calc_p_rho_phi)."""
    kde = len(v["znw"])
    c1h, c2h, c1f, c2f, c3f, c4f, c3h, c4h = (hy[n] for n in ("c1h", "c2h", "c1f", "c2f", "c3f", "c4f", "c3h", "c4h"))
    dnw, rdnw, rdn = v["dnw"], v["rdnw"], v["rdn"]
    qv, t, mu, mub, pb, alb = col["qv"], col["t"], col["mu"], col["mub"], col["pb"], col["alb"]
    p = np.zeros_like(pb)
    kk = kde - 2
    k = kk + 1
    qvf1 = F(0.5) * (qv[..., kk] + qv[..., kk])
    qvf2 = F(1.0) / (F(1.0) + qvf1)
    qvf1 = qvf1 * qvf2
    p[..., kk] = -F(0.5) * ((c1f[k] * mu) + qvf1 * (c1f[k] * mub + c2f[k])) / rdnw[kk] / qvf2
    for kk in range(kde - 3, -1, -1):
        k = kk + 1
        qvf1 = F(0.5) * (qv[..., kk] + qv[..., kk + 1])
        qvf2 = F(1.0) / (F(1.0) + qvf1)
        qvf1 = qvf1 * qvf2
        p[..., kk] = p[..., kk + 1] - ((c1f[k] * mu) + qvf1 * (c1f[k] * mub + c2f[k])) / qvf2 / rdn[kk + 1]
    qvf = F(1.0) + rvovrd * qv
    with np.errstate(divide="ignore", invalid="ignore"):   # level kde holds no mass point; zeroed next
        alt = ((r_d / p1000mb) * (t + t0) * qvf * (((p + pb) / p1000mb) ** cvpm)).astype(F)
    alt[..., kde - 1] = 0
    al = (alt - alb).astype(F)
    al[..., kde - 1] = 0
    ph = np.zeros_like(pb)
    muts = (mub + mu).astype(F)
    if hypsometric_opt == 1:
        for kk in range(1, kde):
            k = kk - 1
            ph[..., kk] = ph[..., kk - 1] - dnw[kk - 1] * (((c1h[k] * mub + c2h[k]) + (c1h[k] * mu)) * al[..., kk - 1] +
                                                          (c1h[k] * mu) * alb[..., kk - 1])
    else:
        phf = col["phb"][..., 0].copy()
        for k in range(1, kde):
            pfu = c3f[k] * muts + c4f[k] + F(p_top)
            pfd = c3f[k - 1] * muts + c4f[k - 1] + F(p_top)
            phm = c3h[k - 1] * muts + c4h[k - 1] + F(p_top)
            phf = (phf + alt[..., k - 1] * phm * np.log(pfd / pfu)).astype(F)
            ph[..., k] = phf - col["phb"][..., k]
    col.update(p=p.astype(F), alt=alt, al=al, ph=ph.astype(F))
    return col


def _hash_noise(shape, seed=1234):
    """Deterministic uniform [-1,1) noise (small-amplitude noise on theta)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        x = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xFF51AFD7ED558CCD)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xC4CEB9FE1A85EC53)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(33)
    return ((x >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0).astype(F).reshape(shape)


def make_case(name="quarter_ss", nx=None, ny=None, nz=None, **over) -> Case:
    """Build one of the BASELINE.json configs.  nx,ny,nz are STAGGERED dims (e_we,e_sn,e_vert).
    Keyword overrides go into the wrfgpu_config (e.g. moist_adv_opt=0, periodic=True)."""
    periodic = over.pop("periodic", None)
    noise = over.pop("noise", None)
    seed = over.pop("seed", 1234)
    bubble_lattice = over.pop("bubble_lattice", 0.0)
    u_scale = over.pop("u_scale", 1.0)
    calm = over.pop("calm", False)      # no environmental wind (needed next to symmetric walls)
    if name == "quarter_ss":
        nx, ny, nz = nx or 42, ny or 42, nz or 41
        c = default_config(dt=12.0, dx=2000.0, dy=2000.0, time_step_sound=6, time_step=12, damp_opt=2, zdamp=5000.0,
                           dampcoef=0.003, mp_physics=1, num_moist=3, open_xs=1, open_xe=1, open_ys=1, open_ye=1)
        ztop, hybrid_opt, snd, delt, stretch = 20000.0, 0, wk_sounding(u_scale=u_scale), 3.0, True
    elif name == "hill2d_x":
        nx, ny, nz = nx or 202, ny or 3, nz or 41
        c = default_config(dt=20.0, dx=2000.0, dy=2000.0, time_step_sound=6, time_step=20, damp_opt=3, zdamp=20000.0,
                           dampcoef=0.1, mp_physics=0, num_moist=1, open_xs=1, open_xe=1, periodic_y=1)
        ztop, hybrid_opt, snd, delt, stretch = 30000.0, 2, const_n_sounding(), 0.0, True
    elif name == "b_wave":
        nx, ny, nz = nx or 41, ny or 81, nz or 65
        c = default_config(dt=600.0, dx=100000.0, dy=100000.0, time_step_sound=4, time_step=600, damp_opt=0,
                           mp_physics=0, num_moist=1, periodic_x=1, symmetric_ys=1, symmetric_ye=1)
        ztop, hybrid_opt, snd, delt, stretch = 16000.0, 0, wk_sounding(zmax=20000.0, qv0=0.0, u_scale=0.0), 0.0, True
    elif name == "les":
        nx, ny, nz = nx or 40, ny or 40, nz or 40
        c = default_config(dt=1.0, dx=100.0, dy=100.0, time_step_sound=6, time_step=1, damp_opt=0, mp_physics=0,
                           num_moist=1, num_scalar=over.pop("num_scalar", 2), periodic_x=1, periodic_y=1, pert_coriolis=1)
        ztop, hybrid_opt, snd, delt, stretch = 2000.0, 0, les_sounding(), 3.0, False
    elif name == "synthetic":
        nx, ny, nz = nx or 2000, ny or 2000, nz or 80
        per = True if periodic is None else periodic
        c = default_config(dt=12.0, dx=2000.0, dy=2000.0, time_step_sound=6, time_step=12, damp_opt=0, mp_physics=0,
                           num_moist=1)
        if per:
            c.periodic_x = c.periodic_y = 1
        else:
            c.open_xs = c.open_xe = c.open_ys = c.open_ye = 1
        ztop, hybrid_opt, snd, delt, stretch = 20000.0, 0, wk_sounding(u_scale=u_scale), 3.0, True
        noise = 1.0e-3 if noise is None else noise
        bubble_lattice = bubble_lattice or 100000.0
    else:
        raise ValueError(name)
    hybrid_opt = over.pop("hybrid_opt", hybrid_opt)
    for k, v_ in over.items():
        setattr(c, k, v_)
    if c.mp_physics == 0:
        c.num_moist = max(1, min(c.num_moist, 1)) if c.num_moist else 0
    d = memory_dims(nx, ny, nz)
    v, sc = vertical_coords(nz, ztop, stretch)
    for k_, val in sc.items():
        setattr(c, k_, float(val))
    dry_s = get_sounding(snd, dry=True)
    moist_s = get_sounding(snd, dry=False)
    p_top = float(interp_0(dry_s["p"], dry_s["zk"], F(ztop)))
    c.p_top = p_top
    hy = hybrid_coefs(v, p_top, hybrid_opt)

    nj, nk, ni = d.jme - d.jms + 1, nz, d.ime - d.ims + 1
    # terrain over the full memory range (so that base-state halos are exact)
    ii = np.arange(d.ims, d.ime + 1)
    ht1 = np.zeros(ni, F)
    if name == "hill2d_x":
        icm = d.ide // 2
        ht1 = (F(100.0) / (F(1.0) + ((ii - icm).astype(F) / F(5.0)) ** 2)).astype(F)
    col = _columns(ht1, v, hy, p_top, dry_s, moist_s, True, c.hypsometric_opt)   # (ni, k)
    if c.num_moist == 0:
        col["qv"][:] = 0

    def bcast(a):   # (ni,k) -> (nj,k,ni)
        return np.ascontiguousarray(np.broadcast_to(a.T[None, :, :], (nj, nk, ni))).astype(F)

    fl = {}
    T = bcast(col["t"])
    QV = bcast(col["qv"])
    PB = bcast(col["pb"])
    ALB = bcast(col["alb"])
    PHB = bcast(col["phb"])
    MUB = np.ascontiguousarray(np.broadcast_to(col["mub"][None, :], (nj, ni))).astype(F)
    MU = np.ascontiguousarray(np.broadcast_to(col["mu"][None, :], (nj, ni))).astype(F)
    HT = np.ascontiguousarray(np.broadcast_to(ht1[None, :], (nj, ni))).astype(F)
    # perturbation: bubble(s) + noise on theta (then rebalanced below)
    jj = np.arange(d.jms, d.jme + 1)
    colb = balance(dict(col), v, hy, p_top, c.hypsometric_opt)
    zmid = (0.5 * (colb["ph"][:, :-1] + colb["ph"][:, 1:] + col["phb"][:, :-1] + col["phb"][:, 1:]) / g).astype(F)  # (ni,k-1)
    if delt != 0.0:
        nxc, nyc = (d.ide - d.ids) // 2, (d.jde - d.jds) // 2
        if bubble_lattice:
            lx = max(1, int(round(bubble_lattice / c.dx)))
            xr = ((((ii - nxc) % lx) + lx // 2) % lx - lx // 2).astype(F) * F(c.dx) / F(10000.0)
            yr = ((((jj - nyc) % lx) + lx // 2) % lx - lx // 2).astype(F) * F(c.dy) / F(10000.0)
        else:
            xr = (F(c.dx) * (ii - nxc).astype(F) / F(10000.0)) * (F(10000.0 / 10000.0) if name != "les" else F(10000.0 / 500.0))
            yr = (F(c.dy) * (jj - nyc).astype(F) / F(10000.0)) * (F(1.0) if name != "les" else F(10000.0 / 500.0))
        zc, zr = (1500.0, 1500.0) if name != "les" else (400.0, 300.0)
        zrad = ((zmid - F(zc)) / F(zr)).T  # (k-1, ni)
        rad = np.sqrt(xr[None, None, :] ** 2 + yr[:, None, None] ** 2 + zrad[None, :, :] ** 2).astype(F)
        bub = np.where(rad <= 1.0, F(delt) * np.cos(F(0.5 * math.pi) * rad) ** 2, F(0.0)).astype(F)
        T[:, :-1, :] += bub
    if noise:
        T[:, :-1, :] += F(noise) * _hash_noise((nj, nk - 1, ni), seed)
    # rebalance every column with the perturbed theta (vectorised over (j,i))
    full = dict(qv=np.moveaxis(QV, 1, -1), t=np.moveaxis(T, 1, -1), mu=MU, mub=MUB, pb=np.moveaxis(PB, 1, -1),
                alb=np.moveaxis(ALB, 1, -1), phb=np.moveaxis(PHB, 1, -1))
    full = balance(full, v, hy, p_top, c.hypsometric_opt)
    back = lambda a: np.ascontiguousarray(np.moveaxis(a, -1, 1)).astype(F)  # noqa: E731
    P, AL, ALT, PH = back(full["p"]), back(full["al"]), back(full["alt"]), back(full["ph"])

    # winds (:1500-1555): interpolate on the (moist) sounding pressure at the level pressure
    def wind(prof, ht_at):
        ps = interp_0(moist_s["p"], moist_s["zk"], ht_at)
        out = np.zeros((len(ht_at), nk), F)
        for k in range(nz):
            pl = (hy["c3h"][k] * (ps - F(p_top)) + hy["c4h"][k] + F(p_top)).astype(F)
            out[:, k] = interp_0(prof, moist_s["p"], pl)
        return out
    ht_u = ht1.copy()
    ht_u[1:] = F(0.5) * (ht1[1:] + ht1[:-1])
    U = bcast(wind(moist_s["u"], ht_u))
    V = bcast(wind(moist_s["v"], ht1))
    if name == "b_wave":
        # analytic jet: u(y,z) = u0 sin^2(pi y/Ly) * z/zt below the tropopause, f-plane
        yy = (jj - d.jds + 0.5) / float(d.jde - d.jds)
        zprof = np.clip(zmid.T / 12000.0, 0, 1.2)            # (k-1, ni)
        jet = (35.0 * np.sin(math.pi * np.clip(yy, 0, 1)) ** 2)[:, None, None] * (zprof * np.exp(1 - zprof))[None, :, :]
        U[:, :-1, :] = jet.astype(F)
        V[:] = 0
    if calm:
        U[:] = 0
        V[:] = 0
    U[:, nz - 1, :] = 0
    V[:, nz - 1, :] = 0
    base_i = 0 - d.ims + d.ids  # column (ids, jds)
    b1 = {n: np.zeros(nz, F) for n in ("u_base", "v_base", "t_base", "qv_base", "z_base")}
    b1["t_base"][:-1] = T[base_i, :-1, base_i]
    b1["qv_base"][:-1] = QV[base_i, :-1, base_i]
    b1["u_base"][:-1] = U[base_i, :-1, base_i]
    b1["v_base"][:-1] = V[base_i, :-1, base_i]
    b1["z_base"][:-1] = F(0.5) * (PHB[base_i, :-1, base_i] + PHB[base_i, 1:, base_i] + PH[base_i, :-1, base_i] + PH[base_i, 1:, base_i]) / g
    # moist theta (:1726)
    if c.use_theta_m == 1 and c.num_moist >= 1:
        T[:, :-1, :] = (T[:, :-1, :] + t0) * (F(1.0) + (r_v / r_d) * QV[:, :-1, :]) - t0
    T[:, nz - 1, :] = 0
    for n, a in (("u", U), ("v", V), ("t", T), ("ph", PH)):
        fl[n + "_1"] = a.copy()
        fl[n + "_2"] = a.copy()
    fl["w_1"] = np.zeros_like(T)
    fl["w_2"] = np.zeros_like(T)
    fl["mu_1"], fl["mu_2"], fl["mub"], fl["ht"] = MU.copy(), MU.copy(), MUB, HT
    fl["phb"], fl["pb"], fl["alb"], fl["p"], fl["al"], fl["alt"] = PHB, PB, ALB, P, AL, ALT
    fl["t_init"] = bcast(col["t_init"])      # dry base-state theta - t0 (module_initialize_ideal.F), the base_3d of the 3dmp diffusion
    if c.num_moist >= 1:
        fl["moist_qv"] = QV
    ones = np.ones((nj, ni), F)
    for n in ("msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy"):
        fl[n] = ones.copy()
    fcor = 1.0e-4 if name in ("b_wave", "les") else 0.0
    fl["f"] = np.full((nj, ni), fcor, F)
    fl["e"] = np.zeros((nj, ni), F)
    fl["sina"] = np.zeros((nj, ni), F)
    fl["cosa"] = np.ones((nj, ni), F)
    for n in ("znu", "znw", "fnm", "fnp", "rdnw", "rdn", "dnw", "dn"):
        fl[n] = v[n]
    for n in ("c1h", "c2h", "c1f", "c2f", "c3h", "c4h", "c3f", "c4f"):
        fl[n] = hy[n]
    fl.update(b1)
    # passive scalars: top-hats a la module_advect_em.F:56-80 (tophat)
    for s in range(1, c.num_scalar + 1):
        a = np.zeros_like(T)
        i0 = (d.ide // 4) * s % max(1, d.ide - 12) - d.ims + 3
        a[max(0, -d.jms + 3):, 2:8 + s, i0:i0 + 8] = 1.0
        a[:, nz - 1, :] = 0
        fl[f"scalar_{s}"] = a
    return Case(name=name, dims=d, config=c, fields={k: np.ascontiguousarray(a, dtype=F) for k, a in fl.items()})
