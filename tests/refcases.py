"""Configurations and helpers shared by tests/test_reference_source.py (which runs the reference's own Fortran here and needs
/root/reference) and tests/test_reference_source_golden.py (which replays what those runs established from a committed fixture and
runs anywhere).  Nothing in this file touches the reference tree."""
import numpy as np

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import decomp, make_case

f32 = np.float32


def small_case(name="quarter_ss", **kw):
    return parity.with_map_factors(make_case(name, nx=12, ny=10, nz=9, **kw))


# The configurations of tests/parity.py's catalogue at sizes the translated Fortran (a Python loop nest) finishes in a second or two
Q = dict(nx=12, ny=10, nz=9)
S = dict(nx=14, ny=12, nz=9, bubble_lattice=8000.0, noise=0.05)
H = dict(nx=20, ny=3, nz=9)
CASES = {
    "small": lambda: small_case(),
    "open_mapfac": lambda: parity.with_map_factors(make_case("quarter_ss", w_damping=1, **Q)),
    "b_wave_lid": lambda: make_case("b_wave", nx=9, ny=13, nz=9, top_lid=1, phi_adv_z=2, w_damping=1, hypsometric_opt=1),
    "opts": lambda: make_case("quarter_ss", use_theta_m=0, hypsometric_opt=1, damp_opt=3, zdamp=8000.0, dampcoef=0.2, scalar_adv_opt=0,
                              num_scalar=1, u_scale=1.5, **Q),
    "periodic_mapfac": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=1, **S)),
    "rk2_sym_x": lambda: make_case("synthetic", periodic=False, calm=True, rk_ord=2, open_xs=0, open_xe=0, symmetric_xs=1, symmetric_xe=1,
                                   time_step_sound=4, **S),
    "diffusion": lambda: make_case("quarter_ss", diff_opt=1, km_opt=1, khdif=500.0, kvdif=100.0, **Q),
    "diffusion_periodic": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=1, diff_opt=1, km_opt=1, khdif=300.0, kvdif=50.0, **S)),
    "hill": lambda: make_case("hill2d_x", **H),
    "order64": lambda: make_case("quarter_ss", h_mom_adv_order=6, v_mom_adv_order=4, h_sca_adv_order=6, v_sca_adv_order=4, **Q),
    "weno": lambda: make_case("quarter_ss", momentum_adv_opt=3, moist_adv_opt=3, **Q),
    "mono": lambda: make_case("quarter_ss", moist_adv_opt=2, **Q),
    "mono_periodic": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=2, scalar_adv_opt=2, moist_adv_opt=2, **S)),
    "wenopd": lambda: make_case("quarter_ss", moist_adv_opt=4, **Q),
    "wenopd_periodic": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=2, scalar_adv_opt=4, moist_adv_opt=3, **S)),
    "smag2d": lambda: make_case("quarter_ss", diff_opt=1, km_opt=4, **Q),
    "smag2d_hill": lambda: make_case("hill2d_x", diff_opt=1, km_opt=4, mix_full_fields=1, **H),
    "diff2": lambda: make_case("quarter_ss", diff_opt=2, km_opt=1, khdif=300.0, kvdif=100.0, tke_drag_coefficient=0.003, tke_heat_flux=0.05, **Q),
    "diff2_periodic": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=2, diff_opt=2, km_opt=4, c_s=0.5, mix_full_fields=1, **S)),
    "diff2_hill": lambda: parity.with_surface_fluxes(make_case("hill2d_x", diff_opt=2, km_opt=1, khdif=100.0, kvdif=30.0, isfflx=1, **H)),
    "smag3d": lambda: parity.with_cloud(parity.with_updraft(make_case("quarter_ss", diff_opt=2, km_opt=3, c_s=0.25, **Q), 30.0)),
    "smag3d_periodic": lambda: parity.with_updraft(parity.with_map_factors(make_case("synthetic", num_scalar=1, diff_opt=2, km_opt=3, mix_isotropic=1,
                                                                                     mix_full_fields=1, **S)), 30.0),
    "tke": lambda: parity.with_tke(parity.with_cloud(parity.with_updraft(make_case("quarter_ss", diff_opt=2, km_opt=2, mix_full_fields=1, **Q), 30.0))),
    "tke_periodic": lambda: parity.with_tke(parity.with_updraft(parity.with_map_factors(make_case(
        "synthetic", num_scalar=1, diff_opt=2, km_opt=2, mix_isotropic=1, tke_adv_opt=2, tke_drag_coefficient=0.003, tke_heat_flux=0.05, **S)), 30.0)),
    "tke_hill": lambda: parity.with_tke(parity.with_surface_fluxes(make_case("hill2d_x", diff_opt=2, km_opt=2, isfflx=1, **H))),
    "specified_tke": lambda: parity.with_boundary_zone(parity.with_tke(parity.with_updraft(make_case(
        "quarter_ss", nx=16, ny=14, nz=9, specified=1, diff_opt=2, km_opt=2, mix_full_fields=1, **parity.SPEC), 20.0))),
    "specified_real": lambda: parity.with_boundary_zone(parity.with_map_factors(make_case(
        "quarter_ss", nx=16, ny=14, nz=9, specified=1, diff_opt=1, km_opt=4, w_damping=1, damp_opt=3, zdamp=8000.0, dampcoef=0.2, num_scalar=1,
        mp_physics=0, num_moist=1, **parity.SPEC))),
    "ieva_wdamp": lambda: parity.with_updraft(make_case("quarter_ss", nx=12, ny=10, nz=17, zadvect_implicit=1, dt=40.0, time_step=40, w_damping=1,
                                                        w_crit_cfl=2.0, use_adaptive_time_step=1), 80.0, conv=60.0),
    "adaptive": lambda: parity.with_updraft(parity.with_map_factors(make_case("quarter_ss", w_damping=1, use_adaptive_time_step=1, **Q)), 20.0),
    "diff6_periodic": lambda: parity.with_map_factors(make_case("synthetic", num_scalar=2, diff_6th_opt=1, diff_6th_factor=0.2, scalar_mix6_off=1, **S)),
    "order_mixed": lambda: make_case("b_wave", nx=9, ny=13, nz=9, h_mom_adv_order=6, v_sca_adv_order=4, moist_adv_opt=4, momentum_adv_opt=3),
    "rk2_everything": lambda: parity.with_updraft(make_case("synthetic", nx=14, ny=12, nz=17, periodic=False, bubble_lattice=8000.0, noise=0.05, rk_ord=2,
                                                            time_step_sound=4, zadvect_implicit=1, momentum_adv_opt=3, moist_adv_opt=3, num_scalar=1,
                                                            scalar_adv_opt=4, dt=20.0, time_step=20, top_lid=1, diff_6th_opt=2, diff_opt=1, km_opt=1,
                                                            khdif=200.0, kvdif=20.0), 40.0, conv=20.0),
    "specified_bcs_moist": lambda: parity.with_boundary_zone(make_case("quarter_ss", nx=16, ny=14, nz=9, specified=1, have_bcs_moist=1, diff_6th_opt=2,
                                                                       h_mom_adv_order=6, v_mom_adv_order=4, h_sca_adv_order=6, v_sca_adv_order=4,
                                                                       moist_adv_opt=0, **parity.SPEC)),
    "smag2d_b_wave": lambda: make_case("b_wave", nx=9, ny=13, nz=9, diff_opt=1, km_opt=4, kvdif=5.0),
    "diff2_b_wave": lambda: make_case("b_wave", nx=9, ny=13, nz=9, diff_opt=2, km_opt=4, isfflx=2, tke_heat_flux=0.02),
    # a strong updraft under the 3-D Smagorinsky closure: two points where glibc's powf(x, 0.5) and sqrtf(x) differ (tmp**0.5 in smag_km)
    "smag3d_pow": lambda: parity.with_updraft(make_case("quarter_ss", nx=12, ny=10, nz=17, diff_opt=2, km_opt=3, khdif=300.0, kvdif=50.0, use_theta_m=0,
                                                        v_sca_adv_order=4, moist_adv_opt=3, momentum_adv_opt=3, damp_opt=0, rk_ord=2, zadvect_implicit=1,
                                                        dt=40.0, time_step=40), 80.0),
    # momentum and scalar orders apart, on a domain with a non-empty interior box: w (scalar orders) through its k-marching kernel next to
    # u, v through the gather kernels, and the other way round
    "w_fast_uv_gather": lambda: parity.with_map_factors(parity.with_updraft(make_case("quarter_ss", nx=20, ny=18, nz=9, h_mom_adv_order=6), 15.0)),
    "uv_fast_w_gather": lambda: parity.with_map_factors(parity.with_updraft(make_case("quarter_ss", nx=20, ny=18, nz=9, h_sca_adv_order=6, v_sca_adv_order=4), 15.0)),
    "les": lambda: parity.with_map_factors(make_case("les", **Q)),          # pert_coriolis, PD tracers, passive qv
    "diff6": lambda: make_case("quarter_ss", diff_6th_opt=2, **Q),
    "rain_open_mapfac": lambda: rainy(CASES["open_mapfac"]()),
    "rain_opts": lambda: rainy(CASES["opts"]()),
    "rain_specified": lambda: rainy(CASES["specified"]()),
    "rain_weno": lambda: rainy(CASES["weno"]()),
    "diff6_hill": lambda: make_case("hill2d_x", diff_6th_opt=2, diff_6th_slopeopt=1, diff_6th_thresh=0.004, **H),
    "ieva": lambda: parity.with_updraft(make_case("quarter_ss", nx=12, ny=10, nz=17, zadvect_implicit=1, dt=40.0, time_step=40), 80.0),
    "backward": lambda: make_case("synthetic", num_scalar=1, scalar_adv_opt=3, momentum_adv_opt=3, time_step=-12, **S),
    "specified": lambda: parity.with_boundary_zone(make_case("quarter_ss", nx=16, ny=14, nz=9, specified=1, **parity.SPEC)),
    "nested": lambda: parity.with_boundary_zone(make_case("quarter_ss", nx=16, ny=14, nz=9, nested=1, num_scalar=1, **parity.SPEC)),
}


def build(name):
    """a fresh Case of the catalogue, remembering its name (the golden fixture refers to cases by name)"""
    case = CASES[name]()
    case.refname = name
    return case


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).hexdigest()[:16]


def rainy(case):
    """cloud above the autoconversion threshold and rain falling through cloudy and through dry air: every branch of Kessler"""
    case = parity.with_cloud(case)
    d = case.dims
    qc = case.fields["moist_qc"]
    qc[qc > 0] = 2.5e-3
    qr = case.fields.setdefault("moist_qr", np.zeros_like(qc))
    qr[:, 2:7, d.ide // 2 - d.ims:] = 1.5e-3
    return case


def at_phase(case, target, rk=None, occurrence=0):
    """an oracle advanced to just before the occurrence-th `target` phase (of Runge-Kutta stage rk) of the first step"""
    o = Oracle(case)
    o.start_domain()
    if case.config.specified or case.config.nested:
        o.set_dtbc(case.config.dt)
    n = 0
    for ph, args in parity.phase_sequence(case):
        if ph == target and (rk is None or (args and args[0] == rk)):
            if n == occurrence:
                return o, args
            n += 1
        o.call(ph, *args)
    raise KeyError(target)


def rk_scalars(c, rk):
    """dt_rk, dts_rk, number_of_small_timesteps of solve_em.F:585-615"""
    nss = c.time_step_sound
    dt = f32(c.dt)
    dts = f32(dt / f32(nss))
    if c.rk_ord == 3 and rk == 1:
        return dict(dt_rk=f32(dt / f32(3.0)), dts_rk=f32(dt / f32(3.0)), number_of_small_timesteps=1)
    if rk < c.rk_ord:
        return dict(dt_rk=f32(f32(0.5) * dt), dts_rk=dts, number_of_small_timesteps=nss // 2)
    return dict(dt_rk=dt, dts_rk=dts, number_of_small_timesteps=nss)


def own(case, name, a):
    """the points of the field the domain owns (the halo is the boundary fills' and exchanges' business)"""
    d = case.dims
    ie = d.ide if name in decomp.XSTAG else d.ide - 1
    je = d.jde if name in decomp.YSTAG else d.jde - 1
    if a.ndim == 1:
        return a
    return a[d.jds - d.jms:je - d.jms + 1, ..., d.ids - d.ims:ie - d.ims + 1]




def oracle_feeder(adv_opt, loops, at=()):
    """rk_scalar_step(1..3) per loop with the periodic fill at the end of the loop only, as the program does -> {loop: column}, row"""
    o = Oracle(parity.tophat_case(adv_opt))
    o.start_domain()
    d = o.case.dims
    row = lambda: np.asarray(o.get("scalar_1"))[2 - d.jms, 0, 1 - d.ims:91 - d.ims + 1]
    cols = {}
    for n in range(1, loops + 1):
        for rk in (1, 2, 3):
            o.call("rk_scalar_step", rk)
        o.bc("scalar_1", "p")
        if n in at:
            cols[n] = [int(x) for x in np.rint((row() * f32(100)).astype(f32))]
    return cols, row().copy()
