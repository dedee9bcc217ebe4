"""Shared machinery of the GPU-vs-oracle parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

from tests.oracle_lib import Oracle
from wrf_b200 import Solver, make_case

# The device evaluates powf/logf/expf/sinf with a restatement of the host libm's own
# algorithm (wrf_b200/csrc/libm_port.h, checked against libm in tests/test_libm_port.py),
# so every phase -- including calc_p_rho_phi, Kessler and the damping profiles -- and every
# whole step is held to BIT equality with the oracle.  The relative-L2 numbers are still
# computed for the reports; the north-star bar (<= 1e-5 after one step) is the fallback
# the tests would apply on a host whose libm is not the glibc >= 2.28 family.
TOL_REL_L2 = 0.0
STEP_TOL_REL_L2 = 1.0e-5


def with_map_factors(case, amp=0.04):
    """Give a case smoothly varying map factors and a full set of Coriolis / rotation fields (f, e, sina,
    cosa).  Every shipped ideal case has msf == 1 and e == 0, which would hide a wrong map factor or a
    dropped curvature term in a kernel; here the seven staggered factors all differ from one another."""
    d = case.dims
    jj, ii = np.meshgrid(np.arange(d.jms, d.jme + 1, dtype=np.float64), np.arange(d.ims, d.ime + 1, dtype=np.float64), indexing="ij")
    nx, ny = float(d.ide - d.ids), float(d.jde - d.jds)

    def m(pi, pj, sx=0.0, sy=0.0):
        return (1.0 + amp * np.sin(2 * np.pi * (ii + sx) / nx * pi + 0.3) * np.cos(2 * np.pi * (jj + sy) / ny * pj + 0.7)).astype(np.float32)

    f = case.fields
    f["msftx"], f["msfty"] = m(1, 1), m(1, 2)
    f["msfux"], f["msfuy"] = m(1, 1, -0.5), m(1, 2, -0.5)
    f["msfvx"], f["msfvy"] = m(1, 1, 0.0, -0.5), m(1, 2, 0.0, -0.5)
    f["msfvx_inv"] = (np.float32(1.0) / f["msfvx"]).astype(np.float32)
    lat = (0.6 + 0.1 * np.sin(2 * np.pi * jj / ny)).astype(np.float64)
    f["f"] = (1.4584e-4 * np.sin(lat)).astype(np.float32)
    f["e"] = (1.4584e-4 * np.cos(lat)).astype(np.float32)
    rot = 0.2 * np.sin(2 * np.pi * ii / nx)
    f["sina"], f["cosa"] = np.sin(rot).astype(np.float32), np.cos(rot).astype(np.float32)
    return case


def with_surface_fluxes(case):
    """ustm, hfx, qfx as a host-side surface-layer scheme would hand them over (isfflx 1 / 2): smooth, non-zero patterns"""
    d = case.dims
    jj, ii = np.meshgrid(np.arange(d.jms, d.jme + 1), np.arange(d.ims, d.ime + 1), indexing="ij")
    case.fields["ustm"] = (0.25 + 0.1 * np.sin(0.37 * ii + 0.9 * jj)).astype(np.float32)
    case.fields["hfx"] = (80.0 + 40.0 * np.cos(0.21 * ii - 0.5 * jj)).astype(np.float32)
    case.fields["qfx"] = (2.0e-5 * (1.0 + 0.5 * np.sin(0.13 * ii + 0.4 * jj))).astype(np.float32)
    return case


def shipped_case(name):
    import os
    from wrf_b200 import shipped
    return shipped.init_domain(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", name))


def with_tke(case):
    """a smooth, positive initial tke (both time levels) so that every term of tke_rhs acts from the first step"""
    d, c = case.dims, case.config
    kk, jj, ii = np.meshgrid(np.arange(d.kms, d.kme + 1), np.arange(d.jms, d.jme + 1), np.arange(d.ims, d.ime + 1), indexing="ij")
    a = (0.6 * np.exp(-kk / 6.0) * (1.0 + 0.5 * np.sin(0.41 * ii + 0.23 * jj))).astype(np.float32).transpose(1, 0, 2)
    a[:, -1, :] = 0
    n = c.num_scalar + 1
    case.fields["scalar_%d" % n] = a.copy()
    case.fields["scalar_old_%d" % n] = a.copy()
    return case


def with_cloud(case):
    """cloud water above the threshold of calculate_N2's saturated branch (qc >= 1e-5) in a box of the domain"""
    d = case.dims
    qc = case.fields.setdefault("moist_qc", np.zeros_like(case.fields["moist_qv"]))
    j0, i0 = (d.jde // 3) - d.jms, (d.ide // 3) - d.ims
    qc[j0:j0 + d.jde // 3, 1:9, i0:i0 + d.ide // 3] = 2.0e-4
    return case


def with_updraft(case, wmax, conv=25.0):
    """Add an updraft core to a case: a Gaussian blob of `wmax` m/s in w and, because the first step diagnoses ww from
    the divergence of (u, v) and not from w, winds converging below and diverging aloft (`conv` m/s) that imply the same
    kind of core.  The vertical Courant number then exceeds IEVA's threshold from the first step on, so that WW_SPLIT
    hands part of ww to the implicit solver (wwI /= 0)."""
    d = case.dims
    nz = d.kde
    jj = np.arange(d.jms, d.jme + 1, dtype=np.float64)[:, None, None]
    ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)[None, None, :]
    kk = np.arange(case.fields["w_2"].shape[1], dtype=np.float64)[None, :, None]
    ic, jc = 0.5 * (d.ids + d.ide), 0.5 * (d.jds + d.jde)
    xh, yh = (ii - ic) / 4.0, (jj - jc) / 4.0
    blob = (wmax * np.exp(-(xh ** 2 + yh ** 2 + ((kk - nz / 3.0) / (nz / 5.0)) ** 2))).astype(np.float32)
    blob[:, 0, :] = 0
    blob[:, nz - 1:, :] = 0
    prof = np.cos(np.pi * (kk + 0.5) / (nz - 1))
    prof[:, nz - 1:, :] = 0
    up = (-conv * (xh - 0.125) * np.exp(-((xh - 0.125) ** 2 + yh ** 2)) * prof).astype(np.float32)     # u points sit half a cell west
    vp = (-conv * (yh - 0.125) * np.exp(-(xh ** 2 + (yh - 0.125) ** 2)) * prof).astype(np.float32)
    if d.jde - d.jds < 4:      # 2-D (x, z) cases keep v = 0 and no y variation
        blob = np.broadcast_to(blob[blob.shape[0] // 2][None], blob.shape).copy()
        up = np.broadcast_to((-conv * (xh - 0.125) * np.exp(-((xh - 0.125) ** 2)) * prof).astype(np.float32), up.shape).copy()
        vp = np.zeros_like(vp)
    for n, a in (("w_2", blob), ("w_1", blob), ("u_2", up), ("u_1", up), ("v_2", vp), ("v_1", vp)):
        case.fields[n] = (case.fields[n] + a).astype(np.float32)
    return case


def with_boundary_zone(case):
    """Boundary arrays for a specified / nested case (wrf_b200/bdy.py): the values real.exe would store for the first
    boundary time (the case's own mass-coupled initial state) and tendencies that are not zero -- a slow drift of the
    boundary values with a wave along each side, so that the relaxation, the specified-zone tendencies, the small-step
    updates and the end-of-step reset all have something to do."""
    from wrf_b200 import bdy

    def tend(name, v):
        wave = np.sin(np.arange(v.shape[-1]) * 0.3 + len(name))[None, None, :]
        lev = np.cos(np.arange(v.shape[1]) * 0.2)[None, :, None]
        return (1.0e-4 * v * (0.5 + 0.5 * wave * lev)).astype(np.float32)

    return bdy.add_boundary_arrays(case, tend, spec_exp=0.33)


SPEC = dict(open_xs=0, open_xe=0, open_ys=0, open_ye=0, spec_bdy_width=5, spec_zone=1, relax_zone=4)


class LazyCases(dict):
    """name -> Case, each built the first time it is asked for (building all ~55 states takes several seconds and most
    tests want one): the values of the dict literal in _all_cases are compiled one by one and evaluated on demand."""

    def __init__(self):
        super().__init__()
        import ast
        import inspect
        import textwrap
        tree = ast.parse(textwrap.dedent(inspect.getsource(_all_cases)))
        lit = next(n for n in ast.walk(tree) if isinstance(n, ast.Return)).value
        self._code = {}
        for k, v in zip(lit.keys, lit.values):
            self._code[k.value] = compile(ast.fix_missing_locations(ast.Expression(v)), f"<case {k.value}>", "eval")

    def __missing__(self, name):
        if name not in self._code:
            raise KeyError(name)
        self[name] = case = eval(self._code[name], globals())
        return case

    def __contains__(self, name):
        return name in self._code

    def __iter__(self):
        return iter(self._code)

    def __len__(self):
        return len(self._code)

    def keys(self):
        return self._code.keys()

    def items(self):
        return [(k, self[k]) for k in self._code]

    def values(self):
        return [self[k] for k in self._code]


def small_cases():
    """name -> Case (built on first access); sizes the oracle finishes in about a second per step.  Together they cover
    every lateral boundary type, both advection options, Kessler, the three damping options and
    the namelist switches the path honours (SURVEY.md 8b "Namelist semantics")."""
    return LazyCases()


def _all_cases():
    return {
        # C1 em_hill2d_x: open x / periodic y, hybrid_opt 2, damp_opt 3, passive qv
        "hill2d_x": make_case("hill2d_x", nx=62, ny=3, nz=21),
        # C2 em_quarter_ss: open on all sides, Kessler, damp_opt 2, PD moist advection
        "quarter_ss": make_case("quarter_ss", nx=26, ny=24, nz=21),
        "quarter_ss_nopd": make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=0),
        # C3 em_b_wave: periodic x / symmetric y, f-plane
        "b_wave": make_case("b_wave", nx=21, ny=31, nz=17),
        # C4 em_les: doubly periodic, pert_coriolis, passive tracers with PD advection
        "les": make_case("les", nx=24, ny=22, nz=19),
        # C5 synthetic (the bench workload, small): doubly periodic, bubbles + noise, 2 tracers
        "periodic": make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2),
        # namelist switches away from their defaults
        "b_wave_lid": make_case("b_wave", nx=21, ny=31, nz=17, top_lid=1, phi_adv_z=2, w_damping=1, hypsometric_opt=1),
        "quarter_ss_opts": make_case("quarter_ss", nx=26, ny=24, nz=21, use_theta_m=0, hypsometric_opt=1, damp_opt=3,
                                     zdamp=8000.0, dampcoef=0.2, scalar_adv_opt=0, num_scalar=1, u_scale=1.5),
        "rk2_sym_x": make_case("synthetic", nx=30, ny=28, nz=15, periodic=False, bubble_lattice=20000.0, noise=0.05,
                               calm=True, rk_ord=2, open_xs=0, open_xe=0, symmetric_xs=1, symmetric_xe=1, time_step_sound=4),
        # non-trivial map factors, f, e, sina, cosa: doubly periodic (the interior kernels do every point) and open
        "mapfac": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1)),
        "mapfac_open": with_map_factors(make_case("quarter_ss", nx=30, ny=28, nz=19, w_damping=1)),
        # constant-K diffusion on coordinate surfaces (diff_opt 1 / km_opt 1): open + Kessler (qv uses
        # the base-state form), doubly periodic with map factors and a passive scalar, and em_b_wave as shipped (K = 0)
        "diffusion": make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=1, km_opt=1, khdif=500.0, kvdif=100.0),
        "diffusion_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1,
                                                         diff_opt=1, km_opt=1, khdif=300.0, kvdif=50.0)),
        "b_wave_shipped": make_case("b_wave", nx=21, ny=31, nz=17, diff_opt=1, km_opt=1, khdif=0.0, kvdif=0.0),
        # monotonic scalar advection (*_adv_opt = 2): open + Kessler species, and doubly periodic tracers with map factors
        "mono": make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=2),
        "mono_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                    scalar_adv_opt=2, moist_adv_opt=2)),
        # 5th-order WENO scalar advection (*_adv_opt = 3; theta follows on the last RK substep) and WENO inside the
        # positive-definite limiter (*_adv_opt = 4): open + Kessler species, and doubly periodic tracers with map factors
        "weno": make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=3),
        "weno_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                    scalar_adv_opt=3, moist_adv_opt=3)),
        "wenopd": make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=4),
        "wenopd_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                      scalar_adv_opt=4, moist_adv_opt=3)),
        # WENO momentum advection (momentum_adv_opt = 3, advect_weno_u/v/w on the last RK substep): open + Kessler,
        # doubly periodic with map factors, and periodic x / symmetric y with the rigid lid (w flux at the top)
        "weno_mom": make_case("quarter_ss", nx=26, ny=24, nz=21, momentum_adv_opt=3, moist_adv_opt=4),
        "weno_mom_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1,
                                                        momentum_adv_opt=3)),
        "weno_mom_b_wave": make_case("b_wave", nx=21, ny=31, nz=17, momentum_adv_opt=3, top_lid=1),
        # implicit-explicit vertical advection (zadvect_implicit = 1, module_ieva_em.F) with an updraft strong enough to
        # make wwI non-zero: open + Kessler + PD, doubly periodic with map factors + monotonic / PD tracers, and over
        # the hill (terrain slopes in advect_w_implicit's lower boundary row) together with the WENO options
        # 6th-order horizontal filter (diff_6th_opt 1 / 2, sixth_order_diffusion): open + Kessler with the up-gradient fluxes
        # dropped, doubly periodic with map factors (scalars exempted by scalar_mix6_off), over the hill with the slope taper
        "diff6": make_case("quarter_ss", nx=26, ny=24, nz=21, diff_6th_opt=2),
        "diff6_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                     diff_6th_opt=1, diff_6th_factor=0.2, scalar_mix6_off=1)),
        "diff6_hill": make_case("hill2d_x", nx=62, ny=3, nz=21, diff_6th_opt=2, diff_6th_slopeopt=1, diff_6th_thresh=0.004),
        "ieva": with_updraft(make_case("quarter_ss", nx=26, ny=24, nz=41, zadvect_implicit=1, dt=20.0, time_step=20), 60.0),
        # w_damp with IEVA: damping starts at w_crit_cfl instead of w_beta = 1 (module_big_step_utilities_em.F:2602-2608);
        # the updraft has vertical Courant numbers on both sides of 2 and between 1 and 2; also the CFL maxima (adaptive dt)
        "ieva_wdamp": with_updraft(make_case("quarter_ss", nx=26, ny=24, nz=41, zadvect_implicit=1, dt=20.0, time_step=20, w_damping=1,
                                             w_crit_cfl=2.0, use_adaptive_time_step=1), 60.0, conv=60.0),
        "ieva_periodic": with_updraft(with_map_factors(make_case("synthetic", nx=34, ny=30, nz=33, bubble_lattice=20000.0, noise=0.05,
                                                                  num_scalar=2, scalar_adv_opt=2, moist_adv_opt=1, zadvect_implicit=1,
                                                                  dt=20.0, time_step=20)), 60.0),
        # edges of the envelope: the most levels a Geo carries (126), a domain with no point further than 4 from an open
        # boundary (every fast-kernel box empty), rk_ord = 2 with every optional scheme at once under a rigid lid, and a
        # negative time_step (sign(1,time_step) flips the upwind bias of every odd-order flux, module_advect_em.F:204)
        "tall": make_case("quarter_ss", nx=14, ny=13, nz=126),
        "tiny": make_case("quarter_ss", nx=10, ny=9, nz=11),
        "rk2_everything": with_updraft(make_case("synthetic", nx=30, ny=28, nz=33, periodic=False, bubble_lattice=20000.0, noise=0.05,
                                                 rk_ord=2, time_step_sound=4, zadvect_implicit=1, momentum_adv_opt=3, moist_adv_opt=3,
                                                 num_scalar=1, scalar_adv_opt=4, dt=20.0, time_step=20, top_lid=1, diff_6th_opt=2,
                                                 diff_opt=1, km_opt=1, khdif=200.0, kvdif=20.0), 40.0, conv=20.0),
        "backward": make_case("synthetic", nx=30, ny=28, nz=15, bubble_lattice=20000.0, noise=0.05, num_scalar=1, scalar_adv_opt=3,
                              momentum_adv_opt=3, time_step=-12),
        # even advection orders (h 6 / v 4: the same code with flux6 / flux4): all four at once with open boundaries, Kessler
        # and the PD limiter; doubly periodic with map factors and h 6 only; mixed per family with periodic x / symmetric y
        "order64": make_case("quarter_ss", nx=26, ny=24, nz=21, h_mom_adv_order=6, v_mom_adv_order=4, h_sca_adv_order=6, v_sca_adv_order=4),
        "order63_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                       scalar_adv_opt=0, h_mom_adv_order=6, h_sca_adv_order=6)),
        "order_mixed": make_case("b_wave", nx=21, ny=31, nz=17, h_mom_adv_order=6, v_sca_adv_order=4, moist_adv_opt=4, momentum_adv_opt=3),
        # 2-D Smagorinsky closure on coordinate surfaces (diff_opt 1 / km_opt 4, the usual real-data setting):
        # open + Kessler, doubly periodic with map factors and tracers, over the hill (terrain slopes zx in the deformation),
        # periodic x / symmetric y with vertical constant-K mixing on top
        "smag2d": make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=1, km_opt=4),
        "smag2d_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                      diff_opt=1, km_opt=4, c_s=0.5)),
        "smag2d_hill": make_case("hill2d_x", nx=62, ny=3, nz=21, diff_opt=1, km_opt=4, mix_full_fields=1),
        "smag2d_b_wave": make_case("b_wave", nx=21, ny=31, nz=17, diff_opt=1, km_opt=4, kvdif=5.0),
        # diffusion in physical space (diff_opt 2: vertical_diffusion_2 + horizontal_diffusion_2 with the cal_titau stresses):
        # constant K with open boundaries, Kessler, namelist drag and heat flux; 2-D Smagorinsky K, doubly periodic with map
        # factors, tracers and full-field mixing; constant K over the hill (the terrain-slope terms) with surface fluxes from a
        # host-side surface scheme (isfflx 1); Smagorinsky K with periodic x / symmetric y
        "diff2": make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=1, khdif=300.0, kvdif=100.0, tke_drag_coefficient=0.003,
                           tke_heat_flux=0.05),
        "diff2_periodic": with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=2,
                                                     diff_opt=2, km_opt=4, c_s=0.5, mix_full_fields=1)),
        "diff2_hill": with_surface_fluxes(make_case("hill2d_x", nx=62, ny=3, nz=21, diff_opt=2, km_opt=1, khdif=100.0, kvdif=30.0, isfflx=1)),
        "diff2_b_wave": make_case("b_wave", nx=21, ny=31, nz=17, diff_opt=2, km_opt=4, isfflx=2, tke_heat_flux=0.02),
        # 3-D Smagorinsky closure (km_opt 3: calculate_N2 with its saturated branch, smag_km) feeding the physical-space diffusion:
        # open + Kessler with a cloud and an updraft; doubly periodic, map factors, isotropic length scale; and feeding the
        # coordinate-surface diffusion over the hill
        "smag3d": with_cloud(with_updraft(make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=3, c_s=0.25), 30.0)),
        "smag3d_periodic": with_updraft(with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1,
                                                                   diff_opt=2, km_opt=3, mix_isotropic=1, mix_full_fields=1)), 30.0),
        "smag3d_hill": with_updraft(make_case("hill2d_x", nx=62, ny=3, nz=21, diff_opt=1, km_opt=3, kvdif=10.0), 20.0, conv=70.0),
        # 1.5-order TKE closure (km_opt 2 with diff_opt 2: tke_km, tke_rhs, the prognostic tke advected, mixed and bounded) -- the
        # closure of the shipped em_quarter_ss and em_les namelists: open + Kessler as shipped (mix_full_fields), doubly periodic
        # with the isotropic length scale, namelist drag / heat flux and a tracer, over the hill with fluxes from a surface scheme
        "tke": with_tke(with_cloud(with_updraft(make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=2, mix_full_fields=1), 30.0))),
        "tke_periodic": with_tke(with_updraft(with_map_factors(make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05,
                                                                         num_scalar=1, diff_opt=2, km_opt=2, mix_isotropic=1, tke_adv_opt=2,
                                                                         tke_drag_coefficient=0.003, tke_heat_flux=0.05)), 30.0)),
        "tke_hill": with_tke(with_surface_fluxes(make_case("hill2d_x", nx=62, ny=3, nz=21, diff_opt=2, km_opt=2, isfflx=1))),
        # C1-C4 exactly as shipped: test/em_*/namelist.input + input_sounding / input_jet (tests/fixtures, unchanged copies) through
        # wrf_b200/shipped.py (= ideal.exe + start_em), at the shipped dimensions 202x3x41, 42x42x41, 41x81x65, 40x40x40.  em_les
        # selects sf_sfclay_physics = 1 (host-side physics): its friction velocity is stood in for by a smooth pattern.
        "shipped_hill2d_x": shipped_case("em_hill2d_x"),
        "shipped_quarter_ss": shipped_case("em_quarter_ss"),
        "shipped_b_wave": shipped_case("em_b_wave"),
        "shipped_les": with_surface_fluxes(shipped_case("em_les")),
        # specified / nested lateral boundaries: relaxation + specified zone driven by boundary arrays with non-zero
        # tendencies.  Kessler with flow-dependent qc, qr; the real-data dynamics (2-D Smagorinsky on coordinate surfaces, w_damping,
        # damp_opt 3, map factors, a tracer without boundary arrays); a nest (w and every species relaxed too); boundary arrays for
        # every moist species with the 6th-order filter and even-order advection; the TKE closure in physical space
        "specified": with_boundary_zone(make_case("quarter_ss", nx=30, ny=28, nz=19, specified=1, **SPEC)),
        "specified_real": with_boundary_zone(with_map_factors(make_case("quarter_ss", nx=30, ny=28, nz=19, specified=1, diff_opt=1, km_opt=4,
                                                                         w_damping=1, damp_opt=3, zdamp=8000.0, dampcoef=0.2, num_scalar=1,
                                                                         mp_physics=0, num_moist=1, **SPEC))),
        "nested": with_boundary_zone(make_case("quarter_ss", nx=30, ny=28, nz=19, nested=1, num_scalar=1, **SPEC)),
        "specified_bcs_moist": with_boundary_zone(make_case("quarter_ss", nx=30, ny=28, nz=19, specified=1, have_bcs_moist=1, diff_6th_opt=2,
                                                            h_mom_adv_order=6, v_mom_adv_order=4, h_sca_adv_order=6, v_sca_adv_order=4,
                                                            moist_adv_opt=0, **SPEC)),
        "specified_tke": with_boundary_zone(with_tke(with_updraft(make_case("quarter_ss", nx=30, ny=28, nz=19, specified=1, diff_opt=2, km_opt=2,
                                                                            mix_full_fields=1, **SPEC), 20.0))),
        "ieva_weno_hill": with_updraft(make_case("hill2d_x", nx=62, ny=3, nz=41, zadvect_implicit=1, momentum_adv_opt=3,
                                                 moist_adv_opt=3), 60.0, conv=70.0),
    }


def common_fields(o: Oracle, s: Solver):
    on = set(o.names())
    c = o.case.config
    skip = set()
    if (c.specified or c.nested) and c.mp_physics and c.diff_opt == 2 and c.km_opt in (2, 3):
        # th_phy is scratch with two writers in the oracle: phy_prep for calculate_N2 (every mass point) and moist_physics_prep
        # (all but the specified rows, solve_em.F:3623-3645).  The device evaluates the first in registers (NOT_MATERIALISED),
        # so its specified rows keep Kessler's untouched zeros; nothing reads them.
        skip.add("th_phy")
    return [n for n in s.names() if n in on and n not in skip]


def rel_l2(a, b):
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    den = np.sqrt((b * b).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return 0.0 if num == 0.0 else (num / den if den > 0 else np.inf)


# perturbation fields are differences of O(1e5 Pa) / O(1 m3/kg) totals: their error is
# measured against the full field the model actually uses (p+pb, al+alb, ph+phb, mu+mub)
BASE_OF = {"p": "pb", "al": "alb", "ph_2": "phb", "ph_1": "phb", "mu_2": "mub", "mu_1": "mub", "t_2": None, "t_1": None}


def rel_l2_field(o: Oracle, name, a, b):
    base = BASE_OF.get(name, "")
    if base == "":
        return rel_l2(a, b)
    if base is None:       # theta perturbation: full theta = t + 300
        return rel_l2(a.astype(np.float64) + 300.0, b.astype(np.float64) + 300.0)
    bb = o.get(base).astype(np.float64)
    return rel_l2(a.astype(np.float64) + bb, b.astype(np.float64) + bb)


def diff_report(o: Oracle, s: Solver, fields=None):
    """returns list of (field, n_mismatch, max_abs, rel_l2)"""
    out = []
    for n in fields or common_fields(o, s):
        a, b = s.get(n), o.get(n)
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        if bad.any():
            with np.errstate(invalid="ignore"):
                out.append((n, int(bad.sum()), float(np.nanmax(np.abs(a.astype(np.float64) - b))), rel_l2_field(o, n, a, b)))
    return out


def sync_from_oracle(o: Oracle, s: Solver, fields=None):
    for n in fields or common_fields(o, s):
        s.set(n, o.get(n))


def phase_sequence(case):
    """(phase, args) list of one full solve_em step -- solve_em.F:578-4848."""
    c = case.config
    nss = c.time_step_sound
    seq = []
    for rk in range(1, c.rk_ord + 1):
        ns = 1 if (c.rk_ord == 3 and rk == 1) else (nss // 2 if rk < c.rk_ord else nss)
        seq += [("rk_step_prep", (rk,)), ("rk_phys_bc_dry_1", ())]
        if rk == 1:
            seq += [("init_zero_tendency", ()), ("diffusion_prep", ())]
        seq += [("rk_tendency", (rk,)), ("small_step_prep", (rk,))]
        for it in range(1, ns + 1):
            seq += [("advance_uv", (rk,)), ("advance_mu_t", (rk,)), ("advance_w", (rk,)), ("sumflux", (rk, it)), ("calc_p_rho", (it,))]
        seq += [("small_step_finish", (rk,)), ("rk_scalar_step", (rk,)), ("calc_p_rho_phi", ())]
        if rk < c.rk_ord:
            seq.append(("rk_phys_bc_dry_2", ()))
    seq += [("microphysics", ()), ("end_of_step", ())]
    return seq


# Scratch arrays of the oracle the device never materialises in a phase: the closures' phy_prep (th_phy of
# module_big_step_utilities_em.F:5394, read by calculate_N2) is evaluated in registers by k_smag3d / k_tke_km; the device's
# th_phy only exists for Kessler and is written by its own moist_physics_prep in the microphysics phase (compared there).
NOT_MATERIALISED = {("diffusion_prep", "th_phy")}


def run_phase_parity(case, verbose=False):
    """Run one step phase by phase; the GPU input of every phase is the oracle's state.
    Returns list of failures (phase, field, n, maxabs, rel)."""
    o, s = Oracle(case), Solver(case)
    o.start_domain(); s.start_domain()
    if case.config.specified or case.config.nested:   # grid%dtbc as solve_em leaves it for the first step (solve_em.F:377)
        o.set_dtbc(case.config.dt); s.set_dtbc(case.config.dt)
    fails = []
    rep = diff_report(o, s)
    for r in rep:
        fails.append(("start_domain",) + r)
    fields = common_fields(o, s)
    for phase, args in phase_sequence(case):
        o.call(phase, *args)
        s.call(phase, *args)
        rep = diff_report(o, s, fields)
        for r in rep:
            ok = (phase, r[0]) in NOT_MATERIALISED   # any differing bit of anything else is a failure
            if verbose or not ok:
                print(f"  {'ok ' if ok else 'BAD'} {phase}{args}: {r[0]:12s} n={r[1]:7d} maxabs={r[2]:.3e} rel_l2={r[3]:.3e}")
            if not ok:
                fails.append((phase,) + r)
        if rep:
            sync_from_oracle(o, s, [r[0] for r in rep])
    s.close()
    return fails


def tophat_case(adv_opt, nx=91, ny=3, nz=10):
    """The reference's own kernel harness, PROGRAM feeder (dyn_em/module_advect_em.F:10567-10870,
    built with -DADVECT_KERNEL): 91x3x10 doubly periodic box, top-hat scalar between i=36 and 54
    (tophat, :56-80), ru = 90, rv = rom = 0, mu = 1, rdx = rdy = 1/1000, dt = 5, levels
    znw = 1-(k-1)/(kte-1).  One feeder `loop` is the three advect_scalar calls with dt/3, dt/2, dt,
    i.e. exactly rk_scalar_step(1..3) of this path with scalar_adv_opt = 0 (1: the third call is
    advect_scalar_pd)."""
    from wrf_b200 import ideal
    F = np.float32
    d = ideal.memory_dims(nx, ny, nz)
    c = ideal.default_config(dt=5.0, dx=1000.0, dy=1000.0, time_step=5, time_step_sound=4, periodic_x=1, periodic_y=1,
                             num_moist=0, num_scalar=1, scalar_adv_opt=adv_opt, mp_physics=0)
    v, sc = ideal.vertical_coords(nz, 10000.0, stretch=False)
    for k_, val in sc.items():
        setattr(c, k_, float(val))
    nj, nk, ni = d.jme - d.jms + 1, nz, d.ime - d.ims + 1
    fl = {n: v[n] for n in ("znu", "znw", "fnm", "fnp", "rdnw", "rdn", "dnw", "dn")}
    one1, zero1 = np.ones(nz, F), np.zeros(nz, F)
    fl.update(c1h=one1, c1f=one1, c2h=zero1, c2f=zero1, c3h=v["znu"], c3f=v["znw"], c4h=zero1, c4f=zero1)
    ones2 = np.ones((nj, ni), F)
    for n in ("msftx", "msfty", "msfux", "msfuy", "msfvx", "msfvx_inv", "msfvy", "mub", "muts", "mut"):
        fl[n] = ones2.copy()
    fl["ru_m"] = np.full((nj, nk, ni), 90.0, F)
    q = np.zeros((nj, nk, ni), F)
    ii = np.arange(d.ims, d.ime + 1)
    jj = np.arange(d.jms, d.jme + 1)
    hat = ((ii > 35) & (ii < 55) & (ii >= 1) & (ii <= nx))
    q[np.ix_((jj >= 1) & (jj <= ny), np.arange(nk), hat)] = 1.0
    fl["scalar_1"] = q
    fl["scalar_old_1"] = q.copy()
    return ideal.Case(name="tophat", dims=d, config=c, fields={k: np.ascontiguousarray(a, dtype=F) for k, a in fl.items()})


def feeder_loop(m, nloops, fill_halo=True):
    """m: Oracle or Solver.  fill_halo=False reproduces the feeder literally (it never refreshes the halo)."""
    for _ in range(nloops):
        for rk in (1, 2, 3):
            m.call("rk_scalar_step", rk)
            if fill_halo:
                m.bc("scalar_1", "p")
