"""Random namelists through the reference-source comparison (tests/test_reference_source.py): for every seed a configuration is drawn
(case, lateral boundaries, advection orders and options for momentum / moisture / scalars / tke, damping, closures, filter, lid, theta_m,
hypsometric option, Coriolis form, RK order, sign of time_step, acoustic coefficients; with a third argument also specified / nested
boundaries, IEVA and the TKE closure), and every phase of a step is run from the reference's Fortran and from the oracle on the same state;
any field that differs in a single bit is reported.  Build-container tool (needs /root/reference):

    python tools/fuzz_reference.py <first seed> <count> [x [bdy]]

Equal or default values hide shared transliteration slips (w advected with the momentum orders went unnoticed as long as every case had
h_mom = h_sca); drawing the options independently is what finds them.  Known, explained reports: IEVA with periodic boundaries (the
reference reads never-written elements of its local wwI / mut_new at the periodic edge; the translator fills locals with NaN), advect_tend
with tke and scalars together (scratch left by the last species, the two sides order the species differently)."""
import sys, random, time, traceback
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from tests import parity
from tests.test_reference_source import (run, tendency_program, scalar_program, acoustic_program, diffusion_program, final_program)
from wrf_b200 import make_case

seed0 = int(sys.argv[1]); n = int(sys.argv[2])
for seed in range(seed0, seed0 + n):
    R = random.Random(seed)
    kw = {}
    base = R.choice(["quarter_ss", "synthetic", "hill2d_x", "b_wave"])
    dims = dict(nx=12, ny=10, nz=9)
    if base == "hill2d_x": dims = dict(nx=20, ny=3, nz=9)
    if base == "b_wave": dims = dict(nx=9, ny=13, nz=9)
    if base == "synthetic":
        dims = dict(nx=14, ny=12, nz=9); kw.update(bubble_lattice=8000.0, noise=0.05, num_scalar=R.choice([0, 1, 2]))
    kw.update(h_mom_adv_order=R.choice([5, 6]), v_mom_adv_order=R.choice([3, 4]), h_sca_adv_order=R.choice([5, 6]), v_sca_adv_order=R.choice([3, 4]))
    kw["moist_adv_opt"] = R.choice([0, 1, 3, 4]); kw["scalar_adv_opt"] = R.choice([0, 1, 3, 4]); kw["momentum_adv_opt"] = R.choice([1, 3])
    if kw["h_sca_adv_order"] == 5 and kw["v_sca_adv_order"] == 3 and R.random() < 0.3:
        kw["moist_adv_opt"] = 2; kw["scalar_adv_opt"] = 2
    kw["damp_opt"] = R.choice([0, 2, 3]); kw["w_damping"] = R.choice([0, 1])
    do, ko = R.choice([(0, 0), (1, 1), (1, 4), (2, 1), (2, 4), (2, 3)])
    if do: kw.update(diff_opt=do, km_opt=ko, khdif=R.choice([0.0, 300.0]), kvdif=R.choice([0.0, 50.0]))
    kw["diff_6th_opt"] = R.choice([0, 1, 2])
    kw["top_lid"] = R.choice([0, 1]); kw["phi_adv_z"] = R.choice([1, 2]); kw["hypsometric_opt"] = R.choice([1, 2]); kw["use_theta_m"] = R.choice([0, 1])
    kw["pert_coriolis"] = R.choice([0, 1]); kw["rk_ord"] = R.choice([2, 3, 3]); kw["mix_full_fields"] = R.choice([0, 1])
    if R.random() < 0.2: kw["time_step"] = -12
    kw["epssm"] = R.choice([0.1, 0.3]); kw["smdiv"] = R.choice([0.1, 0.0]); kw["emdiv"] = R.choice([0.01, 0.0])
    kw["hybrid_opt"] = R.choice([0, 2]); kw["time_step_sound"] = R.choice([2, 4, 6])
    if base == "quarter_ss" and R.random() < 0.3: kw.update(mp_physics=0, num_moist=1)
    mode = R.choice(["plain", "spec", "nest", "ieva", "tke"]) if len(sys.argv) > 3 else "plain"
    wrap = []
    if len(sys.argv) > 4 and sys.argv[4] == "bdy":          # every seed a specified / nested configuration
        mode, base, dims = R.choice(["spec", "nest"]), "quarter_ss", dict(nx=16, ny=14, nz=9)
        kw.pop("bubble_lattice", None); kw.pop("noise", None); kw.pop("num_scalar", None)
    if mode in ("spec", "nest") and base in ("quarter_ss",):
        dims = dict(nx=16, ny=14, nz=9)
        kw.update(parity.SPEC); kw["specified" if mode == "spec" else "nested"] = 1
        kw["momentum_adv_opt"] = 1
        for k_ in ("moist_adv_opt", "scalar_adv_opt"):
            if kw[k_] in (2, 3, 4): kw[k_] = R.choice([0, 1])
        if mode == "spec": kw["have_bcs_moist"] = R.choice([0, 1])
        if mode == "nest": kw["num_scalar"] = 1
        wrap.append(parity.with_boundary_zone)
    if mode == "ieva" and base != "b_wave":
        kw.update(zadvect_implicit=1, dt=40.0, time_step=40 if kw.get("time_step", 1) > 0 else -40)
        dims = dict(dims, nz=17)
    if mode == "tke":
        kw.update(diff_opt=2, km_opt=2, mix_isotropic=R.choice([0, 1]), tke_adv_opt=R.choice([1, 2]), tke_drag_coefficient=R.choice([0.0, 0.003]),
                  tke_heat_flux=R.choice([0.0, 0.05]))
        wrap.append(parity.with_tke)
    try:
        case = make_case(base, **dims, **kw)
        if mode == "ieva": case = parity.with_updraft(case, 80.0)
        for w in wrap: case = w(case)
        if R.random() < 0.5: case = parity.with_map_factors(case)
        if R.random() < 0.5: case = parity.with_updraft(case, 15.0)
        case.refname = f"fuzz{seed}"
        c = case.config
    except Exception as e:
        print(seed, "SKIP make_case", str(e)[:100], flush=True); continue
    t0 = time.time(); bad = []
    def go(phase, rk, prog, **k):
        try:
            run(case, phase, rk, prog, **k)
        except AssertionError as e:
            if "vacuous" not in str(e): bad.append((phase, rk, str(e)[:300]))
        except Exception as e:
            bad.append((phase, rk, "ERR " + type(e).__name__ + " " + str(e)[:200]))
    last = c.rk_ord
    if c.diff_opt and not (c.diff_opt == 1 and c.km_opt == 1) and not (c.diff_opt == 2 and c.km_opt == 1 and c.khdif == 0 and c.kvdif == 0):
        skip = ["pi_phy", "z_at_w", "dz8w"] + (["th_phy", "p_phy", "t_phy", "bn2"] if c.km_opt not in (2, 3) else []) + (["rho"] if c.diff_opt != 2 else [])
        go("diffusion_prep", None, diffusion_program(c), not_compared=skip)
    for rk in sorted({1, last}):
        go("rk_step_prep", rk, [("rk_step_prep", 0)])
        prog, skip = tendency_program(c, rk)
        go("rk_tendency", rk, prog, not_compared=skip)
        go("small_step_prep", rk, [("small_step_prep", 0), ("calc_p_rho", 0), ("calc_coef_w", 0)])
        for r in ("advance_uv", "advance_mu_t", "advance_w", "sumflux"):
            go(r, rk, acoustic_program(c, r))
        go("small_step_finish", rk, [("calc_mu_uv_1", 0), ("small_step_finish", 0)])
        if c.num_moist + c.num_scalar:
            go("rk_scalar_step", rk, scalar_program(c, rk, set(case.fields)), not_compared=["h_tendency", "z_tendency", "xkmh", "xkmv", "xkhh", "xkhv"])
    if c.mp_physics == 1 and not (c.specified or c.nested):
        from tests.refcases import rainy
        import copy
        d = case.dims
        tile = dict(its=d.ips, ite=min(d.ipe, d.ide - 1), jts=d.jps, jte=min(d.jpe, d.jde - 1))
        alias = {"moist(ims,kms,jms,p_qv)": "moist_qv", "moist(ims,kms,jms,p_qc)": "moist_qc"}
        kess = {"t": "th_phy", "qv": "moist_qv", "qc": "moist_qc", "qr": "moist_qr", "pii": "pi_phy", "dt_in": "dt", "xlv": "xlv", "cp": "cp",
                "ep2": "ep_2", "kte": d.kde - 1}
        f32 = np.float32
        keep = case
        case = rainy(copy.deepcopy(case)); case.refname = keep.refname
        go("microphysics", None, [("moist_physics_prep_em", 0, {"alias": alias, "scalars": {"dtm": f32(c.dt), **tile}}),
                                   ("kessler", 0, {"actuals": kess, "scalars": tile}),
                                   ("moist_physics_finish_em", 0, {"alias": alias, "scalars": {"dtm": f32(c.dt), **tile}})])
        case = keep
    go("calc_p_rho_phi", None, [("calc_p_rho_phi", 0)], occurrence=1)
    go("end_of_step", None, final_program(c, set(case.fields)))
    print(seed, base, mode, "OK" if not bad else "DIFF", round(time.time() - t0, 1), {k: v for k, v in kw.items() if k not in ("bubble_lattice", "noise")} if bad else "", flush=True)
    for b in bad: print("    ", b, flush=True)
