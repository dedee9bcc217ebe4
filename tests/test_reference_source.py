"""The oracle against the reference's OWN SOURCE, executed here.

tests/fortran_subset.py translates the routines of dyn_em/module_em.F, module_small_step_em.F, module_big_step_utilities_em.F,
module_advect_em.F, ... (read where they lie under /root/reference; nothing is copied) into Python with real*4 semantics,
tests/refrun.py binds each one as the CALL statement in dyn_em/solve_em.F binds it, and every test below

  1. advances an Oracle to the point just before one phase of a solve_em step,
  2. runs the reference routines of that phase on a copy of the oracle's state,
  3. runs the oracle's phase,
  4. requires every field the reference routines were handed to be equal BIT FOR BIT on the domain's own points.

That pins oracle/ (and through the GPU parity tests the device code) to the reference's arithmetic, statement by statement, rather
than to a reading of it.  The reference tree only exists in the build container: without it the tests skip, and
tests/golden/reference_source.json (written by this file with REFSRC_WRITE=1, read by test_reference_source_golden.py) keeps
what they established: a checksum of every compared field, which the oracle must reproduce wherever the suite runs."""
import json
import os

import numpy as np
import pytest

from tests import parity, refrun
from tests.refcases import CASES, at_phase, build, digest, oracle_feeder, own, rk_scalars
from wrf_b200 import decomp

pytestmark = pytest.mark.skipif(not refrun.have_reference(), reason="needs the reference tree (/root/reference)")
f32 = np.float32


GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_source.json")
RECORD = [] if os.environ.get("REFSRC_WRITE") == "1" else None
# The full matrix (REFSRC_FULL=1, implied by REFSRC_WRITE=1: the fixture always holds all of it) takes 2.5 minutes; the default run leaves
# out the combinations that repeat a code path already run (rk_tendency at the middle Runge-Kutta stage, and at the last stage for the
# configurations whose last stage has nothing of its own), which the fixture replay of test_reference_source_golden.py still covers.
FULL = RECORD is not None or os.environ.get("REFSRC_FULL") == "1"
# the configurations of the default run: one per code path family (the others repeat paths with other boundary types or sizes)
CORE = {"w_fast_uv_gather", "uv_fast_w_gather", "open_mapfac", "periodic_mapfac", "b_wave_lid", "rk2_sym_x", "hill", "diffusion_periodic", "order_mixed",
        "weno", "ieva_wdamp", "specified", "nested", "smag2d", "smag3d_pow", "tke_periodic", "les", "mono", "wenopd", "diff6_hill", "rk2_everything"}
LAST_STAGE_DIFFERS = {"w_fast_uv_gather", "uv_fast_w_gather", "rk2_everything", "order_mixed", "weno", "backward", "ieva", "ieva_wdamp", "wenopd", "wenopd_periodic", "rk2_sym_x", "specified", "open_mapfac",
                      "periodic_mapfac", "smag2d", "tke", "les"}


@pytest.fixture(scope="module", autouse=True)
def _write_golden():
    """REFSRC_WRITE=1 python -m pytest tests/test_reference_source.py  regenerates tests/golden/reference_source.json: for every
    comparison made here, a digest of each field as the REFERENCE's routines left it (own points, or the whole array for the boundary
    fills).  tests/test_reference_source_golden.py holds the oracle to those digests where the reference tree is absent."""
    yield
    if RECORD is not None:
        with open(GOLDEN, "w") as f:
            ran = {k: r.calls for k, r in sorted(refrun.library(1).cache.items()) if r.calls}
            json.dump({"generator": "REFSRC_WRITE=1 python -m pytest tests/test_reference_source.py", "records": RECORD,
                       "routines_executed": ran}, f, separators=(",", ":"), sort_keys=True)


def compare(case, state, oracle, phase, whole=False):
    """-> (n compared, [names that differ]); whole: halo included (the boundary fills)"""
    bad, n = [], 0
    for k in sorted(state.fields):
        if k in state.absent or decomp.bdy_side(k):      # boundary arrays are inputs
            continue
        a, b = state.result(k), np.asarray(oracle.get(k))
        if not whole:
            a, b = own(case, k, a), own(case, k, b)
        n += 1
        if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
            both_nan = np.isnan(a) & np.isnan(b)
            d = np.argwhere((a.view(np.uint32) != b.view(np.uint32)) & ~both_nan)
            if len(d):
                bad.append((k, len(d), tuple(d[0]), float(a[tuple(d[0])]), float(b[tuple(d[0])])))
    return n, bad


def run(case, phase, rk, program, occurrence=0, extra=None, not_compared=(), whole=False):
    o, args = at_phase(case, phase, rk, occurrence)
    rk_in, rk = rk, rk or 1
    scal = rk_scalars(case.config, rk)
    scal.update(rk_step=rk, rk_order=case.config.rk_ord, iteration=args[1] if len(args) > 1 else 1)
    if case.config.specified or case.config.nested:
        scal["dtbc"] = f32(o.dtbc())
    scal.update(extra or {})
    st = refrun.RefState(o)
    for step in program:
        routine, occ = step[0], step[1]
        opts = step[2] if len(step) > 2 else {}
        ret = refrun.RefCall(routine, occ, opts.get("path", "dyn_em/solve_em.F"), opts.get("alias"), opts.get("actuals")).run(
            st, **{**scal, **opts.get("scalars", {})})
        st.returns.append((routine, ret))            # the scalar dummies after the call (max_vert_cfl, ...)
    changed = [k for k in st.fields if k not in st.absent and not decomp.bdy_side(k) and not np.array_equal(st.result(k), np.asarray(o.get(k)), equal_nan=True)]
    assert changed, f"{phase}: the reference routines changed nothing (a vacuous comparison)"
    o.call(phase, *args)
    for k in not_compared:
        st.absent.add(k)
    n, bad = compare(case, st, o, phase, whole)
    assert not bad, f"{phase}: reference source and oracle differ in {bad}"
    if RECORD is not None:
        fields = {}
        for k in sorted(st.fields):
            if k in st.absent or decomp.bdy_side(k):
                continue
            a = st.result(k)
            fields[k] = digest(a if whole else own(case, k, a))
        RECORD.append({"case": case.refname, "phase": phase, "rk": rk_in, "occurrence": occurrence, "whole": whole, "fields": fields})
    return n, st, o


# ---------------------------------------------------------------------------------------------------------------------
# the acoustic loop and its brackets (module_small_step_em.F), at every Runge-Kutta stage

@pytest.mark.parametrize("routine,rk,least", [("advance_uv", 1, 30), ("advance_uv", 2, 30), ("advance_uv", 3, 30), ("advance_mu_t", 1, 35),
                                              ("advance_mu_t", 3, 35), ("advance_w", 1, 35), ("advance_w", 3, 35), ("sumflux", 1, 20),
                                              ("sumflux", 2, 20), ("sumflux", 3, 20)])
def test_acoustic_step(routine, rk, least):
    n, *_ = run(build("small"), routine, rk, [(routine, 0)], occurrence=1 if rk > 1 else 0)
    assert n >= least


# ---------------------------------------------------------------------------------------------------------------------
# the large-step tendencies: rk_tendency (module_em.F:188) with everything it calls -- advection of u, v, w, ph, theta,
# curvature, Coriolis, pressure gradient and buoyancy, horizontal/vertical diffusion, Rayleigh damping -- then rk_addtend_dry

def tendency_program(c, rk):
    """the reference calls between rk_step_prep's boundary fills and small_step_prep (solve_em.F:790-1100) -> (program, fields left out)"""
    prog, skip = [], ()
    if c.diff_opt == 1 and c.km_opt == 1 and rk == 1:
        # first_rk_step_part2 -> calculate_km_kh -> isotropic_km fills the K arrays with the namelist constants once per step; the
        # oracle and the device hand the constants to the diffusion routines directly and never materialise the arrays
        prog.append(("isotropic_km", 0, {"actuals": {}}))
        for k in ("xkmh", "xkhh"):      # phy_bc (module_diffusion_em.F:5901): the halo the horizontal operators read
            prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": k, "variable_in": "'t'"}}))
        skip = ("xkmh", "xkmv", "xkhh", "xkhv")
    prog.append(("rk_tendency", 0))
    if (c.specified or c.nested) and rk == 1:        # solve_em.F:937
        prog.append(("relax_bdy_dry", 0))
    prog.append(("rk_addtend_dry", 0))
    if c.specified or c.nested:
        prog.append(("spec_bdy_dry", 0))
    return prog, skip


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("rk", [1, 2, 3])
def test_rk_tendency(name, rk):
    if not FULL and (name not in CORE or rk == 2 or (rk == 3 and name not in LAST_STAGE_DIFFERS)):
        pytest.skip("REFSRC_FULL=1 runs it; its digests are in the fixture")
    case = build(name)
    if rk > case.config.rk_ord:
        pytest.skip("rk_ord 2")
    prog, skip = tendency_program(case.config, rk)
    n, st, o = run(case, "rk_tendency", rk, prog, not_compared=skip)
    assert n >= 70
    if name == "ieva" and rk == 3:
        assert np.any(np.asarray(o.get("wwI"))), "the IEVA case must move some of the vertical transport to the implicit solver"


# ---------------------------------------------------------------------------------------------------------------------
# the brackets of the acoustic loop and the diagnostics

@pytest.mark.parametrize("name", ["ieva_wdamp", "adaptive"])
def test_w_damp_cfl_maxima(name):
    """max_vert_cfl / max_horiz_cfl as w_damp returns them through rk_tendency (module_big_step_utilities_em.F:2503-2711), the input of
    the adaptive time step (solve_em.F:914-930); with IEVA the damping starts at w_crit_cfl"""
    case = build(name)
    prog, skip = tendency_program(case.config, 1)
    o, args = at_phase(case, "rk_tendency", 1)
    o.max_cfl(reset=True)
    n, st, o = run(case, "rk_tendency", 1, prog, not_compared=skip)
    ret = dict(st.returns)["rk_tendency"]
    mv, mh = o.max_cfl()
    assert f32(ret["max_vert_cfl"]) == f32(mv) and f32(ret["max_horiz_cfl"]) == f32(mh) and mv > 0 and mh > 0
    if name == "ieva_wdamp":
        assert mv > 1.0, "the case must have vertical Courant numbers above the explicit limit"


SOME = ["les", "open_mapfac", "b_wave_lid", "opts", "periodic_mapfac", "rk2_sym_x", "hill", "specified"]


@pytest.mark.parametrize("name", SOME)
@pytest.mark.parametrize("rk", [1, 2])
def test_rk_step_prep(name, rk):
    n, *_ = run(build(name), "rk_step_prep", rk, [("rk_step_prep", 0)])
    assert n >= 25


@pytest.mark.parametrize("name", SOME)
@pytest.mark.parametrize("rk", [1, 2])
def test_small_step_prep(name, rk):
    """small_step_prep, calc_p_rho(step 0) and calc_coef_w (solve_em.F:1100-1175)"""
    n, *_ = run(build(name), "small_step_prep", rk, [("small_step_prep", 0), ("calc_p_rho", 0), ("calc_coef_w", 0)])
    assert n >= 40


@pytest.mark.parametrize("name", SOME)
def test_calc_p_rho(name):
    """the pressure / inverse density diagnosis inside the acoustic loop (solve_em.F:1623), second small step of stage 2 or 3"""
    case = build(name)
    n, *_ = run(case, "calc_p_rho", None, [("calc_p_rho", 1)], occurrence=2, extra={"iteration": 2})
    assert n >= 10


@pytest.mark.parametrize("name", SOME)
@pytest.mark.parametrize("rk", [1, 3])
def test_small_step_finish(name, rk):
    case = build(name)
    rk = min(rk, case.config.rk_ord)
    n, *_ = run(case, "small_step_finish", rk, [("calc_mu_uv_1", 0), ("small_step_finish", 0)])
    assert n >= 30


@pytest.mark.parametrize("name", SOME)
def test_calc_p_rho_phi(name):
    n, *_ = run(build(name), "calc_p_rho_phi", None, [("calc_p_rho_phi", 0)], occurrence=1)
    assert n >= 10


def acoustic_program(c, routine):
    """a routine of the acoustic loop with the specified-zone updates solve_em.F makes right after it (:1344-1365, :1459-1490,
    :1582-1620)"""
    prog = [(routine, 0)]
    if not (c.specified or c.nested):
        return prog
    if routine == "advance_uv":
        prog += [("spec_bdyupdate", 0), ("spec_bdyupdate", 1)]
    if routine == "advance_mu_t":
        prog += [("spec_bdyupdate", 2), ("spec_bdyupdate", 3), ("spec_bdyupdate", 4)]
    if routine == "sumflux":
        prog += [("spec_bdyupdate_ph", 0), ("zero_grad_bdy", 0) if c.specified else ("spec_bdyupdate", 5)]
    return prog


@pytest.mark.parametrize("name", ["specified", "nested", "periodic_mapfac", "b_wave_lid", "hill", "rk2_sym_x"])
@pytest.mark.parametrize("routine", ["advance_uv", "advance_mu_t", "advance_w", "sumflux"])
def test_acoustic_step_configurations(name, routine):
    case = build(name)
    rk = case.config.rk_ord
    n, *_ = run(case, routine, rk, acoustic_program(case.config, routine), occurrence=1)
    assert n >= 20


# ---------------------------------------------------------------------------------------------------------------------
# the scalar step (solve_em.F:1845-2955): rk_update_scalar_pd, rk_scalar_tend with every advection option, the species'
# boundary-zone tendencies, rk_update_scalar, flow_dep_bdy -- species by species as the moist and scalar loops run them

def species_alias(fam, member, index_var):
    """the slab actuals of one species, moist(ims,kms,jms,im) ..., as the oracle's per-species field names"""
    tail = member[len(fam) + 1:]                      # qv | 1
    al = {f"{fam}(ims,kms,jms,{index_var})": member, f"{fam}_old(ims,kms,jms,{index_var})": f"{fam}_old_{tail}",
          f"{fam}_tend(ims,kms,jms,{index_var})": f"{fam}_tend_{tail}"}
    for side in ("xs", "xe", "ys", "ye"):
        first = "jms" if side[0] == "x" else "ims"
        al[f"{fam}_b{side}({first},kms,1,{index_var})"] = f"{member}_b{side}"
        al[f"{fam}_bt{side}({first},kms,1,{index_var})"] = f"{member}_bt{side}"
    return al


def scalar_program(c, rk, have):
    prog = []
    if c.diff_opt == 1 and c.km_opt == 1 and rk == 1:          # the K arrays of isotropic_km, as in tendency_program
        prog.append(("isotropic_km", 0, {"actuals": {}}))
        for k in ("xkmh", "xkhh"):
            prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": k, "variable_in": "'t'"}}))
    fams = [("moist", ["moist_" + q for q in ("qv", "qc", "qr")[:c.num_moist]], "im", c.moist_adv_opt, 0),
            ("scalar", ["scalar_%d" % n for n in range(1, c.num_scalar + 1)], "is", c.scalar_adv_opt, 1)]
    last = rk == c.rk_ord
    for fam, members, iv, adv, nth in fams:                 # solve_em.F:1849-1977
        if last and adv in (1, 2, 4):
            for n, m in enumerate(members):
                prog.append(("rk_update_scalar_pd", fam + "_tend(", {"alias": species_alias(fam, m, "im"), "scalars": {"im": n + 2}}))
    for fam, members, iv, adv, nth in fams:
        if last and adv in (1, 2, 4):
            for m in members:                              # set_physical_bc3d on the _old fields (:1897-1907, :1960-1970)
                old = species_alias(fam, m, iv)[f"{fam}_old(ims,kms,jms,{iv})"]
                prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": old, "variable_in": "'p'"}}))
    if c.km_opt == 2:          # the prognostic tke (solve_em.F:2100-2142, :2376-2456): the oracle's last scalar species
        n = c.num_scalar + 1
        al = {"grid%tke_2": f"scalar_{n}", "grid%tke_1": f"scalar_old_{n}", "tke_tend(ims,kms,jms)": f"scalar_tend_{n}"}
        tk = {"alias": al, "scalars": {"tenddec": False}}
        if last and c.tke_adv_opt in (1, 2, 4):
            prog.append(("rk_update_scalar_pd", "grid%tke_1,", tk))
            prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": f"scalar_old_{n}", "variable_in": "'p'"}}))
    for fam, members, iv, adv, nth in fams:
        for n, m in enumerate(members):
            opts = {"alias": species_alias(fam, m, iv), "scalars": {iv: n + 2, "tenddec": False}}
            occ = fam + "_tend("
            prog.append(("rk_scalar_tend", occ, opts))
            spec = c.specified or c.nested
            if spec and rk == 1 and (m + "_bxs") in have:
                prog.append(("relax_bdy_scalar", occ, opts))
                prog.append(("spec_bdy_scalar", occ, opts))
            prog.append(("rk_update_scalar", occ, opts))
            if fam == "moist" and c.specified and not c.have_bcs_moist and m != "moist_qv":
                prog.append(("flow_dep_bdy", 0, opts))
        if fam == "moist" and c.km_opt == 2:
            prog += [("rk_scalar_tend", "grid%tke_1,", tk), ("rk_update_scalar", "grid%tke_1,", tk), ("bound_tke", 0, tk)]
            if c.specified or c.nested:
                prog.append(("flow_dep_bdy", "grid%tke_2,", tk))
    return prog


SCALAR_CASES = ["diffusion_periodic", "diff6_periodic", "rk2_everything", "specified_bcs_moist", "les", "smag2d", "diff2_periodic", "tke", "tke_periodic", "tke_hill", "specified_tke", "mono", "mono_periodic", "wenopd", "wenopd_periodic", "order64", "open_mapfac", "opts", "periodic_mapfac", "rk2_sym_x", "weno", "diffusion", "diff6", "specified", "nested", "ieva"]


@pytest.mark.parametrize("name", SCALAR_CASES)
@pytest.mark.parametrize("rk", [1, 3])
def test_rk_scalar_step(name, rk):
    # the first stage has code of its own only where something is added to the tendency there: diffusion, the 6th-order filter, the
    # boundary-zone tendencies
    if not FULL and (name not in CORE or (rk == 1 and name not in ("diffusion_periodic", "specified", "nested", "rk2_everything", "smag2d", "tke"))):
        pytest.skip("REFSRC_FULL=1 runs it; its digests are in the fixture")
    case = build(name)
    c = case.config
    rk = min(rk, c.rk_ord)
    if c.num_moist + c.num_scalar == 0:
        pytest.skip("no species")
    skip = ["h_tendency", "z_tendency", "xkmh", "xkmv", "xkhh", "xkhv"]
    if c.km_opt == 2 and c.num_scalar:
        # advect_tend is the per-species scratch rk_scalar_tend zeroes first: what it holds after the phase is the LAST species' flux
        # divergence, and solve_em.F steps the tke before the scalar family while the oracle keeps it as the family's last member
        skip.append("advect_tend")
    n, *_ = run(case, "rk_scalar_step", rk, scalar_program(c, rk, set(case.fields)), not_compared=skip)
    assert n >= 15


# ---------------------------------------------------------------------------------------------------------------------
# the physical boundary fills (share/module_bc.F set_physical_bc3d / 2d through module_bc_em.F's rk_phys_bc_dry_1 / _2): periodic,
# symmetric, open and specified edges; the WHOLE memory arrays are compared, halo and corners included

BC_CASES = ["open_mapfac", "periodic_mapfac", "b_wave_lid", "rk2_sym_x", "specified", "hill"]


@pytest.mark.parametrize("name", BC_CASES)
def test_rk_phys_bc_dry_1(name):
    prog = [("rk_phys_bc_dry_1", 0), ("set_physical_bc3d", "grid%rho,"), ("set_physical_bc3d", "grid%al,"), ("set_physical_bc3d", "grid%ph_2,'w'")]
    n, *_ = run(build(name), "rk_phys_bc_dry_1", None, prog, occurrence=1, whole=True)
    assert n >= 12


@pytest.mark.parametrize("name", BC_CASES)
def test_rk_phys_bc_dry_2(name):
    case = build(name)
    c = case.config
    prog = [("rk_phys_bc_dry_2", 0)]
    for fam, members, iv in (("moist", ["moist_" + q for q in ("qv", "qc", "qr")[:c.num_moist]], "im"),
                             ("scalar", ["scalar_%d" % n for n in range(1, c.num_scalar + 1)], "is")):
        for m in members:
            prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": m, "variable_in": "'p'"}}))
    n, *_ = run(case, "rk_phys_bc_dry_2", None, prog, whole=True)
    assert n >= 6


# ---------------------------------------------------------------------------------------------------------------------
# microphysics (solve_em.F:3623-4115): moist_physics_prep_em, phys/module_mp_kessler.F as microphysics_driver calls it, and
# moist_physics_finish_em -- expf / logf / powf / sqrtf through the host libm on both sides

@pytest.mark.parametrize("name", ["rain_open_mapfac", "rain_opts", "rain_specified", "rain_weno"])
def test_microphysics(name):
    case = build(name)
    c, d = case.config, case.dims
    assert c.mp_physics == 1
    sz = c.spec_zone if (c.specified or c.nested) else 0          # solve_em.F:3623-3645
    tile = dict(its=max(d.ips, d.ids + sz), ite=min(d.ipe, d.ide - 1 - sz), jts=max(d.jps, d.jds + sz), jte=min(d.jpe, d.jde - 1 - sz))
    alias = {"moist(ims,kms,jms,p_qv)": "moist_qv", "moist(ims,kms,jms,p_qc)": "moist_qc"}
    kess = {"t": "th_phy", "qv": "moist_qv", "qc": "moist_qc", "qr": "moist_qr", "pii": "pi_phy", "dt_in": "dt", "xlv": "xlv", "cp": "cp", "ep2": "ep_2",
            "kte": min(d.kde, d.kde - 1)}
    prog = [("moist_physics_prep_em", 0, {"alias": alias, "scalars": {"dtm": f32(c.dt), **tile}}),
            ("kessler", 0, {"actuals": kess, "scalars": tile}),
            ("moist_physics_finish_em", 0, {"alias": alias, "scalars": {"dtm": f32(c.dt), **tile}})]
    n, st, o = run(case, "microphysics", None, prog)
    assert n >= 20
    assert np.any(np.asarray(o.get("rainncv")) > 0) and np.any(np.asarray(o.get("moist_qc")) > 0)


# ---------------------------------------------------------------------------------------------------------------------
# the end of the step (solve_em.F:4175-4759): calc_p_rho_phi, the boundary fills of both time levels, spec_bdy_final on every
# field with boundary arrays, set_w_surface

def final_program(c, have):
    prog = [("calc_p_rho_phi", 1), ("set_phys_bc_dry_2", 0)]
    species = [("moist", "moist_" + q, "im", n + 2) for n, q in enumerate(("qv", "qc", "qr")[:c.num_moist])]
    species += [("scalar", "scalar_%d" % n, "is", n + 1) for n in range(1, c.num_scalar + 1)]
    for fam, m, iv, idx in species:
        prog.append(("set_physical_bc3d", 0, {"actuals": {"dat": m, "variable_in": "'p'"}}))
    if c.specified or c.nested:
        for f in ("grid%u_2,", "grid%v_2,") + (("grid%w_2,",) if c.nested else ()) + ("grid%t_2,", "grid%ph_2,", "grid%mu_2,"):
            prog.append(("spec_bdy_final", f))
        for fam, m, iv, idx in species:
            if (m + "_bxs") in have and (fam == "moist" or c.nested):
                prog.append(("spec_bdy_final", fam + "(ims,kms,jms," + iv + ")", {"alias": species_alias(fam, m, iv), "scalars": {iv: idx}}))
    prog.append(("set_w_surface", 0, {"scalars": {"fill_w_flag": False}}))
    return prog


@pytest.mark.parametrize("name", ["open_mapfac", "periodic_mapfac", "hill", "b_wave_lid", "specified", "nested"])
def test_end_of_step(name):
    case = build(name)
    n, *_ = run(case, "end_of_step", None, final_program(case.config, set(case.fields)))
    assert n >= 20


# ---------------------------------------------------------------------------------------------------------------------
# first_rk_step_part1 / part2 as far as the path goes: phy_prep, compute_diff_metrics and its boundary fills, cal_deform_and_div,
# calculate_km_kh (calculate_N2 + isotropic / 2-D and 3-D Smagorinsky / TKE closures), phy_bc, tke_rhs, vertical_diffusion_2,
# horizontal_diffusion_2 (module_diffusion_em.F)

PART1, PART2 = "dyn_em/module_first_rk_step_part1.F", "dyn_em/module_first_rk_step_part2.F"


def tke_alias(c):
    n = c.num_scalar + 1
    return {"grid%tke_2": f"scalar_{n}", "grid%tke_1": f"scalar_old_{n}", "tke_tend": f"scalar_tend_{n}"} if c.km_opt == 2 else {}


def diffusion_program(c):
    al = tke_alias(c)
    al.update({"moist(ims,kms,jms,p_qv)": "moist_qv"})
    o2 = {"path": PART2, "alias": al, "scalars": {"num_chem": 1, "num_tracer": 1}}
    prog = [("phy_prep", 0, {"path": PART1})]
    prog.append(("compute_diff_metrics", 0, o2))
    for text in ("grid%rdzw,", "grid%rdz,", "grid%z,", "grid%zx,", "grid%zy,"):
        prog.append(("set_physical_bc3d", text, o2))
    if c.diff_opt == 2:
        prog.append(("set_physical_bc2d", "grid%ustm,", o2))
    prog += [("cal_deform_and_div", 0, o2), ("calculate_km_kh", 0, o2), ("phy_bc", 0, o2)]
    if c.diff_opt == 2:
        if c.km_opt == 2:
            prog.append(("tke_rhs", 0, o2))
        prog += [("vertical_diffusion_2", 0, o2), ("horizontal_diffusion_2", 0, o2)]
    return prog


DIFFUSION_CASES = ["smag3d_pow", "smag2d_b_wave", "diff2_b_wave", "smag2d", "smag2d_hill", "diff2", "diff2_periodic", "diff2_hill", "smag3d", "smag3d_periodic", "tke", "tke_periodic", "tke_hill"]


@pytest.mark.parametrize("name", DIFFUSION_CASES)
def test_diffusion_prep(name):
    if not FULL and name not in CORE:
        pytest.skip("REFSRC_FULL=1 runs it; its digests are in the fixture")
    case = build(name)
    c = case.config
    # phy_prep's outputs the closures of this configuration never read are not materialised by the oracle (nor by the device)
    skip = ["pi_phy", "z_at_w", "dz8w"]
    if c.km_opt not in (2, 3):
        skip += ["th_phy", "p_phy", "t_phy", "bn2"]
    if c.diff_opt != 2:
        skip += ["rho"]
    n, *_ = run(case, "diffusion_prep", None, diffusion_program(c), not_compared=skip)
    assert n >= 25


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own kernel harness, PROGRAM feeder (module_advect_em.F:10567-10870, -DADVECT_KERNEL), with its five
# scalar_adv_opt settings: its advection kernels run from source inside the program's own loop (tests/feeder_ref.py), against
# the oracle's rk_scalar_step driven the same way.  The harness's output -- the integer columns of NNNNNN.txt -- must be
# reproduced exactly; the float rows agree to rounding (the program updates with dt*tendency/3., the model with (dt/3)*tendency).

FEEDER_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "feeder_columns.json")


@pytest.mark.parametrize("adv_opt", [0, 1, 2, 3, 4])
def test_feeder_harness(adv_opt):
    from tests.feeder_ref import Feeder
    write = os.environ.get("REFSRC_FEEDER") == "1"         # REFSRC_FEEDER=1: the program's first output file, 000200.txt (10 minutes)
    marks = (100, 200) if write else (2,)
    F = Feeder(adv_opt)
    ref = {}
    for n in range(1, marks[-1] + 1):
        F.loop()
        if n in marks:
            ref[n] = F.column()
    cols, row = oracle_feeder(adv_opt, marks[-1], marks)
    assert cols == ref, f"scalar_adv_opt {adv_opt}: the harness's output columns differ"
    assert np.abs(row - F.row()).max() < 1e-5 and np.abs(F.row()).max() > 0.9
    if write:
        have = json.load(open(FEEDER_GOLDEN)) if os.path.exists(FEEDER_GOLDEN) else {}
        have[str(adv_opt)] = {str(k): v for k, v in ref.items()}
        have["generator"] = "REFSRC_FEEDER=1 python -m pytest tests/test_reference_source.py -k feeder_harness"
        json.dump(have, open(FEEDER_GOLDEN, "w"), separators=(",", ":"), sort_keys=True)


# ---------------------------------------------------------------------------------------------------------------------
# the boundary fills solve_em.F makes around the acoustic loop, each taken from its own CALL in solve_em.F (a fill the oracle makes that
# solve_em.F does not have would not be found there), whole arrays compared

def fills(pairs):
    return [("set_physical_bc%dd" % rank, f"grid%{name},'{var}'") for name, var, rank in pairs]


@pytest.mark.parametrize("name", ["open_mapfac", "periodic_mapfac", "b_wave_lid", "rk2_sym_x", "hill"])
def test_fills_after_small_step_prep(name):
    """solve_em.F:1180-1260"""
    prog = [("small_step_prep", 0), ("calc_p_rho", 0), ("calc_coef_w", 0)] + fills(
        [("ru_tend", "u", 3), ("rv_tend", "v", 3), ("ph_2", "w", 3), ("al", "p", 3), ("p", "p", 3), ("t_1", "p", 3), ("t_save", "t", 3),
         ("mu_1", "t", 2), ("mu_2", "t", 2), ("mudf", "t", 2)])
    n, *_ = run(build(name), "small_step_prep", 2, prog, whole=True)
    assert n >= 40


@pytest.mark.parametrize("name", ["open_mapfac", "periodic_mapfac", "b_wave_lid", "rk2_sym_x", "hill"])
def test_fills_inside_the_acoustic_loop(name):
    """solve_em.F:1660-1720"""
    prog = [("calc_p_rho", 1)] + fills([("ph_2", "w", 3), ("al", "p", 3), ("p", "p", 3), ("muts", "t", 2), ("mu_2", "t", 2), ("mudf", "t", 2)])
    n, *_ = run(build(name), "calc_p_rho", None, prog, occurrence=2, extra={"iteration": 2}, whole=True)
    assert n >= 10


@pytest.mark.parametrize("name", ["open_mapfac", "periodic_mapfac", "b_wave_lid", "hill"])
def test_fills_after_the_last_small_step(name):
    """solve_em.F:1800-1850: ru_m, rv_m, ww_m, mut, muts for the scalar step"""
    case = build(name)
    prog = [("calc_mu_uv_1", 0), ("small_step_finish", 0)] + fills([("ru_m", "u", 3), ("rv_m", "v", 3), ("ww_m", "w", 3), ("mut", "t", 2), ("muts", "t", 2)])
    n, *_ = run(case, "small_step_finish", case.config.rk_ord, prog, whole=True)
    assert n >= 30


# ---------------------------------------------------------------------------------------------------------------------
# tiles: the reference's routines given ONE tile of the domain (grid%i_start(ij) ..., an interior tile and one on the domain edge) against
# the oracle's whole-domain phase on that tile's points -- the loop-range logic that depends on where a tile lies (its > ids + 3 switches the
# boundary treatment off, ...), which is what a patch of a decomposed run exercises

@pytest.mark.parametrize("tile", [(5, 9, 4, 8), (1, 4, 1, 10), (8, 12, 6, 10)], ids=["interior", "west", "north-east"])
@pytest.mark.parametrize("name,phase,rk", [("open_mapfac", "rk_tendency", 1), ("open_mapfac", "rk_tendency", 3), ("periodic_mapfac", "rk_tendency", 3),
                                           ("order64", "rk_tendency", 3), ("weno", "rk_tendency", 3), ("open_mapfac", "advance_uv", 3),
                                           ("open_mapfac", "advance_mu_t", 3), ("open_mapfac", "advance_w", 3), ("periodic_mapfac", "advance_uv", 3),
                                           ("open_mapfac", "rk_scalar_step", 1), ("periodic_mapfac", "rk_scalar_step", 1)])
def test_one_tile_of_the_reference(name, phase, rk, tile):
    case = build(name)
    d, c = case.dims, case.config
    if phase == "rk_tendency":
        prog = [p for p in tendency_program(c, rk)[0] if p[0] not in ("isotropic_km", "set_physical_bc3d")]
    elif phase == "rk_scalar_step":
        prog = [p for p in scalar_program(c, rk, set(case.fields)) if p[0] not in ("isotropic_km", "set_physical_bc3d")]
    else:
        prog = [(phase, 0)]
    o, args = at_phase(case, phase, rk)
    st = refrun.RefState(o)
    scal = rk_scalars(c, rk)
    i0, i1, j0, j1 = tile
    scal.update(rk_step=rk, rk_order=c.rk_ord, iteration=1)
    scal.update({"grid%i_start(ij)": i0, "grid%i_end(ij)": i1, "grid%j_start(ij)": j0, "grid%j_end(ij)": j1})
    for step in prog:
        opts = step[2] if len(step) > 2 else {}
        refrun.RefCall(step[0], step[1], opts.get("path", "dyn_em/solve_em.F"), opts.get("alias"), opts.get("actuals")).run(
            st, **{**scal, **opts.get("scalars", {})})
    before = {k: np.asarray(o.get(k)).copy() for k in st.fields if k not in st.absent}
    o.call(phase, *args)
    sl = (slice(j0 - d.jms, j1 - d.jms + 1), Ellipsis, slice(i0 - d.ims, i1 - d.ims + 1))
    bad, changed = [], 0
    for k in sorted(before):
        a, b = st.result(k), np.asarray(o.get(k))
        if a.ndim < 2:
            continue
        changed += not np.array_equal(a[sl], before[k][sl], equal_nan=True)
        if np.count_nonzero(~((a[sl] == b[sl]) | (np.isnan(a[sl]) & np.isnan(b[sl])))):
            bad.append(k)
    assert not bad and changed > 0, (bad, changed)
