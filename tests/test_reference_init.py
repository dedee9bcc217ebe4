"""wrf_b200/shipped.py (the host-side restatement of ideal.exe's sounding handling) against the reference's own routines of
dyn_em/module_initialize_ideal.F, executed from source (tests/fortran_subset.py) on the SHIPPED inputs (tests/fixtures/: unchanged copies
of test/em_*/input_sounding and test/em_b_wave/input_jet): get_sounding (:1767), get_sounding_b_wave (:1969) with calc_jet_sounding
(:2117).  Bit equality of every output level: the hydrostatic integration (ten iterations per level through x**cvpm), the dry
pressure integrated downwards, the jet's moisture profile (EXP) and its surface extrapolation.  Needs /root/reference."""
import os

import numpy as np
import pytest

from tests import refrun
from wrf_b200 import shipped

pytestmark = pytest.mark.skipif(not refrun.have_reference(), reason="needs the reference tree (/root/reference)")
f32 = np.float32
FIX = os.path.join(os.path.dirname(__file__), "fixtures")
PAIRS = (("zk", "zk"), ("p", "p"), ("p_dry", "pd"), ("theta", "theta"), ("rho", "rho"), ("u", "u"), ("v", "v"), ("qv", "qv"))


def differing(ref, mine, nl):
    out = {}
    for k_ref, k_me in PAIRS:
        n = int(np.count_nonzero(ref[k_ref][:nl].view(np.uint32) != np.asarray(mine[k_me], f32).view(np.uint32)))
        if n:
            out[k_ref] = n
    return out


@pytest.mark.parametrize("case", ["em_quarter_ss", "em_hill2d_x", "em_les"])
@pytest.mark.parametrize("dry", [False, True])
def test_get_sounding(case, dry):
    L = refrun.library(1)
    snd = shipped.read_sounding(os.path.join(FIX, case, "input_sounding"))

    def read_sounding(ps, ts, qvs, h, th, qv, u, v, n, nl, debug):      # read_sounding (:1929): the list-directed READs of the file
        k = len(snd["h"])
        h[:k], th[:k], qv[:k], u[:k], v[:k] = snd["h"], snd["th"], snd["qv"], snd["u"], snd["v"]
        return {0: f32(snd["ps"]), 1: f32(snd["ts"]), 2: f32(snd["qvs"]), 9: k}
    L.hooks["read_sounding"] = read_sounding
    nl_max = 1000
    a = {n: np.zeros(nl_max, f32) for n, _ in PAIRS}
    ret = L.routine("get_sounding")(dry=dry, nl_max=nl_max, nl_in=0, th_surf=f32(0), **a)
    mine = shipped.get_sounding(snd, dry)
    nl = ret["nl_in"]
    assert nl == len(mine["zk"]) >= 40 and f32(ret["th_surf"]) == f32(mine["th_surf"])
    assert not differing(a, mine, nl)
    assert a["p"][0] > 9.0e4 and a["p"][nl - 1] < a["p"][0] and (dry or case == "em_hill2d_x" or a["qv"][:nl].max() > 0)


@pytest.mark.parametrize("jp", [1, 17, 40, 41, 64, 80])
@pytest.mark.parametrize("dry", [False, True])
def test_get_sounding_b_wave(jp, dry):
    L = refrun.library(1)
    jet = shipped.read_input_jet(os.path.join(FIX, "em_b_wave", "input_jet"))
    nz, ny = jet["u"].shape
    nl_max = 1000
    a = {n: np.zeros(nl_max, f32) for n, _ in PAIRS}
    ret = L.routine("get_sounding_b_wave")(dry=dry, nl_max=nl_max, nl_in=0, u_jet=jet["u"].copy(), rho_jet=jet["rho"].copy(),
                                           th_jet=jet["th"].copy(), z_jet=jet["z"].copy(), nz_jet=nz, ny_jet=ny, j_point=jp, debug=False, **a)
    mine = shipped.get_sounding(shipped.jet_sounding(jet, jp, dry), dry)
    nl = ret["nl_in"]
    assert nl == nz == len(mine["zk"])
    assert not differing(a, mine, nl)
    assert a["u"][:nl].max() > 5.0 or jp in (1, 80)          # the jet


# ---------------------------------------------------------------------------------------------------------------------
# the whole initial state: init_domain_rk (dyn_em/module_initialize_ideal.F:70-1738, 1200 statements on TYPE(domain) components) and the
# first-trip diagnosis of start_domain_em (dyn_em/start_em.F:793-843: al from ph with the model's hypsometric option, p from the
# equation of state), executed from source for the shipped run directories, against wrf_b200/shipped.py: every field of the Case that
# solve_em starts from, bit for bit.

import json
import time

from tests.fortran_subset import FArr, module_parameters
from tests.refcases import digest
from wrf_b200 import ideal

INIT_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "shipped_init.json")
# Case field <- TYPE(domain) component the reference leaves it in
COMPARED = {"u_2": "u_2", "v_2": "v_2", "w_2": "w_2", "t_2": "t_2", "ph_2": "ph_2", "u_1": "u_1", "v_1": "v_1", "t_1": "t_1", "ph_1": "ph_1",
            "phb": "phb", "pb": "pb", "alb": "alb", "t_init": "t_init", "mu_1": "mu_1", "mu_2": "mu_2", "mub": "mub", "ht": "ht",
            "al": "al", "p": "p", "alt": "alt",
            "znw": "znw", "znu": "znu", "dnw": "dnw", "rdnw": "rdnw", "dn": "dn", "rdn": "rdn", "fnm": "fnm", "fnp": "fnp",
            "c1h": "c1h", "c2h": "c2h", "c3h": "c3h", "c4h": "c4h", "c1f": "c1f", "c2f": "c2f", "c3f": "c3f", "c4f": "c4f",
            "u_base": "u_base", "v_base": "v_base", "t_base": "t_base", "qv_base": "qv_base", "z_base": "z_base",
            "f": "f", "e": "e", "sina": "sina", "cosa": "cosa", "msftx": "msftx", "msfty": "msfty", "msfux": "msfux", "msfuy": "msfuy",
            "msfvx": "msfvx", "msfvx_inv": "msfvx_inv", "msfvy": "msfvy"}
SCALARS = ("p_top", "cf1", "cf2", "cf3", "cfn", "cfn1")


class Grid(dict):
    """TYPE(domain) as far as the two routines touch it: components are created on first use with the rank they are indexed with"""

    def __init__(self, d):
        super().__init__()
        self.d = d

    def make(self, comp, rank):
        d = self.d
        ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
        if rank is None:
            self[comp] = f32(0)
        elif rank == 1:
            self[comp] = FArr(np.zeros(max(nk, 1000), f32), (1,))
        else:
            shape, lo = {3: ((ni, nk, nj), (d.ims, d.kms, d.jms)), 2: ((ni, nj), (d.ims, d.jms)), 4: ((ni, nk, nj, 4), (d.ims, d.kms, d.jms, 1))}[rank]
            self[comp] = FArr(np.zeros(shape, f32), lo)


def reference_initial_state(name):
    """-> (grid, moist(:,:,:,P_QV), Config, Dims) after init_domain_rk and the first-trip block of start_domain_em"""
    L = refrun.library(1)
    L.objects["grid"] = {"*": "r", **{k: "i" for k in ("id", "step_number", "itimestep", "iswater", "isice", "num_traj", "imask_nostag") + tuple(
        a + b + c for a in ("s", "e") for b in ("d", "m", "p") for c in ("31", "32", "33"))}}
    L.objects["model_config_rec"] = {"*": "i", "eta_levels": "r"}
    # dummy_new_args.inc / dummy_new_decl.inc are generated at build time: the 4-D arrays the routine uses
    four_d = (["moist", "num_moist", "tracer", "num_tracer"], ["integer :: num_moist, num_tracer", "real, dimension(ims:ime,kms:kme,jms:jme,num_moist) :: moist",
                                                               "real, dimension(ims:ime,kms:kme,jms:jme,num_tracer) :: tracer"])
    L.inject["init_domain_rk"] = four_d
    for k, v in module_parameters(os.path.join(refrun.REF, "frame", "module_driver_constants.F")).items():
        L.constants.setdefault(k, v)
    L.constants["model_data_order"] = (5, "i")            # model_data_order.inc of the EM core: DATA_ORDER_XZY
    L.constants["tracer_test1"] = (2, "i")                # Registry.EM: package tracer_test1 tracer_opt==2
    L.globals.update({"p_tr17_%d" % n: 1 for n in range(1, 9)})
    mcr = L.globals.setdefault("model_config_rec", {})
    rundir = os.path.join(FIX, name)
    nl = shipped.read_namelist(os.path.join(rundir, "namelist.input"))
    c, gp = shipped.config_from_namelist(nl, 1)
    nx, ny, nz = gp["e_we"], gp["e_sn"], gp["e_vert"]
    d = ideal.memory_dims(nx, ny, nz)
    g = Grid(d)
    g["etac"] = f32(gp["etac"])
    g.update(id=1, sd31=1, ed31=nx, sd32=1, ed32=nz, sd33=1, ed33=ny, sm31=d.ims, em31=d.ime, sm32=d.kms, em32=d.kme, sm33=d.jms, em33=d.jme,
             sp31=1, ep31=nx, sp32=1, ep32=nz, sp33=1, ep33=ny)
    mcr.clear()
    mcr.update(ideal_case=gp["ideal_case"], eta_levels=FArr(np.full(1000, -1.0, f32), (1,)), e_vert=FArr(np.array([nz]), (1,)),
               sf_surface_physics=FArr(np.array([0]), (1,)), num_soil_layers=5)

    def model_to_grid_config_rec(idn, rec, cfg):
        cfg.update(dx=f32(c.dx), dy=f32(c.dy), ztop=f32(gp["ztop"]), hybrid_opt=int(gp["hybrid_opt"]), use_theta_m=int(c.use_theta_m),
                   mp_physics=int(c.mp_physics), hypsometric_opt=int(c.hypsometric_opt), tracer_opt=0, traj_opt=0, num_traj=0)
    nothing = lambda *a, **k: {}
    L.hooks.update(model_to_grid_config_rec=model_to_grid_config_rec, boundary_condition_check=nothing, wrf_dm_bcast_bytes=nothing,
                   wrf_dm_bcast_real=nothing, wrf_dm_bcast_integer=nothing, nl_get_iswater=lambda *a: {1: 0}, wrf_dm_on_monitor=lambda: True,
                   random_seed=nothing)
    for n in ("mminlu", "iswater", "cen_lat", "cen_lon", "truelat1", "truelat2", "moad_cen_lat", "stand_lon", "pole_lon", "pole_lat", "map_proj", "isice"):
        L.hooks["nl_set_" + n] = nothing
    if os.path.exists(os.path.join(rundir, "input_sounding")):
        snd = shipped.read_sounding(os.path.join(rundir, "input_sounding"))

        def read_sounding(ps, ts, qvs, h, th, qv, u, v, n, nl_, debug):
            k = len(snd["h"])
            h[:k], th[:k], qv[:k], u[:k], v[:k] = snd["h"], snd["th"], snd["qv"], snd["u"], snd["v"]
            return {0: f32(snd["ps"]), 1: f32(snd["ts"]), 2: f32(snd["qvs"]), 9: k}
        L.hooks["read_sounding"] = read_sounding
    else:
        jet = shipped.read_input_jet(os.path.join(rundir, "input_jet"))

        def read_input_jet(u, r, t, zk, nz_, ny_):
            u[...], r[...], t[...], zk[...] = jet["u"], jet["rho"], jet["th"], jet["z"]
            return {}
        L.hooks["read_input_jet"] = read_input_jet
    count = [0]

    def random_number(x):
        # RANDOM_NUMBER's stream is the compiler's: on both sides the same deterministic hash of the global (i, j) stands in for it
        # (em_les draws once per column, j outer, i inner, module_initialize_ideal.F:1378-1385)
        n = count[0]
        count[0] += 1
        return {0: f32(shipped._hash01(1 + n % (nx - 1), 1 + n // (nx - 1)))}
    L.hooks["random_number"] = random_number
    ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
    moist, tracer = np.zeros((ni, nk, nj, 4), f32), np.zeros((ni, nk, nj, 1), f32)
    refrun.set_species(c)
    L.routine("init_domain_rk")(grid=g, moist=moist, num_moist=4, tracer=tracer, num_tracer=1)
    # wrf.exe reads the _2 time level from wrfinput; start_domain_em's first trip copies it to _1 (start_em.F:643-668) ...
    sem = os.path.join(refrun.REF, "dyn_em", "start_em.F")
    dims = ["ims", "ime", "jms", "jme", "kms", "kme", "its", "ite", "jts", "jte", "kts", "kte", "ide", "jde"]
    L.add_fragment("start_em_copy_levels", sem, r"data that is expected to be zero must be explicitly initialized", r"^\s*END IF\s*$", ["grid"] + dims,
                   ["type(domain) :: grid", "integer :: " + ", ".join(dims) + ", i, j, k"], trim=1)
    g["imask_nostag"] = FArr(np.ones((d.ime - d.ims + 1, d.jme - d.jms + 1), np.int64), (d.ims, d.jms))
    L.routine("start_em_copy_levels")(grid=g, ims=d.ims, ime=d.ime, jms=d.jms, jme=d.jme, kms=d.kms, kme=d.kme, its=1, ite=nx, jts=1, jte=ny,
                                      kts=1, kte=nz, ide=nx, jde=ny)
    # ... and then "Use equations from calc_p_rho_phi to derive p and al from ph" (:793-843), up to the ENDIF of the rebalancing options
    L.add_fragment("start_em_first_trip", sem, r"Use equations from calc_p_rho_phi to derive p and al from ph",
                   r"ENDIF ! Various rebalancing options", ["grid", "config_flags", "moist", "num_moist", "ims", "ime", "jms", "jme", "kms", "kme",
                                                          "its", "ite", "jts", "jte", "kts", "kte", "ide", "jde"],
                   ["type(domain) :: grid", "type(grid_config_rec_type) :: config_flags",
                    "integer :: num_moist, ims, ime, jms, jme, kms, kme, its, ite, jts, jte, kts, kte, ide, jde, i, j, k",
                    "real :: pfu, pfd, phm, qvf", "real, dimension(ims:ime,kms:kme,jms:jme,num_moist) :: moist"], trim=1)
    cfg = refrun._Flags(hypsometric_opt=int(c.hypsometric_opt), use_theta_m=int(c.use_theta_m))
    L.routine("start_em_first_trip")(grid=g, config_flags=cfg, moist=moist, num_moist=4, ims=d.ims, ime=d.ime, jms=d.jms, jme=d.jme, kms=d.kms,
                                     kme=d.kme, its=1, ite=nx, jts=1, jte=ny, kts=1, kte=nz, ide=nx, jde=ny)
    return g, moist[..., L.globals["p_qv"] - 1], c, d


def reference_fields(g, qv, d):
    """the reference's arrays in the Case's layout, cut to the domain"""
    out = {}
    sj, si = slice(1 - d.jms, d.jde - d.jms + 1), slice(1 - d.ims, d.ide - d.ims + 1)
    for field, comp in list(COMPARED.items()) + [("moist_qv", None)]:
        a = qv if comp is None else g[comp].a
        a = a.transpose(2, 1, 0)[sj, :, si] if a.ndim == 3 else (a.T[sj, si] if a.ndim == 2 else a[:d.kde])
        out[field] = np.ascontiguousarray(a)
    return out


def shipped_fields(case):
    d, out = case.dims, {}
    sj, si = slice(1 - d.jms, d.jde - d.jms + 1), slice(1 - d.ims, d.ide - d.ims + 1)
    for field in list(COMPARED) + ["moist_qv"]:
        a = case.fields[field]
        out[field] = np.ascontiguousarray(a[sj, ..., si] if a.ndim >= 2 else a)
    return out


@pytest.mark.parametrize("name", ["em_quarter_ss", "em_hill2d_x", "em_les", "em_b_wave"])
def test_initial_state(name):
    write = os.environ.get("REFSRC_WRITE") == "1"
    if not (write or os.environ.get("REFSRC_FULL") == "1"):
        pytest.skip("25 000 - 216 000 points take 20 s - over a minute through the translator: REFSRC_FULL=1 runs it; the digests of what the "
                    "reference produced are in tests/golden/shipped_init.json and checked by test_reference_source_golden.py")
    g, qv, c, d = reference_initial_state(name)
    ref = reference_fields(g, qv, d)
    case = shipped.init_domain(os.path.join(FIX, name))
    mine = shipped_fields(case)
    bad = [k for k in ref if not np.array_equal(ref[k].view(np.uint32), mine[k].view(np.uint32))]
    assert not bad, f"{name}: fields of shipped.init_domain that differ from the reference's initial state: {bad}"
    for s in SCALARS:
        assert f32(g[s]) == f32(getattr(case.config, s)), s
    assert np.abs(ref["t_2"]).max() > 1 and ref["ph_2"].any() and ref["al"].any() and ref["p"].any()
    if write:
        have = json.load(open(INIT_GOLDEN)) if os.path.exists(INIT_GOLDEN) else {}
        have[name] = {k: digest(v) for k, v in ref.items()}
        have[name].update({s: float(f32(g[s])) for s in SCALARS})
        have["generator"] = "REFSRC_WRITE=1 python -m pytest tests/test_reference_init.py -k initial_state"
        json.dump(have, open(INIT_GOLDEN, "w"), separators=(",", ":"), sort_keys=True)
