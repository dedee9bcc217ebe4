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
