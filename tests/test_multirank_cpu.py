"""The N>1 path on CPU: decomposition rules, exchange tables and index boxes of the product's
host logic, exercised by running the ORACLE on N ranks (threads, and world_size-2 gloo
processes) with halos moved exactly as the product's tables prescribe.  Bar: every rank's owned
points equal the single-rank run bit for bit (tools/regtest.csh BIT4BIT), starting from states
whose non-owned points are NaN-poisoned."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import halo_emul, parity
from tests.oracle_lib import Oracle
from wrf_b200 import decomp, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROG = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "al", "u_1", "t_1", "ww", "ru_m"]


def test_compute_mesh_is_mpaspect():
    """module_dm.F:98-116: most square factorisation with x <= y"""
    assert decomp.compute_mesh(1) == (1, 1)
    assert decomp.compute_mesh(2) == (1, 2)
    assert decomp.compute_mesh(4) == (2, 2)
    assert decomp.compute_mesh(8) == (2, 4)
    assert decomp.compute_mesh(6) == (2, 3)
    assert decomp.compute_mesh(8, nproc_x=4) == (4, 2)
    with pytest.raises(ValueError):
        decomp.compute_mesh(8, nproc_x=3)


def test_patches_tile_the_domain():
    """task_for_point.c:6-14: every point has exactly one owner; remainders go to both ends"""
    case = make_case("quarter_ss", nx=43, ny=38, nz=9)
    for ntx, nty in [(1, 2), (2, 2), (3, 2), (2, 4), (5, 3)]:
        cover = np.zeros((case.dims.jde + 1, case.dims.ide + 1), int)
        sizes_x, sizes_y = set(), set()
        for r in range(ntx * nty):
            pd = decomp.patch_dims(case.dims, ntx, nty, r)
            cover[pd.jps:pd.jpe + 1, pd.ips:pd.ipe + 1] += 1
            sizes_x.add(pd.ipe - pd.ips + 1); sizes_y.add(pd.jpe - pd.jps + 1)
            assert pd.ims <= pd.ips - 4 and pd.ime >= pd.ipe + 4 and pd.jms <= pd.jps - 4 and pd.jme >= pd.jpe + 4
        assert (cover[1:, 1:] == 1).all()
        assert max(sizes_x) - min(sizes_x) <= 1 and max(sizes_y) - min(sizes_y) <= 1


def test_halo_boxes_pair_up():
    """what rank A packs towards B has the shape of what B fills from A, in both phases, for open and periodic meshes"""
    for name, kw in [("quarter_ss", {}), ("periodic", {})]:
        case = parity.small_cases()[name]
        for ntx, nty in [(1, 2), (2, 1), (2, 2), (3, 2)]:
            patches = decomp.split_case(case, ntx, nty)
            for w in (1, 2, 3, 4):
                for pc in patches:
                    for q in range(4):
                        sb = halo_emul.halo_box(pc.dims, pc.config, w, q, True)
                        if sb is None:
                            continue
                        peer = patches[sb[0]]
                        rb = halo_emul.halo_box(peer.dims, peer.config, w, q ^ 1, False)
                        assert rb is not None and rb[0] == pc.config.mytask
                        assert (sb[2] - sb[1], sb[4] - sb[3]) == (rb[2] - rb[1], rb[4] - rb[3]), (name, ntx, nty, w, q)
                        assert rb[1] >= peer.dims.ims and rb[2] <= peer.dims.ime and rb[3] >= peer.dims.jms and rb[4] <= peer.dims.jme


def _check(case, ntx, nty, nsteps=2):
    ref = Oracle(case)
    ref.start_domain(); ref.step(nsteps)
    patches = decomp.split_case(case, ntx, nty, poison=True)
    oracles, emus = halo_emul.run_ranks_in_threads(patches, nsteps, Oracle)
    names = PROG + [n for n in ref.names() if n.startswith(("moist_q", "scalar_")) and "old" not in n and "tend" not in n and not decomp.bdy_side(n)]
    names += [n for n in ("xkmh", "xkhh", "defor11", "defor22", "defor12", "defor13", "defor23", "div", "z", "rdz", "rdzw", "zx", "zy",
                          "ru_tendf", "rv_tendf", "rw_tendf", "t_tendf", "bn2") if n in ref.names()]
    for o in oracles:
        pd = o.case.dims
        for n in names:
            mine = decomp.owned(pd, o.get(n), name=n)
            glob = decomp.owned(pd, ref.get(n), case.dims, name=n)
            assert np.array_equal(mine, glob, equal_nan=False), (o.case.name, n, int((mine != glob).sum()))
    assert "HALO_EM_C2" in emus[0].calls and "HALO_EM_INIT" in emus[0].calls


@pytest.mark.parametrize("mesh", [(1, 2), (2, 1), (2, 2)])
@pytest.mark.parametrize("name", ["quarter_ss", "periodic", "b_wave", "hill2d_x_wide", "les", "diffusion", "diffusion_periodic", "mono", "mono_periodic", "weno", "wenopd_periodic", "weno_mom", "ieva", "ieva_periodic", "diff6", "diff6_periodic", "order64",
                                  "smag2d", "smag2d_periodic", "smag2d_b_wave", "diff2", "diff2_periodic", "diff2_hill_wide", "diff2_b_wave", "smag3d", "smag3d_periodic", "tke", "tke_periodic",
                                  "specified", "specified_real", "nested", "specified_bcs_moist", "specified_tke"])
def test_n_rank_oracle_equals_single_rank(name, mesh):
    cases = parity.small_cases()
    cases["hill2d_x_wide"] = make_case("hill2d_x", nx=62, ny=9, nz=21)
    cases["diff2_hill_wide"] = parity.with_surface_fluxes(make_case("hill2d_x", nx=62, ny=9, nz=21, diff_opt=2, km_opt=1, khdif=100.0, kvdif=30.0, isfflx=1))
    _check(cases[name], *mesh)


def test_three_by_two_mesh_with_uneven_patches():
    _check(make_case("quarter_ss", nx=29, ny=27, nz=15), 3, 2, nsteps=1)


def test_world_size_2_gloo():
    """the same check with the ranks as two processes exchanging over torch.distributed / gloo"""
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "gloo_rank_main.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("GLOO_RANK_OK") == 2, r.stdout[-2000:]
