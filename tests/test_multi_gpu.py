"""N ranks on the device: every rank's owned points equal the single-rank GPU run bit for bit
(the reference's BIT4BIT criterion, tools/regtest.csh:14-18).  The ranks are solvers of
one process exchanging through peer memory (wrfgpu_comm_init_local), so the test also runs on a
box with a single GPU; states start NaN-poisoned outside the owned points, so every value a rank
reads must have been delivered by a halo exchange or a boundary fill."""
import numpy as np
import pytest

import wrf_b200
from tests import parity
from wrf_b200 import Solver, decomp

pytestmark = pytest.mark.gpu


def _ndev():
    import ctypes
    n = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n))
    return max(1, n.value)


def _run(case, ntx, nty, nsteps, spread=False):
    ref = Solver(case)
    ref.start_domain(); ref.step(nsteps)
    patches = decomp.split_case(case, ntx, nty, poison=True)
    nd = _ndev() if spread else 1
    ranks = [Solver(p, device=r % nd) for r, p in enumerate(patches)]
    wrf_b200.comm_init_local(ranks)

    def go(s):
        s.start_domain()
        s.step(nsteps)
        s.sync()
    wrf_b200.run_ranks(ranks, go)
    names = [n for n in ref.names() if ref.size(n) > case.dims.kde and not decomp.bdy_side(n)]
    bad = []
    for s in ranks:
        pd = s.case.dims
        for n in names:
            a = decomp.owned(pd, s.get(n), name=n)
            b = decomp.owned(pd, ref.get(n), case.dims, name=n)
            if not np.array_equal(a, b, equal_nan=True):
                bad.append((s.case.name, n, int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())))
    sent = [s.halo_bytes() for s in ranks]
    for s in ranks:
        s.close()
    ref.close()
    return bad, sent


@pytest.mark.parametrize("mesh", [(1, 2), (2, 1), (2, 2)])
@pytest.mark.parametrize("name", ["quarter_ss", "periodic", "b_wave", "les", "diffusion", "mapfac", "mono_periodic", "wenopd_periodic", "weno_mom_periodic", "ieva_periodic", "diff6_periodic", "smag2d", "smag2d_periodic", "diff2", "diff2_periodic",
                                  "specified", "specified_real", "nested", "specified_tke"])
def test_n_rank_gpu_equals_single_rank(name, mesh):
    bad, sent = _run(parity.small_cases()[name], *mesh, nsteps=2)
    assert not bad, bad[:8]
    assert all(b > 0 for b in sent)


@pytest.mark.parametrize("mesh", [(1, 2), (2, 2)])
def test_shipped_b_wave_on_2_and_4_ranks(mesh):
    """BASELINE config C3: em_b_wave as shipped (41x81x65, periodic x / symmetric y, input_jet) at 2 and 4 ranks"""
    bad, sent = _run(parity.small_cases()["shipped_b_wave"], *mesh, nsteps=2)
    assert not bad, bad[:8]


def test_2x4_mesh_like_the_8_gpu_bench():
    """the mesh mpaspect gives 8 tasks (module_dm.F:98-116), on a periodic domain with uneven patches"""
    case = wrf_b200.make_case("synthetic", nx=43, ny=54, nz=13, bubble_lattice=20000.0, noise=0.05, num_scalar=1)
    bad, _ = _run(case, 2, 4, nsteps=2, spread=True)
    assert not bad, bad[:8]


def test_halo_is_neighbour_interior():
    """halo contents after an exchange are the neighbour's interior, bit for bit"""
    case = parity.small_cases()["periodic"]
    patches = decomp.split_case(case, 2, 2, poison=True)
    ranks = [Solver(p) for p in patches]
    wrf_b200.comm_init_local(ranks)
    wrf_b200.run_ranks(ranks, lambda s: (s.halo("HALO_EM_INIT"), s.sync()))
    gd = case.dims
    for s in ranks:
        pd = s.case.dims
        for n in ("t_2", "u_2", "phb", "mub"):
            a = s.get(n)
            g = case.fields[n]
            sj, si = decomp.window(gd, pd)
            gw = g[sj, :, si] if g.ndim == 3 else g[sj, si]
            # every in-domain point within 4 of the owned box (ghost zones beyond the domain are the
            # periodic closure's business and are covered by the whole-step tests)
            jj = np.arange(pd.jms, pd.jme + 1)
            ii = np.arange(pd.ims, pd.ime + 1)
            mj = (jj >= max(pd.jps - 4, gd.jds)) & (jj <= min(pd.jpe + 4, gd.jde - 1))
            mi = (ii >= max(pd.ips - 4, gd.ids)) & (ii <= min(pd.ipe + 4, gd.ide - 1))
            aa = a[mj][..., mi]
            gg = gw[mj][..., mi]
            assert np.array_equal(aa, gg), (s.case.name, n, int((aa != gg).sum()))
    for s in ranks:
        s.close()


@pytest.mark.parametrize("mesh", [(1, 2), (2, 2)])
@pytest.mark.parametrize("name", ["quarter_ss", "periodic", "tke_periodic"])
def test_peer_window_transport_between_solvers_of_one_process(name, mesh, monkeypatch):
    """the transport the multi-process runs use (k_halo_direct: packs into the neighbour's receive buffer, ready / ack
    counters in the neighbour's window) between solvers that share this GPU: same bits as the single-rank run"""
    monkeypatch.setenv("WRFGPU_HALO_DIRECT", "1")
    bad, sent = _run(parity.small_cases()[name], *mesh, nsteps=3)
    assert not bad, bad[:8]
    assert all(b > 0 for b in sent)


def test_processes_exchange_through_cuda_ipc_windows():
    """ranks as processes, one per GPU (what bench.py runs under torchrun and wrf.exe under mpirun): the neighbours' windows
    are mapped through CUDA IPC handles exchanged over NCCL.  Needs two GPUs."""
    import os, subprocess, sys
    if _ndev() < 2:
        pytest.skip("one GPU on this box: the IPC path needs one GPU per process")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, WRFGPU_HALO="direct")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(root, "tests", "ipc_rank_main.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("IPC_RANK_OK") == 2, r.stdout[-2000:]


@pytest.mark.parametrize("name", ["quarter_ss", "periodic"])
def test_exchange_overlapped_with_the_interior_of_its_producer(name, monkeypatch):
    """WRFGPU_OVERLAP=1: advance_uv / advance_w run on the ring of the patch first, HALO_EM_C / HALO_EM_C2 then travel on a
    priority stream while the interior is computed.  Same bits as the single-rank run."""
    monkeypatch.setenv("WRFGPU_OVERLAP", "1")
    bad, sent = _run(parity.small_cases()[name], 2, 2, nsteps=3)
    assert not bad, bad[:8]


def test_n_ranks_equal_one_rank_at_a_size_where_every_kernel_runs_its_interior_path():
    """the BIT4BIT criterion at 500 x 500 x 40 (the parity cases are at most 62 points wide: there the TMA-staged tendency
    kernel, the k-marching kernels and the halo transport only ever see a handful of tiles)"""
    case = wrf_b200.make_case("synthetic", nx=501, ny=501, nz=41, bubble_lattice=100000.0, num_scalar=1)
    bad, sent = _run(case, 2, 2, nsteps=2)
    assert not bad, bad[:8]
    assert all(b > 1.0e7 for b in sent)
