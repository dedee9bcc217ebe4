"""One rank of tests/test_multirank_cpu.py::test_world_size_2_gloo (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import halo_emul, parity  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402
from wrf_b200 import decomp  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for name in ("quarter_ss", "periodic"):
    case = parity.small_cases()[name]
    ntx, nty = decomp.compute_mesh(world)
    ref = Oracle(case)
    ref.start_domain(); ref.step(2)
    mine = decomp.split_case(case, ntx, nty, poison=True)[rank]
    o = Oracle(mine)
    emu = halo_emul.HaloEmulator(o, halo_emul.GlooTransport())
    o.start_domain(); o.step(2)
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "moist_qv"]:
        a = decomp.owned(mine.dims, o.get(n), name=n)
        b = decomp.owned(mine.dims, ref.get(n), case.dims, name=n)
        assert np.array_equal(a, b), (name, rank, n, int((a != b).sum()))
    dist.barrier()
print("GLOO_RANK_OK", rank, flush=True)
dist.destroy_process_group()
