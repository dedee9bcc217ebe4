"""One rank of tests/test_multi_gpu.py::test_processes_exchange_through_cuda_ipc_windows (launched by torch.distributed.run):
one process per GPU, halos through CUDA IPC peer windows; owned points must equal the single-rank device run bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity  # noqa: E402
from wrf_b200 import Solver, decomp  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for name in ("quarter_ss", "periodic", "tke_periodic"):
    case = parity.small_cases()[name]
    ntx, nty = decomp.compute_mesh(world)
    ref = Solver(case, device=local)
    ref.start_domain(); ref.step(3)
    mine = decomp.split_case(case, ntx, nty, poison=True)[rank]
    s = Solver(mine, device=local)
    ident = [Solver.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    s.comm_init(ident[0], rank, world)
    s.start_domain(); s.step(3); s.sync()
    for n in [m for m in ref.names() if ref.size(m) > case.dims.kde]:
        a = decomp.owned(mine.dims, s.get(n), name=n)
        b = decomp.owned(mine.dims, ref.get(n), case.dims, name=n)
        assert np.array_equal(a, b, equal_nan=True), (name, rank, n, int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum()))
    assert s.halo_bytes() > 0
    s.close(); ref.close()
    dist.barrier()
print("IPC_RANK_OK", rank, flush=True)
dist.destroy_process_group()
