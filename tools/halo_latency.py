"""per-exchange latency of the halo transport, ranks in lockstep (torchrun --nproc-per-node N tools/halo_latency.py)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import tile_case_on_device
from wrf_b200 import Solver, decomp, make_case
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ntx, nty = decomp.compute_mesh(world)
small = make_case("synthetic", nx=51, ny=51, nz=80, bubble_lattice=100000.0)
case, f = tile_case_on_device(small, 2000, 2000, 80, dev, ntx, nty, rank)
s = Solver(case, device=local, upload=False)
stream = torch.cuda.current_stream(); s.set_stream(stream.cuda_stream)
ident = [Solver.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ident, src=0); s.comm_init(ident[0], rank, world)
for k, t in f.items(): s.set(k, t)
s.start_domain(); s.step(1); s.sync()
out = {}
for name in ["HALO_EM_C", "HALO_EM_C2", "HALO_EM_A", "HALO_EM_B", "HALO_EM_D3_5"]:
    for _ in range(20): s.halo(name)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0 = s.halo_bytes()
    e0.record(stream)
    for _ in range(200): s.halo(name)
    e1.record(stream); torch.cuda.synchronize()
    out[name] = {"us_per_exchange": round(e0.elapsed_time(e1) * 1000 / 200, 1), "bytes_per_exchange": (s.halo_bytes() - b0) // 200}
if rank == 0: print(json.dumps({"world": world, "mode": os.environ.get("WRFGPU_HALO", "direct"), "halo": out}), flush=True)
s.close(); dist.barrier(); dist.destroy_process_group()
