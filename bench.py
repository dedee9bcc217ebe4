#!/usr/bin/env python
"""bench.py -- grid-point RK3 steps/s of the solve_em hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # CPU restatement on the host cores

One "step" is one full solve_em call (3 RK substeps, 1+ns/2+ns acoustic steps, scalar
advection, boundary fills) on the synthetic 2000x2000x80 em_quarter_ss-like domain of
BASELINE.json configs[4] (it fits one B200).  `value` is whole-job grid-pt-steps/s with the
state resident in HBM; `e2e` is the same metric through the drop-in call wrfgpu_solve_em
with HOST buffers (pinned) -- prognostic state uploaded and downloaded every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic 3-D field accesses (4 B each) per grid point per launch, SURVEY.md 8(d):
# distinct fields read + distinct fields written by the routine(s) the kernel implements.
ALGO_ACCESSES = {
    "k_advance_w": 19, "k_advance_uv": 14, "k_advance_mu_t": 11, "k_sumflux": 9, "k_calc_p_rho": 9,
    "k_tend_u": 6, "k_tend_v": 6, "k_tend_w": 6, "k_tend_t": 6, "k_tend_ph": 6,      # rk_tendency+addtend = 30 in all
    "k_ssp_3d": 26, "k_calc_coef_w": 5, "k_ssf": 18, "k_prep_3d": 13, "k_calc_ww": 3,
    "k_calc_p_rho_phi": 7, "k_advect_scalar": 5, "k_update_scalar": 3, "k_pd_scale": 6, "k_pd_tend": 7,
    "k_update_scalar_pd": 2, "k_zero3": 5, "k_invert_cqw": 2, "k_kessler": 20,
    # fused forms (what wrfgpu_step launches): DISTINCT fields of the fused routines, so lower than the sum of the parts
    "k_advance_mu_t<true>": 11, "k_advance_mu_t<false>": 11,
    "k_advance_uv<true>": 18,          # advance_uv 14 + ru_m, rv_m read and written
    "k_advance_w<true>": 25,           # advance_w 19 + calc_p_rho (al, p, pm1 r/w) + sumflux (ww_m r/w); alt, c2a, t_2, t_1 shared
    "k_advect_scalar_fast<theta>": 7, "k_advect_scalar_fast": 5,
    "k_mom_u_fast": 6, "k_mom_v_fast": 6, "k_mom_w_fast": 6, "k_ph_fast": 6,     # interior forms of k_tend_u/v/w/ph
    "k_pd_scale_fast": 6, "k_pd_tend_fast": 7,
    # the whole rk_tendency + rk_addtend_dry phase of the interior in one TMA-staged kernel: 24 fields in, 5 tendencies + cqw out
    "k_tend_tile": 30,
}


def algorithmic_bytes_per_point_step(n_m, n_s, ns, mp):
    """B = 4*[3*(104+n_m+8*n_s) + 62*N_ac + 20*mp]."""
    n_ac = 1 + ns // 2 + ns
    return 4 * (3 * (104 + n_m + 8 * n_s) + 62 * n_ac + 20 * (1 if mp else 0))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop = threading.Event()

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 3 + k and s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
def tile_case_on_device(small, nx, ny, nz, device, ntx=1, nty=1, rank=0):
    """Build this rank's patch of the big periodic state on the GPU by tiling a small periodic block
    (same levels, same column physics).  Every value is a pure function of the GLOBAL index, so any
    decomposition holds the same global state.  Returns (Case with empty fields, {name: tensor})."""
    import torch
    from wrf_b200 import decomp, ideal
    gd = ideal.memory_dims(nx, ny, nz)
    d = decomp.patch_dims(gd, ntx, nty, rank) if ntx * nty > 1 else gd
    c = decomp.patch_config(small.config, ntx, nty, rank)
    sd = small.dims
    per_i, per_j = sd.ide - sd.ids, sd.jde - sd.jds      # mass points of the block = period
    ii = torch.arange(d.ims, d.ime + 1, device=device)
    jj = torch.arange(d.jms, d.jme + 1, device=device)
    # periodic domain index -> block index -> block memory index
    imap = ((ii - 1) % (nx - 1)) % per_i + 1 - sd.ims
    jmap = ((jj - 1) % (ny - 1)) % per_j + 1 - sd.jms
    out = {}
    for k, a in small.fields.items():
        t = torch.from_numpy(a).to(device)
        if t.ndim == 3:
            t = t.index_select(0, jmap).index_select(2, imap).contiguous()
        elif t.ndim == 2:
            t = t.index_select(0, jmap).index_select(1, imap).contiguous()
        out[k] = t
    big = ideal.Case(name="synthetic", dims=d, config=c, fields={})
    return big, out


def state_checksum(s, names, dev, world):
    """Decomposition-independent checksum of the owned points of `names` + a finiteness check.  Every point contributes
    bits(value) x odd_hash(global i, k, j) in wrap-around 64-bit arithmetic; the per-rank sums add up to the same number
    for any patch mesh, so the N-GPU run can be compared with the 1-GPU run at the size that is timed (the reference's
    BIT4BIT criterion, tools/regtest.csh:14-18)."""
    import torch
    import torch.distributed as dist
    from wrf_b200 import decomp
    d = s.case.dims
    ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
    out, finite = {}, True
    for n in names:
        t = torch.empty(s.size(n), dtype=torch.float32, device=dev)
        s.get(n, out=t)
        t = t.view(nj, nk, ni) if t.numel() == ni * nk * nj else t.view(nj, 1, ni)
        ie = d.ipe if n in decomp.XSTAG else min(d.ipe, d.ide - 1)
        je = d.jpe if n in decomp.YSTAG else min(d.jpe, d.jde - 1)
        own = t[d.jps - d.jms:je - d.jms + 1, :, d.ips - d.ims:ie - d.ims + 1]
        finite = finite and bool(torch.isfinite(own).all().item())
        gj = torch.arange(d.jps, je + 1, device=dev, dtype=torch.int64)[:, None, None]
        gk = torch.arange(1, own.shape[1] + 1, device=dev, dtype=torch.int64)[None, :, None]
        gi = torch.arange(d.ips, ie + 1, device=dev, dtype=torch.int64)[None, None, :]
        h = ((gi * 73856093) ^ (gk * 19349663) ^ (gj * 83492791)) | 1
        bits = own.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        acc = (bits * h).sum()
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        out[n] = "%016x" % (int(acc.item()) & 0xFFFFFFFFFFFFFFFF)
    if world > 1:
        f = torch.tensor([1 if finite else 0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        finite = bool(f.item())
    return out, finite


def measured_traffic():
    """per-kernel DRAM bytes per grid point from the last committed `ncu --set full` capture"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from wrf_b200 import Solver, decomp, make_case

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if args.gpus != world:
            raise SystemExit("--gpus must equal WORLD_SIZE")
    ntx, nty = decomp.compute_mesh(world)        # mpaspect: 1x1, 1x2, 2x2, 2x4 (module_dm.F:98-116)

    nx, ny, nz = args.nx, args.ny, args.nz
    lat = 50
    moist = args.workload == "moist"
    if moist:   # the shipped em_quarter_ss physics: Kessler (qv, qc, qr), diff_opt 2 / km_opt 2 (prognostic tke), PD advection
        small = make_case("synthetic", nx=lat + 1, ny=lat + 1, nz=nz, bubble_lattice=lat * 2000.0, mp_physics=1, num_moist=3,
                          diff_opt=2, km_opt=2, mix_full_fields=1, damp_opt=2, zdamp=5000.0, dampcoef=0.003)
    else:
        small = make_case("synthetic", nx=lat + 1, ny=lat + 1, nz=nz, bubble_lattice=lat * 2000.0)
    case, dev_fields = tile_case_on_device(small, nx, ny, nz, dev, ntx, nty, rank)
    s = Solver(case, device=local, upload=False)
    stream = torch.cuda.current_stream()
    s.set_stream(stream.cuda_stream)
    if world > 1:
        ident = [Solver.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        s.comm_init(ident[0], rank, world)
    for k, t in dev_fields.items():
        s.set(k, t)
    del dev_fields
    torch.cuda.empty_cache()
    s.start_domain()
    s.sync()
    c = case.config
    npts = (nx - 1) * (ny - 1) * (nz - 1)          # the whole job's grid points (strong scaling)
    d = case.dims
    my_pts = (min(d.ipe, d.ide - 1) - d.ips + 1) * (min(d.jpe, d.jde - 1) - d.jps + 1) * (nz - 1)
    ns = c.time_step_sound
    n_adv = c.num_moist + c.num_scalar + (1 if (c.km_opt == 2 and c.diff_opt == 2) else 0)      # species advected (tke is one)
    bytes_pt = algorithmic_bytes_per_point_step(c.num_moist, n_adv, ns, c.mp_physics) + (584 if moist else 0)   # + DESIGN.md diffusion family

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-state throughput
    for _ in range(args.warmup):
        s.step(1)
    barrier()
    l0, hb0 = s.launch_count(), s.halo_bytes()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        s.step(1)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop.set()
    sampler.join(timeout=3)
    launches = s.launch_count() - l0
    halo_bytes = (s.halo_bytes() - hb0) / args.steps
    ms_per_step = ms / args.steps
    value = npts / (ms_per_step * 1e-3)
    # the state after exactly warmup + steps steps: finite, and the same numbers on any number of GPUs
    checksum, finite = state_checksum(s, ["u_2", "w_2", "t_2", "mu_2"], dev, world)
    if not finite:
        raise SystemExit("bench.py: the state is not finite after the timed steps -- no throughput number")

    # ---- per-kernel times (events on the same stream, outside the timed region)
    s.phase_timing(True)
    s.step(1)
    rep = s.kernel_report()
    s.phase_timing(False)
    barrier()
    tot = sum(v[1] for v in rep.values())
    # the dominant COMPUTE kernel (the exchange kernels' times include waiting for the neighbour)
    top = max(((k, v) for k, v in rep.items() if "halo" not in k), key=lambda kv: kv[1][1])
    peak, peak_src = measured_peak()
    kname, (kn, kms) = top
    acc = ALGO_ACCESSES.get(kname, ALGO_ACCESSES.get(kname.split("<")[0], 0))
    achieved = acc * 4 * my_pts / ((kms / kn) * 1e-3) / 1e9 if kn else 0.0
    tr = measured_traffic().get(kname, measured_traffic().get(kname.split("<")[0]))
    roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": (round(tr["dram_bytes_per_point"] * my_pts) if tr else None),
                "traffic_source": (tr.get("source") if tr else None), "peak_source": peak_src,
                "launches_per_step": kn, "avg_launch_ms": round(kms / kn, 4), "share_of_step": round(kms / tot, 4),
                "algorithmic_accesses_per_point": acc,
                "step": {"algorithmic_bytes_per_point": bytes_pt, "achieved_GBs_per_gpu": round(value * bytes_pt / 1e9 / world, 1),
                         "frac": round(value * bytes_pt / 1e9 / world / peak, 4)},
                "kernels_ms_per_step": {k: round(v[1], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])}}
    if world > 1:
        roofline["halo"] = {"bytes_sent_per_rank_per_step": int(halo_bytes),
                            "GBs_per_rank": round(halo_bytes / (ms_per_step * 1e-3) / 1e9, 2),
                            "nvlink_peak_GBs_per_direction": 770.0,
                            "transport": ("ncclSend/ncclRecv" if os.environ.get("WRFGPU_HALO") == "nccl" else
                                          "peer windows over CUDA IPC: packed straight into the neighbour's buffer by one kernel per phase")}

    # ---- end to end through the drop-in call with host buffers
    if args.no_e2e:      # tuning runs only: the line is not a bench result without e2e
        if rank == 0:
            print(json.dumps({"tuning_run": True, "workload": args.workload, "ms_per_step": round(ms_per_step, 3), "value": value,
                              "gpu_launches": int(launches), "step_frac": roofline["step"]["frac"], "checksum": checksum,
                              "kernels_ms_per_step": roofline["kernels_ms_per_step"]}), flush=True)
        s.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    prog = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2"] + [n for n in s.names() if n.startswith("moist_q")]
    host = {}
    for n in prog:
        host[n] = torch.empty(s.size(n), dtype=torch.float32, pin_memory=True)
        s.get(n, out=host[n])
    ptrs = [host[n].data_ptr() for n in prog]
    h2d = sum(host[n].numel() * 4 for n in prog)
    e2e_steps = max(1, min(args.steps, 3))
    s.solve_em(prog, ptrs, prog, ptrs)     # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.solve_em(prog, ptrs, prog, ptrs)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": npts / e2e_s, "unit": "grid-pt-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
           "ms_per_step": round(e2e_s * 1e3, 2),
           "call": "wrfgpu_solve_em (pinned host prognostic state of the rank's patch in and out every step; bytes are per rank)"}

    line = {"metric": "grid-pt RK3 steps/s", "value": value, "unit": "grid-pt-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic em_quarter_ss {nx}x{ny}x{nz} (e_we x e_sn x e_vert), dx=2km, dt=12s, ns={ns}, periodic x/y, "
                                   + ("Kessler + diff_opt=2/km_opt=2 (qv, qc, qr, tke; the shipped em_quarter_ss physics), " if moist else
                                      "dry (passive qv, moist_adv_opt=1), ") + f"{ntx}x{nty} patch mesh, 1 patch per GPU",
                       "grid_points": npts, "l2_policy": "working set (tens of GB per step) is far larger than the 126 MB L2",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "checksum": {"after_steps": args.warmup + args.steps, "finite": finite, "fields": checksum,
                         "note": "sum over owned points of bits(value) x odd_hash(global i,k,j) mod 2^64: identical for every patch mesh"}}
    s.close()
    del s
    torch.cuda.empty_cache()
    if world == 1 and not moist and not args.no_second_workload:
        line["second_workload"] = second_workload(args)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, quick=True)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def second_workload(args):
    """the shipped em_quarter_ss physics (Kessler + TKE closure) at the same size, a short resident-state measurement in a
    child process (fresh device memory); reported inside the dry line, it is not the headline number"""
    cmd = [sys.executable, os.path.abspath(__file__), "--workload", "moist", "--steps", "3", "--warmup", "3", "--no-e2e",
           "--nx", str(args.nx), "--ny", str(args.ny), "--nz", str(args.nz)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        top = sorted(d["kernels_ms_per_step"].items(), key=lambda kv: -kv[1])[:8]
        return {"workload": "synthetic em_quarter_ss, Kessler + diff_opt=2/km_opt=2 (qv, qc, qr, tke)", "ms_per_step": d["ms_per_step"],
                "value": d["value"], "unit": "grid-pt-steps/s", "step_roofline_frac": d["step_frac"], "steps": 3, "warmup": 3,
                "top_kernels_ms_per_step": dict(top), "checksum": d["checksum"]}
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


# ---------------------------------------------------------------------------
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(args, quick=False):
    """The CPU restatement (oracle/) on a bounded sample of the same workload."""
    from tests.oracle_lib import Oracle
    from wrf_b200 import make_case
    n = args.cpu_n
    lat = 50
    case = make_case("synthetic", nx=n + 1, ny=n + 1, nz=args.nz, bubble_lattice=lat * 2000.0)
    o = Oracle(case)
    threads = o.set_threads(args.cpu_threads or os.cpu_count() or 1)
    o.start_domain()
    o.step(1)
    steps = 8 if quick else max(1, args.steps)      # ~5-10 s of CPU work on the sample
    t0 = time.perf_counter()
    o.step(steps)
    dt = (time.perf_counter() - t0) / steps
    npts = n * n * (args.nz - 1)
    return {"value": npts / dt, "unit": "grid-pt-steps/s", "cores": threads, "kind": "port",
            "sample": f"oracle/ (C++ restatement, -O2 -ffp-contract=off, j-tiles on {threads} threads) on {n + 1}x{n + 1}x{args.nz} "
                      f"of the same synthetic workload, {steps} step(s), {dt:.2f} s/step",
            "host_cpus": os.cpu_count(), "cpu_model": cpu_model(),
            "label": "vs-port: the repo's own CPU restatement of solve_em on a cache-friendly sample, NOT wrf.exe (no Fortran compiler here)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(max(0, args.warmup - 1)):
        pass
    b = cpu_baseline(args, quick=False)
    ns = 6
    line = {"impl": "reference", "metric": "grid-pt RK3 steps/s", "value": b["value"], "unit": "grid-pt-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic em_quarter_ss {args.nx}x{args.ny}x{args.nz}, dx=2km, dt=12s, ns={ns}, periodic x/y, dry "
                                   f"(bounded sample: {b['sample']})"},
            "cpu_baseline": b, "e2e": {"value": b["value"], "unit": "grid-pt-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "no Fortran compiler exists in the build image, so the reference arm is the repo's CPU restatement of "
                    "solve_em (oracle/), not wrf.exe"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=2000)
    ap.add_argument("--nz", type=int, default=80)
    ap.add_argument("--cpu-n", type=int, default=200, help="mass points per side of the CPU sample")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: skip the end-to-end leg and print a short line")
    ap.add_argument("--workload", default="dry", choices=["dry", "moist"],
                    help="dry = the headline configuration (BASELINE.json configs[4]); moist = the same domain with the shipped em_quarter_ss physics")
    ap.add_argument("--no-second-workload", action="store_true", help="skip the short moist + TKE measurement appended to the 1-GPU line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
