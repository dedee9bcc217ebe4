#!/usr/bin/env python
"""run_shipped_case.py -- run a WRF ideal-case directory (namelist.input + input_sounding / input_jet, e.g. the reference's
test/em_quarter_ss) through this repo and write a wrfout file in the reference's layout, for tools/compare_wrfout.py.

    python tools/run_shipped_case.py tests/fixtures/em_quarter_ss --steps 10 --out wrfout_ours [--backend gpu|oracle]

backend gpu = the CUDA path (libwrfgpu.so through the C-ABI); backend oracle = the CPU restatement (test infrastructure;
useful where no GPU exists).  Records: the initial state (like wrf.exe's first history write) and the state after --steps."""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rundir, steps, out, backend="gpu", every=0):
    from wrf_b200 import shipped, wrfout
    case = shipped.init_domain(rundir)
    if backend == "gpu":
        from wrf_b200 import Solver
        m = Solver(case)
    else:
        from tests.oracle_lib import Oracle      # the checker, on request only
        m = Oracle(case)
    m.start_domain()
    names = wrfout.stream_fields(m.names(), "h")
    w = wrfout.WrfOut(out, case.dims, case.config)
    w.write(wrfout.time_string("0001-01-01_00:00:00", 0.0), {n: m.get(n) for n in names})
    done = 0
    while done < steps:
        n = min(every or steps, steps - done)
        m.step(n)
        done += n
        w.write(wrfout.time_string("0001-01-01_00:00:00", done * float(case.config.dt)), {n_: m.get(n_) for n_ in names})
    w.close()
    return case


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("rundir")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--every", type=int, default=0, help="also write a record every N steps")
    ap.add_argument("--out", default="wrfout_d01_wrf_b200")
    ap.add_argument("--backend", default="gpu", choices=["gpu", "oracle"])
    a = ap.parse_args()
    run(a.rundir, a.steps, a.out, a.backend, a.every)
    print(f"wrote {a.out}")


if __name__ == "__main__":
    main()
