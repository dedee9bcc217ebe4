"""A/B of an FMA-contracted build (FMAD=true bash wrf_b200/csrc/variant.sh fma) against the oracle: relative L2 error of the
prognostic fields after 1 and N steps.  Run with WRFGPU_LIB=wrf_b200/variants/libwrfgpu_fma.so (GPU box)."""
import sys

import numpy as np

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import Solver

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for name in ("quarter_ss", "periodic", "b_wave", "specified_real"):
    case = parity.small_cases()[name]
    o, s = Oracle(case), Solver(case)
    o.set_threads(8)
    o.start_domain(); s.start_domain()
    done = 0
    for target in (1, nsteps):
        o.step(target - done); s.step(target - done); done = target
        worst = max((parity.rel_l2_field(o, n, s.get(n), o.get(n)), n) for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "al"])
        nbad = sum(int((s.get(n) != o.get(n)).sum()) for n in ["u_2", "w_2", "t_2"])
        print(f"{name:16s} after {target:4d} steps: worst rel-L2 {worst[0]:.3e} ({worst[1]}), differing values in u, w, t: {nbad}, max|w| {np.abs(s.get('w_2')).max():.3f}", flush=True)
    s.close()
