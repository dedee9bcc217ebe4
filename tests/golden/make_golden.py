"""Regenerates tests/golden/oracle_step.json: per case and per prognostic field, the CRC32 of the
float32 bit patterns and the L2 norm of the ORACLE's state after start_domain + 2 solve_em steps.
The reference ships no golden outputs and cannot be built here (no Fortran compiler), so these are
regression fixtures of the restatement, not outputs of wrf.exe.   python tests/golden/make_golden.py"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import parity  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402

FIELDS = ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "p", "al", "ww", "ru_m"]
out = {"_libc": os.confstr("CS_GNU_LIBC_VERSION"), "_what": "oracle after start_domain + 2 steps: [crc32(float32 bits), l2]", "cases": {}}
for name, case in parity.small_cases().items():
    o = Oracle(case)
    o.start_domain(); o.step(2)
    names = FIELDS + [n for n in o.names() if n.startswith(("moist_q", "scalar_")) and "old" not in n and "tend" not in n]
    rec = {}
    for n in names:
        a = o.get(n)
        rec[n] = [zlib.crc32(np.ascontiguousarray(a, dtype=np.float32).tobytes()) & 0xFFFFFFFF,
                  float(np.sqrt((a.astype(np.float64) ** 2).sum()))]
    out["cases"][name] = rec
with open(os.path.join(ROOT, "tests", "golden", "oracle_step.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["cases"]), "cases")
