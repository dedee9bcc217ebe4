"""gpurun helper: phase-by-phase diff table and whole-step diff for the small cases."""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import Solver

names = sys.argv[1:] or ["periodic", "hill2d_x", "quarter_ss", "quarter_ss_nopd", "b_wave"]
cs = parity.small_cases()
for n in names:
    t = time.time()
    print("=== case", n, flush=True)
    try:
        fails = parity.run_phase_parity(cs[n], verbose=True)
        print(f"    -> {len(fails)} failures in {time.time()-t:.1f}s", flush=True)
        for nsteps in (1, 10):
            o, s = Oracle(cs[n]), Solver(cs[n])
            o.start_domain(); s.start_domain()
            o.step(nsteps); s.step(nsteps)
            rep = parity.diff_report(o, s)
            print(f"    full {nsteps} step(s): {len(rep)} fields differ")
            for r in rep:
                print(f"       {r[0]:14s} n={r[1]:7d} maxabs={r[2]:.3e} rel_l2={r[3]:.3e}")
            s.close()
    except Exception as e:
        print("    EXC", repr(e), flush=True)
