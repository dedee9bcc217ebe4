"""gpurun helper: print the phase-by-phase diff table for the small cases."""
import sys, time
sys.path.insert(0, ".")
from tests import parity

names = sys.argv[1:] or ["periodic", "hill2d_x", "quarter_ss", "b_wave"]
cs = parity.small_cases()
for n in names:
    t = time.time()
    print("=== case", n, flush=True)
    try:
        fails = parity.run_phase_parity(cs[n], verbose=False)
        print(f"    -> {len(fails)} failures in {time.time()-t:.1f}s", flush=True)
    except Exception as e:
        print("    EXC", repr(e), flush=True)
