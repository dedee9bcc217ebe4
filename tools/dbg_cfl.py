"""debug: the oracle's CFL maxima for the ieva_wdamp case, alone in a fresh process"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.parity import small_cases
from tests.oracle_lib import Oracle
import numpy as np
c = small_cases()["ieva_wdamp"]
print("u_2 sum", float(np.abs(c.fields["u_2"]).sum()), "w_2 max", float(c.fields["w_2"].max()))
o = Oracle(c); o.start_domain(); o.step(1); print("oracle alone", o.max_cfl())
try:
    from wrf_b200 import Solver
    o = Oracle(c); s = Solver(c); o.start_domain(); s.start_domain(); o.step(1); s.step(1); print("oracle with solver", o.max_cfl(), "solver", s.max_cfl())
except Exception as e:
    print("no solver:", e)
