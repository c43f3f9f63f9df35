"""Cold vs warm whole-solve time of cfg2 in one process (module load / graph instantiation show up in the first)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params
inst = sdpa.maxcut(100_000, 500_000, 3)
for k in range(3):
    S = Solver(inst)
    t = time.time()
    r = S.solve(default_params())
    print(k, "solve %.3f s (alm %.3f) wall %.3f its %d launches %d" % (r["solveSeconds"], r["almSeconds"], time.time() - t, r["almInnerIter"], r["kernelLaunches"]), flush=True)
    S.close()
