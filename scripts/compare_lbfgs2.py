import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params
tot = 0
for seed in range(10, 18):
    r = Solver(sdpa.maxcut(100000, 500000, seed)).solve(default_params())
    tot += r["almInnerIter"]
    print("seed", seed, "vf" if os.environ.get("LORADS_B200_VF_LBFGS") else "exact", r["almInnerIter"], r["almOuterIter"], "%.8e" % r["pObj"], "%.2e %.2e" % (r["pInfeasL1"], r["pdGap"]), r["status"], "alm %.2fs" % r["almSeconds"], flush=True)
print("TOTAL inner", tot)
