"""exact two-loop vs vector-free L-BFGS: iteration counts / objectives on a few instances (run under each env)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params
cases = [("mc20k_s%d" % s, lambda s=s: sdpa.maxcut(20000, 100000, s)) for s in (1, 2, 3)] + \
        [("mc5k_s%d" % s, lambda s=s: sdpa.maxcut(5000, 40000, s)) for s in (4, 5)] + \
        [("mc100k_s%d" % s, lambda s=s: sdpa.maxcut(100000, 500000, s)) for s in (3, 6)] + \
        [("mcomp", lambda: sdpa.matrix_completion(2000, 2000, 100000, 3, 7)), ("theta1000", lambda: sdpa.lovasz_theta(1000, 8000, 5))]
for name, make in cases:
    r = Solver(make()).solve(default_params())
    print(name, "vf" if os.environ.get("LORADS_B200_VF_LBFGS") else "exact", r["almInnerIter"], r["admmIter"], r["cgIter"], "%.8e %.8e" % (r["pObj"], r["dObj"]), "%.2e %.2e" % (r["pInfeasL1"], r["pdGap"]), r["status"], "%.2fs" % r["solveSeconds"], flush=True)
