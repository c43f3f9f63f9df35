"""Larger mixed SDP + LP solve (timing of the LP kernels at scale)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params
n, e, nlp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
inst = sdpa.add_lp_block(sdpa.maxcut(n, e, 2), nlp, 3)
t = time.time(); S = Solver(inst); print("setup %.2f s, rank %d, LP cols %d" % (time.time() - t, S.rank(), S.info(19)), flush=True)
t = time.time(); r = S.solve(default_params(timeSecLimit=120.0))
print(json.dumps({k: r[k] for k in ("pObj", "dObj", "pInfeasL1", "dInfeasL1", "pdGap", "almInnerIter", "admmIter", "cgIter", "solveSeconds", "almSeconds", "admmSeconds", "status")}))
x = S.get_lp("R") ** 2
print("LP x: min %.3e max %.3e nonzero(>1e-8) %d" % (x.min(), x.max(), int((x > 1e-8).sum())))
