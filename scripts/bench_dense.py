"""Timing of the dense-scratch kernels (cones with a genuinely dense objective)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lorads_b200 import sdpa
from lorads_b200.capi import Solver
n, e = int(sys.argv[1]), int(sys.argv[2])
tlr = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
inst = sdpa.lovasz_theta(n, e, 5)
c = inst.cones[0]
elem = c.elem.copy()
elem[: c.beg[1]] += 0.01 * np.random.default_rng(0).standard_normal(c.beg[1])     # dense, not constant
inst.cones[0] = sdpa.Cone(n=c.n, beg=c.beg, idx=c.idx, elem=elem)
S = Solver(inst, times_log_rank=tlr)
print("n", n, "rank", S.rank(), "dense", S.info(6), "psize", S.info(4))
names = {0: "dual pass: dense_uvt_dual + 2x auv<FROMZ>", 1: "A-only: dense_uvt + auv<FROMZ>", 2: "dense_wsum", 3: "dense_symm"}
for w in (0, 1, 2, 3):
    print(f"{names[w]:45s} {S.bench_kernel(w, 20)*1e3:10.1f} us")
rho = S.dinfo(6)
S.alm_prepare(rho)
sec, done = S.time_alm_inner_iters(rho, 100)
print("inner iteration", sec / done * 1e3, "ms")
