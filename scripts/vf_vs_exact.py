import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lorads_b200 import sdpa
from lorads_b200.capi import Solver
n, e, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
inst = sdpa.maxcut(n, e, seed)
os.environ.pop("LORADS_B200_VF_LBFGS", None)
A = Solver(inst)
os.environ["LORADS_B200_VF_LBFGS"] = "1"
B = Solver(inst)
rho = A.dinfo(6)
A.alm_prepare(rho); B.alm_prepare(rho)
for k in range(int(sys.argv[4])):
    ra, oa = A.alm_inner_iter(rho, k); rb, ob = B.alm_inner_iter(rho, k)
    if k < 12 or k % 10 == 0:
        Ra, Rb = A.get_factor("R"), B.get_factor("R")
        Da, Db = A.get_factor("U"), B.get_factor("U")
        print(k, ra, rb, "tau %.15e %.15e" % (oa["tau"], ob["tau"]), "R rel %.2e" % (np.linalg.norm(Ra - Rb) / np.linalg.norm(Ra)), "D rel %.2e" % (np.linalg.norm(Da - Db) / np.linalg.norm(Da)), flush=True)
