"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): vertex-centric path (MaxCut,
matrix completion, Lovasz theta with rank-one objective and a residual constraint), generic item path (forced), dense
scratch path, LP block, ADMM/CG, rank augmentation, dual infeasibility."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from lorads_b200 import sdpa  # noqa: E402
from lorads_b200.capi import Solver, default_params  # noqa: E402


def exercise(inst, label, solve=True, **kw):
    S = Solver(inst, device=0, **kw)
    rho = S.dinfo(6)
    S.auv("U", "V", with_obj=True)
    S.wsum_mulrk(np.random.default_rng(0).standard_normal(inst.m), True, "V")
    S.alm_prepare(rho)
    for k in range(3):
        S.alm_inner_iter(rho, k)
    S.update_sdp_var_one("V", "U", 0.5, 1e-6, 30, 0)
    S.dual_infeasibility()
    S.close()
    if solve:
        S = Solver(inst, device=0, **kw)
        p = default_params()
        p.maxALMIter = 12
        p.maxADMMIter = 30
        res = S.solve(p)
        S.close()
        print(label, "status", res["status"], "pObj", res["pObj"], flush=True)
    else:
        print(label, "ok", flush=True)


exercise(sdpa.maxcut(300, 1500, 1), "maxcut")
exercise(sdpa.matrix_completion(60, 50, 700, 3, 7), "mcomp")
exercise(sdpa.lovasz_theta(120, 600, 5), "theta rank-one + residual")
exercise(sdpa.lovasz_theta(40, 150, 4), "theta dense scratch")
exercise(sdpa.maxcut(200, 900, 3), "maxcut rank 40", times_log_rank=7.6)
os.environ["LORADS_B200_VC"] = "0"
exercise(sdpa.maxcut(300, 1500, 1), "maxcut generic item path")
os.environ.pop("LORADS_B200_VC")
exercise(sdpa.add_lp_block(sdpa.maxcut(60, 200, 1), 25, 7), "maxcut + LP block")
print("SANITIZER_TARGET_DONE")
