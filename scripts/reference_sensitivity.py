"""How far do the REFERENCE's own whole-solve results move under a last-bit perturbation?

The same SDP is solved twice by the untouched reference (oracle/_ref): once as generated, once with its constraints
listed in a different order (b permuted alike).  In exact arithmetic both runs are the same computation -- the
constraint order only changes the order of the floating-point sums over constraints (sum_i w_i A_i, the m-vector
dots) -- so the difference between the two results is what rounding alone does to a LoRADS solve.
usage: python scripts/reference_sensitivity.py [case ...]"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from lorads_b200 import sdpa  # noqa: E402


def permuted(inst, seed):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(inst.m)
    cones = []
    for cone in inst.cones:
        counts = np.diff(cone.beg)
        take = np.concatenate([np.arange(cone.beg[0], cone.beg[1])] + [np.arange(cone.beg[1 + p], cone.beg[2 + p]) for p in perm])
        beg = np.concatenate([[0], np.cumsum([counts[0]] + [counts[1 + p] for p in perm])]).astype(np.int64)
        cones.append(sdpa.Cone(n=cone.n, beg=beg, idx=cone.idx[take], elem=cone.elem[take]))
    return sdpa.Instance(m=inst.m, b=inst.b[perm], cones=cones, name=inst.name + "_perm")


def solve_ref(inst, bits=32):
    from oracle import ref
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    path = os.path.join(tempfile.mkdtemp(prefix="lorads_sens_"), inst.name + ".dat-s")
    sdpa.write_dat_s(inst, path)
    saved = os.dup(1)
    os.dup2(2, 1)                      # the reference prints its log with printf
    try:
        return ref.RefSolver(path, bits).solve()
    finally:
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)


def sensitivity(case, seed=1):
    from make_golden import CASES, build_instance
    kind, kw = CASES[case]
    inst = build_instance(kind, kw)
    a, b = solve_ref(inst), solve_ref(permuted(inst, seed))
    scale = 1.0 + abs(a["pobj"])
    return dict(case=case, pobj=(a["pobj"], b["pobj"]), dobj=(a["dobj"], b["dobj"]),
                rel_pobj=abs(a["pobj"] - b["pobj"]) / scale, rel_dobj=abs(a["dobj"] - b["dobj"]) / scale,
                alm_inner=(a["alm_inner"], b["alm_inner"]), admm_iter=(a["admm_iter"], b["admm_iter"]),
                status=(a["status"], b["status"]), gap=(a["gap"], b["gap"]), pinf=(a["pinf"], b["pinf"]))


if __name__ == "__main__":
    cases = sys.argv[1:] or ["maxcut_n120", "maxcut_n800", "mcomp_60x50", "theta_n60", "twoblock", "theta_n200"]
    for c in cases:
        print(json.dumps(sensitivity(c)), flush=True)
