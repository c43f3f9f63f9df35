"""The C restatement (oracle/lorads_oracle.c) against the golden vectors produced by the untouched reference
(tests/golden/make_golden.py).  CPU only.  Also pins the host-side logic of the CUDA library (pre-solve,
rank rule, line search) against the same vectors."""
import json

import numpy as np
import pytest

from conftest import rel_err
from lorads_b200 import capi
from oracle import restate

TOL = 1e-12


def make_oracle(g, inst):
    O = restate.OracleSolver(inst)
    for c in range(O.n_cones):
        assert O.rank(c) == int(g["ranks"][c])
        for f in "RUV":
            # the libc rand() stream reproduces the reference start exactly; the stored factors are authoritative
            assert np.array_equal(O.factor(f, c), g[f"{f}{c}"])
    return O


def test_constants(golden):
    name, g, inst = golden
    O = make_oracle(g, inst)
    assert np.allclose([O.dinfo(k) for k in range(6)], g["norms"], rtol=1e-14, atol=0)
    assert O.dinfo(6) == float(g["rho0"])
    for c in range(O.n_cones):
        assert O.info(10, c) == int(g["rank_max"][c])
        assert O.info(6, c) == int(g[f"dense{c}"])
        if not O.info(6, c):
            pr, pc = O.pattern(c)
            assert np.array_equal(pr, g[f"prow{c}"]) and np.array_equal(pc, g[f"pcol{c}"])


def test_operators(golden):
    name, g, inst = golden
    O = make_oracle(g, inst)
    for c in range(O.n_cones):
        assert rel_err(O.auv("R", "R", c), g[f"auv_RR{c}"]) < TOL
        assert rel_err(O.auv("U", "V", c), g[f"auv_UV{c}"]) < TOL
        assert abs(O.obj_auv("R", "R", c) - float(g[f"obj_RR{c}"])) <= TOL * max(1.0, abs(float(g[f"obj_RR{c}"])))
        assert abs(O.obj_auv("U", "V", c) - float(g[f"obj_UV{c}"])) <= TOL * max(1.0, abs(float(g[f"obj_UV{c}"])))
        assert rel_err(O.wsum_mulrk(g["w"], True, "V", c), g[f"wsum_C{c}"]) < TOL
        assert rel_err(O.wsum_mulrk(g["w"], False, "V", c), g[f"wsum_noC{c}"]) < TOL
        assert rel_err(O.cg_matvec(g[f"cgx{c}"], "V", c), g[f"cgmv{c}"]) < TOL


def test_gradient_and_cg(golden):
    name, g, inst = golden
    O = make_oracle(g, inst)
    O.vec("l")[:] = g["lam"]
    lag = O.alm_prepare(float(g["rho0"]))
    assert abs(lag - float(g["lag_sq"])) <= TOL * float(g["lag_sq"])
    assert rel_err(O.vec("s"), g["constr_sum"]) < TOL
    for c in range(O.n_cones):
        assert rel_err(O.factor("G", c), g[f"grad{c}"]) < TOL
    it = O.update_sdp_var_one("U", "V", 1.0, 1e-8, 800, 0)
    assert it == int(g["cg_iters"])
    assert rel_err(O.factor("U", 0), g["U_after_cg0"]) < 1e-9     # a converged CG solve: conditioning-limited


def test_alm_inner_iterations(golden):
    name, g, inst = golden
    O = make_oracle(g, inst)
    rho = float(g["rho0"])
    O.alm_prepare(rho)
    for k in range(len(g["it_tau"])):
        root, o = O.alm_inner_iter(rho, k)
        assert root == int(g["it_root"][k])
        # rounding differences are amplified by the L-BFGS recursion: tolerance grows with the iteration index
        tol = 1e-11 * 10 ** k
        assert abs(o["tau"] - g["it_tau"][k]) <= tol * max(1.0, abs(g["it_tau"][k]))
        assert abs(o["lag_norm_sq"] - g["it_lag"][k]) <= tol * abs(g["it_lag"][k])
        assert abs(o["pinf"] - g["it_pinf"][k]) <= tol * abs(g["it_pinf"][k])


# ---- host side of the CUDA library -----------------------------------------------------------------

def test_host_presolve_matches_reference(golden):
    name, g, inst = golden
    for c, cone in enumerate(inst.cones):
        info, rows, cols = capi.host_presolve(cone, inst.m)
        if info["rank_one_objective"]:
            # C = c ee^T (+ remainder) keeps the cone on the sparse scratch although the reference goes dense
            assert int(g[f"dense{c}"]) == 1 and info["dense_path"] == 0 and info["nnzC"] == 0
            assert info["psize"] == cone.n + (inst.m - 1)      # theta: diagonal + one entry per edge
        else:
            assert info["dense_path"] == int(g[f"dense{c}"])
        if info["rank_one_objective"]:
            pass
        elif not info["dense_path"]:
            assert info["psize"] == int(g[f"psize{c}"])
            assert np.array_equal(rows, g[f"prow{c}"]) and np.array_equal(cols, g[f"pcol{c}"])
        else:
            assert info["psize"] == cone.n * (cone.n + 1) // 2
        r, cap = capi.host_rank_rule(cone.n, info["n_nonzero_coeff"], len(inst.cones))
        assert r == int(g["ranks"][c]) and cap == int(g["rank_max"][c])


def test_host_line_search_matches_reference(golden):
    """Recomputes the five line-search sums with the oracle and feeds them to the library's host line search."""
    name, g, inst = golden
    O = make_oracle(g, inst)
    rho = float(g["rho0"])
    O.alm_prepare(rho)
    for k in range(3):
        s_before = O.vec("s").copy()
        lam = O.vec("l").copy()
        root, o = O.alm_inner_iter(rho, k)
        q1, q2 = O.vec("q"), O.vec("Q")
        q0 = inst.b - s_before + lam / rho
        sums = [q2 @ q2, q1 @ q2, q0 @ q2, q1 @ q1, q0 @ q1]
        root2, tau2 = capi.host_line_search(rho, sums, o["p1"], o["p2"])
        assert root2 == root
        assert abs(tau2 - o["tau"]) <= 1e-9 * max(1.0, abs(o["tau"]))


def test_golden_solves_reached_tolerance():
    """Sanity of the fixtures themselves: the reference solved every case to its DIMACS tolerance."""
    import os
    from conftest import GOLDEN_CASES, GOLDEN_DIR
    for name in GOLDEN_CASES:
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        sol = json.loads(str(g["solve"]))
        assert sol["status"] in (1.0, 2.0)
        assert sol["pinf"] <= 1e-5 and sol["gap"] <= 5e-5


def test_host_presolve_is_order_independent():
    """The union pattern must not depend on the order of the constraints in the file: the pre-solve merges sorted
    runs when the keys ascend and falls back to a sort / binary search when they do not."""
    from lorads_b200 import sdpa
    inst = sdpa.lovasz_theta(120, 900, 3)
    cone = inst.cones[0]
    rng = np.random.default_rng(1)
    perm = rng.permutation(inst.m)
    counts = np.diff(cone.beg)
    cols = [np.arange(cone.beg[c], cone.beg[c + 1]) for c in range(inst.m + 1)]
    order = [cols[0]] + [cols[1 + p] for p in perm]
    take = np.concatenate(order)
    beg = np.concatenate([[0], np.cumsum([counts[0]] + [counts[1 + p] for p in perm])]).astype(np.int64)
    shuffled = sdpa.Cone(n=cone.n, beg=beg, idx=cone.idx[take], elem=cone.elem[take])
    a, ra, ca = capi.host_presolve(cone, inst.m)
    b, rb, cb = capi.host_presolve(shuffled, inst.m)
    assert a == b
    assert np.array_equal(ra, rb) and np.array_equal(ca, cb)
    # a sparse objective too (no rank-one shortcut): MaxCut with shuffled constraints
    mc = sdpa.maxcut(300, 1500, 5).cones[0]
    counts = np.diff(mc.beg)
    perm = rng.permutation(300)
    take = np.concatenate([np.arange(mc.beg[0], mc.beg[1])] + [np.arange(mc.beg[1 + p], mc.beg[2 + p]) for p in perm])
    beg = np.concatenate([[0], np.cumsum([counts[0]] + [counts[1 + p] for p in perm])]).astype(np.int64)
    sh = sdpa.Cone(n=mc.n, beg=beg, idx=mc.idx[take], elem=mc.elem[take])
    a, ra, ca = capi.host_presolve(mc, 300)
    b, rb, cb = capi.host_presolve(sh, 300)
    assert a == b and np.array_equal(ra, rb) and np.array_equal(ca, cb)
