"""Size-independent properties of the CUDA operators at BASELINE.json's full single-GPU size (MaxCut n = 1e5),
where the CPU oracle is too slow to be the checker, plus determinism."""
import numpy as np
import pytest

from conftest import have_gpu, rel_err
from lorads_b200 import sdpa

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def big():
    from lorads_b200.capi import Solver
    inst = sdpa.maxcut(100_000, 500_000, 3)
    return inst, Solver(inst)


def test_maxcut_constraint_values_are_row_norms(big):
    inst, G = big
    R = G.get_factor("R")
    assert rel_err(G.auv("R", "R"), (R * R).sum(1)) < 1e-13
    U, V = G.get_factor("U"), G.get_factor("V")
    assert rel_err(G.auv("U", "V"), (U * V).sum(1)) < 1e-13


def test_symmetry_and_linearity(big):
    inst, G = big
    a1, o1 = G.auv("U", "V", with_obj=True)
    a2, o2 = G.auv("V", "U", with_obj=True)
    assert rel_err(a1, a2) < 1e-14 and abs(o1 - o2) <= 1e-13 * abs(o1)
    U = G.get_factor("U")
    G.set_factor("U", 3.0 * U)
    a3, o3 = G.auv("U", "V", with_obj=True)
    G.set_factor("U", U)
    assert rel_err(a3, 3.0 * a1) < 1e-14 and abs(o3 - 3.0 * o1) <= 1e-13 * abs(o1)


def test_adjoint_identity(big):
    """<A(sym(U V^T)), w> + <C, sym(U V^T)> == <(C + sum_i w_i A_i) V, U>: ties the A(UV^T) kernel to the
    weighted-sum + SpMM kernels without any CPU reference."""
    inst, G = big
    w = np.random.default_rng(0).standard_normal(inst.m)
    a, o = G.auv("U", "V", with_obj=True)
    Y = G.wsum_mulrk(w, True, "V")
    lhs = float(a @ w + o)
    rhs = float((Y * G.get_factor("U")).sum())
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), 1.0)


def test_objective_matches_explicit_sum(big):
    inst, G = big
    cone = inst.cones[0]
    r, c = sdpa.unpack_idx(cone.n, cone.idx[cone.beg[0]:cone.beg[1]])
    v = cone.elem[cone.beg[0]:cone.beg[1]]
    R = G.get_factor("R")
    z = np.einsum("ij,ij->i", R[r], R[c])
    expect = float(np.sum(np.where(r == c, v * z, 2 * v * z)))
    assert abs(G.obj_auv("R", "R") - expect) <= 1e-12 * abs(expect)


def test_cg_operator_is_symmetric_positive(big):
    inst, G = big
    rng = np.random.default_rng(1)
    shape = G.get_factor("U").shape
    x, y = rng.standard_normal(shape), rng.standard_normal(shape)
    Mx, My = G.cg_matvec(x, "V"), G.cg_matvec(y, "V")
    assert abs((Mx * y).sum() - (My * x).sum()) <= 1e-11 * abs((Mx * y).sum())
    assert (Mx * x).sum() >= (x * x).sum() * (1 - 1e-12)          # M = I + A_V^T A_V


def test_bitwise_determinism(big):
    inst, G = big
    rho = G.dinfo(6)
    R0 = G.get_factor("R")
    outs = []
    for _ in range(2):
        G.set_factor("R", R0)
        G.set_vec("l", np.zeros(inst.m))
        G.alm_prepare(rho)
        G.time_alm_inner_iters(rho, 5)
        outs.append(G.get_factor("R").copy())
    assert np.array_equal(outs[0], outs[1])


def test_inner_iterations_decrease_the_merit(big):
    inst, G = big
    rho = G.dinfo(6)
    G.set_vec("l", np.zeros(inst.m))
    lag0 = G.alm_prepare(rho)
    lags = []
    for k in range(12):
        root, o = G.alm_inner_iter(rho, k)
        assert root != 0 and 0.0 <= o["tau"] <= 1.0
        lags.append(o["lag_norm_sq"])
    assert lags[-1] < lag0


@pytest.mark.parametrize("case", ["theta", "mcomp", "maxcut"])
def test_reopt_path_reaches_the_same_optimum(case):
    """A solve pushed into the re-optimisation loop (reopt -> objScale_dualvar, lorads_solver.c:1040-1052, main.c:386-435)
    by a tiny ADMM iteration limit has to land on the optimum of the default solve: the objective rescaling touches every
    copy of C the kernels read (pattern values, item weights, the vertex-centric adjacency, the rank-one coefficient)."""
    from lorads_b200.capi import Solver, default_params
    inst = {"theta": lambda: sdpa.lovasz_theta(60, 300, 5), "mcomp": lambda: sdpa.matrix_completion(60, 50, 700, 3, 7),
            "maxcut": lambda: sdpa.maxcut(120, 600, 1)}[case]()
    ref = Solver(inst).solve(default_params())
    p = default_params()
    p.maxALMIter = 6          # leave phase 1 early ...
    p.maxADMMIter = 20        # ... and phase 2 before it converges, so that reopt has to finish the job
    res = Solver(inst).solve(p)
    assert ref["status"] in (1, 2)
    if case == "theta":
        # the hard instance may run out of its (tiny) iteration budget; an inconsistent rescaling would be off by the
        # reopt factor 5, not by a percent
        assert res["status"] in (1, 2, 3)
        assert abs(res["pObj"] - ref["pObj"]) <= 1e-2 * (1 + abs(ref["pObj"]))
    else:
        assert res["status"] in (1, 2)
        assert res["pInfeasL1"] <= 1e-5 and res["pdGap"] <= 5e-5
        assert abs(res["pObj"] - ref["pObj"]) <= 1e-4 * (1 + abs(ref["pObj"]))
