"""The numpy LP restatement (oracle/lp_restate.py) together with the C restatement of the cone operators
(oracle/lorads_oracle.c) against the golden vectors the compiled reference produced with its LP function set
(tests/golden/lp_*.npz).  CPU only."""
import numpy as np
import pytest

from conftest import LP_GOLDEN_CASES, load_golden, rel_err
from lorads_b200 import sdpa
from oracle import restate
from oracle.lp_restate import LpOracle

TOL = 1e-12


@pytest.fixture(scope="module", params=LP_GOLDEN_CASES)
def lp_golden(request):
    return (request.param,) + load_golden(request.param)


def cones_only(inst):
    return sdpa.Instance(m=inst.m, b=inst.b, cones=inst.cones, name=inst.name)


def test_lp_norms_and_start(lp_golden):
    name, g, inst = lp_golden
    O = restate.OracleSolver(cones_only(inst))
    L = LpOracle(inst.lp, inst.m)
    l1, l2sq, linf = L.obj_norm_terms()
    # the cone-only restatement knows nothing of the LP block: add the LP terms the way the reference does
    n1 = O.dinfo(0) + l1
    n2 = np.sqrt(O.dinfo(1) ** 2 + l2sq)
    ninf = max(O.dinfo(2), linf)
    assert np.allclose([n1, n2, ninf], g["norms"][:3], rtol=1e-14, atol=0)
    assert np.allclose([O.dinfo(k) for k in (3, 4, 5)], g["norms"][3:], rtol=1e-14, atol=0)
    # cone factors come first in the srand(925) stream, so the cone-only start still matches for R
    for c in range(O.n_cones):
        assert np.array_equal(O.factor("R", c), g[f"R{c}"])


def test_lp_gradient_and_constraint_sum(lp_golden):
    name, g, inst = lp_golden
    O = restate.OracleSolver(cones_only(inst))
    L = LpOracle(inst.lp, inst.m)
    r = g["lpR"]
    s = sum(O.auv("R", "R", c) for c in range(O.n_cones)) + L.constr(r, r)
    assert rel_err(s, g["constr_sum"]) < TOL
    rho = float(g["rho0"])
    w = -g["lam"] - rho * inst.b + rho * s
    glp = L.grad(w, r)
    assert rel_err(glp, g["lpG"]) < TOL
    gsq = float(glp @ glp)
    for c in range(O.n_cones):
        G = O.wsum_mulrk(w, True, "R", c) * 2.0
        assert rel_err(G, g[f"grad{c}"]) < TOL
        gsq += float((G * G).sum())
    assert abs(gsq - float(g["lag_sq"])) <= 1e-11 * float(g["lag_sq"])


def test_lp_sweep(lp_golden):
    """LP half of the ADMM sweep: start from the reference's cone factors after their CG solves, rebuild
    constrValSum, run the sequential column loop, compare with the reference's uLp / vLp / constrValSum."""
    name, g, inst = lp_golden
    O = restate.OracleSolver(cones_only(inst))
    L = LpOracle(inst.lp, inst.m)
    for c in range(O.n_cones):
        O.factor("U", c)[:] = g[f"sweep_U{c}"]
        O.factor("V", c)[:] = g[f"sweep_V{c}"]
    r = g["lpR"]
    x = r * r
    cvs = sum(O.auv("U", "V", c) for c in range(O.n_cones)) + L.constr(r, r)
    u, v = r.copy(), r.copy()
    L.sweep(float(g["sweep_rho"]), inst.b, g["lam"], cvs, x, u, v)
    assert rel_err(u, g["sweep_lpU"]) < 1e-10
    assert rel_err(v, g["sweep_lpV"]) < 1e-10
    assert rel_err(cvs, g["sweep_s"]) < 1e-10
