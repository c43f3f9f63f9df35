"""LP cone: the CUDA path through the C ABI against golden vectors produced by the reference's
own LP function set (ALMCalGradLP, ALMCalq12p12LP, LORADSUpdateSDPLPVar, ... driven by oracle/ref_harness.c;
tests/golden/lp_*.npz), and against the compiled reference itself when oracle/_ref travelled to the box.
Kernel outputs: 1e-12 relative (north_star); whole solves: objectives to 1e-6 relative."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import LP_GOLDEN_CASES, ROOT, have_gpu, load_golden, rel_err
from lorads_b200 import sdpa

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]

KTOL = 1e-12


@pytest.fixture(scope="module", params=LP_GOLDEN_CASES)
def lp_golden(request):
    return (request.param,) + load_golden(request.param)


def gpu_solver(inst, **kw):
    from lorads_b200.capi import Solver
    return Solver(inst, **kw)


def test_lp_start_and_norms(lp_golden):
    name, g, inst = lp_golden
    G = gpu_solver(inst)
    assert G.info(19) == inst.lp.n
    # srand(925) draw order with an LP block: R of every cone, rLp, uLp, vLp, then U, V of every cone
    for f in "RUV":
        assert np.array_equal(G.get_lp(f), g[f"lp{f}"])
        for c in range(G.n_cones):
            assert np.array_equal(G.get_factor(f, c), g[f"{f}{c}"])
    # objective norms include the LP block (with the reference's |c|_1^2 in the 2-norm and its idamax_ indexing)
    assert np.allclose([G.dinfo(k) for k in range(6)], g["norms"], rtol=1e-14, atol=0)


def test_lp_gradient(lp_golden):
    name, g, inst = lp_golden
    G = gpu_solver(inst)
    G.set_vec("l", g["lam"])
    lag = G.alm_prepare(float(g["rho0"]))
    assert abs(lag - float(g["lag_sq"])) <= KTOL * float(g["lag_sq"])
    assert rel_err(G.get_vec("s"), g["constr_sum"]) < KTOL            # A(RR^T) + A_lp (r.r)
    assert rel_err(G.get_lp("G"), g["lpG"]) < KTOL                    # 2 (c + A_lp^T w) . r
    for c in range(G.n_cones):
        assert rel_err(G.get_factor("G", c), g[f"grad{c}"]) < KTOL


def test_lp_admm_sweep(lp_golden):
    """One Gauss-Seidel sweep over the cones (CG) and the LP columns (closed form, level-scheduled on the device):
    the level schedule must reproduce the reference's strictly sequential column loop."""
    name, g, inst = lp_golden
    G = gpu_solver(inst)
    G.set_vec("l", g["lam"])
    for c in range(G.n_cones):
        G.set_factor("U", g[f"R{c}"], c)
        G.set_factor("V", g[f"R{c}"], c)
    G.set_lp("U", g["lpR"])
    G.set_lp("V", g["lpR"])
    G.admm_init_constr()
    G.admm_update_var(float(g["sweep_rho"]), 1e-10, 800)
    for c in range(G.n_cones):
        assert rel_err(G.get_factor("U", c), g[f"sweep_U{c}"]) < 1e-8      # CG solves, tolerance 1e-10
        assert rel_err(G.get_factor("V", c), g[f"sweep_V{c}"]) < 1e-8
    assert rel_err(G.get_lp("U"), g["sweep_lpU"]) < 1e-8
    assert rel_err(G.get_lp("V"), g["sweep_lpV"]) < 1e-8
    assert rel_err(G.get_vec("s"), g["sweep_s"]) < 1e-8


def test_lp_sweep_alone_is_exact(lp_golden):
    """The LP part of the sweep in isolation (a CG budget of zero iterations freezes the cone factors): column
    updates to 1e-12 against a sequential numpy restatement of LORADSUpdateLPVarOne (lorads_admm.c:595-628) + the
    constrValSum bookkeeping (lorads_alg_common.c:229-247)."""
    name, g, inst = lp_golden
    G = gpu_solver(inst)
    lp = inst.lp
    m, n = inst.m, lp.n
    rng = np.random.default_rng(5)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    lam = 0.3 * rng.standard_normal(m)
    G.set_vec("l", lam)
    G.set_lp("U", u); G.set_lp("V", v)
    G.admm_init_constr()
    s0 = G.get_vec("s")
    # numpy restatement
    cost = np.zeros(n); cost[lp.idx[lp.beg[0]:lp.beg[1]]] = lp.elem[lp.beg[0]:lp.beg[1]]
    cols = [[] for _ in range(n)]
    for i in range(m):
        for k in range(lp.beg[i + 1], lp.beg[i + 2]):
            cols[lp.idx[k]].append((i, lp.elem[k]))
    rho = 3.7
    s = s0.copy(); x = u * v; uu, vv = u.copy(), v.copy()
    for j in range(n):
        for upd, noupd in ((uu, vv), (vv, uu)):
            w = cost[j] + sum(a * (((s[i] - inst.b[i]) - a * x[j]) * rho - lam[i]) for i, a in cols[j])
            M2 = w * noupd[j] - rho * noupd[j]
            nrm2 = np.sqrt(sum(a * a for _, a in cols[j])) ** 2
            upd[j] = (-M2 / rho) / (1 + nrm2 * noupd[j] ** 2)
            xn = uu[j] * vv[j]
            for i, a in cols[j]:
                s[i] = (s[i] - a * x[j]) + a * xn
            x[j] = xn
    G.admm_update_var(rho, 1e-8, 0)       # zero CG iterations: the cone factors stay as they are
    assert rel_err(G.get_lp("U"), uu) < KTOL
    assert rel_err(G.get_lp("V"), vv) < KTOL
    assert rel_err(G.get_lp("x"), x) < KTOL


def test_lp_alm_inner_iterations(lp_golden):
    name, g, inst = lp_golden
    G = gpu_solver(inst)
    rho = float(g["rho0"])
    G.alm_prepare(rho)
    for k in range(len(g["it_tau"])):
        root, o = G.alm_inner_iter(rho, k)
        assert root == int(g["it_root"][k])
        tol = 1e-11 * 10 ** k      # rounding noise is amplified by the L-BFGS recursion iteration after iteration
        assert abs(o["tau"] - g["it_tau"][k]) <= tol * max(1.0, abs(g["it_tau"][k]))
        assert abs(o["lag_norm_sq"] - g["it_lag"][k]) <= tol * abs(g["it_lag"][k])
        assert abs(o["pinf"] - g["it_pinf"][k]) <= tol * abs(g["it_pinf"][k])
        assert abs(o["p1"] - g["it_p1"][k]) <= tol * max(1.0, abs(g["it_p1"][k]))
        assert abs(o["p2"] - g["it_p2"][k]) <= tol * max(1.0, abs(g["it_p2"][k]))
    assert rel_err(G.get_lp("R"), g["lpR_after_iters"]) < 1e-4


def test_lp_whole_solve(lp_golden):
    from lorads_b200.capi import default_params
    name, g, inst = lp_golden
    ref = json.loads(str(g["solve"]))
    G = gpu_solver(inst)
    res = G.solve(default_params())
    assert res["status"] in (1, 2) and ref["status"] in (1.0, 2.0)
    assert res["pInfeasL1"] <= 1e-5 and res["pdGap"] <= 5e-5
    scale = 1 + abs(ref["pobj"])
    if ref["alm_inner"] < 1000:
        # short runs follow the reference iteration for iteration
        assert res["almInnerIter"] == int(ref["alm_inner"]) and res["admmIter"] == int(ref["admm_iter"])
        assert abs(res["cgIter"] - int(ref["cg_iter"])) <= 2
        assert abs(res["pObj"] - ref["pobj"]) <= 1e-6 * scale and abs(res["dObj"] - ref["dobj"]) <= 1e-6 * scale
        assert abs(res["dInfeasL1"] - ref["dinf"]) <= 1e-6
        assert np.allclose(G.get_lp("R") ** 2, g["lp_x_solution"], atol=1e-6)
    else:
        assert abs(res["pObj"] - ref["pobj"]) <= 3e-5 * scale and abs(res["dObj"] - ref["dobj"]) <= 3e-5 * scale
    assert (G.get_lp("R") ** 2 >= 0).all()


BIN = os.path.join(ROOT, "oracle", "_ref", "lorads_gpu32")
REF = os.path.join(ROOT, "oracle", "_ref", "lorads_ref32")


@pytest.mark.skipif(not (os.path.exists(BIN) and os.path.exists(REF)), reason="drop-in / reference binaries not built")
def test_lp_dropin_matches_reference_binary(tmp_path):
    """Same mixed SDP + LP .dat-s file through the reference CPU binary and through the reference's main.c linked
    against the CUDA library: same printed result block."""
    from test_gpu_dropin import parse
    g, inst = load_golden("lp_maxcut_n300")
    path = str(tmp_path / "mix.dat-s")
    sdpa.write_dat_s(inst, path)
    gpu = subprocess.run([BIN, path], capture_output=True, text=True, timeout=300)
    cpu = subprocess.run([REF, path], capture_output=True, text=True, timeout=300)
    assert gpu.returncode == 0 and cpu.returncode == 0, gpu.stderr[-2000:]
    assert "lp Cols = 400" in gpu.stdout
    a, b = parse(gpu.stdout), parse(cpu.stdout)
    for k in ("pobj", "dobj"):
        assert abs(a[k] - b[k]) <= 1e-6 * (1 + abs(b[k]))
    assert a["pinf"] <= 1e-5 and a["gap"] <= 5e-5


def test_lp_file_through_the_standalone_cli(tmp_path):
    """Library reader (LP block included) + solve, no reference tree involved; same result block as the golden solve."""
    from test_gpu_dropin import parse
    cli = os.path.join(ROOT, "lorads_b200", "lorads_b200_cli")
    assert os.path.exists(cli), "build with python -m lorads_b200.build"
    g, inst = load_golden("lp_twoblock")
    path = str(tmp_path / "mix.dat-s")
    sdpa.write_dat_s(inst, path)
    out = subprocess.run([cli, path, "--quiet"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "lp Cols = 60" in out.stdout
    got, ref = parse(out.stdout), json.loads(str(g["solve"]))
    assert abs(got["pobj"] - ref["pobj"]) <= 1e-6 * (1 + abs(ref["pobj"]))
    assert abs(got["dobj"] - ref["dobj"]) <= 1e-6 * (1 + abs(ref["dobj"]))


def test_lp_rejected_with_sharding_and_bad_input():
    from lorads_b200.capi import Lb2Error
    inst = sdpa.add_lp_block(sdpa.maxcut(20, 40, 1), 5, 2)
    bad = sdpa.LpBlock(n=5, beg=inst.lp.beg, idx=inst.lp.idx + 3, elem=inst.lp.elem)      # column index out of range
    with pytest.raises(Lb2Error):
        gpu_solver(sdpa.Instance(m=inst.m, b=inst.b, cones=inst.cones, lp=bad))
