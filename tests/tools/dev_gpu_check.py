"""Development check on a GPU box: CUDA path vs the compiled reference (oracle/_ref) on small instances."""
import os, sys, time, tempfile, functools
print = functools.partial(print, flush=True)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params
from oracle import ref

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

def check(inst, bits=32, solve=True):
    d = tempfile.mkdtemp()
    path = os.path.join(d, inst.name + ".dat-s")
    sdpa.write_dat_s(inst, path)
    R = ref.RefSolver(path, bits)
    G = Solver(inst)
    print(f"== {inst.name}: m={R.m} n={R.dim()} r_ref={R.rank()} r_gpu={G.rank()} P_ref={R.info(4)} P_gpu={G.info(4)} dense={G.info(6)}")
    for w in "RUV":
        G.set_factor(w, R.factor(w))
    print("  factor roundtrip", rel(G.get_factor("R"), R.factor("R")))
    for u, v in (("R", "R"), ("U", "V")):
        a_ref = R.auv(u, v); o_ref = R.obj_auv(u, v)
        a_gpu, o_gpu = G.auv(u, v, with_obj=True)
        print(f"  auv({u},{v}) rel={rel(a_gpu, a_ref):.2e} obj rel={abs(o_gpu - o_ref) / abs(o_ref):.2e}  A-only rel={rel(G.auv(u, v), a_ref):.2e}")
    rng = np.random.default_rng(0)
    w = rng.standard_normal(R.m)
    for addc in (True, False):
        print(f"  wsum_mulrk addC={addc} rel={rel(G.wsum_mulrk(w, addc, 'V'), R.wsum_mulrk(w, addc, 'V')):.2e}")
    lam = rng.standard_normal(R.m) * 0.1
    R.vec("l")[:] = lam; G.set_vec("l", lam)
    rho = R.dinfo(6)
    l_ref = R.alm_prepare(rho); l_gpu = G.alm_prepare(rho)
    print(f"  alm_prepare lagSq rel={abs(l_gpu - l_ref) / l_ref:.2e}  grad rel={rel(G.get_factor('G'), R.factor('G')):.2e}")
    x = rng.standard_normal(R.factor("U").shape)
    print(f"  cg_matvec rel={rel(G.cg_matvec(x, 'V'), R.cg_matvec(x, 'V')):.2e}")
    # ADMM block solve from the same state
    R.vec("l")[:] = lam; G.set_vec("l", lam)
    R.alm_prepare(rho); G.alm_prepare(rho)   # constrVal/constrValSum = A(RR^T) in both
    it_ref = R.update_sdp_var_one("U", "V", 1.0, 1e-8, 800); it_gpu = G.update_sdp_var_one("U", "V", 1.0, 1e-8, 800)
    print(f"  update_sdp_var_one iters ref={it_ref} gpu={it_gpu} U rel={rel(G.get_factor('U'), R.factor('U')):.2e}")
    # ALM inner iterations
    R.vec("l")[:] = 0; G.set_vec("l", np.zeros(R.m))
    R.alm_prepare(rho); G.alm_prepare(rho)
    for k in range(6):
        rr, orf = R.alm_inner_iter(rho, k); rg, og = G.alm_inner_iter(rho, k)
        print(f"  inner {k}: tau {orf['tau']:.12e} / {og['tau']:.12e}  lag {orf['lag_norm_sq']:.10e} / {og['lag_norm_sq']:.10e} pinf {orf['pinf']:.8e}/{og['pinf']:.8e} R rel={rel(G.get_factor('R'), R.factor('R')):.2e}")
    if solve:
        R2 = ref.RefSolver(path, bits)
        t = time.time(); sr = R2.solve(); tr = time.time() - t
        G2 = Solver(inst)
        t = time.time(); sg = G2.solve(default_params(verbose=0)); tg = time.time() - t
        print("  ref solve:", {k: (f"{v:.8e}" if isinstance(v, float) else v) for k, v in sr.items()}, f"{tr:.2f}s")
        print("  gpu solve:", {k: (f"{v:.8e}" if isinstance(v, float) else v) for k, v in sg.items()}, f"{tg:.2f}s")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "mc800"
    solve = "nosolve" not in sys.argv
    if which == "mc800": check(sdpa.maxcut(800, 19176, 1), solve=solve)
    if which == "mcomp": check(sdpa.matrix_completion(150, 150, 4000, 3, 7), solve=solve)
    if which == "theta": check(sdpa.lovasz_theta(120, 900, 5), solve=solve)
    if which == "mc20k": check(sdpa.maxcut(20000, 100000, 2), solve=False)
