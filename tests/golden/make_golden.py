"""Generates tests/golden/*.npz from the UNTOUCHED reference compiled in oracle/_ref (run where /root/reference
exists: `make -C oracle ref && python tests/golden/make_golden.py`).

Each fixture holds, for one small seeded instance, the reference's own outputs of the hot-path functions on its
srand(925) start (factors included, so no libc rand() reproduction is needed to use the fixture) plus the
result of a whole solve.  The instance itself is rebuilt from its generator arguments, which are stored too.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lorads_b200 import sdpa  # noqa: E402
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "maxcut_n120": ("maxcut", dict(n=120, n_edges=600, seed=1)),
    "maxcut_n800": ("maxcut", dict(n=800, n_edges=19176, seed=1)),
    "mcomp_60x50": ("matrix_completion", dict(n1=60, n2=50, n_samples=700, rank=3, seed=7)),
    "theta_n60": ("lovasz_theta", dict(n=60, n_edges=300, seed=5)),
    "twoblock": ("two_block", dict(n1=30, e1=90, n2=120, e2=500)),
    # dense scratch in the reference, rank-one objective + sparse scratch in the CUDA library
    "theta_n200": ("lovasz_theta", dict(n=200, n_edges=1000, seed=5)),
}


# mixed SDP + LP instances: an LP block appended to a base instance (sdpa.add_lp_block)
LP_CASES = {
    "lp_maxcut_n60": ("with_lp", dict(base="maxcut", base_kw=dict(n=60, n_edges=200, seed=1), n_lp=25, seed=7)),
    "lp_maxcut_n300": ("with_lp", dict(base="maxcut", base_kw=dict(n=300, n_edges=1500, seed=2), n_lp=400, seed=3)),
    "lp_theta_n40": ("with_lp", dict(base="lovasz_theta", base_kw=dict(n=40, n_edges=150, seed=4), n_lp=30, seed=5)),
    "lp_twoblock": ("with_lp", dict(base="two_block", base_kw=dict(n1=30, e1=90, n2=120, e2=500), n_lp=60, seed=9)),
}


def build_instance(kind, kw):
    if kind == "with_lp":
        return sdpa.add_lp_block(build_instance(kw["base"], kw["base_kw"]), kw["n_lp"], kw["seed"])
    if kind == "two_block":
        a = sdpa.maxcut(kw["n1"], kw["e1"], 11)
        b = sdpa.maxcut(kw["n2"], kw["e2"], 12)
        # two cones sharing the constraint index space; block 2 only touches the first n2 constraints
        m = max(a.m, b.m)
        def widen(cone, m_old):
            beg = np.concatenate([cone.beg, np.full(m - m_old, cone.beg[-1], dtype=np.int64)])
            return sdpa.Cone(n=cone.n, beg=beg, idx=cone.idx, elem=cone.elem)
        cones = [widen(a.cones[0], a.m), widen(b.cones[0], b.m)]
        bb = np.zeros(m)
        bb[: a.m] += a.b
        bb[: b.m] += b.b
        return sdpa.Instance(m=m, b=bb, cones=cones, name="twoblock")
    return getattr(sdpa, kind)(**kw)


def main():
    only = sys.argv[1:]
    for name, (kind, kw) in CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        inst = build_instance(kind, kw)
        d = tempfile.mkdtemp()
        path = os.path.join(d, name + ".dat-s")
        sdpa.write_dat_s(inst, path)
        R = ref.RefSolver(path, 32)
        out = {"meta": json.dumps({"kind": kind, "kw": kw})}
        nc = R.n_cones
        rho = R.dinfo(6)
        out["rho0"] = rho
        out["norms"] = np.array([R.dinfo(k) for k in range(6)])
        out["ranks"] = np.array([R.rank(c) for c in range(nc)])
        out["rank_max"] = np.array([R.info(10, c) for c in range(nc)])
        w = rng.standard_normal(R.m)
        lam = 0.1 * rng.standard_normal(R.m)
        out["w"] = w
        out["lam"] = lam
        for c in range(nc):
            for f in "RUV":
                out[f"{f}{c}"] = R.factor(f, c).copy()
            out[f"dense{c}"] = R.info(6, c)
            out[f"psize{c}"] = R.info(4, c)
            if not R.info(6, c):
                pr, pc = R.pattern(c)
                out[f"prow{c}"], out[f"pcol{c}"] = pr, pc
            out[f"auv_RR{c}"] = R.auv("R", "R", c)
            out[f"auv_UV{c}"] = R.auv("U", "V", c)
            out[f"obj_RR{c}"] = R.obj_auv("R", "R", c)
            out[f"obj_UV{c}"] = R.obj_auv("U", "V", c)
            out[f"wsum_C{c}"] = R.wsum_mulrk(w, True, "V", c).copy()
            out[f"wsum_noC{c}"] = R.wsum_mulrk(w, False, "V", c).copy()
            x = rng.standard_normal(R.factor("U", c).shape)
            out[f"cgx{c}"] = x
            out[f"cgmv{c}"] = R.cg_matvec(x, "V", c).copy()
        # gradient with a non-trivial dual vector
        R.vec("l")[:] = lam
        out["lag_sq"] = R.alm_prepare(rho)
        for c in range(nc):
            out[f"grad{c}"] = R.factor("G", c).copy()
        out["constr_sum"] = R.vec("s").copy()
        # one ADMM block solve from that state (cone 0, update U with V fixed)
        out["cg_iters"] = R.update_sdp_var_one("U", "V", 1.0, 1e-8, 800, 0)
        out["U_after_cg0"] = R.factor("U", 0).copy()
        # ALM inner iterations from lambda = 0
        R.vec("l")[:] = 0.0
        R.alm_prepare(rho)
        taus, lags, pinfs, roots, p1s = [], [], [], [], []
        for k in range(8):
            root, o = R.alm_inner_iter(rho, k)
            roots.append(root); taus.append(o["tau"]); lags.append(o["lag_norm_sq"]); pinfs.append(o["pinf"]); p1s.append(o["p1"])
        out["it_tau"], out["it_lag"], out["it_pinf"], out["it_root"], out["it_p1"] = map(np.array, (taus, lags, pinfs, roots, p1s))
        for c in range(nc):
            out[f"R_after_iters{c}"] = R.factor("R", c).copy()
        # whole solve with a fresh context
        R2 = ref.RefSolver(path, 32)
        sol = R2.solve()
        out["solve"] = json.dumps(sol)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ok", {k: sol[k] for k in ("pobj", "dobj", "pinf", "gap", "dinf", "alm_inner", "admm_iter", "cg_iter", "status")})


def main_lp():
    """LP fixtures: the reference's own LP function set (ALMCalGradLP, ALMCalq12p12LP, LORADSUpdateSDPLPVar, ...)
    driven through oracle/ref_harness.c."""
    only = sys.argv[1:]
    for name, (kind, kw) in LP_CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(sum(map(ord, name)))
        inst = build_instance(kind, kw)
        d = tempfile.mkdtemp()
        path = os.path.join(d, name + ".dat-s")
        sdpa.write_dat_s(inst, path)
        R = ref.RefSolver(path, 32)
        assert R.n_lp == inst.lp.n
        out = {"meta": json.dumps({"kind": kind, "kw": kw})}
        nc = R.n_cones
        rho = R.dinfo(6)
        out["rho0"] = rho
        out["norms"] = np.array([R.dinfo(k) for k in range(6)])
        out["ranks"] = np.array([R.rank(c) for c in range(nc)])
        for c in range(nc):
            for f in "RUV":
                out[f"{f}{c}"] = R.factor(f, c).copy()
        for f in "RUV":
            out[f"lp{f}"] = R.lp(f).copy()
        lam = 0.1 * rng.standard_normal(R.m)
        out["lam"] = lam
        R.vec("l")[:] = lam
        out["lag_sq"] = R.alm_prepare(rho)
        out["lpG"] = R.lp("G").copy()
        for c in range(nc):
            out[f"grad{c}"] = R.factor("G", c).copy()
        out["constr_sum"] = R.vec("s").copy()
        # one full ADMM sweep (cones by CG, then the LP columns) from U = V = R, uLp = vLp = rLp
        for c in range(nc):
            R.factor("U", c)[:] = R.factor("R", c)
            R.factor("V", c)[:] = R.factor("R", c)
        R.lp("U")[:] = R.lp("R")
        R.lp("V")[:] = R.lp("R")
        R.admm_init_constr()
        out["sweep_rho"] = 10.0 * rho
        R.admm_update_var(10.0 * rho, 1e-10, 800)
        out["sweep_lpU"], out["sweep_lpV"] = R.lp("U").copy(), R.lp("V").copy()
        out["sweep_s"] = R.vec("s").copy()
        for c in range(nc):
            out[f"sweep_U{c}"], out[f"sweep_V{c}"] = R.factor("U", c).copy(), R.factor("V", c).copy()
        # ALM inner iterations from lambda = 0 on a fresh context
        R = ref.RefSolver(path, 32)
        R.alm_prepare(rho)
        taus, lags, pinfs, roots, p1s, p2s = [], [], [], [], [], []
        for k in range(8):
            root, o = R.alm_inner_iter(rho, k)
            roots.append(root); taus.append(o["tau"]); lags.append(o["lag_norm_sq"]); pinfs.append(o["pinf"])
            p1s.append(o["p1"]); p2s.append(o["p2"])
        out["it_tau"], out["it_lag"], out["it_pinf"], out["it_root"], out["it_p1"], out["it_p2"] = map(
            np.array, (taus, lags, pinfs, roots, p1s, p2s))
        out["lpR_after_iters"] = R.lp("R").copy()
        R2 = ref.RefSolver(path, 32)
        sol = R2.solve()
        out["solve"] = json.dumps(sol)
        out["lp_x_solution"] = R2.lp("R").copy() ** 2
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ok", {k: sol[k] for k in ("pobj", "dobj", "pinf", "gap", "dinf", "alm_inner", "admm_iter", "cg_iter", "status")})


if __name__ == "__main__":
    if not any(a.startswith("lp_") for a in sys.argv[1:]):
        main()
    if not sys.argv[1:] or any(a.startswith("lp_") for a in sys.argv[1:]):
        main_lp()
