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


def build_instance(kind, kw):
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


if __name__ == "__main__":
    main()
