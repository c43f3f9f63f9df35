"""Prints the relative deviation of the first inner iterations from the reference's golden vectors, per iteration
(used to set the tolerances of tests/test_gpu_parity.py from measurements instead of a guessed growth law)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN_CASES, load_golden  # noqa: E402
from lorads_b200.capi import Solver  # noqa: E402

for name in GOLDEN_CASES:
    g, inst = load_golden(name)
    G = Solver(inst)
    for c in range(G.n_cones):
        for f in "RUV":
            G.set_factor(f, g[f"{f}{c}"], c)
    rho = float(g["rho0"])
    G.alm_prepare(rho)
    rows = []
    for k in range(len(g["it_tau"])):
        root, o = G.alm_inner_iter(rho, k)
        e = lambda a, b: abs(a - b) / max(1.0, abs(b))
        rows.append((k, int(root == int(g["it_root"][k])), e(o["tau"], g["it_tau"][k]), abs(o["lag_norm_sq"] - g["it_lag"][k]) / abs(g["it_lag"][k]),
                     abs(o["pinf"] - g["it_pinf"][k]) / abs(g["it_pinf"][k]), e(o["p1"], g["it_p1"][k])))
    print(name)
    for r in rows:
        print("   k=%d root_ok=%d tau %.1e lag %.1e pinf %.1e p1 %.1e" % r)
    G.close()
