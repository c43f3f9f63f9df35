"""Launches the two gather kernels of the ALM step a few times on a BASELINE workload (for ncu / quick timing).
usage: python scripts/kprof.py [cfg2|cfg5] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lorads_b200.capi import Solver  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench.WORKLOAD = bench.WORKLOADS[wl]
S = Solver(bench.make_instance(), device=0)
rho = S.dinfo(6)
S.alm_prepare(rho)
S.time_alm_inner_iters(rho, 3)
names = {0: "A(UV^T) TRI", 1: "A(RR^T)", 3: "SpMM G=2SR", 4: "BLAS-1"}
for w in (0, 3, 1, 4):
    hot = S.bench_kernel(w, reps) * 1e3
    cold = S.bench_kernel(w, reps, True) * 1e3
    print(f"{names[w]:14s} hot {hot:8.1f} us   cold {cold:8.1f} us", flush=True)
S.close()
