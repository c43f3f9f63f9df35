"""Factors wider than 32 columns (two to four passes per row): times the two gather kernels of MaxCut n = 1e5 at ranks
64 / 96 / 128 with the register budget of the one-pass kernels (LORADS_B200_VC_WIDE_MINB=0: 4-5 resident CTAs per SM, 48-64
registers, accumulators spilled) against 3 and 2 resident CTAs per SM (80 / 128 registers, no spills), and checks that
the three builds of each kernel return bit-identical results.
usage: python scripts/wide_minb.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lorads_b200.capi import Solver  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
inst = bench.make_instance()
rng = np.random.default_rng(11)
w = rng.standard_normal(inst.m)
for rank, tlr in ((64, 5.55), (96, 8.33), (128, 11.11)):
    S = Solver(inst, device=0, times_log_rank=tlr)
    rho = S.dinfo(6)
    S.alm_prepare(rho)
    ref = None
    for minb in (0, 3, 2):
        os.environ["LORADS_B200_VC_WIDE_MINB"] = str(minb)      # 0 = budget of the one-pass kernels
        a, o = S.auv("U", "V", with_obj=True)
        y = S.wsum_mulrk(w, True, "U")
        if ref is None:
            ref = (a.copy(), o, y.copy())
        # the objective is a sum over CTAs and the persistent grid follows the CTA count per SM: compared to 1e-14 only
        same = np.array_equal(a, ref[0]) and np.array_equal(y, ref[2])
        obj_rel = abs(o - ref[1]) / abs(ref[1])
        tri_hot, spmm_hot = S.bench_kernel(0, reps) * 1e3, S.bench_kernel(3, reps) * 1e3
        tri_cold, spmm_cold = S.bench_kernel(0, 8, True) * 1e3, S.bench_kernel(3, 8, True) * 1e3
        print(f"rank {S.rank(0):4d} ld {S.info(17):4d} minb {minb}: TRI hot {tri_hot:7.1f} cold {tri_cold:7.1f} us   "
              f"SpMM hot {spmm_hot:7.1f} cold {spmm_cold:7.1f} us   A() and SpMM bit-identical to minb 0: {same}, "
              f"objective rel. diff {obj_rel:.1e}", flush=True)
    S.close()
