"""Times the hot kernels of the ALM step on a BASELINE workload: back-to-back (L2 warm) and with the L2 flushed before
every launch, with the algorithmic bytes of bench.kernel_bytes and the fraction of the measured HBM peak.
usage: python scripts/kprof.py [cfg2|cfg3|cfg4|cfg5] [reps] [times_log_rank]   (also the ncu target: few launches with
reps = 1; a small times_log_rank gives the narrow factors a column-sharded rank works on)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lorads_b200 import sdpa  # noqa: E402
from lorads_b200.capi import Solver  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if wl in bench.WORKLOADS:
    bench.WORKLOAD = bench.WORKLOADS[wl]
    inst = bench.make_instance()
elif wl == "cfg3":      # BASELINE configs[2]: Lovasz theta n = 5000, 5e4 edges (rank-one objective layout)
    inst = sdpa.lovasz_theta(5000, 50_000, 5)
elif wl == "cfg4":      # BASELINE configs[3]: matrix completion 20k x 20k, 2M samples
    inst = sdpa.matrix_completion(20_000, 20_000, 2_000_000, 3, 7)
else:
    raise SystemExit("unknown workload")
tlr = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
S = Solver(inst, device=0, times_log_rank=tlr)
rho = S.dinfo(6)
S.alm_prepare(rho)
S.time_alm_inner_iters(rho, 3)
peak, _ = bench.measured_peak_gbs()
kb = bench.kernel_bytes(S)
out = {"workload": wl, "n": S.dim(0), "m": S.m, "rank": S.rank(0), "ld": S.info(17), "vc": S.info(21), "kernels": {}}
ovh = S.bench_kernel(99, 20, True) * 1e3
print(f"cold timing overhead (one-thread kernel timed the same way): {ovh:.1f} us, subtracted from the cold figures")
for w in (0, 3, 1, 4):
    name, nbytes = kb[w]
    if nbytes == 0:
        continue
    hot = S.bench_kernel(w, reps) * 1e3
    cold = max(S.bench_kernel(w, max(3, reps // 4), True) * 1e3 - ovh, 1e-3)
    out["kernels"][name] = {"hot_us": hot, "cold_us": cold, "alg_bytes": nbytes, "frac_hot": nbytes / hot / 1e3 / peak, "frac_cold": nbytes / cold / 1e3 / peak}
    print(f"{name[:58]:58s} hot {hot:8.1f} us ({nbytes / hot / 1e3 / peak:.3f})  cold {cold:8.1f} us ({nbytes / cold / 1e3 / peak:.3f})  alg {nbytes / 1e6:8.1f} MB", flush=True)
sec, done = S.time_alm_inner_iters(rho, 50)
out["alm_inner_iterations_per_second"] = done / sec
print(f"ALM inner iterations/s {done / sec:.1f} ({1e3 * sec / done:.3f} ms/step)")
print("KPROF " + json.dumps(out))
S.close()
