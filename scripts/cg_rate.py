"""CG iterations per second of the ADMM block solve (LORADSUpdateSDPVarOne) on BASELINE configs[1]; tolerance 0 forces the
iteration count."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from lorads_b200.capi import Solver  # noqa: E402

S = Solver(bench.make_instance(), device=0)
rho = S.dinfo(6)
S.admm_init_constr()
S.update_sdp_var_one("U", "V", rho, 0.0, 5)
for n in (60, 200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its = S.update_sdp_var_one("U", "V", rho, 0.0, n)
    torch.cuda.synchronize()
    print(f"{its} CG iterations: {its / (time.perf_counter() - t0):.0f} it/s")
