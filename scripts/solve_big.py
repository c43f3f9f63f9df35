"""Whole solve of a BASELINE.json configuration on the GPU (development driver)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
limit = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
t = time.time()
if cfg == "cfg1": inst = sdpa.maxcut(800, 19176, 1)
elif cfg == "cfg2": inst = sdpa.maxcut(100_000, 500_000, 3)
elif cfg == "cfg3": inst = sdpa.lovasz_theta(5000, 50_000, 5)
elif cfg == "cfg3s": inst = sdpa.lovasz_theta(1000, 8000, 5)
elif cfg == "cfg4": inst = sdpa.matrix_completion(20_000, 20_000, 2_000_000, 3, 7)
elif cfg == "cfg4s": inst = sdpa.matrix_completion(2000, 2000, 100_000, 3, 7)
elif cfg == "cfg5": inst = sdpa.maxcut(1_000_000, 5_000_000, 5)
print(f"instance {inst.name} generated in {time.time()-t:.1f}s", flush=True)
t = time.time()
S = Solver(inst)
print(f"setup (presolve + upload) {time.time()-t:.1f}s rank={S.rank()} P={S.info(4)} dense={S.info(6)}", flush=True)
t = time.time()
res = S.solve(default_params(verbose=1, timeSecLimit=limit))
print(json.dumps({k: (float(f"{v:.10g}") if isinstance(v, float) else v) for k, v in res.items()}), f"wall {time.time()-t:.2f}s", flush=True)
