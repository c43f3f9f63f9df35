"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun), the sharded CUDA path against the ORACLE.

Every rank builds the same instance, joins the library communicator and evaluates the hot-path operators through the
C ABI; rank 0 also evaluates them with the C restatement of the reference (oracle/lorads_oracle.c, CPU) and compares.
Prints one line `SHARDED_RESULT {...json...}` on rank 0."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lorads_b200 import sdpa  # noqa: E402
from lorads_b200.capi import Solver, default_params, load_library  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_library()


def new_comm():
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        assert lib.lb2_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    return (C.create_string_buffer(bytes(uid.cpu().numpy().tobytes()), 128), rank, world)


def complete(x, mode):
    """Column sharding returns this rank's factor columns (zeros elsewhere): the caller sums over ranks."""
    if mode == "rows":
        return x
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def build(case):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import build_instance
    return {
        "maxcut": lambda: sdpa.maxcut(3000, 15000, 4),
        "mcomp": lambda: sdpa.matrix_completion(150, 120, 3000, 3, 47),
        "theta_rank_one": lambda: sdpa.lovasz_theta(300, 1500, 50),
        "theta_dense": lambda: sdpa.lovasz_theta(90, 500, 48),
        "two_block": lambda: build_instance("two_block", dict(n1=25, e1=70, n2=140, e2=600)),
    }[case]()


def check(case):
    from oracle import restate
    inst = build(case)
    G = Solver(inst, device=local, comm=new_comm())
    mode = "rows" if G.info(28) == 1 else "cols"
    out = {"case": case, "mode": mode, "owned_rows": G.info(26)}
    O = restate.OracleSolver(inst) if rank == 0 else None
    nc = len(inst.cones)
    errs = {}
    rng = np.random.default_rng(7)
    w = rng.standard_normal(inst.m)
    for c in range(nc):
        a, o = G.auv("U", "V", c, with_obj=True)
        y = complete(G.wsum_mulrk(w, True, "V", c), mode)
        x = rng.standard_normal((inst.cones[c].n, G.rank(c)))
        mv = complete(G.cg_matvec(x, "V", c), mode)
        if rank == 0:
            errs[f"auv{c}"] = rel(a, O.auv("U", "V", c))
            ro = O.obj_auv("U", "V", c)
            errs[f"obj{c}"] = abs(o - ro) / max(1.0, abs(ro))
            errs[f"wsum{c}"] = rel(y, O.wsum_mulrk(w, True, "V", c))
            errs[f"cgmv{c}"] = rel(mv, O.cg_matvec(x, "V", c))
    lam = 0.3 * rng.standard_normal(inst.m)
    G.set_vec("l", lam)
    rho = G.dinfo(6)
    lg = G.alm_prepare(rho)
    grads = [complete(G.get_factor("G", c), mode) for c in range(nc)]
    if rank == 0:
        O.vec("l")[:] = lam
        lo = O.alm_prepare(rho)
        errs["lag"] = abs(lg - lo) / lo
        for c in range(nc):
            errs[f"grad{c}"] = rel(grads[c], O.factor("G", c))
    # three inner iterations from lambda = 0 (the sharded path runs the Gram-table L-BFGS: same mathematics, other
    # rounding, hence 1e-9)
    G.set_vec("l", np.zeros(inst.m))
    G.alm_prepare(rho)
    its = [G.alm_inner_iter(rho, k) for k in range(3)]
    if rank == 0:
        O.vec("l")[:] = 0
        O.alm_prepare(rho)
        for k in range(3):
            ro, oo = O.alm_inner_iter(rho, k)
            rg, og = its[k]
            errs[f"it{k}_root"] = float(rg != ro)
            errs[f"it{k}_tau"] = abs(og["tau"] - oo["tau"]) / max(1.0, abs(oo["tau"]))
            errs[f"it{k}_lag"] = abs(og["lag_norm_sq"] - oo["lag_norm_sq"]) / oo["lag_norm_sq"]
            errs[f"it{k}_pinf"] = abs(og["pinf"] - oo["pinf"]) / max(oo["pinf"], 1e-300)
    it_cg = G.update_sdp_var_one("V", "U", 0.5, 1e-9, 800, 0)
    Vn = complete(G.get_factor("V", 0), mode)
    if rank == 0:
        it_o = O.update_sdp_var_one("V", "U", 0.5, 1e-9, 800)
        errs["cg_iters"] = float(abs(it_cg - it_o))
        out["cg_iters_oracle"] = int(it_o)
        errs["cg_V"] = rel(Vn, O.factor("V"))
    G.close()
    if rank == 0:
        out["errs"] = errs
        print("SHARDED_RESULT " + json.dumps(out), flush=True)
    dist.barrier()


def solve_golden(name):
    """Whole sharded solve of a golden instance against the REFERENCE's own result stored in tests/golden."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    g, inst = load_golden(name)
    ref = json.loads(str(g["solve"]))
    S = Solver(inst, device=local, comm=new_comm())
    res = S.solve(default_params())
    mode = "rows" if S.info(28) == 1 else "cols"
    S.close()
    if rank == 0:
        keys = ("status", "pObj", "dObj", "pInfeasL1", "pdGap", "dInfeasL1", "almInnerIter", "admmIter", "cgIter")
        print("SHARDED_RESULT " + json.dumps({"case": "solve:" + name, "mode": mode, "solve": {k: res[k] for k in keys}, "ref": ref}), flush=True)
    dist.barrier()


for case in sys.argv[1:]:
    if case.startswith("solve:"):
        solve_golden(case[6:])
    else:
        check(case)
dist.destroy_process_group()
