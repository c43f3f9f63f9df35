"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : column-sharded solver vs the single-GPU solver."""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, load_library

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_library()


def new_comm():
    """A fresh NCCL id per communicator: rank 0 creates it, everybody receives it."""
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        assert lib.lb2_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda(); dist.broadcast(uid, 0)
    return (C.create_string_buffer(bytes(uid.cpu().numpy().tobytes()), 128), rank, world)


comm = new_comm()
inst = sdpa.maxcut(20000, 100000, 2)
S = Solver(inst, device=local, comm=comm)
rho = S.dinfo(6)
lag = S.alm_prepare(rho)
outs = [S.alm_inner_iter(rho, k)[1] for k in range(6)]
a = S.auv("R", "R")
if rank == 0:
    S1 = Solver(inst, device=local)
    lag1 = S1.alm_prepare(rho)
    outs1 = [S1.alm_inner_iter(rho, k)[1] for k in range(6)]
    a1 = S1.auv("R", "R")
    print("lag", lag, lag1, abs(lag - lag1) / lag1)
    for o, o1 in zip(outs, outs1):
        print("tau %.12e %.12e  lag %.10e %.10e pinf %.8e %.8e" % (o["tau"], o1["tau"], o["lag_norm_sq"], o1["lag_norm_sq"], o["pinf"], o1["pinf"]))
    print("auv rel", float(np.linalg.norm(a - a1) / np.linalg.norm(a1)))
    ok = all(abs(o["tau"] - o1["tau"]) < 1e-8 for o, o1 in zip(outs, outs1))
    print("MULTI_GPU_CHECK", "OK" if ok else "FAIL")
# whole solve, sharded vs single GPU (rank growth + ADMM + dual infeasibility under column sharding)
from lorads_b200.capi import default_params
inst2 = sdpa.maxcut(3000, 15000, 4)
S.close()
Sm = Solver(inst2, device=local, comm=new_comm())
rm = Sm.solve(default_params())
if rank == 0:
    r1 = Solver(inst2, device=local).solve(default_params())
    keys = ("pObj", "dObj", "pInfeasL1", "pdGap", "dInfeasL1", "almInnerIter", "admmIter", "cgIter", "finalRank0", "status")
    print("sharded:", {k: rm[k] for k in keys})
    print("single :", {k: r1[k] for k in keys})
    ok2 = rm["status"] in (1, 2) and abs(rm["pObj"] - r1["pObj"]) <= 1e-5 * abs(r1["pObj"])
    print("MULTI_GPU_SOLVE", "OK" if ok2 else "FAIL")
Sm.close()
dist.barrier()
dist.destroy_process_group()
