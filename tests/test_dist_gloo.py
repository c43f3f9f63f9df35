"""Host-side logic of the multi-GPU path on CPU: factor COLUMNS are sharded round-robin over ranks, every rank
evaluates A(sym(U V^T)) and dot products on its column slice, and one all-reduce (sum) restores the full value.
world_size = 2, gloo backend, the oracle's numpy arithmetic standing in for the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lorads_b200 import sdpa
    from oracle import restate
    inst = sdpa.maxcut(64, 300, 5)
    O = restate.OracleSolver(inst)
    U, V = O.factor("U").copy(), O.factor("V").copy()
    full = O.auv("U", "V")
    obj_full = O.obj_auv("U", "V")
    r = U.shape[1]
    cols = [k for k in range(r) if k % world == rank]          # lorads_b200/csrc/solver.cu: assign_columns
    O.factor("U")[:] = 0.0
    O.factor("V")[:] = 0.0
    O.factor("U")[:, cols] = U[:, cols]
    O.factor("V")[:, cols] = V[:, cols]
    part = torch.from_numpy(np.concatenate([O.auv("U", "V"), [O.obj_auv("U", "V")], [float((U[:, cols] * V[:, cols]).sum())]]))
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    got = part.numpy()
    ok = (np.allclose(got[:-2], full, rtol=1e-12, atol=1e-12) and abs(got[-2] - obj_full) <= 1e-11 * abs(obj_full)
          and abs(got[-1] - float((U * V).sum())) <= 1e-11 * abs(float((U * V).sum())))
    q.put((rank, bool(ok), len(cols)))
    dist.destroy_process_group()


def test_column_sharded_constraint_evaluation_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == 9        # rank of the n=64 MaxCut instance: min(ceil(2 ln 64), sqrt(2m)+1) = 9


def test_column_assignment_covers_every_column_once():
    for r in (1, 7, 24, 28, 36):
        for world in (1, 2, 4, 8):
            owned = [[k for k in range(r) if k % world == rank] for rank in range(world)]
            flat = sorted(c for o in owned for c in o)
            assert flat == list(range(r))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
