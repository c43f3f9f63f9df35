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


def _two_loop(g, sa, ya, sb, yb):
    """L-BFGS two-loop recursion with two history pairs (a = newest) and H0 = I, as LBFGSDirection does it
    (reference lorads_alm.c:230-391): D = -H g."""
    ba, bb = 1.0 / float(ya @ sa), 1.0 / float(yb @ sb)
    q = g.copy()
    aa = ba * float(sa @ q); q -= aa * ya
    ab = bb * float(sb @ q); q -= ab * yb
    wb = ab - bb * float(yb @ q); q += wb * sb
    wa = aa - ba * float(ya @ q); q += wa * sa
    return -q


def _gram_worker(rank, world, port, q):
    """Sharded Gram-table L-BFGS (what column-sharded runs use): every rank reduces its slice to the 8 dots of
    launch_lbfgs_pair (+ y_b.y_b kept from the previous step), ONE all-reduce completes them, and the direction is a
    linear combination with coefficients computed from the table alone."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    n, r = 200, 10
    g, sa, ya, sb, yb = (rng.standard_normal((n, r)) for _ in range(5))
    ya += 2.0 * sa      # keep y.s away from zero
    yb += 2.0 * sb
    ref = _two_loop(*(x.ravel() for x in (g, sa, ya, sb, yb))).reshape(n, r)
    cols = [k for k in range(r) if k % world == rank]
    G, SA, YA, SB, YB = (x[:, cols].ravel() for x in (g, sa, ya, sb, yb))
    table = torch.tensor([YA @ SA, YA @ YA, YA @ YB, YA @ SB, SA @ G, SB @ G, YA @ G, YB @ G, YB @ YB, YB @ SB], dtype=torch.float64)
    dist.all_reduce(table, op=dist.ReduceOp.SUM)
    d = table.numpy()
    ba, bb = 1.0 / d[0], 1.0 / d[9]
    aa = ba * d[4]
    ab = bb * (d[5] - aa * d[3])
    wb = ab - bb * (d[7] - aa * d[2] - ab * d[8])
    wa = aa - ba * (d[6] - aa * d[1] - ab * d[2] + wb * d[3])
    mine = -(G - aa * YA - ab * YB + wb * SB + wa * SA)
    err = float(np.abs(mine - ref[:, cols].ravel()).max() / np.abs(ref).max())
    q.put((rank, err))
    dist.destroy_process_group()


def test_sharded_gram_lbfgs_direction_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gram_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(err < 1e-11 for _, err in res), res
