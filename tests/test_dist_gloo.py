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


# ---------------------------------------------------------------------------------------------------------------------
# row-slab / cone-block sharding (the default scheme): every rank owns a contiguous slab of rows, evaluates the pattern
# entries of its COLUMNS in A(sym(U V^T)) and its ROWS of (C + A^*(w)) V from the vertex-centric layout, one all-reduce
# (sum) completes the m-vector and the objective, an all-gather completes the product.
# ---------------------------------------------------------------------------------------------------------------------
def _row_slab_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lorads_b200 import capi, sdpa
    from oracle import restate
    ok = True
    for inst in (sdpa.maxcut(90, 400, 6), sdpa.matrix_completion(40, 30, 500, 2, 8), sdpa.lovasz_theta(70, 300, 9)):
        O = restate.OracleSolver(inst)
        lay = capi.host_layout(inst.cones[0], inst.m)
        if not lay.get("vc_on"):
            continue
        n = lay["n"]
        U, V = O.factor("U").copy(), O.factor("V").copy()
        # slabs of equal work: the weights Solver::setup_row_partition uses
        wgt = 8 + np.diff(lay["vc_u_ptr"]) + 2 * np.diff(lay["vc_l_ptr"])
        cum = np.cumsum(wgt)
        cuts = [0] + [int(np.searchsorted(cum, cum[-1] * k / world) + 1) for k in range(1, world)] + [n]
        lo, hi = cuts[rank], cuts[rank + 1]
        mine = np.zeros(n, bool)
        mine[lo:hi] = True
        # --- A(sym(U V^T)) of the owned columns + objective partial
        out = np.zeros(lay["n_act"])
        rows = np.repeat(np.arange(n), np.diff(lay["vc_u_ptr"]))
        sel = (lay["vc_u_tag"] == -1) & mine[rows]
        T = np.zeros_like(V)
        np.add.at(T, rows[sel], lay["vc_u_val"][sel][:, None] * V[lay["vc_u_col"][sel]])
        obj = float(np.einsum("ij,ij->", U[lo:hi], T[lo:hi]))
        d = lay["vc_d_con"]
        if d.size:
            has = (d >= 0) & mine
            out[d[has]] = lay["vc_d_coef"][has] * np.einsum("ij,ij->i", U[has], V[has])
        if lay["vc_l_row"].size:
            j = np.repeat(np.arange(n), np.diff(lay["vc_l_ptr"]))
            keep = mine[j]
            i = lay["vc_l_row"].astype(np.int64)[keep]
            jj = j[keep]
            z = 0.5 * (np.einsum("ij,ij->i", U[i], V[jj]) + np.einsum("ij,ij->i", U[jj], V[i]))
            out[lay["vc_l_con"][keep]] = lay["vc_l_coef"][keep] * z
        if lay["vc_nnz_res"] and rank == 0:                 # the residual item list is evaluated by the cone's lead rank
            i, j = lay["vc_res_irow"].astype(np.int64), lay["vc_res_icol"].astype(np.int64)
            z = 0.5 * (np.einsum("ij,ij->i", U[i], V[j]) + np.einsum("ij,ij->i", U[j], V[i]))
            r = np.repeat(np.arange(lay["n_act"]), np.diff(lay["vc_res_ptr"]))
            np.add.at(out, r, lay["vc_res_coef"] * z)
        if rank == 0:
            obj += lay["c_rank1"] * float(U.sum(axis=0) @ V.sum(axis=0))
        part = torch.from_numpy(np.concatenate([out, [obj]]))
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        full = np.zeros(inst.m)
        full[lay["act_idx"]] = part.numpy()[:-1]
        ref = O.auv("U", "V")
        ok &= bool(np.linalg.norm(full - ref) <= 1e-12 * np.linalg.norm(ref))
        ok &= bool(abs(part.numpy()[-1] - O.obj_auv("U", "V")) <= 1e-12 * max(1.0, abs(O.obj_auv("U", "V"))))
        # --- owned rows of (C + A^*(w)) V, completed by an all-gather
        from test_host_layout import vc_product
        w = np.random.default_rng(3).standard_normal(inst.m)
        Y = vc_product(lay, w[lay["act_idx"]], V) + lay["c_rank1"] * V.sum(axis=0)[None, :]
        Y[~mine] = 0.0                                          # a rank only computes its slab
        t = torch.from_numpy(Y)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                # disjoint slabs: the sum is the all-gather
        refY = O.wsum_mulrk(w, True, "V")
        ok &= bool(np.linalg.norm(t.numpy() - refY) <= 1e-12 * np.linalg.norm(refY))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_row_slab_sharded_operators_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_slab_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
