"""Host logic of the product, checked on the CPU: the device layouts the pre-solve builds (item lists of A and
[A;C], the constraint table transposed by pattern position, C on the pattern, the symmetric adjacency, the rank-one
objective split) are exported by the host-only lb2_layout_* entry points and evaluated HERE in numpy against the
golden outputs of the reference's A(UV^T), objective and (C + A^*(w)) X.  No GPU, no library arithmetic."""
import numpy as np
import pytest

from conftest import rel_err
from lorads_b200 import capi, sdpa

TOL = 1e-12


def item_values(lay, which, U, V):
    irow, icol, coef, ptr = lay[which + "_irow"], lay[which + "_icol"], lay[which + "_coef"], lay[which + "_ptr"]
    n = lay["n"]
    if lay["dense_path"]:
        i, j = sdpa.unpack_idx(n, irow.astype(np.int64))
    else:
        i, j = irow.astype(np.int64), icol.astype(np.int64)
    z = 0.5 * (np.einsum("ij,ij->i", U[i], V[j]) + np.einsum("ij,ij->i", U[j], V[i]))
    rows = np.repeat(np.arange(ptr.size - 1), np.diff(ptr))
    out = np.zeros(ptr.size - 1)
    np.add.at(out, rows, coef * z)
    return out


def vc_values(lay, U, V):
    """A(sym(U V^T)) (compact rows) and <C, sym(U V^T)> evaluated from the vertex-centric layout exactly as
    vc_auv_kernel does: objective = sum_j U_j . (C V)_j over the objective part of the adjacency, diagonal singleton
    constraints from d_con/d_coef, the other singletons from the lower-triangular list, multi-entry constraints from
    the residual item list."""
    n = lay["n"]
    out = np.zeros(lay["n_act"])
    rows = np.repeat(np.arange(n), np.diff(lay["vc_u_ptr"]))
    static = lay["vc_u_tag"] == -1
    T = np.zeros_like(V)
    np.add.at(T, rows[static], lay["vc_u_val"][static][:, None] * V[lay["vc_u_col"][static]])
    obj = float(np.einsum("ij,ij->", U, T))
    d = lay["vc_d_con"]
    if d.size:
        has = d >= 0
        out[d[has]] = lay["vc_d_coef"][has] * np.einsum("ij,ij->i", U[has], V[has])
    if lay["vc_l_row"].size:
        j = np.repeat(np.arange(n), np.diff(lay["vc_l_ptr"]))
        assert np.array_equal(j, lay["vc_l_col"])            # flat list in (col,row) order + its column pointers
        i = lay["vc_l_row"].astype(np.int64)
        z = 0.5 * (np.einsum("ij,ij->i", U[i], V[j]) + np.einsum("ij,ij->i", U[j], V[i]))
        assert np.all(out[lay["vc_l_con"]] == 0.0)
        out[lay["vc_l_con"]] = lay["vc_l_coef"] * z
    if lay["vc_nnz_res"]:
        i, j = lay["vc_res_irow"].astype(np.int64), lay["vc_res_icol"].astype(np.int64)
        z = 0.5 * (np.einsum("ij,ij->i", U[i], V[j]) + np.einsum("ij,ij->i", U[j], V[i]))
        r = np.repeat(np.arange(lay["n_act"]), np.diff(lay["vc_res_ptr"]))
        np.add.at(out, r, lay["vc_res_coef"] * z)
    return out, obj


def vc_product(lay, wc, X, use_c=True):
    """(C + A^*(w)) X from the vertex-centric adjacency as vc_spmm_kernel evaluates it (wc = compact weights)."""
    n, npat = lay["n"], lay["psize"]
    tag, val, col = lay["vc_u_tag"], lay["vc_u_val"], lay["vc_u_col"]
    Sres = np.zeros(npat)
    if lay["vc_nnz_res"]:
        pos = np.repeat(np.arange(npat), np.diff(lay["vc_Tr_ptr"]))
        np.add.at(Sres, pos, wc[lay["vc_Tr_con"]] * lay["vc_Tr_val"])
    s = np.where(tag == -1, val if use_c else 0.0, 0.0)
    dyn = tag >= 0
    s[dyn] = wc[tag[dyn]] * val[dyn]
    res = tag <= -2
    s[res] = Sres[-2 - tag[res]]
    rows = np.repeat(np.arange(n), np.diff(lay["vc_u_ptr"]))
    # weight-dependent entries come first in every row
    assert np.all(lay["vc_u_mid"] >= lay["vc_u_ptr"][:-1]) and np.all(lay["vc_u_mid"] <= lay["vc_u_ptr"][1:])
    first_static = np.repeat(lay["vc_u_mid"], np.diff(lay["vc_u_ptr"]))
    assert np.array_equal(tag == -1, np.arange(tag.size) >= first_static)
    Y = np.zeros_like(X)
    np.add.at(Y, rows, s[:, None] * X[col])
    return Y


def check_vc(lay, U, V, w_full, ref_auv_full, ref_obj, ref_wsum, m):
    compact, obj = vc_values(lay, U, V)
    full = np.zeros(m)
    full[lay["act_idx"]] = compact
    assert rel_err(full, ref_auv_full) < TOL
    obj += lay["c_rank1"] * float(U.sum(axis=0) @ V.sum(axis=0))
    assert abs(obj - ref_obj) <= TOL * max(1.0, abs(ref_obj))
    Y = vc_product(lay, w_full[lay["act_idx"]], V) + lay["c_rank1"] * V.sum(axis=0)[None, :]
    assert rel_err(Y, ref_wsum) < TOL
    # both row orders are permutations
    for k in ("vc_order", "vc_order_l"):
        assert np.array_equal(np.sort(lay[k]), np.arange(lay["n"]))


def test_layout_reproduces_reference_operators(golden):
    name, g, inst = golden
    w = g["w"]
    for c, cone in enumerate(inst.cones):
        lay = capi.host_layout(cone, inst.m)
        U, V = g[f"U{c}"], g[f"V{c}"]
        # tiles: short lists use 128-item tiles, rows never straddle more tiles than recorded
        assert lay["tile_A"] in (128, 512) and lay["tile_AC"] in (128, 512)
        # A(sym(U V^T)) from the constraint item list, scattered to the global constraint index
        compact = item_values(lay, "A", U, V)
        full = np.zeros(inst.m)
        full[lay["act_idx"]] = compact
        assert rel_err(full, g[f"auv_UV{c}"]) < TOL
        # [A;C] list: same constraint rows, last row = <C, sym(U V^T)> (+ the rank-one part kept outside the list)
        ac = item_values(lay, "AC", U, V)
        assert rel_err(ac[:-1], compact) < TOL
        obj = ac[-1] + lay["c_rank1"] * float(U.sum(axis=0) @ V.sum(axis=0))
        assert abs(obj - float(g[f"obj_UV{c}"])) <= TOL * max(1.0, abs(float(g[f"obj_UV{c}"])))
        if lay["dense_path"]:
            continue
        # S = C + A^*(w) on the pattern from the transposed table, then Y = S V through the adjacency
        npat = lay["psize"]
        wc = w[lay["act_idx"]]
        S = lay["C_onP"].copy()
        pos = np.repeat(np.arange(npat), np.diff(lay["T_ptr"]))
        np.add.at(S, pos, wc[lay["T_con"]] * lay["T_val"])
        rows = np.repeat(np.arange(lay["n"]), np.diff(lay["adj_ptr"]))
        Y = np.zeros_like(V)
        np.add.at(Y, rows, S[lay["adj_pos"]][:, None] * V[lay["adj_col"]])
        Y += lay["c_rank1"] * V.sum(axis=0)[None, :]
        assert rel_err(Y, g[f"wsum_C{c}"]) < TOL
        # the same three operators from the vertex-centric class split
        check_vc(lay, U, V, w, g[f"auv_UV{c}"], float(g[f"obj_UV{c}"]), g[f"wsum_C{c}"], inst.m)
        # the adjacency lists every off-diagonal pattern entry twice and every diagonal entry once, rows sorted
        assert lay["adj_col"].size == 2 * npat - int((lay["P_row"] == lay["P_col"]).sum())
        for i in range(0, lay["n"], max(1, lay["n"] // 50)):
            seg = lay["adj_col"][lay["adj_ptr"][i]:lay["adj_ptr"][i + 1]]
            assert np.all(np.diff(seg) > 0)


def test_layout_of_shuffled_constraints():
    """Constraint order in the file must not change the operators (sorted-run merge vs sort fallback in the pre-solve)."""
    inst = sdpa.matrix_completion(40, 30, 400, 2, 3)
    cone = inst.cones[0]
    rng = np.random.default_rng(2)
    perm = rng.permutation(inst.m)
    counts = np.diff(cone.beg)
    take = np.concatenate([np.arange(cone.beg[0], cone.beg[1])] + [np.arange(cone.beg[1 + p], cone.beg[2 + p]) for p in perm])
    beg = np.concatenate([[0], np.cumsum([counts[0]] + [counts[1 + p] for p in perm])]).astype(np.int64)
    sh = sdpa.Cone(n=cone.n, beg=beg, idx=cone.idx[take], elem=cone.elem[take])
    a, b = capi.host_layout(cone, inst.m), capi.host_layout(sh, inst.m)
    U, V = rng.standard_normal((cone.n, 5)), rng.standard_normal((cone.n, 5))
    va, vb = item_values(a, "A", U, V), item_values(b, "A", U, V)
    fa, fb = np.zeros(inst.m), np.zeros(inst.m)
    fa[a["act_idx"]] = va
    fb[b["act_idx"]] = vb
    assert rel_err(fb, fa[perm]) < TOL          # new constraint k is old constraint perm[k]
    assert np.array_equal(a["P_row"], b["P_row"]) and np.array_equal(a["P_col"], b["P_col"])


def test_layout_is_independent_of_the_thread_count(monkeypatch):
    """The pre-solve splits its passes over threads on large cones; every array must come out identical."""
    inst = sdpa.lovasz_theta(150, 1200, 8)
    mc = sdpa.maxcut(4000, 30000, 4)
    for cone, m in ((inst.cones[0], inst.m), (mc.cones[0], mc.m)):
        monkeypatch.setenv("LORADS_B200_PRESOLVE_THREADS", "1")
        a = capi.host_layout(cone, m)
        monkeypatch.setenv("LORADS_B200_PRESOLVE_THREADS", "7")
        b = capi.host_layout(cone, m)
        assert a.keys() == b.keys()
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], k


EDGE = {
    "maxcut_split_rows": lambda: sdpa.maxcut(500, 1800, 41),
    "maxcut_n2000": lambda: sdpa.maxcut(2000, 19000, 42),
    "maxcut_isolated": lambda: sdpa.maxcut(200, 60, 46),
    "mcomp": lambda: sdpa.matrix_completion(150, 120, 3000, 3, 47),
    "theta_dense": lambda: sdpa.lovasz_theta(90, 500, 48),
    "tiny_dense": lambda: sdpa.maxcut(16, 40, 49),
    "theta_rank_one": lambda: sdpa.lovasz_theta(300, 1500, 50),
}


@pytest.mark.parametrize("case", sorted(EDGE))
def test_layout_edge_shapes_against_the_restatement(case):
    """The edge shapes of the GPU parity suite, checked on the CPU: layout evaluated in numpy vs the C restatement of
    the reference operators (oracle/lorads_oracle.c) on its srand(925) start."""
    from oracle import restate
    inst = EDGE[case]()
    O = restate.OracleSolver(inst)
    lay = capi.host_layout(inst.cones[0], inst.m)
    U, V = O.factor("U").copy(), O.factor("V").copy()
    compact = item_values(lay, "A", U, V)
    full = np.zeros(inst.m)
    full[lay["act_idx"]] = compact
    assert rel_err(full, O.auv("U", "V")) < TOL
    ac = item_values(lay, "AC", U, V)
    obj = ac[-1] + lay["c_rank1"] * float(U.sum(axis=0) @ V.sum(axis=0))
    ref_obj = O.obj_auv("U", "V")
    assert abs(obj - ref_obj) <= TOL * max(1.0, abs(ref_obj))
    if lay["dense_path"]:
        return
    w = np.random.default_rng(3).standard_normal(inst.m)
    S = lay["C_onP"].copy()
    pos = np.repeat(np.arange(lay["psize"]), np.diff(lay["T_ptr"]))
    np.add.at(S, pos, w[lay["act_idx"]][lay["T_con"]] * lay["T_val"])
    rows = np.repeat(np.arange(lay["n"]), np.diff(lay["adj_ptr"]))
    Y = np.zeros_like(V)
    np.add.at(Y, rows, S[lay["adj_pos"]][:, None] * V[lay["adj_col"]])
    Y += lay["c_rank1"] * V.sum(axis=0)[None, :]
    assert rel_err(Y, O.wsum_mulrk(w, True, "V")) < TOL
    check_vc(lay, U, V, w, O.auv("U", "V"), ref_obj, O.wsum_mulrk(w, True, "V"), inst.m)


def _with_extra_constraints(inst, extra, seed):
    """Appends constraints given as lists of (row, col, value) to the single cone of `inst`: repeated singleton
    constraints on one position, multi-entry (residual) constraints that share positions with singletons."""
    cone = inst.cones[0]
    n = cone.n
    beg, idx, elem = list(cone.beg), list(cone.idx), list(cone.elem)
    for entries in extra:
        for (i, j, v) in sorted(entries, key=lambda e: (min(e[0], e[1]), max(e[0], e[1]))):
            r, c = max(i, j), min(i, j)
            idx.append(c * n - c * (c - 1) // 2 + (r - c))      # packed lower-triangular index, column-major
            elem.append(v)
        beg.append(len(idx))
    rng = np.random.default_rng(seed)
    out = sdpa.Instance(m=inst.m + len(extra), cones=[sdpa.Cone(n=n, beg=np.array(beg), idx=np.array(idx), elem=np.array(elem))],
                        b=np.concatenate([inst.b, rng.standard_normal(len(extra))]))
    return out


def _vc_build_cases():
    cases = {k: f for k, f in EDGE.items() if k not in ("theta_dense", "tiny_dense")}
    cases["two_block_cone1"] = None
    base = sdpa.maxcut(200, 500, 77)
    cases["repeated_singletons_and_residuals"] = lambda: _with_extra_constraints(base, [
        [(3, 3, 2.0)], [(3, 3, -1.0)],                       # two more singleton constraints on a diagonal position
        [(7, 2, 0.5)], [(7, 2, 1.5)],                        # two singleton constraints on one off-diagonal position
        [(7, 2, 1.0), (9, 9, 2.0), (11, 4, -3.0)],           # residual constraint sharing positions with singletons
        [(5, 5, 1.0), (6, 6, 1.0)], [(20, 1, 1.0), (20, 1, 2.0)],   # residual; a constraint listing one position twice
    ], 5)
    return cases


@pytest.mark.parametrize("case", sorted(_vc_build_cases()))
def test_vc_adjacency_constructions_agree(case, monkeypatch):
    """The class-split adjacency is built by walking the pattern adjacency row by row (default, threaded); the
    construction from sorted entry lists (LORADS_B200_VC_BUILD=sort) must give the very same arrays, entry for entry --
    the order inside a row fixes the order of the floating-point sums of the gather kernels."""
    if case == "two_block_cone1":
        import sys
        from conftest import GOLDEN_DIR
        sys.path.insert(0, GOLDEN_DIR)
        from make_golden import build_instance
        inst = build_instance("two_block", dict(n1=25, e1=70, n2=140, e2=600))
        cones = [(inst.cones[1], inst.m), (inst.cones[0], inst.m)]
    else:
        inst = _vc_build_cases()[case]()
        cones = [(inst.cones[0], inst.m)]
    for cone, m in cones:
        for threads in ("1", "5"):
            monkeypatch.setenv("LORADS_B200_PRESOLVE_THREADS", threads)
            monkeypatch.delenv("LORADS_B200_VC_BUILD", raising=False)
            a = capi.host_layout(cone, m)
            monkeypatch.setenv("LORADS_B200_VC_BUILD", "sort")
            b = capi.host_layout(cone, m)
            assert a.keys() == b.keys()
            assert a["dense_path"] or any(k.startswith("vc_") for k in a)
            for k in a:
                if isinstance(a[k], np.ndarray):
                    assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
                else:
                    assert a[k] == b[k], k


def _dense_reference(cone, m, U, V, w):
    """A(sym(UV^T)), <C, sym(UV^T)> and (C + A^*(w)) V from dense symmetric matrices built straight from the reader
    arrays (column 0 = C, column i = A_i; packed lower-triangular indices)."""
    n = cone.n
    Z = 0.5 * (U @ V.T + V @ U.T)
    mats = []
    for c in range(m + 1):
        k = np.arange(cone.beg[c], cone.beg[c + 1])
        i, j = sdpa.unpack_idx(n, np.asarray(cone.idx)[k].astype(np.int64))
        A = np.zeros((n, n))
        np.add.at(A, (i, j), np.asarray(cone.elem)[k])
        A = A + np.tril(A, -1).T
        mats.append(A)
    auv = np.array([float((A * Z).sum()) for A in mats[1:]])
    obj = float((mats[0] * Z).sum())
    S = mats[0] + sum(wi * A for wi, A in zip(w, mats[1:]))
    return auv, obj, S @ V


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_vc_layout_on_random_mixed_constraints(seed):
    """Singleton constraints (several on one position), multi-entry constraints sharing positions with them and with the
    objective: the class-split layout evaluated in numpy against dense matrices built from the same reader arrays."""
    rng = np.random.default_rng(100 + seed)
    base = sdpa.maxcut(120, 300, 60 + seed)
    cone = base.cones[0]
    i0, j0 = sdpa.unpack_idx(cone.n, np.asarray(cone.idx[cone.beg[0]:cone.beg[1]]).astype(np.int64))
    extra = []
    for _ in range(30):
        cnt = int(rng.integers(1, 5))
        pick = rng.integers(0, i0.size, size=cnt)
        ent = {(int(i0[p]), int(j0[p])): float(rng.standard_normal()) for p in pick}      # distinct positions of C's pattern
        if rng.random() < 0.3:
            a, b = sorted(rng.integers(0, cone.n, size=2))
            ent[(int(b), int(a))] = float(rng.standard_normal())                           # a position outside C's pattern
        extra.append([(i, j, v) for (i, j), v in ent.items()])
    inst = _with_extra_constraints(base, extra, seed)
    cone = inst.cones[0]
    lay = capi.host_layout(cone, inst.m)
    assert not lay["dense_path"] and lay["vc_nnz_res"] > 0 and lay["vc_l_row"].size > 0
    U, V = rng.standard_normal((cone.n, 6)), rng.standard_normal((cone.n, 6))
    w = rng.standard_normal(inst.m)
    auv, obj, SV = _dense_reference(cone, inst.m, U, V, w)
    check_vc(lay, U, V, w, auv, obj, SV, inst.m)
    compact = item_values(lay, "A", U, V)
    full = np.zeros(inst.m)
    full[lay["act_idx"]] = compact
    assert rel_err(full, auv) < TOL
