"""SDPA ``.dat-s`` instances in the reference reader's in-memory format.

The reference's reader ``LReadSDPA`` (reference ``src_semi/io/lorads_file_io.c:21``)
hands every PSD block to the solver as a CSC matrix with ``m + 1`` columns over
the packed lower-triangular index ``PACK_IDX(n, i, j) = (2n - j - 1) * j / 2 + i``
(``src_semi/lorads_utils.h:45``): column 0 is the objective block (negated on
read, ``lorads_file_io.c:279-281``) and column ``i`` is constraint ``A_i``.
Those three arrays per block (``coneMatBeg``, ``coneMatIdx``, ``coneMatElem``) and
the right-hand side ``rowRHS`` are the *input format* of the device data layer, so
this module produces exactly them:

* seeded synthetic generators for the BASELINE.json configurations (MaxCut,
  Lovasz theta, matrix completion; forms in SURVEY.md section 8d),
* a ``.dat-s`` writer (so the reference CPU binary can be run on the same
  instance) and a reader that mirrors ``LReadSDPA``'s output.

Everything here is host-side numpy; nothing touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def pack_idx(n: int, row, col):
    """Packed lower-tri position of (row >= col); reference ``lorads_utils.h:45``."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    return (2 * n - col - 1) * col // 2 + row


def unpack_idx(n: int, packed):
    """Inverse of :func:`pack_idx` (reference ``tsp_decompress``, ``lorads_sparse_opts.c:37``)."""
    packed = np.asarray(packed, dtype=np.int64)
    # column j starts at s(j) = j*n - j*(j-1)/2 ; solve s(j) <= p
    nn = float(n)
    col = np.floor(((2 * nn + 1) - np.sqrt((2 * nn + 1) ** 2 - 8.0 * packed)) / 2.0).astype(np.int64)
    start = col * n - col * (col - 1) // 2
    bad = start > packed
    col[bad] -= 1
    start = col * n - col * (col - 1) // 2
    nxt = (col + 1) * n - (col + 1) * col // 2
    bad = packed >= nxt
    col[bad] += 1
    start = col * n - col * (col - 1) // 2
    row = packed - start + col
    return row, col


@dataclass
class Cone:
    """One PSD block in reader-output form (column 0 = C, column i = A_i)."""

    n: int
    beg: np.ndarray   # int64, length m + 2
    idx: np.ndarray   # int64 packed lower-tri positions
    elem: np.ndarray  # float64


@dataclass
class LpBlock:
    """The LP (diagonal) block in reader-output form (``LpMatBeg/Idx/Elem``, ``lorads_file_io.c:342-355``):
    CSC with m + 1 columns, column 0 = objective (negated on read), column i = constraint i; the row index is the
    LP column."""

    n: int
    beg: np.ndarray   # int64, length m + 2
    idx: np.ndarray   # int64 LP column of the entry
    elem: np.ndarray  # float64


@dataclass
class Instance:
    m: int
    b: np.ndarray
    cones: List[Cone]
    name: str = "instance"
    meta: dict = field(default_factory=dict)
    lp: Optional[LpBlock] = None

    @property
    def blk_dims(self):
        return [c.n for c in self.cones]


def _sorted_unique(x: np.ndarray) -> np.ndarray:
    """np.unique for int64 keys by sort + neighbour compare (numpy's hash-based unique is ~10x slower at 5e6 keys)."""
    x = np.sort(x)
    if x.size < 2:
        return x
    keep = np.empty(x.size, dtype=bool)
    keep[0] = True
    np.not_equal(x[1:], x[:-1], out=keep[1:])
    return x[keep]


def _csc_from_triplets(n: int, m: int, con, row, col, val) -> Cone:
    """Triplets (constraint index 0..m, row>=col, value) -> reader CSC, rows sorted inside a column."""
    con = np.asarray(con, dtype=np.int64)
    packed = pack_idx(n, row, col)
    val = np.asarray(val, dtype=np.float64)
    keep = np.abs(val) >= 1e-12           # the reader drops tiny entries, lorads_file_io.c:250-256
    con, packed, val = con[keep], packed[keep], val[keep]
    order = np.lexsort((packed, con))
    con, packed, val = con[order], packed[order], val[order]
    beg = np.cumsum(np.bincount(con + 1, minlength=m + 2)).astype(np.int64)
    return Cone(n=n, beg=beg, idx=packed.astype(np.int64), elem=val)


def random_graph(n: int, n_edges: int, seed: int):
    """Uniform random simple graph: returns (i, j) with i > j, exactly n_edges distinct pairs."""
    rng = np.random.default_rng(seed)
    have = np.empty(0, dtype=np.int64)
    while have.size < n_edges:
        need = int((n_edges - have.size) * 1.1) + 16
        a = rng.integers(0, n, size=need, dtype=np.int64)
        b = rng.integers(0, n, size=need, dtype=np.int64)
        ok = a != b
        hi = np.maximum(a[ok], b[ok])
        lo = np.minimum(a[ok], b[ok])
        key = hi * n + lo
        have = _sorted_unique(np.concatenate([have, key]))
        if have.size > n_edges:
            have = rng.permutation(have)[:n_edges]
            have.sort()
    return have // n, have % n


def maxcut(n: int, n_edges: int, seed: int) -> Instance:
    """MaxCut SDP relaxation: m = n, A_i = e_i e_i^T, b = 1, F_0 = L/4 so C = -L/4 (unit weights)."""
    hi, lo = random_graph(n, n_edges, seed)
    deg = np.bincount(hi, minlength=n) + np.bincount(lo, minlength=n)
    diag = np.arange(n, dtype=np.int64)
    con = np.concatenate([np.zeros(n, np.int64), np.zeros(hi.size, np.int64), diag + 1])
    row = np.concatenate([diag, hi, diag])
    col = np.concatenate([diag, lo, diag])
    val = np.concatenate([-deg / 4.0, np.full(hi.size, 0.25), np.ones(n)])
    cone = _csc_from_triplets(n, n, con, row, col, val)
    return Instance(m=n, b=np.ones(n), cones=[cone], name=f"maxcut_n{n}_e{n_edges}_s{seed}",
                    meta={"kind": "maxcut", "n": n, "edges": int(n_edges), "seed": seed})


def lovasz_theta(n: int, n_edges: int, seed: int) -> Instance:
    """Lovasz theta: constraint 1 = I (b = 1), constraint k = single entry (i, j) in E (b = 0), F_0 = J."""
    hi, lo = random_graph(n, n_edges, seed)
    m = n_edges + 1
    tr_r, tr_c = np.tril_indices(n)
    diag = np.arange(n, dtype=np.int64)
    con = np.concatenate([np.zeros(tr_r.size, np.int64), np.ones(n, np.int64), np.arange(2, m + 1, dtype=np.int64)])
    row = np.concatenate([tr_r, diag, hi])
    col = np.concatenate([tr_c, diag, lo])
    val = np.concatenate([-np.ones(tr_r.size), np.ones(n), np.ones(hi.size)])
    cone = _csc_from_triplets(n, m, con, row, col, val)
    b = np.zeros(m)
    b[0] = 1.0
    return Instance(m=m, b=b, cones=[cone], name=f"theta_n{n}_e{n_edges}_s{seed}",
                    meta={"kind": "theta", "n": n, "edges": int(n_edges), "seed": seed})


def matrix_completion(n1: int, n2: int, n_samples: int, rank: int, seed: int) -> Instance:
    """Nuclear-norm matrix completion as an SDP on an (n1+n2) block: F_0 = -I (C = I), constraint k =
    entry (i, n1 + j) with value 0.5, b_k = M_ij for a rank-`rank` Gaussian ground truth M."""
    rng = np.random.default_rng(seed)
    left = rng.standard_normal((n1, rank))
    right = rng.standard_normal((n2, rank))
    key = np.empty(0, dtype=np.int64)
    while key.size < n_samples:
        need = int((n_samples - key.size) * 1.1) + 16
        k = rng.integers(0, n1 * n2, size=need, dtype=np.int64)
        key = _sorted_unique(np.concatenate([key, k]))
        if key.size > n_samples:
            key = rng.permutation(key)[:n_samples]
            key.sort()
    si, sj = key // n2, key % n2
    b = np.einsum("ij,ij->i", left[si], right[sj])
    n = n1 + n2
    m = n_samples
    diag = np.arange(n, dtype=np.int64)
    con = np.concatenate([np.zeros(n, np.int64), np.arange(1, m + 1, dtype=np.int64)])
    row = np.concatenate([diag, n1 + sj])
    col = np.concatenate([diag, si])
    val = np.concatenate([np.ones(n), np.full(m, 0.5)])
    cone = _csc_from_triplets(n, m, con, row, col, val)
    return Instance(m=m, b=b, cones=[cone], name=f"mc_{n1}x{n2}_s{n_samples}_r{rank}_seed{seed}",
                    meta={"kind": "matrix_completion", "n1": n1, "n2": n2, "samples": int(n_samples), "seed": seed})


def _lp_from_triplets(n_lp: int, m: int, con, col, val) -> LpBlock:
    con = np.asarray(con, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    keep = np.abs(val) >= 1e-12
    con, col, val = con[keep], col[keep], val[keep]
    order = np.lexsort((col, con))
    con, col, val = con[order], col[order], val[order]
    beg = np.zeros(m + 2, dtype=np.int64)
    np.add.at(beg, con + 1, 1)
    return LpBlock(n=n_lp, beg=np.cumsum(beg), idx=col, elem=val)


def add_lp_block(inst: Instance, n_lp: int, seed: int, max_rows_per_col: int = 3) -> Instance:
    """Mixed SDP + LP instance: `n_lp` non-negative variables x are appended to an SDP instance.  Column j touches
    1..max_rows_per_col random constraints with coefficients in [0.5, 1.5] and costs c_j in [0.1, 1]; the right-hand
    side grows by A_lp x0 for a random x0 >= 0, so the mixed problem stays feasible and bounded:
        min <C, X> + c^T x   s.t.  A(X) + A_lp x = b + A_lp x0,  X psd,  x >= 0."""
    rng = np.random.default_rng(seed)
    m = inst.m
    cols, rows, vals = [], [], []
    for j in range(n_lp):
        k = int(rng.integers(1, max_rows_per_col + 1))
        r = rng.choice(m, size=min(k, m), replace=False)
        rows.append(np.sort(r)); cols.append(np.full(r.size, j)); vals.append(rng.uniform(0.5, 1.5, r.size))
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    cost = rng.uniform(0.1, 1.0, n_lp)
    x0 = rng.uniform(0.0, 1.0, n_lp)
    b = inst.b.copy()
    np.add.at(b, rows, vals * x0[cols])
    lp = _lp_from_triplets(n_lp, m, np.concatenate([np.zeros(n_lp, np.int64), rows + 1]),
                           np.concatenate([np.arange(n_lp), cols]), np.concatenate([cost, vals]))
    return Instance(m=m, b=b, cones=inst.cones, name=f"{inst.name}_lp{n_lp}_s{seed}", meta=dict(inst.meta, lp=n_lp), lp=lp)


def multi_block(instances: List[Instance], name: str = "multiblock") -> Instance:
    """Stack single-block instances with the same m into one multi-cone instance (constraints are shared:
    A_i = blkdiag(A_i^1, A_i^2, ...), b = sum of the b's)."""
    m = instances[0].m
    assert all(i.m == m for i in instances)
    cones = [c for inst in instances for c in inst.cones]
    b = np.sum([inst.b for inst in instances], axis=0)
    return Instance(m=m, b=b, cones=cones, name=name, meta={"kind": "multi_block"})


def write_dat_s(inst: Instance, path: str) -> None:
    """Write the instance as SDPA sparse text.  The objective block is written as F_0 = -C because the
    reference reader negates constraint index 0 (``lorads_file_io.c:279-281``)."""
    chunks = []
    for k, cone in enumerate(inst.cones):
        counts = np.diff(cone.beg)
        con = np.repeat(np.arange(inst.m + 1, dtype=np.int64), counts)
        row, col = unpack_idx(cone.n, cone.idx)
        val = np.where(con == 0, -cone.elem, cone.elem)
        blk = np.full(con.size, k + 1, dtype=np.int64)
        # SDPA entries are upper-triangular 1-based (i <= j)
        chunks.append((con, blk, col + 1, row + 1, val))
    if inst.lp is not None:
        # the LP block is the trailing diagonal block of (negative) dimension -nLpCols; entries are (j, j)
        lp = inst.lp
        con = np.repeat(np.arange(inst.m + 1, dtype=np.int64), np.diff(lp.beg))
        val = np.where(con == 0, -lp.elem, lp.elem)
        blk = np.full(con.size, len(inst.cones) + 1, dtype=np.int64)
        chunks.append((con, blk, lp.idx + 1, lp.idx + 1, val))
    with open(path, "w") as f:
        dims = [str(c.n) for c in inst.cones] + ([str(-inst.lp.n)] if inst.lp is not None else [])
        f.write(f"{inst.m}\n{len(dims)}\n")
        f.write(" ".join(dims) + "\n")
        f.write(" ".join(repr(float(x)) for x in inst.b) + "\n")
        for con, blk, i, j, val in chunks:
            lines = [f"{a} {b_} {c} {d} {e!r}" for a, b_, c, d, e in zip(con.tolist(), blk.tolist(), i.tolist(), j.tolist(), val.tolist())]
            f.write("\n".join(lines))
            f.write("\n")


def read_dat_s(path: str) -> Instance:
    """Host-side reader with the output convention of ``LReadSDPA`` (a trailing negative dimension is the LP
    block, ``lorads_file_io.c:139-155``)."""
    with open(path) as f:
        lines = f.read().split("\n")
    pos = 0
    while lines[pos].startswith("*") or lines[pos].startswith('"'):
        pos += 1
    m = int(lines[pos].split()[0]); pos += 1
    nblk = int(lines[pos].split()[0]); pos += 1
    dims = [int(t) for t in lines[pos].replace("{", " ").replace("}", " ").replace("(", " ").replace(")", " ").replace(",", " ").split()]
    pos += 1
    assert len(dims) == nblk
    n_lp = 0
    if dims and dims[-1] < 0:
        n_lp = -dims.pop()
    if any(d <= 0 for d in dims):
        raise ValueError("only one diagonal (LP) block is supported and it must be the last one")
    b = np.array([float(t) for t in lines[pos].replace(",", " ").replace("{", " ").replace("}", " ").split()], dtype=np.float64)
    pos += 1
    assert b.size == m
    body = [ln for ln in lines[pos:] if ln.strip()]
    data = np.array([ln.split()[:5] for ln in body], dtype=np.float64) if body else np.zeros((0, 5))
    con = data[:, 0].astype(np.int64)
    blk = data[:, 1].astype(np.int64) - 1
    i = data[:, 2].astype(np.int64) - 1
    j = data[:, 3].astype(np.int64) - 1
    val = data[:, 4].copy()
    val[con == 0] *= -1.0
    cones = []
    for k, n in enumerate(dims):
        sel = blk == k
        hi = np.maximum(i[sel], j[sel])
        lo = np.minimum(i[sel], j[sel])
        cones.append(_csc_from_triplets(n, m, con[sel], hi, lo, val[sel]))
    lp = None
    if n_lp > 0:
        sel = blk == len(dims)
        keep = np.abs(val[sel]) >= 1e-12
        c_, i_, v_ = con[sel][keep], i[sel][keep], val[sel][keep]
        order = np.argsort(c_, kind="stable")          # the reader keeps file order inside a constraint
        beg = np.zeros(m + 2, dtype=np.int64)
        np.add.at(beg, c_ + 1, 1)
        lp = LpBlock(n=n_lp, beg=np.cumsum(beg), idx=i_[order], elem=v_[order])
    return Instance(m=m, b=b, cones=cones, name=path, lp=lp)
