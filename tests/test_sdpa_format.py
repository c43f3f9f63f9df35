"""Host-side data formats: packed lower-triangular index, generators, .dat-s writer/reader round trip
(reader convention of reference src_semi/io/lorads_file_io.c:21-340)."""
import numpy as np
import pytest

from lorads_b200 import sdpa


def test_pack_unpack_roundtrip():
    for n in (1, 2, 7, 64, 1001):
        p = np.arange(n * (n + 1) // 2, dtype=np.int64)
        r, c = sdpa.unpack_idx(n, p)
        assert (r >= c).all() and (c >= 0).all() and (r < n).all()
        assert np.array_equal(sdpa.pack_idx(n, r, c), p)


def test_unpack_large_n():
    n = 1_000_000
    rng = np.random.default_rng(0)
    r = rng.integers(0, n, 1000)
    c = rng.integers(0, n, 1000)
    hi, lo = np.maximum(r, c), np.minimum(r, c)
    rr, cc = sdpa.unpack_idx(n, sdpa.pack_idx(n, hi, lo))
    assert np.array_equal(rr, hi) and np.array_equal(cc, lo)


def test_maxcut_structure():
    inst = sdpa.maxcut(50, 200, 3)
    cone = inst.cones[0]
    assert inst.m == 50 and cone.n == 50
    # constraint i is e_i e_i^T with value 1
    for i in range(inst.m):
        a, b = cone.beg[i + 1], cone.beg[i + 2]
        assert b - a == 1 and cone.elem[a] == 1.0
        r, c = sdpa.unpack_idx(50, cone.idx[a:b])
        assert r[0] == i and c[0] == i
    # C = -L/4: the full symmetric matrix has zero row sums
    r, c = sdpa.unpack_idx(50, cone.idx[cone.beg[0]:cone.beg[1]])
    C = np.zeros((50, 50))
    C[r, c] = cone.elem[cone.beg[0]:cone.beg[1]]
    C = C + C.T - np.diag(np.diag(C))
    assert np.allclose(C.sum(1), 0.0)
    assert (np.diff(cone.beg) >= 0).all()


def test_random_graph_is_simple():
    hi, lo = sdpa.random_graph(100, 700, 1)
    assert hi.size == 700 and (hi > lo).all()
    assert np.unique(hi * 100 + lo).size == 700


@pytest.mark.parametrize("make", [
    lambda: sdpa.maxcut(30, 80, 1),
    lambda: sdpa.lovasz_theta(12, 20, 2),
    lambda: sdpa.matrix_completion(9, 7, 25, 2, 3),
])
def test_dat_s_roundtrip(tmp_path, make):
    inst = make()
    path = str(tmp_path / "x.dat-s")
    sdpa.write_dat_s(inst, path)
    back = sdpa.read_dat_s(path)
    assert back.m == inst.m and np.allclose(back.b, inst.b)
    for a, b in zip(inst.cones, back.cones):
        assert a.n == b.n
        assert np.array_equal(a.beg, b.beg) and np.array_equal(a.idx, b.idx) and np.allclose(a.elem, b.elem)


def test_reader_takes_a_trailing_lp_block(tmp_path):
    """A trailing negative dimension is the LP block (lorads_file_io.c:139-155): entries (j, j) of that block become
    LpMatBeg/Idx/Elem, objective negated; a diagonal block anywhere else is rejected."""
    p = tmp_path / "lp.dat-s"
    p.write_text("2\n2\n2 -3\n1.0 2.0\n0 1 1 1 1.0\n0 2 2 2 -4.0\n1 1 1 1 1.0\n1 2 1 1 0.5\n2 2 3 3 2.5\n2 2 1 1 1.5\n")
    inst = sdpa.read_dat_s(str(p))
    assert inst.lp is not None and inst.lp.n == 3 and len(inst.cones) == 1
    assert inst.lp.beg.tolist() == [0, 1, 2, 4]
    assert inst.lp.idx.tolist() == [1, 0, 2, 0]          # file order inside a constraint, like the reference reader
    assert inst.lp.elem.tolist() == [4.0, 0.5, 2.5, 1.5]
    q = tmp_path / "bad.dat-s"
    q.write_text("1\n2\n-3 2\n1.0\n0 2 1 1 1.0\n")
    with pytest.raises(ValueError):
        sdpa.read_dat_s(str(q))


def test_lp_instance_roundtrip(tmp_path):
    inst = sdpa.add_lp_block(sdpa.maxcut(30, 80, 1), 12, 3)
    assert inst.lp.n == 12 and inst.lp.beg[1] == 12          # every LP column has a cost
    path = str(tmp_path / "mix.dat-s")
    sdpa.write_dat_s(inst, path)
    back = sdpa.read_dat_s(path)
    assert np.array_equal(back.lp.beg, inst.lp.beg) and np.array_equal(back.lp.idx, inst.lp.idx)
    assert np.allclose(back.lp.elem, inst.lp.elem) and np.allclose(back.b, inst.b)
    assert np.array_equal(back.cones[0].idx, inst.cones[0].idx)


def test_tiny_entries_dropped():
    cone = sdpa._csc_from_triplets(4, 1, [0, 1, 1], [0, 1, 2], [0, 1, 0], [1e-13, 2.0, 1e-14])
    assert cone.elem.tolist() == [2.0]
    assert cone.beg.tolist() == [0, 0, 1]
