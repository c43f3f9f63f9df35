"""The library's .dat-s reader (lb2_read_sdpa) against the Python reader, the generators and -- where oracle/_ref
is present -- the reference's own LReadSDPA output."""
import os
import time

import numpy as np
import pytest

from lorads_b200 import capi, sdpa


def same_csc(a, b, m):
    assert a.n == b.n and np.array_equal(a.beg, b.beg)
    for c in range(m + 1):
        s, e = a.beg[c], a.beg[c + 1]
        oa, ob = np.argsort(a.idx[s:e], kind="stable"), np.argsort(b.idx[s:e], kind="stable")
        assert np.array_equal(a.idx[s:e][oa], b.idx[s:e][ob])
        assert np.array_equal(a.elem[s:e][oa], b.elem[s:e][ob])


@pytest.mark.parametrize("make", [
    lambda: sdpa.maxcut(60, 200, 1),
    lambda: sdpa.lovasz_theta(15, 30, 2),
    lambda: sdpa.matrix_completion(12, 9, 40, 2, 3),
    lambda: sdpa.multi_block([sdpa.maxcut(10, 20, 4), sdpa.maxcut(10, 25, 5)]),
])
def test_matches_generator_and_python_reader(tmp_path, make):
    inst = make()
    path = str(tmp_path / "x.dat-s")
    sdpa.write_dat_s(inst, path)
    got = capi.read_sdpa(path)
    ref = sdpa.read_dat_s(path)
    assert got.m == inst.m and np.array_equal(got.b, inst.b)
    for a, b, c in zip(got.cones, inst.cones, ref.cones):
        same_csc(a, b, inst.m)
        same_csc(a, c, inst.m)


def test_format_variants(tmp_path):
    p = tmp_path / "v.dat-s"
    p.write_text('"a comment\n* another\n2 = m\n1 = nblocks\n{3}\n{1.5, -2}\n0 1 1 1 2.0\n0 1 2 1 1e-13\n1 1 3 1 0.5\n1 1 2 2 -1\n2 1 3 3 4\n')
    inst = capi.read_sdpa(str(p))
    assert inst.m == 2 and inst.blk_dims == [3] and inst.b.tolist() == [1.5, -2.0]
    c = inst.cones[0]
    assert c.beg.tolist() == [0, 1, 3, 4]          # the 1e-13 entry is dropped
    assert c.elem.tolist() == [-2.0, 0.5, -1.0, 4.0]   # objective negated
    assert c.idx.tolist() == [0, 2, 3, 5]          # PACK_IDX(3, row, col): (0,0)=0 (2,0)=2 (1,1)=3 (2,2)=5


def test_lp_block_is_returned(tmp_path):
    p = tmp_path / "lp.dat-s"
    p.write_text("1\n2\n2 -3\n1.0\n0 1 1 1 1.0\n0 2 3 3 2.0\n1 1 1 1 1.0\n1 2 2 2 1.0\n")
    inst = capi.read_sdpa(str(p))
    assert inst.lp.n == 3 and inst.lp.beg.tolist() == [0, 1, 2]
    assert inst.lp.idx.tolist() == [2, 1] and inst.lp.elem.tolist() == [-2.0, 1.0]
    # same arrays as the pure-python reader on a generated mixed instance
    mix = sdpa.add_lp_block(sdpa.maxcut(40, 100, 2), 30, 4)
    path = str(tmp_path / "mix.dat-s")
    sdpa.write_dat_s(mix, path)
    a, b = capi.read_sdpa(path), sdpa.read_dat_s(path)
    assert np.array_equal(a.lp.beg, b.lp.beg) and np.array_equal(a.lp.idx, b.lp.idx) and np.array_equal(a.lp.elem, b.lp.elem)
    assert np.array_equal(a.lp.beg, mix.lp.beg) and np.array_equal(a.lp.idx, mix.lp.idx)


def test_errors(tmp_path):
    with pytest.raises(capi.Lb2Error):
        capi.read_sdpa(str(tmp_path / "missing.dat-s"))
    p = tmp_path / "bad.dat-s"
    p.write_text("2\n1\n3\n1 2\n0 1 5 1 1.0\n")
    with pytest.raises(capi.Lb2Error):
        capi.read_sdpa(str(p))


def test_against_reference_reader_and_speed(tmp_path):
    from oracle import ref
    if not ref.available(32):
        pytest.skip("oracle/_ref not built")
    inst = sdpa.maxcut(3000, 20000, 8)
    path = str(tmp_path / "r.dat-s")
    sdpa.write_dat_s(inst, path)
    t0 = time.perf_counter()
    got = capi.read_sdpa(path)
    t_lib = time.perf_counter() - t0
    R = ref.RefSolver(path, 32)
    beg, idx, elem = R.reader_csc()
    same_csc(got.cones[0], sdpa.Cone(n=3000, beg=beg, idx=idx, elem=elem), inst.m)
    assert t_lib < 1.0


def test_binary_cache_roundtrip(tmp_path):
    """lb2_sdpa_save writes the reader's arrays; lb2_read_sdpa recognises the file by its magic."""
    mix = sdpa.add_lp_block(sdpa.maxcut(50, 160, 3), 20, 6)
    txt, binp = str(tmp_path / "mix.dat-s"), str(tmp_path / "mix.lb2")
    sdpa.write_dat_s(mix, txt)
    a = capi.read_sdpa(txt, save_binary=binp)
    b = capi.read_sdpa(binp)
    assert b.m == a.m and np.array_equal(a.b, b.b) and len(a.cones) == len(b.cones)
    for x, y in zip(a.cones, b.cones):
        assert x.n == y.n and np.array_equal(x.beg, y.beg) and np.array_equal(x.idx, y.idx) and np.array_equal(x.elem, y.elem)
    assert b.lp.n == a.lp.n and np.array_equal(a.lp.beg, b.lp.beg) and np.array_equal(a.lp.idx, b.lp.idx) and np.array_equal(a.lp.elem, b.lp.elem)
    # truncated cache files are rejected
    data = open(binp, "rb").read()
    bad = tmp_path / "bad.lb2"
    bad.write_bytes(data[: len(data) // 2])
    with pytest.raises(capi.Lb2Error):
        capi.read_sdpa(str(bad))


def test_threaded_parse_keeps_file_order(tmp_path):
    """Files above 8 MB are parsed by several threads; the arrays must equal the single-thread result (checked against
    the pure-python reader, which keeps file order inside a constraint)."""
    inst = sdpa.maxcut(120000, 500000, 9)
    path = str(tmp_path / "big.dat-s")
    sdpa.write_dat_s(inst, path)
    assert os.path.getsize(path) > (8 << 20)
    a = capi.read_sdpa(path)
    c0 = inst.cones[0]
    assert np.array_equal(a.cones[0].beg, c0.beg) and np.array_equal(a.cones[0].idx, c0.idx) and np.allclose(a.cones[0].elem, c0.elem)
