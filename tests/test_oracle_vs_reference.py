"""Live cross-check of the C restatement against the compiled reference (oracle/_ref) on fresh seeds and edge
shapes.  Skipped where oracle/_ref is absent (it is built by `make -C oracle ref` only where /root/reference
exists); the committed golden vectors cover the same functions everywhere else."""
import os
import tempfile

import numpy as np
import pytest

from conftest import rel_err
from lorads_b200 import sdpa
from oracle import ref, restate

pytestmark = pytest.mark.skipif(not ref.available(32), reason="oracle/_ref not built")

CASES = {
    "maxcut_sparse": lambda: sdpa.maxcut(260, 1500, 21),
    "maxcut_isolated_nodes": lambda: sdpa.maxcut(90, 30, 22),          # zero diagonal entries of C are dropped
    "mcomp_sparse_scratch": lambda: sdpa.matrix_completion(120, 100, 900, 2, 23),
    "theta_dense_C": lambda: sdpa.lovasz_theta(40, 150, 24),
    "tiny_dense_n_lt_20": lambda: sdpa.maxcut(12, 20, 25),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_reference(name):
    inst = CASES[name]()
    d = tempfile.mkdtemp()
    path = os.path.join(d, name + ".dat-s")
    sdpa.write_dat_s(inst, path)
    R = ref.RefSolver(path, 32)
    O = restate.OracleSolver(inst)
    assert R.rank() == O.rank() and R.info(6) == O.info(6) and R.info(4) == O.info(4)
    for f in "RUV":
        assert np.array_equal(R.factor(f), O.factor(f))
    assert rel_err(O.auv("U", "V"), R.auv("U", "V")) < 1e-13
    assert abs(O.obj_auv("U", "V") - R.obj_auv("U", "V")) <= 1e-12 * max(1.0, abs(R.obj_auv("U", "V")))
    w = np.random.default_rng(1).standard_normal(R.m)
    assert rel_err(O.wsum_mulrk(w, True, "R"), R.wsum_mulrk(w, True, "R")) < 1e-13
    x = np.random.default_rng(2).standard_normal(R.factor("U").shape)
    assert rel_err(O.cg_matvec(x, "V"), R.cg_matvec(x, "V")) < 1e-13
    rho = R.dinfo(6)
    assert abs(O.alm_prepare(rho) - R.alm_prepare(rho)) <= 1e-12 * R.alm_prepare(rho)
    assert O.update_sdp_var_one("U", "V", 0.7, 1e-9, 800) == R.update_sdp_var_one("U", "V", 0.7, 1e-9, 800)
    assert rel_err(O.factor("U"), R.factor("U")) < 1e-9
    R.alm_prepare(rho); O.alm_prepare(rho)
    for k in range(4):
        ra, oa = R.alm_inner_iter(rho, k), O.alm_inner_iter(rho, k)
        assert ra[0] == oa[0]
        assert abs(ra[1]["tau"] - oa[1]["tau"]) <= 1e-10 * 10 ** k * max(1.0, abs(ra[1]["tau"]))


def test_reader_output_matches_generator():
    inst = sdpa.maxcut(70, 300, 31)
    d = tempfile.mkdtemp()
    path = os.path.join(d, "r.dat-s")
    sdpa.write_dat_s(inst, path)
    R = ref.RefSolver(path, 32)
    beg, idx, elem = R.reader_csc()
    c = inst.cones[0]
    assert np.array_equal(beg, c.beg)
    for col in range(inst.m + 1):
        a, b = beg[col], beg[col + 1]
        o1, o2 = np.argsort(idx[a:b]), np.argsort(c.idx[a:b])
        assert np.array_equal(idx[a:b][o1], c.idx[a:b][o2]) and np.allclose(elem[a:b][o1], c.elem[a:b][o2])
