"""Justifies the whole-solve tolerances of tests/test_gpu_parity.py by the REFERENCE itself (CPU, no GPU).

north_star asks for primal and dual objectives within 1e-6 relative of the reference.  Short solves (MaxCut: ~150
inner iterations) meet that bar and are tested at it.  Long solves (matrix completion, Lovasz theta: thousands of
L-BFGS steps with rank updates) amplify last-bit differences: this test shows that the untouched reference, solving the
SAME problem with its constraints merely listed in another order (only the order of floating-point sums over the
constraints changes), lands on objectives that differ by MORE than 1e-6 -- so agreement to 1e-6 on those instances
is not a property the reference has with itself, and the GPU tests compare them at 3e-5 (the solver's own 1e-5 gap
tolerance times three) instead.  All six golden instances: profiles/r02_reference_sensitivity.jsonl."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def _sens(case):
    from oracle import ref
    if not ref.available(32):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from reference_sensitivity import sensitivity
    return sensitivity(case)


def test_short_solves_are_reproducible_to_rounding():
    s = _sens("maxcut_n800")
    assert s["alm_inner"][0] == s["alm_inner"][1] and s["admm_iter"][0] == s["admm_iter"][1]
    assert s["rel_pobj"] <= 1e-12 and s["rel_dobj"] <= 1e-12


def test_long_solves_of_the_reference_move_by_more_than_1e_6():
    s = _sens("mcomp_60x50")
    assert s["status"] == (1.0, 1.0) or list(s["status"]) == [1.0, 1.0]
    # same problem, same start, same code: different iteration counts and objectives
    assert s["alm_inner"][0] != s["alm_inner"][1]
    assert max(s["rel_pobj"], s["rel_dobj"]) > 1e-6
    assert max(s["rel_pobj"], s["rel_dobj"]) < 3e-5          # ... but inside the tolerance the GPU tests use
