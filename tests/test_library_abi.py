"""The C-ABI library: loads, exports every symbol include/lorads_b200.h declares, host-only entry points work,
and the compute path fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu
from lorads_b200 import capi, sdpa


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lorads_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lb2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lorads_b200.h but not exported"
    assert set(capi.EXPORTS) <= set(names)


def test_version_and_default_params():
    lib = capi.load_library()
    assert b"sm_100a" in lib.lb2_version()
    p = capi.default_params()
    # defaults of initCommandLineArgs, reference src_semi/main.c:19-43
    assert p.rhoMax == 5000.0 and p.rhoCellingADMM == 5000.0 * 200 and p.maxALMIter == 200 and p.maxADMMIter == 10000
    assert p.timesLogRank == 2.0 and p.rhoFreq == 5 and p.rhoFactor == 1.2 and p.ALMRhoFactor == 2.0
    assert p.phase1Tol == 1e-3 and p.phase2Tol == 1e-5 and p.lbfgsListLength == 2 and p.reoptLevel == 2 and p.dyrankLevel == 2


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(capi.Lb2Error) as e:
        capi.Solver(sdpa.maxcut(20, 40, 1))
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_are_rejected():
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.lb2_create(C.byref(h), 0, 1, None, None, 0) == -1
    assert lib.lb2_preprocess(None) == -1
    assert lib.lb2_info(None, 0, 0) == -1


def test_rank_rule_matches_reference_formula():
    # LORADSDetermineRank, reference src_semi/data/lorads_solver.c:290-319
    assert capi.host_rank_rule(800, 800, 1) == (14, 41)        # ceil(2 ln 800) = 14 ; floor(sqrt(1600)) + 1 = 41
    assert capi.host_rank_rule(100000, 100000, 1)[0] == 24
    assert capi.host_rank_rule(1000000, 1000000, 1)[0] == 28
    assert capi.host_rank_rule(40000, 2000000, 1)[0] == 22
    assert capi.host_rank_rule(200, 5000, 1) == (101, 101)     # m/n >= 20, n <= 400, <= 3 cones -> sqrt(2m)+1
    assert capi.host_rank_rule(5, 3, 1) == (3, 3)
    assert capi.host_rank_rule(300, 0, 1, 2.0)[0] == 1
