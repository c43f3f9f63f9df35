"""The link-time drop-in: the reference's UNMODIFIED main.c + reader linked against liblorads_b200.so through
integration/lorads_dropin.c (oracle/_ref/lorads_gpu32, built by `make -C oracle dropin` where /root/reference
exists) solves a .dat-s file on the GPU and prints the reference's own result block."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, have_gpu, load_golden
from lorads_b200 import sdpa

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]

BIN = os.path.join(ROOT, "oracle", "_ref", "lorads_gpu32")
REF = os.path.join(ROOT, "oracle", "_ref", "lorads_ref32")


def parse(out):
    def grab(label):
        m = re.search(re.escape(label) + r"\s*:\s*([-+0-9.eE]+)", out)
        return float(m.group(1)) if m else None
    return dict(pobj=grab("1.Primal Objective:"), dobj=grab("2.Dual Objective:"), pinf=grab("1.Constraint Violation(1)"),
                dinf=grab("2.Dual Infeasibility(1)"), gap=grab("3.Primal Dual Gap"))


@pytest.mark.skipif(not os.path.exists(BIN), reason="drop-in binary not built (needs /root/reference at build time)")
def test_reference_driver_runs_on_the_cuda_library(tmp_path):
    g, inst = load_golden("maxcut_n800")
    path = str(tmp_path / "mc800.dat-s")
    sdpa.write_dat_s(inst, path)
    out = subprocess.run([BIN, path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = parse(out.stdout)
    ref = json.loads(str(g["solve"]))
    assert "End Program due to reaching `Official terminate criteria`" in out.stdout
    assert abs(got["pobj"] - ref["pobj"]) <= 1e-6 * (1 + abs(ref["pobj"]))
    assert abs(got["dobj"] - ref["dobj"]) <= 1e-6 * (1 + abs(ref["dobj"]))
    assert got["pinf"] <= 1e-5 and got["gap"] <= 5e-5
    if os.path.exists(REF):   # same file through the reference CPU binary, same printed block
        cpu = subprocess.run([REF, path], capture_output=True, text=True, timeout=300)
        c = parse(cpu.stdout)
        assert abs(got["pobj"] - c["pobj"]) <= 1e-6 * (1 + abs(c["pobj"]))
        assert abs(got["dobj"] - c["dobj"]) <= 1e-6 * (1 + abs(c["dobj"]))


def test_standalone_cli_solves_a_file(tmp_path):
    """lorads_b200_cli: the library's own reader + solve, no reference tree involved."""
    cli = os.path.join(ROOT, "lorads_b200", "lorads_b200_cli")
    assert os.path.exists(cli), "build with python -m lorads_b200.build"
    g, inst = load_golden("maxcut_n120")
    path = str(tmp_path / "mc120.dat-s")
    sdpa.write_dat_s(inst, path)
    out = subprocess.run([cli, path, "--quiet"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = parse(out.stdout)
    ref = json.loads(str(g["solve"]))
    assert abs(got["pobj"] - ref["pobj"]) <= 1e-6 * (1 + abs(ref["pobj"]))
    assert abs(got["dobj"] - ref["dobj"]) <= 1e-6 * (1 + abs(ref["dobj"]))
