"""Builds lorads_b200/liblorads_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblorads_b200.so")
CLI = os.path.join(HERE, "lorads_b200_cli")
SOURCES = ["kernels.cu", "solver.cu", "alg.cu", "capi.cu", "sdpa_reader.cpp"]
HEADERS = ["common.cuh", "kernels.cuh", "layout.hpp", "solver.hpp", "main_cli.cpp", os.path.join("..", "..", "include", "lorads_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2", "--shared", "-Wno-deprecated-gpu-targets",
]


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building liblorads_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    # the stand-alone command-line driver (host only, links the library)
    cli = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(CSRC, "main_cli.cpp"), "-o", CLI, "-L" + HERE, "-llorads_b200",
                          "-Wl,--disable-new-dtags,-rpath,$ORIGIN"], capture_output=True, text=True)
    if cli.returncode != 0:
        sys.stderr.write(cli.stdout + cli.stderr)
        raise RuntimeError("g++ failed building lorads_b200_cli")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
