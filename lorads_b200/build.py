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
SOURCES = ["kernels.cu", "vc_kernels.cu", "xfer_kernels.cu", "solver.cu", "alg.cu", "capi.cu", "sdpa_reader.cpp"]
HEADERS = ["common.cuh", "kernels.cuh", "layout.hpp", "solver.hpp", "main_cli.cpp", os.path.join("..", "..", "include", "lorads_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2", "-Wno-deprecated-gpu-targets",
]
OBJDIR = os.path.join(HERE, "build")


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
    os.makedirs(OBJDIR, exist_ok=True)
    # one translation unit per process (the kernels have no cross-file device calls), objects rebuilt only when stale
    hdr_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if not h.endswith(".cpp"))
    hdr_time = max(hdr_time, os.path.getmtime(os.path.join(HERE, "build.py")))
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_time):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, srcp]
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out)
            failed = True
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building liblorads_b200.so")
    link = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"],
                          capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("linking liblorads_b200.so failed")
    # the stand-alone command-line driver (host only, links the library)
    cli = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(CSRC, "main_cli.cpp"), "-o", CLI, "-L" + HERE, "-llorads_b200",
                          "-Wl,--disable-new-dtags,-rpath,$ORIGIN"], capture_output=True, text=True)
    if cli.returncode != 0:
        sys.stderr.write(cli.stdout + cli.stderr)
        raise RuntimeError("g++ failed building lorads_b200_cli")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
