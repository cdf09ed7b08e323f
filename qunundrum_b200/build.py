"""Build the CUDA library in-tree: qunundrum_b200/libqunundrum_b200.so.

nvcc cross-compiles for sm_100a without a GPU. The built .so is git-ignored but
travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqunundrum_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # FMAs are explicit; error-free transforms must not be contracted
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared",
]


def sources():
    return [os.path.join(CSRC, "qb200.cu"), os.path.join(CSRC, "hostconst.cpp")]


def deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "qunundrum_b200.h"))
    return out


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    stale = force or not os.path.exists(LIB) or any(
        os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps())
    if stale:
        cmd = [nvcc_path(), *NVCC_FLAGS, *(extra or []), *sources(), "-o", LIB]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else None)
    print(LIB)
