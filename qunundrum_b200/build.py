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
]
OBJ = os.path.join(HERE, "_obj")


def sources():
    return [os.path.join(CSRC, f) for f in
            ("qb200.cu", "qb200_text.cu", "qb200_sampler.cu", "qb200_diagk.cu", "qb200_exact.cu", "qb200_client.cu", "hostconst.cpp",
             "text_tables.cpp")]


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


def _includes(src: str) -> list[str]:
    """Headers of csrc/ a source file includes, transitively (quoted includes only)."""
    import re
    seen, todo = set(), [src]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        for inc in re.findall(r'#include\s+"([^"]+)"', open(f).read()):
            todo.append(os.path.normpath(os.path.join(os.path.dirname(f), inc)))
    return sorted(seen)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    """One object per source (rebuilt only when it or a header it includes changed,
    in parallel), then one link."""
    os.makedirs(OBJ, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or any(
            os.path.getmtime(d) > os.path.getmtime(obj) for d in _includes(src) + [__file__])
        if stale:
            cmd = [nvcc_path(), *NVCC_FLAGS, *(extra or []), "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if procs or not os.path.exists(LIB):
        cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else None)
    print(LIB)
