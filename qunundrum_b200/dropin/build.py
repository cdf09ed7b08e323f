"""Compile-test of the reference-side forwarding TU (dropin.cpp).

Needs the reference's headers (/root/reference/src, used in place, never copied)
and -- only because this image lacks libgmp-dev -- the declaration shim
integration/shims/gmp.h. Output: qunundrum_b200/dropin/libqunundrum_dropin.so, which
exports the six C++ entry points of the reference with their mangled names and
depends on ../libqunundrum_b200.so and libgmp. It also compiles the reference's
src/errors.c (critical()) in place so that the test library is self-contained.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(HERE, "libqunundrum_dropin.so")


def build(reference_root: str = "/root/reference", force: bool = False) -> str | None:
    src = os.path.join(reference_root, "src")
    if not os.path.isdir(src):
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(HERE, "dropin.cpp"), os.path.join(ROOT, "include", "qunundrum_b200.h")]
    if not force and os.path.exists(LIB) and all(
            os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gmp = "/lib/x86_64-linux-gnu/libgmp.so.10"
    obj = os.path.join(HERE, "_errors.o")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", "-iquote", src, "-c",
                           os.path.join(src, "errors.c"), "-o", obj])
    subprocess.check_call(
        ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w",
         "-I", os.path.join(ROOT, "integration", "shims"), "-I", os.path.join(ROOT, "integration", "minimpi"),
         "-I", os.path.join(ROOT, "include"),
         "-iquote", src, os.path.join(HERE, "dropin.cpp"), obj,
         "-o", LIB, "-L", PKG, "-lqunundrum_b200", "-Wl,-rpath,$ORIGIN/..", gmp])
    os.remove(obj)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
