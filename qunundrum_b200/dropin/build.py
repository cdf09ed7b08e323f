"""Compile-test of the reference-side forwarding TUs (dropin.cpp: the six integrators;
dropin_tau.cpp: tau_estimate / tau_estimate_linear).

Needs the reference's headers (/root/reference/src, used in place, never copied)
and -- only because this image lacks libgmp-dev -- the declaration shim
integration/shims/gmp.h. Output: qunundrum_b200/dropin/libqunundrum_dropin.so, which
exports the six C++ entry points of the reference with their mangled names and
depends on ../libqunundrum_b200.so and libgmp. It also compiles the reference's
src/errors.c (critical()), generator (src/random.c, src/keccak*.c) and slice enumerators
(src/*_enumerator.cpp, src/math.cpp) in place so that the test library is self-contained.
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
    deps = [os.path.join(HERE, "dropin.cpp"), os.path.join(HERE, "dropin_tau.cpp"),
            os.path.join(ROOT, "include", "qunundrum_b200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(
            os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gmp = "/lib/x86_64-linux-gnu/libgmp.so.10"
    # the reference's own errors.c (critical()) and generator (random_generate(), which
    # dropin_tau.cpp draws the caller's stream with), compiled where they lie
    objs = []
    for f in ("errors", "random", "keccak", "keccak_random", "debug_common"):
        o = os.path.join(HERE, f"_{f}.o")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", "-I", os.path.join(ROOT, "integration", "shims"),
                               "-iquote", src, "-c", os.path.join(src, f + ".c"), "-o", o])
        objs.append(o)
    # ... and its enumerators (the drop-in asks them which slices a generator will request)
    inc = ["-I", os.path.join(ROOT, "integration", "shims"), "-I", os.path.join(ROOT, "integration", "minimpi"),
           "-I", os.path.join(ROOT, "integration", "stubs"), "-iquote", src]
    for f in ("distribution_enumerator", "linear_distribution_enumerator", "diagonal_distribution_enumerator",
              "math"):
        o = os.path.join(HERE, f"_{f}.o")
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-fPIC", "-w", "-include", "cmath", *inc, "-c",
                               os.path.join(src, f + ".cpp"), "-o", o])
        objs.append(o)
    subprocess.check_call(
        ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w",
         "-I", os.path.join(ROOT, "integration", "shims"), "-I", os.path.join(ROOT, "integration", "minimpi"),
         "-I", os.path.join(ROOT, "integration", "stubs"), "-I", os.path.join(ROOT, "include"),
         "-iquote", src, os.path.join(HERE, "dropin.cpp"), os.path.join(HERE, "dropin_tau.cpp"), *objs,
         "-o", LIB, "-L", PKG, "-lqunundrum_b200", "-Wl,-rpath,$ORIGIN/..", gmp])
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
