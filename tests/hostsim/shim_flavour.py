"""TEST-ONLY "shim" flavour of the reference's importing executables.

filter_distribution, info_distribution and compare_*_distributions linked with the reference's
own integrators, qunundrum_b200/dropin/dropin_text.cpp and a CPU stand-in for the two text entry
points (tests/hostsim/abi_shim.cpp: the CPU compile of textfmt.cuh / textparse.cuh), so that the
HOST logic of the text drop-in (block reads, re-reads, seeking, fwrite) runs in the GPU-less suite
(tests/test_text_dropin_host_logic.py). Nothing here is part of, built by, or reachable from the
product or the integration build: it only reuses the object files integration/build.py left in
integration/_build/obj. Output: tests/hostsim/_build/shim/.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
OUT = os.path.join(_HERE, "_build", "shim")
TOOLS = ("filter_distribution", "compare_distributions", "compare_linear_distributions",
         "compare_diagonal_distributions", "info_distribution")


def build(force: bool = False) -> bool:
    from integration import build as ib
    obj = os.path.join(ib.OUT, "obj")
    if not os.path.exists(os.path.join(obj, "dropin_text.o")):
        return os.path.exists(os.path.join(OUT, TOOLS[0]))
    srcs = [os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp")]
    deps = srcs + [os.path.join(_ROOT, "qunundrum_b200", "csrc", f) for f in ("textfmt.cuh", "textparse.cuh")]
    deps += [os.path.join(obj, "dropin_text.o"), os.path.abspath(__file__)]
    last = os.path.join(OUT, TOOLS[-1])
    if not force and os.path.exists(last) and all(
            os.path.getmtime(d) <= os.path.getmtime(last) for d in deps):
        return True
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(OUT, "libqb200_textshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++", *srcs,
                           "-o", shim])
    common = [os.path.join(obj, f + ".o") for f in ib.COMMON_CPP + ib.COMMON_C + ["lattice_stub", "minimpi"]]
    libs = [os.path.join(ib.LIBDIR, "libmpfr.so.6"), os.path.join(ib.LIBDIR, "libgmp.so.10"),
            "-lpthread", "-lm"]
    for m in TOOLS:
        subprocess.check_call(["g++", os.path.join(obj, "main_" + m + ".o"), *common,
                               *[os.path.join(obj, f + ".o") for f in ib.INTEGRATORS],
                               os.path.join(obj, "dropin_text.o"), shim, "-Wl,-rpath,$ORIGIN", *libs,
                               "-o", os.path.join(OUT, m)])
    return True


if __name__ == "__main__":
    print(build(force=True))
