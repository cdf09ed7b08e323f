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


DROPIN_TAU_SHIM = os.path.join(OUT, "libdropin_tau_shim.so")


def build_tau(reference_root: str = "/root/reference", force: bool = False):
    """TEST-ONLY: qunundrum_b200/dropin/dropin_tau.cpp linked against the CPU stand-in of the sampler
    entry points (abi_shim.cpp) and the reference's own generator / errors sources compiled in
    place, so that the drop-in's host logic and csrc/sampler_host.hpp run without a GPU
    (tests/test_dropin_gpu.py::test_tau_dropin_host_logic_on_the_cpu_shim)."""
    src = os.path.join(reference_root, "src")
    if not os.path.isdir(src):
        return DROPIN_TAU_SHIM if os.path.exists(DROPIN_TAU_SHIM) else None
    deps = [os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"), os.path.abspath(__file__),
            os.path.join(_ROOT, "qunundrum_b200", "dropin", "dropin_tau.cpp")]
    deps += [os.path.join(_ROOT, "qunundrum_b200", "csrc", f) for f in
             ("sampler.cuh", "x87soft.cuh", "sampler_host.hpp")]
    if not force and os.path.exists(DROPIN_TAU_SHIM) and all(
            os.path.getmtime(d) <= os.path.getmtime(DROPIN_TAU_SHIM) for d in deps):
        return DROPIN_TAU_SHIM
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(_ROOT, "integration", "shims"), "-I", os.path.join(_ROOT, "integration", "minimpi"),
           "-I", os.path.join(_ROOT, "integration", "stubs"), "-I", os.path.join(_ROOT, "include"), "-iquote", src]
    objs = []
    for f in ("errors", "random", "keccak", "keccak_random", "debug_common"):
        o = os.path.join(OUT, f"_{f}.o")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", *inc, "-c", os.path.join(src, f + ".c"), "-o", o])
        objs.append(o)
    o = os.path.join(OUT, "_dropin_tau.o")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-fPIC", "-w", *inc, "-c",
                           os.path.join(_ROOT, "qunundrum_b200", "dropin", "dropin_tau.cpp"), "-o", o])
    objs.append(o)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"),
                           os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
                           os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp"),
                           "-x", "none", *objs, "/lib/x86_64-linux-gnu/libgmp.so.10", "-o", DROPIN_TAU_SHIM])
    for o in objs:
        os.remove(o)
    return DROPIN_TAU_SHIM


TAU_DIAGONAL_CHECK = os.path.join(OUT, "tau_diagonal_check")


def build_tau_diagonal(force: bool = False):
    """TEST-ONLY: integration/tools/tau_diagonal_check.cpp -- the drop-in tau_estimate_diagonal of
    qunundrum_b200/dropin/dropin_tau_diagonal.cpp next to the reference's own in one process --
    linked against the CPU stand-in of the qb200_diagk_* entry points (abi_shim.cpp over the CPU
    compile of csrc/diagk.cuh) and the reference's own text importers, from the object files
    integration/build.py left behind. Returns the executable's path, or None."""
    from integration import build as ib
    obj = os.path.join(ib.OUT, "obj")
    need = [os.path.join(obj, f) for f in ("tau_diagonal_check.o", "dropin_tau_diagonal.o",
                                           "tau_estimate_renamed.o")]
    if not all(os.path.exists(f) for f in need):
        return TAU_DIAGONAL_CHECK if os.path.exists(TAU_DIAGONAL_CHECK) else None
    srcs = [os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp")]
    deps = srcs + need + [os.path.abspath(__file__)]
    deps += [os.path.join(_ROOT, "qunundrum_b200", "csrc", f) for f in ("diagk.cuh", "diagk_host.hpp")]
    if not force and os.path.exists(TAU_DIAGONAL_CHECK) and all(
            os.path.getmtime(d) <= os.path.getmtime(TAU_DIAGONAL_CHECK) for d in deps):
        return TAU_DIAGONAL_CHECK
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(OUT, "libqb200_diagkshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++", *srcs,
                           "-o", shim])
    common = [os.path.join(obj, f + ".o") for f in ib.COMMON_CPP + ib.COMMON_C + ["lattice_stub", "minimpi"]]
    libs = [os.path.join(ib.LIBDIR, "libmpfr.so.6"), os.path.join(ib.LIBDIR, "libgmp.so.10"),
            "-lpthread", "-lm"]
    subprocess.check_call(["g++", *need, *common,
                           *[os.path.join(obj, f + ".o") for f in ib.INTEGRATORS + ib.TEXT_IO],
                           shim, "-Wl,-rpath,$ORIGIN", *libs, "-o", TAU_DIAGONAL_CHECK])
    return TAU_DIAGONAL_CHECK


SAMPLE_K_KAT_CHECK = os.path.join(OUT, "sample_k_kat_check")


def build_sample_k_kat(reference_root: str = "/root/reference", force: bool = False, flavour: str = "dropin"):
    """TEST-ONLY: integration/tools/sample_k_kat_check.cpp -- the reference's own known-answer test
    function of the diagonal k sampler (src/test/test_sample.cpp, compiled in place) linked against
    the drop-in's sample_k_from_diagonal_j_eta_pivot (dropin_tau_diagonal.cpp with
    -DQB200_DROPIN_SAMPLE_K), the reference's sample.cpp with that one function renamed, and the CPU
    stand-in of qb200_diagk_*. flavour = "reference": the same driver over the unmodified sample.cpp
    (to show that the vectors and the driver's comparison are sound). Returns the executable's path,
    or None."""
    if flavour == "reference":
        return _build_sample_k_kat_reference(reference_root, force)
    from integration import build as ib
    src = os.path.join(reference_root, "src")
    obj = os.path.join(ib.OUT, "obj")
    if not os.path.isdir(src) or not os.path.exists(os.path.join(obj, "math.o")):
        return SAMPLE_K_KAT_CHECK if os.path.exists(SAMPLE_K_KAT_CHECK) else None
    dropin = os.path.join(_ROOT, "qunundrum_b200", "dropin", "dropin_tau_diagonal.cpp")
    driver = os.path.join(_ROOT, "integration", "tools", "sample_k_kat_check.cpp")
    srcs = [os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp")]
    deps = srcs + [dropin, driver, os.path.abspath(__file__)]
    deps += [os.path.join(_ROOT, "qunundrum_b200", "csrc", f) for f in ("diagk.cuh", "diagk_host.hpp")]
    if not force and os.path.exists(SAMPLE_K_KAT_CHECK) and all(
            os.path.getmtime(d) <= os.path.getmtime(SAMPLE_K_KAT_CHECK) for d in deps):
        return SAMPLE_K_KAT_CHECK
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(_ROOT, "integration", "minimpi"), "-I", os.path.join(_ROOT, "integration", "stubs"),
           "-I", os.path.join(_ROOT, "integration", "shims"), "-I", os.path.join(_ROOT, "include"), "-iquote", src]
    cxx = ["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *inc, "-c"]
    tmp = []

    def cc(source, name, *flags):
        o = os.path.join(OUT, name)
        subprocess.check_call([*cxx, *flags, source, "-o", o])
        tmp.append(o)
        return o

    objs = [cc(driver, "_kat_driver.o"),
            cc(os.path.join(src, "test", "test_sample.cpp"), "_kat_test_sample.o"),
            cc(os.path.join(src, "test", "test_common.cpp"), "_kat_test_common.o",
               "-Dtest_cmp_ld=test_cmp_ld_of_the_reference"),
            cc(os.path.join(src, "test", "test_diagonal_probability.cpp"), "_kat_test_diagonal_probability.o"),
            cc(os.path.join(src, "sample.cpp"), "_kat_sample_renamed.o",
               "-Dsample_k_from_diagonal_j_eta_pivot=sample_k_from_diagonal_j_eta_pivot_cpu_unused"),
            cc(os.path.join(src, "diagonal_probability.cpp"), "_kat_diagonal_probability_renamed.o",
               "-Ddiagonal_probability_approx_h=diagonal_probability_approx_h_cpu_unused"),
            cc(dropin, "_kat_dropin.o", "-DQB200_DROPIN_SAMPLE_K")]
    shim = os.path.join(OUT, "libqb200_diagkshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++", *srcs,
                           "-o", shim])
    common = [os.path.join(obj, f + ".o") for f in ib.COMMON_CPP + ib.COMMON_C + ["lattice_stub", "minimpi"]
              if f not in ("sample", "diagonal_probability")]
    libs = [os.path.join(ib.LIBDIR, "libmpfr.so.6"), os.path.join(ib.LIBDIR, "libgmp.so.10"),
            "-lpthread", "-lm"]
    subprocess.check_call(["g++", *objs, *common,
                           *[os.path.join(obj, f + ".o") for f in ib.INTEGRATORS + ib.TEXT_IO],
                           shim, "-Wl,-rpath,$ORIGIN", *libs, "-o", SAMPLE_K_KAT_CHECK])
    for o in tmp:
        os.remove(o)
    return SAMPLE_K_KAT_CHECK


def _build_sample_k_kat_reference(reference_root, force):
    from integration import build as ib
    src = os.path.join(reference_root, "src")
    obj = os.path.join(ib.OUT, "obj")
    exe = SAMPLE_K_KAT_CHECK + "_reference"
    if not os.path.isdir(src) or not os.path.exists(os.path.join(obj, "sample.o")):
        return exe if os.path.exists(exe) else None
    driver = os.path.join(_ROOT, "integration", "tools", "sample_k_kat_check.cpp")
    if not force and os.path.exists(exe) and os.path.getmtime(driver) <= os.path.getmtime(exe):
        return exe
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(_ROOT, "integration", "minimpi"), "-I", os.path.join(_ROOT, "integration", "stubs"),
           "-I", os.path.join(_ROOT, "integration", "shims"), "-I", os.path.join(_ROOT, "include"), "-iquote", src]
    cxx = ["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *inc, "-c"]
    objs = []
    for source, name, flags in ((driver, "_katr_driver.o", []),
                                (os.path.join(src, "test", "test_sample.cpp"), "_katr_test_sample.o", []),
                                (os.path.join(src, "test", "test_diagonal_probability.cpp"),
                                 "_katr_test_diagonal_probability.o", []),
                                (os.path.join(src, "test", "test_common.cpp"), "_katr_test_common.o",
                                 ["-Dtest_cmp_ld=test_cmp_ld_of_the_reference"])):
        o = os.path.join(OUT, name)
        subprocess.check_call([*cxx, *flags, source, "-o", o])
        objs.append(o)
    common = [os.path.join(obj, f + ".o") for f in ib.COMMON_CPP + ib.COMMON_C + ["lattice_stub", "minimpi"]]
    libs = [os.path.join(ib.LIBDIR, "libmpfr.so.6"), os.path.join(ib.LIBDIR, "libgmp.so.10"),
            "-lpthread", "-lm"]
    subprocess.check_call(["g++", *objs, *common,
                           *[os.path.join(obj, f + ".o") for f in ib.INTEGRATORS + ib.TEXT_IO],
                           *libs, "-o", exe])
    for o in objs:
        os.remove(o)
    return exe


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


GENERATORS = ("generate_distribution", "generate_linear_distribution", "generate_diagonal_distribution")


def build_generators(force: bool = False):
    """TEST-ONLY: the reference's own generator executables linked with dropin.cpp, dropin_text.cpp and
    -- flavour "shim" -- dropin_collapse.cpp over the CPU stand-in of the C ABI (abi_shim.cpp: SYNTHETIC
    cells, real collapse / text arithmetic), so that the host logic of the three drop-ins (prefetching
    the enumerator list, speculating on the dimension upgrades, the resident collapse and export) runs
    inside the real master-worker protocol in the GPU-less suite. Flavour "shim_refcollapse": the same
    with the reference's own collapse functions, the checker for the collapsed files.
    Returns the two directories, or None."""
    from integration import build as ib
    obj = os.path.join(ib.OUT, "obj")
    need = [os.path.join(obj, f) for f in ("dropin.o", "dropin_text.o", "dropin_collapse.o",
                                           "linear_distribution_renamed.o", "linear_distribution.o")]
    outs = [os.path.join(OUT, "gen"), os.path.join(OUT, "gen_refcollapse")]
    last = os.path.join(outs[1], GENERATORS[-1])
    if not all(os.path.exists(f) for f in need):
        return outs if os.path.exists(last) else None
    srcs = [os.path.join(_HERE, "abi_shim.cpp"), os.path.join(_HERE, "hostsim.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "hostconst.cpp"),
            os.path.join(_ROOT, "qunundrum_b200", "csrc", "text_tables.cpp")]
    deps = srcs + need + [os.path.abspath(__file__),
                          os.path.join(_ROOT, "qunundrum_b200", "csrc", "client_math.cuh")]
    if not force and os.path.exists(last) and all(os.path.getmtime(d) <= os.path.getmtime(last) for d in deps):
        return outs
    for o in outs:
        os.makedirs(o, exist_ok=True)
    shim = os.path.join(OUT, "libqb200_genshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-x", "c++", *srcs,
                           "-o", shim])
    common = [os.path.join(obj, f + ".o") for f in ib.COMMON_CPP + ib.COMMON_C + ["lattice_stub", "minimpi"]]
    common_gpu = [o for o in common if not o.endswith(os.sep + "linear_distribution.o")] + [
        os.path.join(obj, "linear_distribution_renamed.o"), os.path.join(obj, "dropin_collapse.o")]
    libs = [os.path.join(ib.LIBDIR, "libmpfr.so.6"), os.path.join(ib.LIBDIR, "libgmp.so.10"),
            "-lpthread", "-lm"]
    for m in GENERATORS:
        for out, objs in ((outs[0], common_gpu), (outs[1], common)):
            subprocess.check_call(["g++", os.path.join(obj, "main_" + m + ".o"), *objs,
                                   os.path.join(obj, "dropin.o"), os.path.join(obj, "dropin_text.o"), shim,
                                   "-Wl,-rpath,$ORIGIN/..", *libs, "-o", os.path.join(out, m)])
    return outs


if __name__ == "__main__":
    print(build(force=True))
