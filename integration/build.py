"""Build the reference's own generator executables, unmodified, in two flavours:

  integration/_build/ref/generate_*   the reference's slice integrators (MPFR, CPU)
  integration/_build/gpu/generate_*   the six integrator TUs replaced by
                                      qunundrum_b200/dropin/dropin.cpp, the three
                                      *_slice_import_export TUs by dropin_text.cpp, the two
                                      collapse functions of linear_distribution.cpp by
                                      dropin_collapse.cpp, + libqunundrum_b200.so

plus, in both flavours, the importing executables filter_distribution, info_distribution and
compare_[linear_|diagonal_]distributions (they load stored distributions: the importer path), and
estimate_runs_distribution / estimate_runs_linear_distribution (gpu flavour: tau_estimate and
tau_estimate_linear from qunundrum_b200/dropin/dropin_tau.cpp; the reference's tau_estimate.cpp
stays in the build compiled with three -D renames; tau_estimate_diagonal from
qunundrum_b200/dropin/dropin_tau_diagonal.cpp).

INTEGRATION-TEST INFRASTRUCTURE. Sources are compiled where they lie under /root/reference/src
(never copied); neither OpenMPI nor fpLLL nor the GMP/MPFR development headers exist in this
image, so the build uses integration/minimpi (a minimal single-node MPI over socket pairs),
integration/stubs (a compile-only fpLLL stand-in: generation never reduces a lattice) and the
declaration shims of integration/shims. With the real toolchain a maintainer follows INTEGRATION.md
instead. The built binaries travel to the GPU box; /root/reference is not needed at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
LIBDIR = "/lib/x86_64-linux-gnu"

COMMON_CPP = """math rsa parameters diagonal_parameters parameters_selection sample
 probability linear_probability diagonal_probability
 distribution distribution_enumerator distribution_info distribution_mpi distribution_slice
 distribution_slice_mpi
 linear_distribution linear_distribution_enumerator linear_distribution_info linear_distribution_mpi
 linear_distribution_slice linear_distribution_slice_mpi
 diagonal_distribution diagonal_distribution_enumerator diagonal_distribution_info
 diagonal_distribution_mpi diagonal_distribution_slice diagonal_distribution_slice_mpi
 distribution_loader linear_distribution_loader diagonal_distribution_loader tau_ordered_list
 tau_volume_quotient log""".split()
TEXT_IO = """distribution_slice_import_export linear_distribution_slice_import_export
 diagonal_distribution_slice_import_export""".split()
COMMON_C = "errors random keccak keccak_random gmp_mpi mpfr_mpi string_utilities thread_pool debug_common".split()
INTEGRATORS = """distribution_slice_compute distribution_slice_compute_richardson
 linear_distribution_slice_compute linear_distribution_slice_compute_richardson
 diagonal_distribution_slice_compute diagonal_distribution_slice_compute_richardson""".split()
# tau_estimate.cpp: as it is in the "ref" flavour; in the "gpu" flavour compiled with three renames
# next to qunundrum_b200/dropin/dropin_tau.cpp and dropin_tau_diagonal.cpp
TAU = ["tau_estimate"]
# linear_distribution.cpp: the two collapse functions renamed away in the "gpu" flavour, where
# qunundrum_b200/dropin/dropin_collapse.cpp defines them
COLLAPSE_RENAMES = ["-Dlinear_distribution_init_collapse_d=linear_distribution_init_collapse_d_cpu_unused",
                    "-Dlinear_distribution_init_collapse_r=linear_distribution_init_collapse_r_cpu_unused"]
TAU_RENAMES = ["-Dtau_estimate=tau_estimate_cpu_unused", "-Dtau_estimate_linear=tau_estimate_linear_cpu_unused",
               "-Dtau_estimate_diagonal=tau_estimate_diagonal_cpu_unused"]
ESTIMATORS = ["estimate_runs_distribution", "estimate_runs_linear_distribution",
              "estimate_runs_diagonal_distribution"]
MAINS = ["generate_distribution", "generate_linear_distribution", "generate_linear_distribution_rsa",
         "generate_diagonal_distribution", "filter_distribution", "info_distribution",
         "compare_distributions", "compare_linear_distributions", "compare_diagonal_distributions"]


def build(reference_root: str = "/root/reference", force: bool = False) -> bool:
    src = os.path.join(reference_root, "src")
    done = os.path.join(OUT, ".done")
    if not os.path.isdir(src):
        return os.path.exists(done)
    deps = [os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin.cpp"),
            os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_text.cpp"),
            os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_tau.cpp"),
            os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_tau_diagonal.cpp"),
            os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_collapse.cpp"),
            os.path.join(HERE, "tools", "tau_diagonal_check.cpp"),
            os.path.join(HERE, "minimpi", "minimpi.c"), os.path.join(HERE, "minimpi", "mpi.h"),
            os.path.join(HERE, "build.py"), os.path.join(ROOT, "include", "qunundrum_b200.h"),
            ]
    if not force and os.path.exists(done) and all(
            os.path.getmtime(d) <= os.path.getmtime(done) for d in deps):
        return True
    obj = os.path.join(OUT, "obj")
    for d in (obj, os.path.join(OUT, "ref"), os.path.join(OUT, "gpu")):
        os.makedirs(d, exist_ok=True)
    inc = ["-I", os.path.join(HERE, "minimpi"), "-I", os.path.join(HERE, "stubs"),
           "-I", os.path.join(HERE, "shims"), "-I", os.path.join(ROOT, "include"),
           "-iquote", src]
    jobs = []
    for f in COMMON_CPP + INTEGRATORS + TEXT_IO + TAU + ["main_" + m for m in MAINS + ESTIMATORS]:
        jobs.append(["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *inc, "-c",
                     os.path.join(src, f + ".cpp"), "-o", os.path.join(obj, f + ".o")])
    for f in COMMON_C:
        jobs.append(["gcc", "-O2", "-w", *inc, "-c", os.path.join(src, f + ".c"),
                     "-o", os.path.join(obj, f + ".o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", *inc, "-c",
                 os.path.join(HERE, "stubs", "lattice_stub.cpp"), "-o", os.path.join(obj, "lattice_stub.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", *inc, "-c",
                 os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin.cpp"),
                 "-o", os.path.join(obj, "dropin.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", *inc, "-c",
                 os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_text.cpp"),
                 "-o", os.path.join(obj, "dropin_text.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", *inc, "-c",
                 os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_collapse.cpp"),
                 "-o", os.path.join(obj, "dropin_collapse.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *COLLAPSE_RENAMES, *inc, "-c",
                 os.path.join(src, "linear_distribution.cpp"),
                 "-o", os.path.join(obj, "linear_distribution_renamed.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *TAU_RENAMES, *inc, "-c",
                 os.path.join(src, "tau_estimate.cpp"), "-o", os.path.join(obj, "tau_estimate_renamed.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", *inc, "-c",
                 os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_tau.cpp"),
                 "-o", os.path.join(obj, "dropin_tau.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *inc, "-c",
                 os.path.join(ROOT, "qunundrum_b200", "dropin", "dropin_tau_diagonal.cpp"),
                 "-o", os.path.join(obj, "dropin_tau_diagonal.o")])
    jobs.append(["g++", "-std=c++11", "-O2", "-w", "-include", "cmath", *inc, "-c",
                 os.path.join(HERE, "tools", "tau_diagonal_check.cpp"),
                 "-o", os.path.join(obj, "tau_diagonal_check.o")])
    jobs.append(["gcc", "-O2", "-c", os.path.join(HERE, "minimpi", "minimpi.c"),
                 "-o", os.path.join(obj, "minimpi.o")])
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(subprocess.check_call, jobs))
    subprocess.check_call(["gcc", "-O2", os.path.join(HERE, "minimpi", "minimpirun.c"),
                           "-o", os.path.join(OUT, "minimpirun")])
    common = [os.path.join(obj, f + ".o") for f in COMMON_CPP + COMMON_C + ["lattice_stub", "minimpi"]]
    common_gpu = [o for o in common if not o.endswith(os.sep + "linear_distribution.o")] + [
        os.path.join(obj, "linear_distribution_renamed.o"), os.path.join(obj, "dropin_collapse.o")]
    libs = [os.path.join(LIBDIR, "libmpfr.so.6"), os.path.join(LIBDIR, "libgmp.so.10"), "-lpthread", "-lm"]
    for m in MAINS:
        main_o = os.path.join(obj, "main_" + m + ".o")
        subprocess.check_call(["g++", main_o, *common,
                               *[os.path.join(obj, f + ".o") for f in INTEGRATORS + TEXT_IO],
                               *libs, "-o", os.path.join(OUT, "ref", m)])
        subprocess.check_call(["g++", main_o, *common_gpu, os.path.join(obj, "dropin.o"),
                               os.path.join(obj, "dropin_text.o"),
                               "-L", os.path.join(ROOT, "qunundrum_b200"), "-lqunundrum_b200",
                               "-Wl,-rpath,$ORIGIN/../../../qunundrum_b200", *libs,
                               "-o", os.path.join(OUT, "gpu", m)])
    # estimate_runs_*: the reference's tau estimators (ref) against dropin_tau.cpp (gpu); the
    # integrators and the text format are the same drop-ins as above in the gpu flavour
    for m in ESTIMATORS:
        main_o = os.path.join(obj, "main_" + m + ".o")
        subprocess.check_call(["g++", main_o, *common, os.path.join(obj, "tau_estimate.o"),
                               *[os.path.join(obj, f + ".o") for f in INTEGRATORS + TEXT_IO],
                               *libs, "-o", os.path.join(OUT, "ref", m)])
        subprocess.check_call(["g++", main_o, *common_gpu, os.path.join(obj, "tau_estimate_renamed.o"),
                               os.path.join(obj, "dropin_tau.o"), os.path.join(obj, "dropin_tau_diagonal.o"),
                               os.path.join(obj, "dropin.o"), os.path.join(obj, "dropin_text.o"),
                               "-L", os.path.join(ROOT, "qunundrum_b200"), "-lqunundrum_b200",
                               "-Wl,-rpath,$ORIGIN/../../../qunundrum_b200", *libs,
                               "-o", os.path.join(OUT, "gpu", m)])
    # test driver: the drop-in tau_estimate_diagonal against the reference's (renamed) in one process
    subprocess.check_call(["g++", os.path.join(obj, "tau_diagonal_check.o"), *common_gpu,
                           os.path.join(obj, "tau_estimate_renamed.o"), os.path.join(obj, "dropin_tau.o"),
                           os.path.join(obj, "dropin_tau_diagonal.o"), os.path.join(obj, "dropin.o"),
                           os.path.join(obj, "dropin_text.o"),
                           "-L", os.path.join(ROOT, "qunundrum_b200"), "-lqunundrum_b200",
                           "-Wl,-rpath,$ORIGIN/../../../qunundrum_b200", *libs,
                           "-o", os.path.join(OUT, "gpu", "tau_diagonal_check")])
    open(done, "w").write("ok\n")
    return True


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
