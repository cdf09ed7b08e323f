"""Host logic of the integrator / collapse / export drop-ins inside the reference's own generator
executables, on a GPU-less machine.

tests/hostsim/shim_flavour.build_generators() links the unmodified generator mains, the reference's
containers / enumerators / MPI protocol (integration/minimpi), qunundrum_b200/dropin/dropin.cpp,
dropin_text.cpp and dropin_collapse.cpp against tests/hostsim/abi_shim.cpp, a CPU stand-in of the
C ABI whose integrators return SYNTHETIC cells (a pure function of coordinate, dimension and cell
index) while collapse and text use the real arithmetic (CPU twins of client_math.cuh / textfmt.cuh).
What is checked is therefore the drop-ins' own logic, not the mathematics:

  * one-slice-at-a-time calls are served from batches of the whole enumerator list, and the output
    files are byte-identical to a run with QB200_PREFETCH=0;
  * the collapsed marginals written through dropin_collapse.cpp (resident distribution) are
    byte-identical to those of the reference's own linear_distribution_init_collapse_d / _r on the
    same slices (flavour gen_refcollapse), and so is the two-dimensional file exported from the
    resident copy.

The dimension-heuristic run (slices at 128 / 256 / 512, 4.7 GB of text, a minute per run on the CPU
twin) is exercised the same way on the GPU (tests/test_generators_end_to_end.py); run by hand on the
CPU twin it served 3912 calls from 7 C-ABI calls with identical files (profiles/README.md)."""
import filecmp
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "integration", "_build")

CASES = [
    ("generate_distribution", ["-det", "-dim", "32", "128", "2"], 3),
    ("generate_distribution", ["-det", "-approx-quick", "-dim", "16", "-l", "64", "40", "96", "50"], 2),
    ("generate_linear_distribution", ["-d", "-dim", "512", "-det", "128", "2"], 3),
    ("generate_diagonal_distribution", ["-dim", "128", "-det", "-eta-bound", "2", "128", "5", "2"], 4),
]


def _dirs():
    from tests.hostsim import shim_flavour
    return shim_flavour.build_generators()


def _run(exe_dir, exe, args, ranks, cwd, env=None):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(ranks), os.path.join(exe_dir, exe), *args],
                       cwd=cwd, env=dict(os.environ, **(env or {})), capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p


@pytest.mark.parametrize("exe,args,ranks", CASES, ids=[c[0] + str(i) for i, c in enumerate(CASES)])
def test_prefetch_collapse_and_resident_export_host_logic(exe, args, ranks):
    dirs = _dirs()
    if not dirs or not os.path.exists(os.path.join(B, "minimpirun")):
        pytest.skip("integration/_build missing (needs /root/reference at build time)")
    gen, gen_ref = dirs
    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb, tempfile.TemporaryDirectory() as tc:
        p = _run(gen, exe, args, ranks, ta, {"QB200_DROPIN_STATS": "1"})
        stats = [l for l in p.stderr.splitlines() if "slices per call" in l]
        assert stats, p.stderr[-2000:]
        served = sum(int(re.search(r"(\d+) calls served from", l).group(1)) for l in stats)
        calls = sum(int(re.search(r"drop-in: (\d+) slice calls", l).group(1)) for l in stats)
        abi = sum(int(re.search(r"; (\d+) C-ABI calls", l).group(1)) for l in stats)
        assert served == calls and abi <= 2 * (ranks - 1) * (2 if "-l" in args else 1), stats
        if exe == "generate_distribution":
            col = [l for l in p.stderr.splitlines() if "collapse drop-in" in l]
            assert col and re.search(r"serving (\d+) slice exports", col[0]).group(1) != "0", p.stderr[-2000:]
        _run(gen, exe, args, ranks, tb, {"QB200_PREFETCH": "0"})
        _run(gen_ref, exe, args, ranks, tc)
        files = sorted(f for f in os.listdir(os.path.join(ta, "distributions")) if f.endswith(".txt"))
        assert files
        if exe == "generate_distribution":
            assert any(f.startswith("collapsed-d-") for f in files) and any(f.startswith("collapsed-r-") for f in files)
        for f in files:
            for other in (tb, tc):
                assert filecmp.cmp(os.path.join(ta, "distributions", f), os.path.join(other, "distributions", f),
                                   shallow=False), (f, other)
