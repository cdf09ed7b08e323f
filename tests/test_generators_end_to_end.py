"""The reference's own, unmodified generator executables (argv parsing, MPI master-worker
protocol, enumerators, containers, sort, text export) with the six slice-integration TUs
replaced by the drop-in, against the same executables with the reference's integrators:
the .txt distribution files must agree slice for slice.

Binaries are built by integration/build.py where /root/reference is mounted and travel to the
GPU box; MPI is integration/minimpi (no MPI implementation is installed in this image)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "integration", "_build")


def _have():
    return os.path.exists(os.path.join(B, ".done"))


def _run(flavour, exe, args, np_, cwd, env=None):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(np_), os.path.join(B, flavour, exe), *args],
                       cwd=cwd, env=e, capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _files(cwd):
    d = os.path.join(cwd, "distributions")
    return sorted(f for f in os.listdir(d) if f.endswith(".txt"))


def test_minimpi_runs_the_reference_generator_on_cpu():
    """BASELINE config 0 shape (linear, m = 128, s = 2, MPI world_size = 2), reference integrators,
    small dimension so that it takes a second: protocol + export work under integration/minimpi."""
    if not _have():
        pytest.skip("integration/_build missing (needs /root/reference at build time)")
    from integration import distfile
    with tempfile.TemporaryDirectory() as t:
        out = _run("ref", "generate_linear_distribution", ["-d", "-dim", "64", "-det", "128", "2"], 2, t)
        assert "Stopping node 1" in out
        f = _files(t)
        assert f == ["linear-distribution-det-dim-64-d-m-128-s-2.txt"]
        d = distfile.read(os.path.join(t, "distributions", f[0]), "linear")
        assert d.header["m"] == 128 and d.header["l"] == 64 and len(d.slices) == 82
        mass = float(sum(s["cells"].sum() for s in d.slices.values()))
        assert abs(mass - 0.9999038946) < 1e-4    # docs/pages/info-linear-distribution.md (dimension 2048)


CASES = [
    # exe, args, kind, ranks  (ranks = 1 server + clients, all clients share cuda:0 here)
    ("generate_linear_distribution", ["-d", "-dim", "2048", "-det", "128", "2"], "linear", 2),
    ("generate_linear_distribution", ["-r", "-dim", "512", "-det", "128", "2"], "linear", 3),
    ("generate_diagonal_distribution", ["-dim", "512", "-det", "-eta-bound", "1", "128", "5", "2"], "diagonal", 3),
    ("generate_distribution", ["-det", "-dim", "16", "128", "2"], "2d", 3),
    # the other two slice methods of the two-dimensional generator (src/distribution_slice.h:31-78)
    ("generate_distribution", ["-det", "-approx-quick", "-dim", "16", "128", "2"], "2d", 3),
    ("generate_distribution", ["-det", "-sigma-optimal", "-dim", "16", "64", "2"], "2d", 3),
    # two distributions in one run, explicit l (parameters_explicit_m_l): the drop-in sees the
    # parameter set change between calls
    ("generate_distribution", ["-det", "-dim", "16", "-l", "64", "40", "96", "50"], "2d", 3),
    # BASELINE config 3 shape: Ekera-Hastad factoring, m = n / 2 - 1, l = m - 20, always target d
    ("generate_linear_distribution_rsa", ["-dim", "256", "-max", "256"], "linear", 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("exe,args,kind,ranks", CASES, ids=[c[0] + "-" + c[2] + str(i) for i, c in enumerate(CASES)])
def test_generator_with_dropin_matches_reference_generator(exe, args, kind, ranks):
    if not _have():
        pytest.skip("integration/_build missing")
    from integration import distfile
    ta, tb = tempfile.mkdtemp(), tempfile.mkdtemp()
    try:
        _run("gpu", exe, args, ranks, ta, env={"QB200_DEVICE": "0"})
        _run("ref", exe, args, 9, tb)
        fa, fb = _files(ta), _files(tb)
        assert fa == fb and len(fa) >= 1
        for f in fa:
            # collapsed-d / collapsed-r marginals of a 2D distribution are linear distributions;
            # filtered-* files hold the slices that survive the error filter (selection parity)
            k = "linear" if f.startswith("collapsed-") else kind
            a = distfile.read(os.path.join(ta, "distributions", f), k)
            b = distfile.read(os.path.join(tb, "distributions", f), k)
            rep = distfile.compare(a, b)
            print(f, rep)
            assert rep["slices"] > 10
    finally:
        shutil.rmtree(ta, ignore_errors=True)
        shutil.rmtree(tb, ignore_errors=True)


PREFETCH = [
    ("generate_distribution", ["-det", "128", "2"], 3),                       # dimension heuristic: upgrades
    ("generate_distribution", ["-det", "-dim", "32", "-l", "64", "40", "96", "50"], 2),   # two parameter sets
    ("generate_linear_distribution", ["-r", "-dim", "512", "-det", "128", "2"], 3),
    ("generate_diagonal_distribution", ["-dim", "256", "-det", "-eta-bound", "2", "128", "5", "2"], 4),
]


@pytest.mark.gpu
@pytest.mark.parametrize("exe,args,ranks", PREFETCH, ids=[c[0] + str(i) for i, c in enumerate(PREFETCH)])
def test_prefetching_dropin_is_invisible_to_the_generator(exe, args, ranks):
    """the drop-in integrates the whole enumerator list in one C-ABI call
    on the first request and serves the client's one-slice-at-a-time calls from that batch (the
    dimension upgrades of the heuristic speculatively); collapse and export run from a device copy.
    Every exported file must be byte-identical to a run with QB200_PREFETCH=0 (one slice per call),
    and the statistics must show that the calls were served from batches."""
    if not _have():
        pytest.skip("integration/_build missing")
    import filecmp
    ta, tb = tempfile.mkdtemp(), tempfile.mkdtemp()
    try:
        e = {"QB200_DEVICE": "0", "QB200_DROPIN_STATS": "1"}
        os.makedirs(os.path.join(ta, "distributions"), exist_ok=True)
        p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(ranks), os.path.join(B, "gpu", exe), *args],
                           cwd=ta, env=dict(os.environ, **e), capture_output=True, text=True, timeout=3000)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        stats = [l for l in p.stderr.splitlines() if "slices per call" in l]
        assert stats, p.stderr[-2000:]
        import re
        served = sum(int(re.search(r"(\d+) calls served from", l).group(1)) for l in stats)
        calls = sum(int(re.search(r"drop-in: (\d+) slice calls", l).group(1)) for l in stats)
        abi = sum(int(re.search(r"; (\d+) C-ABI calls", l).group(1)) for l in stats)
        assert served == calls and abi < calls / 4, stats
        _run("gpu", exe, args, ranks, tb, env={"QB200_DEVICE": "0", "QB200_PREFETCH": "0"})
        fa, fb = _files(ta), _files(tb)
        assert fa == fb and len(fa) >= 1
        for f in fa:
            assert filecmp.cmp(os.path.join(ta, "distributions", f), os.path.join(tb, "distributions", f),
                               shallow=False), f
    finally:
        shutil.rmtree(ta, ignore_errors=True)
        shutil.rmtree(tb, ignore_errors=True)


FULL = [
    # BASELINE config 4 shape: two-dimensional, m = 3072, s = 4 (l = 768, sigma = 391), Richardson
    ["--m", "3072", "--s", "4", "--dim", "64", "--sample", "16", "--clients", "2"],
    # config 1 / 2 shape at a size the reference finishes quickly, dimension heuristic (upgrades)
    ["--m", "256", "--s", "2", "--dim", "0", "--sample", "12", "--clients", "2"],
]


@pytest.mark.gpu
@pytest.mark.parametrize("args", FULL, ids=["m3072-s4-dim64", "m256-s2-heuristic"])
def test_full_distribution_sampled_against_reference(args):
    """A complete two-dimensional distribution from the reference's generator with both drop-ins;
    a random sample of its slices is re-derived with the reference itself (oracle/_ref) on the
    host cores, following the client's dimension heuristic where no -dim is given."""
    from oracle import ref
    if not _have() or not ref.available():
        pytest.skip("integration/_build or oracle/_ref missing")
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "full_distribution.py"), *args],
                       capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["linear", "rsa", "sweep", "diagonal"])
def test_one_dimensional_baseline_configs_at_full_size(config):
    """BASELINE configs 1, 3 (RSA-2048 and the s = 1..8 tradeoff sweep) and 5 (diagonal, m = 2048,
    sigma sweep) at their full sizes (dimension 2048) through the reference's generators with both
    drop-ins; sampled slices re-computed by the reference on the host cores (tests/tools/full_1d.py)."""
    from oracle import ref
    if not _have() or not ref.available():
        pytest.skip("integration/_build or oracle/_ref missing")
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "full_1d.py"), config, "8"],
                       capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
