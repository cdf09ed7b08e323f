"""The reference's own executables with the TEXT drop-in (qunundrum_b200/dropin/dropin_text.cpp
in place of the three *_slice_import_export.cpp translation units), against the same
executables with the reference's exporters / importers:

* a file written through the GPU exporter is canonical: every number line is exactly what
  libc prints for the value libc reads from it;
* filter_distribution (import -> filter -> sort -> export, src/main_filter_distribution.cpp)
  produces byte-identical files in both flavours from the same input;
* compare_*_distributions (import x 2) print the same report in both flavours.

Binaries: integration/build.py (built where /root/reference is mounted; they travel to the GPU box)."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import text as ot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "integration", "_build")
pytestmark = pytest.mark.gpu


def _have():
    return os.path.exists(os.path.join(B, ".done")) and os.path.exists(os.path.join(B, "gpu", "filter_distribution"))


def _generate(exe, args, ranks, cwd):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    env = dict(os.environ, QB200_DEVICE="0")
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", str(ranks), os.path.join(B, "gpu", exe), *args],
                       cwd=cwd, env=env, capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = os.path.join(cwd, "distributions")
    return {f: os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".txt")}


def _tool(flavour, exe, args, cwd):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    env = dict(os.environ, QB200_DEVICE="0")
    p = subprocess.run([os.path.join(B, flavour, exe), *args], cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _number_lines(path):
    """The lines of a distribution file that hold '%.24Lg' numbers (contain '.', 'e', or are
    plain decimals of a cell) -- selected by position: everything the exporters print with
    %.24Lg has either a '.' or an exponent, or is a bare integer < 2^63 such as 0."""
    out = []
    for line in open(path, "rb"):
        s = line.strip()
        if b"." in s or b"e" in s or s in (b"inf", b"nan", b"-inf", b"-nan"):
            out.append(s)
    return out


def _assert_canonical(path):
    toks = _number_lines(path)
    assert len(toks) > 100
    text = b"\n".join(toks) + b"\n"
    vals = ot.parse_ld(text, len(toks))
    assert ot.format_ld24(vals) == text, path


@pytest.fixture(scope="module")
def workdir():
    if not _have():
        pytest.skip("integration/_build missing (needs /root/reference at build time)")
    t = tempfile.mkdtemp()
    yield t
    shutil.rmtree(t, ignore_errors=True)


def test_two_dimensional_files_roundtrip_through_both_flavours(workdir):
    files = _generate("generate_distribution", ["-det", "-dim", "64", "256", "2"], 3, workdir)
    main = next(p for f, p in files.items() if f.startswith("distribution-"))
    for p in files.values():
        _assert_canonical(p)
    ta, tb = os.path.join(workdir, "a"), os.path.join(workdir, "b")
    os.makedirs(ta), os.makedirs(tb)
    _tool("ref", "filter_distribution", [main], ta)
    _tool("gpu", "filter_distribution", [main], tb)
    name = "filtered-" + os.path.basename(main)
    fa = open(os.path.join(ta, "distributions", name), "rb").read()
    fb = open(os.path.join(tb, "distributions", name), "rb").read()
    assert len(fa) > 100000 and fa == fb, "GPU import + export differs from the reference's"
    # and the importers agree when comparing the original with its filtered copy
    ra = _tool("ref", "compare_distributions", [main, os.path.join(ta, "distributions", name)], ta)
    rb = _tool("gpu", "compare_distributions", [main, os.path.join(ta, "distributions", name)], tb)
    assert ra == rb and len(ra) > 0
    # the same file with 50 blanks in front of every line: the importer's first block is too
    # short for a slice and it reads on; both flavours must still agree byte for byte
    padded = os.path.join(workdir, "padded-" + os.path.basename(main))
    open(padded, "wb").write(open(main, "rb").read().replace(b"\n", b"\n" + b" " * 50))
    tc, td = os.path.join(workdir, "c"), os.path.join(workdir, "d")
    os.makedirs(tc), os.makedirs(td)
    _tool("ref", "filter_distribution", [padded], tc)
    _tool("gpu", "filter_distribution", [padded], td)
    name2 = "filtered-" + os.path.basename(padded)
    fc = open(os.path.join(tc, "distributions", name2), "rb").read()
    assert fc == open(os.path.join(td, "distributions", name2), "rb").read() and fc == fa
    ia = _tool("ref", "info_distribution", [main], ta)
    ib = _tool("gpu", "info_distribution", [main], tb)
    assert ia == ib and "Total probability" in ia


def test_linear_and_diagonal_files_are_canonical_and_import_identically(workdir):
    t = os.path.join(workdir, "lin")
    files = _generate("generate_linear_distribution", ["-d", "-dim", "2048", "-det", "128", "2"], 2, t)
    files.update(_generate("generate_linear_distribution", ["-r", "-dim", "512", "-det", "128", "2"], 2, t))
    lin = list(files.values())
    for p in lin:
        _assert_canonical(p)
    ra = _tool("ref", "compare_linear_distributions", lin[:2], t)
    rb = _tool("gpu", "compare_linear_distributions", lin[:2], t)
    assert ra == rb and len(ra) > 0
    t = os.path.join(workdir, "diag")
    files = _generate("generate_diagonal_distribution",
                      ["-dim", "512", "-det", "-eta-bound", "1", "128", "5", "2"], 3, t)
    diag = list(files.values())
    for p in diag:
        _assert_canonical(p)
    ra = _tool("ref", "compare_diagonal_distributions", [diag[0], diag[0]], t)
    rb = _tool("gpu", "compare_diagonal_distributions", [diag[0], diag[0]], t)
    assert ra == rb and len(ra) > 0
