"""Host logic of the reference-side text translation unit (qunundrum_b200/dropin/dropin_text.cpp)
WITHOUT a GPU: block reads, re-reads when a block is too short or ends inside a number, seeking
to where fscanf would have stopped, running sums, fwrite.

The reference's own importing executables are linked with dropin_text.cpp and a TEST-ONLY CPU
stand-in for the two text entry points (tests/hostsim/abi_shim.cpp: the CPU compile of
textfmt.cuh / textparse.cuh) -- tests/hostsim/shim_flavour.py -- and must produce
the same bytes as the same executables with the reference's own *_slice_import_export.cpp."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "integration", "_build")
SHIM = os.path.join(ROOT, "tests", "hostsim", "_build", "shim")


def _have():
    return os.path.exists(os.path.join(SHIM, "filter_distribution"))


def _tool(flavour, exe, args, cwd):
    os.makedirs(os.path.join(cwd, "distributions"), exist_ok=True)
    p = subprocess.run([os.path.join(SHIM if flavour == "shim" else os.path.join(B, flavour), exe), *args], cwd=cwd, capture_output=True, text=True,
                       timeout=3000)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.fixture(scope="module")
def generated():
    if not _have():
        pytest.skip("tests/hostsim/_build/shim missing (needs /root/reference at build time)")
    t = tempfile.mkdtemp()
    os.makedirs(os.path.join(t, "distributions"))
    # the reference's integrators on the CPU: small enough for a few seconds
    p = subprocess.run([os.path.join(B, "minimpirun"), "-np", "9", os.path.join(B, "ref", "generate_distribution"),
                        "-det", "-approx-quick", "-dim", "16", "64", "2"], cwd=t, capture_output=True, text=True,
                       timeout=3000)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = os.path.join(t, "distributions")
    main = next(os.path.join(d, f) for f in os.listdir(d) if f.startswith("distribution-") and f.endswith(".txt"))
    lin = sorted(os.path.join(d, f) for f in os.listdir(d) if f.startswith("collapsed-") and f.endswith(".txt"))
    yield t, main, lin
    shutil.rmtree(t, ignore_errors=True)


@pytest.mark.parametrize("pad", [0, 7, 9, 10, 11, 50])
def test_filter_distribution_is_byte_identical_with_the_text_dropin(generated, pad):
    """import -> filter -> sort -> export (src/main_filter_distribution.cpp). pad = blanks in front
    of every line: around 9-11 the importer's first block (40 bytes per number) ends inside or
    right behind the last number of a slice; at 50 it is far too short."""
    t, main, _ = generated
    src = main
    if pad:
        src = os.path.join(t, f"pad{pad}-" + os.path.basename(main))
        open(src, "wb").write(open(main, "rb").read().replace(b"\n", b"\n" + b" " * pad))
    outs = []
    for flavour in ("ref", "shim"):
        w = os.path.join(t, f"{flavour}-{pad}")
        _tool(flavour, "filter_distribution", [src], w)
        outs.append(open(os.path.join(w, "distributions", "filtered-" + os.path.basename(src)), "rb").read())
    assert len(outs[0]) > 100000 and outs[0] == outs[1]


def test_compare_and_info_agree(generated):
    t, main, lin = generated
    for exe, args in (("compare_distributions", [main, main]), ("info_distribution", [main]),
                      ("compare_linear_distributions", lin[:2])):
        a = _tool("ref", exe, args, os.path.join(t, "ca"))
        b = _tool("shim", exe, args, os.path.join(t, "cb"))
        assert a == b and len(a) > 0, exe
