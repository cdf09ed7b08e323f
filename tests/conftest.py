import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


class GoldenSlice:
    def __init__(self, meta, cells):
        self.meta = meta
        self.cells = cells  # np.longdouble
        self.total_probability = np.longdouble(meta["tp_hi"]) + np.longdouble(meta["tp_lo"])
        self.total_error = np.ldexp(np.longdouble(meta["te_mant"]), meta["te_exp"])
        self.flags = meta["flags"]
        self.d = int(meta["d"])
        self.r = int(meta["r"])

    def __repr__(self):
        return f"GoldenSlice({self.meta['name']})"


def load_golden():
    meta = json.load(open(os.path.join(GOLDEN, "slices_meta.json")))
    z = np.load(os.path.join(GOLDEN, "slices.npz"))
    out = []
    for m in meta:
        hi = z[m["name"] + "/cells_hi"].astype(np.longdouble)
        lo = z[m["name"] + "/cells_lo"].astype(np.longdouble)
        out.append(GoldenSlice(m, hi + lo))
    return out


_GOLDEN = None


def golden_slices():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = load_golden()
    return _GOLDEN


@pytest.fixture(scope="session")
def golden():
    return golden_slices()


def ref_or_none():
    """The compiled reference (oracle/_ref) if it has been built, else None."""
    from oracle import ref
    if not ref.available() and os.path.isdir("/root/reference/src"):
        try:
            ref.build()
        except Exception:
            return None
    return ref if ref.available() else None


@pytest.fixture(scope="session")
def refmod():
    return ref_or_none()


@pytest.fixture(scope="session")
def gpu_ctx():
    import qunundrum_b200 as qb
    ctx = qb.Context(0)
    yield ctx
    ctx.close()
