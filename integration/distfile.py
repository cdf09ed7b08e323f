"""Reader / comparer of the reference's distribution text format.

  2D       distribution_export           src/distribution.cpp:305-323
  linear   linear_distribution_export    src/linear_distribution.cpp:498-519
  diagonal diagonal_distribution_export  src/diagonal_distribution.cpp:230-255
  slices   *_slice_export                src/*_slice_import_export.cpp:89-103
  params   parameters_export / diagonal_parameters_export
"""
from __future__ import annotations

import numpy as np


class Dist:
    def __init__(self, kind):
        self.kind = kind
        self.header = {}
        self.slices = {}     # key (coordinates) -> dict(dimension, flags, cells, total_error)
        self.order = []      # keys in file order


def read(path: str, kind: str) -> Dist:
    tok = open(path).read().split()
    pos = 0

    def nxt():
        nonlocal pos
        pos += 1
        return tok[pos - 1]

    d = Dist(kind)
    h = d.header
    h["precision"] = int(nxt())
    if kind in ("linear", "diagonal"):
        h["flags"] = int(nxt(), 16)
    h["m"] = int(nxt())
    if kind == "diagonal":
        h["sigma"] = int(nxt())
    h["s"] = int(nxt())
    h["l"] = int(nxt())
    h["r"] = int(nxt())
    h["d"] = int(nxt())
    h["t"] = int(nxt())
    if kind == "diagonal":
        h["eta_bound"] = int(nxt())
        h["min_alpha_r"], h["max_alpha_r"] = int(nxt()), int(nxt())
    else:
        h["min_alpha_d"], h["max_alpha_d"] = int(nxt()), int(nxt())
        h["min_alpha_r"], h["max_alpha_r"] = int(nxt()), int(nxt())
    count = int(nxt())
    for _ in range(count):
        dim = int(nxt())
        if kind == "2d":
            key = (int(nxt()), int(nxt()))
            n = dim * dim
        elif kind == "linear":
            key = (int(nxt()),)
            n = dim
        else:
            key = (int(nxt()), int(nxt()))   # (min_log_alpha_r, eta)
            n = dim
        flags = int(nxt(), 16)
        cells = np.array(tok[pos:pos + n], dtype=np.longdouble)
        pos += n
        te = np.longdouble(nxt())
        if key in d.slices:
            raise ValueError(f"duplicate slice {key}")
        d.slices[key] = dict(dimension=dim, flags=flags, cells=cells, total_error=te)
        d.order.append(key)
    if pos != len(tok):
        raise ValueError("trailing tokens")
    return d


def compare(a: Dist, b: Dist, cell_rtol=1e-9, mass_atol=1e-12, error_rtol=1e-9):
    """a: ours, b: reference. Returns a report dict; raises AssertionError on a mismatch."""
    assert a.header == b.header, (a.header, b.header)
    assert set(a.slices) == set(b.slices), "slice selection differs"
    worst_cell = worst_mass = worst_err = 0.0
    for key, sb in b.slices.items():
        sa = a.slices[key]
        assert sa["dimension"] == sb["dimension"] and sa["flags"] == sb["flags"], key
        ref = sb["cells"]
        floor = np.longdouble(1e-15) * np.max(np.abs(ref))
        e = float(np.max(np.abs(sa["cells"] - ref) / (np.abs(ref) + floor)))
        worst_cell = max(worst_cell, e)
        assert e <= cell_rtol, (key, e)
        dm = abs(float(sa["cells"].sum() - ref.sum()))
        worst_mass = max(worst_mass, dm)
        assert dm <= mass_atol, (key, dm)
        if sb["total_error"] != 0:
            de = abs(float((sa["total_error"] - sb["total_error"]) / sb["total_error"]))
            worst_err = max(worst_err, de)
            assert de <= error_rtol, (key, de)
        else:
            assert sa["total_error"] == 0
    # ordering: both files are sorted by slice probability (qsort, unstable): the order may only
    # differ among slices whose totals agree to the comparison tolerance
    ta = [float(a.slices[k]["cells"].sum()) for k in a.order]
    assert all(x >= y - mass_atol for x, y in zip(ta, ta[1:])), "our file is not sorted by probability"
    moved = sum(1 for ka, kb in zip(a.order, b.order) if ka != kb)
    for ka, kb in zip(a.order, b.order):
        if ka != kb:
            assert abs(float(b.slices[ka]["cells"].sum() - b.slices[kb]["cells"].sum())) <= 2 * mass_atol, \
                (ka, kb)
    return dict(slices=len(b.slices), worst_cell=worst_cell, worst_mass=worst_mass,
                worst_error=worst_err, reordered_ties=moved,
                total_mass=float(sum(s["cells"].sum() for s in a.slices.values())))
