"""Comparison helpers shared by the parity tests.

Tolerances (BASELINE.json north_star): every slice cell within 1e-9 relative,
total captured probability mass within 1e-12, flags / coordinates exact.
A cell that the reference's own Richardson step cancels to (almost) nothing
(2 * fine - coarse with both terms ~1e6 times larger) gets an absolute floor of
1e-15 of the slice's largest cell.
"""
import numpy as np

CELL_RTOL = 1e-9
MASS_ATOL = 1e-12
ERROR_RTOL = 1e-9


def cell_errors(got, ref):
    ref = np.asarray(ref, dtype=np.longdouble)
    got = np.asarray(got, dtype=np.longdouble)
    floor = np.longdouble(1e-15) * np.max(np.abs(ref))
    err = np.abs(got - ref) / (np.abs(ref) + floor)
    return float(np.max(err))


def assert_slice_matches(got_cells, got_tp, got_te, got_flags, g, check_error=True):
    e = cell_errors(got_cells, g.cells)
    assert e <= CELL_RTOL, f"{g}: cell error {e:.3e}"
    dtp = abs(float(np.longdouble(got_tp) - g.total_probability))
    assert dtp <= MASS_ATOL, f"{g}: total probability off by {dtp:.3e}"
    if check_error:
        if g.total_error == 0:
            assert got_te == 0, f"{g}: expected zero total_error"
        else:
            rel = abs(float((np.longdouble(got_te) - g.total_error) / g.total_error))
            assert rel <= ERROR_RTOL, f"{g}: total_error rel {rel:.3e}"
    assert int(got_flags) == int(g.flags), f"{g}: flags {int(got_flags):x} != {int(g.flags):x}"


def group_by(slices, keys):
    groups = {}
    for g in slices:
        k = tuple(g.meta.get(x) for x in keys)
        groups.setdefault(k, []).append(g)
    return groups
