"""Randomised differential testing against the reference itself (oracle/_ref): random bit sizes
m in [32, 3072], tradeoff factors s in [1, 80], arbitrary d < r, random coordinates over the whole
admissible region (|alpha| up to 2^(m+29)), random small dimensions INCLUDING non-powers of two
(the reference's grid step is the double 1.0 / dimension, reproduced in hostconst.cpp), all
three 2D methods, single pass and Richardson, linear d / r and diagonal with eta in [-25, 25].

CPU: the kernels' mathematics through tests/hostsim.  GPU: the CUDA path through the C ABI."""
import math
import random

import numpy as np
import pytest

from tests.conftest import ref_or_none
from tests.util import CELL_RTOL, cell_errors

REF = ref_or_none()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built")


def cases(seed, n):
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        m = rnd.choice([32, 48, 64, 100, 128, 200, 256, 513, 1024, 2048, 3072])
        s = rnd.choice([1, 1, 2, 3, 4, 5, 8, 10, 20, 30, 50, 80])
        l = math.ceil(m / s)
        if l < 8:
            continue
        r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
        d = r // 2 + rnd.randrange(r // 2)
        if rnd.random() < 0.2:
            d = 1 + rnd.randrange(r - 1)
        D = rnd.choice([2, 3, 4, 5, 8, 12])
        kind = rnd.choice(["2d", "2d", "2d", "lin_d", "lin_r", "diag"])
        lo, hi = max(1, m - 30), min(m + 29, m + l - 2)
        c = dict(m=m, s=s, l=l, d=d, r=r, D=D, kind=kind)
        if kind == "2d":
            c["method"] = rnd.choice([0, 0, 2, 1]) if l >= 12 else rnd.choice([0, 2])
            if c["method"] == 0 and REF.heuristic_sigma(l) > l:
                c["method"] = 2   # the reference returns NaN cells for sigma > l (unsigned l - sigma)
            c["a_d"] = rnd.randint(lo, hi) * rnd.choice([1, -1])
            c["a_r"] = rnd.randint(lo, hi) * rnd.choice([1, 1, -1])
            c["rich"] = rnd.choice([1, 1, 0])
        elif kind in ("lin_d", "lin_r"):
            if kind == "lin_d" and m >= 2048:
                c["D"] = min(D, 4)       # 6144-bit MPFR in the reference
            c["a"] = rnd.randint(lo, hi) * rnd.choice([1, -1])
        else:
            c["sigma"] = rnd.choice([0, 1, 2, 5, 8, 12, 20])
            c["eta"] = rnd.choice([0, 0, 1, -1, 2, -3, 25, -25])
            hi2 = (m + c["sigma"] - 2) if 30 >= c["sigma"] else m + 29
            if hi2 < lo:
                continue
            if m >= 2048:
                c["D"] = min(D, 4)
            c["a"] = rnd.randint(lo, hi2) * rnd.choice([1, -1])
        out.append(c)
    return out


def reference(c):
    if c["kind"] == "2d":
        P = REF.RefParameters(c["m"], c["s"], c["d"], c["r"])
        return REF.distribution_slice_compute(P, c["D"], c["a_d"], c["a_r"], method=c["method"],
                                              richardson=bool(c["rich"]))
    if c["kind"] in ("lin_d", "lin_r"):
        P = REF.RefParameters(c["m"], c["s"], c["d"], c["r"])
        return REF.linear_distribution_slice_compute(P, c["D"], c["a"], 0 if c["kind"] == "lin_d" else 1)
    P = REF.RefDiagonalParameters(c["m"], c["sigma"], c["s"], c["d"], c["r"], eta_bound=25)
    return REF.diagonal_distribution_slice_compute(P, c["D"], c["a"], c["eta"])


def check(c, cells, tp, te, fl):
    R = reference(c)
    e = cell_errors(cells, R.cells)
    assert e <= CELL_RTOL, (c, e)
    mass = abs(float(R.total_probability))
    assert abs(float(tp - R.total_probability)) <= 1e-12 * max(1.0, mass), c   # under-resolved far-out slices have |mass| >> 1
    if R.total_error != 0:
        assert abs(float((te - R.total_error) / R.total_error)) <= 1e-9, c
    assert int(fl) == R.flags, c


def test_kernel_math_vs_reference_random():
    from tests import hostsim as hs
    for c in cases(20260101, 300):
        if c["kind"] == "2d":
            cells, tp, te, fl = hs.slice2d(c["m"], c["l"], c["d"], c["r"], c["method"], c["rich"], c["D"],
                                           [c["a_d"]], [c["a_r"]])
            check(c, cells[0], tp[0], te[0], fl[0])
        else:
            kind = {"lin_d": 0, "lin_r": 1, "diag": 2}[c["kind"]]
            cells, tp, fl = hs.slice1d(c["m"], c["l"], c.get("sigma", 0), c["d"], c["r"], kind, 1, c["D"],
                                       [c["a"]], [c.get("eta", 0)])
            check(c, cells[0], tp[0], 0, fl[0])


@pytest.mark.gpu
def test_cuda_vs_reference_random(gpu_ctx):
    import qunundrum_b200 as qb
    for c in cases(20260102, 200):
        if c["kind"] == "2d":
            P = qb.Parameters(c["m"], c["s"], c["d"], c["r"])
            cells, tp, te, fl = gpu_ctx.slice2d_batch(P, c["method"], bool(c["rich"]), c["D"], [c["a_d"]],
                                                      [c["a_r"]])
            check(c, cells[0], tp[0], te[0], fl[0])
        elif c["kind"] == "diag":
            P = qb.Diagonal_Parameters(c["m"], c["sigma"], c["s"], c["d"], c["r"], eta_bound=25)
            cells, tp, fl = gpu_ctx.slice1d_batch(P, 2, True, c["D"], [c["a"]], [c["eta"]])
            check(c, cells[0], tp[0], 0, fl[0])
        else:
            P = qb.Parameters(c["m"], c["s"], c["d"], c["r"])
            cells, tp, fl = gpu_ctx.slice1d_batch(P, 0 if c["kind"] == "lin_d" else 1, True, c["D"], [c["a"]])
            check(c, cells[0], tp[0], 0, fl[0])


@pytest.mark.gpu
def test_fused_kernel_random(gpu_ctx):
    """The fused kernel (dimension a multiple of 32, Richardson) on random parameters: against the
    reference on a few slices, and against the plain kernels (validated above) on many, with
    coordinates over the whole region so that all slice classes, both modes of the inner-sine
    treatment, the bound-test and second-moment variants are hit."""
    import torch
    import qunundrum_b200 as qb

    def run(plan, algo):
        plan.set_algorithm(algo)
        cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
        summ = torch.empty(plan.n * 8, dtype=torch.float64, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            plan.run(cells.data_ptr(), summ.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        tp, te, fl = plan.finish(summ.cpu().numpy())
        return cells.cpu().numpy().reshape(plan.n, -1), tp, te, fl

    rnd = random.Random(77)
    n_ref = 0
    for it in range(24):
        m = rnd.choice([64, 128, 200, 256, 513, 1024, 2048, 3072])
        s = rnd.choice([1, 1, 2, 3, 4, 8, 20, 30])
        l = math.ceil(m / s)
        if l < 24:
            continue
        r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1)
        d = r // 2 + rnd.randrange(r // 2)
        method = rnd.choice([0, 0, 2])
        D = rnd.choice([32, 32, 64])
        lo, hi = max(1, m - 30), min(m + 29, m + l - 2)
        n = 12
        a_d = [rnd.randint(lo, hi) * rnd.choice([1, -1]) for _ in range(n)]
        a_r = [rnd.randint(lo, hi) * rnd.choice([1, 1, -1]) for _ in range(n)]
        a_d[0], a_r[0] = m, m                      # the ridge
        a_d[1], a_r[1] = -(m + 1), m + 1
        P = qb.Parameters(m, s, d, r)
        plan = gpu_ctx.plan2d(P, method, True, D, a_d, a_r)
        if plan.algorithm != 2:                    # l - sigma too small for the fused kernel
            plan.close()
            continue
        c2, tp2, te2, fl2 = run(plan, 2)
        c1, tp1, te1, fl1 = run(plan, 1)
        plan.close()
        assert cell_errors(c2, c1) <= 1e-10, (m, s, method, D)
        for i in range(n):
            assert abs(float(tp2[i] - tp1[i])) <= 1e-13 * max(1.0, abs(float(tp1[i]))), (m, s, i)
            if te1[i] != 0:
                assert abs(float((te2[i] - te1[i]) / te1[i])) <= 1e-10, (m, s, method, i)
        assert np.array_equal(fl1, fl2), (m, s, method, fl1, fl2)
        if D == 32 and n_ref < 6:                  # and the reference itself on the first two slices
            RP = REF.RefParameters(m, s, d, r)
            for i in range(2):
                R = REF.distribution_slice_compute(RP, D, a_d[i], a_r[i], method=method)
                assert cell_errors(c2[i], R.cells) <= CELL_RTOL, (m, s, method, i)
                assert int(fl2[i]) == R.flags
            n_ref += 1
    assert n_ref >= 3
