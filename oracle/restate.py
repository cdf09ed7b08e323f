"""CPU restatement of the reference's slice-integration path (mpmath).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never from qunundrum_b200/ (the product has no CPU
path).

Every function follows the reference line by line at the reference's own
working precisions (MPFR round-to-nearest == mpmath's default rounding):

  function here                          reference (file:line)
  deterministic_d_r                      src/parameters_selection.cpp:21-77
  Parameters / DiagonalParameters        src/parameters.cpp:30-133, src/diagonal_parameters.cpp:30-131
  heuristic_sigma                        src/distribution_slice_compute.cpp:149-158
  probability_approx                     src/probability.cpp:150-288
  probability_approx_quick               src/probability.cpp:290-372
  linear_probability_d                   src/linear_probability.cpp:21-168
  linear_probability_r                   src/linear_probability.cpp:170-263
  diagonal_probability_approx_f_eta      src/diagonal_probability.cpp:18-97
  distribution_slice_compute             src/distribution_slice_compute.cpp:38-453
  distribution_slice_compute_richardson  src/distribution_slice_compute_richardson.cpp:17-73
  linear_distribution_slice_compute      src/linear_distribution_slice_compute.cpp:30-245
  linear_..._compute_richardson          src/linear_distribution_slice_compute_richardson.cpp:17-65
  diagonal_distribution_slice_compute    src/diagonal_distribution_slice_compute.cpp:30-210
  diagonal_..._compute_richardson        src/diagonal_distribution_slice_compute_richardson.cpp:17-65

The arithmetic library of the reference (MPFR 4.2.1 / GMP 6.3.0, system
libraries, not vendored under /root/reference) is replaced by mpmath 1.3.0.
Pinning: tests/test_oracle_kat.py checks this file against the reference's
known-answer vectors (committed sample under tests/golden/kat/) and against
oracle/_ref (the reference itself, compiled), so parity is PINNED.

It is slow (pure Python): use it for small dimensions only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import mpmath as mp
import numpy as np

PRECISION = 192  # src/common.h:13

SLICE_FLAGS_ERROR_BOUND_WARNING = 0x00000001
SLICE_FLAGS_METHOD_SIMPSON = 0x00020000
SLICE_FLAGS_METHOD_RICHARDSON = 0x00080000
SLICE_FLAGS_MASK_METHOD = 0x000F0000

METHOD_HEURISTIC_SIGMA = 0
METHOD_OPTIMAL_LOCAL_SIGMA = 1
METHOD_QUICK = 2
TARGET_D = 0
TARGET_R = 1

# (double)0.01f, src/probability.cpp:281
_BOUND = float(np.float32(0.01))


def deterministic_d_r(m: int) -> tuple[int, int]:
    """src/parameters_selection.cpp:21-77."""
    if m > 8192:
        raise ValueError("m must be <= 8192")
    precision = 3 * 8192
    with mp.workprec(precision):
        data = mp.floor(mp.catalan * mp.mpf(2) ** precision)
        data_z = int(data)
    length = data_z.bit_length()
    r = 0
    for i in range(m - 1):
        if (data_z >> ((length - 1) - i)) & 1:
            r |= 1 << (m - 2 - i)
    d = 0
    for i in range(m - 1):
        if (data_z >> ((length - 1) - (8192 - 1) - i)) & 1:
            d |= 1 << (m - 2 - i)
    d %= r
    r |= 1 << (m - 1)
    d |= 1 << (m - 1)
    return d, r


@dataclass
class Parameters:
    """src/parameters.h:33-114; regions per src/parameters.cpp:30-51."""
    m: int
    s: int
    d: int
    r: int
    t: int = 30
    l: int = 0

    def __post_init__(self):
        if self.l == 0:
            self.l = int(math.ceil(self.m / self.s))
        else:
            self.s = 0
        self.min_alpha_d = 0 if self.t > self.m else self.m - self.t
        if self.t >= self.l:
            self.max_alpha_d = self.m + self.l - 2
        else:
            self.max_alpha_d = self.m + self.t - 1
        self.min_alpha_r = self.min_alpha_d
        self.max_alpha_r = self.max_alpha_d


@dataclass
class DiagonalParameters:
    """src/diagonal_parameters.h:32-106; src/diagonal_parameters.cpp:30-131."""
    m: int
    sigma: int
    s: int
    d: int
    r: int
    eta_bound: int = 0
    t: int = 30
    l: int = 0

    def __post_init__(self):
        if self.l == 0:
            self.l = int(math.ceil(self.m / self.s))
        else:
            self.s = 0
        self.min_alpha_r = 0 if self.t > self.m else self.m - self.t
        if self.t >= self.sigma:
            self.max_alpha_r = self.m + self.sigma - 2
        else:
            self.max_alpha_r = self.m + self.t - 1


def heuristic_sigma(l: int) -> int:
    f = np.float32
    v = (f(l) + f(11) + f(4) - f(1.6515)) / f(2.0)
    return int(math.floor(float(v) + 0.5))


def _pow2(e: int):
    return mp.ldexp(mp.mpf(1), e)


# --------------------------------------------------------------------------- #
# Integrands                                                                  #
# --------------------------------------------------------------------------- #

def probability_approx(sigma: int, theta_d, theta_r, p: Parameters):
    """src/probability.cpp:150-288. Returns (norm, error, bounded)."""
    with mp.workprec(PRECISION):
        tmp = _pow2(sigma)
        tmp2 = -tmp
        tmp2 = tmp2 * p.d            # mpfr_mul_z, rounded to 192 bits
        tmp2 = tmp2 / p.r            # mpfr_div_z
        tmp2 = mp.ceil(tmp2)
        tmp2 = theta_r * tmp2
        tmp = theta_d * tmp
        tmp = tmp + tmp2
        tmp2 = _pow2(p.l - sigma)
        if tmp == 0:
            norm = tmp2 * tmp2
        else:
            tmp = tmp / 2
            tmp2 = tmp2 * tmp
            tmp2 = mp.sin(tmp2)
            tmp = mp.sin(tmp)
            tmp = tmp2 / tmp
            norm = tmp * tmp
        tmp2 = _pow2(p.m + p.l)
        tmp2 = tmp2 / p.r
        tmp2 = mp.ceil(tmp2)
        if theta_r == 0:
            tmp2 = tmp2 * tmp2
            norm = norm * tmp2
        else:
            tmp = theta_r / 2
            tmp2 = tmp * tmp2
            tmp2 = mp.sin(tmp2)
            tmp = mp.sin(tmp)
            tmp = tmp2 / tmp
            tmp = tmp * tmp
            norm = norm * tmp
        tmp2 = _pow2(2 * sigma - 2 * (p.m + 2 * p.l))
        tmp2 = tmp2 * p.r
        norm = norm * tmp2

        tmp = abs(theta_d)
        tmp2 = abs(theta_r)
        tmp = tmp + tmp2
        tmp2 = _pow2(sigma)
        tmp2 = tmp2 / 2
        tmp = tmp2 * tmp
        tmp2 = tmp + 2
        tmp = tmp * tmp2
        tmp = tmp * norm
        tmp2 = _pow2(4 - (p.m + sigma))
        tmp = tmp + tmp2
        tmp2 = _pow2(3 - (p.m + p.l))
        error = tmp + tmp2
        tmp = error / norm
        bounded = bool(tmp <= mp.mpf(_BOUND))
        return norm, error, bounded


def probability_approx_optimal_sigma(theta_d, theta_r, p: Parameters):
    """src/probability.cpp:102-148. Returns (norm, error, sigma, bounded)."""
    with mp.workprec(PRECISION):
        best_norm = mp.nan
        best_error = mp.mpf(1)
        best_sigma = 0
        best_bounded = False
        for sigma in range(1, p.l - 1):
            norm, error, bounded = probability_approx(sigma, theta_d, theta_r, p)
            if error < best_error:
                best_error, best_norm, best_sigma, best_bounded = error, norm, sigma, bounded
        return best_norm, best_error, best_sigma, best_bounded


def probability_approx_adjust_sigma(best_sigma: int, theta_d, theta_r, p: Parameters):
    """src/probability.cpp:20-100. Returns (norm, error, sigma, bounded).

    Note the reference's control flow: the first non-improving DECREASE returns
    at once (`best_sigma != sigma` always holds there, :59-67), so the increasing
    search (:77-93) only runs when sigma has walked down to 1."""
    with mp.workprec(PRECISION):
        norm, error, bounded = probability_approx(best_sigma, theta_d, theta_r, p)
        best_norm, best_error, best_bounded = norm, error, bounded
        sigma = best_sigma - 1
        while sigma >= 1:
            norm, error, bounded = probability_approx(sigma, theta_d, theta_r, p)
            if error >= best_error:
                return best_norm, best_error, best_sigma, best_bounded
            best_norm, best_error, best_bounded, best_sigma = norm, error, bounded, sigma
            sigma -= 1
        sigma = best_sigma + 1
        while sigma < p.l - 1:
            norm, error, bounded = probability_approx(sigma, theta_d, theta_r, p)
            if error >= best_error:
                break
            best_norm, best_error, best_bounded, best_sigma = norm, error, bounded, sigma
            sigma += 1
        return best_norm, best_error, best_sigma, best_bounded


def probability_approx_quick(theta_d, theta_r, p: Parameters):
    """src/probability.cpp:290-372."""
    with mp.workprec(PRECISION):
        tmp = theta_r * p.d
        tmp = tmp / p.r
        tmp = theta_d - tmp
        if tmp == 0:
            norm = _pow2(2 * p.l)
        else:
            tmp = tmp / 2
            tmp2 = _pow2(p.l)
            tmp2 = tmp * tmp2
            tmp = mp.sin(tmp)
            tmp2 = mp.sin(tmp2)
            tmp = tmp2 / tmp
            norm = tmp * tmp
        tmp = _pow2(p.m + p.l)
        tmp = tmp / p.r
        tmp = mp.ceil(tmp)
        if theta_r == 0:
            tmp2 = tmp * tmp
            norm = norm * tmp2
        else:
            tmp2 = theta_r / 2
            tmp = tmp * tmp2
            tmp = mp.sin(tmp)
            tmp2 = mp.sin(tmp2)
            tmp = tmp / tmp2
            tmp = tmp * tmp
            norm = norm * tmp
        norm = norm * p.r
        tmp = _pow2(2 * (p.m + 2 * p.l))
        norm = norm / tmp
        return norm


def linear_probability_d(theta_d, p: Parameters):
    """src/linear_probability.cpp:21-168 (precision 3 * max(m, 192))."""
    precision = 3 * max(p.m, PRECISION)
    with mp.workprec(precision):
        theta_d = mp.mpf(theta_d)
        if theta_d == 0:
            tmp = _pow2(p.l) - 1
            tmp = tmp * p.d
            tmp2 = _pow2(p.l + 1) - 1
            tmp3 = tmp * tmp2
            tmp2 = _pow2(p.l)
            tmp3 = tmp3 * tmp2
            tmp3 = tmp3 / 3
            tmp2 = _pow2(p.l + p.m)
            tmp = tmp2 - tmp
            tmp2 = _pow2(2 * p.l)
            tmp2 = tmp * tmp2
            tmp3 = tmp3 + tmp2
            tmp = _pow2(2 * (p.m + 2 * p.l))
            norm = tmp3 / tmp
        else:
            tmp = theta_d / 2
            tmp2 = mp.sin(tmp)
            tmp2 = tmp2 * tmp2
            one_minus_cos_theta = tmp2 * 2

            tmp3 = _pow2(p.l)
            tmp2 = tmp3 * tmp
            tmp2 = mp.sin(tmp2)
            tmp2 = tmp2 * tmp2
            one_minus_cos_2l_theta = tmp2 * 2

            tmp3 = tmp3 - 1
            tmp2 = tmp3 * tmp
            tmp2 = mp.sin(tmp2)
            tmp2 = tmp2 * tmp2
            tmp2 = tmp2 * 2

            tmp2 = one_minus_cos_2l_theta - tmp2
            tmp2 = tmp2 / one_minus_cos_theta
            tmp2 = tmp2 - 1
            tmp2 = tmp2 / 2
            tmp2 = tmp3 - tmp2
            tmp2 = tmp2 * 2
            tmp2 = tmp2 * p.d

            tmp3 = tmp3 * p.d
            tmp = _pow2(p.l + p.m)
            tmp3 = tmp - tmp3
            tmp3 = tmp3 * one_minus_cos_2l_theta

            tmp = tmp3 + tmp2
            tmp = tmp / one_minus_cos_theta
            tmp2 = _pow2(2 * (p.m + 2 * p.l))
            norm = tmp / tmp2
    with mp.workprec(PRECISION):
        return +norm  # stored into a PRECISION-bit mpfr_t by the caller


def linear_probability_r(theta_r, p: Parameters):
    """src/linear_probability.cpp:170-263."""
    with mp.workprec(PRECISION):
        beta = (1 << (p.l + p.m)) % p.r
        N = _pow2(p.l + p.m)
        N = N / p.r
        N = mp.floor(N)
        if theta_r == 0:
            tmp = N + 1
            tmp = tmp * tmp
            tmp = tmp * beta
            beta = p.r - beta
            tmp2 = N * N
            tmp2 = tmp2 * beta
            tmp = tmp + tmp2
        else:
            tmp3 = theta_r / 2
            tmp = N + 1
            tmp = tmp * tmp3
            tmp = mp.sin(tmp)
            tmp2 = mp.sin(tmp3)
            tmp3 = tmp3 * N
            tmp3 = mp.sin(tmp3)
            tmp = tmp / tmp2
            tmp = tmp * tmp
            tmp3 = tmp3 / tmp2
            tmp3 = tmp3 * tmp3
            tmp = tmp * beta
            beta = p.r - beta
            tmp3 = tmp3 * beta
            tmp = tmp + tmp3
        tmp2 = _pow2(2 * (p.l + p.m))
        return tmp / tmp2


def diagonal_probability_approx_f_eta(theta_r, eta: int, p: DiagonalParameters):
    """src/diagonal_probability.cpp:18-97."""
    if theta_r == 0 and eta == 0:
        with mp.workprec(PRECISION):
            return mp.mpf(1) / p.r
    precision = max(2 * (p.m + p.sigma), PRECISION)
    with mp.workprec(precision):
        tmp2 = +mp.pi
        tmp2 = tmp2 * 2
        tmp2 = tmp2 * eta
        tmp2 = theta_r - tmp2
        tmp = _pow2(p.m + p.sigma)
        tmp = tmp / p.r
        tmp = tmp2 * tmp
        tmp = tmp / 2
        tmp = mp.sin(tmp)
        tmp = tmp * tmp
        tmp = tmp * 4
        tmp2 = tmp2 * tmp2
        tmp = tmp / tmp2
        tmp = tmp * p.r
        tmp2 = _pow2(2 * (p.m + p.sigma))
        norm = tmp / tmp2
    with mp.workprec(PRECISION):
        return +norm


def diagonal_probability_approx_h(phi, p: DiagonalParameters):
    """src/diagonal_probability.cpp:99-162: (1 - cos(2^l phi)) / (2^2l (1 - cos phi))."""
    if phi == 0:
        return mp.mpf(1)
    precision = 2 * max(p.m + p.sigma, PRECISION)
    with mp.workprec(precision):
        tmp2 = phi / 2
        tmp = _pow2(p.l)
        tmp = phi * tmp
        tmp = tmp / 2
        tmp = mp.sin(tmp)
        tmp = tmp * tmp
        tmp = tmp * 2
        tmp2 = mp.sin(tmp2)
        tmp2 = tmp2 * tmp2
        tmp2 = tmp2 * 2
        tmp = tmp / tmp2
        tmp2 = _pow2(2 * p.l)
        return tmp / tmp2


def _mod_reduce(x: int, n: int) -> int:
    """mod_reduce (src/math.cpp): x mod n on [-n/2, n/2)."""
    x %= n
    return x - n if x >= (n >> 1) else x


def sample_k_from_diagonal_j_eta_pivot(p: DiagonalParameters, pivot, j: int, eta: int,
                                       delta_bound: int):
    """src/sample.cpp:412-646. pivot: np.longdouble. Returns (ok, k, alpha_phi) with alpha_phi
    an mpf at 3 max(m + sigma, PRECISION) bits."""
    pivot = np.longdouble(pivot)
    assert 0 <= pivot <= 1
    precision = 3 * max(p.m + p.sigma, PRECISION)
    pow2_m_sigma = 1 << (p.m + p.sigma)
    pow2_m_sigma_l = 1 << (p.m + p.sigma - p.l)
    pow2_l = 1 << p.l
    alpha_r = _mod_reduce(p.r * j, pow2_m_sigma)                      # :477-478
    with mp.workprec(precision):
        tmp_z = alpha_r - pow2_m_sigma * eta                           # :481-482
        tmp = mp.mpf(tmp_z)
        tmp = tmp * p.d
        tmp = tmp / p.r
        tmp = tmp - p.d * j                                            # :490-492
        tmp = tmp / pow2_m_sigma_l
        k0 = int(_round_int(tmp)) % pow2_l                             # :498-500
        scale = +mp.pi
        scale = scale * 2
        scale = scale / pow2_m_sigma                                   # :514-517
        term = mp.mpf(tmp_z)
        term = term * p.d
        term = term / p.r
        term = term - p.d * j
        term = -term                                                   # :525-537
        half = mp.mpf(pow2_m_sigma) / 2
        delta_abs = 0
        while delta_abs <= delta_bound:
            for sgn in (1, -1):
                if delta_abs == 0 and sgn == -1:
                    continue
                k = (k0 + sgn * delta_abs) % pow2_l                    # :552-559
                phi = term + pow2_m_sigma_l * k
                phi = mp.fmod(phi, mp.mpf(pow2_m_sigma))               # :566-568
                if phi >= half:
                    phi = phi - pow2_m_sigma
                alpha_phi = phi
                phi = phi * scale
                h = diagonal_probability_approx_h(phi, p)
                pivot = pivot - _get_ld(h)                             # :594
                if pivot <= 0:
                    return True, k, alpha_phi
            delta_abs += 1
    return False, 0, mp.mpf(0)


# --------------------------------------------------------------------------- #
# Slices                                                                      #
# --------------------------------------------------------------------------- #

def _get_ld(x) -> np.longdouble:
    """mpfr_get_ld: round to the 64-bit significand of x87 long double."""
    with mp.workprec(64):
        y = +x
    hi = float(y)
    with mp.workprec(128):
        lo = float(y - mp.mpf(hi))
    return np.longdouble(hi) + np.longdouble(lo)


def _round_int(x):
    """mpfr_round: nearest integer, halfway cases away from zero."""
    return mp.floor(x + mp.mpf(0.5)) if x >= 0 else -mp.floor(-x + mp.mpf(0.5))


def _grid(min_log_alpha: int, dimension: int, precision: int):
    """Signed alpha at the 2 * dimension + 1 interleaved main / average points
    (src/distribution_slice_compute.cpp:196-243), and the cell widths
    (src/distribution_slice_compute.cpp:331-347)."""
    with mp.workprec(precision):
        step = 1.0 / float(dimension)
        pow_2step = mp.power(2, mp.mpf(step))
        sgn = -1 if min_log_alpha < 0 else 1
        alphas = []
        max_alpha = _pow2(abs(min_log_alpha))
        min_alpha = max_alpha
        for i in range(2 * dimension + 1):
            if i % 2 == 0:
                min_alpha = max_alpha
                alpha = _round_int(min_alpha)
            else:
                max_alpha = min_alpha * pow_2step
                alpha = min_alpha + max_alpha
                alpha = alpha / 2
                alpha = _round_int(alpha)
            alphas.append(sgn * alpha)
    with mp.workprec(PRECISION):
        pow_2step = mp.power(2, mp.mpf(1.0 / float(dimension)))
        widths = []
        max_alpha = _pow2(abs(min_log_alpha))
        for i in range(dimension):
            min_alpha = max_alpha
            max_alpha = min_alpha * pow_2step
            widths.append(max_alpha - min_alpha)
    return alphas, widths


class Slice:
    def __init__(self, dimension: int, ndim: int):
        self.dimension = dimension
        self.cells = np.zeros(dimension ** ndim, dtype=np.longdouble)
        self.total_probability = np.longdouble(0)
        self.total_error = np.longdouble(0)
        self.flags = 0


def distribution_slice_compute(p: Parameters, dimension: int,
                               min_log_alpha_d: int, min_log_alpha_r: int,
                               method: int = METHOD_HEURISTIC_SIGMA) -> Slice:
    """src/distribution_slice_compute.cpp:38-453. For the sigma-optimal method the
    sigma chosen at every point is kept in the returned slice's `sigmas`."""
    if method not in (METHOD_HEURISTIC_SIGMA, METHOD_OPTIMAL_LOCAL_SIGMA, METHOD_QUICK):
        raise ValueError("unknown method")
    sl = Slice(dimension, 2)
    sl.sigmas = []
    n = 2 * dimension + 1
    with mp.workprec(PRECISION):
        scale = 2 * mp.pi / _pow2(p.l + p.m)
        pow_2m = _pow2(p.m)
        sigma = heuristic_sigma(p.l)
        ad, wd = _grid(min_log_alpha_d, dimension, PRECISION)
        ar, wr = _grid(min_log_alpha_r, dimension, PRECISION)
        norm = [[None] * n for _ in range(n)]
        err = [[None] * n for _ in range(n)] if method != METHOD_QUICK else None
        bounded = True
        for i in range(n):
            theta_d = ad[i] * scale
            for j in range(n):
                theta_r = ar[j] * scale
                if method == METHOD_QUICK:
                    norm[i][j] = probability_approx_quick(theta_d, theta_r, p)
                elif method == METHOD_OPTIMAL_LOCAL_SIGMA:
                    if i == 0 and j == 0:
                        nn, ee, sigma, bb = probability_approx_optimal_sigma(theta_d, theta_r, p)
                    else:
                        nn, ee, sigma, bb = probability_approx_adjust_sigma(sigma, theta_d, theta_r, p)
                    norm[i][j], err[i][j] = nn, ee
                    bounded = bounded and bb
                    sl.sigmas.append(sigma)
                else:
                    nn, ee, bb = probability_approx(sigma, theta_d, theta_r, p)
                    norm[i][j], err[i][j] = nn, ee
                    bounded = bounded and bb

        def simpson(a, i, j):
            v = 4 * a[i + 1][j + 1]
            v = v + a[i][j + 1]
            v = v + a[i + 2][j + 1]
            v = v + a[i + 1][j]
            v = v + a[i + 1][j + 2]
            v = v * 4
            v = v + a[i][j]
            v = v + a[i + 2][j]
            v = v + a[i][j + 2]
            v = v + a[i + 2][j + 2]
            return v / 36

        for i in range(0, 2 * dimension, 2):
            for j in range(0, 2 * dimension, 2):
                avg = simpson(norm, i, j)
                avg = avg * wd[i // 2]
                avg = avg * wr[j // 2]
                avg = avg / pow_2m
                v = _get_ld(avg)
                sl.cells[(i // 2) + dimension * (j // 2)] = v
                sl.total_probability += v
                if err is not None:
                    e = simpson(err, i, j)
                    e = e * wd[i // 2]
                    e = e * wr[j // 2]
                    e = e / pow_2m
                    sl.total_error += _get_ld(e)
    sl.flags &= ~SLICE_FLAGS_MASK_METHOD
    sl.flags |= SLICE_FLAGS_METHOD_SIMPSON
    if not bounded:
        sl.flags |= SLICE_FLAGS_ERROR_BOUND_WARNING
    return sl


def distribution_slice_compute_richardson(p: Parameters, dimension: int,
                                          min_log_alpha_d: int,
                                          min_log_alpha_r: int,
                                          method: int = METHOD_HEURISTIC_SIGMA) -> Slice:
    """src/distribution_slice_compute_richardson.cpp:17-73."""
    sl = distribution_slice_compute(p, dimension, min_log_alpha_d, min_log_alpha_r, method)
    db = distribution_slice_compute(p, 2 * dimension, min_log_alpha_d, min_log_alpha_r, method)
    D = dimension
    sl.total_probability = np.longdouble(0)
    for i in range(D):
        for j in range(D):
            probability = sl.cells[D * j + i]
            dp = (db.cells[(2 * D) * (2 * j) + (2 * i)] +
                  db.cells[(2 * D) * (2 * j) + (2 * i + 1)] +
                  db.cells[(2 * D) * (2 * j + 1) + (2 * i)] +
                  db.cells[(2 * D) * (2 * j + 1) + (2 * i + 1)])
            sl.cells[D * j + i] = 2 * dp - probability
            sl.total_probability += sl.cells[D * j + i]
    sl.total_error = 2 * db.total_error - sl.total_error
    sl.flags |= SLICE_FLAGS_METHOD_RICHARDSON
    return sl


def linear_distribution_slice_compute(p: Parameters, dimension: int,
                                      min_log_alpha: int, target: int) -> Slice:
    """src/linear_distribution_slice_compute.cpp:30-245."""
    sl = Slice(dimension, 1)
    with mp.workprec(PRECISION):
        scale = 2 * mp.pi / _pow2(p.l + p.m)
        pow_2l = _pow2(p.l)
        al, w = _grid(min_log_alpha, dimension, PRECISION)
        norm = []
        for i in range(2 * dimension + 1):
            theta = al[i] * scale
            if target == TARGET_D:
                norm.append(linear_probability_d(theta, p))
            else:
                norm.append(linear_probability_r(theta, p))
        for i in range(0, 2 * dimension, 2):
            avg = 4 * norm[i + 1]
            avg = avg + norm[i]
            avg = avg + norm[i + 2]
            avg = avg / 6
            avg = avg * w[i // 2]
            if target == TARGET_D:
                avg = avg * pow_2l
            v = _get_ld(avg)
            sl.cells[i // 2] = v
            sl.total_probability += v
    sl.flags &= ~SLICE_FLAGS_MASK_METHOD
    sl.flags |= SLICE_FLAGS_METHOD_SIMPSON
    return sl


def _richardson_1d(sl: Slice, db: Slice) -> Slice:
    D = sl.dimension
    sl.total_probability = np.longdouble(0)
    for i in range(D):
        probability = sl.cells[i]
        dp = db.cells[2 * i] + db.cells[2 * i + 1]
        sl.cells[i] = 2 * dp - probability
        sl.total_probability += sl.cells[i]
    sl.total_error = 2 * db.total_error - sl.total_error
    sl.flags |= SLICE_FLAGS_METHOD_RICHARDSON
    return sl


def linear_distribution_slice_compute_richardson(p: Parameters, dimension: int,
                                                 min_log_alpha: int,
                                                 target: int) -> Slice:
    """src/linear_distribution_slice_compute_richardson.cpp:17-65."""
    sl = linear_distribution_slice_compute(p, dimension, min_log_alpha, target)
    db = linear_distribution_slice_compute(p, 2 * dimension, min_log_alpha, target)
    return _richardson_1d(sl, db)


def diagonal_distribution_slice_compute(p: DiagonalParameters, dimension: int,
                                        min_log_alpha_r: int, eta: int) -> Slice:
    """src/diagonal_distribution_slice_compute.cpp:30-210."""
    sl = Slice(dimension, 1)
    precision = 2 * max(p.m + p.sigma, PRECISION)
    with mp.workprec(precision):
        scale = 2 * mp.pi / _pow2(p.m + p.sigma)
    al, w = _grid(min_log_alpha_r, dimension, precision)
    norm = []
    for i in range(2 * dimension + 1):
        with mp.workprec(precision):
            theta = scale * al[i]
        norm.append(diagonal_probability_approx_f_eta(theta, eta, p))
    with mp.workprec(PRECISION):
        for i in range(0, 2 * dimension, 2):
            avg = 4 * norm[i + 1]
            avg = avg + norm[i]
            avg = avg + norm[i + 2]
            avg = avg / 6
            avg = avg * w[i // 2]
            v = _get_ld(avg)
            sl.cells[i // 2] = v
            sl.total_probability += v
    sl.flags &= ~SLICE_FLAGS_MASK_METHOD
    sl.flags |= SLICE_FLAGS_METHOD_SIMPSON
    sl.eta = eta
    return sl


def diagonal_distribution_slice_compute_richardson(p: DiagonalParameters,
                                                   dimension: int,
                                                   min_log_alpha_r: int,
                                                   eta: int) -> Slice:
    """src/diagonal_distribution_slice_compute_richardson.cpp:17-65."""
    sl = diagonal_distribution_slice_compute(p, dimension, min_log_alpha_r, eta)
    db = diagonal_distribution_slice_compute(p, 2 * dimension, min_log_alpha_r, eta)
    out = _richardson_1d(sl, db)
    out.eta = eta
    return out
