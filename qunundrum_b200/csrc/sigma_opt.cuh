// sigma_opt.cuh -- the sigma-optimal method (DISTRIBUTION_SLICE_COMPUTE_METHOD_OPTIMAL_LOCAL_SIGMA).
//
// Reference: distribution_slice_compute (src/distribution_slice_compute.cpp:265-287) walks the
// (2 D + 1)^2 points in loop order (alpha_d outer, alpha_r inner); the first point takes the
// global arg-min of the error bound over sigma in [1, l - 2] (probability_approx_optimal_sigma,
// src/probability.cpp:102-148), every later point starts from its predecessor's sigma and searches
// locally (probability_approx_adjust_sigma, src/probability.cpp:20-100). As written there, the
// first non-improving DECREASE returns immediately, so sigma never increases along the walk
// (the increasing loop is reachable only once sigma has walked down to 1). The walk is therefore
//     sigma_p = f_p(sigma_{p-1}),   f_p(s) <= s,   f_p monotone in s,
// and sigma_p is the fixed point reached by iterating  sigma <- f(prefix-min(sigma))  in
// parallel over all points (kernels_sigma_opt.cuh); this file holds the per-point pieces,
// __host__ __device__ so that tests/hostsim can walk the chain serially with the same code.
//
// The error bound at m = 2048 is ~2^-1000-sized and cannot be held in a double, so error values
// are compared as (mantissa, exponent) pairs; the choice of sigma is a comparison of error
// values, not a rounding-sensitive quantity, except at exact ties.
#pragma once

#include "integrands.cuh"

namespace qb200 {

// ---- extended-range positive numbers: f * 2^e, f in [1, 2) or f == 0 -----------------
struct xd {
  double f;
  int e;
};

QHD xd xd_make(double v, int e) {  // v >= 0 finite
  xd r;
  if (v == 0.0) {
    r.f = 0.0;
    r.e = -(1 << 28);
    return r;
  }
  int k;
  const double fr = frexp(v, &k);  // v = fr * 2^k, fr in [0.5, 1)
  r.f = fr * 2.0;
  r.e = e + k - 1;
  return r;
}
QHD xd xd_add(xd a, xd b) {
  if (a.f == 0.0) return b;
  if (b.f == 0.0) return a;
  if (a.e < b.e) {
    const xd t = a;
    a = b;
    b = t;
  }
  const int dlt = a.e - b.e;
  if (dlt > 120) return a;
  return xd_make(a.f + ldexp(b.f, -dlt), a.e);
}
QHD bool xd_less(xd a, xd b) {  // a < b
  if (a.f == 0.0 || b.f == 0.0) return b.f != 0.0 && a.f == 0.0;
  return a.e != b.e ? a.e < b.e : a.f < b.f;
}

struct SigmaOptConsts {
  unsigned long long q_mant[3];  // Q = mant * 2^q_exp (see hostconst.hpp)
  int q_exp;
};

// kappa_sigma = K_sigma / 2^sigma = -floor(Q * 2^sigma) / 2^sigma as a double-double.
QHD dd kappa_of_sigma(const SigmaOptConsts& q, int sigma) {
  // floor(mant * 2^(q_exp + sigma)): drop t = -(q_exp + sigma) low bits when t > 0
  unsigned long long w0 = q.q_mant[0], w1 = q.q_mant[1], w2 = q.q_mant[2];
  const int t = -(q.q_exp + sigma);
  if (t >= 192) return make_dd(0.0, 0.0);
  if (t > 0) {  // clear the low t bits (value stays scaled by 2^q_exp)
    if (t >= 128) {
      w0 = 0;
      w1 = 0;
      w2 &= ~0ull << (t - 128);
    } else if (t >= 64) {
      w0 = 0;
      w1 &= ~0ull << (t - 64);
    } else {
      w0 &= ~0ull << t;
    }
  }
  // (w2, w1, w0) -> double-double: split into 48-bit pieces so every partial sum is exact
  const double p5 = (double)(w2 >> 32), p4 = (double)(w2 & 0xffffffffull);
  const double p3 = (double)(w1 >> 32), p2 = (double)(w1 & 0xffffffffull);
  const double p1 = (double)(w0 >> 32), p0 = (double)(w0 & 0xffffffffull);
  dd acc = make_dd(p5, 0.0);
  acc = dd_add_d(dd_mul_pow2(acc, 4294967296.0), p4);
  acc = dd_add_d(dd_mul_pow2(acc, 4294967296.0), p3);
  acc = dd_add_d(dd_mul_pow2(acc, 4294967296.0), p2);
  acc = dd_add_d(dd_mul_pow2(acc, 4294967296.0), p1);
  acc = dd_add_d(dd_mul_pow2(acc, 4294967296.0), p0);
  // scale by 2^q_exp (|q_exp| ~ 192..200 for d ~ r: within range)
  const int e = q.q_exp;
  const double s1 = pow2i(e / 2), s2 = pow2i(e - e / 2);
  acc = dd_mul_pow2(dd_mul_pow2(acc, s1), s2);
  return dd_neg(acc);
}

struct SoPoint {  // everything of a grid point that does not depend on sigma
  dd xd_;      // x_d
  dd xr_;      // x_r
  double t2;   // second factor at x_r
  double h;    // |x_d| + |x_r|
};

// norm' = T1 T2 (norm * 2^(2m) / r) and the error bound e = error * 2^m (extended range):
//   e = s (2 + s) norm' r/2^m + 2^(4 - sigma) + 2^(3 - l),  s = pi 2^(sigma - l) h
// (src/probability.cpp:252-277).
QHD void so_eval(const DevConsts& c, const SigmaOptConsts& q, const SoPoint& p, int sigma,
                 double* norm, xd* err) {
  const dd kappa = kappa_of_sigma(q, sigma);
  const dd y = dd_mul(kappa, p.xr_);
  const double t1 = t1_value(p.xd_, y, c.l - sigma);
  const double n = t1 * p.t2;
  *norm = n;
  const int sl = sigma - c.l;  // < 0
  const double ph = 3.14159265358979323846 * p.h;
  const double s = sl > -1000 ? ldexp(ph, sl) : 0.0;
  xd e = xd_make(ph * (2.0 + s) * n * c.r_m, sl);
  e = xd_add(e, xd_make(1.0, 4 - sigma));
  e = xd_add(e, xd_make(1.0, 3 - c.l));
  *err = e;
}

// error / norm <= (double)0.01f  (src/probability.cpp:280-281)
QHD bool so_bounded(const DevConsts& c, double norm, xd err) {
  return !xd_less(xd_make(QB_ERROR_BOUND * norm * c.r_m, 0), err);
}

// probability_approx_adjust_sigma (src/probability.cpp:20-100): returns the new sigma.
// `increased` is set when the increasing loop moved sigma up (only possible from sigma == 1).
QHD int so_adjust(const DevConsts& c, const SigmaOptConsts& q, const SoPoint& p, int start,
                  double* norm, xd* err, bool* increased) {
  double bn, n;
  xd be, e;
  so_eval(c, q, p, start, &bn, &be);
  int best = start;
  *increased = false;
  for (int sigma = best - 1; sigma >= 1; sigma--) {
    so_eval(c, q, p, sigma, &n, &e);
    if (!xd_less(e, be)) {  // error >= best_error: the reference returns here
      *norm = bn;
      *err = be;
      return best;
    }
    bn = n;
    be = e;
    best = sigma;
  }
  for (int sigma = best + 1; sigma < c.l - 1; sigma++) {
    so_eval(c, q, p, sigma, &n, &e);
    if (!xd_less(e, be)) break;
    bn = n;
    be = e;
    best = sigma;
    *increased = true;
  }
  *norm = bn;
  *err = be;
  return best;
}

// ---- the walk in closed form (large l) -----------------------------------------------------------
// With a = pi h n r/2^m, the bound at a point is  e(sigma) = a 2^(sigma-l) (2 + pi h 2^(sigma-l))
// + 2^(4-sigma) + 2^(3-l)  (so_eval). Two facts make the serial walk a prefix minimum when l is
// large (l >= 256; the walk then stays near sigma = l / 2):
//  (1) n does not depend on sigma there: kappa_sigma = -d/r + O(2^-sigma) moves u by less than
//      2^(11-sigma) and Lambda sin(pi u / Lambda) = pi u to 2^-56, so for sigma >= 64 and
//      l - sigma >= 60 the norm is the quick method's, to the last bits of a double;
//  (2) e is then convex in 2^sigma, and the descent of probability_approx_adjust_sigma from any
//      start s stops at the first sigma with e(sigma - 1) >= e(sigma), i.e.
//          16 >= a 2^(2 sigma - l) (1 + 0.75 pi h 2^(sigma - l))           [the bracket is 1 to 2^-60]
//      which holds exactly for sigma <= sigma* := max { sigma : a 2^(2 sigma - l) <= 16 }.
//      Hence f_p(s) = min(s, sigma*_p) and sigma_p = min(sigma_0, sigma*_1, ..., sigma*_p).
// sigma*_p needs no logarithm: a = f 2^e with f in [0.5, 1) gives 2 sigma - l + e <= 4 (5 if f = 0.5).
// The range assumptions are CHECKED on the device for every point (64 <= sigma_p <= l - 60); a
// slice that leaves them is computed by the general fixed-point iteration instead.
QHD int so_fast_sigma_star(int l, double a) {
  if (!(a > 0.0)) return 0x3fffffff;  // n = 0: e decreases with sigma for ever, the descent never moves
  int e;
  bool half;
#if defined(__CUDA_ARCH__)
  const unsigned long long bits = (unsigned long long)__double_as_longlong(a);
  const int field = (int)(bits >> 52);  // a > 0: no sign bit
  if (field != 0 && field != 0x7ff) {   // a normal number: frexp without its special cases
    e = field - 1022;
    half = (bits & 0xfffffffffffffull) == 0;
  } else
#endif
  {
    half = frexp(a, &e) == 0.5;
  }
  const int kmax = half ? 5 : 4;
  const int t = kmax + l - e;  // 2 sigma <= t
  return t >= 0 ? t / 2 : -((1 - t) / 2);  // floor(t / 2)
}

// The first point takes the arg-min over all sigma, the SMALLEST among equal values
// (probability_approx_optimal_sigma, src/probability.cpp:102-148: strict improvement while scanning
// upwards): sigma* itself unless e(sigma* - 1) == e(sigma*), i.e. a 2^(2 sigma* - l) == 16 exactly.
QHD int so_fast_sigma_first(int l, double a) {
  const int s = so_fast_sigma_star(l, a);
  if (s >= 0x3fffffff) return s;
  return ldexp(a, 2 * s - l) == 16.0 ? s - 1 : s;
}

// e(sigma) for a given norm (so_eval without the evaluation of the norm).
QHD xd so_error_given_norm(const DevConsts& c, double ph, double n, int sigma) {
  const int sl = sigma - c.l;
  const double s = sl > -1000 ? ldexp(ph, sl) : 0.0;
  xd e = xd_make(ph * (2.0 + s) * n * c.r_m, sl);
  e = xd_add(e, xd_make(1.0, 4 - sigma));
  e = xd_add(e, xd_make(1.0, 3 - c.l));
  return e;
}

// A point's terms of the two error sums of a slice, both relative to sigma_0 of the pass (s0 >= sg):
//   *era = pi h (2 + s) n r/2^m 2^(sg - s0),  s = pi h 2^(sg - l);    *cterm = 2^(s0 - sg)
// i.e. 2^(l - s0) times the first and 2^(s0 - 4) times the second term of e(sg). Scaling by a power
// of two is a multiplication here (exact, or rounded once where the product is subnormal, as ldexp).
QHD void so_fast_terms(const DevConsts& c, double ph, double n, int sg, int s0, double* era, double* cterm) {
  const int sl = sg - c.l;
  const double sv = sl > -1000 ? ph * pow2i(sl) : 0.0;
  const int down = sg - s0 > -1000 ? sg - s0 : -1000, up = s0 - sg < 1000 ? s0 - sg : 1000;
  *era = ph * (2.0 + sv) * n * c.r_m * pow2i(down);
  *cterm = pow2i(up);
}

// so_bounded(c, n, e(sg)) for a point of the closed-form walk, where 64 <= sg <= sigma* (the prefix
// minimum). There a 2^(2 sg - l) <= 16 with a = pi h n r/2^m, so
//     e(sg) 2^sg = a (2 + s) 2^(2 sg - l) + 16 + 2^(3 - l + sg) < 49,
// and the bound holds whenever 0.01 n r/2^m 2^sg >= 49: decided without extended-range arithmetic
// for every point whose norm is not vanishingly small; the others take the comparison itself.
QHD bool so_fast_bounded(const DevConsts& c, double ph, double n, int sg) {
  const double nr = n * c.r_m;
  // 49 / 0.01f = 4900.0002; beyond sigma = 1000 the (stricter) threshold of sigma = 1000
  const double need = 4900.001 * pow2i(sg <= 1000 ? -sg : -1000);
  if (nr >= need) return true;
  return so_bounded(c, n, so_error_given_norm(c, ph, n, sg));
}

// Simpson weight of abscissa i (0 .. 2 Dp) of one axis over the cells it belongs to:
// 4 w[I] for the mid-point of cell I, w[I - 1] + w[I] for a point shared by two cells.
QHD double so_axis_weight(const double* w, int Dp, int i) {
  if (i & 1) return 4.0 * w[i >> 1];
  const int I = i >> 1;
  return (I > 0 ? w[I - 1] : 0.0) + (I < Dp ? w[I] : 0.0);
}

}  // namespace qb200
