// integrands.cuh -- the reference's closed-form probabilities in scale-free form.
//
// Notation: x = alpha / 2^m (signed), L = 2^l, Lambda = 2^(l - sigma).
// All functions return the integrand in "mass density" units, i.e. such that
//
//   cell mass = Simpson average of the value * (Delta alpha / 2^m) [* ...]
//
// exactly as the reference's slice drivers weight their point values
// (src/distribution_slice_compute.cpp:386-390,
//  src/linear_distribution_slice_compute.cpp:189-198,
//  src/diagonal_distribution_slice_compute.cpp:185).
//
//   reference function (file:line)                         here
//   probability_approx           src/probability.cpp:150   t1_value * t2_value (* r/2^m)
//   probability_approx_quick     src/probability.cpp:290   same, kappa = -d/r, Lambda = L
//   linear_probability_d         src/linear_probability.cpp:21    linear_d_value
//   linear_probability_r         src/linear_probability.cpp:170   linear_r_value
//   diagonal_probability_approx_f_eta  src/diagonal_probability.cpp:18  diagonal_value
//
// Derivations are in DESIGN.md ("Scale-free integrands").
#pragma once

#include "qmath.cuh"

namespace qb200 {

// Constants of one distribution, prepared on the host (hostconst.hpp) and
// passed to kernels by value.
struct DevConsts {
  dd kappa;       // 2D: K_sigma / 2^sigma (heuristic sigma) or -d/r (quick)
  dd c_over_L;    // ceil(2^(m+l)/r) / 2^l
  dd n_over_L;    // floor(2^(m+l)/r) / 2^l
  dd n1_over_L;   // (N + 1) / 2^l
  dd rho;         // 2^m / r
  double r_m;     // r / 2^m
  double d_m;     // d / 2^m
  double omd_m;   // 1 - d / 2^m
  double beta_m;  // (2^(l+m) mod r) / 2^m
  double rbeta_m; // (r - beta) / 2^m
  double cs;      // pi * 2^(sigma - l): s = cs * (|x_d| + |x_r|)  (src/probability.cpp:252-260)
  double e0s;     // 2^(4 - sigma) + 2^(3 - l)                     (src/probability.cpp:270-277)
  int m, l, sigma;
  int lam_exp;    // l - sigma, or l for the quick method
};

// (double)0.01f, the relative error bound of src/probability.cpp:281.
#define QB_ERROR_BOUND 0.00999999977648258209228515625

// [ sin(pi a) / (2^e sin(pi b / 2^e)) ]^2. For e > 200 the inner sine is its
// argument to far below double precision (|b| < 2^60 is enforced by the host).
QHD double ratio_sinpi_sq(dd a, dd b, int e) {
  double den;
  if (e > 200) {
    den = QB_PI_HI * (b.hi + b.lo);
  } else {
    const dd w = dd_mul_pow2(b, pow2i(-e));
    den = sinpi_dd(w) * pow2i(e);
  }
  const double q = sinpi_dd(a) / den;
  return q * q;
}

// First factor of probability_approx (src/probability.cpp:165-214), divided by
// Lambda^2:  T1 = [ sin(pi u) / (Lambda sin(pi u / Lambda)) ]^2,
// u = x_d + kappa x_r  (phi / 2 = pi u / Lambda).  y = kappa * x_r.
QHD double t1_value(dd xd, dd y, int lam_exp) {
  const dd u = dd_add(xd, y);
  const double uv = u.hi + u.lo;
  if (fabs(uv) < 1e-30) return 1.0;  // phi == 0 branch (:181-182) and its neighbourhood
  const double t = ratio_sinpi_sq(u, u, lam_exp);
  if (!(t == t) || t > 1e300) return 1.0;  // u / Lambda an exact non-zero integer
  return t;
}

// Second factor (src/probability.cpp:216-242), divided by L^2:
// T2 = [ sin(pi x_r C / L) / (L sin(pi x_r / L)) ]^2.
QHD double t2_value(dd xr, dd c_over_L, int l) {
  const dd v = dd_mul(xr, c_over_L);
  return ratio_sinpi_sq(v, xr, l);
}

// linear_probability_d (src/linear_probability.cpp:85-161) times 2^(m+l):
//   A / (pi x sinc(e))^2,  e = pi x / L,
//   A = (1 - delta) sin^2(pi x) + delta [ (1 - sinc(2 pi x)) + sinc(2 pi x) (1 - e cot e) ].
QHD double linear_d_value(dd x, double d_m, double omd_m, int l) {
  double s1, c1;
  sincospi_dd(x, &s1, &c1);
  const double oms = one_minus_sinc_2pi(x);
  double corr = 0.0, den;
  if (l > 200) {
    den = QB_PI_HI * (x.hi + x.lo);
  } else {
    const dd w = dd_mul_pow2(x, pow2i(-l));
    corr = (1.0 - oms) * one_minus_ecote_pi(w);
    den = sinpi_dd(w) * pow2i(l);
  }
  const double A = fma(omd_m, s1 * s1, d_m * (oms + corr));
  return A / (den * den);
}

// linear_probability_r (src/linear_probability.cpp:210-255) times 2^(2m):
//   beta/2^m [sin(pi x (N+1)/L) / (L sin e)]^2 + (r-beta)/2^m [sin(pi x N/L) / (L sin e)]^2.
QHD double linear_r_value(dd x, const DevConsts& k) {
  const dd a1 = dd_mul(x, k.n1_over_L);
  const dd a0 = dd_mul(x, k.n_over_L);
  double den;
  if (k.l > 200) {
    den = QB_PI_HI * (x.hi + x.lo);
  } else {
    const dd w = dd_mul_pow2(x, pow2i(-k.l));
    den = sinpi_dd(w) * pow2i(k.l);
  }
  const double qa = sinpi_dd(a1) / den;
  const double qb = sinpi_dd(a0) / den;
  return fma(k.beta_m, qa * qa, k.rbeta_m * (qb * qb));
}

// diagonal_probability_approx_f_eta (src/diagonal_probability.cpp:24-96) times
// 2^m:  rho sinc^2(pi (x - eta 2^sigma) rho),  rho = 2^m / r.
QHD double diagonal_value(dd x, double eta_shift, dd rho) {
  const dd z = dd_mul(dd_add_d(x, -eta_shift), rho);
  const double sc = sincpi_dd(z);
  return sc * sc * rho.hi;
}

// ---------------------------------------------------------------------------
// Grid abscissae (src/distribution_slice_compute.cpp:196-243): interleaved main
// points 2^(|k| + i/D) and arithmetic means of neighbours, rounded to integers,
// signed, then divided by 2^m.  g = main / mean value of 2^(i/D) in [1, 2].
// ---------------------------------------------------------------------------
QHD dd dd_round_half_away(dd a) {  // a >= 0, nearest integer, ties away from zero
  const double ih = floor(a.hi);
  const double fh = a.hi - ih;  // exact, in [0, 1)
  const double il = floor(a.lo);
  const double fl = a.lo - il;  // in [0, 1]
  double f = fh + fl;           // in [0, 2]
  double extra = floor(f);
  f -= extra;
  if (f >= 0.5) extra += 1.0;
  return quick_two_sum(ih, il + extra);
}

QHD dd grid_x(dd g, int k_abs, int sign, int m) {
  dd a;
  if (k_abs < 100) {
    // alpha has fractional bits at the reference's 192-bit precision: round.
    a = dd_mul_pow2(g, pow2i(k_abs));
    a = dd_round_half_away(a);
    a = dd_mul_pow2(a, pow2i(-k_abs));
  } else {
    a = g;
  }
  const double sc = pow2i(k_abs - m);
  a = dd_mul_pow2(a, sign < 0 ? -sc : sc);
  return a;
}

}  // namespace qb200
