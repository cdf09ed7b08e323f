// qmath.cuh -- double-double helpers and pi-scaled trigonometry.
//
// Everything here is __host__ __device__ so that tests/hostsim (a test-only CPU
// twin of the device functions) can exercise exactly the code the kernels run.
// The shipped library only ever calls these from device code.
//
// Why this exists: the reference evaluates its integrands at 192..6144 bits of
// MPFR precision (src/common.h:13, src/linear_probability.cpp:26,
// src/diagonal_probability.cpp:38-39) because it forms theta = 2 pi alpha /
// 2^(l+m) ~ 1e-600 and differences of cosines. In the scale-free variable
// x = alpha / 2^m every trigonometric argument is pi * (an O(2^+-40) number), so
// a double-double argument reduction modulo 1/2 followed by FP64 polynomials
// reproduces the reference to ~1e-15 relative.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QHD __host__ __device__ __forceinline__
#else
#define QHD inline
#endif

namespace qb200 {

struct dd {
  double hi, lo;
};

QHD dd make_dd(double hi, double lo) {
  dd r;
  r.hi = hi;
  r.lo = lo;
  return r;
}

QHD dd two_sum(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  return make_dd(s, (a - (s - bb)) + (b - bb));
}
// Requires |a| >= |b| (or a == 0).
QHD dd quick_two_sum(double a, double b) {
  const double s = a + b;
  return make_dd(s, b - (s - a));
}
QHD dd two_prod(double a, double b) {
  const double p = a * b;
  return make_dd(p, fma(a, b, -p));
}
QHD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
QHD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo += a.lo;
  return quick_two_sum(s.hi, s.lo);
}
QHD dd dd_neg(dd a) { return make_dd(-a.hi, -a.lo); }
QHD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += fma(a.hi, b.lo, a.lo * b.hi);
  return quick_two_sum(p.hi, p.lo);
}
QHD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo = fma(a.lo, b, p.lo);
  return quick_two_sum(p.hi, p.lo);
}
// Multiply by an exact power of two given as a double (no rounding unless the
// result leaves the normal range).
QHD dd dd_mul_pow2(dd a, double p2) { return make_dd(a.hi * p2, a.lo * p2); }

// 2^e as a double for -1022 <= e <= 1023.
QHD double pow2i(int e) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)(e + 1023) << 52);
#else
  return ldexp(1.0, e);
#endif
}

#define QB_PI_HI 3.141592653589793116
#define QB_PI_LO 1.2246467991473532072e-16

// sin(pi r) and cos(pi r) for |r| <= 1/4 (a little beyond is fine). Taylor
// coefficients (-1)^k pi^(2k+1) / (2k+1)! and (-1)^k pi^(2k) / (2k)!; the
// truncation errors at |r| = 1/4 are 8e-20 and 3e-20.
QHD double sinpi_kernel(double r) {
  const double z = r * r;
  double p = 7.9520540014755127848e-7;          // k = 8
  p = fma(p, z, -2.1915353447830215827e-5);     // k = 7
  p = fma(p, z, 4.6630280576761256442e-4);      // k = 6
  p = fma(p, z, -7.3704309457143507773e-3);     // k = 5
  p = fma(p, z, 8.2145886611128228799e-2);      // k = 4
  p = fma(p, z, -5.9926452932079207689e-1);     // k = 3
  p = fma(p, z, 2.5501640398773454439);         // k = 2
  p = fma(p, z, -5.1677127800499700292);        // k = 1
  // r * (pi_hi + pi_lo + z * p)
  return fma(r, QB_PI_HI, r * fma(z, p, QB_PI_LO));
}
QHD double cospi_kernel(double r) {
  const double z = r * r;
  double p = 3.604730797462500934e-09;
  p = fma(p, z, -1.387895246221377211e-07);
  p = fma(p, z, 4.303069587032947007e-06);
  p = fma(p, z, -1.046381049248457071e-04);
  p = fma(p, z, 1.929574309403923048e-03);
  p = fma(p, z, -2.580689139001406001e-02);
  p = fma(p, z, 2.353306303588932045e-01);
  p = fma(p, z, -1.335262768854589496e+00);
  p = fma(p, z, 4.058712126416768218e+00);
  p = fma(p, z, -4.934802200544679309e+00);
  return fma(z, p, 1.0);
}

// sin(pi x), cos(pi x) for a double-double x, |x| < 2^51. The reduction
// x - n/2 is exact, so the results are accurate to about one ulp RELATIVE, also
// next to the zeros of either function (as far as x itself is accurate).
QHD void sincospi_dd(dd x, double* s, double* c) {
  const double n = rint(2.0 * x.hi);
  const double rh0 = fma(-0.5, n, x.hi);  // exact
  const dd r = two_sum(rh0, x.lo);
  const double sk = sinpi_kernel(r.hi);
  const double ck = cospi_kernel(r.hi);
  const double t = QB_PI_HI * r.lo;
  const double S = fma(t, ck, sk);
  const double Cc = fma(-t, sk, ck);
  const long long q = (long long)n & 3;
  if (q == 0) {
    *s = S;
    *c = Cc;
  } else if (q == 1) {
    *s = Cc;
    *c = -S;
  } else if (q == 2) {
    *s = -S;
    *c = -Cc;
  } else {
    *s = -Cc;
    *c = S;
  }
}
QHD double sinpi_dd(dd x) {
  double s, c;
  sincospi_dd(x, &s, &c);
  return s;
}

// sin(pi r) / (pi r) for |r| <= 1/4: sum (-1)^k pi^(2k) / (2k+1)! r^(2k).
QHD double sincpi_small(double r) {
  const double z = r * r;
  double p = -7.3047118222177747971e-9;         // k = 9
  p = fma(p, z, 2.5312174041370276514e-7);      // k = 8
  p = fma(p, z, -6.9758736616563804745e-6);     // k = 7
  p = fma(p, z, 1.4842879303107100368e-4);      // k = 6
  p = fma(p, z, -2.3460810354558236375e-3);     // k = 5
  p = fma(p, z, 2.6147847817654800505e-2);      // k = 4
  p = fma(p, z, -1.907518241220842137e-1);      // k = 3
  p = fma(p, z, 8.1174242528335364364e-1);      // k = 2
  p = fma(p, z, -1.6449340668482264365);        // k = 1
  return fma(z, p, 1.0);
}

// sin(pi u) / (pi u) for a double-double u (1 at u = 0).
QHD double sincpi_dd(dd u) {
  const double uv = u.hi + u.lo;
  if (fabs(uv) <= 0.25) return sincpi_small(uv);
  return sinpi_dd(u) / (QB_PI_HI * uv);
}

// 1 - sin(y)/y with y = 2 pi x, for a double-double x: series
// sum_{k>=1} (-1)^(k+1) (2 pi)^(2k) / (2k+1)! x^(2k) below |x| = 1/8, and
// 1 - sin(pi x) cos(pi x) / (pi x) above (no cancellation there).
QHD double one_minus_sinc_2pi(dd x) {
  const double xv = x.hi + x.lo;
  if (fabs(xv) <= 0.125) {
    const double z = xv * xv;
    double p = 1.9148863759234563564e-3;        // k = 9
    p = fma(p, z, -1.6588586379752424416e-2);   // k = 8
    p = fma(p, z, 1.1429271407257813769e-1);    // k = 7
    p = fma(p, z, -6.0796433625526683109e-1);   // k = 6
    p = fma(p, z, 2.4023869803067634048);       // k = 5
    p = fma(p, z, -6.6938490413196289292);      // k = 4
    p = fma(p, z, 1.2208116743813389677e+1);    // k = 3
    p = fma(p, z, -1.2987878804533658298e+1);   // k = 2
    p = fma(p, z, 6.5797362673929057459);       // k = 1
    return z * p;
  }
  double s, c;
  sincospi_dd(x, &s, &c);
  return 1.0 - (s * c) / (QB_PI_HI * xv);
}

// 1 - e cot(e) for e = pi * w, w double-double with |w| < 1/2: the series
// sum_{k>=1} 2^(2k) |B_2k| / (2k)! e^(2k) below |e| = 1/2, direct above.
QHD double one_minus_ecote_pi(dd w) {
  const double wv = w.hi + w.lo;
  const double e = QB_PI_HI * wv;
  if (fabs(e) <= 0.5) {
    const double z = e * e;
    double p = 2.3411706819824883959e-12;       // k = 12
    p = fma(p, z, 2.3106432599002624097e-11);   // k = 11
    p = fma(p, z, 2.2805151204592182866e-10);   // k = 10
    p = fma(p, z, 2.2507846516808992854e-9);    // k = 9
    p = fma(p, z, 2.2214608789979679076e-8);    // k = 8
    p = fma(p, z, 2.19259478518737778e-7);      // k = 7
    p = fma(p, z, 2.1644042808063972085e-6);    // k = 6
    p = fma(p, z, 2.1377799155576933355e-5);    // k = 5
    p = fma(p, z, 2.1164021164021164021e-4);    // k = 4
    p = fma(p, z, 2.1164021164021164021e-3);    // k = 3
    p = fma(p, z, 2.2222222222222222222e-2);    // k = 2
    p = fma(p, z, 3.3333333333333333333e-1);    // k = 1
    return z * p;
  }
  double s, c;
  sincospi_dd(w, &s, &c);
  return 1.0 - e * c / s;
}

}  // namespace qb200
