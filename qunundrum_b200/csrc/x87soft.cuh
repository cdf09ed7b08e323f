// x87soft.cuh -- the x87 extended-precision operations the reference's samplers perform,
// emulated bit for bit in integer arithmetic.
//
// The reference walks a stored distribution with `long double` arithmetic
//   pivot = value / (2^64 - 1)            random_generate_pivot_inclusive  src/random.c:116-135
//   pivot *= total_probability            src/distribution.cpp:373-381, src/distribution_slice.cpp:183
//   pivot -= slice->total_probability     src/distribution.cpp:384-401
//   pivot -= slice->norm_matrix[i]        src/distribution_slice.cpp:191-223
// and stops at the first non-positive pivot. Each operation rounds to a 64-bit mantissa
// (round to nearest, ties to even). A GPU has no such type; the fast path of the sampler works
// on exact double-double images of these numbers and proves, with a rounding-error band, which
// element the sequential walk stops at; the rare pivot that falls inside the band (probability
// ~1e-10 per sample) is decided by replaying the walk with the functions below.
//
// A number is sign * mant * 2^(exp - 63) with bit 63 of mant set, or zero (mant == 0).
// Denormals, infinities and NaNs are not represented: x87_decode() reports them and the
// sampler refuses a distribution that contains one (probabilities below 2^-16382 do not occur).
//
// __host__ __device__ so that tests/hostsim can check every operation against the host's own
// long double arithmetic.
#pragma once

#include <string.h>

#include "qmath.cuh"

namespace qb200 {

struct X87 {
  uint64_t mant;
  int32_t exp;
  int32_t neg;
};

QHD X87 x87_zero() {
  X87 r;
  r.mant = 0;
  r.exp = 0;
  r.neg = 0;
  return r;
}

QHD bool x87_is_zero(X87 a) { return a.mant == 0; }
// a <= 0
QHD bool x87_nonpositive(X87 a) { return a.mant == 0 || a.neg; }
QHD X87 x87_neg(X87 a) {
  if (a.mant) a.neg = !a.neg;
  return a;
}

QHD int qb_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

QHD void qb_mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
  *hi = __umul64hi(a, b);
  *lo = a * b;
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *hi = (uint64_t)(p >> 64);
  *lo = (uint64_t)p;
#endif
}

// The 16 bytes of an x86-64 long double: 64-bit mantissa (explicit integer bit), then sign and
// 15-bit exponent. Returns false for denormals / pseudo-denormals / infinities / NaNs.
QHD bool x87_decode(uint64_t mant, uint32_t sign_exp, X87* out) {
  const uint32_t e = sign_exp & 0x7fffu;
  out->neg = (sign_exp >> 15) & 1u;
  if (e == 0) {
    out->mant = 0;
    out->exp = 0;
    out->neg = 0;
    return mant == 0;  // +-0 is fine, denormals are not
  }
  if (e == 0x7fffu || !(mant >> 63)) return false;
  out->mant = mant;
  out->exp = (int32_t)e - 16383;
  return true;
}

// (long double)value / (long double)(2^64 - 1), rounded to nearest: the quotient is
// value 2^-64 (1 + 2^-64 + ...), i.e. the normalised mantissa of value plus a fraction in
// (1/2, 1) of its last place -- always one step up.
QHD X87 x87_pivot_inclusive(uint64_t value) {
  X87 r = x87_zero();
  if (value == 0) return r;
  const int lz = qb_clz64(value);
  r.mant = value << lz;
  r.exp = -1 - lz;
  r.mant += 1;
  if (r.mant == 0) {
    r.mant = 0x8000000000000000ull;
    r.exp += 1;
  }
  return r;
}

// Round the 128-bit product / 192-bit sum tail: `rest_hi` holds the bits right below the
// mantissa (its top bit is the rounding bit), `sticky` any bit below that.
QHD void x87_round(X87* r, uint64_t rest_hi, bool sticky) {
  const bool round_bit = (rest_hi >> 63) != 0;
  const bool below = ((rest_hi << 1) != 0) || sticky;
  if (round_bit && (below || (r->mant & 1ull))) {
    r->mant += 1;
    if (r->mant == 0) {
      r->mant = 0x8000000000000000ull;
      r->exp += 1;
    }
  }
}

// RN64(a * b)
QHD X87 x87_mul(X87 a, X87 b) {
  if (a.mant == 0 || b.mant == 0) return x87_zero();
  uint64_t hi, lo;
  qb_mul64(a.mant, b.mant, &hi, &lo);
  X87 r;
  r.neg = a.neg ^ b.neg;
  r.exp = a.exp + b.exp + 1;
  if (!(hi >> 63)) {
    hi = (hi << 1) | (lo >> 63);
    lo <<= 1;
    r.exp -= 1;
  }
  r.mant = hi;
  x87_round(&r, lo, false);
  return r;
}

// RN64(a + b), any signs.
QHD X87 x87_add(X87 a, X87 b) {
  if (a.mant == 0) return b;
  if (b.mant == 0) return a;
  // |a| >= |b|
  if (b.exp > a.exp || (b.exp == a.exp && b.mant > a.mant)) {
    const X87 t = a;
    a = b;
    b = t;
  }
  const int d = a.exp - b.exp;
  if (d > 66) return a;  // |b| is below a quarter of a's last place, also right under a power of two
  // 192-bit fixed point: a.mant in the top limb; b.mant shifted right by d (no bit is lost: 64 + 66 <= 192)
  uint64_t a2 = a.mant, a1 = 0, a0 = 0;
  uint64_t b2, b1, b0;
  if (d == 0) {
    b2 = b.mant; b1 = 0; b0 = 0;
  } else if (d < 64) {
    b2 = b.mant >> d; b1 = b.mant << (64 - d); b0 = 0;
  } else if (d == 64) {
    b2 = 0; b1 = b.mant; b0 = 0;
  } else {
    b2 = 0; b1 = b.mant >> (d - 64); b0 = b.mant << (128 - d);
  }
  X87 r;
  r.neg = a.neg;
  r.exp = a.exp;
  uint64_t r2, r1, r0;
  if (a.neg == b.neg) {
    r0 = a0 + b0;
    uint64_t c = r0 < a0;
    r1 = a1 + b1;
    uint64_t c1 = r1 < a1;
    r1 += c;
    c1 |= (r1 < c);
    r2 = a2 + b2;
    uint64_t c2 = r2 < a2;
    r2 += c1;
    c2 |= (r2 < c1);
    if (c2) {  // carry out: shift right by one
      const bool lost = (r0 & 1ull) != 0;
      r0 = (r0 >> 1) | (r1 << 63);
      r1 = (r1 >> 1) | (r2 << 63);
      r2 = (r2 >> 1) | 0x8000000000000000ull;
      r0 |= lost ? 1ull : 0ull;
      r.exp += 1;
    }
  } else {
    uint64_t bw = a0 < b0;
    r0 = a0 - b0;
    uint64_t t1 = a1 - b1;
    uint64_t bw1 = a1 < b1;
    r1 = t1 - bw;
    bw1 |= (t1 < bw);
    r2 = a2 - b2 - bw1;
    if ((r2 | r1 | r0) == 0) return x87_zero();
    // normalise
    int sh = 0;
    if (r2) {
      sh = qb_clz64(r2);
    } else if (r1) {
      sh = 64 + qb_clz64(r1);
    } else {
      sh = 128 + qb_clz64(r0);
    }
    r.exp -= sh;
    while (sh >= 64) {
      r2 = r1;
      r1 = r0;
      r0 = 0;
      sh -= 64;
    }
    if (sh) {
      r2 = (r2 << sh) | (r1 >> (64 - sh));
      r1 = (r1 << sh) | (r0 >> (64 - sh));
      r0 <<= sh;
    }
  }
  r.mant = r2;
  x87_round(&r, r1, r0 != 0);
  return r;
}

// (long double)v for a finite double, exact (53 <= 64 bits; denormal doubles are normalised).
QHD X87 x87_from_double(double v) {
  X87 r = x87_zero();
  uint64_t b;
#if defined(__CUDA_ARCH__)
  b = (uint64_t)__double_as_longlong(v);
#else
  memcpy(&b, &v, 8);
#endif
  const int e = (int)((b >> 52) & 0x7ff);
  uint64_t f = b & 0xfffffffffffffull;
  if (e == 0) {
    if (f == 0) return r;
    const int lz = qb_clz64(f);       // denormal: value = f 2^-1074
    r.mant = f << lz;
    r.exp = -1074 + 63 - lz;
  } else {
    r.mant = (f | (1ull << 52)) << 11;
    r.exp = e - 1023;
  }
  r.neg = (int32_t)(b >> 63);
  return r;
}

// The 16 bytes of the x86-64 long double (mantissa word, then sign and biased exponent).
// Results below the normal range (|x| < 2^-16382) are not represented by X87: *ok = false.
QHD void x87_encode(X87 a, uint64_t* mant, uint64_t* sign_exp, bool* ok) {
  if (a.mant == 0) {
    *mant = 0;
    *sign_exp = 0;
    return;
  }
  const int e = a.exp + 16383;
  if (e < 1 || e > 0x7ffe) *ok = false;
  *mant = a.mant;
  *sign_exp = (uint64_t)((uint32_t)(e & 0x7fff) | (a.neg ? 0x8000u : 0u));
}

// a / 2^k, k >= 0 (the divisions by the power-of-two `divisor` of
// src/linear_distribution.cpp:224-227: exact while the result stays normal).
QHD X87 x87_div_pow2(X87 a, int k) {
  if (a.mant) a.exp -= k;
  return a;
}

// Exact double-double image (64 <= 106 bits); magnitudes below 2^-959 flush to zero, which
// the sampler's error band covers.
QHD double qb_bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

QHD dd x87_to_dd(X87 a) {
  if (a.mant == 0) return make_dd(0.0, 0.0);
  // the top 53 bits are a double as they stand (no rounding); the low 11 bits follow 2^-52 below
  const int eh = a.exp + 1023, el = a.exp - 63 + 1023;
  if (el < 1 || eh > 2046) return make_dd(0.0, 0.0);  // below 2^-959 (or absurdly large): flushed
  const uint64_t sign = a.neg ? 0x8000000000000000ull : 0ull;
  const double hi = qb_bits_to_double(sign | ((uint64_t)eh << 52) | ((a.mant >> 11) & 0xfffffffffffffull));
  double lo = (double)(uint32_t)(a.mant & 0x7ffull) * qb_bits_to_double((uint64_t)el << 52);
  if (a.neg) lo = -lo;
  return quick_two_sum(hi, lo);
}

}  // namespace qb200
