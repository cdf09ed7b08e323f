// textfmt.cuh -- "%.24Lg\n" of an x87 long double, byte for byte as glibc prints it.
//
// Replaces the per-cell fprintf(file, "%.24Lg\n", ...) of the reference's slice
// exporters (src/distribution_slice_import_export.cpp:89-103,
// src/linear_distribution_slice_import_export.cpp:82-97,
// src/diagonal_distribution_slice_import_export.cpp:87-103): 1.05e8 calls and
// 3.1 GB of text for one m = 2048 two-dimensional distribution, 87 % of the
// wall clock once the integration runs on the GPU.
//
// Semantics restated (ISO C %g, precision P = 24, glibc rounding = exact value,
// round-half-even): let X be the decimal exponent after rounding to 24
// significant digits; if -4 <= X < 24 the fixed style is used, else d.ddde+XX
// with at least two exponent digits; trailing zeros and a trailing point are
// removed; "inf" / "nan" with sign; zero prints as "0" / "-0".
//
// Method. value = M * 2^q with a normalised 64-bit M. With X1 = floor((e2+1)
// log10 2) >= X and the 192-bit table entry T of 10^(-1-X1) (truncated),
// F = M * T is a 320-bit binary FRACTION in [0.01, 1); the 24 digits are the
// integer parts of three multiplications by 10^8, the remainder decides the
// rounding. For 0 <= -1-X1 <= 82 the table entry is exact and so is every step
// (values in [1e-83, 1)). Otherwise the remainder is off by < 2^213 / 2^320,
// one-sided; only a remainder inside that band below one half is undecided, and
// is then resolved by exact integer arithmetic (exact_round_up: compares
// M * 5^k * 2^s with 2 * digits + 1). Everything is __host__ __device__ so the
// very same code is checked against snprintf on a GPU-less machine (tests/hostsim).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#define QT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define QT_HD inline
#define QT_HD_NOINLINE inline
#endif

namespace qb200 {
namespace text {

constexpr int K_MIN = -5000;       // powers of ten tabulated: 10^K_MIN .. 10^K_MAX
constexpr int K_MAX = 4970;        // (formatter: -4934..4951; parser: -4992..4940)
constexpr int K_EXACT_MAX = 82;    // 5^82 < 2^191: 10^k is exact in 192 bits for 0 <= k <= 82
constexpr int MAX_TEXT = 33;       // "-d.ddddddddddddddddddddddde-4932\n"
constexpr int BIG_LIMBS = 376;     // exact path: (96 + log2(5) * 5000) / 32 = 366 limbs

// 10^k ~ T * 2^(e2 - 191), T = w[5]..w[0] (little-endian 32-bit limbs), 2^191 <= T < 2^192,
// T = floor(true): exact iff 0 <= k <= K_EXACT_MAX.
struct Pow10Entry {
  uint32_t w[6];
  int32_t e2;
  uint32_t exact;
};

struct Dec24 {
  uint32_t c0, c1, c2;  // 3 x 8 decimal digits, most significant group first
  int32_t x;            // decimal exponent of the first digit
  uint32_t up;        // fast-path rounding decision (final unless undecided)
  uint32_t undecided;   // 1: remainder inside the error band below one half
};

QT_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}

QT_HD uint32_t mulhi_lo(uint32_t a, uint32_t b, uint32_t carry, uint32_t* hi) {
  const uint64_t t = (uint64_t)a * b + carry;
  *hi = (uint32_t)(t >> 32);
  return (uint32_t)t;
}

// floor(n * log10(2)) for |n| <= 16500 (checked exhaustively in tests/test_text_format.py).
QT_HD int32_t floor_log10_pow2(int32_t n) {
  return (int32_t)(((int64_t)n * 1292913987LL) >> 32);
}

// F (10 limbs, a fraction of 2^320) *= c; returns the integer part.
QT_HD uint32_t frac_mul(uint32_t (&F)[10], uint32_t c) {
  uint32_t carry = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) F[i] = mulhi_lo(F[i], c, carry, &carry);
  return carry;
}

// Four decimal digits of h < 10^4 as ASCII, first digit in the lowest byte.
QT_HD uint32_t ascii4(uint32_t h) {
  const uint32_t q = (h * 5243u) >> 19;          // h / 100
  const uint32_t y = q | ((h - q * 100u) << 16);  // two 2-digit numbers in 16-bit lanes
  const uint32_t tens = ((y * 103u) >> 10) & 0x000F000Fu;
  return tens | ((y - tens * 10u) << 8) | 0x30303030u;
}

// Eight decimal digits of c < 10^8 as ASCII, first digit in the lowest byte.
QT_HD uint64_t ascii8(uint32_t c) {
  const uint32_t hi = c / 10000u, lo = c - hi * 10000u;
  return (uint64_t)ascii4(hi) | ((uint64_t)ascii4(lo) << 32);
}

// Number of trailing '0' characters of an ascii8 group (8 if all).
QT_HD int trailing_zeros8(uint64_t a) {
  const uint64_t d = a ^ 0x3030303030303030ULL;
  return d ? (clz64(d) >> 3) : 8;
}

// Sign of (z0 * 5^n  -  other * 2^t) by exact integer arithmetic; z0, other < 2^96
// (three limbs, little-endian), 0 <= n <= 5000, any t. scratch: BIG_LIMBS words.
QT_HD_NOINLINE int cmp_pow5(const uint32_t z0[3], int n, const uint32_t other[3], int t,
                            uint32_t* Z) {
  int len = 3;
  Z[0] = z0[0];
  Z[1] = z0[1];
  Z[2] = z0[2];
  while (len > 0 && Z[len - 1] == 0) len--;
  while (n > 0) {
    const int step = n >= 13 ? 13 : n;
    uint32_t mulc = 1;
    for (int i = 0; i < step; i++) mulc *= 5u;  // 5^13 < 2^32
    uint32_t carry = 0;
    for (int i = 0; i < len; i++) Z[i] = mulhi_lo(Z[i], mulc, carry, &carry);
    if (carry) Z[len++] = carry;
    n -= step;
  }
  // bit lengths
  int olen = 3;
  while (olen > 0 && other[olen - 1] == 0) olen--;
  if (len == 0 || olen == 0) return (len == 0 ? 0 : 1) - (olen == 0 ? 0 : 1);
  const long zbits = 32L * (len - 1) + (32 - (clz64(Z[len - 1]) - 32));
  const long obits = 32L * (olen - 1) + (32 - (clz64(other[olen - 1]) - 32)) + t;
  if (zbits != obits) return zbits > obits ? 1 : -1;
  // Same bit length: compare bit by bit from the top over the extent of `other`
  // (96 bits), then any lower bit of Z decides.
  // bit j of (other << t) = bit (j - t) of other.
  for (long j = zbits - 1; j >= 0; j--) {
    const uint32_t zb = (Z[j >> 5] >> (j & 31)) & 1u;
    const long oj = j - t;
    const uint32_t ob = (oj >= 0 && oj < 96) ? ((other[oj >> 5] >> (oj & 31)) & 1u) : 0u;
    if (zb != ob) return zb ? 1 : -1;
    if (oj < 0) {
      // below the extent of `other`: only Z can have bits; scan whole words
      for (long w = j >> 5; w >= 0; w--) {
        uint32_t v = Z[w];
        if (w == (j >> 5) && (j & 31) != 31) v &= (2u << (j & 31)) - 1u;
        if (v) return 1;
      }
      return 0;
    }
  }
  // all integer bits equal; fractional bits of other * 2^t (t < 0) make it larger
  if (t < 0) {
    const long nb = -t < 96 ? -t : 96;
    for (long b = 0; b < nb; b++)
      if ((other[b >> 5] >> (b & 31)) & 1u) return -1;
  }
  return 0;
}

// Exact rounding decision: is M * 2^q * 10^(23 - x) above digits + 1/2 (or equal
// with odd digits)? digits = c0 * 10^16 + c1 * 10^8 + c2 before rounding.
QT_HD_NOINLINE bool exact_round_up(uint64_t Mn, int qn, int x, uint32_t c0, uint32_t c1,
                                   uint32_t c2, uint32_t* scratch) {
  // b = 2 * digits + 1 < 2^82
  uint32_t b[3] = {c0, 0, 0};
  uint32_t carry = 0;
  for (int r = 0; r < 2; r++) {
    carry = r == 0 ? c1 : c2;
    for (int i = 0; i < 3; i++) b[i] = mulhi_lo(b[i], 100000000u, carry, &carry);
  }
  carry = 1;
  for (int i = 0; i < 3; i++) {
    const uint64_t t = ((uint64_t)b[i] << 1) + carry;
    b[i] = (uint32_t)t;
    carry = (uint32_t)(t >> 32);
  }
  const uint32_t a[3] = {(uint32_t)Mn, (uint32_t)(Mn >> 32), 0};
  const int k = 23 - x;
  int sgn;
  if (k >= 0) {
    // 2W = Mn * 5^k * 2^(qn + k + 1)  vs  b      <=>  Mn * 5^k  vs  b * 2^-(qn + k + 1)
    sgn = cmp_pow5(a, k, b, -(qn + k + 1), scratch);
  } else {
    // 2W = Mn * 2^(qn + k + 1) / 5^|k|  vs  b    <=>  Mn * 2^(qn + k + 1)  vs  b * 5^|k|
    sgn = -cmp_pow5(b, -k, a, qn + k + 1, scratch);
  }
  if (sgn != 0) return sgn > 0;
  return (c2 & 1u) != 0;  // tie: to even
}

// 24 significant digits of Mn * 2^qn (Mn normalised: bit 63 set).
// `force_band`: test hook, sends every value through the exact decision.
QT_HD void digits24(uint64_t Mn, int qn, const Pow10Entry* __restrict__ tab, Dec24* out,
                    bool force_band = false) {
  const int e2 = qn + 63;
  const int x1 = floor_log10_pow2(e2 + 1);
  const int k = -1 - x1;
  const Pow10Entry* ent = tab + (k - K_MIN);
  uint32_t T[6];
  int te2;
#if defined(__CUDA_ARCH__)
  {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(ent));
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(ent) + 1);
    T[0] = v0.x; T[1] = v0.y; T[2] = v0.z; T[3] = v0.w; T[4] = v1.x; T[5] = v1.y;
    te2 = (int)v1.z;
  }
#else
  for (int i = 0; i < 6; i++) T[i] = ent->w[i];
  te2 = ent->e2;
#endif
  // P = Mn * T (256 bits, exact)
  uint32_t P[8];
  {
    const uint32_t m0 = (uint32_t)Mn, m1 = (uint32_t)(Mn >> 32);
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) P[i] = mulhi_lo(T[i], m0, carry, &carry);
    P[6] = carry;
    carry = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const uint64_t t = (uint64_t)T[i] * m1 + P[i + 1] + carry;
      P[i + 1] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
    P[7] = carry;
  }
  // F = P * 2^L as a 320-bit fraction, L = 58..65: word shift 1 or 2, bit shift L & 31
  const int L = qn + te2 + 129;
  const int bs = L & 31;
  uint32_t S[9];
  S[0] = P[0] << bs;
#pragma unroll
  for (int i = 1; i < 8; i++) S[i] = bs ? ((P[i] << bs) | (P[i - 1] >> (32 - bs))) : P[i];
  S[8] = bs ? (P[7] >> (32 - bs)) : 0u;
  uint32_t F[10];
  const bool two = (L >> 5) == 2;
  F[0] = 0;
  F[1] = two ? 0u : S[0];
#pragma unroll
  for (int i = 2; i < 10; i++) F[i] = two ? S[i - 2] : S[i - 1];

  int x = x1;
  // F < 1/10 ?  (1/10 = 0.1999...9A hex, rounded up)
  bool small = F[9] < 0x19999999u;
  if (F[9] == 0x19999999u) {
    bool decided = false;
    small = false;
#pragma unroll
    for (int i = 8; i >= 1; i--) {
      if (!decided && F[i] != 0x99999999u) {
        small = F[i] < 0x99999999u;
        decided = true;
      }
    }
    if (!decided) small = F[0] < 0x9999999Au;
  }
  if (small) {
    frac_mul(F, 10u);
    x = x1 - 1;
  }
  out->c0 = frac_mul(F, 100000000u);
  out->c1 = frac_mul(F, 100000000u);
  out->c2 = frac_mul(F, 100000000u);
  out->x = x;
  // remainder against one half
  const bool top = (F[9] >> 31) != 0;
  uint32_t rest = F[9] & 0x7fffffffu;
#pragma unroll
  for (int i = 0; i < 9; i++) rest |= F[i];
  const bool exact = (k >= 0 && k <= K_EXACT_MAX);
  bool up, band = false;
  if (exact) {
    up = top && (rest != 0 || (out->c2 & 1u));
  } else {
    // the table entry is a strict truncation: true remainder in (F, F + 2^213)
    up = top;
    band = (F[9] == 0x7fffffffu) && (F[8] == 0xffffffffu) && (F[7] == 0xffffffffu) &&
           ((F[6] >> 21) == 0x7ffu);
  }
  out->up = up ? 1u : 0u;
  out->undecided = (band || force_band) ? 1u : 0u;
}

// Apply the rounding decision; may carry into the exponent.
QT_HD void round_digits(Dec24* d, bool up) {
  if (!up) return;
  if (++d->c2 == 100000000u) {
    d->c2 = 0;
    if (++d->c1 == 100000000u) {
      d->c1 = 0;
      if (++d->c0 == 100000000u) {
        d->c0 = 10000000u;
        d->x += 1;
      }
    }
  }
}

// What to print.
struct Piece {
  uint64_t a0, a1, a2;  // 24 ASCII digits
  int32_t x;
  int32_t ndig;         // significant digits kept (1..24), 0 for the special forms
  uint32_t neg;
  uint32_t special;     // 0 number, 1 zero, 2 inf, 3 nan
};

QT_HD int piece_length(const Piece& p) {
  int len = (int)p.neg + 1;  // sign, newline
  if (p.special) return len + (p.special == 1 ? 1 : 3);
  const int x = p.x;
  if (x < -4 || x >= 24) {
    const int ax = x < 0 ? -x : x;
    len += p.ndig + (p.ndig > 1 ? 1 : 0) + 2 + (ax >= 1000 ? 4 : ax >= 100 ? 3 : 2);
  } else if (x >= 0) {
    len += p.ndig > x + 1 ? p.ndig + 1 : x + 1;
  } else {
    len += 1 - x + p.ndig;  // "0." + (-x - 1) zeros + digits
  }
  return len;
}

QT_HD char piece_digit(const Piece& p, int j) {
  const uint64_t a = j < 8 ? p.a0 : (j < 16 ? p.a1 : p.a2);
  const uint32_t w = (j & 4) ? (uint32_t)(a >> 32) : (uint32_t)a;
  return (char)(w >> (8 * (j & 3)));
}

// "e+XX\n" / "e-XXXX\n" at out[pos...]
template <class Ptr>
QT_HD void render_exponent(Ptr out, int pos, int x) {
  const uint32_t ax = (uint32_t)(x < 0 ? -x : x);  // < 5000
  out[pos++] = 'e';
  out[pos++] = x < 0 ? '-' : '+';
  const uint32_t hi2 = (ax * 5243u) >> 19, lo2 = ax - hi2 * 100u;
  const uint32_t t1 = (hi2 * 103u) >> 10, t0 = (lo2 * 103u) >> 10;
  if (ax >= 1000u) out[pos++] = (char)('0' + t1);
  if (ax >= 100u) out[pos++] = (char)('0' + (hi2 - t1 * 10u));
  out[pos++] = (char)('0' + t0);
  out[pos++] = (char)('0' + (lo2 - t0 * 10u));
  out[pos] = '\n';
}

// Writes piece_length(p) characters to out.
template <class Ptr>
QT_HD void piece_render(const Piece& p, Ptr out) {
  int pos = 0;
  if (p.neg) out[pos++] = '-';
  const int x = p.x;
  const bool estyle = !p.special && (x < -4 || x >= 24);
#if defined(__CUDA_ARCH__)
  // warp-uniform choice: no divergence between the two digit loops
  const bool fast = __all_sync(__activemask(), estyle) != 0;
#else
  const bool fast = estyle;
#endif
  if (fast) {
    // d[.ddd]e+XX with every digit offset a compile-time constant
    const int nd = p.ndig;
    out[pos] = piece_digit(p, 0);
    if (nd > 1) out[pos + 1] = '.';
#pragma unroll
    for (int j = 1; j < 24; j++)
      if (j < nd) out[pos + 1 + j] = piece_digit(p, j);
    render_exponent(out, pos + nd + (nd > 1 ? 1 : 0), x);
    return;
  }
  if (p.special) {
    if (p.special == 1) {
      out[pos++] = '0';
    } else {
      const bool inf = p.special == 2;
      out[pos++] = inf ? 'i' : 'n';
      out[pos++] = inf ? 'n' : 'a';
      out[pos++] = inf ? 'f' : 'n';
    }
    out[pos] = '\n';
    return;
  }
  int nd = p.ndig, point = 1;
  if (!estyle) {
    if (x >= 0) {
      point = x + 1;
      if (nd < point) nd = point;
    } else {
      out[pos++] = '0';
      out[pos++] = '.';
#pragma unroll
      for (int z = 0; z < 3; z++)
        if (z < -x - 1) out[pos++] = '0';
      point = 64;  // no point among the digits
    }
  }
  const bool has_point = nd > point;
  // digits before the point at pos + j, after it at pos + 1 + j: two predicated stores
  // with compile-time offsets per digit
  const int before = nd < point ? nd : point;
  const unsigned after = has_point ? (unsigned)(nd - point) : 0u;
#pragma unroll
  for (int j = 0; j < 24; j++) {
    const char c = piece_digit(p, j);
    if (j < before) out[pos + j] = c;
    if ((unsigned)(j - point) < after) out[pos + 1 + j] = c;
  }
  if (has_point) out[pos + point] = '.';
  pos += nd + (has_point ? 1 : 0);
  if (estyle)
    render_exponent(out, pos, x);
  else
    out[pos] = '\n';
}

// Classify an x87 extended value (mant with explicit integer bit, se = sign | exponent).
// Returns false for the special forms (piece filled in), true for a finite non-zero
// number with *Mn, *qn set (value = Mn * 2^qn, Mn normalised).
QT_HD bool classify_x87(uint64_t mant, uint32_t se, Piece* p, uint64_t* Mn, int* qn) {
  p->neg = (se >> 15) & 1u;
  p->ndig = 0;
  p->x = 0;
  p->a0 = p->a1 = p->a2 = 0;
  const int E = (int)(se & 0x7fffu);
  if (E == 0x7fff) {
    p->special = (mant << 1) == 0 ? 2u : 3u;
    return false;
  }
  if (mant == 0) {
    p->special = 1;
    return false;
  }
  p->special = 0;
  const int lz = clz64(mant);
  *Mn = mant << lz;
  *qn = (E ? E : 1) - 16383 - 63 - lz;
  return true;
}

// Same for an IEEE double (bit pattern v), which a slice cell is before the binding
// widens it (include/qunundrum_b200.h: cells are doubles).
QT_HD bool classify_f64(uint64_t v, Piece* p, uint64_t* Mn, int* qn) {
  p->neg = (uint32_t)(v >> 63);
  p->ndig = 0;
  p->x = 0;
  p->a0 = p->a1 = p->a2 = 0;
  const int E = (int)((v >> 52) & 0x7ffu);
  const uint64_t frac = v & ((1ULL << 52) - 1);
  if (E == 0x7ff) {
    p->special = frac == 0 ? 2u : 3u;
    return false;
  }
  if (E == 0 && frac == 0) {
    p->special = 1;
    return false;
  }
  p->special = 0;
  const uint64_t M = E ? (frac | (1ULL << 52)) : frac;
  const int lz = clz64(M);
  *Mn = M << lz;
  *qn = (E ? E : 1) - 1075 - lz;
  return true;
}

QT_HD void piece_from_digits(const Dec24& d, Piece* p) {
  p->a0 = ascii8(d.c0);
  p->a1 = ascii8(d.c1);
  p->a2 = ascii8(d.c2);
  p->x = d.x;
  int tz;
  if (d.c2)
    tz = trailing_zeros8(p->a2);
  else if (d.c1)
    tz = 8 + trailing_zeros8(p->a1);
  else
    tz = 16 + trailing_zeros8(p->a0);
  p->ndig = 24 - tz;
}

}  // namespace text
}  // namespace qb200
