// diagk.cuh -- k given (j, eta) for the diagonal distribution, without multi-thousand-bit floats.
//
// Replaces sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412-646) and its integrand
// diagonal_probability_approx_h (src/diagonal_probability.cpp:99-162). The reference forms, at
// 3 (m + sigma) bits of MPFR precision,
//     term = d j - (d / r) ({r j}_{2^(m+sigma)} - 2^(m+sigma) eta)                  :525-537
//     k0   = round(-term / 2^(m+sigma-l)) mod 2^l                                    :480-500
//     phi  = (2 pi / 2^(m+sigma)) {term + 2^(m+sigma-l) k}_{2^(m+sigma)}             :561-583
// for k = k0, k0 + 1, k0 - 1, ... and subtracts h(phi) = (1 - cos(2^l phi)) / (2^2l (1 - cos phi))
// from the pivot until it is used up. With r j = alpha_r + q 2^(m+sigma) (q an integer) the term
// is (d / r) 2^(m+sigma) (q + eta) EXACTLY, so everything the walk needs is the fraction
//     2^l d (q + eta) / r mod 2^l = Qv + w2 / r,       s = (q + eta) mod r,
//     Qv = (s D' + floor(s rho / r)) mod 2^l, w2 = s rho mod r   with 2^l d = D' r + rho (host constants)
// i.e. integer arithmetic on m-bit numbers: k0 = -(Qv + c) mod 2^l with c = [2 w2 >= r],
// t = w2 / r - c in [-1/2, 1/2), alpha_phi = 2^(m+sigma-l) (t + delta) and
//     h = sin^2(pi t) / (2^l sin(pi (t + delta) / 2^l))^2
// where t is needed to double-double accuracy only. The integers are exact. The walk is decided
// in doubles where the pivot clears a rigorous error band (diagk_walk) and otherwise by the exact
// walk, which subtracts h rounded to the x87 format as the reference's
// `pivot -= mpfr_get_ld(...)` does (:594), h agreeing with the reference's to ~2^-100 -- so k
// matches the reference bit for bit (a pivot used up within 2^-100 of a step, which only the
// reference's own rounding of h to 64 bits could decide, does not occur).
//
// The multi-limb part: the products r j and s rho, one Barrett division by r with a host-prepared
// reciprocal, the low l bits of s D' when k itself is wanted; 32-bit limbs, product scanning with
// 96-bit column accumulators, four columns at a time (mul_columns).
// Operands that are the same for every sample (r, d, mu) are read through `c`; per-sample
// operands live in caller-provided arrays with a stride (1 on the host; on the device the samples
// of a CTA are interleaved so that the threads of a warp read consecutive words).
//
// __host__ __device__ so that tests/hostsim can run exactly this code on the CPU.
#pragma once

#include "x87soft.cuh"

namespace qb200 {

// Zero limbs the constant operands carry below index 0 and above their last limb (mul_columns).
#define QB_DIAGK_PAD 3

struct DiagKConst {
  uint32_t m, sigma, l;
  uint32_t n;        // m + sigma
  uint32_t k;        // limbs of r (top limb non-zero)
  uint32_t wj;       // limbs of j: ceil(n / 32)
  uint32_t wl;       // limbs of k: ceil(l / 32)
  // each with QB_DIAGK_PAD zero limbs below index 0 and above its last limb (mul_columns):
  const uint32_t* r;   // k limbs
  const uint32_t* d;   // k limbs (d < r)
  const uint32_t* mu;  // k + 2 limbs: floor(2^(64 k) / r)
  const uint32_t* rho; // k limbs: 2^l d mod r
  const uint32_t* dq;  // wl limbs: floor(2^l d / r)
  const uint32_t* psi; // k + 4 + wl limbs: floor(2^(32 (k + 4)) 2^l d / r)
  dd r_top;            // the top four limbs of r as a number (limb k - 4 has weight 1)
  int force_exact;     // test switch: the exact walk only (diagk_walk)
  int full_product;    // test switch (QB200_DIAGK_FULL_PRODUCT=1): r j without the skipped columns
  int exact_fraction;  // test switch (QB200_DIAGK_EXACT_FRACTION=1): diagk_fraction_exact for every sample
};

// Words of the constants r, d, mu, rho, dq, psi with their zero limbs, in this order (the kernel's
// shared-memory copy).
QHD uint32_t diagk_const_words(const DiagKConst& c) {
  return 3 * (c.k + 2 * QB_DIAGK_PAD) + (c.k + 2 + 2 * QB_DIAGK_PAD) + (c.wl + 2 * QB_DIAGK_PAD) +
         (c.k + 4 + c.wl + 2 * QB_DIAGK_PAD);
}

// Scratch words one sample needs (times the stride).
QHD uint32_t diagk_scratch_limbs(uint32_t k) { return (2 * k + 2) + (k + 3) + (k + 2) + (k + 2); }

#define QB_DIAGK_OK 0
#define QB_DIAGK_OK_NEGATIVE_PHI 2  // success, and alpha_phi = 2^(m+sigma-l) (x - 2^l): see diagk_sample
#define QB_DIAGK_OUT_OF_BOUNDS 1  // the reference returns FALSE: pivot not used up within delta_bound
#define QB_DIAGK_GAVE_UP 4        // more than QB_DIAGK_MAX_STEPS steps (the caller's delta_bound allows more)
#define QB_DIAGK_MAX_STEPS (1ull << 22)

#if defined(__CUDACC__)
#define QHD_NOINLINE __host__ __device__ __noinline__
#else
#define QHD_NOINLINE inline
#endif

// ---- double-double pieces -----------------------------------------------------------

QHD dd dd_div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
  const double q2 = r.hi / b.hi;
  r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
  const double q3 = r.hi / b.hi;
  return dd_add_d(quick_two_sum(q1, q2), q3);
}

// sin(pi t) for a double-double |t| <= 1/2 to double-double accuracy (~2^-100 relative): Taylor
// series of sin on |pi t| <= pi/4, of cos on the rest (sin(pi t) = sgn(t) cos(pi (1/2 - |t|))).
QHD dd sinpi_acc(dd t) {
  const double SC[16][2] = {
      {1.0, 0.0},
      {-0.16666666666666666, -9.25185853854297e-18},
      {0.008333333333333333, 1.1564823173178714e-19},
      {-0.0001984126984126984, -1.7209558293420705e-22},
      {2.7557319223985893e-06, -1.858393274046472e-22},
      {-2.505210838544172e-08, 1.448814070935912e-24},
      {1.6059043836821613e-10, 1.2585294588752098e-26},
      {-7.647163731819816e-13, -7.03872877733453e-30},
      {2.8114572543455206e-15, 1.6508842730861433e-31},
      {-8.22063524662433e-18, -2.2141894119604265e-34},
      {1.9572941063391263e-20, -1.3643503830087908e-36},
      {-3.868170170630684e-23, 8.843177655482344e-40},
      {6.446950284384474e-26, -1.9330404233703465e-42},
      {-9.183689863795546e-29, -1.4303150396787322e-45},
      {1.1309962886447716e-31, 1.0498015412959506e-47},
      {-1.216125041553518e-34, -5.586290567888806e-51}};
  const double CC[16][2] = {
      {1.0, 0.0},
      {-0.5, 0.0},
      {0.041666666666666664, 2.3129646346357427e-18},
      {-0.001388888888888889, 5.300543954373577e-20},
      {2.48015873015873e-05, 2.1511947866775882e-23},
      {-2.755731922398589e-07, -2.3767714622250297e-23},
      {2.08767569878681e-09, -1.20734505911326e-25},
      {-1.1470745597729725e-11, -2.0655512752830745e-28},
      {4.779477332387385e-14, 4.399205485834081e-31},
      {-1.5619206968586225e-16, -1.1910679660273754e-32},
      {4.110317623312165e-19, 1.4412973378659527e-36},
      {-8.896791392450574e-22, 7.911402614872376e-38},
      {1.6117375710961184e-24, -3.6846573564509766e-41},
      {-2.4795962632247976e-27, 1.2953730964765229e-43},
      {3.279889237069838e-30, 1.5117542744029879e-46},
      {-3.7699876288159054e-33, -2.5870347832750324e-49}};
  const dd pi = make_dd(QB_PI_HI, QB_PI_LO);
  const bool neg = t.hi < 0;
  const dd a = neg ? dd_neg(t) : t;
  if (a.hi <= 0.25) {
    const dd y = dd_mul(pi, a);
    const dd z = dd_mul(y, y);
    dd p = make_dd(SC[15][0], SC[15][1]);
    for (int i = 14; i >= 0; i--) p = dd_add(dd_mul(p, z), make_dd(SC[i][0], SC[i][1]));
    p = dd_mul(p, y);
    return neg ? dd_neg(p) : p;
  }
  const dd y = dd_mul(pi, dd_add_d(dd_neg(a), 0.5));
  const dd z = dd_mul(y, y);
  dd p = make_dd(CC[15][0], CC[15][1]);
  for (int i = 14; i >= 0; i--) p = dd_add(dd_mul(p, z), make_dd(CC[i][0], CC[i][1]));
  return neg ? dd_neg(p) : p;
}

QHD uint64_t qb_double_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  memcpy(&b, &d, 8);
  return b;
#endif
}

// RN64 of a positive double-double (mpfr_get_ld of the reference's h); zero for anything
// non-positive or below the normal doubles.
QHD X87 x87_from_dd(dd a) {
  X87 r = x87_zero();
  if (!(a.hi > 0)) return r;
  const uint64_t bh = qb_double_bits(a.hi);
  const int ebh = (int)((bh >> 52) & 0x7ff);
  if (ebh == 0 || ebh == 0x7ff) return r;
  int eh = ebh - 1023;
  uint64_t Mh = ((bh & 0xfffffffffffffull) | (1ull << 52)) << 11, Ml = 0;  // a.hi = M 2^(eh - 127)
  bool sticky = false;
  const uint64_t bl = qb_double_bits(a.lo);
  const int ebl = (int)((bl >> 52) & 0x7ff);
  if (ebl != 0 && a.lo != 0.0) {
    const uint64_t ml = (bl & 0xfffffffffffffull) | (1ull << 52);
    const int sh = (ebl - 1023) - eh + 75;  // a.lo = ml 2^(sh) in units of 2^(eh - 127)
    uint64_t Lh = 0, Ll = 0;
    if (sh >= 0) {
      // |lo| <= ulp(hi) / 2 keeps sh <= 22
      Ll = ml << sh;
      Lh = sh ? (ml >> (64 - sh)) : 0;
    } else if (sh > -64) {
      Ll = ml >> (-sh);
      sticky = (ml << (64 + sh)) != 0;
    } else {
      sticky = true;
    }
    if (!(bl >> 63)) {
      const uint64_t s0 = Ml + Ll;
      const uint64_t c0 = s0 < Ml;
      const uint64_t s1 = Mh + Lh + c0;
      const bool carry = (s1 < Mh) || (c0 && s1 == Mh && Lh == ~0ull);
      Ml = s0;
      Mh = s1;
      if (carry) {  // hi had an all-ones mantissa and lo half a last place
        sticky = sticky || (Ml & 1ull);
        Ml = (Ml >> 1) | (Mh << 63);
        Mh = (Mh >> 1) | (1ull << 63);
        eh += 1;
      }
    } else {
      // M - L - (a fraction if sticky)
      uint64_t b0 = Ml < Ll;
      uint64_t s0 = Ml - Ll;
      uint64_t s1 = Mh - Lh - b0;
      if (sticky) {
        b0 = s0 == 0;
        s0 -= 1;
        s1 -= b0;
      }
      Ml = s0;
      Mh = s1;
      if (!(Mh >> 63)) {  // hi was a power of two
        Mh = (Mh << 1) | (Ml >> 63);
        Ml <<= 1;  // the shifted-in bit is below the sticky fraction: leave it zero, sticky covers it
        eh -= 1;
      }
    }
  }
  r.mant = Mh;
  r.exp = eh;
  r.neg = 0;
  x87_round(&r, Ml, sticky);
  return r;
}

// ---- multi-limb pieces ---------------------------------------------------------------

// Limb i of a per-sample number: word i * S of its array (S = 1 on the host; on the device the
// samples of a CTA are interleaved, S = the CTA's threads, a compile-time constant so that the
// unrolled loops address their operands with immediate offsets).
#define QB_L(p, i) (p)[(size_t)(i) * (size_t)S]

// 96-bit column accumulator of the product scanning.
#if defined(__CUDA_ARCH__)
struct Acc96 {
  uint32_t a0, a1, a2;
};
QHD void acc_zero(Acc96& a) { a.a0 = a.a1 = a.a2 = 0; }
QHD void acc_mad(Acc96& a, uint32_t x, uint32_t y) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a.a0), "+r"(a.a1), "+r"(a.a2)
      : "r"(x), "r"(y));
}
QHD uint32_t acc_pop(Acc96& a) {
  const uint32_t out = a.a0;
  a.a0 = a.a1;
  a.a1 = a.a2;
  a.a2 = 0;
  return out;
}
QHD void acc_add(Acc96& a, const Acc96& b) {
  asm("add.cc.u32 %0, %0, %3;\n\t"
      "addc.cc.u32 %1, %1, %4;\n\t"
      "addc.u32 %2, %2, %5;"
      : "+r"(a.a0), "+r"(a.a1), "+r"(a.a2)
      : "r"(b.a0), "r"(b.a1), "r"(b.a2));
}
#else
struct Acc96 {
  uint64_t lo;
  uint32_t hi;
};
QHD void acc_zero(Acc96& a) {
  a.lo = 0;
  a.hi = 0;
}
QHD void acc_mad(Acc96& a, uint32_t x, uint32_t y) {
  const uint64_t p = (uint64_t)x * (uint64_t)y;
  a.lo += p;
  a.hi += (a.lo < p) ? 1u : 0u;
}
QHD uint32_t acc_pop(Acc96& a) {
  const uint32_t out = (uint32_t)a.lo;
  a.lo = (a.lo >> 32) | ((uint64_t)a.hi << 32);
  a.hi = 0;
  return out;
}
QHD void acc_add(Acc96& a, const Acc96& b) {
  a.lo += b.lo;
  a.hi += b.hi + ((a.lo < b.lo) ? 1u : 0u);
}
#endif

// Columns [first, last] of the product U V (U: nu limbs at stride SU, V: nv limbs at stride SV;
// column c = sum of U[a] V[b] over a + b = c, plus the carry of the columns before): the limbs of
// the columns >= store_from go to out[(c - store_from) * SO]; `acc` holds the carry on entry (zero
// for a product that starts here).
//
// Four columns at a time: for a given a, the columns c .. c + 3 need U[a] and V[c - a .. c + 3 - a],
// and the next a needs the same window of V moved down by one -- so one pass over a loads U[a] and
// ONE new limb of V per step and feeds four accumulators (two loads per four multiply-adds). The
// pass runs over the UNION of the four columns' ranges of a: V MUST BE READABLE AND ZERO at the
// QB_DIAGK_PAD indices below 0 and above nv - 1, so that the terms a column does not have are
// products with zero. (Round 1 ran the pass over the intersection and added the ragged ends term
// by term: a fifth of the kernel's instructions, ncu source view, for 4 % of its multiply-adds.)
// Columns past `last` that the last group of four covers are computed and not stored; the carry
// that is returned is the one after that group.
#ifndef QB_MULCOL_ATTR
#define QB_MULCOL_ATTR QHD
#endif
#ifndef QB_DIAGK_UNROLL
#define QB_DIAGK_UNROLL 8
#endif
// Loads of V. In device code V lies in shared memory (k_diagk stages r, d and mu there) and is read
// with ld.shared instead of a generic load (+3 %; QB_DIAGK_LDS=0: generic loads, for the A/B).
#ifndef QB_DIAGK_LDS
#define QB_DIAGK_LDS 1
#endif
#if defined(__CUDA_ARCH__) && QB_DIAGK_LDS
typedef unsigned QB_VPTR;
__device__ __forceinline__ unsigned qb_vptr_of(const uint32_t* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t qb_vload(unsigned p, int i) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p + 4u * (unsigned)i));
  return v;
}
#define QB_VPTR_OF(p) qb_vptr_of(p)
#define QB_VLOAD(p, i) qb_vload(p, (int)(i))
#define QB_VSTEP(sv) (4u * (unsigned)(sv))
#else
typedef const uint32_t* QB_VPTR;
#define QB_VPTR_OF(p) (p)
#define QB_VLOAD(p, i) (p)[(size_t)(i)]
#define QB_VSTEP(sv) (sv)
#endif
template <int SU, int SV, int SO>
QB_MULCOL_ATTR Acc96 mul_columns(const uint32_t* __restrict__ U, uint32_t nu, const uint32_t* __restrict__ V,
                                 uint32_t nv, uint32_t first, uint32_t last, uint32_t store_from,
                                 uint32_t* __restrict__ out, Acc96 acc) {
#define QB_EMIT(c, limb)                                                                  \
  do {                                                                                    \
    if ((c) >= store_from && (c) <= last) out[(size_t)((c) - store_from) * SO] = (limb);  \
  } while (0)
  for (uint32_t c = first; c <= last; c += 4) {
    Acc96 a1, a2, a3;
    acc_zero(a1);
    acc_zero(a2);
    acc_zero(a3);
    const uint32_t alo = c >= nv ? c - nv + 1 : 0u;    // first a of column c
    const uint32_t ahi = c + 3 < nu ? c + 3 : nu - 1;  // last a of column c + 3
    if (alo <= ahi) {
      const uint32_t* pu = U + (size_t)alo * SU;
      QB_VPTR pv = QB_VPTR_OF(V + (size_t)(c - alo) * SV);  // c - alo <= nv - 1
      uint32_t w1 = QB_VLOAD(pv, 1 * SV), w2 = QB_VLOAD(pv, 2 * SV), w3 = QB_VLOAD(pv, 3 * SV);
      uint32_t n = ahi - alo + 1;
      constexpr int kUnroll = QB_DIAGK_UNROLL;
#pragma unroll kUnroll
      for (; n; n--) {
        const uint32_t u = *pu, v0 = QB_VLOAD(pv, 0);
        acc_mad(acc, u, v0);
        acc_mad(a1, u, w1);
        acc_mad(a2, u, w2);
        acc_mad(a3, u, w3);
        w3 = w2;
        w2 = w1;
        w1 = v0;
        pu += SU;
        pv -= QB_VSTEP(SV);  // down to V[c - ahi] >= V[-3]
      }
    }
    uint32_t limb = acc_pop(acc);
    QB_EMIT(c, limb);
    acc_add(acc, a1);
    limb = acc_pop(acc);
    QB_EMIT(c + 1, limb);
    acc_add(acc, a2);
    limb = acc_pop(acc);
    QB_EMIT(c + 2, limb);
    acc_add(acc, a3);
    limb = acc_pop(acc);
    QB_EMIT(c + 3, limb);
  }
#undef QB_EMIT
  return acc;
}

// The same product NC columns at a time (NC = 4: mul_columns; NC = 8 halves the loads of U and the loop
// overhead per multiply-add for 8 more accumulator words and a few more products with the zero
// limbs): V MUST BE READABLE AND ZERO at the NC - 1 indices below 0 and above nv - 1.
template <int NC, int SU, int SV, int SO>
QB_MULCOL_ATTR Acc96 mul_columns_wide(const uint32_t* __restrict__ U, uint32_t nu, const uint32_t* __restrict__ V,
                                      uint32_t nv, uint32_t first, uint32_t last, uint32_t store_from,
                                      uint32_t* __restrict__ out, Acc96 acc) {
  for (uint32_t c = first; c <= last; c += NC) {
    Acc96 a[NC];
    a[0] = acc;
#pragma unroll
    for (int k = 1; k < NC; k++) acc_zero(a[k]);
    const uint32_t alo = c >= nv ? c - nv + 1 : 0u;              // first a of column c
    const uint32_t ahi = c + (NC - 1) < nu ? c + (NC - 1) : nu - 1;  // last a of column c + NC - 1
    if (alo <= ahi) {
      const uint32_t* pu = U + (size_t)alo * SU;
      QB_VPTR pv = QB_VPTR_OF(V + (size_t)(c - alo) * SV);  // c - alo <= nv - 1
      uint32_t w[NC];
#pragma unroll
      for (int k = 1; k < NC; k++) w[k] = QB_VLOAD(pv, k * SV);
      uint32_t n = ahi - alo + 1;
      constexpr int kUnrollWide = NC >= 8 ? NC : 2 * NC;  // a multiple of NC: the window rotates without moves
#pragma unroll kUnrollWide
      for (; n; n--) {
        const uint32_t u = *pu, v0 = QB_VLOAD(pv, 0);
        acc_mad(a[0], u, v0);
#pragma unroll
        for (int k = 1; k < NC; k++) acc_mad(a[k], u, w[k]);
#pragma unroll
        for (int k = NC - 1; k > 1; k--) w[k] = w[k - 1];
        w[1] = v0;
        pu += SU;
        pv -= QB_VSTEP(SV);  // down to V[c - ahi] >= V[-(NC - 1)]
      }
    }
#pragma unroll
    for (int k = 0; k < NC; k++) {
      if (k) acc_add(a[0], a[k]);
      const uint32_t limb = acc_pop(a[0]);
      if (c + k >= store_from && c + k <= last) out[(size_t)(c + k - store_from) * SO] = limb;
    }
    acc = a[0];
  }
  return acc;
}

// W (k + 1 limbs, strided) >= r (k limbs)?
template <int S>
QHD bool limbs_ge_r(const uint32_t* W, const uint32_t* r, uint32_t k) {
  if (QB_L(W, k)) return true;
  for (uint32_t i = k; i-- > 0;) {
    const uint32_t a = QB_L(W, i), b = r[i];
    if (a != b) return a > b;
  }
  return true;
}
template <int S>
QHD void limbs_sub_r(uint32_t* W, const uint32_t* r, uint32_t k) {
  uint32_t borrow = 0;
  for (uint32_t i = 0; i < k; i++) {
    const uint64_t v = (uint64_t)QB_L(W, i) - r[i] - borrow;
    QB_L(W, i) = (uint32_t)v;
    borrow = (uint32_t)(v >> 63);
  }
  QB_L(W, k) -= borrow;
}

// Barrett division of x = A[0, 2k) < 2^(64 k) by r (Handbook of Applied Cryptography 14.42):
// Q (k + 2 limbs) = floor(x / r), W (k + 1 limbs, the top one zero on return) = x mod r. The
// columns of q1 mu below k - 1 are dropped (the estimate loses at most one more unit, which the
// final loop restores).
template <int S>
QHD_NOINLINE void diagk_barrett(const DiagKConst& c, const uint32_t* A, uint32_t* Q, uint32_t* W) {
  const uint32_t k = c.k;
  Acc96 acc;
  acc_zero(acc);
  // q2 = q1 mu with q1 = A[k - 1, 2k): k + 1 limbs, mu: k + 2 limbs; columns k + 1 .. 2k + 2 are Q
  // (column 2k + 2 is the carry out of the product's last column)
  acc = mul_columns<S, 1, S>(&QB_L(A, k - 1), k + 1, c.mu, k + 2, k - 1, 2 * k + 2, k + 1, Q, acc);
  // W = (x - Q r) mod 2^(32 (k + 1))
  acc_zero(acc);
  acc = mul_columns<S, 1, S>(Q, k + 2, c.r, k, 0, k, 0, W, acc);
  uint32_t borrow = 0;
  for (uint32_t col = 0; col <= k; col++) {
    const uint64_t v = (uint64_t)QB_L(A, col) - QB_L(W, col) - borrow;
    QB_L(W, col) = (uint32_t)v;
    borrow = (uint32_t)(v >> 63);
  }
  while (limbs_ge_r<S>(W, c.r, k)) {
    limbs_sub_r<S>(W, c.r, k);
    for (uint32_t i = 0; i <= k + 1; i++) {
      if (++QB_L(Q, i)) break;
    }
  }
}

// Four limbs from index `top` downwards as a double-double (limb top - 3 has weight 1).
template <int S>
QHD dd limbs_top_dd(const uint32_t* p, uint32_t top) {
  dd v = make_dd(0.0, 0.0);
  double w = 79228162514264337593543950336.0;  // 2^96
  for (int i = 0; i < 4; i++) {
    const uint32_t limb = (top >= (uint32_t)i) ? QB_L(p, top - i) : 0u;
    v = dd_add_d(v, (double)limb * w);
    w *= 2.3283064365386962890625e-10;  // 2^-32
  }
  return v;
}

// h = S / (2^l sin(pi x / 2^l))^2 with S = sin^2(pi x) (diagonal_probability_approx_h at
// phi = 2 pi x / 2^l, src/diagonal_probability.cpp:99-162), x on [-2^(l-1), 2^(l-1)).
//   l >= 110: (2^l sin(pi x / 2^l))^2 = (pi x)^2 to double-double accuracy for |x| < 2^33.
QHD dd diagk_h(uint32_t l, dd S, dd x) {
  if (x.hi == 0.0) return make_dd(1.0, 0.0);  // :104-108
  dd D;
  if (l >= 110) {
    D = dd_mul(make_dd(QB_PI_HI, QB_PI_LO), x);
  } else {
    D = sinpi_acc(dd_mul_pow2(x, ldexp(1.0, -(int)l)));
    D = dd_mul_pow2(D, ldexp(1.0, (int)l));
  }
  return dd_div(S, dd_mul(D, D));
}

// x = t + delta folded as the reference folds k and phi: k = (k0 + delta) mod 2^l and phi on
// [-2^(m+sigma-1), 2^(m+sigma-1)) (src/sample.cpp:552-574). Only l < 62 can wrap (|delta| < 2^32).
QHD dd diagk_walk_x(uint32_t l, dd t, int64_t delta) {
  if (l >= 62) return dd_add_d(t, (double)delta);
  const int64_t M = (int64_t)1 << l;
  const double two_l = (double)M;
  int64_t dm = delta;
  if (delta >= M / 2 || delta < -(M / 2)) {
    dm = ((delta % M) + M) % M;
    if (dm >= M / 2) dm -= M;  // |dm| <= 2^32 or < 2^53: exact as a double
  }
  dd x = dd_add_d(t, (double)dm);
  if (x.hi >= 0.5 * two_l) x = dd_add_d(x, -two_l);
  if (x.hi < -0.5 * two_l) x = dd_add_d(x, two_l);
  return x;
}

// Step idx of the walk: delta = 0, +1, -1, +2, -2, ... (src/sample.cpp:539-559); the last one is
// idx = 2 delta_bound.
QHD int64_t diagk_step_delta(uint64_t idx) {
  const int64_t da = (int64_t)((idx + 1) >> 1);
  return (idx & 1) ? da : -da;
}

// h in plain doubles (a few units in the last place): Sd = sin^2(pi t), two_l = 2^l and
// inv_two_l = 2^-l for l < 110.
QHD double diagk_quick_h(uint32_t l, double two_l, double inv_two_l, double Sd, dd x) {
  if (x.hi == 0.0) return 1.0;
  double D;
  if (l >= 110) {
    D = QB_PI_HI * x.hi;
  } else {
    const double y = x.hi * inv_two_l;
    if (fabs(y) < 1.220703125e-4) {  // 2^-13: sin(e) / e = 1 - e^2/6 (1 - e^2/20) to 1e-24
      const double e = QB_PI_HI * y, z = e * e;
      D = QB_PI_HI * x.hi * (1.0 - z * (1.0 / 6.0) * (1.0 - z * 0.05));
    } else {
      D = two_l * sinpi_dd(dd_mul_pow2(x, inv_two_l));
    }
  }
  return Sd / (D * D);
}

// State of the pass in doubles: the pivot after the steps before idx (accumulated as an unevaluated
// sum of two doubles, so that the additions themselves do not contribute to the error).
struct DiagKQuick {
  dd p;
  uint64_t idx;
};
// The error band after N steps: 2^-47 for the h (each to ~8 units in the last place of a double,
// their sum at most 2 because the walk stops once it reaches the pivot <= 1) and the partial sums
// inside a warp step, plus N 2^-63 for the reference's own N roundings to 64 bits of a number <= 1.
QHD double diagk_band(uint64_t steps) { return 7.105427357601002e-15 + (double)steps * 1.0842021724855044e-19; }
#define QB_DIAGK_CONTINUE 100   // internal: not decided before idx_end
#define QB_DIAGK_UNDECIDED 101  // internal: the pivot ended inside the error band: exact walk

// Steps q->idx .. idx_end - 1 of the pass in doubles (see diagk_walk). QB_DIAGK_OK (with delta and
// x), QB_DIAGK_OUT_OF_BOUNDS, QB_DIAGK_GAVE_UP, QB_DIAGK_UNDECIDED or QB_DIAGK_CONTINUE.
QHD int diagk_quick_steps(uint32_t l, dd t, double Sd, uint64_t delta_bound, uint64_t idx_end, DiagKQuick* q,
                          int64_t* delta_out, dd* x_out) {
  const double two_l = l < 110 ? ldexp(1.0, (int)l) : 0.0, inv_two_l = l < 110 ? ldexp(1.0, -(int)l) : 0.0;
  const uint64_t last = 2 * delta_bound;
  while (q->idx < idx_end) {
    if (q->idx > last) {  // above the band after the last step: out of bounds in the reference as well
      *delta_out = 0;
      *x_out = make_dd(0.0, 0.0);
      return QB_DIAGK_OUT_OF_BOUNDS;
    }
    if (q->idx + 1 > QB_DIAGK_MAX_STEPS) return QB_DIAGK_GAVE_UP;
    const int64_t delta = diagk_step_delta(q->idx);
    const dd x = diagk_walk_x(l, t, delta);
    q->p = dd_add_d(q->p, -diagk_quick_h(l, two_l, inv_two_l, Sd, x));
    const double pv = q->p.hi + q->p.lo, band = diagk_band(q->idx + 1);
    if (pv < -band) {
      *delta_out = delta;
      *x_out = x;
      return QB_DIAGK_OK;
    }
    if (pv <= band) return QB_DIAGK_UNDECIDED;
    q->idx++;
  }
  return QB_DIAGK_CONTINUE;
}

// The exact walk: h in double-double, rounded to the x87 format, subtracted with the x87 rounding
// (pivot -= mpfr_get_ld(h), src/sample.cpp:594).
QHD int diagk_walk_exact(uint32_t l, dd t, dd S, X87 pivot, uint64_t delta_bound, int64_t* delta_out, dd* x_out) {
  const uint64_t last = 2 * delta_bound;
  for (uint64_t idx = 0; idx <= last; idx++) {
    if (idx + 1 > QB_DIAGK_MAX_STEPS) return QB_DIAGK_GAVE_UP;
    const int64_t delta = diagk_step_delta(idx);
    const dd x = diagk_walk_x(l, t, delta);
    pivot = x87_add(pivot, x87_neg(x87_from_dd(diagk_h(l, S, x))));
    if (x87_nonpositive(pivot)) {
      *delta_out = delta;
      *x_out = x;
      return QB_DIAGK_OK;
    }
  }
  *delta_out = 0;
  *x_out = make_dd(0.0, 0.0);
  return QB_DIAGK_OUT_OF_BOUNDS;
}

// The walk of src/sample.cpp:539-604 on x = t + delta: pivot -= h(x) for delta = 0, 1, -1, 2, ...
// until the pivot is used up.
//
// Only the step at which that happens is an output, not the pivot. A first pass therefore runs in
// doubles (h in plain doubles, the pivot as a compensated sum): after N steps it differs from the
// reference's sequence of long double subtractions by less than diagk_band(N), so a pivot that
// passes below -band at a step while it was above +band at all earlier steps stops at that step
// in the reference as well. A pivot inside the band at some step (probability ~1e-4 for a walk of
// 10^5 steps, far less for short ones) sends the sample through the exact walk: h in double-double, rounded to the
// x87 format, subtracted with the x87 rounding -- from the start. force_exact: the exact walk
// only (tests). (The kernel runs the first steps of the pass in doubles per thread and the rest
// thirty-two steps at a time across the warp, kernels_diagk.cuh: any order of summation stays
// inside the same band.)
QHD int diagk_walk(uint32_t l, dd t, X87 pivot, uint64_t delta_bound, int force_exact, int64_t* delta_out,
                   dd* x_out) {
  const dd st = sinpi_acc(t);
  const dd S = dd_mul(st, st);
  if (!force_exact) {
    DiagKQuick q;
    q.p = x87_to_dd(pivot);  // exact
    q.idx = 0;
    const int status = diagk_quick_steps(l, t, S.hi, delta_bound, ~(uint64_t)0, &q, delta_out, x_out);
    if (status != QB_DIAGK_UNDECIDED) return status;
  }
  return diagk_walk_exact(l, t, S, pivot, delta_bound, delta_out, x_out);
}

// (Qv, t, c) with Qv + t + c = 2^l d s / r mod 2^l, exactly: Qv to k_out (if wanted), t = w2 / r - c,
// c = [2 w2 >= r], for s in [0, r) in Sq. 2^l d = D' r + rho (host constants: D' = floor(2^l d / r)
// < 2^l, rho < r), so 2^l d s = (s D' + floor(s rho / r)) r + (s rho mod r): one full product
// s rho, one Barrett division, and -- only if k is wanted -- the low l bits of s D'. (Round 1
// formed w = d s mod r and then divmod(2^l w, r), 32 k bits of the shift at a time: two Barrett
// divisions and more for l > 32 k. The quotients differ by a multiple of 2^l.)
// This is the path of the few samples diagk_fraction_fixed_point leaves open.
template <int S>
QHD_NOINLINE void diagk_fraction_exact(const DiagKConst& c, const uint32_t* Sq, uint32_t* A, uint32_t* Q,
                                       uint32_t* W, uint32_t* k_out, dd* t_out, bool* cflag_out) {
  const uint32_t k = c.k;
  Acc96 acc;
  acc_zero(acc);
  acc = mul_columns<S, 1, S>(Sq, k, c.rho, k, 0, 2 * k - 1, 0, A, acc);
  diagk_barrett<S>(c, A, Q, W);  // Q = floor(s rho / r) < r, W = w2
  if (k_out) {
    acc_zero(acc);
    acc = mul_columns<S, 1, S>(Sq, k, c.dq, c.wl, 0, c.wl - 1, 0, k_out, acc);  // columns past l bits: dropped
    uint32_t carry = 0;
    for (uint32_t i = 0; i < c.wl; i++) {
      const uint64_t v = (uint64_t)QB_L(k_out, i) + (i < k ? QB_L(Q, i) : 0u) + carry;
      QB_L(k_out, i) = (uint32_t)v;
      carry = (uint32_t)(v >> 32);
    }  // (bits above l: masked in diagk_finish)
  }
  // ---- t = w2 / r - c ----
  // 2 w2 >= r  <=>  w2 >= r - w2: form r - w2 in A and compare
  uint32_t borrow = 0;
  for (uint32_t i = 0; i < k; i++) {
    const uint64_t v = (uint64_t)c.r[i] - QB_L(W, i) - borrow;
    QB_L(A, i) = (uint32_t)v;
    borrow = (uint32_t)(v >> 63);
  }
  bool cflag = true;  // w2 >= r - w2
  for (uint32_t i = k; i-- > 0;) {
    const uint32_t a = QB_L(W, i), b = QB_L(A, i);
    if (a != b) {
      cflag = a > b;
      break;
    }
  }
  const uint32_t* N = cflag ? A : W;
  int top = -1;
  for (uint32_t i = k; i-- > 0;) {
    if (QB_L(N, i)) {
      top = (int)i;
      break;
    }
  }
  dd t = make_dd(0.0, 0.0);
  if (top >= 0) {
    const int e = 32 * (top - (int)k + 1);
    if (e >= -960) {
      t = dd_div(limbs_top_dd<S>(N, (uint32_t)top), c.r_top);
      t = dd_mul_pow2(t, pow2i(e));
      if (cflag) t = dd_neg(t);
    }
  }
  *t_out = t;
  *cflag_out = cflag;
}

// The same from ONE truncated product, for all but ~2^-31 of the samples. psi = floor(2^(32 fl)
// 2^l d / r) is 2^l d / r in fixed point with fl = k + 4 fractional limbs (error < 2^(-32 fl)), so
// s psi / 2^(32 fl) misses 2^l d s / r by less than 2^-128: its integer part is Qv + c's carrier,
// its four leading fractional limbs F' give the fraction F = w2 / r with F' <= F < F' + 2^-127.
// Computed: the columns fl - 8 .. fl - 1 of s psi (four guard columns as for r j, four fractional
// limbs) and, if k is wanted, the wl columns of the integer part. Returns false -- nothing decided,
// k_out untouched -- when the guard, the integer part (F' within 2^-127 of 1), c (F' within 2^-127
// of 1/2) or the relative accuracy of t (|t| < 2^-31: t must be good to 2^-96 of itself) is in doubt.
template <int S>
QHD bool diagk_fraction_fixed_point(const DiagKConst& c, const uint32_t* Sq, uint32_t* A, uint32_t* k_out,
                                    dd* t_out, bool* cflag_out) {
  const uint32_t k = c.k, fl = c.k + 4, np = fl + c.wl;
  if (c.exact_fraction || fl < 8) return false;  // (r of fewer than four limbs: the exact path)
  Acc96 acc;
  acc_zero(acc);
  acc = mul_columns<S, 1, S>(Sq, k, c.psi, np, fl - 8, fl - 1, fl - 8, A, acc);  // carry: into column fl
  if (QB_L(A, 3) == 0xffffffffu) return false;  // the skipped columns' carry may pass the guards
  const uint32_t f0 = QB_L(A, 4), f1 = QB_L(A, 5), f2 = QB_L(A, 6), f3 = QB_L(A, 7);  // f3: leading
  const bool run = f2 == 0xffffffffu && f1 == 0xffffffffu && f0 >= 0xfffffffeu;
  if (run && (f3 == 0xffffffffu || f3 == 0x7fffffffu)) return false;  // F' within 2^-127 of 1 or of 1/2
  const bool cflag = (f3 >> 31) != 0;
  // |t| 2^128 = F' or 2^128 - F' as an integer, then as a double-double
  uint32_t n[4] = {f0, f1, f2, f3};
  if (cflag) {
    uint32_t one = 1;
    for (int i = 0; i < 4; i++) {
      const uint64_t v = (uint64_t)(~n[i]) + one;
      n[i] = (uint32_t)v;
      one = (uint32_t)(v >> 32);
    }
  }
  dd t = limbs_top_dd<1>(n, 3);
  t = dd_mul_pow2(t, 2.938735877055718769921841343055614194546e-39);  // 2^-128
  if (!(t.hi >= 4.656612873077392578125e-10)) return false;  // |t| < 2^-31 (or zero)
  if (cflag) t = dd_neg(t);
  if (k_out) acc = mul_columns<S, 1, S>(Sq, k, c.psi, np, fl, fl + c.wl - 1, fl, k_out, acc);
  *t_out = t;
  *cflag_out = cflag;
  return true;
}

struct DiagKFraction {
  dd t;        // w2 / r - c
  bool cflag;  // c = [2 w2 >= r]
  bool whole;  // q + eta < 0 and d |q + eta| >= r: the unreduced phi is negative
};

// The integer part of one sample: k_out = Qv (if wanted), f = (t, c, whole).
template <int S>
QHD void diagk_fraction(const DiagKConst& c, const uint32_t* j, int32_t eta, uint32_t* scratch, uint32_t* k_out,
                        DiagKFraction* f) {
  const uint32_t k = c.k;
  uint32_t* A = scratch;                              // 2k + 2
  uint32_t* Q = A + (size_t)(2 * k + 2) * (size_t)S;  // k + 3
  uint32_t* W = Q + (size_t)(k + 3) * (size_t)S;      // k + 2
  uint32_t* Sq = W + (size_t)(k + 2) * (size_t)S;     // k + 2: s = q + eta

  // ---- q = round-to-centred quotient of r j by 2^n (alpha_r = {r j}_{2^n}, :477-478) ----
  const uint32_t cs = c.n >> 5, sh = c.n & 31;
  Acc96 acc;
  acc_zero(acc);
  // Z = r j: only the columns from cs - 1 on are needed (they go to A, free here, then to Sq). The
  // product therefore starts three columns further down, with no carry from the columns it skips:
  // that carry is below k 2^32, so it can reach past the three guard columns only if the third one
  // comes out as 0xffffffff (one sample in 2^32) -- then the product is formed in full.
  const uint32_t ncol = k + c.wj;
  const uint32_t zs = cs >= 2 ? cs - 2 : cs - 1;          // first column stored: the last guard (or cs - 1)
  uint32_t first = (zs == cs - 2 && zs >= 2 && !c.full_product) ? zs - 2 : 0u;
  for (;;) {
    acc_zero(acc);
    acc = mul_columns<S, 1, S>(j, c.wj, c.r, k, first, ncol - 1, zs, A, acc);
    if (first == 0 || QB_L(A, 0) != 0xffffffffu) break;
    first = 0;
  }
  const uint32_t* Az = &QB_L(A, cs - 1 - zs);  // Az[i]: column cs - 1 + i
  const uint32_t below = QB_L(Az, 0);
  for (uint32_t i = 0; i + cs < ncol; i++) QB_L(Sq, i) = QB_L(Az, i + 1);
  const uint32_t ns = ncol - cs;  // <= k + 1 limbs hold Z >> (32 cs)
  for (uint32_t i = ns; i < k + 2; i++) QB_L(Sq, i) = 0;
  uint32_t half_bit;
  if (sh == 0) {
    half_bit = below >> 31;
  } else {
    half_bit = (QB_L(Sq, 0) >> (sh - 1)) & 1u;
    for (uint32_t i = 0; i < k + 1; i++)
      QB_L(Sq, i) = (QB_L(Sq, i) >> sh) | (QB_L(Sq, i + 1) << (32 - sh));
    QB_L(Sq, k + 1) >>= sh;
  }
  // ---- s = q + eta, brought to [0, r): adding r to s adds the integer d to d s / r ----
  int64_t carry = (int64_t)half_bit + (int64_t)eta;
  for (uint32_t i = 0; i < k + 2; i++) {
    const int64_t v = (int64_t)QB_L(Sq, i) + carry;
    QB_L(Sq, i) = (uint32_t)v;
    carry = v >> 32;
    if (carry == 0) break;
  }
  const bool s_negative = carry < 0;
  bool whole = false;  // q + eta < 0 and d |q + eta| >= r: the unreduced phi is negative
  if (s_negative) {
    const uint32_t abs_s = 0u - QB_L(Sq, 0);  // |q + eta| <= |eta|
    uint64_t cw = 0;
    for (uint32_t i = 0; i < k; i++) {
      const uint64_t v = (uint64_t)c.d[i] * abs_s + cw;
      QB_L(A, i) = (uint32_t)v;
      cw = v >> 32;
    }
    QB_L(A, k) = (uint32_t)cw;
    whole = limbs_ge_r<S>(A, c.r, k);
    uint64_t cy = 0;
    for (uint32_t i = 0; i < k + 2; i++) {
      const uint64_t v = (uint64_t)QB_L(Sq, i) + (i < k ? c.r[i] : 0u) + cy;
      QB_L(Sq, i) = (uint32_t)v;
      cy = v >> 32;
    }
  }
  for (int guard = 0; guard < 64; guard++) {
    if (QB_L(Sq, k + 1) == 0 && !limbs_ge_r<S>(Sq, c.r, k)) break;
    // s >= r: s <= r + |eta| + 1 for j < 2^n, so one round is the rule
    uint32_t borrow = 0;
    for (uint32_t i = 0; i < k + 2; i++) {
      const uint64_t v = (uint64_t)QB_L(Sq, i) - (i < k ? c.r[i] : 0u) - borrow;
      QB_L(Sq, i) = (uint32_t)v;
      borrow = (uint32_t)(v >> 63);
    }
  }
  // ---- Qv + t + c = 2^l d s / r mod 2^l ----
  dd t = make_dd(0.0, 0.0);
  bool cflag = false;
  if (!diagk_fraction_fixed_point<S>(c, Sq, A, k_out, &t, &cflag)) diagk_fraction_exact<S>(c, Sq, A, Q, W, k_out, &t, &cflag);
  f->t = t;
  f->cflag = cflag;
  f->whole = whole;
}

// After the walk: k = (k0 + delta) mod 2^l from Qv in k_out, the status, x.
template <int S>
QHD int diagk_finish(const DiagKConst& c, const DiagKFraction& f, int status, int64_t delta, dd x, uint32_t* k_out,
                     dd* x_out, int64_t* delta_out) {
  if (status != QB_DIAGK_OK) {
    if (k_out)
      for (uint32_t i = 0; i < c.wl; i++) QB_L(k_out, i) = 0;
    *x_out = make_dd(0.0, 0.0);
    *delta_out = 0;
    return status;
  }
  // mpfr_fmod keeps the sign of a negative dividend (:566-574): with q + eta < 0 and
  // d |q + eta| >= r the unreduced phi is negative and a positive residue is not folded back
  // (j < |eta| 2^(m+sigma) / r: never drawn in practice). 2^l may not be a double: reported as
  // a status, the subtraction is the caller's.
  const int ok_status = (f.whole && x.hi > 0.0) ? QB_DIAGK_OK_NEGATIVE_PHI : QB_DIAGK_OK;
  if (k_out) {
    // k = (-(Qv + c) + delta) mod 2^l = -(Qv + c - delta) mod 2^l
    int64_t cy = (int64_t)(f.cflag ? 1 : 0) - delta;
    for (uint32_t i = 0; i < c.wl; i++) {
      const int64_t v = (int64_t)QB_L(k_out, i) + cy;
      QB_L(k_out, i) = (uint32_t)v;
      cy = v >> 32;
    }
    uint32_t one = 1;
    for (uint32_t i = 0; i < c.wl; i++) {
      const uint64_t v = (uint64_t)(~QB_L(k_out, i)) + one;
      QB_L(k_out, i) = (uint32_t)v;
      one = (uint32_t)(v >> 32);
    }
    if (c.l & 31) QB_L(k_out, c.wl - 1) &= (1u << (c.l & 31)) - 1u;
  }
  *x_out = x;
  *delta_out = delta;
  return ok_status;
}

// One sample. j: c.wj limbs (j < 2^n); scratch: diagk_scratch_limbs(k) words; k_out: c.wl limbs
// (may be null); all three with stride S. x_out = alpha_phi / 2^(m + sigma - l).
template <int S>
QHD int diagk_sample(const DiagKConst& c, const uint32_t* j, int32_t eta, X87 pivot, uint64_t delta_bound,
                     uint32_t* scratch, uint32_t* k_out, dd* x_out, int64_t* delta_out) {
  DiagKFraction f;
  diagk_fraction<S>(c, j, eta, scratch, k_out, &f);
  int64_t delta = 0;
  dd x = make_dd(0.0, 0.0);
  const int status = diagk_walk(c.l, f.t, pivot, delta_bound, c.force_exact, &delta, &x);
  return diagk_finish<S>(c, f, status, delta, x, k_out, x_out, delta_out);
}

}  // namespace qb200
