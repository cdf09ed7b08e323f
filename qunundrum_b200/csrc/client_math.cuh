// client_math.cuh -- the per-element arithmetic of kernels_client.cuh (copy_scale, collapse) as
// __host__ __device__ functions, so that tests/hostsim can run the identical code on the CPU
// against the reference's own distribution_slice_copy_scale / linear_distribution_init_collapse_*.
#pragma once

#include "x87soft.cuh"

namespace qb200 {

// RN64(a / q), q >= 1 (probability / (long double)divisor, src/linear_distribution.cpp:224-227):
// long division of mant * 2^64 by q, 32 bits at a time; the quotient's top 64 bits are the
// mantissa, the next bit and the remainder decide the rounding.
QHD X87 x87_div_u32(X87 a, uint32_t q) {
  if (a.mant == 0 || q == 1) return a;
  if ((q & (q - 1)) == 0) {  // exact for powers of two
    int k = 0;
    while ((1u << k) < q) k++;
    return x87_div_pow2(a, k);
  }
  // dividend limbs (most significant first): mant_hi, mant_lo, 0, 0, 0 -> 160 bits
  const uint32_t limb[5] = {(uint32_t)(a.mant >> 32), (uint32_t)a.mant, 0u, 0u, 0u};
  uint32_t quo[5];
  uint64_t rem = 0;
  for (int i = 0; i < 5; i++) {
    const uint64_t cur = (rem << 32) | limb[i];
    quo[i] = (uint32_t)(cur / q);
    rem = cur % q;
  }
  // quotient = quo[0..4] as a 160-bit number = floor(mant 2^96 / q); value = quotient 2^(exp - 63 - 96)
  uint64_t hi = ((uint64_t)quo[0] << 32) | quo[1];
  uint64_t mid = ((uint64_t)quo[2] << 32) | quo[3];
  uint64_t lo = (uint64_t)quo[4] << 32;
  // hi is non-zero: mant >= 2^63 and q < 2^32
  const int lz = qb_clz64(hi);
  X87 r;
  r.neg = a.neg;
  r.exp = a.exp - lz;
  uint64_t rest;
  if (lz) {
    r.mant = (hi << lz) | (mid >> (64 - lz));
    rest = (mid << lz) | (lo >> (64 - lz));
    lo <<= lz;
  } else {
    r.mant = hi;
    rest = mid;
  }
  x87_round(&r, rest, lo != 0 || rem != 0);
  return r;
}

QHD X87 x87_load(uint64_t mant, uint64_t sign_exp, bool* ok) {
  X87 v;
  if (!x87_decode(mant, (uint32_t)sign_exp, &v)) *ok = false;
  return v;
}

// One destination cell of distribution_slice_copy_scale (src/distribution_slice.cpp:249-261):
// 0 + the scale x scale block of the source, alpha_d offset outermost. src: D x D doubles.
QHD X87 scale_cell_x87(const double* src, int D, int store, int idx) {
  const int scale = D / store;
  const int i1 = idx % store, j1 = idx / store;
  X87 acc = x87_zero();
  for (int i2 = 0; i2 < scale; i2++)
    for (int j2 = 0; j2 < scale; j2++)
      acc = x87_add(acc, x87_from_double(src[(size_t)(i1 * scale + i2) + (size_t)D * (j1 * scale + j2)]));
  return acc;
}

// norm_vector[...] += probability / (long double)divisor  (src/linear_distribution.cpp:224-227)
QHD X87 collapse_step(X87 acc, uint64_t mant, uint64_t sign_exp, uint32_t divisor, bool* ok) {
  return x87_add(acc, x87_div_u32(x87_load(mant, sign_exp, ok), divisor));
}

}  // namespace qb200
