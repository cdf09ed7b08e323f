// exact_host.hpp -- per-distribution constants of the exact samplers (no CUDA in this file).
//
// Once per (kind, m, l or sigma, d, r, largest dimension): kappa_d, kappa_r (the powers of two in d
// and r; kappa(), src/math.cpp), (r / 2^kappa_r)^-1 and (d / 2^kappa_d)^-1 modulo 2^n -- what
// sample_j_from_alpha_r, sample_j_k_from_alpha_d[_r] and sample_j_from_diagonal_alpha_r recompute
// with mpz_invert for every sample (src/sample.cpp:185-189, 244-248, 314-318, 380-385) -- and the
// table 2^(i / D_max), i = 0 .. D_max - 1, in fixed point with emax + 128 fractional bits, from which
// the kernels form the bounds round(2^|log alpha|) of a region (exact.cuh) that
// sample_alpha_from_region computes with two mpfr_exp2 per sample (src/sample.cpp:97-124).
// Own big integers (bigint.hpp): the library has no GMP / MPFR dependency.
// Shared by the CUDA library (qb200_exact.cu) and the test-only CPU twin of tests/hostsim.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "bigint.hpp"
#include "diagk_host.hpp"
#include "exact.cuh"

namespace qb200 {

#define QB_EXACT_KIND_TWO_DIMENSIONAL 0
#define QB_EXACT_KIND_DIAGONAL 1
#define QB_EXACT_MAX_KAPPA 64

struct ExactHost {
  std::vector<uint32_t> inv_r, inv_d, d;  // QB_EXACT_PAD zero limbs, the number, QB_EXACT_PAD zero limbs
  std::vector<uint32_t> table;
  ExactConst c;
};

inline BigUInt big_low_bits(const BigUInt& x, size_t n) {
  BigUInt r = x;
  const size_t words = (n + 63) / 64;
  if (r.w.size() > words) r.w.resize(words);
  if ((n % 64) && r.w.size() == words) r.w[words - 1] &= (uint64_t(1) << (n % 64)) - 1;
  r.trim();
  return r;
}

inline uint32_t big_trailing_zeros(const BigUInt& x) {
  uint32_t k = 0;
  while (!x.bit(k)) k++;
  return k;
}

// a^-1 modulo 2^n for odd a: Newton's iteration x <- x (2 - a x), doubling the correct low bits.
inline BigUInt big_inverse_mod_pow2(const BigUInt& a, size_t n) {
  BigUInt x(1);
  const BigUInt two_n = BigUInt::pow2(n);
  for (size_t good = 1; good < n; good *= 2) {
    const BigUInt ax = big_low_bits(BigUInt::mul(a, x), n);
    // 2 - a x modulo 2^n
    const BigUInt u = big_low_bits(BigUInt::sub(BigUInt::add(two_n, BigUInt(2)), ax), n);
    x = big_low_bits(BigUInt::mul(x, u), n);
  }
  return x;
}

// floor(sqrt(x)).
inline BigUInt big_isqrt(const BigUInt& x) {
  if (x.is_zero()) return BigUInt();
  BigUInt y = BigUInt::pow2((x.bit_length() + 1) / 2);  // >= sqrt(x)
  for (;;) {
    BigUInt q, r;
    BigUInt::divmod(x, y, q, r);
    const BigUInt z = BigUInt::add(y, q).shr(1);
    if (BigUInt::cmp(z, y) >= 0) return y;
    y = z;
  }
}

// Entry i = 2^(i / D) 2^P, i = 0 .. D - 1, within 2 units (entry 0 exact), each in `tw` 32-bit words.
// 2^(1/D) by log2(D) square roots with 64 bits more (each floor costs one unit of 2^-(P+64) and
// halves the error it inherits), entry i = entry i - 1 times that root, truncated: entry i is low
// by at most 3 i + 1 < 2^16 units of 2^-(P+64).
inline void exact_exp2_table(uint32_t log_d, uint32_t P, uint32_t tw, std::vector<uint32_t>* out) {
  const size_t Q = (size_t)P + 64;
  const uint32_t D = 1u << log_d;
  BigUInt root = BigUInt::pow2(Q + 1);  // 2 2^Q
  for (uint32_t k = 1; k <= log_d; k++) root = big_isqrt(root.shl(Q));
  out->assign((size_t)D * tw, 0u);
  BigUInt acc = BigUInt::pow2(Q);
  for (uint32_t i = 0; i < D; i++) {
    if (i) acc = BigUInt::mul(acc, root).shr(Q);
    const BigUInt v = acc.shr(64);
    for (uint32_t w = 0; w < tw; w++) {
      const size_t k = w / 2;
      if (k < v.w.size()) (*out)[(size_t)i * tw + w] = (uint32_t)(v.w[k] >> (32 * (w % 2)));
    }
  }
}

// 0, or a negative code with *err set. emax = 0: m + 64 (a generator's slices end below m + 60),
// never above n. The pointers of h->c refer to the vectors of *h; the CUDA library replaces them by
// device copies.
inline int exact_prepare(int kind, uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d_be, size_t d_len,
                         const uint8_t* r_be, size_t r_len, uint32_t dimension_max, uint32_t emax, ExactHost* h,
                         std::string* err) {
  if (!d_be || !r_be) {
    *err = "null argument";
    return -1;
  }
  if (kind != QB_EXACT_KIND_TWO_DIMENSIONAL && kind != QB_EXACT_KIND_DIAGONAL) {
    *err = "exact sampler: unknown kind";
    return -2;
  }
  const BigUInt d = BigUInt::from_bytes_be(d_be, d_len), r = BigUInt::from_bytes_be(r_be, r_len);
  if (r.is_zero() || d.is_zero()) {
    *err = "exact sampler: need d, r > 0";
    return -2;
  }
  const uint32_t n = kind == QB_EXACT_KIND_DIAGONAL ? m + sigma : m + l;
  if (m < 8 || n <= m || n > (1u << 15) || r.bit_length() > n || d.bit_length() > n) {
    *err = "exact sampler: need m >= 8, l (sigma) > 0, m + l (m + sigma) <= 32768, d, r < 2^n";
    return -3;
  }
  if (dimension_max == 0 || (dimension_max & (dimension_max - 1)) != 0 || dimension_max > (1u << 14)) {
    *err = "exact sampler: the largest dimension must be a power of two up to 16384";
    return -3;
  }
  ExactConst& c = h->c;
  c.m = m;
  c.l = l;
  c.sigma = sigma;
  c.n = n;
  c.kbits = kind == QB_EXACT_KIND_DIAGONAL ? 0 : l;
  c.kappa_d = big_trailing_zeros(d);
  c.kappa_r = big_trailing_zeros(r);
  if (c.kappa_d > QB_EXACT_MAX_KAPPA || c.kappa_r > QB_EXACT_MAX_KAPPA) {
    *err = "exact sampler: d or r divisible by more than 2^64 is not supported";
    return -4;
  }
  c.wn = (n + 31) / 32;
  c.wd = (uint32_t)((d.bit_length() + 31) / 32);
  c.wk = (c.kbits + 31) / 32;
  if (emax == 0) emax = m + 64;
  if (emax > n) emax = n;
  if (emax < 8) {
    *err = "exact sampler: emax below 8";
    return -3;
  }
  c.emax = emax;
  c.wa = (emax + 1 + 31) / 32;
  c.table_dim = dimension_max;
  c.table_log = 0;
  while ((1u << c.table_log) < dimension_max) c.table_log++;
  c.P = emax + QB_EXACT_GUARD;
  c.tw = (c.P + 1 + 31) / 32;
  h->inv_r = limbs32_padded(big_inverse_mod_pow2(r.shr(c.kappa_r), n), c.wn, QB_EXACT_PAD);
  h->inv_d = limbs32_padded(big_inverse_mod_pow2(d.shr(c.kappa_d), n), c.wn, QB_EXACT_PAD);
  h->d = limbs32_padded(d, c.wd, QB_EXACT_PAD);
  exact_exp2_table(c.table_log, c.P, c.tw, &h->table);
  c.inv_r = h->inv_r.data() + QB_EXACT_PAD;
  c.inv_d = h->inv_d.data() + QB_EXACT_PAD;
  c.d = h->d.data() + QB_EXACT_PAD;
  c.table = h->table.data();
  return 0;
}

}  // namespace qb200
