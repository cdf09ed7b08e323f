// diagk_host.hpp -- per-distribution constants of the diagonal k sampler (no CUDA in this file).
//
// Once per (m, sigma, l, d, r): the limbs of r and d, the Barrett reciprocal
// mu = floor(2^(64 k) / r), quotient and remainder of 2^l d by r, 2^l d / r in fixed point, and
// the top of r as a double-double (diagk.cuh). Shared by the CUDA
// library (qb200_diagk.cu) and the test-only CPU twin of tests/hostsim.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "bigint.hpp"
#include "diagk.cuh"

namespace qb200 {

struct DiagKHost {
  std::vector<uint32_t> r, d, mu, rho, dq, psi;  // QB_DIAGK_PAD zero limbs, the number, QB_DIAGK_PAD zero limbs
  DiagKConst c;
};

// x as n limbs with the zero limbs mul_columns reads around a constant operand.
inline std::vector<uint32_t> limbs32_padded(const BigUInt& x, size_t n, size_t pad = QB_DIAGK_PAD) {
  std::vector<uint32_t> out(n + 2 * pad, 0);
  for (size_t i = 0; i < n; i++) {
    const size_t w = i / 2;
    if (w < x.w.size()) out[pad + i] = (uint32_t)(x.w[w] >> (32 * (i % 2)));
  }
  return out;
}

// 0, or a negative code with *err set. The pointers of h->c refer to the vectors of *h (host
// side); the CUDA library replaces them by device copies.
inline int diagk_prepare(uint32_t m, uint32_t sigma, uint32_t l, const uint8_t* d_be, size_t d_len,
                         const uint8_t* r_be, size_t r_len, DiagKHost* h, std::string* err) {
  if (!d_be || !r_be) {
    *err = "null argument";
    return -1;
  }
  const BigUInt d = BigUInt::from_bytes_be(d_be, d_len), r = BigUInt::from_bytes_be(r_be, r_len);
  if (r.is_zero() || d.is_zero() || BigUInt::cmp(d, r) >= 0) {
    *err = "diagonal k sampler: need 0 < d < r";
    return -2;
  }
  if (m < 32 || l == 0 || l > m + sigma || r.bit_length() > m || m + sigma > (1u << 20)) {
    *err = "diagonal k sampler: need m >= 32, 0 < l <= m + sigma, r < 2^m";
    return -3;
  }
  const uint32_t k = (uint32_t)((r.bit_length() + 31) / 32);
  h->r = limbs32_padded(r, k);
  h->d = limbs32_padded(d, k);
  BigUInt q, rem;
  BigUInt::divmod(BigUInt::pow2(64ull * k), r, q, rem);
  h->mu = limbs32_padded(q, k + 2);
  const uint32_t wl = (l + 31) / 32;
  BigUInt dq, rho;
  BigUInt::divmod(d.shl(l), r, dq, rho);  // 2^l d = dq r + rho
  h->rho = limbs32_padded(rho, k);
  h->dq = limbs32_padded(dq, wl);
  BigUInt psi, psi_rem;
  BigUInt::divmod(d.shl((size_t)l + 32 * (size_t)(k + 4)), r, psi, psi_rem);  // 2^l d / r, k + 4 fractional limbs
  h->psi = limbs32_padded(psi, k + 4 + wl);
  DiagKConst& c = h->c;
  c.m = m;
  c.sigma = sigma;
  c.l = l;
  c.n = m + sigma;
  c.k = k;
  c.wj = (c.n + 31) / 32;
  c.wl = (l + 31) / 32;
  c.r = h->r.data() + QB_DIAGK_PAD;
  c.d = h->d.data() + QB_DIAGK_PAD;
  c.mu = h->mu.data() + QB_DIAGK_PAD;
  c.rho = h->rho.data() + QB_DIAGK_PAD;
  c.dq = h->dq.data() + QB_DIAGK_PAD;
  c.psi = h->psi.data() + QB_DIAGK_PAD;
  c.r_top = limbs_top_dd<1>(c.r, k - 1);
  c.force_exact = 0;
  {
    const char* fp = getenv("QB200_DIAGK_FULL_PRODUCT");
    c.full_product = (fp && *fp == '1') ? 1 : 0;
    const char* ef = getenv("QB200_DIAGK_EXACT_FRACTION");
    c.exact_fraction = (ef && *ef == '1') ? 1 : 0;
  }
  return 0;
}

}  // namespace qb200
