// hostconst.cpp -- see hostconst.hpp.
#include "hostconst.hpp"

#include <cmath>
#include <cstring>
#include <vector>

#include "bigint.hpp"

namespace qb200 {

namespace {

const size_t kMpfrPrec = 192;  // PRECISION, src/common.h:13

double bf_to_double(const BigFloat& v) {
  if (v.mant.is_zero()) return 0.0;
  const BigFloat r = v.rounded(53);
  const uint64_t mant = r.mant.w.empty() ? 0 : r.mant.w[0];  // <= 2^53
  return std::ldexp((double)mant, (int)r.exp);
}

// Exact a - b; returns magnitude, sets neg.
BigFloat bf_sub(const BigFloat& a, const BigFloat& b, bool* neg) {
  const long e = a.exp < b.exp ? a.exp : b.exp;
  const BigUInt am = a.mant.shl((size_t)(a.exp - e));
  const BigUInt bm = b.mant.shl((size_t)(b.exp - e));
  if (BigUInt::cmp(am, bm) >= 0) {
    *neg = false;
    return BigFloat(BigUInt::sub(am, bm), e);
  }
  *neg = true;
  return BigFloat(BigUInt::sub(bm, am), e);
}

BigFloat bf_from_double(double x) {  // x >= 0
  if (x == 0.0) return BigFloat();
  int e;
  const double f = std::frexp(x, &e);  // x = f * 2^e, f in [0.5, 1)
  const uint64_t mant = (uint64_t)std::ldexp(f, 53);
  return BigFloat(BigUInt(mant), (long)e - 53);
}

// Non-negative big float -> double-double (hi = nearest double, lo = nearest
// double of the remainder).
DD bf_to_dd(const BigFloat& v, bool negative = false) {
  DD r;
  r.hi = bf_to_double(v);
  bool neg = false;
  const BigFloat rem = bf_sub(v, bf_from_double(r.hi), &neg);
  r.lo = bf_to_double(rem);
  if (neg) r.lo = -r.lo;
  if (negative) {
    r.hi = -r.hi;
    r.lo = -r.lo;
  }
  return r;
}

// a / b as a 256-bit-accurate big float (for plain rational constants).
BigFloat bf_ratio(const BigUInt& a, long ea, const BigUInt& b) {
  return BigFloat::div_rounded(a, ea, b, 256);
}

}  // namespace

namespace {
int host_consts_compute_uncached(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d_be,
                                 size_t d_len, const uint8_t* r_be, size_t r_len, HostConsts* out);

// The constants are a pure function of (m, l, sigma, d, r), and a generator client asks for them
// thousands of times with the same parameters (one slice per call): remember the last answer
// per thread (18 us of big-integer division at m = 2048 otherwise, a sixth of a single-slice call).
struct LastConsts {
  bool valid = false;
  uint32_t m = 0, l = 0, sigma = 0;
  std::vector<uint8_t> d, r;
  HostConsts value;
};
thread_local LastConsts g_last;
}  // namespace

int host_consts_compute(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d_be,
                        size_t d_len, const uint8_t* r_be, size_t r_len, HostConsts* out) {
  LastConsts& c = g_last;
  if (c.valid && c.m == m && c.l == l && c.sigma == sigma && c.d.size() == d_len && c.r.size() == r_len &&
      0 == memcmp(c.d.data(), d_be, d_len) && 0 == memcmp(c.r.data(), r_be, r_len)) {
    *out = c.value;
    return 0;
  }
  const int rc = host_consts_compute_uncached(m, l, sigma, d_be, d_len, r_be, r_len, out);
  if (rc == 0) {
    c.m = m;
    c.l = l;
    c.sigma = sigma;
    c.d.assign(d_be, d_be + d_len);
    c.r.assign(r_be, r_be + r_len);
    c.value = *out;
    c.valid = true;
  }
  return rc;
}

namespace {
int host_consts_compute_uncached(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d_be,
                                 size_t d_len, const uint8_t* r_be, size_t r_len, HostConsts* out) {
  if (m == 0 || m > 65536 || l == 0 || l > 65536 || sigma > 65536) return -3;
  const BigUInt d = BigUInt::from_bytes_be(d_be, d_len);
  const BigUInt r = BigUInt::from_bytes_be(r_be, r_len);
  if (d.is_zero() || r.is_zero()) return -1;
  if (d.bit_length() > m || r.bit_length() > m) return -2;

  out->m = m;
  out->l = l;
  out->sigma = sigma;

  // K_sigma = ceil(rnd(rnd(-2^sigma * d) / r))        src/probability.cpp:165-170
  //         = -floor(rnd(rnd(2^sigma d) / r))         (round-to-nearest is symmetric)
  {
    const BigFloat t = BigFloat(d, (long)sigma).rounded(kMpfrPrec);
    const BigFloat q = BigFloat::div_rounded(t.mant, t.exp, r, kMpfrPrec);
    const BigUInt k = q.floor_int();
    out->kappa = bf_to_dd(BigFloat(k, -(long)sigma), /*negative=*/true);
  }
  // quick: theta_r * d / r with both operations rounded  src/probability.cpp:302-304
  {
    const BigFloat t = BigFloat(d, 0).rounded(kMpfrPrec);
    BigFloat q = BigFloat::div_rounded(t.mant, t.exp, r, kMpfrPrec);
    out->kappa_q = bf_to_dd(q, /*negative=*/true);
    // the same Q serves every sigma: rnd(rnd(2^sigma d) / r) == Q * 2^sigma exactly
    while (q.mant.bit_length() > kMpfrPrec) {  // rounding carried into bit 192
      q.mant = q.mant.shr(1);
      q.exp += 1;
    }
    for (int i = 0; i < 3; i++) out->q_mant[i] = (size_t)i < q.mant.w.size() ? q.mant.w[i] : 0;
    out->q_exp = (int)q.exp;
  }
  // Q = rnd(2^(m+l) / r); C = ceil(Q), N = floor(Q)     src/probability.cpp:216-220,
  //                                                     src/linear_probability.cpp:194-197
  {
    const BigFloat q = BigFloat::div_rounded(BigUInt(1), (long)(m + l), r, kMpfrPrec);
    const BigUInt c = q.ceil_int();
    const BigUInt n = q.floor_int();
    out->c_over_L = bf_to_dd(BigFloat(c, -(long)l));
    out->n_over_L = bf_to_dd(BigFloat(n, -(long)l));
    // mpfr_add_ui(tmp, N, 1): rounded to 192 bits       src/linear_probability.cpp:200,217
    const BigFloat n1 = BigFloat(BigUInt::add(n, BigUInt(1)), 0).rounded(kMpfrPrec);
    out->n1_over_L = bf_to_dd(BigFloat(n1.mant, n1.exp - (long)l));
  }
  // beta = 2^(l+m) mod r (exact)                        src/linear_probability.cpp:190-192
  {
    BigUInt q, beta;
    BigUInt::divmod(BigUInt::pow2((uint64_t)m + l), r, q, beta);
    out->beta_m = bf_to_dd(BigFloat(beta, -(long)m));
    out->rbeta_m = bf_to_dd(BigFloat(BigUInt::sub(r, beta), -(long)m));
  }
  out->r_m = bf_to_dd(BigFloat(r, -(long)m));
  out->d_m = bf_to_dd(BigFloat(d, -(long)m));
  out->rho = bf_to_dd(bf_ratio(BigUInt(1), (long)m, r));
  return 0;
}
}  // namespace

// ---------------------------------------------------------------------------
// 2^(i/n) table. 256-bit fixed point; g = 2^(1/n) from the exponential series,
// then table[i] = table[i-1] * g. The accumulated error is below n * 2^-254.
// The reference forms the same grid by repeated 192-bit multiplication with
// exp2(1/dimension) (src/distribution_slice_compute.cpp:110-113, 196-212).
// ---------------------------------------------------------------------------
namespace {

const size_t kFix = 256;

BigUInt fix_mul(const BigUInt& a, const BigUInt& b) { return BigUInt::mul(a, b).shr(kFix); }

BigUInt fix_ln2() {
  // floor(ln 2 * 2^320)
  static const uint64_t limbs[5] = {0xe7b876206debac98ull, 0x8a0d175b8baafa2bull,
                                    0x40f343267298b62dull, 0xc9e3b39803f2f6afull,
                                    0xb17217f7d1cf79abull};
  BigUInt v;
  v.w.assign(limbs, limbs + 5);
  return v.shr(320 - kFix);
}

}  // namespace

void exp2_table_dd(uint32_t n, DD* table) {
  // The reference's step is the DOUBLE 1.0 / dimension (src/distribution_slice_compute.cpp:48,
  // 110-113: mpfr_set_d(pow_2step, step); exp2), exact only for powers of two; its grid is
  // (2^step)^i. Reproduce that: t = ln2 * fl(1/n) with fl(1/n) taken exactly.
  const double step = (double)1 / (double)n;
  int e2;
  const double fr = std::frexp(step, &e2);                 // step = fr * 2^e2, fr in [0.5, 1)
  const uint64_t mant = (uint64_t)std::ldexp(fr, 53);      // exact 53-bit integer
  // step in 256-bit fixed point: mant * 2^(e2 - 53 + kFix)
  const BigUInt step_fix = BigUInt(mant).shl((size_t)((long)kFix + e2 - 53));
  const BigUInt t = fix_mul(fix_ln2(), step_fix);
  BigUInt q, rem;
  const BigUInt one = BigUInt::pow2(kFix);
  BigUInt g = one, term = one;
  for (uint64_t k = 1; k < 200; k++) {
    term = fix_mul(term, t);
    BigUInt::divmod(term, BigUInt(k), q, rem);
    term = q;
    if (term.is_zero()) break;
    g = BigUInt::add(g, term);
  }
  const bool pow2 = (n & (n - 1)) == 0;
  BigUInt cur = one;
  for (uint32_t i = 0; i <= n; i++) {
    if (i == n && pow2) cur = BigUInt::pow2(kFix + 1);  // exactly 2 when the step is exact
    table[i] = bf_to_dd(BigFloat(cur, -(long)kFix));
    cur = fix_mul(cur, g);
  }
}

}  // namespace qb200
