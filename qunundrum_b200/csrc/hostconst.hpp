// hostconst.hpp -- per-distribution constants as double-double values.
//
// The reference recomputes, inside every integrand evaluation and at 192 bits
// (or more) of MPFR precision, a few quantities that depend only on
// (m, l, sigma, d, r):
//
//   K_sigma = ceil(-2^sigma d / r)            src/probability.cpp:165-170
//   C       = ceil(2^(m+l) / r)               src/probability.cpp:216-220, 332-336
//   N       = floor(2^(m+l) / r), N + 1       src/linear_probability.cpp:194-200, 217
//   beta    = 2^(l+m) mod r, r - beta         src/linear_probability.cpp:190-192, 244
//   r, d, 2^(m+sigma) / r                     src/probability.cpp:247, src/linear_probability.cpp:133-140,
//                                             src/diagonal_probability.cpp:56-60
//
// The kernels work in the scale-free variable x = alpha / 2^m, so each constant
// is divided by the matching power of two and rounded ONCE, here on the host,
// to a double-double (about 106 bits). The MPFR roundings that decide an
// integer (the 192-bit roundings before ceil / floor, and the rounding of
// N + 1) are reproduced exactly; see DESIGN.md "Host-prepared constants".
#pragma once

#include <cstddef>
#include <cstdint>

namespace qb200 {

struct DD {
  double hi, lo;
};

struct HostConsts {
  uint32_t m, l, sigma;
  DD kappa;      // K_sigma / 2^sigma                       (2D, error-bounded approximation)
  DD kappa_q;    // -rnd(rnd(d) / r)                        (2D, quick approximation)
  DD c_over_L;   // ceil(rnd(2^(m+l) / r)) / 2^l
  DD n_over_L;   // floor(rnd(2^(m+l) / r)) / 2^l
  DD n1_over_L;  // rnd(N + 1) / 2^l
  DD beta_m;     // (2^(l+m) mod r) / 2^m
  DD rbeta_m;    // (r - beta) / 2^m
  DD r_m;        // r / 2^m
  DD d_m;        // d / 2^m
  DD rho;        // 2^m / r
  // Q = rnd(rnd(d) / r) at 192 bits as mant * 2^q_exp (mant < 2^192, little-endian limbs):
  // K_sigma = -floor(Q * 2^sigma) for EVERY sigma (the scaling by 2^sigma is exact in MPFR),
  // which is what the sigma-optimal method needs (src/probability.cpp:20-148, 165-170).
  uint64_t q_mant[3];
  int q_exp;
};

// d, r: big-endian magnitude bytes (what mpz_export(buf, &n, 1, 1, 1, 0, z)
// writes). Returns 0, or a negative error code:
//   -1 r == 0 or d == 0;  -2 r or d does not fit in m bits;  -3 bad m / l / sigma.
int host_consts_compute(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d_be,
                        size_t d_len, const uint8_t* r_be, size_t r_len, HostConsts* out);

// (2^fl(1/n))^i for i = 0..n as double-double, fl(1/n) the double the reference uses as its
// step (equal to 2^(i/n) when n is a power of two); n >= 1.
void exp2_table_dd(uint32_t n, DD* table);

}  // namespace qb200
