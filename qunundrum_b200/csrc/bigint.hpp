// bigint.hpp -- minimal unsigned big integers and binary big floats.
//
// Host-side only. Used once per distribution to turn the m-bit integers d and r
// of the reference's Parameters (src/parameters.h:33-114) into the handful of
// double-double constants the kernels consume (see hostconst.hpp). The
// reference does this arithmetic with GMP/MPFR inside every integrand call
// (e.g. src/probability.cpp:165-170, 216-220); here it is hoisted out of the
// hot path, so a small schoolbook implementation is all that is needed and the
// library has no GMP dependency.
#pragma once

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace qb200 {

typedef unsigned __int128 u128;

class BigUInt {
 public:
  std::vector<uint64_t> w;  // little-endian limbs, no leading zero limbs

  BigUInt() {}
  explicit BigUInt(uint64_t v) {
    if (v) w.push_back(v);
  }

  static BigUInt pow2(uint64_t e) {
    BigUInt r;
    r.w.assign(e / 64 + 1, 0);
    r.w[e / 64] = uint64_t(1) << (e % 64);
    return r;
  }

  // Big-endian magnitude bytes (the layout mpz_export(.., 1, 1, 1, 0, ..) gives).
  static BigUInt from_bytes_be(const uint8_t* p, size_t n) {
    BigUInt r;
    r.w.assign((n + 7) / 8, 0);
    for (size_t i = 0; i < n; i++) {
      const size_t k = n - 1 - i;  // significance of byte i, in bytes
      r.w[k / 8] |= uint64_t(p[i]) << (8 * (k % 8));
    }
    r.trim();
    return r;
  }

  void trim() {
    while (!w.empty() && w.back() == 0) w.pop_back();
  }
  bool is_zero() const { return w.empty(); }

  size_t bit_length() const {
    if (w.empty()) return 0;
    return 64 * (w.size() - 1) + (64 - __builtin_clzll(w.back()));
  }
  bool bit(size_t i) const {
    const size_t k = i / 64;
    return k < w.size() && ((w[k] >> (i % 64)) & 1);
  }
  // True if any bit strictly below position i is set.
  bool any_below(size_t i) const {
    const size_t k = i / 64;
    for (size_t j = 0; j < std::min(k, w.size()); j++)
      if (w[j]) return true;
    if (k < w.size() && (i % 64)) return (w[k] & ((uint64_t(1) << (i % 64)) - 1)) != 0;
    return false;
  }

  static int cmp(const BigUInt& a, const BigUInt& b) {
    if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
    for (size_t i = a.w.size(); i-- > 0;)
      if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
  }

  BigUInt shl(size_t s) const {
    if (w.empty()) return BigUInt();
    BigUInt r;
    const size_t ls = s / 64, bs = s % 64;
    r.w.assign(w.size() + ls + 1, 0);
    for (size_t i = 0; i < w.size(); i++) {
      r.w[i + ls] |= w[i] << bs;
      if (bs) r.w[i + ls + 1] |= w[i] >> (64 - bs);
    }
    r.trim();
    return r;
  }
  BigUInt shr(size_t s) const {
    const size_t ls = s / 64, bs = s % 64;
    if (ls >= w.size()) return BigUInt();
    BigUInt r;
    r.w.assign(w.size() - ls, 0);
    for (size_t i = 0; i < r.w.size(); i++) {
      r.w[i] = w[i + ls] >> bs;
      if (bs && i + ls + 1 < w.size()) r.w[i] |= w[i + ls + 1] << (64 - bs);
    }
    r.trim();
    return r;
  }

  static BigUInt add(const BigUInt& a, const BigUInt& b) {
    BigUInt r;
    const size_t n = std::max(a.w.size(), b.w.size());
    r.w.assign(n + 1, 0);
    u128 c = 0;
    for (size_t i = 0; i < n; i++) {
      c += (i < a.w.size() ? a.w[i] : 0);
      c += (i < b.w.size() ? b.w[i] : 0);
      r.w[i] = (uint64_t)c;
      c >>= 64;
    }
    r.w[n] = (uint64_t)c;
    r.trim();
    return r;
  }
  // a - b, requires a >= b.
  static BigUInt sub(const BigUInt& a, const BigUInt& b) {
    if (cmp(a, b) < 0) throw std::logic_error("BigUInt::sub underflow");
    BigUInt r;
    r.w.assign(a.w.size(), 0);
    uint64_t borrow = 0;
    for (size_t i = 0; i < a.w.size(); i++) {
      const uint64_t bi = i < b.w.size() ? b.w[i] : 0;
      const u128 t = (u128)a.w[i] - bi - borrow;
      r.w[i] = (uint64_t)t;
      borrow = (uint64_t)(t >> 64) ? 1 : 0;
    }
    r.trim();
    return r;
  }
  static BigUInt mul(const BigUInt& a, const BigUInt& b) {
    if (a.w.empty() || b.w.empty()) return BigUInt();
    BigUInt r;
    r.w.assign(a.w.size() + b.w.size(), 0);
    for (size_t i = 0; i < a.w.size(); i++) {
      u128 c = 0;
      for (size_t j = 0; j < b.w.size(); j++) {
        c += (u128)a.w[i] * b.w[j] + r.w[i + j];
        r.w[i + j] = (uint64_t)c;
        c >>= 64;
      }
      r.w[i + b.w.size()] += (uint64_t)c;
    }
    r.trim();
    return r;
  }
  static BigUInt mul_small(const BigUInt& a, uint64_t b) { return mul(a, BigUInt(b)); }

  // Quotient and remainder (Knuth algorithm D, 64-bit limbs).
  static void divmod(const BigUInt& a, const BigUInt& b, BigUInt& q, BigUInt& r) {
    if (b.w.empty()) throw std::logic_error("BigUInt::divmod by zero");
    if (cmp(a, b) < 0) {
      q = BigUInt();
      r = a;
      return;
    }
    if (b.w.size() == 1) {
      q.w.assign(a.w.size(), 0);
      u128 rem = 0;
      for (size_t i = a.w.size(); i-- > 0;) {
        const u128 cur = (rem << 64) | a.w[i];
        q.w[i] = (uint64_t)(cur / b.w[0]);
        rem = cur % b.w[0];
      }
      q.trim();
      r = BigUInt((uint64_t)rem);
      return;
    }
    const int s = __builtin_clzll(b.w.back());
    const BigUInt v = b.shl(s);
    BigUInt u = a.shl(s);
    const size_t n = v.w.size();
    if (u.w.size() == a.w.size()) u.w.push_back(0);
    while (u.w.size() < a.w.size() + 1) u.w.push_back(0);
    const size_t mq = u.w.size() - n - 1;
    q.w.assign(mq + 1, 0);
    for (size_t j = mq + 1; j-- > 0;) {
      const u128 num = ((u128)u.w[j + n] << 64) | u.w[j + n - 1];
      u128 qhat = num / v.w[n - 1];
      u128 rhat = num % v.w[n - 1];
      while ((qhat >> 64) != 0 ||
             qhat * v.w[n - 2] > ((rhat << 64) | u.w[j + n - 2])) {
        qhat--;
        rhat += v.w[n - 1];
        if ((rhat >> 64) != 0) break;
      }
      // multiply and subtract
      u128 borrow = 0, carry = 0;
      for (size_t i = 0; i < n; i++) {
        const u128 p = qhat * v.w[i] + carry;
        carry = p >> 64;
        const u128 t = (u128)u.w[i + j] - (uint64_t)p - borrow;
        u.w[i + j] = (uint64_t)t;
        borrow = (t >> 64) ? 1 : 0;
      }
      const u128 t = (u128)u.w[j + n] - (uint64_t)carry - borrow;
      u.w[j + n] = (uint64_t)t;
      if (t >> 64) {  // add back
        qhat--;
        u128 c = 0;
        for (size_t i = 0; i < n; i++) {
          c += (u128)u.w[i + j] + v.w[i];
          u.w[i + j] = (uint64_t)c;
          c >>= 64;
        }
        u.w[j + n] += (uint64_t)c;
      }
      q.w[j] = (uint64_t)qhat;
    }
    q.trim();
    u.trim();
    r = u.shr(s);
  }

  std::string to_hex() const {
    if (w.empty()) return "0";
    static const char* H = "0123456789abcdef";
    std::string s;
    for (size_t i = w.size(); i-- > 0;)
      for (int k = 60; k >= 0; k -= 4) s.push_back(H[(w[i] >> k) & 15]);
    const size_t nz = s.find_first_not_of('0');
    return s.substr(nz);
  }
};

// Non-negative binary float: mant * 2^exp (sign tracked by the caller).
struct BigFloat {
  BigUInt mant;
  long exp = 0;

  BigFloat() {}
  BigFloat(const BigUInt& m, long e) : mant(m), exp(e) {}

  // Round to `prec` significant bits, nearest-even (MPFR_RNDN semantics).
  // `sticky` says the true value is slightly above mant * 2^exp.
  BigFloat rounded(size_t prec, bool sticky = false) const {
    const size_t bl = mant.bit_length();
    if (bl <= prec) return *this;  // exact (sticky cannot flip a kept bit)
    const size_t drop = bl - prec;
    BigUInt keep = mant.shr(drop);
    const bool half = mant.bit(drop - 1);
    const bool below = mant.any_below(drop - 1) || sticky;
    if (half && (below || keep.bit(0))) keep = BigUInt::add(keep, BigUInt(1));
    return BigFloat(keep, exp + (long)drop);  // keep may be 2^prec: still fine
  }

  // rnd_prec(a / b) for integers a, b > 0 (times 2^e2), nearest-even.
  static BigFloat div_rounded(const BigUInt& a, long ea, const BigUInt& b, size_t prec) {
    const long need = (long)prec + 3 + (long)b.bit_length() - (long)a.bit_length();
    const size_t s = need > 0 ? (size_t)need : 0;
    BigUInt q, r;
    BigUInt::divmod(a.shl(s), b, q, r);
    return BigFloat(q, ea - (long)s).rounded(prec, !r.is_zero());
  }

  // ceil / floor to an integer (as BigUInt); value must be >= 0.
  BigUInt ceil_int() const {
    if (exp >= 0) return mant.shl((size_t)exp);
    const size_t s = (size_t)(-exp);
    BigUInt q = mant.shr(s);
    if (mant.any_below(s)) q = BigUInt::add(q, BigUInt(1));
    return q;
  }
  BigUInt floor_int() const {
    if (exp >= 0) return mant.shl((size_t)exp);
    return mant.shr((size_t)(-exp));
  }
};

}  // namespace qb200
