// textparse.cuh -- one "%Lg" token -> x87 long double, bit for bit as glibc's strtold.
//
// Replaces the per-cell fscanf(file, "%Lg\n", &norm_matrix[i]) of the reference's
// slice importers (src/distribution_slice_import_export.cpp:38-50,
// src/linear_distribution_slice_import_export.cpp:33-45,
// src/diagonal_distribution_slice_import_export.cpp:38-50): every consumer of a
// stored distribution (sampling, estimation, solving, comparison, plotting)
// parses ~1e8 numbers on start-up.
//
// Semantics restated: optional sign; decimal digits with an optional point;
// optional exponent; or inf / infinity / nan (any case). The result is the x87
// value nearest to the decimal number, ties to even; overflow gives infinity,
// underflow a denormal or zero (correct rounding at the denormal position).
//
// Method. The first 28 significant digits form D < 2^96; value = D * 10^k. With
// the 192-bit table entry T of 10^k (textfmt.cuh), P = D * T is a 288-bit
// product; the mantissa is its top bits. If T is exact (0 <= k <= 82) and no
// digit was dropped, rounding P is exact. Otherwise the true product lies in
// (P, P + err): if rounding P and P + err agree the result is decided, else the
// decision is made by exact integer arithmetic (cmp_pow5). Tokens with more than
// 28 significant digits that fall into that undecided band are reported as
// unsupported (the exporter writes 24).
#pragma once

#include "textfmt.cuh"

namespace qb200 {
namespace text {

constexpr int MAX_SIG_DIGITS = 28;  // 10^28 < 2^96

enum ParseStatus : uint32_t {
  PARSE_OK = 0,
  PARSE_MALFORMED = 1,    // not a number the reference's fscanf("%Lg") would accept
  PARSE_UNSUPPORTED = 2,  // hexadecimal float, or > 28 digits on a rounding boundary
};

struct Decimal {
  uint32_t d[3];     // D, little-endian limbs
  int32_t k;         // value = D * 10^k
  uint32_t neg;
  uint32_t sticky;   // non-zero digits beyond the 28th were dropped
  uint32_t special;  // 0 number, 2 inf, 3 nan
  int32_t ndig;      // significant digits in D
};

QT_HD bool is_space(uint32_t c) { return c == ' ' || (c >= 9 && c <= 13); }
QT_HD uint32_t lower(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

QT_HD uint32_t pow10_u32(int k) {  // 10^k, 0 <= k <= 9
  uint32_t r = 1;
#pragma unroll
  for (int j = 0; j < 9; j++)
    if (j < k) r *= 10u;
  return r;
}

// D = D * mul + add on three limbs.
QT_HD void muladd96(uint32_t (&d)[3], uint32_t mul, uint32_t add) {
  uint32_t carry = add;
#pragma unroll
  for (int j = 0; j < 3; j++) d[j] = mulhi_lo(d[j], mul, carry, &carry);
}

// Parse ONE number starting at s[0]; it ends at the first white-space byte or after
// `avail` bytes (the end of the text). Single pass. Digits are gathered nine at a time
// in a 32-bit accumulator (one multiply-add per digit) and the groups are combined
// into the 96-bit D at the end. Returns a ParseStatus; *used = length of the token.
template <class Ptr>
QT_HD uint32_t parse_number(Ptr s, long avail, Decimal* out, int* used) {
  long i = 0;
  uint32_t c = avail > 0 ? (uint32_t)(unsigned char)s[0] : 32u;
#define QT_NEXT() (c = (++i < avail) ? (uint32_t)(unsigned char)s[i] : 32u)
  out->neg = 0;
  out->sticky = 0;
  out->special = 0;
  out->k = 0;
  out->ndig = 0;
  out->d[0] = out->d[1] = out->d[2] = 0;
  if (c == '+' || c == '-') {
    out->neg = c == '-';
    QT_NEXT();
  }
  uint32_t status = PARSE_OK;
  if (lower(c) == 'i' || lower(c) == 'n') {
    // inf, infinity, nan, nan(chars): gather up to the white space
    char w[12];
    int m = 0;
    char last = 0;
    while (!is_space(c)) {
      if (m < 12) w[m] = (char)lower(c);
      last = (char)c;
      m++;
      QT_NEXT();
    }
    const char* inf = "infinity";
    bool ok;
    if (w[0] == 'i') {
      ok = m == 3 || m == 8;
      for (int j = 0; ok && j < m; j++) ok = w[j] == inf[j];
      out->special = 2;
    } else {
      ok = m >= 3 && w[1] == 'a' && w[2] == 'n' && (m == 3 || (w[3] == '(' && last == ')'));
      out->special = 3;
    }
    *used = (int)(i > 0x7fffffffL ? 0x7fffffffL : i);
    return ok ? PARSE_OK : PARSE_MALFORMED;
  }
  uint32_t qa = 0, qb = 0, qc = 0, cur = 0;  // full groups of nine digits (oldest first), current group
  int cnt = 0, nsig = 0, adj = 0;
  uint32_t sticky = 0;
  bool point = false, any = false, hex = false;
  if (c == '0' && i + 1 < avail && lower((uint32_t)(unsigned char)s[i + 1]) == 'x') hex = true;
  for (;;) {
    const uint32_t v = c - '0';
    if (v <= 9u) {
      any = true;
      if (nsig < MAX_SIG_DIGITS) {
        if ((nsig | (int)v) != 0) {  // significant (not a leading zero)
          cur = cur * 10u + v;
          nsig++;
          if (++cnt == 9) {
            qa = qb;
            qb = qc;
            qc = cur;
            cur = 0;
            cnt = 0;
          }
        }
        if (point) adj--;
      } else {
        sticky |= v;
        if (!point) adj++;
      }
    } else if (c == '.' && !point) {
      point = true;
    } else {
      break;
    }
    QT_NEXT();
  }
  if (!any) status = PARSE_MALFORMED;
  int ex = 0;
  if (status == PARSE_OK && lower(c) == 'e') {
    QT_NEXT();
    bool eneg = false;
    if (c == '+' || c == '-') {
      eneg = c == '-';
      QT_NEXT();
    }
    if (c - '0' > 9u) status = PARSE_MALFORMED;
    while (c - '0' <= 9u) {
      if (ex < 100000) ex = ex * 10 + (int)(c - '0');
      QT_NEXT();
    }
    if (eneg) ex = -ex;
  }
  if (!is_space(c)) {  // trailing garbage: skip to the end of the token
    status = hex ? PARSE_UNSUPPORTED : PARSE_MALFORMED;
    while (!is_space(c) && i < 0x7fffffffL) QT_NEXT();
  }
#undef QT_NEXT
  *used = (int)(i > 0x7fffffffL ? 0x7fffffffL : i);
  // D = ((qa * 10^9 + qb) * 10^9 + qc) * 10^cnt + cur
  uint32_t d[3] = {qa, 0, 0};
  muladd96(d, 1000000000u, qb);
  muladd96(d, 1000000000u, qc);
  muladd96(d, pow10_u32(cnt), cur);
  out->d[0] = d[0];
  out->d[1] = d[1];
  out->d[2] = d[2];
  out->ndig = nsig;
  out->sticky = sticky ? 1u : 0u;
  out->k = ex + adj;
  return status;
}

// Round the 288-bit P (9 limbs, top bit 287 set) after dropping `drop` low bits
// (drop >= 224), ties to even. Returns the kept integer (< 2^64) + rounding;
// *carry = 1 if it reached 2^64.
QT_HD uint64_t round288(const uint32_t (&P)[9], int drop, uint32_t* carry) {
  *carry = 0;
  if (drop > 288) return 0;
  // kept = P >> drop
  uint64_t kept = 0;
  uint32_t half = 0, rest = 0;
  const int hb = drop - 1;  // position of the rounding bit
#pragma unroll
  for (int w = 0; w < 9; w++) {
    const int lo = 32 * w;  // bit position of this limb
    // bits of this limb at or above `drop` go to kept
    if (lo + 31 >= drop) {
      if (lo >= drop) {
        const int sh = lo - drop;
        if (sh < 64) kept |= (uint64_t)P[w] << sh;
      } else {
        kept |= (uint64_t)(P[w] >> (drop - lo));
      }
    }
    if (hb >= lo && hb <= lo + 31) {
      half = (P[w] >> (hb - lo)) & 1u;
      rest |= P[w] & ((1u << (hb - lo)) - 1u);
    } else if (lo + 31 < hb) {
      rest |= P[w];
    }
  }
  if (half && (rest || (kept & 1ULL))) {
    kept++;
    if (kept == 0) *carry = 1;
  }
  return kept;
}

// D * 10^k -> x87 (mant, se without sign). Returns 0 decided, 1 undecided (then
// *mant holds the truncated mantissa M and *q its exponent for the exact decision:
// value vs (2M + 1) * 2^(q - 1)), 2 unsupported.
QT_HD uint32_t decimal_to_x87(const Decimal& dec, const Pow10Entry* __restrict__ tab,
                              uint64_t* mant, uint32_t* se, int* q_out, bool force_band = false) {
  uint32_t d0 = dec.d[0], d1 = dec.d[1], d2 = dec.d[2];
  if ((d0 | d1 | d2) == 0) {
    *mant = 0;
    *se = 0;
    return 0;
  }
  // decimal magnitude guards (table range; far outside the x87 range anyway)
  const int x10 = dec.k + dec.ndig - 1;  // decimal exponent of the first digit
  if (x10 > 4940) {
    *mant = 1ULL << 63;
    *se = 0x7fff;
    return 0;
  }
  if (x10 < -4965) {
    *mant = 0;
    *se = 0;
    return 0;
  }
  // normalise D to 96 bits
  int lz;
  if (d2)
    lz = clz64(d2) - 32;
  else if (d1)
    lz = 32 + clz64(d1) - 32;
  else
    lz = 64 + clz64(d0) - 32;
  {
    const int ws = lz >> 5, bs = lz & 31;
    uint32_t a0 = d0, a1 = d1, a2 = d2;
    if (ws == 1) {
      a2 = a1; a1 = a0; a0 = 0;
    } else if (ws == 2) {
      a2 = a0; a1 = 0; a0 = 0;
    }
    if (bs) {
      a2 = (a2 << bs) | (a1 >> (32 - bs));
      a1 = (a1 << bs) | (a0 >> (32 - bs));
      a0 <<= bs;
    }
    d0 = a0; d1 = a1; d2 = a2;
  }
  const Pow10Entry* ent = tab + (dec.k - K_MIN);
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
  // P = Dn * T (288 bits)
  uint32_t P[9];
  {
    const uint32_t dn[3] = {d0, d1, d2};
#pragma unroll
    for (int i = 0; i < 9; i++) P[i] = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      uint32_t carry = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const uint64_t t = (uint64_t)T[i] * dn[j] + P[i + j] + carry;
        P[i + j] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
      }
      P[6 + j] = carry;
    }
  }
  // value = P * 2^(te2 - 191 - lz); make bit 287 the top bit
  int pe = te2 - 191 - lz;  // exponent of P's unit
  if (!(P[8] >> 31)) {
#pragma unroll
    for (int i = 8; i > 0; i--) P[i] = (P[i] << 1) | (P[i - 1] >> 31);
    P[0] <<= 1;
    pe -= 1;
  }
  // value in [2^(pe+287), 2^(pe+288)); biased exponent if normal:
  const int e = pe + 287 + 16383;
  int drop = 224;
  if (e < 1) drop = 224 + (1 - e);  // denormal: fewer mantissa bits
  const bool exact = (dec.k >= 0 && dec.k <= K_EXACT_MAX) && !dec.sticky;
  uint32_t carry = 0;
  uint64_t M;
  uint32_t status = 0;
  if (drop == 224 && !dec.sticky && !force_band) {
    // The common case (a normal number, no dropped digits) without the generic bit loops:
    // 64 kept bits, the rounding bit is the top bit of P[6].
    const uint64_t kept = ((uint64_t)P[8] << 32) | P[7];
    const uint32_t half = P[6] >> 31;
    uint32_t rest = P[6] & 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 6; i++) rest |= P[i];
    bool up;
    if (exact) {
      up = half && (rest || (kept & 1ULL));
    } else {
      // true remainder in (r, r + 2^98), strictly above r: undecided iff r in [half - 2^98, half);
      // r >= half rounds up (r == half is not a tie: the true value lies above it)
      const bool below = P[6] == 0x7fffffffu && P[5] == 0xffffffffu && P[4] == 0xffffffffu &&
                         (P[3] >> 2) == 0x3fffffffu;
      if (below) {
        *mant = kept;
        *q_out = pe + 224;
        return 1;
      }
      up = half != 0;
    }
    M = kept + (up ? 1ULL : 0ULL);
    if (up && M == 0) carry = 1;
  } else {
  M = round288(P, drop, &carry);
  if (!exact || force_band) {
    // upper end of the interval the true product lies in
    uint32_t Q[9];
    const int eb = dec.sticky ? 200 : 98;  // err < 2^eb
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const uint32_t add = (i == (eb >> 5)) ? (1u << (eb & 31)) : 0u;
      const uint64_t t = (uint64_t)P[i] + add + c;
      Q[i] = (uint32_t)t;
      c = (uint32_t)(t >> 32);
    }
    // If P + err reached 2^288, P's bits 200..287 are all ones and P itself already
    // rounds up to 2^(288 - drop): decided.
    uint32_t carry2 = carry;
    uint64_t M2 = M;
    if (!c) M2 = round288(Q, drop, &carry2);
    if (M2 != M || carry2 != carry || force_band) {
      if (dec.sticky) return 2;
      // truncated mantissa and its exponent for the exact decision
      uint32_t cz = 0;
      uint32_t Z[9];
#pragma unroll
      for (int i = 0; i < 9; i++) Z[i] = P[i];
      // clear the rounding bit and below, then round288 = truncation
      if (drop <= 288) {
        const int hb = drop - 1;
#pragma unroll
        for (int w = 0; w < 9; w++) {
          const int lo = 32 * w;
          if (lo + 31 <= hb)
            Z[w] = 0;
          else if (lo <= hb)
            Z[w] &= ~((2u << (hb - lo)) - 1u);
        }
        *mant = round288(Z, drop, &cz);
      } else {
        *mant = 0;
      }
      *q_out = pe + drop;
      status = 1;
      return status;
    }
  }
  }
  // assemble
  int ee = e < 1 ? 0 : e;
  if (carry) {  // mantissa overflowed to 2^64 (normal case only)
    M = 1ULL << 63;
    ee += 1;
  }
  if (ee == 0 && (M >> 63)) ee = 1;  // rounded up into the normal range
  if (ee >= 0x7fff) {
    M = 1ULL << 63;
    ee = 0x7fff;
  }
  *mant = M;
  *se = (uint32_t)ee;
  return 0;
}

// Finish an undecided conversion exactly: value = D * 10^k against the midpoint
// (2M + 1) * 2^(q - 1), M the truncated mantissa at exponent q.
QT_HD_NOINLINE void finish_exact(const Decimal& dec, uint64_t M, int q, uint32_t* scratch,
                                 uint64_t* mant, uint32_t* se) {
  // mid = 2M + 1 (up to 65 bits)
  uint32_t mid[3];
  mid[0] = (uint32_t)(M << 1) | 1u;
  mid[1] = (uint32_t)(M >> 31);
  mid[2] = (uint32_t)(M >> 63);
  int sgn;
  if (dec.k >= 0) {
    // D * 5^k * 2^k  vs  mid * 2^(q-1)
    sgn = cmp_pow5(dec.d, dec.k, mid, q - 1 - dec.k, scratch);
  } else {
    // D  vs  mid * 5^|k| * 2^(q - 1 + |k|)
    sgn = -cmp_pow5(mid, -dec.k, dec.d, -(q - 1 - dec.k), scratch);
  }
  const bool up = sgn > 0 || (sgn == 0 && (M & 1ULL));
  uint64_t R = M + (up ? 1ULL : 0ULL);
  // exponent field: value = R * 2^q
  int ee;
  if (up && R == 0) {  // 2^64
    R = 1ULL << 63;
    q += 1;
  }
  if (R == 0) {
    *mant = 0;
    *se = 0;
    return;
  }
  if (R >> 63) {
    ee = q + 63 + 16383;
    if (ee < 1) ee = 1;  // q is never below the denormal position
  } else {
    ee = 0;  // denormal: q == -16445
  }
  if (ee >= 0x7fff) {
    R = 1ULL << 63;
    ee = 0x7fff;
  }
  *mant = R;
  *se = (uint32_t)ee;
}

}  // namespace text
}  // namespace qb200
