// exact.cuh -- the exact samplers: an integer alpha from a region and the pair (j, k) from it.
//
// Replaces, for a batch of samples, what the reference does per sample with MPFR at 3 m bits and
// GMP ("missing" #3):
//   sample_alpha_from_region            src/sample.cpp:78-158    alpha = min + (v mod (max - min))
//   sample_j_from_alpha_r               src/sample.cpp:160-208   j from alpha_r (linear / two-dimensional)
//   sample_j_k_from_alpha_d             src/sample.cpp:210-273   j from (alpha_d, k)
//   sample_j_k_from_alpha_d_r           src/sample.cpp:275-352   (j, k) from (alpha_d, alpha_r)
//   sample_j_from_diagonal_alpha_r      src/sample.cpp:354-410   j modulo 2^(m + sigma)
// and therefore the second halves of linear_distribution_sample_alpha
// (src/linear_distribution.cpp:668-724), diagonal_distribution_sample_alpha_r / _j_eta
// (src/diagonal_distribution.cpp:355-472) and distribution_sample_pair_j_k
// (src/distribution.cpp:615-680; lattice_alpha_map, src/lattice_sample.cpp:104-135, sits between
// the two halves and stays the reference's fpLLL code: out of scope).
//
// What the reference computes and how it is done here, per sample:
//   * the region's integer bounds round(2^|log alpha|) (src/sample.cpp:97-124: mpfr_exp2 at
//     3 ceil(|max log alpha|) bits, mpfr_round). |log alpha| = e + i / D with the slice's integer e
//     and the region's index i (src/distribution_slice.cpp:130-165, the linear and diagonal twins):
//     the bound is round(2^e 2^(i/D)) from a table of 2^(i/D) in fixed point with 128 bits more
//     than the largest e needs (exact_host.hpp). The reference's own rounding of 2^(e + i/D) to
//     3 ceil(e + 1) bits moves the value by less than 2^(-2 e); a bound is therefore the
//     reference's unless 2^(e + i/D) lies within 2^-64 of a half-integer, which is detected
//     (QB_EXACT_AMBIGUOUS; never observed). For e < 31 (toy parameters: m < 61) that argument does not
//     hold and both of the reference's roundings are carried out as it does them (exact_bound).
//   * v mod (max - min) with v the (bits(max - min) + 72) / 8 bytes random_generate_mpz
//     (src/random.c:158-181) reads, big-endian: v exceeds the modulus by at most 73 bits, so the
//     division is Knuth's algorithm D with at most four 32-bit quotient digits (exact_mod).
//   * j = floor(inv alpha_r / 2^kappa_r) + t_r 2^(n - kappa_r) mod 2^n with inv =
//     (r / 2^kappa_r)^-1 mod 2^n (a host constant, exact_host.hpp; the reference calls mpz_invert
//     per sample): one truncated product; negative alpha_r in two's complement (mpz_div floors,
//     mpz_mod is non-negative: bits [kappa_r, kappa_r + n) of the two's complement product).
//   * k = floor((alpha_d - d j) / 2^m) mod 2^l: bits [m, m + l) of alpha_d - d j in two's
//     complement, from the low m + l bits of d j: a second truncated product.
// 32-bit limbs, products by mul_columns_wide (diagk.cuh). Per-sample numbers are strided (S = 1 on the
// host, the CTA's thread count on the device, see diagk.cuh). Integer arithmetic throughout: the
// results are the reference's integers bit for bit.
//
// __host__ __device__ so that tests/hostsim runs exactly this code on the CPU.
#pragma once

#include "diagk.cuh"

namespace qb200 {

#define QB_EXACT_OK 0
#define QB_EXACT_LENGTH 1       // the bytes given are not what random_generate_mpz reads for this modulus
#define QB_EXACT_AMBIGUOUS 2    // a bound within 2^-64 of a half-integer: only the reference's own rounding decides
#define QB_EXACT_UNSUPPORTED 3  // e < 8, e + 1 above the table, dimension not a power of two the table holds, max = min

// Guard bits of the table below the rounding position of the largest bound.
#define QB_EXACT_GUARD 128
// Zero limbs the constant operands carry below index 0 and above their last limb (mul_columns_wide
// with up to 8 columns at a time).
#define QB_EXACT_PAD 7

struct ExactConst {
  uint32_t m, l, sigma;
  uint32_t n;          // j is reduced modulo 2^n: m + l (two-dimensional, linear) or m + sigma (diagonal)
  uint32_t kbits;      // bits of k: l (two-dimensional); 0: no k (diagonal)
  uint32_t kappa_d, kappa_r;
  uint32_t wn;         // limbs of j: ceil(n / 32)
  uint32_t wa;         // limbs of |alpha|: ceil((emax + 1) / 32)
  uint32_t wd;         // limbs of d
  uint32_t wk;         // limbs of k
  uint32_t emax;       // bounds up to 2^emax
  uint32_t table_dim;  // D_max, a power of two
  uint32_t table_log;  // log2(D_max)
  uint32_t tw;         // words per table entry
  uint32_t P;          // entry i = 2^(i / D_max) 2^P, truncated (error below 2 units): emax + QB_EXACT_GUARD bits
  // each with QB_EXACT_PAD zero limbs below index 0 and above its last limb (mul_columns_wide):
  const uint32_t* inv_r;  // wn limbs: (r / 2^kappa_r)^-1 mod 2^n
  const uint32_t* inv_d;  // wn limbs: (d / 2^kappa_d)^-1 mod 2^n
  const uint32_t* d;      // wd limbs
  const uint32_t* table;  // table_dim entries of tw words, entry after entry (global memory)
};

// Words of the constants inv_r, inv_d, d with their zero limbs, in this order (shared-memory copy).
QHD uint32_t exact_const_words(const ExactConst& c) { return 2 * (c.wn + 2 * QB_EXACT_PAD) + (c.wd + 2 * QB_EXACT_PAD); }

// One region of a slice: |log alpha| on [e + region / dimension, e + (region + 1) / dimension],
// e = |min_log_alpha| (the slice's coordinate), the sign that of min_log_alpha; the sample's bytes
// are stream[offset, offset + length).
struct ExactRegion {
  int32_t min_log_alpha;
  uint32_t region;
  uint32_t dimension;
  uint32_t length;
  uint64_t offset;
};

// Scratch words per sample (times the stride) of exact_alpha: V (wa + 5), M (wa + 1).
QHD uint32_t exact_alpha_scratch_limbs(const ExactConst& c) { return (c.wa + 5) + (c.wa + 1); }
// ... of the (j, k) functions: a product of up to wn + 3 limbs, a second operand of wn + 1.
QHD uint32_t exact_jk_scratch_limbs(const ExactConst& c) { return (c.wn + 4) + (c.wn + 4); }

// ---- bounds from the table --------------------------------------------------------------

// 32 bits of the entry T (tw words) from bit position `bit` upwards.
QHD uint32_t exact_table_bits(const uint32_t* T, uint32_t tw, uint32_t bit) {
  const uint32_t w = bit >> 5, off = bit & 31u;
  const uint32_t lo = w < tw ? T[w] : 0u;
  if (!off) return lo;
  const uint32_t hi = w + 1 < tw ? T[w + 1] : 0u;
  return (lo >> off) | (hi << (32u - off));
}

// 64 bits of the entry T from bit position `bit` upwards.
QHD uint64_t exact_table_bits64(const uint32_t* T, uint32_t tw, uint32_t bit) {
  return (uint64_t)exact_table_bits(T, tw, bit) | ((uint64_t)exact_table_bits(T, tw, bit + 32) << 32);
}

// Is the value within 2^-64 of the rounding boundary below bit position `pos` of T (rounding bit at
// pos - 1)? -- all zero above a set rounding bit or all one below a clear one in the 64 bits that
// follow (an entry is good to 2 units of 2^-P, far below).
QHD bool exact_near_boundary(const uint32_t* T, uint32_t tw, uint32_t pos) {
  const uint32_t round_bit = exact_table_bits(T, tw, pos - 1) & 1u;
  const uint64_t guard = exact_table_bits64(T, tw, pos - 65);
  return round_bit ? guard == 0ull : guard == ~0ull;
}

// Below this e the reference's own rounding of 2^(e + i/D) to 3 (e + 1) bits (src/sample.cpp:97-115:
// precision = 3 ceil(|max_log_alpha|), mpfr_exp2 to nearest) is wider than 2^-64 and is reproduced
// step by step instead of being argued away.
#define QB_EXACT_SMALL_E 31
// ... and below this one mpfr_set_d rounds the argument e + i / D itself (3 (e + 1) bits < 4 + 14 for
// the largest dimension): not supported. |alpha| < 256 there.
#define QB_EXACT_MIN_E 8

// out (wa limbs, strided) = round(2^(e + idx / D_max)) as the reference computes it for a region at
// e_region < QB_EXACT_SMALL_E, idx on [0, D_max] (idx = D_max: the bound 2^(e + 1), at the precision of
// the region below it). Returns QB_EXACT_OK or QB_EXACT_AMBIGUOUS. Requires e + 1 <= emax. (Regions at
// e_region >= 31 take the one-pass form inside exact_region_modulus.)
template <int S>
QHD int exact_bound(const ExactConst& c, uint32_t e_region, uint32_t e, uint32_t idx, uint32_t* out) {
  if (idx == c.table_dim) {
    e += 1;
    idx = 0;
  }
  const uint32_t* T = c.table + (size_t)idx * c.tw;
  const uint32_t sh = c.P - e;  // >= QB_EXACT_GUARD: the integer part of T 2^(e - P) starts at bit sh
  int status = QB_EXACT_OK;
  // both roundings of the reference. mpfr_exp2 at p = 3 (e_region + 1) bits: the value has e + 1
  // integer bits, so F = p - (e + 1) fractional bits survive, to nearest ...
  const uint32_t F = 3 * (e_region + 1) - (e + 1);  // on [2 e_region + 1, 61]
  const uint32_t rp = sh - F;                        // > 64: P - e >= 128 + emax - e
  uint64_t frac = exact_table_bits64(T, c.tw, rp);   // low F bits: the fraction; above: integer bits
  uint64_t whole = exact_table_bits64(T, c.tw, sh);  // the integer part (e + 1 <= 32 bits)
  frac &= (1ull << F) - 1ull;
  if (idx != 0) {  // 2^e itself is exact
    if (exact_near_boundary(T, c.tw, rp)) status = QB_EXACT_AMBIGUOUS;
    if (exact_table_bits(T, c.tw, rp - 1) & 1u) {
      frac += 1;
      if (frac >> F) {
        frac = 0;
        whole += 1;
      }
    }
  }
  // ... then mpfr_round: to the nearest integer, halves away from zero
  if ((frac >> (F - 1)) != 0) whole += 1;
  for (uint32_t i = 0; i < c.wa; i++) QB_L(out, i) = i == 0 ? (uint32_t)whole : (i == 1 ? (uint32_t)(whole >> 32) : 0u);
  return status;
}

QHD uint32_t qb_clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)v);
#else
  return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}

// The bytes random_generate_mpz reads for a modulus of `bits` bits (src/random.c:163-164).
QHD uint32_t exact_bytes_for_bits(uint32_t bits) { return (bits + 64u + 8u) / 8u; }

// ---- v mod M ---------------------------------------------------------------------------

// The division steps of Knuth's algorithm D (TAOCP 4.3.1) in 32-bit digits. Vn: the dividend shifted
// left by s, nv + 1 limbs (strided); M: the divisor, wm >= 2 limbs, top limb non-zero, s = clz(top
// limb) -- M is read shifted on the fly (one load per limb, the previous one carried) and never
// written. On return Vn[0, wm) holds the remainder shifted left by s and Vn[wm, nv] is zero.
template <int S>
QHD void exact_div_core(uint32_t* Vn, uint32_t nv, const uint32_t* M, uint32_t wm, uint32_t s) {
  if (nv < wm) return;
  const uint32_t m_top = QB_L(M, wm - 1), m_next = QB_L(M, wm - 2), m_third = wm >= 3 ? QB_L(M, wm - 3) : 0u;
  const uint64_t vtop = s ? ((m_top << s) | (m_next >> (32u - s))) : m_top;
  const uint64_t vnext = s ? ((m_next << s) | (m_third >> (32u - s))) : m_next;
  for (uint32_t j = nv - wm + 1; j-- > 0;) {
    const uint64_t num = ((uint64_t)QB_L(Vn, j + wm) << 32) | QB_L(Vn, j + wm - 1);
    uint64_t qhat = num / vtop, rhat = num % vtop;
    const uint64_t u2 = QB_L(Vn, j + wm - 2);
    while ((qhat >> 32) != 0 || qhat * vnext > ((rhat << 32) | u2)) {
      qhat--;
      rhat += vtop;
      if ((rhat >> 32) != 0) break;
    }
    if (qhat == 0) continue;
    // Vn[j, j + wm] -= qhat (M << s)
    uint64_t carry = 0;
    uint32_t borrow = 0, prev = 0;
    for (uint32_t i = 0; i < wm; i++) {
      const uint32_t cur = QB_L(M, i);
      const uint32_t mn = s ? ((cur << s) | (prev >> (32u - s))) : cur;
      prev = cur;
      const uint64_t p = qhat * (uint64_t)mn + carry;
      carry = p >> 32;
      const uint64_t t = (uint64_t)QB_L(Vn, i + j) - (uint32_t)p - borrow;
      QB_L(Vn, i + j) = (uint32_t)t;
      borrow = (uint32_t)(t >> 63);
    }
    const uint64_t t = (uint64_t)QB_L(Vn, j + wm) - carry - borrow;
    QB_L(Vn, j + wm) = (uint32_t)t;
    if (t >> 63) {  // qhat was one too large: add M << s back
      uint64_t c2 = 0;
      prev = 0;
      for (uint32_t i = 0; i < wm; i++) {
        const uint32_t cur = QB_L(M, i);
        const uint32_t mn = s ? ((cur << s) | (prev >> (32u - s))) : cur;
        prev = cur;
        c2 += (uint64_t)QB_L(Vn, i + j) + mn;
        QB_L(Vn, i + j) = (uint32_t)c2;
        c2 >>= 32;
      }
      QB_L(Vn, j + wm) += (uint32_t)c2;
    }
  }
}

// V (nv limbs, strided, one more limb of room above) modulo M (wm limbs, top limb non-zero): the
// remainder is left in V[0, wm). wm = 1: short division. (The general form, for the tests and for
// callers that hold V unshifted; exact_alpha shifts while it imports the bytes.)
template <int S>
QHD void exact_mod(uint32_t* V, uint32_t nv, const uint32_t* M, uint32_t wm) {
  if (nv < wm) return;
  if (wm == 1) {  // a one-limb modulus (regions of slices at |log alpha| < ~40)
    const uint64_t mod = QB_L(M, 0);
    uint64_t rem = 0;
    for (uint32_t i = nv; i-- > 0;) rem = ((rem << 32) | QB_L(V, i)) % mod;
    QB_L(V, 0) = (uint32_t)rem;
    return;
  }
  const uint32_t s = qb_clz32(QB_L(M, wm - 1));
  if (s) {
    QB_L(V, nv) = QB_L(V, nv - 1) >> (32u - s);
    for (uint32_t i = nv; i-- > 1;) QB_L(V, i) = (QB_L(V, i) << s) | (QB_L(V, i - 1) >> (32u - s));
    QB_L(V, 0) <<= s;
  } else {
    QB_L(V, nv) = 0;
  }
  exact_div_core<S>(V, nv, M, wm, s);
  if (s) {
    for (uint32_t i = 0; i + 1 < wm; i++) QB_L(V, i) = (QB_L(V, i) >> s) | (QB_L(V, i + 1) << (32u - s));
    QB_L(V, wm - 1) >>= s;
  }
}

// ---- alpha from a region (sample_alpha_from_region, src/sample.cpp:78-158) -----------------

// bits(max - min) of a region, or 0 with *status set; M (wa + 1 limbs at stride S) = max - min,
// lo (wa limbs at stride SL) = min. One pass: both table rows are read a word per limb (the previous
// word carried), rounded, subtracted and stored.
template <int S, int SL>
QHD uint32_t exact_region_modulus(const ExactConst& c, const ExactRegion& g, uint32_t* lo, uint32_t* M, int* status) {
  const uint32_t e = (uint32_t)(g.min_log_alpha < 0 ? -(int64_t)g.min_log_alpha : (int64_t)g.min_log_alpha);
  const uint32_t D = g.dimension;
  // e < 8 (|alpha| < 256): the reference's precision of 3 (e + 1) bits no longer holds e + i / D itself
  if (e < QB_EXACT_MIN_E || e + 1 > c.emax || D == 0 || (D & (D - 1)) != 0 || D > c.table_dim || g.region >= D) {
    *status = QB_EXACT_UNSUPPORTED;
    return 0;
  }
  const uint32_t step = c.table_dim / D;
  uint32_t top_i = 0, top_v = 0;
  if (e < QB_EXACT_SMALL_E) {
    // toy sizes: the bounds one after the other, with both of the reference's roundings (exact_bound)
    const int s_lo = exact_bound<SL>(c, e, e, g.region * step, lo);
    const int s_hi = exact_bound<S>(c, e, e, (g.region + 1) * step, M);
    if (s_lo != QB_EXACT_OK || s_hi != QB_EXACT_OK) {
      *status = QB_EXACT_AMBIGUOUS;
      return 0;
    }
    uint32_t borrow = 0;
    for (uint32_t i = 0; i < c.wa; i++) {  // src/sample.cpp:126-128
      const uint64_t t = (uint64_t)QB_L(M, i) - lo[(size_t)i * SL] - borrow;
      QB_L(M, i) = (uint32_t)t;
      borrow = (uint32_t)(t >> 63);
      if ((uint32_t)t) {
        top_i = i;
        top_v = (uint32_t)t;
      }
    }
  } else {
    uint32_t idx_lo = g.region * step, idx_hi = (g.region + 1) * step, e_hi = e;
    if (idx_hi == c.table_dim) {  // the bound 2^(e + 1)
      e_hi = e + 1;
      idx_hi = 0;
    }
    const uint32_t* Tl = c.table + (size_t)idx_lo * c.tw;
    const uint32_t* Th = c.table + (size_t)idx_hi * c.tw;
    const uint32_t sl = c.P - e, sh = c.P - e_hi;  // >= QB_EXACT_GUARD
    if ((idx_lo != 0 && exact_near_boundary(Tl, c.tw, sl)) || (idx_hi != 0 && exact_near_boundary(Th, c.tw, sh))) {
      *status = QB_EXACT_AMBIGUOUS;
      return 0;
    }
    uint32_t cl = exact_table_bits(Tl, c.tw, sl - 1) & 1u, ch = exact_table_bits(Th, c.tw, sh - 1) & 1u;
    const uint32_t wl = sl >> 5, ol = sl & 31u, wh = sh >> 5, oh = sh & 31u;
    uint32_t pl = wl < c.tw ? Tl[wl] : 0u, ph = wh < c.tw ? Th[wh] : 0u;
    uint32_t borrow = 0;
    for (uint32_t i = 0; i < c.wa; i++) {
      const uint32_t nl = wl + i + 1 < c.tw ? Tl[wl + i + 1] : 0u, nh = wh + i + 1 < c.tw ? Th[wh + i + 1] : 0u;
      const uint32_t vl = ol ? ((pl >> ol) | (nl << (32u - ol))) : pl;
      const uint32_t vh = oh ? ((ph >> oh) | (nh << (32u - oh))) : ph;
      pl = nl;
      ph = nh;
      const uint32_t rl = vl + cl, rh = vh + ch;  // the rounding bits rippling up
      cl = rl < vl ? 1u : 0u;
      ch = rh < vh ? 1u : 0u;
      lo[(size_t)i * SL] = rl;
      const uint64_t t = (uint64_t)rh - rl - borrow;  // src/sample.cpp:126-128
      QB_L(M, i) = (uint32_t)t;
      borrow = (uint32_t)(t >> 63);
      if ((uint32_t)t) {
        top_i = i;
        top_v = (uint32_t)t;
      }
    }
  }
  QB_L(M, c.wa) = 0;
  const uint32_t bits = top_v ? 32u * top_i + (32u - qb_clz32(top_v)) : 0u;
  // max = min (tiny regions of a slice at |log alpha| < ~8): the reference divides by zero
  *status = bits ? QB_EXACT_OK : QB_EXACT_UNSUPPORTED;
  return bits;
}

// The second half of exact_alpha: V (nv + 1 limbs, the sample's integer shifted left by s) modulo M
// (wm limbs), and alpha_out (holding min) += remainder >> s with the low kappa bits cleared.
template <int S, int SO>
QHD void exact_alpha_finish(const ExactConst& c, uint32_t* V, uint32_t nv, const uint32_t* M, uint32_t wm, uint32_t s,
                            uint32_t kappa, uint32_t* alpha_out) {
  if (wm >= 2) {
    exact_div_core<S>(V, nv, M, wm, s);  // src/random.c:179
  } else {
    const uint64_t mod = QB_L(M, 0);
    uint64_t rem = 0;
    for (uint32_t i = nv; i-- > 0;) rem = ((rem << 32) | QB_L(V, i)) % mod;
    QB_L(V, 0) = (uint32_t)rem;
    QB_L(V, 1) = 0;
  }
  // alpha = min + (remainder >> s) (src/sample.cpp:131), then the low kappa bits cleared (:133-144)
  uint32_t carry = 0;
  uint32_t cur = QB_L(V, 0);
  for (uint32_t i = 0; i < c.wa; i++) {
    uint32_t rem = 0;
    if (i < wm && i < nv) {
      const uint32_t nxt = QB_L(V, i + 1);  // limb wm of the shifted remainder is zero
      rem = s ? ((cur >> s) | (nxt << (32u - s))) : cur;
      cur = nxt;
    }
    const uint64_t t = (uint64_t)alpha_out[(size_t)i * SO] + rem + carry;
    uint32_t v = (uint32_t)t;
    carry = (uint32_t)(t >> 32);
    if (32u * i < kappa) v &= (kappa - 32u * i >= 32u) ? 0u : ~((1u << (kappa - 32u * i)) - 1u);
    alpha_out[(size_t)i * SO] = v;
  }
}

// The 32-bit limb whose most significant byte is p[0] (big-endian).
QHD uint32_t exact_load_be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// |alpha| (wa limbs at stride SO) of one sample; *negative = the sign. scratch:
// exact_alpha_scratch_limbs words (strided).
template <int S, int SO>
QHD int exact_alpha(const ExactConst& c, const ExactRegion& g, uint32_t kappa, const uint8_t* stream,
                    uint64_t stream_len, uint32_t* scratch, uint32_t* alpha_out, int* negative) {
  uint32_t* V = scratch;                             // wa + 5 limbs
  uint32_t* M = scratch + (size_t)(c.wa + 5) * S;    // wa + 1 limbs
  *negative = g.min_log_alpha < 0 ? 1 : 0;
  int status = QB_EXACT_OK;
  const uint32_t bits = exact_region_modulus<S, SO>(c, g, alpha_out, M, &status);  // alpha_out = min
  if (status == QB_EXACT_OK && (g.length != exact_bytes_for_bits(bits) || g.offset + g.length > stream_len))
    status = QB_EXACT_LENGTH;
  if (status != QB_EXACT_OK) {
    for (uint32_t i = 0; i < c.wa; i++) alpha_out[(size_t)i * SO] = 0u;
    return status;
  }
  const uint32_t wm = (bits + 31) / 32;
  const uint32_t s = wm >= 2 ? qb_clz32(QB_L(M, wm - 1)) : 0u;  // the one-limb division is not normalised
  // v: big-endian bytes (mpz_import(value, length, 1, 1, 1, 0, buffer), src/random.c:178), stored shifted
  // left by s as the division wants it. Limb i = the four bytes ending 4 i bytes before the end.
  const uint32_t nv = (g.length + 3) / 4;  // <= wa + 3
  const uint8_t* b = stream + g.offset;
  const uint32_t full = g.length / 4;      // limbs with all four bytes
  uint32_t prev = 0;
#if defined(__CUDA_ARCH__)
  // aligned 32-bit loads: limb i lies at the byte address q - 4 i, whose alignment is the same for
  // every i -- one aligned word per limb, the other half carried (funnel shift), then a byte swap
  const size_t q = (size_t)(b + g.length - 4);
  const uint32_t mis = (uint32_t)(q & 3u) * 8u;
  const uint32_t* aw = (const uint32_t*)(q & ~(size_t)3);
  uint32_t upper = (mis && full) ? aw[1] : 0u;  // the word above limb 0's aligned word (inside the buffer: mis != 0)
#endif
  for (uint32_t i = 0; i < nv; i++) {
    uint32_t w = 0;
    if (i < full) {
#if defined(__CUDA_ARCH__)
      const uint32_t lower = *(aw - i);
      w = __byte_perm(mis ? __funnelshift_r(lower, upper, mis) : lower, 0u, 0x0123);
      upper = lower;
#else
      w = exact_load_be32(b + g.length - 4 - 4 * (size_t)i);
#endif
    } else {
      for (uint32_t k = 0; k < g.length - 4 * full; k++) w |= (uint32_t)b[g.length - 4 * full - 1 - k] << (8 * k);
    }
    QB_L(V, i) = s ? ((w << s) | (prev >> (32u - s))) : w;
    prev = w;
  }
  QB_L(V, nv) = s ? (prev >> (32u - s)) : 0u;
  exact_alpha_finish<S, SO>(c, V, nv, M, wm, s, kappa, alpha_out);
  return QB_EXACT_OK;
}

// ---- small limb helpers ------------------------------------------------------------------

// p (n limbs, strided) = -p modulo 2^(32 n).
template <int S>
QHD void limbs_negate(uint32_t* p, uint32_t n) {
  uint32_t carry = 1;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t v = ~QB_L(p, i) + carry;
    carry = (carry && v == 0u) ? 1u : 0u;
    QB_L(p, i) = v;
  }
}

// 32 bits of a strided number of n limbs from bit position `bit` upwards (zero above).
template <int S>
QHD uint32_t limbs_bits(const uint32_t* p, uint32_t n, uint32_t bit) {
  const uint32_t w = bit >> 5, off = bit & 31u;
  const uint32_t lo = w < n ? QB_L(p, w) : 0u;
  if (!off) return lo;
  const uint32_t hi = w + 1 < n ? QB_L(p, w + 1) : 0u;
  return (lo >> off) | (hi << (32u - off));
}

// out (wn limbs at stride SO) = (bits [shift, shift + n) of P (np limbs)) + t 2^(n - kappa) mod 2^n,
// t: tw limbs at stride ST (ignored when kappa = 0); only the low kappa bits of t count.
template <int S, int SO, int ST>
QHD void exact_finish_j(const ExactConst& c, const uint32_t* P, uint32_t np, uint32_t shift, uint32_t kappa,
                        const uint32_t* t, uint32_t* out) {
  const uint32_t n = c.n;
  const uint32_t tpos = n - kappa;  // t's bit 0 lands here
  uint32_t carry = 0;
  for (uint32_t i = 0; i < c.wn; i++) {
    uint32_t v = limbs_bits<S>(P, np, shift + 32u * i);
    uint32_t add = 0;
    if (kappa && 32u * i + 32u > tpos) {
      // bits of t at positions [32 i - tpos, 32 i - tpos + 32)
      const int64_t from = (int64_t)32 * i - (int64_t)tpos;
      const uint32_t tl = (kappa + 31) / 32;
      if (from >= 0) {
        add = limbs_bits<ST>(t, tl, (uint32_t)from);
      } else {
        add = limbs_bits<ST>(t, tl, 0) << (uint32_t)(-from);  // -from < 32
      }
      // t counts modulo 2^kappa: bits at or above n fall away below
    }
    const uint64_t s = (uint64_t)v + add + carry;
    v = (uint32_t)s;
    carry = (uint32_t)(s >> 32);
    if (32u * i + 32u > n) v &= (1u << (n - 32u * i)) - 1u;  // n - 32 i on [1, 31]
    out[(size_t)i * SO] = v;
  }
}

// ---- j from alpha_r (sample_j_from_alpha_r, src/sample.cpp:160-208; the diagonal twin :354-410;
//      the first half of sample_j_k_from_alpha_d_r, :290-335, whose t_r is a multiple of
//      2^kappa_t_r: the caller passes t already multiplied) ------------------------------------
// a: |alpha_r|, wa limbs at stride SA. scratch: exact_jk_scratch_limbs words (strided by S).
template <int NC, int S, int SA, int SO, int ST>
QHD void exact_j_from_alpha_r(const ExactConst& c, const uint32_t* a, int negative, const uint32_t* t,
                              uint32_t* scratch, uint32_t* j_out) {
  uint32_t* P = scratch;  // wn + 4 limbs
  const uint32_t cols = (c.n + c.kappa_r + 31) / 32;  // <= wn + ceil(kappa_r / 32)
  Acc96 acc;
  acc_zero(acc);
  mul_columns_wide<NC, SA, 1, S>(a, c.wa, c.inv_r, c.wn, 0, cols - 1, 0, P, acc);  // :386-387 (low columns)
  if (negative) limbs_negate<S>(P, cols);
  exact_finish_j<S, SO, ST>(c, P, cols, c.kappa_r, c.kappa_r, t, j_out);    // :389-406
}

// ---- k from (alpha_d, j) (the second half of sample_j_k_from_alpha_d_r, src/sample.cpp:337-347) --
// j: wn limbs at stride SJ; a: |alpha_d|, wa limbs at stride SA; k_out: wk limbs at stride SO.
template <int NC, int S, int SA, int SJ, int SO>
QHD void exact_k_from_alpha_d_j(const ExactConst& c, const uint32_t* a, int negative, const uint32_t* j,
                                uint32_t* scratch, uint32_t* k_out) {
  uint32_t* P = scratch;  // wn + 4 limbs
  const uint32_t cols = (c.m + c.kbits + 31) / 32;  // = wn for the two-dimensional sampler
  Acc96 acc;
  acc_zero(acc);
  mul_columns_wide<NC, SJ, 1, S>(j, c.wn, c.d, c.wd, 0, cols - 1, 0, P, acc);  // :338 (low columns)
  // P = alpha_d - d j modulo 2^(32 cols) (:339)
  uint32_t borrow = 0, ncarry = 1;
  for (uint32_t i = 0; i < cols; i++) {
    uint32_t x = i < c.wa ? a[(size_t)i * SA] : 0u;
    if (negative) {
      x = ~x + ncarry;
      ncarry = (ncarry && x == 0u) ? 1u : 0u;
    }
    const uint64_t t = (uint64_t)x - QB_L(P, i) - borrow;
    QB_L(P, i) = (uint32_t)t;
    borrow = (uint32_t)(t >> 63);
  }
  for (uint32_t i = 0; i < c.wk; i++) {  // :341-347
    uint32_t v = limbs_bits<S>(P, cols, c.m + 32u * i);
    if (32u * i + 32u > c.kbits) v &= (1u << (c.kbits - 32u * i)) - 1u;
    k_out[(size_t)i * SO] = v;
  }
}

// ---- j from (alpha_d, k) (sample_j_k_from_alpha_d, src/sample.cpp:210-273; k and t_d are drawn
//      by the caller, :230-239) -------------------------------------------------------------
// kk: k, wk limbs at stride SK.
template <int NC, int S, int SA, int SK, int SO, int ST>
QHD void exact_j_from_alpha_d_k(const ExactConst& c, const uint32_t* a, int negative, const uint32_t* kk,
                                const uint32_t* t, uint32_t* scratch, uint32_t* j_out) {
  uint32_t* X = scratch;                             // wn + 4 limbs
  uint32_t* P = scratch + (size_t)(c.wn + 4) * S;    // wn + 4 limbs
  const uint32_t cols = (c.n + c.kappa_d + 31) / 32;
  // X = alpha_d - 2^m k modulo 2^(32 cols) (:250-253)
  uint32_t borrow = 0, ncarry = 1;
  for (uint32_t i = 0; i < cols; i++) {
    uint32_t x = i < c.wa ? a[(size_t)i * SA] : 0u;
    if (negative) {
      x = ~x + ncarry;
      ncarry = (ncarry && x == 0u) ? 1u : 0u;
    }
    // limb i of 2^m k
    uint32_t y = 0;
    if (32u * i + 32u > c.m) {
      const int64_t from = (int64_t)32 * i - (int64_t)c.m;
      y = from >= 0 ? limbs_bits<SK>(kk, c.wk, (uint32_t)from) : (limbs_bits<SK>(kk, c.wk, 0) << (uint32_t)(-from));
    }
    const uint64_t tt = (uint64_t)x - y - borrow;
    QB_L(X, i) = (uint32_t)tt;
    borrow = (uint32_t)(tt >> 63);
  }
  // Y = floor(X / 2^kappa_d) modulo 2^n (:255-257), in place
  for (uint32_t i = 0; i < c.wn; i++) {
    uint32_t v = limbs_bits<S>(X, cols, c.kappa_d + 32u * i);
    if (32u * i + 32u > c.n) v &= (1u << (c.n - 32u * i)) - 1u;
    QB_L(X, i) = v;
  }
  Acc96 acc;
  acc_zero(acc);
  mul_columns_wide<NC, S, 1, S>(X, c.wn, c.inv_d, c.wn, 0, c.wn - 1, 0, P, acc);  // :259 (low columns)
  exact_finish_j<S, SO, ST>(c, P, c.wn, 0, c.kappa_d, t, j_out);           // :261-268
}

}  // namespace qb200
