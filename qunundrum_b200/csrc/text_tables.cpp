// text_tables.cpp -- see text_tables.hpp.
#include "text_tables.hpp"

#include "bigint.hpp"

namespace qb200 {
namespace text {

namespace {

void store(const BigUInt& T, long e2, bool exact, Pow10Entry* e) {
  for (int i = 0; i < 3; i++) {
    const uint64_t limb = (size_t)i < T.w.size() ? T.w[i] : 0;
    e->w[2 * i] = (uint32_t)limb;
    e->w[2 * i + 1] = (uint32_t)(limb >> 32);
  }
  e->e2 = (int32_t)e2;
  e->exact = exact ? 1u : 0u;
}

}  // namespace

void build_pow10_table(std::vector<Pow10Entry>& out) {
  out.assign((size_t)(K_MAX - K_MIN + 1), Pow10Entry());
  const int top = K_MAX > -K_MIN ? K_MAX : -K_MIN;
  BigUInt N(1);  // 5^n
  for (int n = 0; n <= top; n++) {
    if (n) N = BigUInt::mul_small(N, 5);
    const size_t b = N.bit_length();
    if (n <= K_MAX) {
      // 10^n = N * 2^n; T = the top 192 bits of N (truncated)
      BigUInt T;
      bool exact = true;
      if (b <= 192) {
        T = N.shl(192 - b);
      } else {
        T = N.shr(b - 192);
        exact = !N.any_below(b - 192);
      }
      store(T, (long)b - 1 + n, exact, &out[(size_t)(n - K_MIN)]);
    }
    if (n > 0 && -n >= K_MIN) {
      // 10^-n = 2^-n / N; T = floor(2^(191 + b) / N) in [2^191, 2^192)
      BigUInt q, r;
      BigUInt::divmod(BigUInt::pow2(191 + b), N, q, r);
      store(q, -(long)b - n, r.is_zero(), &out[(size_t)(-n - K_MIN)]);
    }
  }
}

}  // namespace text
}  // namespace qb200
