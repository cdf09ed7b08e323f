// text_tables.hpp -- host-side construction of the powers-of-ten table the text
// kernels use (textfmt.cuh). Built once per context with the library's own big
// integers; checked entry by entry against exact Python integers in
// tests/test_text_format.py.
#pragma once

#include <vector>

#include "textfmt.cuh"

namespace qb200 {
namespace text {

// out[k - K_MIN] for k = K_MIN .. K_MAX.
void build_pow10_table(std::vector<Pow10Entry>& out);

}  // namespace text
}  // namespace qb200
