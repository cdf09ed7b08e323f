// tests/hostsim/abi_shim.cpp -- TEST-ONLY stand-in for the TEXT entry points of
// libqunundrum_b200.so, backed by the CPU compile of textfmt.cuh / textparse.cuh
// (tests/hostsim/hostsim.cpp).
//
// Purpose: the host logic of the reference-side translation unit
// qunundrum_b200/dropin/dropin_text.cpp -- block reads, re-reads when a block is too
// short or ends inside a number, seeking to where fscanf would have stopped, the running
// long double sums, fwrite -- can then run inside the reference's own executables
// (filter_distribution, compare_*_distributions) in the GPU-less test suite.
// It is NOT part of the product: nothing under qunundrum_b200/ references it, the
// product library has no CPU path, and the "shim" flavour of the integration build is
// used by tests/test_text_dropin_host_logic.py only.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"

extern "C" {
size_t hostsim_text_format_ld(const long double* v, size_t n, char* out, int force_band,
                              uint64_t* n_exact);
int hostsim_text_parse_ld(const char* textp, size_t len, size_t n, long double* values,
                          size_t* consumed, int force_band, uint64_t* n_exact);
}

struct qb200_context {
  std::vector<char> text;
};

namespace {
thread_local std::string g_err;
}

extern "C" {

int qb200_device_count(void) { return 1; }
const char* qb200_last_error(void) { return g_err.c_str(); }

int qb200_create(int, qb200_context** ctx) {
  *ctx = new qb200_context;
  return 0;
}
void qb200_destroy(qb200_context* ctx) { delete ctx; }

int qb200_text_format_ld(qb200_context* ctx, const long double* values, size_t n,
                         const long double* tail, const char** text, size_t* len) {
  ctx->text.resize(34 * (n + 1) + 64);
  size_t pos = hostsim_text_format_ld(values, n, ctx->text.data(), 0, nullptr);
  if (tail) pos += hostsim_text_format_ld(tail, 1, ctx->text.data() + pos, 0, nullptr);
  *text = ctx->text.data();
  *len = pos;
  return 0;
}

int qb200_text_parse_ld(qb200_context*, const char* text, size_t len, size_t n, long double* values,
                        size_t* consumed) {
  const int rc = hostsim_text_parse_ld(text, len, n, values, consumed, 0, nullptr);
  if (rc) g_err = "hostsim parse error " + std::to_string(rc);
  return rc;
}

}  // extern "C"
