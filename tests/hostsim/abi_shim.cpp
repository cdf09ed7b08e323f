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
#include <cmath>
#include <cstdint>
#include <cstdlib>
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

// ---- sampler entry points (TEST-ONLY, same purpose): the host logic of
// qunundrum_b200/dropin/dropin_tau.cpp -- batching, the queue of pre-drawn words, the fast
// Keccak stream, re-use across calls -- and the serial part of qb200_sampler_tau_estimate
// (csrc/sampler_host.hpp: stream layout, failure threshold, log2 in long double) run in the
// GPU-less suite on the reference's own Distribution / Random_State structs. The samples
// themselves come from the CPU compile of sampler.cuh (hostsim_sampler_*).
#include "../../qunundrum_b200/csrc/sampler_host.hpp"

extern "C" {
void* hostsim_sampler_new(int dims, uint32_t m, uint32_t n_slices, const uint32_t* dimension,
                          const int32_t* c0, const int32_t* c1, const long double* cells,
                          const long double* slice_total, const long double* total);
void hostsim_sampler_free(void* h);
int hostsim_sampler_bad(void* h);
void hostsim_sampler_sample(void* hh, uint32_t k, const uint64_t* words, int force_exact, double* out,
                            int32_t* status, int32_t* exact);
}

struct qb200_sampler {
  void* h = nullptr;
  int dims = 2;
  int m = 0;
  qb200::FailureThreshold fail;
  qb200::TauLayout layout;
};

extern "C" {

int qb200_sampler_create(qb200_context*, int dims, uint32_t m, uint32_t n_slices, const uint32_t* dimension,
                         const int32_t* c0, const int32_t* c1, const long double* const* cells,
                         const long double* slice_total, long double total_probability,
                         qb200_sampler** out) {
  std::vector<long double> flat;
  for (uint32_t i = 0; i < n_slices; i++) {
    const size_t nc = dims == 2 ? (size_t)dimension[i] * dimension[i] : dimension[i];
    flat.insert(flat.end(), cells[i], cells[i] + nc);
  }
  qb200_sampler* s = new qb200_sampler;
  s->dims = dims;
  s->m = (int)m;
  s->h = hostsim_sampler_new(dims, m, n_slices, dimension, c0, c1, flat.data(), slice_total,
                             &total_probability);
  if (hostsim_sampler_bad(s->h)) {
    hostsim_sampler_free(s->h);
    delete s;
    g_err = "unsupported probability value";
    return -14;
  }
  s->fail = qb200::find_failure_threshold(slice_total, n_slices, total_probability);
  *out = s;
  return 0;
}

void qb200_sampler_destroy(qb200_sampler* s) {
  if (!s) return;
  hostsim_sampler_free(s->h);
  delete s;
}

uint32_t qb200_sampler_words_per_sample(const qb200_sampler* s) { return (uint32_t)s->dims + 2u; }

int qb200_sampler_tau_estimate(qb200_sampler* s, uint32_t n, uint32_t count, const uint64_t* words,
                               size_t n_words, size_t* words_used, uint32_t* done, long double* tau0,
                               long double* tau1, uint8_t* ok) {
  if (n == 0) {
    g_err = "n must be positive";
    return -15;
  }
  const uint32_t wps = (uint32_t)s->dims + 2u;
  qb200::tau_layout(s->fail, wps, n, count, words, n_words, &s->layout);
  if (done) *done = s->layout.done;
  if (words_used) *words_used = s->layout.words_used;
  std::vector<double> sums(4 * (size_t)s->layout.done, 0.0);
  std::vector<int> status(s->layout.done, 0);
  std::vector<double> out(8 * (size_t)n);
  std::vector<int32_t> st(n), ex(n);
  for (uint32_t t = 0; t < s->layout.done; t++) {
    if (s->layout.off[t] == QB_TAU_SKIP) continue;
    hostsim_sampler_sample(s->h, n, words + s->layout.off[t], 0, out.data(), st.data(), ex.data());
    long double a = 0, b = 0;   // (the device sums in double-double; long double is as good here)
    for (uint32_t i = 0; i < n; i++) {
      if (st[i] != 0) {
        status[t] = st[i];
        break;
      }
      a += (long double)out[8 * i] + (long double)out[8 * i + 1];
      b += (long double)out[8 * i + 2] + (long double)out[8 * i + 3];
    }
    sums[4 * (size_t)t] = (double)a;
    sums[4 * (size_t)t + 1] = (double)(a - (long double)(double)a);
    sums[4 * (size_t)t + 2] = (double)b;
    sums[4 * (size_t)t + 3] = (double)(b - (long double)(double)b);
  }
  std::string err;
  const int rc = qb200::tau_finish(s->dims, s->m, n, s->layout, sums.data(), status.data(), tau0, tau1, ok, &err);
  if (rc) g_err = err;
  return rc;
}

}  // extern "C"

// ---- diagonal k sampler entry points (TEST-ONLY, same purpose): the host logic of
// qunundrum_b200/dropin/dropin_tau_diagonal.cpp -- the stream layout of the draws, the replay after a
// failing sample, the MPFR sum -- runs in the GPU-less suite
// inside integration/tools/tau_diagonal_check.cpp, next to the reference's own
// tau_estimate_diagonal. k and alpha_phi come from the CPU compile of diagk.cuh.
extern "C" {
void* hostsim_diagk_new(uint32_t m, uint32_t sigma, uint32_t l, const uint8_t* d, size_t dn,
                        const uint8_t* r, size_t rn);
void hostsim_diagk_free(void* h);
void hostsim_diagk_dims(void* hh, uint32_t* out3);
int hostsim_diagk_sample(void* hh, uint32_t n, const uint32_t* j, const int32_t* eta,
                         const long double* pivot, uint64_t delta_bound, uint32_t* k_out, double* x,
                         int64_t* delta, int32_t* status);
const char* hostsim_last_error();
void hostsim_diagk_h(uint32_t l, uint32_t n, const double* x, long double* out);
}

struct qb200_diagk {
  void* h = nullptr;
  uint32_t dims[3] = {0, 0, 0};
  uint32_t l = 0;
};

extern "C" {

int qb200_diagk_create(qb200_context*, const qb200_params* p, qb200_diagk** out) {
  *out = nullptr;
  void* h = hostsim_diagk_new(p->m, p->sigma, p->l, p->d_be, p->d_len, p->r_be, p->r_len);
  if (!h) {
    g_err = hostsim_last_error();
    return -2;
  }
  qb200_diagk* s = new qb200_diagk;
  s->h = h;
  s->l = p->l;
  hostsim_diagk_dims(h, s->dims);
  *out = s;
  return 0;
}
void qb200_diagk_destroy(qb200_diagk* s) {
  if (!s) return;
  hostsim_diagk_free(s->h);
  delete s;
}
uint32_t qb200_diagk_j_limbs(const qb200_diagk* s) { return s->dims[1]; }
uint32_t qb200_diagk_k_limbs(const qb200_diagk* s) { return s->dims[2]; }

int qb200_diagk_sample(qb200_diagk* s, uint32_t n, const uint32_t* j, const int32_t* eta,
                       const long double* pivot, uint32_t delta_bound, uint32_t* k, double* x_hi,
                       double* x_lo, int64_t* delta, int32_t* status) {
  std::vector<uint32_t> kk((size_t)n * s->dims[2]);
  std::vector<double> x(2 * (size_t)n);
  std::vector<int64_t> dl(n);
  std::vector<int32_t> st(n);
  if (hostsim_diagk_sample(s->h, n, j, eta, pivot, delta_bound, kk.data(), x.data(), dl.data(), st.data())) {
    g_err = "The pivot is out of bounds.";
    return -42;
  }
  for (uint32_t i = 0; i < n; i++) {
    if (x_hi) x_hi[i] = x[2 * i];
    if (x_lo) x_lo[i] = x[2 * i + 1];
    if (delta) delta[i] = dl[i];
    if (status) status[i] = st[i];
  }
  if (k) memcpy(k, kk.data(), kk.size() * 4);
  return 0;
}

int qb200_diagk_h(qb200_diagk* s, uint32_t n, const double* x_hi, const double* x_lo, long double* h) {
  std::vector<double> x(2 * (size_t)n);
  for (uint32_t i = 0; i < n; i++) {
    x[2 * i] = x_hi[i];
    x[2 * i + 1] = x_lo[i];
  }
  hostsim_diagk_h(s->l, n, x.data(), h);
  return 0;
}

}  // extern "C"

// ---- exact sampler entry points (TEST-ONLY, same purpose): the stream layout of
// dropin_tau_diagonal.cpp (regions, byte counts, t_r) runs in the GPU-less suite; alpha_r, j and k come
// from the CPU compile of exact.cuh and diagk.cuh.
extern "C" {
void* hostsim_exact_new(int kind, uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d, size_t dn,
                        const uint8_t* r, size_t rn, uint32_t dimension_max, uint32_t emax);
void hostsim_exact_free(void* h);
void hostsim_exact_dims(void* hh, uint32_t* out9);
uint32_t hostsim_exact_region_bytes(void* hh, int32_t min_log_alpha, uint32_t region, uint32_t dimension,
                                    int32_t* status);
void hostsim_exact_alpha(void* hh, uint32_t n, const void* regions, uint32_t kappa, const uint8_t* stream,
                         uint64_t stream_len, uint32_t* alpha, int32_t* negative, int32_t* status);
void hostsim_exact_jk(void* hh, int mode, uint32_t n, const uint32_t* alpha_d, const int32_t* neg_d,
                      const uint32_t* alpha_r, const int32_t* neg_r, const uint32_t* t, uint32_t* j,
                      uint32_t* k);
}

struct qb200_exact {
  void* h = nullptr;
  uint32_t dims[9] = {0};
};

extern "C" {

int qb200_exact_create(qb200_context*, const qb200_params* p, int kind, uint32_t dimension_max, uint32_t emax,
                       qb200_exact** out) {
  *out = nullptr;
  void* h = hostsim_exact_new(kind, p->m, p->l, p->sigma, p->d_be, p->d_len, p->r_be, p->r_len, dimension_max, emax);
  if (!h) {
    g_err = hostsim_last_error();
    return -2;
  }
  qb200_exact* s = new qb200_exact;
  s->h = h;
  hostsim_exact_dims(h, s->dims);
  *out = s;
  return 0;
}
void qb200_exact_destroy(qb200_exact* s) {
  if (!s) return;
  hostsim_exact_free(s->h);
  delete s;
}
void qb200_exact_dims(const qb200_exact* s, uint32_t out[6]) {
  out[0] = s->dims[0];
  out[1] = s->dims[1];
  out[2] = s->dims[2];
  out[3] = s->dims[3];
  out[4] = s->dims[4];
  out[5] = s->dims[8];
}
int qb200_exact_region_bytes(const qb200_exact* s, int32_t min_log_alpha, uint32_t region, uint32_t dimension,
                             uint32_t* bytes) {
  int32_t st = 0;
  *bytes = hostsim_exact_region_bytes(s->h, min_log_alpha, region, dimension, &st);
  if (st == 3) {
    g_err = "exact sampler: region outside the sampler's range";
    return -50;
  }
  if (st) {
    g_err = "exact sampler: a bound within 2^-64 of a half-integer";
    return -51;
  }
  return 0;
}

int qb200_diagk_sample_drawn(qb200_diagk* s, qb200_exact* ex, uint32_t n, const qb200_exact_region* regions,
                             const uint32_t* t_r, const uint8_t* stream, uint64_t stream_len, const int32_t* eta,
                             const long double* pivot, uint32_t delta_bound, uint32_t* k, double* x_hi,
                             double* x_lo, int64_t* delta, int32_t* status, int32_t* exact_status) {
  const uint32_t wa = ex->dims[0], wn = ex->dims[1], kappa_r = ex->dims[4];
  std::vector<uint32_t> alpha((size_t)n * wa), j((size_t)n * wn);
  std::vector<int32_t> neg(n);
  hostsim_exact_alpha(ex->h, n, regions, kappa_r, stream, stream_len, alpha.data(), neg.data(), exact_status);
  hostsim_exact_jk(ex->h, 0, n, nullptr, nullptr, alpha.data(), neg.data(), kappa_r ? t_r : nullptr, j.data(),
                   nullptr);
  return qb200_diagk_sample(s, n, j.data(), eta, pivot, delta_bound, k, x_hi, x_lo, delta, status);
}

// ---- integrators and the resident distribution: HOST-LOGIC stand-ins -------------------------------
// The slice integrators return SYNTHETIC cells here: a pure function of the coordinate, the
// dimension and the cell index (no integration). That is all the host logic of dropin.cpp (batching
// the enumerator list, serving one-slice calls from a batch, speculating on the dimension
// heuristic's upgrades), dropin_collapse.cpp and the resident export of dropin_text.cpp needs to be
// exercised inside the reference's own generator executables on a GPU-less machine
// (tests/test_dropin_host_logic_cpu.py). The numbers mean nothing; the collapse and the text use
// the real arithmetic (client_math.cuh / textfmt.cuh twins).

int hostsim_collapse(int axis, uint32_t max_dim, uint32_t n_src, const uint32_t* dims,
                     const long double* const* cells, long double* out);

void* qb200_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void qb200_host_free(void* p) { free(p); }
size_t qb200_text_bound(size_t n) { return 33 * n; }

static uint64_t shim_mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

static double shim_cell(uint32_t m, int32_t a, int32_t b, uint32_t D, uint64_t idx, uint64_t cells) {
  const int da = abs(abs(a) - (int)m), db = abs(abs(b) - (int)m);
  const double base = ldexp(1e-3, -(da + db)) * (a < 0 ? 0.5 : 1.0);
  const uint64_t h = shim_mix(((uint64_t)(uint32_t)a << 32) ^ (uint32_t)b ^ shim_mix(idx * 1315423911ull + D));
  const double u = 0.5 + (double)(h >> 11) * (1.0 / 9007199254740992.0);
  return (h & 63) == 0 ? -1e-3 * base * u / (double)cells : base * u / (double)cells;
}

int qb200_slice2d_compute(qb200_context*, const qb200_params* p, int, int richardson, uint32_t D, uint32_t n,
                          const int32_t* a_d, const int32_t* a_r, double* cells, long double* tp,
                          long double* te, uint32_t* flags) {
  const uint64_t per = (uint64_t)D * D;
  for (uint32_t i = 0; i < n; i++) {
    long double t = 0;
    for (uint64_t k = 0; k < per; k++) {
      cells[i * per + k] = shim_cell(p->m, a_d[i], a_r[i], D, k, per);
      t += cells[i * per + k];
    }
    if (tp) tp[i] = t;
    if (te) te[i] = (long double)t * 1e-9L;
    if (flags) flags[i] = QB200_FLAG_METHOD_SIMPSON | (richardson ? QB200_FLAG_METHOD_RICHARDSON : 0u);
  }
  return 0;
}

int qb200_slice1d_compute(qb200_context*, const qb200_params* p, int kind, int richardson, uint32_t D,
                          uint32_t n, const int32_t* a, const int32_t* eta, double* cells, long double* tp,
                          uint32_t* flags) {
  for (uint32_t i = 0; i < n; i++) {
    long double t = 0;
    for (uint32_t k = 0; k < D; k++) {
      cells[(size_t)i * D + k] = shim_cell(p->m, a[i], (int32_t)p->m + (eta ? eta[i] : 0) + kind, D, k, D);
      t += cells[(size_t)i * D + k];
    }
    if (tp) tp[i] = t;
    if (flags) flags[i] = QB200_FLAG_METHOD_SIMPSON | (richardson ? QB200_FLAG_METHOD_RICHARDSON : 0u);
  }
  return 0;
}

struct qb200_resident {
  std::vector<std::vector<long double>> cells;
  std::vector<long double> tails;
  std::vector<char> text[2];
  int cur = 0;
};

int qb200_resident_create(qb200_context*, uint32_t n, const uint64_t* n_cells, const long double* const* cells,
                          const long double* tails, qb200_resident** out) {
  qb200_resident* r = new qb200_resident;
  for (uint32_t i = 0; i < n; i++) {
    r->cells.emplace_back(cells[i], cells[i] + n_cells[i]);
    r->tails.push_back(tails ? tails[i] : 0.0L);
  }
  *out = r;
  return 0;
}
void qb200_resident_destroy(qb200_resident* r) { delete r; }

int qb200_resident_collapse2d(qb200_resident* r, int axis, const uint32_t* dimension, uint32_t n_dst,
                              const uint32_t* src_begin, const uint32_t* src_index, uint32_t max_dimension,
                              long double* out) {
  for (uint32_t k = 0; k < n_dst; k++) {
    std::vector<uint32_t> dims;
    std::vector<const long double*> ptrs;
    for (uint32_t s = src_begin[k]; s < src_begin[k + 1]; s++) {
      dims.push_back(dimension[src_index[s]]);
      ptrs.push_back(r->cells[src_index[s]].data());
    }
    if (!hostsim_collapse(axis, max_dimension, (uint32_t)dims.size(), dims.data(), ptrs.data(),
                          out + (size_t)k * max_dimension)) {
      g_err = "collapse: value outside the normal long double range";
      return -4;
    }
  }
  return 0;
}

int qb200_resident_format(qb200_resident* r, uint32_t first, uint32_t count, const char** text, size_t* offsets,
                          size_t* lengths) {
  size_t cap = 64;
  for (uint32_t i = 0; i < count; i++) cap += 34 * (r->cells[first + i].size() + 1);
  std::vector<char>& buf = r->text[r->cur ^= 1];   // two sets, as the library
  buf.resize(cap);
  size_t pos = 0;
  for (uint32_t i = 0; i < count; i++) {
    offsets[i] = pos;
    const std::vector<long double>& c = r->cells[first + i];
    pos += hostsim_text_format_ld(c.data(), c.size(), buf.data() + pos, 0, nullptr);
    pos += hostsim_text_format_ld(&r->tails[first + i], 1, buf.data() + pos, 0, nullptr);
    lengths[i] = pos - offsets[i];
  }
  *text = buf.data();
  return 0;
}

int qb200_resident_format_prefetch(qb200_resident*, uint32_t, uint32_t) { return 0; }

}  // extern "C"
