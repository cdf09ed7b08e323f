// qb200_sampler.cu -- C ABI of the distribution sampler (include/qunundrum_b200.h, "sampling").
//
// Replaces the per-sample walks of the reference's tau_estimate / tau_estimate_linear
// (src/tau_estimate.cpp:23-133) and of distribution_sample_approximate_alpha_d_r /
// linear_distribution_sample_approximate_alpha (src/distribution.cpp:464,
// src/linear_distribution.cpp:618) by batched kernels over a device-resident copy of the
// distribution. There is no CPU sampling path: the entry points fail without a CUDA device like
// the rest of the library. The host does what is inherently serial and tiny: which estimate
// starts at which word of the random stream (an estimate stops at its first out-of-bounds
// sample, src/tau_estimate.cpp:46-55), and the final log2 in long double.
#include <cuda_runtime.h>

#include <cstdlib>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "hostconst.hpp"
#include "kernels_sampler.cuh"
#include "sampler_host.hpp"

using namespace qb200;

#define QS_CUDA(call)                                                                   \
  do {                                                                                  \
    const cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess)                                                              \
      return set_error(-100, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

namespace {

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  bool host = false;
  ~Buf() { release(); }
  void release() {
    if (!p) return;
    if (host)
      cudaFreeHost(p);
    else
      cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int reserve(size_t n) {
    if (n <= bytes) return 0;
    release();
    const size_t want = n + n / 4 + 256;
    QS_CUDA(host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want));
    bytes = want;
    return 0;
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

}  // namespace

struct qb200_sampler {
  qb200_context* ctx = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t* launches = nullptr;
  SamplerView view;
  Buf cells, coarse, slices, totals, geo, scratch, guide, cells_d, totals_d;
  // per-call staging
  Buf d_words, d_off, d_out, d_sums, d_status;
  Buf h_sums, h_status, h_words, h_off, h_out;
  unsigned long long* d_exact = nullptr;
  int force_exact = 0;
  uint64_t exact_last = 0;
  uint64_t n_cells = 0;
  // the serial part of the reference's semantics (sampler_host.hpp): the smallest pivot word
  // whose walk over the slices runs out of bounds, decided once in the host's own long double
  FailureThreshold fail;
  TauLayout layout;
  qb200_sampler() {
    h_sums.host = h_status.host = h_words.host = h_off.host = h_out.host = true;
  }
};

// All samples of a call: one thread per sample, in draw order.
static int enqueue_samples(qb200_sampler* s, const uint64_t* d_words, const uint64_t* d_off, uint32_t n,
                           uint64_t total, cudaStream_t st) {
  k_sample<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(s->view, d_words, d_off, n, total,
                                                           s->force_exact, s->d_out.as<SampleOut>());
  *s->launches += 1;
  return 0;
}

extern "C" {

int qb200_sampler_create(qb200_context* ctx, int dims, uint32_t m, uint32_t n_slices,
                         const uint32_t* dimension, const int32_t* c0, const int32_t* c1,
                         const long double* const* cells, const long double* slice_total,
                         long double total_probability, qb200_sampler** out) {
  *out = nullptr;
  if (!ctx || !dimension || !c0 || !cells || !slice_total || (dims == 2 && !c1))
    return set_error(-1, "null argument");
  if (dims != QB200_SAMPLER_LINEAR && dims != QB200_SAMPLER_2D)
    return set_error(-11, "unknown distribution kind");
  if (n_slices == 0) return set_error(-12, "the distribution has no slices");
  const CtxView cv = ctx_view(ctx);
  QS_CUDA(cudaSetDevice(cv.device));
  std::unique_ptr<qb200_sampler> s(new qb200_sampler);
  s->ctx = ctx;
  s->device = cv.device;
  s->stream = cv.stream;
  s->launches = cv.launches;
  // layout
  std::vector<SamplerSlice> hs(n_slices);
  std::map<uint32_t, uint32_t> geo_off;
  std::vector<DD> geo;
  uint64_t cell_off = 0, coarse_off = 0, guide_off = 0;
  for (uint32_t i = 0; i < n_slices; i++) {
    const uint32_t D = dimension[i];
    if (D == 0 || (D & (D - 1)) != 0 || D > (1u << 20) || (dims == 2 && D > 4096))
      return set_error(-12, "slice dimension must be a power of two (at most 4096 for two-dimensional slices)");
    for (int ax = 0; ax < dims; ax++) {
      const long k = std::labs((long)(ax == 0 ? c0[i] : c1[i]));
      if (k - (long)m < -400 || k - (long)m > 59)
        return set_error(-13, "slice coordinate outside the supported range m - 400 <= |min_log_alpha| <= m + 59");
    }
    if (!geo_off.count(D)) {
      geo_off[D] = (uint32_t)geo.size();
      std::vector<DD> t((size_t)D + 1);
      exp2_table_dd(D, t.data());
      geo.insert(geo.end(), t.begin(), t.end());
    }
    SamplerSlice& sl = hs[i];
    sl.cell_off = cell_off;
    sl.coarse_off = coarse_off;
    sl.n_cells = dims == 2 ? D * D : D;
    sl.D = D;
    sl.c0 = c0[i];
    sl.c1 = dims == 2 ? c1[i] : 0;
    sl.geo_off = geo_off[D];
    sl.guide_off = (uint32_t)guide_off;
    sl.abs_sum = 0.0;
    cell_off += sl.n_cells;
    const uint32_t nb = (sl.n_cells + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
    coarse_off += nb + 1;
    guide_off += seg_guide_size(nb) + 1;
  }
  const uint64_t totals_coarse_off = coarse_off, totals_guide_off = guide_off;
  coarse_off += (n_slices + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK + 1;
  guide_off += seg_guide_size((n_slices + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK) + 1;
  if (guide_off > 0xffffffffull) return set_error(-12, "too many slices for the sampler's guide tables");
  s->n_cells = cell_off;
  if (int rc = s->cells.reserve(cell_off * sizeof(RawX87))) return rc;
  if (int rc = s->coarse.reserve(coarse_off * sizeof(SegCoarse))) return rc;
  if (int rc = s->slices.reserve((size_t)n_slices * sizeof(SamplerSlice))) return rc;
  if (int rc = s->totals.reserve((size_t)n_slices * sizeof(RawX87))) return rc;
  if (int rc = s->geo.reserve(geo.size() * sizeof(DD))) return rc;
  if (int rc = s->guide.reserve(guide_off * sizeof(uint32_t))) return rc;
  // the elements once more as doubles, for the quick pass (one block of slack: seg_block_doubles
  // reads whole blocks); QB200_SAMPLER_DOUBLES=0: not kept (A/B, tests)
  const char* dbl_env = getenv("QB200_SAMPLER_DOUBLES");
  const bool keep_doubles = !(dbl_env && *dbl_env == '0');
  if (keep_doubles) {
    if (int rc = s->cells_d.reserve((cell_off + QB_SEG_BLOCK) * sizeof(double))) return rc;
    if (int rc = s->totals_d.reserve(((size_t)n_slices + QB_SEG_BLOCK) * sizeof(double))) return rc;
    QS_CUDA(cudaMemsetAsync((char*)s->cells_d.p + cell_off * sizeof(double), 0, QB_SEG_BLOCK * sizeof(double),
                            s->stream));
    QS_CUDA(cudaMemsetAsync((char*)s->totals_d.p + (size_t)n_slices * sizeof(double), 0,
                            QB_SEG_BLOCK * sizeof(double), s->stream));
  }
  // scratch: segment descriptors, the totals' abs sum, the "bad value" flag
  const size_t seg_bytes = ((size_t)n_slices + 1) * sizeof(SegDesc);
  if (int rc = s->scratch.reserve(seg_bytes + 64)) return rc;
  static_assert(sizeof(long double) == sizeof(RawX87), "x86-64 long double expected");
  for (uint32_t i = 0; i < n_slices; i++)
    QS_CUDA(cudaMemcpyAsync(s->cells.as<RawX87>() + hs[i].cell_off, cells[i],
                            (size_t)hs[i].n_cells * sizeof(RawX87), cudaMemcpyHostToDevice, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->totals.p, slice_total, (size_t)n_slices * sizeof(RawX87),
                          cudaMemcpyHostToDevice, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->slices.p, hs.data(), (size_t)n_slices * sizeof(SamplerSlice),
                          cudaMemcpyHostToDevice, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->geo.p, geo.data(), geo.size() * sizeof(DD), cudaMemcpyHostToDevice, s->stream));
  std::vector<SegDesc> segs(n_slices + 1);
  double* d_tot_abs = (double*)((char*)s->scratch.p + seg_bytes);
  int* d_bad = (int*)((char*)s->scratch.p + seg_bytes + 8);
  for (uint32_t i = 0; i < n_slices; i++) {
    segs[i].vals = s->cells.as<RawX87>() + hs[i].cell_off;
    segs[i].vals_d = keep_doubles ? s->cells_d.as<double>() + hs[i].cell_off : nullptr;
    segs[i].coarse = s->coarse.as<SegCoarse>() + hs[i].coarse_off;
    segs[i].abs_out = &(s->slices.as<SamplerSlice>()[i].abs_sum);
    segs[i].guide = s->guide.as<uint32_t>() + hs[i].guide_off;
    segs[i].n = hs[i].n_cells;
    segs[i].pad = 0;
  }
  segs[n_slices].vals = s->totals.as<RawX87>();
  segs[n_slices].vals_d = keep_doubles ? s->totals_d.as<double>() : nullptr;
  segs[n_slices].coarse = s->coarse.as<SegCoarse>() + totals_coarse_off;
  segs[n_slices].abs_out = d_tot_abs;
  segs[n_slices].guide = s->guide.as<uint32_t>() + totals_guide_off;
  segs[n_slices].n = n_slices;
  segs[n_slices].pad = 0;
  QS_CUDA(cudaMemsetAsync(d_tot_abs, 0, 16, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->scratch.p, segs.data(), seg_bytes, cudaMemcpyHostToDevice, s->stream));
  k_seg_build<<<n_slices + 1, 256, 0, s->stream>>>(s->scratch.as<SegDesc>(), d_bad);
  *s->launches += 1;
  QS_CUDA(cudaGetLastError());
  struct {
    double abs;
    int bad, pad;
  } back;
  QS_CUDA(cudaMemcpyAsync(&back, d_tot_abs, 16, cudaMemcpyDeviceToHost, s->stream));
  QS_CUDA(cudaStreamSynchronize(s->stream));
  if (back.bad)
    return set_error(-14, "the distribution holds a denormal, infinite or NaN probability (not supported)");
  QS_CUDA(cudaMalloc(&s->d_exact, sizeof(unsigned long long)));
  SamplerView& v = s->view;
  v.cells = s->cells.as<RawX87>();
  v.coarse = s->coarse.as<SegCoarse>();
  v.slices = s->slices.as<SamplerSlice>();
  v.totals = s->totals.as<RawX87>();
  v.cells_d = keep_doubles ? s->cells_d.as<double>() : nullptr;
  v.totals_d = keep_doubles ? s->totals_d.as<double>() : nullptr;
  v.totals_coarse = s->coarse.as<SegCoarse>() + totals_coarse_off;
  v.geo = s->geo.as<dd>();
  {
    const char* g = getenv("QB200_SAMPLER_GUIDE");   // 0: plain binary searches (A/B, tests)
    v.guide = (g && *g == '0') ? nullptr : s->guide.as<uint32_t>();
  }
  v.totals_guide_off = (uint32_t)totals_guide_off;
  v.pad0 = 0;
  v.totals_abs_sum = back.abs;
  std::memset(&v.dist_total, 0, sizeof(v.dist_total));
  std::memcpy(&v.dist_total, &total_probability, 10);
  v.n_slices = n_slices;
  v.scale_by_total = total_probability > 1 ? 1 : 0;
  v.m = (int)m;
  v.dims = dims;
  {
    X87 t;
    if (!x87_decode(v.dist_total.mant, (uint32_t)v.dist_total.se & 0xffffu, &t))
      return set_error(-14, "the distribution's total probability is denormal, infinite or NaN");
  }
  s->fail = find_failure_threshold(slice_total, n_slices, total_probability);
  *out = s.release();
  return 0;
}

void qb200_sampler_destroy(qb200_sampler* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->d_exact) cudaFree(s->d_exact);
  delete s;
}

uint32_t qb200_sampler_words_per_sample(const qb200_sampler* s) { return (uint32_t)s->view.dims + 2u; }
uint64_t qb200_sampler_cells(const qb200_sampler* s) { return s->n_cells; }

int qb200_sampler_set_force_exact(qb200_sampler* s, int on) {
  s->force_exact = (on == 1 || on == 2) ? on : 0;
  return 0;
}
uint64_t qb200_sampler_exact_count(const qb200_sampler* s) { return s->exact_last; }

int qb200_sampler_first_failing_word(const qb200_sampler* s, uint64_t* word) {
  if (word) *word = s->fail.first;
  return s->fail.any ? 1 : 0;
}

int qb200_sampler_sample(qb200_sampler* s, uint32_t k, const uint64_t* words, int32_t* slice,
                         int32_t* cell, double* x0, double* x1, int32_t* status) {
  if (!s || (k && !words)) return set_error(-1, "null argument");
  QS_CUDA(cudaSetDevice(s->device));
  s->exact_last = 0;
  if (k == 0) return 0;
  const uint32_t wps = qb200_sampler_words_per_sample(s);
  if (int rc = s->d_words.reserve((size_t)k * wps * 8)) return rc;
  if (int rc = s->d_out.reserve((size_t)k * sizeof(SampleOut))) return rc;
  if (int rc = s->h_out.reserve((size_t)k * sizeof(SampleOut))) return rc;
  QS_CUDA(cudaMemcpyAsync(s->d_words.p, words, (size_t)k * wps * 8, cudaMemcpyHostToDevice, s->stream));
  k_sample<<<(unsigned)((k + 127) / 128), 128, 0, s->stream>>>(s->view, s->d_words.as<uint64_t>(), nullptr,
                                                              1, k, s->force_exact, s->d_out.as<SampleOut>());
  *s->launches += 1;
  QS_CUDA(cudaGetLastError());
  QS_CUDA(cudaMemcpyAsync(s->h_out.p, s->d_out.p, (size_t)k * sizeof(SampleOut), cudaMemcpyDeviceToHost,
                          s->stream));
  QS_CUDA(cudaStreamSynchronize(s->stream));
  const SampleOut* o = s->h_out.as<SampleOut>();
  for (uint32_t i = 0; i < k; i++) {
    if (slice) slice[i] = o[i].slice;
    if (cell) cell[i] = o[i].cell;
    if (x0) x0[i] = o[i].x0;
    if (x1) x1[i] = o[i].x1;
    if (status) status[i] = o[i].status;
    s->exact_last += (uint64_t)o[i].exact;
  }
  return 0;
}

int qb200_sampler_tau_device(qb200_sampler* s, uint32_t n, uint32_t count, const uint64_t* d_words,
                             double* d_sums, int32_t* d_status, void* stream) {
  if (!s || !d_words || !d_sums || !d_status) return set_error(-1, "null argument");
  if (n == 0) return set_error(-15, "n must be positive");
  QS_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
  const uint64_t total = (uint64_t)n * count;
  if (total == 0) return 0;
  if (int rc = s->d_out.reserve(total * sizeof(SampleOut))) return rc;
  if (int rc = enqueue_samples(s, d_words, nullptr, n, total, st)) return rc;
  k_tau_reduce<<<(count + 127) / 128, 128, 0, st>>>(s->d_out.as<SampleOut>(), n, count, d_sums, d_status,
                                                   s->d_exact);
  *s->launches += 1;
  QS_CUDA(cudaGetLastError());
  return 0;
}

int qb200_sampler_tau_estimate(qb200_sampler* s, uint32_t n, uint32_t count, const uint64_t* words,
                               size_t n_words, size_t* words_used, uint32_t* done, long double* tau0,
                               long double* tau1, uint8_t* ok) {
  if (!s || !words || !tau0 || !ok || (s->view.dims == 2 && !tau1)) return set_error(-1, "null argument");
  if (n == 0) return set_error(-15, "n must be positive");
  QS_CUDA(cudaSetDevice(s->device));
  const uint32_t wps = qb200_sampler_words_per_sample(s);
  // ---- the serial part: where each estimate starts in the stream -------------------------
  tau_layout(s->fail, wps, n, count, words, n_words, &s->layout);
  const uint32_t nt = s->layout.done;
  const size_t cur = s->layout.words_used;
  const uint64_t* off = s->layout.off.data();
  if (done) *done = nt;
  if (words_used) *words_used = cur;
  s->exact_last = 0;
  if (nt == 0) return 0;
  // ---- device: all samples of all estimates -------------------------------------------------
  const uint64_t total = (uint64_t)n * nt;
  if (int rc = s->d_words.reserve(std::max<size_t>(8, cur * 8))) return rc;
  if (int rc = s->d_off.reserve((size_t)nt * 8)) return rc;
  if (int rc = s->d_out.reserve(total * sizeof(SampleOut))) return rc;
  if (int rc = s->d_sums.reserve((size_t)nt * 32)) return rc;
  if (int rc = s->d_status.reserve((size_t)nt * 4)) return rc;
  if (int rc = s->h_sums.reserve((size_t)nt * 32)) return rc;
  const size_t exact_at = ((size_t)nt * 4 + 7) & ~(size_t)7;  // the replay counter follows the statuses
  if (int rc = s->h_status.reserve(exact_at + 8)) return rc;
  if (cur) QS_CUDA(cudaMemcpyAsync(s->d_words.p, words, cur * 8, cudaMemcpyHostToDevice, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->d_off.p, off, (size_t)nt * 8, cudaMemcpyHostToDevice, s->stream));
  QS_CUDA(cudaMemsetAsync(s->d_exact, 0, sizeof(unsigned long long), s->stream));
  if (int rc = enqueue_samples(s, s->d_words.as<uint64_t>(), s->d_off.as<uint64_t>(), n, total, s->stream))
    return rc;
  k_tau_reduce<<<(nt + 127) / 128, 128, 0, s->stream>>>(s->d_out.as<SampleOut>(), n, nt,
                                                       s->d_sums.as<double>(), s->d_status.as<int>(),
                                                       s->d_exact);
  *s->launches += 1;
  QS_CUDA(cudaGetLastError());
  QS_CUDA(cudaMemcpyAsync(s->h_sums.p, s->d_sums.p, (size_t)nt * 32, cudaMemcpyDeviceToHost, s->stream));
  QS_CUDA(cudaMemcpyAsync(s->h_status.p, s->d_status.p, (size_t)nt * 4, cudaMemcpyDeviceToHost, s->stream));
  unsigned long long* h_exact = (unsigned long long*)(s->h_status.as<char>() + exact_at);
  QS_CUDA(cudaMemcpyAsync(h_exact, s->d_exact, 8, cudaMemcpyDeviceToHost, s->stream));
  QS_CUDA(cudaStreamSynchronize(s->stream));
  s->exact_last = *h_exact;
  // ---- host: tau = log2(mean alpha^2) / 2 - m (src/tau_estimate.cpp:63-71) ------------------
  std::string err;
  const int rc = tau_finish(s->view.dims, s->view.m, n, s->layout, s->h_sums.as<double>(),
                            s->h_status.as<int>(), tau0, tau1, ok, &err);
  return rc ? set_error(rc, err) : 0;
}

}  // extern "C"
