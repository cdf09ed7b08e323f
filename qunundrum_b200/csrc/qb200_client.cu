// qb200_client.cu -- C ABI of the server-side work on a finished distribution (SURVEY.md
// section 8(f) #2): a stored distribution's cells resident in device memory, collapsed to its
// marginals and formatted for export without a second upload. See include/qunundrum_b200.h
// ("a stored distribution on the device") and kernels_client.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "kernels_client.cuh"

using namespace qb200;

#define QC_CUDA(call)                                                                     \
  do {                                                                                    \
    const cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                                \
      return set_error(-100, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

struct qb200_resident {
  qb200_context* ctx = nullptr;
  CtxView view;
  uint32_t n = 0;
  std::vector<uint64_t> offset;   // first cell of slice i (16-byte units); its tail follows its cells
  std::vector<uint64_t> cells_of;  // cells of slice i
  uint64_t total = 0;             // 16-byte units
  ulonglong2* d_cells = nullptr;
  int* d_status = nullptr;
  // export: two buffer sets, so that a batch can be formatted while the caller still writes the
  // previous one (qb200_resident_format_prefetch)
  struct TextSet {
    char* d_text = nullptr;
    size_t d_text_bytes = 0;
    unsigned long long* d_lens = nullptr;
    size_t lens_count = 0;
    char* h_text = nullptr;
    size_t h_text_bytes = 0;
    unsigned long long* h_lens = nullptr;
    cudaEvent_t done = nullptr;
    bool in_flight = false;
    uint32_t first = 0, count = 0;
    std::vector<size_t> at;   // slice i of the batch starts at at[i] (stride: its capacity)
  } text[2];
  int text_cur = 0;           // the set the last qb200_resident_format returned
  void* h_stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
};

namespace {

const size_t kStageBytes = size_t(32) << 20;

struct Segment {   // bytes [dst, dst + len) of the resident buffer come from src
  uint64_t dst;
  const char* src;
  uint64_t len;
};

// Copy the bytes [lo, hi) of the resident image into stage (which starts at image offset base).
void fill_range(const std::vector<Segment>& segs, uint64_t lo, uint64_t hi, uint64_t base, char* stage) {
  size_t a = 0, b = segs.size();
  while (a < b) {  // first segment that ends after lo
    const size_t mid = (a + b) / 2;
    if (segs[mid].dst + segs[mid].len <= lo) a = mid + 1; else b = mid;
  }
  for (size_t i = a; i < segs.size() && segs[i].dst < hi; i++) {
    const uint64_t s = std::max(lo, segs[i].dst), e = std::min(hi, segs[i].dst + segs[i].len);
    if (e > s) memcpy(stage + (s - base), segs[i].src + (s - segs[i].dst), e - s);
  }
}

void resident_free(qb200_resident* r) {
  if (!r) return;
  cudaSetDevice(r->view.device);
  if (r->d_cells) cudaFree(r->d_cells);
  if (r->d_status) cudaFree(r->d_status);
  for (int k = 0; k < 2; k++) {
    qb200_resident::TextSet& t = r->text[k];
    if (t.in_flight) cudaEventSynchronize(t.done);
    if (t.d_text) cudaFree(t.d_text);
    if (t.d_lens) cudaFree(t.d_lens);
    if (t.h_text) cudaFreeHost(t.h_text);
    if (t.h_lens) cudaFreeHost(t.h_lens);
    if (t.done) cudaEventDestroy(t.done);
  }
  for (int k = 0; k < 2; k++) {
    if (r->h_stage[k]) cudaFreeHost(r->h_stage[k]);
    if (r->stage_done[k]) cudaEventDestroy(r->stage_done[k]);
  }
  delete r;
}

}  // namespace

extern "C" {

int qb200_resident_create(qb200_context* ctx, uint32_t n_slices, const uint64_t* n_cells,
                          const long double* const* cells, const long double* tails,
                          qb200_resident** out) {
  *out = nullptr;
  if (!ctx || (n_slices && (!n_cells || !cells))) return set_error(-1, "null argument");
  qb200_resident* r = new qb200_resident;
  r->ctx = ctx;
  r->view = ctx_view(ctx);
  r->n = n_slices;
  if (cudaSetDevice(r->view.device) != cudaSuccess) {
    delete r;
    return set_error(-100, "cudaSetDevice failed");
  }
  std::vector<Segment> segs;
  segs.reserve(2 * (size_t)n_slices);
  static const long double zero_tail = 0.0L;
  uint64_t pos = 0;
  for (uint32_t i = 0; i < n_slices; i++) {
    r->offset.push_back(pos);
    r->cells_of.push_back(n_cells[i]);
    if (n_cells[i]) segs.push_back(Segment{pos * 16, (const char*)cells[i], n_cells[i] * 16});
    pos += n_cells[i];
    segs.push_back(Segment{pos * 16, (const char*)(tails ? tails + i : &zero_tail), 16});
    pos += 1;
  }
  r->total = pos;
  const uint64_t bytes = pos * 16;
  cudaError_t e = cudaMalloc(&r->d_cells, std::max<uint64_t>(16, bytes));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_status, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(r->d_status, 0, sizeof(int), r->view.stream);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaHostAlloc(&r->h_stage[k], kStageBytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->stage_done[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    resident_free(r);
    return set_error(-100, std::string("resident distribution: ") + cudaGetErrorString(e));
  }
  // The slices are pageable (malloc'ed by the reference's containers): worker threads gather them
  // into one of two pinned staging buffers while the other one is in flight.
  const unsigned workers = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
  int k = 0;
  for (uint64_t base = 0; base < bytes; base += kStageBytes, k ^= 1) {
    const uint64_t len = std::min<uint64_t>(kStageBytes, bytes - base);
    e = cudaEventSynchronize(r->stage_done[k]);   // the copy that last used this buffer
    if (e != cudaSuccess) break;
    char* stage = (char*)r->h_stage[k];
    const uint64_t per = ((len + workers - 1) / workers + 63) & ~uint64_t(63);
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < workers; w++) {
      const uint64_t lo = base + std::min<uint64_t>(len, w * per), hi = base + std::min<uint64_t>(len, (w + 1) * per);
      if (hi > lo) pool.emplace_back(fill_range, std::cref(segs), lo, hi, base, stage);
    }
    fill_range(segs, base, base + std::min<uint64_t>(len, per), base, stage);
    for (std::thread& t : pool) t.join();
    e = cudaMemcpyAsync((char*)r->d_cells + base, stage, len, cudaMemcpyHostToDevice, r->view.stream);
    if (e == cudaSuccess) e = cudaEventRecord(r->stage_done[k], r->view.stream);
    if (e != cudaSuccess) break;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(r->view.stream);
  if (e != cudaSuccess) {
    resident_free(r);
    return set_error(-100, std::string("resident distribution upload: ") + cudaGetErrorString(e));
  }
  *out = r;
  return 0;
}

void qb200_resident_destroy(qb200_resident* r) { resident_free(r); }

uint64_t qb200_resident_cells(const qb200_resident* r) { return r ? r->total - r->n : 0; }

int qb200_resident_collapse2d(qb200_resident* r, int axis, const uint32_t* dimension, uint32_t n_dst,
                              const uint32_t* src_begin, const uint32_t* src_index,
                              uint32_t max_dimension, long double* out) {
  if (!r || !dimension || !src_begin || !src_index || !out) return set_error(-1, "null argument");
  if (axis != 0 && axis != 1) return set_error(-2, "axis must be 0 (alpha_d) or 1 (alpha_r)");
  if (n_dst == 0 || max_dimension == 0) return 0;
  QC_CUDA(cudaSetDevice(r->view.device));
  std::vector<CollapseSrc> srcs(r->n);
  for (uint32_t i = 0; i < r->n; i++) {
    const uint64_t D = dimension[i];
    if (D == 0 || D * D != r->cells_of[i] || max_dimension % D != 0)
      return set_error(-3, "collapse: slice dimensions must match the resident cells and divide the "
                           "maximum dimension");
    srcs[i].offset = r->offset[i];
    srcs[i].dimension = (unsigned)D;
    srcs[i].divisor = (unsigned)(max_dimension / D);
  }
  const uint32_t n_src = src_begin[n_dst];
  for (uint32_t i = 0; i < n_src; i++)
    if (src_index[i] >= r->n) return set_error(-3, "collapse: slice index out of range");
  CollapseSrc* d_srcs = nullptr;
  unsigned *d_begin = nullptr, *d_index = nullptr;
  ulonglong2* d_out = nullptr;
  const size_t out_bytes = (size_t)n_dst * max_dimension * 16;
  cudaError_t e = cudaMalloc(&d_srcs, std::max<size_t>(16, srcs.size() * sizeof(CollapseSrc)));
  if (e == cudaSuccess) e = cudaMalloc(&d_begin, (n_dst + 1) * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&d_index, std::max<size_t>(4, n_src * sizeof(unsigned)));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, out_bytes);
  cudaStream_t st = r->view.stream;
  if (e == cudaSuccess && !srcs.empty())
    e = cudaMemcpyAsync(d_srcs, srcs.data(), srcs.size() * sizeof(CollapseSrc), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_begin, src_begin, (n_dst + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_src)
    e = cudaMemcpyAsync(d_index, src_index, n_src * sizeof(unsigned), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->d_status, 0, sizeof(int), st);
  int status = 0;
  if (e == cudaSuccess) {
    const unsigned per_block = 32 * QB_COLLAPSE_WARPS;
    k_collapse<<<dim3((max_dimension + per_block - 1) / per_block, n_dst), per_block, 0, st>>>(
        axis, max_dimension, d_begin, d_index, d_srcs, r->d_cells, d_out, r->d_status);
    (*r->view.launches)++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&status, r->d_status, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_srcs);
  cudaFree(d_begin);
  cudaFree(d_index);
  cudaFree(d_out);
  if (e != cudaSuccess) return set_error(-100, std::string("collapse: ") + cudaGetErrorString(e));
  if (status)
    return set_error(-4, "collapse: a cell or a sum is outside the normal long double range "
                         "(denormal, infinity or NaN)");
  return 0;
}

// Enqueue the export of the slices [first, first + count) into buffer set `k`: one exporter launch
// per slice (its cells and its tail), the lengths and the text (at the slices' capacities) copied
// to pinned memory behind them; no synchronisation.
static int format_enqueue(qb200_resident* r, int k, uint32_t first, uint32_t count) {
  qb200_resident::TextSet& t = r->text[k];
  cudaStream_t st = r->view.stream;
  if (t.in_flight) {
    QC_CUDA(cudaEventSynchronize(t.done));
    t.in_flight = false;
  }
  t.first = first;
  t.count = count;
  t.at.assign(count + 1, 0);
  for (uint32_t i = 0; i < count; i++)
    t.at[i + 1] = t.at[i] + ((qb200_text_bound(r->cells_of[first + i] + 1) + 63) & ~size_t(63));
  const size_t total = t.at[count];
  if (t.d_text_bytes < total + 64) {
    const size_t prev = t.d_text_bytes;
    if (t.d_text) cudaFree(t.d_text);
    if (t.h_text) cudaFreeHost(t.h_text);
    t.d_text = t.h_text = nullptr;
    t.d_text_bytes = t.h_text_bytes = 0;
    // (pinning costs milliseconds per MB on a virtual machine: grow in steps of at least 50 %)
    const size_t want = std::max(total + 64, prev + prev / 2);
    QC_CUDA(cudaMalloc(&t.d_text, want));
    QC_CUDA(cudaHostAlloc(&t.h_text, want, cudaHostAllocDefault));
    t.d_text_bytes = t.h_text_bytes = want;
  }
  if (t.lens_count < count) {
    if (t.d_lens) cudaFree(t.d_lens);
    if (t.h_lens) cudaFreeHost(t.h_lens);
    t.d_lens = t.h_lens = nullptr;
    t.lens_count = 0;
    const size_t room = std::max<size_t>(count, 64);
    QC_CUDA(cudaMalloc(&t.d_lens, room * sizeof(unsigned long long)));
    QC_CUDA(cudaHostAlloc(&t.h_lens, room * sizeof(unsigned long long), cudaHostAllocDefault));
    t.lens_count = room;
  }
  if (!t.done) QC_CUDA(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
  for (uint32_t i = 0; i < count; i++) {
    const int rc = qb200_text_format_device(r->ctx, QB200_TEXT_X87, r->d_cells + r->offset[first + i],
                                            r->cells_of[first + i] + 1, t.d_text + t.at[i], t.at[i + 1] - t.at[i],
                                            (uint64_t*)(t.d_lens + i), st);
    if (rc) return rc;
  }
  QC_CUDA(cudaMemcpyAsync(t.h_lens, t.d_lens, count * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (total) QC_CUDA(cudaMemcpyAsync(t.h_text, t.d_text, total, cudaMemcpyDeviceToHost, st));
  QC_CUDA(cudaEventRecord(t.done, st));
  t.in_flight = true;
  return 0;
}

int qb200_resident_format_prefetch(qb200_resident* r, uint32_t first, uint32_t count) {
  if (!r) return set_error(-1, "null argument");
  if ((uint64_t)first + count > r->n) return set_error(-2, "slice range outside the resident distribution");
  if (count == 0) return 0;
  QC_CUDA(cudaSetDevice(r->view.device));
  return format_enqueue(r, r->text_cur ^ 1, first, count);
}

int qb200_resident_format(qb200_resident* r, uint32_t first, uint32_t count, const char** text,
                          size_t* offsets, size_t* lengths) {
  if (!r || !text || !offsets || !lengths) return set_error(-1, "null argument");
  if ((uint64_t)first + count > r->n) return set_error(-2, "slice range outside the resident distribution");
  *text = nullptr;
  if (count == 0) return 0;
  QC_CUDA(cudaSetDevice(r->view.device));
  int k = r->text_cur ^ 1;   // a prefetched batch, if it is this one
  if (!(r->text[k].in_flight && r->text[k].first == first && r->text[k].count == count)) {
    if (int rc = format_enqueue(r, k, first, count)) return rc;
  }
  qb200_resident::TextSet& t = r->text[k];
  QC_CUDA(cudaEventSynchronize(t.done));
  t.in_flight = false;
  for (uint32_t i = 0; i < count; i++) {
    if (t.h_lens[i] > t.at[i + 1] - t.at[i]) return set_error(-3, "text buffer overflow (internal)");
    offsets[i] = t.at[i];
    lengths[i] = (size_t)t.h_lens[i];
  }
  r->text_cur = k;
  *text = t.h_text;
  return 0;
}

}  // extern "C"
