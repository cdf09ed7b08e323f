// qb200_text.cu -- C ABI of the text exporter (include/qunundrum_b200.h, "text export").
//
// Replaces the reference's per-cell fprintf("%.24Lg\n") loops
// (src/distribution_slice_import_export.cpp:89-103 and the linear / diagonal
// twins) by one kernel launch per call. There is no CPU formatting path: the
// entry points fail without a CUDA device like the rest of the library.
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "kernels_text.cuh"
#include "text_tables.hpp"

using namespace qb200;
using namespace qb200::text;

#define QT_CUDA(call)                                                                   \
  do {                                                                                  \
    const cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess)                                                              \
      return set_error(-100, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

namespace qb200 {

struct GrowBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool host = false;
  int reserve(size_t n) {
    if (n <= bytes) return 0;
    const size_t want = n + n / 4 + 4096;
    if (p) {
      if (host)
        cudaFreeHost(p);
      else
        cudaFree(p);
      p = nullptr;
      bytes = 0;
    }
    QT_CUDA(host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want));
    bytes = want;
    return 0;
  }
  void release() {
    if (!p) return;
    if (host)
      cudaFreeHost(p);
    else
      cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

struct TextState {
  Pow10Entry* d_tab = nullptr;
  unsigned long long* d_scalars = nullptr;  // [0] total length, [1] exact-path count
  GrowBuf d_in, d_text, d_status;
  GrowBuf d_ptext, d_starts, d_values;       // importer
  unsigned long long* d_info = nullptr;      // importer: INFO_WORDS + 1 (starts[n])
  GrowBuf h_text, h_scalars;
  int force_exact = 0;
  uint64_t exact_total = 0;
  TextState() {
    h_text.host = true;
    h_scalars.host = true;
  }
};

void text_state_destroy(TextState* st) {
  if (!st) return;
  if (st->d_tab) cudaFree(st->d_tab);
  if (st->d_scalars) cudaFree(st->d_scalars);
  st->d_in.release();
  st->d_text.release();
  st->d_status.release();
  st->d_ptext.release();
  st->d_starts.release();
  st->d_values.release();
  if (st->d_info) cudaFree(st->d_info);
  st->h_text.release();
  st->h_scalars.release();
  delete st;
}

}  // namespace qb200

namespace {

const std::vector<Pow10Entry>& host_table() {
  static std::vector<Pow10Entry> t;
  static std::once_flag once;
  std::call_once(once, [] { build_pow10_table(t); });
  return t;
}

int get_state(qb200_context* ctx, CtxView* view, TextState** out) {
  if (!ctx) return set_error(-1, "null context");
  *view = ctx_view(ctx);
  QT_CUDA(cudaSetDevice(view->device));
  if (!*view->text) {
    TextState* st = new TextState;
    const std::vector<Pow10Entry>& t = host_table();
    cudaError_t e = cudaMalloc(&st->d_tab, t.size() * sizeof(Pow10Entry));
    if (e == cudaSuccess) e = cudaMalloc(&st->d_scalars, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&st->d_info, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess)
      e = cudaMemcpy(st->d_tab, t.data(), t.size() * sizeof(Pow10Entry), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(st->d_scalars, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
      text_state_destroy(st);
      return set_error(-100, std::string("text tables: ") + cudaGetErrorString(e));
    }
    if (int rc = st->h_scalars.reserve(64)) {
      text_state_destroy(st);
      return rc;
    }
    *view->text = st;
  }
  *out = *view->text;
  return 0;
}

// Enqueue one formatting pass; d_len receives the text length.
int enqueue_format(const CtxView& view, TextState* st, int kind, const void* d_values, size_t n,
                   char* d_text, size_t cap, unsigned long long* d_len, cudaStream_t stream) {
  if (n == 0) {
    QT_CUDA(cudaMemsetAsync(d_len, 0, sizeof(unsigned long long), stream));
    return 0;
  }
  const size_t n_tiles = (n + TILE - 1) / TILE;
  if (n_tiles > 0xffffffffULL) return set_error(-2, "too many values for one call");
  const size_t st_bytes = (n_tiles + 1) * sizeof(unsigned long long);
  if (int rc = st->d_status.reserve(st_bytes)) return rc;
  QT_CUDA(cudaMemsetAsync(st->d_status.p, 0, st_bytes, stream));
  unsigned long long* status = (unsigned long long*)st->d_status.p;
  unsigned int* ticket = (unsigned int*)(status + n_tiles);
  // the exact-path counter is per call (qb200_text_exact_count: "of the last call")
  QT_CUDA(cudaMemsetAsync(st->d_scalars + 1, 0, sizeof(unsigned long long), stream));
  if (kind == QB200_TEXT_X87)
    k_text_format<SRC_X87><<<(unsigned)n_tiles, TEXT_THREADS, 0, stream>>>(
        d_values, n, st->d_tab, (unsigned char*)d_text, cap, status, ticket, d_len,
        st->d_scalars + 1, st->force_exact);
  else
    k_text_format<SRC_F64><<<(unsigned)n_tiles, TEXT_THREADS, 0, stream>>>(
        d_values, n, st->d_tab, (unsigned char*)d_text, cap, status, ticket, d_len,
        st->d_scalars + 1, st->force_exact);
  QT_CUDA(cudaGetLastError());
  (*view.launches)++;
  return 0;
}

int format_host(qb200_context* ctx, int kind, const void* values, size_t n, const void* tail,
                const char** text, size_t* len) {
  if (!text || !len || (n && !values)) return set_error(-1, "null argument");
  CtxView view;
  TextState* st = nullptr;
  if (int rc = get_state(ctx, &view, &st)) return rc;
  const size_t width = kind == QB200_TEXT_X87 ? 16 : 8;
  const size_t total_n = n + (tail ? 1 : 0);
  if (int rc = st->d_in.reserve(std::max<size_t>(16, total_n * width))) return rc;
  if (int rc = st->d_text.reserve(std::max<size_t>(64, total_n * MAX_TEXT))) return rc;
  if (n)
    QT_CUDA(cudaMemcpyAsync(st->d_in.p, values, n * width, cudaMemcpyHostToDevice, view.stream));
  if (tail)
    QT_CUDA(cudaMemcpyAsync((char*)st->d_in.p + n * width, tail, width, cudaMemcpyHostToDevice,
                            view.stream));
  if (int rc = enqueue_format(view, st, kind, st->d_in.p, total_n, (char*)st->d_text.p,
                              st->d_text.bytes, st->d_scalars, view.stream))
    return rc;
  unsigned long long* hs = (unsigned long long*)st->h_scalars.p;
  QT_CUDA(cudaMemcpyAsync(hs, st->d_scalars, 2 * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, view.stream));
  QT_CUDA(cudaStreamSynchronize(view.stream));
  const size_t total = (size_t)hs[0];
  st->exact_total = hs[1];
  if (total > st->d_text.bytes) return set_error(-3, "text buffer overflow (internal)");
  if (int rc = st->h_text.reserve(std::max<size_t>(64, total))) return rc;
  if (total) {
    QT_CUDA(cudaMemcpyAsync(st->h_text.p, st->d_text.p, total, cudaMemcpyDeviceToHost,
                            view.stream));
    QT_CUDA(cudaStreamSynchronize(view.stream));
  }
  *text = (const char*)st->h_text.p;
  *len = total;
  return 0;
}

// Enqueue the two parser passes. d_info: INFO_WORDS words + one for starts[n].
int enqueue_parse(const CtxView& view, TextState* st, const char* d_text, size_t len, size_t n,
                  void* d_values, unsigned long long* d_info, cudaStream_t stream) {
  const unsigned long long init[INFO_WORDS + 1] = {0, 0, ~0ULL, 0, (unsigned long long)len};
  QT_CUDA(cudaMemcpyAsync(d_info, init, sizeof init, cudaMemcpyHostToDevice, stream));
  if (len == 0) return 0;
  const size_t per_tile = (size_t)TOK_TILE_BYTES;
  const size_t n_tiles = (len + per_tile - 1) / per_tile;
  if (n_tiles > 0xffffffffULL) return set_error(-2, "text too large for one call");
  const size_t st_bytes = (n_tiles + 1) * sizeof(unsigned long long);
  if (int rc = st->d_status.reserve(st_bytes)) return rc;
  if (int rc = st->d_starts.reserve((n + 2) * sizeof(unsigned long long))) return rc;
  QT_CUDA(cudaMemsetAsync(st->d_status.p, 0, st_bytes, stream));
  unsigned long long* status = (unsigned long long*)st->d_status.p;
  unsigned int* ticket = (unsigned int*)(status + n_tiles);
  unsigned long long* starts = (unsigned long long*)st->d_starts.p;
  // starts[n] defaults to len: "everything consumed" when no further number follows
  QT_CUDA(cudaMemcpyAsync(starts + n, &init[INFO_WORDS], sizeof(unsigned long long),
                          cudaMemcpyHostToDevice, stream));
  k_text_tokenize<<<(unsigned)n_tiles, TB, 0, stream>>>((const unsigned char*)d_text, len, n, starts,
                                                        status, ticket, d_info);
  QT_CUDA(cudaGetLastError());
  (*view.launches)++;
  if (n) {
    // tokens beyond those present leave their entries untouched: the host checks the count first
    k_text_parse<<<(unsigned)((n + TB - 1) / TB), TB, 0, stream>>>(
        (const unsigned char*)d_text, len, n, starts, st->d_tab, (ulonglong2*)d_values, d_info,
        st->force_exact);
    QT_CUDA(cudaGetLastError());
    (*view.launches)++;
  }
  // starts[n] (the position after the n-th number and the white space behind it)
  QT_CUDA(cudaMemcpyAsync(d_info + INFO_WORDS, starts + n, sizeof(unsigned long long),
                          cudaMemcpyDeviceToDevice, stream));
  return 0;
}

}  // namespace

extern "C" {

size_t qb200_text_bound(size_t n) { return n * (size_t)MAX_TEXT; }

int qb200_text_format_ld(qb200_context* ctx, const long double* values, size_t n,
                         const long double* tail, const char** text, size_t* len) {
  static_assert(sizeof(long double) == 16, "x86-64 long double expected");
  return format_host(ctx, QB200_TEXT_X87, values, n, tail, text, len);
}

int qb200_text_format_f64(qb200_context* ctx, const double* values, size_t n, const double* tail,
                          const char** text, size_t* len) {
  return format_host(ctx, QB200_TEXT_F64, values, n, tail, text, len);
}

int qb200_text_format_device(qb200_context* ctx, int kind, const void* d_values, size_t n,
                             char* d_text, size_t cap, uint64_t* d_len, void* stream) {
  if (kind != QB200_TEXT_X87 && kind != QB200_TEXT_F64) return set_error(-1, "bad kind");
  CtxView view;
  TextState* st = nullptr;
  if (int rc = get_state(ctx, &view, &st)) return rc;
  return enqueue_format(view, st, kind, d_values, n, d_text, cap, (unsigned long long*)d_len,
                        stream ? (cudaStream_t)stream : view.stream);
}

int qb200_text_parse_ld(qb200_context* ctx, const char* text, size_t len, size_t n,
                        long double* values, size_t* consumed) {
  if ((len && !text) || (n && !values)) return set_error(-1, "null argument");
  CtxView view;
  TextState* st = nullptr;
  if (int rc = get_state(ctx, &view, &st)) return rc;
  const size_t padded = (len + 15) / 16 * 16 + 16;
  if (int rc = st->d_ptext.reserve(padded)) return rc;
  if (int rc = st->d_values.reserve(std::max<size_t>(16, n * 16))) return rc;
  if (int rc = st->d_starts.reserve((n + 2) * sizeof(unsigned long long))) return rc;
  if (len) {
    QT_CUDA(cudaMemcpyAsync(st->d_ptext.p, text, len, cudaMemcpyHostToDevice, view.stream));
    QT_CUDA(cudaMemsetAsync((char*)st->d_ptext.p + len, ' ', padded - len, view.stream));
  }
  if (int rc = enqueue_parse(view, st, (const char*)st->d_ptext.p, len, n, st->d_values.p,
                             st->d_info, view.stream))
    return rc;
  unsigned long long* hs = (unsigned long long*)st->h_scalars.p;
  QT_CUDA(cudaMemcpyAsync(hs, st->d_info, (INFO_WORDS + 1) * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, view.stream));
  QT_CUDA(cudaStreamSynchronize(view.stream));
  st->exact_total = hs[INFO_EXACT];
  if (hs[INFO_TOKENS] < n)
    return set_error(-20, "text holds " + std::to_string(hs[INFO_TOKENS]) + " numbers, " +
                              std::to_string(n) + " expected");
  if (hs[INFO_STATUS] != 0)
    return set_error(hs[INFO_STATUS] == 1 ? -21 : -22,
                     std::string(hs[INFO_STATUS] == 1 ? "malformed number" : "unsupported number form") +
                         " (number " + std::to_string(hs[INFO_FIRST_BAD]) + " of the text)");
  if (n) {
    QT_CUDA(cudaMemcpyAsync(values, st->d_values.p, n * 16, cudaMemcpyDeviceToHost, view.stream));
    QT_CUDA(cudaStreamSynchronize(view.stream));
  }
  if (consumed) *consumed = hs[INFO_TOKENS] > n ? (size_t)hs[INFO_WORDS] : len;
  return 0;
}

int qb200_text_parse_device(qb200_context* ctx, const char* d_text, size_t len, size_t n,
                            void* d_values, uint64_t* d_info, void* stream) {
  CtxView view;
  TextState* st = nullptr;
  if (int rc = get_state(ctx, &view, &st)) return rc;
  return enqueue_parse(view, st, d_text, len, n, d_values, (unsigned long long*)d_info,
                       stream ? (cudaStream_t)stream : view.stream);
}

int qb200_text_pow10(int k, uint32_t w[6], int32_t* e2, uint32_t* exact) {
  if (k < K_MIN || k > K_MAX) return set_error(-1, "power of ten out of range");
  const Pow10Entry& e = host_table()[(size_t)(k - K_MIN)];
  for (int i = 0; i < 6; i++) w[i] = e.w[i];
  *e2 = e.e2;
  *exact = e.exact;
  return 0;
}

int qb200_text_set_force_exact(qb200_context* ctx, int on) {
  CtxView view;
  TextState* st = nullptr;
  if (int rc = get_state(ctx, &view, &st)) return rc;
  st->force_exact = on;  // 1: exact decision for every value; 2: timing experiment (no chain)
  return 0;
}

uint64_t qb200_text_exact_count(qb200_context* ctx) {
  if (!ctx) return 0;
  CtxView view = ctx_view(ctx);
  return *view.text ? (*view.text)->exact_total : 0;
}

}  // extern "C"
