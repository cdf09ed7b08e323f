// qb200_exact.cu -- C ABI of the exact samplers (include/qunundrum_b200.h, "exact samplers").
//
// Replaces sample_alpha_from_region, sample_j_from_alpha_r, sample_j_k_from_alpha_d,
// sample_j_k_from_alpha_d_r and sample_j_from_diagonal_alpha_r (src/sample.cpp:78-410) by batched
// kernels (exact.cuh, kernels_exact.cuh). There is no CPU path: the entry points that compute
// samples fail without a CUDA device like the rest of the library; qb200_exact_region_bytes is
// stream layout (how many bytes random_generate_mpz reads for a region, src/random.c:163-164), which
// the caller needs BEFORE it can hand over a sample's bytes, and runs on the host.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "devbuf.hpp"
#include "exact_api.hpp"
#include "exact_host.hpp"
#include "kernels_exact.cuh"

using namespace qb200;

static_assert(sizeof(qb200_exact_region) == sizeof(ExactRegion), "region record layout");

struct qb200_exact {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  uint64_t* launches = nullptr;
  ExactHost host;
  ExactConst dev;  // pointers into `consts` / `table`
  DBuf consts, table, regions, bytes, scratch, rows, a_d, a_r, neg_d, neg_r, status, t, k, j;
  uint32_t chunk = 0;
  int columns = 8;  // columns per pass of the products (QB200_EXACT_COLUMNS=4: the narrower form, for the A/B)
  // CUDA events around the last k_exact_alpha [0, 1] and k_exact_jk [2, 3] launch (qb200_exact_kernel_ms)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool timed[2] = {false, false};
  ~qb200_exact() {
    for (int i = 0; i < 4; i++)
      if (ev[i]) cudaEventDestroy(ev[i]);
  }
};

namespace {

// Host rows (w words per sample) -> tiles on the device.
int rows_to_tiles(qb200_exact* s, const uint32_t* host_rows, uint32_t w, uint32_t B, DBuf& tiles, cudaStream_t st) {
  const size_t Bp = ((size_t)B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA * QB_DIAGK_CTA;
  if (s->rows.reserve((size_t)B * w * 4) || tiles.reserve(Bp * w * 4)) return -100;
  QD_CUDA(cudaMemcpyAsync(s->rows.p, host_rows, (size_t)B * w * 4, cudaMemcpyHostToDevice, st));
  const uint64_t n = (uint64_t)Bp * w;
  k_exact_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->rows.as<uint32_t>(), w, B, tiles.as<uint32_t>());
  *s->launches += 1;
  QD_CUDA(cudaGetLastError());
  return 0;
}

// Tiles on the device -> host rows (synchronises).
int tiles_to_rows(qb200_exact* s, const DBuf& tiles, uint32_t w, uint32_t B, uint32_t* host_rows, cudaStream_t st) {
  if (s->rows.reserve((size_t)B * w * 4)) return -100;
  const uint64_t n = (uint64_t)B * w;
  k_exact_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tiles.as<uint32_t>(), w, B, s->rows.as<uint32_t>());
  *s->launches += 1;
  QD_CUDA(cudaGetLastError());
  QD_CUDA(cudaMemcpyAsync(host_rows, s->rows.p, (size_t)B * w * 4, cudaMemcpyDeviceToHost, st));
  QD_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int host_ints_to_device(const int32_t* host, uint32_t B, DBuf& buf, cudaStream_t st) {
  if (buf.reserve((size_t)B * 4)) return -100;
  QD_CUDA(cudaMemcpyAsync(buf.p, host, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  return 0;
}

uint32_t tiles_of(uint32_t B) { return (B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA; }

// alpha tiles of B samples from regions (host) and the stream (device) -> `alphaT`, s->neg_r, s->status.
int launch_alpha(qb200_exact* s, uint32_t B, const qb200_exact_region* regions, uint32_t kappa,
                 const uint8_t* d_stream, uint64_t stream_len, DBuf& alphaT, DBuf& neg, cudaStream_t st) {
  const ExactConst& c = s->host.c;
  const uint32_t grid = tiles_of(B);
  if (s->regions.reserve((size_t)B * sizeof(ExactRegion)) || alphaT.reserve((size_t)grid * QB_DIAGK_CTA * c.wa * 4) ||
      neg.reserve((size_t)B * 4) || s->status.reserve((size_t)B * 4) ||
      s->scratch.reserve((size_t)grid * QB_DIAGK_CTA * exact_alpha_scratch_limbs(c) * 4))
    return -100;
  QD_CUDA(cudaMemcpyAsync(s->regions.p, regions, (size_t)B * sizeof(ExactRegion), cudaMemcpyHostToDevice, st));
  QD_CUDA(cudaEventRecord(s->ev[0], st));
  k_exact_alpha<<<grid, QB_DIAGK_CTA, 0, st>>>(s->dev, s->regions.as<ExactRegion>(), kappa, d_stream,
                                               (unsigned long long)stream_len, B, s->scratch.as<uint32_t>(),
                                               alphaT.as<uint32_t>(), neg.as<int32_t>(), s->status.as<int32_t>());
  QD_CUDA(cudaEventRecord(s->ev[1], st));
  s->timed[0] = true;
  *s->launches += 1;
  QD_CUDA(cudaGetLastError());
  return 0;
}

// j (and k) tiles of B samples from alpha tiles.
int launch_jk(qb200_exact* s, int mode, uint32_t B, const uint32_t* adT, const int32_t* neg_d, const uint32_t* arT,
              const int32_t* neg_r, const uint32_t* tT, uint32_t tl, uint32_t* kT, uint32_t* jT, cudaStream_t st) {
  const ExactConst& c = s->host.c;
  const uint32_t grid = tiles_of(B);
  if (s->scratch.reserve((size_t)grid * QB_DIAGK_CTA * exact_jk_scratch_limbs(c) * 4)) return -100;
  const size_t shmem = (size_t)exact_const_words(c) * 4;
  QD_CUDA(cudaEventRecord(s->ev[2], st));
  if (s->columns == 8)
    k_exact_jk<8><<<grid, QB_DIAGK_CTA, shmem, st>>>(s->dev, mode, adT, neg_d, arT, neg_r, tT, tl, kT, B,
                                                     s->scratch.as<uint32_t>(), jT);
  else
    k_exact_jk<4><<<grid, QB_DIAGK_CTA, shmem, st>>>(s->dev, mode, adT, neg_d, arT, neg_r, tT, tl, kT, B,
                                                     s->scratch.as<uint32_t>(), jT);
  QD_CUDA(cudaEventRecord(s->ev[3], st));
  s->timed[1] = true;
  *s->launches += 1;
  QD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

namespace qb200 {

uint32_t exact_chunk(const qb200_exact* s) { return s->chunk; }
uint32_t exact_j_limbs(const qb200_exact* s) { return s->host.c.wn; }
int exact_device(const qb200_exact* s) { return s->device; }

int exact_upload_stream(qb200_exact* s, const uint8_t* stream, uint64_t stream_len, const uint8_t** d_stream,
                        cudaStream_t st) {
  // k_exact_alpha reads aligned 32-bit words: up to three bytes past the last sample's end
  if (s->bytes.reserve(stream_len + 16)) return -100;
  if (stream_len) QD_CUDA(cudaMemcpyAsync(s->bytes.p, stream, stream_len, cudaMemcpyHostToDevice, st));
  *d_stream = s->bytes.as<uint8_t>();
  return 0;
}

// The diagonal pipeline's first half (diagonal_distribution_sample_alpha_r + sample_j_from_diagonal_alpha_r,
// src/diagonal_distribution.cpp:355-472): j tiles of B samples from their regions and bytes. t: host
// rows (NULL when kappa_r = 0). *d_jT, *d_status point into the sampler's buffers (valid until
// its next call).
int exact_draw_j_tiles(qb200_exact* s, uint32_t B, const qb200_exact_region* regions, const uint32_t* t,
                       const uint8_t* d_stream, uint64_t stream_len, const uint32_t** d_jT,
                       const int32_t** d_status, cudaStream_t st) {
  const ExactConst& c = s->host.c;
  const uint32_t tl = c.kappa_r ? (c.kappa_r + 31) / 32 : 1;
  int rc = launch_alpha(s, B, regions, c.kappa_r, d_stream, stream_len, s->a_r, s->neg_r, st);
  if (rc) return rc;
  if (c.kappa_r) {
    if (!t) return set_error(-1, "exact sampler: t_r is needed when r is even");
    rc = rows_to_tiles(s, t, tl, B, s->t, st);
    if (rc) return rc;
  }
  if (s->j.reserve((size_t)tiles_of(B) * QB_DIAGK_CTA * c.wn * 4)) return -100;
  rc = launch_jk(s, QB_EXACT_J_FROM_ALPHA_R, B, nullptr, nullptr, s->a_r.as<uint32_t>(), s->neg_r.as<int32_t>(),
                 c.kappa_r ? s->t.as<uint32_t>() : nullptr, tl, nullptr, s->j.as<uint32_t>(), st);
  if (rc) return rc;
  *d_jT = s->j.as<uint32_t>();
  *d_status = s->status.as<int32_t>();
  return 0;
}

}  // namespace qb200

extern "C" {

int qb200_exact_create(qb200_context* ctx, const qb200_params* params, int kind, uint32_t dimension_max,
                       uint32_t emax, qb200_exact** out) {
  if (out) *out = nullptr;
  if (!ctx || !params || !out) return set_error(-1, "null argument");
  std::unique_ptr<qb200_exact> s(new qb200_exact);
  std::string err;
  const int rc = exact_prepare(kind, params->m, params->l, params->sigma, params->d_be, params->d_len, params->r_be,
                               params->r_len, dimension_max, emax, &s->host, &err);
  if (rc) return set_error(rc, err);
  const CtxView cv = ctx_view(ctx);
  QD_CUDA(cudaSetDevice(cv.device));
  s->device = cv.device;
  s->sm_count = cv.sm_count;
  s->stream = cv.stream;
  s->launches = cv.launches;
  // inv_r, inv_d, d one after the other, each with its zero limbs: the kernel stages the same words
  // in shared memory (exact_const_words)
  const std::vector<uint32_t>* parts[3] = {&s->host.inv_r, &s->host.inv_d, &s->host.d};
  size_t total = 0;
  for (int i = 0; i < 3; i++) total += parts[i]->size();
  if (total != exact_const_words(s->host.c)) return set_error(-100, "exact sampler: constant layout");
  if (s->consts.reserve(total * 4) || s->table.reserve(s->host.table.size() * 4)) return -100;
  uint32_t* c = s->consts.as<uint32_t>();
  size_t at = 0;
  const uint32_t* dev_ptr[3];
  for (int i = 0; i < 3; i++) {
    QD_CUDA(cudaMemcpy(c + at, parts[i]->data(), parts[i]->size() * 4, cudaMemcpyHostToDevice));
    dev_ptr[i] = c + at + QB_EXACT_PAD;
    at += parts[i]->size();
  }
  QD_CUDA(cudaMemcpy(s->table.p, s->host.table.data(), s->host.table.size() * 4, cudaMemcpyHostToDevice));
  s->dev = s->host.c;
  s->dev.inv_r = dev_ptr[0];
  s->dev.inv_d = dev_ptr[1];
  s->dev.d = dev_ptr[2];
  s->dev.table = s->table.as<uint32_t>();
  // samples per launch: enough threads for every SM several times over, buffers of at most ~1 GB
  const size_t per = ((size_t)exact_jk_scratch_limbs(s->host.c) + 2 * s->host.c.wa + 2 * s->host.c.wn) * 4;
  size_t b = (size_t)cv.sm_count * 2048;
  while (b > 4096 && b * per > ((size_t)1 << 30)) b /= 2;
  s->chunk = (uint32_t)(b / QB_DIAGK_CTA * QB_DIAGK_CTA);
  for (int i = 0; i < 4; i++) QD_CUDA(cudaEventCreate(&s->ev[i]));
  {
    const char* v = getenv("QB200_EXACT_COLUMNS");
    if (v && *v == '4') s->columns = 4;
  }
  *out = s.release();
  return 0;
}

void qb200_exact_destroy(qb200_exact* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  delete s;
}

void qb200_exact_dims(const qb200_exact* s, uint32_t out[6]) {
  if (!s || !out) return;
  const ExactConst& c = s->host.c;
  out[0] = c.wa;
  out[1] = c.wn;
  out[2] = c.wk;
  out[3] = c.kappa_d;
  out[4] = c.kappa_r;
  out[5] = c.emax;
}

int qb200_exact_kernel_ms(qb200_exact* s, float out[2]) {
  if (!s || !out) return set_error(-1, "null argument");
  QD_CUDA(cudaSetDevice(s->device));
  for (int i = 0; i < 2; i++) {
    out[i] = 0.0f;
    if (!s->timed[i]) continue;
    QD_CUDA(cudaEventSynchronize(s->ev[2 * i + 1]));
    QD_CUDA(cudaEventElapsedTime(&out[i], s->ev[2 * i], s->ev[2 * i + 1]));
  }
  return 0;
}

int qb200_exact_region_bytes(const qb200_exact* s, int32_t min_log_alpha, uint32_t region, uint32_t dimension,
                             uint32_t* bytes) {
  if (!s || !bytes) return set_error(-1, "null argument");
  const ExactConst& c = s->host.c;
  std::vector<uint32_t> lo(c.wa), M(c.wa + 1);
  ExactRegion g;
  g.min_log_alpha = min_log_alpha;
  g.region = region;
  g.dimension = dimension;
  g.length = 0;
  g.offset = 0;
  int st = 0;
  const uint32_t bits = exact_region_modulus<1, 1>(c, g, lo.data(), M.data(), &st);
  *bytes = 0;
  if (st == QB_EXACT_UNSUPPORTED)
    return set_error(-50, "exact sampler: region outside the sampler's range (|min_log_alpha| < 8 or above emax, "
                          "dimension not a power of two up to the table's)");
  if (st != QB_EXACT_OK) return set_error(-51, "exact sampler: a bound within 2^-64 of a half-integer");
  *bytes = exact_bytes_for_bits(bits);
  return 0;
}

int qb200_exact_alpha(qb200_exact* s, uint32_t n, const qb200_exact_region* regions, uint32_t kappa,
                      const uint8_t* stream, uint64_t stream_len, uint32_t* alpha, int32_t* negative,
                      int32_t* status) {
  if (!s || !regions || !stream || !alpha || !negative || !status) return set_error(-1, "null argument");
  QD_CUDA(cudaSetDevice(s->device));
  const ExactConst& c = s->host.c;
  const uint8_t* d_stream = nullptr;
  int rc = exact_upload_stream(s, stream, stream_len, &d_stream, s->stream);
  if (rc) return rc;
  for (uint32_t done = 0; done < n;) {
    const uint32_t B = n - done < s->chunk ? n - done : s->chunk;
    rc = launch_alpha(s, B, regions + done, kappa, d_stream, stream_len, s->a_r, s->neg_r, s->stream);
    if (rc) return rc;
    QD_CUDA(cudaMemcpyAsync(negative + done, s->neg_r.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s->stream));
    QD_CUDA(cudaMemcpyAsync(status + done, s->status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s->stream));
    rc = tiles_to_rows(s, s->a_r, c.wa, B, alpha + (size_t)done * c.wa, s->stream);
    if (rc) return rc;
    done += B;
  }
  return 0;
}

int qb200_exact_j_k(qb200_exact* s, int mode, uint32_t n, const uint32_t* alpha_d, const int32_t* negative_d,
                    const uint32_t* alpha_r, const int32_t* negative_r, const uint32_t* t, uint32_t* j,
                    uint32_t* k) {
  if (!s || !j) return set_error(-1, "null argument");
  const ExactConst& c = s->host.c;
  const bool need_d = mode == QB_EXACT_J_K_FROM_ALPHA_D_R || mode == QB_EXACT_J_FROM_ALPHA_D_K;
  const bool need_r = mode == QB_EXACT_J_FROM_ALPHA_R || mode == QB_EXACT_J_K_FROM_ALPHA_D_R;
  if (!need_d && !need_r) return set_error(-2, "exact sampler: unknown mode");
  if ((need_d && (!alpha_d || !negative_d)) || (need_r && (!alpha_r || !negative_r)))
    return set_error(-1, "null argument");
  if (need_d && c.kbits == 0) return set_error(-2, "exact sampler: the diagonal sampler has no k (use qb200_diagk)");
  if (need_d && !k) return set_error(-1, "null argument");
  const uint32_t kap = mode == QB_EXACT_J_FROM_ALPHA_D_K ? c.kappa_d : c.kappa_r;
  if (kap && !t) return set_error(-1, "exact sampler: t is needed when the divisor is even");
  const uint32_t tl = kap ? (kap + 31) / 32 : 1;
  QD_CUDA(cudaSetDevice(s->device));
  for (uint32_t done = 0; done < n;) {
    const uint32_t B = n - done < s->chunk ? n - done : s->chunk;
    int rc = 0;
    if (need_d) {
      rc = rows_to_tiles(s, alpha_d + (size_t)done * c.wa, c.wa, B, s->a_d, s->stream);
      if (!rc) rc = host_ints_to_device(negative_d + done, B, s->neg_d, s->stream);
    }
    if (!rc && need_r) {
      rc = rows_to_tiles(s, alpha_r + (size_t)done * c.wa, c.wa, B, s->a_r, s->stream);
      if (!rc) rc = host_ints_to_device(negative_r + done, B, s->neg_r, s->stream);
    }
    if (!rc && kap) rc = rows_to_tiles(s, t + (size_t)done * tl, tl, B, s->t, s->stream);
    if (!rc && mode == QB_EXACT_J_FROM_ALPHA_D_K)
      rc = rows_to_tiles(s, k + (size_t)done * c.wk, c.wk, B, s->k, s->stream);
    if (rc) return rc;
    const size_t Bp = (size_t)tiles_of(B) * QB_DIAGK_CTA;
    if (s->j.reserve(Bp * c.wn * 4) || (need_d && s->k.reserve(Bp * c.wk * 4))) return -100;
    rc = launch_jk(s, mode, B, need_d ? s->a_d.as<uint32_t>() : nullptr, need_d ? s->neg_d.as<int32_t>() : nullptr,
                   need_r ? s->a_r.as<uint32_t>() : nullptr, need_r ? s->neg_r.as<int32_t>() : nullptr,
                   kap ? s->t.as<uint32_t>() : nullptr, tl, need_d ? s->k.as<uint32_t>() : nullptr,
                   s->j.as<uint32_t>(), s->stream);
    if (rc) return rc;
    rc = tiles_to_rows(s, s->j, c.wn, B, j + (size_t)done * c.wn, s->stream);
    if (!rc && mode == QB_EXACT_J_K_FROM_ALPHA_D_R)
      rc = tiles_to_rows(s, s->k, c.wk, B, k + (size_t)done * c.wk, s->stream);
    if (rc) return rc;
    done += B;
  }
  return 0;
}

}  // extern "C"
