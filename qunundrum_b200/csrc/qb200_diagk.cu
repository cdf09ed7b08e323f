// qb200_diagk.cu -- C ABI of the diagonal k sampler (include/qunundrum_b200.h, "diagonal (j, k)").
//
// Replaces sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412-646), the second half of
// diagonal_distribution_sample_pair_j_k (src/diagonal_distribution.cpp:474-552), and the sum of
// tau_estimate_diagonal (src/tau_estimate.cpp:135-210), by batched kernels. There is no CPU
// path: the entry points fail without a CUDA device like the rest of the library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "devbuf.hpp"
#include "diagk_host.hpp"
#include "exact_api.hpp"
#include "kernels_diagk.cuh"

using namespace qb200;

struct qb200_diagk {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t* launches = nullptr;
  DiagKHost host;
  DiagKConst dev;  // pointers into `consts`
  DBuf consts, rows_j, cols_j, eta, pivot, scratch, cols_k, rows_k, out, sums, status, xh, xl, hout;
  uint32_t chunk = 0;
};

// Samples per launch: enough threads for every SM, scratch of at most ~1 GB.
static uint32_t pick_chunk(const qb200_diagk* s, int sm_count) {
  const size_t per = ((size_t)diagk_scratch_limbs(s->host.c.k) + s->host.c.wj + s->host.c.wl) * 4;
  size_t b = (size_t)sm_count * 2048;
  while (b > 4096 && b * per > ((size_t)1 << 30)) b /= 2;
  return (uint32_t)b;
}

extern "C" {

int qb200_diagk_create(qb200_context* ctx, const qb200_params* params, qb200_diagk** out) {
  if (out) *out = nullptr;
  if (!ctx || !params || !out) return set_error(-1, "null argument");
  std::unique_ptr<qb200_diagk> s(new qb200_diagk);
  std::string err;
  const int rc = diagk_prepare(params->m, params->sigma, params->l, params->d_be, params->d_len,
                               params->r_be, params->r_len, &s->host, &err);
  if (rc) return set_error(rc, err);
  if (params->m + params->sigma > 16384)
    return set_error(-4, "diagonal k sampler: m + sigma above 16384 bits is not supported");
  const CtxView cv = ctx_view(ctx);
  QD_CUDA(cudaSetDevice(cv.device));
  s->device = cv.device;
  s->stream = cv.stream;
  s->launches = cv.launches;
  const uint32_t k = s->host.c.k;
  // r, d, mu, rho, dq, psi one after the other, each with its zero limbs (diagk_host.hpp); the kernel
  // stages the same words in shared memory (diagk_const_words)
  const std::vector<uint32_t>* parts[6] = {&s->host.r,   &s->host.d,  &s->host.mu,
                                           &s->host.rho, &s->host.dq, &s->host.psi};
  size_t total_words = 0;
  for (int i = 0; i < 6; i++) total_words += parts[i]->size();
  if (total_words != diagk_const_words(s->host.c)) return set_error(-100, "diagonal k sampler: constant layout");
  if (s->consts.reserve(total_words * 4)) return -100;
  uint32_t* c = s->consts.as<uint32_t>();
  size_t at = 0;
  const uint32_t* dev_ptr[6];
  for (int i = 0; i < 6; i++) {
    QD_CUDA(cudaMemcpy(c + at, parts[i]->data(), parts[i]->size() * 4, cudaMemcpyHostToDevice));
    dev_ptr[i] = c + at + QB_DIAGK_PAD;
    at += parts[i]->size();
  }
  s->dev = s->host.c;
  s->dev.r = dev_ptr[0];
  s->dev.d = dev_ptr[1];
  s->dev.mu = dev_ptr[2];
  s->dev.rho = dev_ptr[3];
  s->dev.dq = dev_ptr[4];
  s->dev.psi = dev_ptr[5];
  s->chunk = pick_chunk(s.get(), cv.sm_count);
  *out = s.release();
  return 0;
}

void qb200_diagk_destroy(qb200_diagk* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  delete s;
}

int qb200_diagk_set_force_exact(qb200_diagk* s, int on) {
  if (!s) return set_error(-1, "null argument");
  s->host.c.force_exact = s->dev.force_exact = on ? 1 : 0;
  return 0;
}

uint32_t qb200_diagk_j_limbs(const qb200_diagk* s) { return s ? s->host.c.wj : 0; }
uint32_t qb200_diagk_k_limbs(const qb200_diagk* s) { return s ? s->host.c.wl : 0; }

// The launches of one chunk of B samples whose rows already lie in device memory.
// d_j_tiles != NULL: j already lies in tiles (k_exact_jk wrote it) and d_j is not read.
static int launch_chunk(qb200_diagk* s, uint32_t B, const uint32_t* d_j, const int32_t* d_eta,
                        const RawX87* d_pivot, uint32_t delta_bound, uint32_t* d_k_rows, DiagKOut* d_out,
                        cudaStream_t stream, const uint32_t* d_j_tiles = nullptr) {
  const DiagKConst& c = s->host.c;
  const size_t scr = diagk_scratch_limbs(c.k);
  const size_t Bp = ((size_t)B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA * QB_DIAGK_CTA;  // whole tiles
  const size_t shmem = (size_t)diagk_const_words(c) * 4;
  // one CTA (and one scratch area) per tile: a persistent wave with per-CTA scratch was measured
  // and is slower (kernels_diagk.cuh)
  const uint32_t grid = (uint32_t)(Bp / QB_DIAGK_CTA);
  if ((!d_j_tiles && s->cols_j.reserve(Bp * c.wj * 4)) || s->scratch.reserve((size_t)grid * QB_DIAGK_CTA * scr * 4))
    return -100;
  if (d_k_rows && s->cols_k.reserve(Bp * c.wl * 4)) return -100;
  if (!d_j_tiles) {
    const uint64_t nj = (uint64_t)Bp * c.wj;
    k_diagk_gather<<<(unsigned)((nj + 255) / 256), 256, 0, stream>>>(d_j, c.wj, B, s->cols_j.as<uint32_t>());
    *s->launches += 1;
  }
  k_diagk<<<grid, QB_DIAGK_CTA, shmem, stream>>>(s->dev, d_j_tiles ? d_j_tiles : s->cols_j.as<uint32_t>(), d_eta,
                                                  d_pivot, (unsigned long long)delta_bound, B,
                                                  s->scratch.as<uint32_t>(),
                                                  d_k_rows ? s->cols_k.as<uint32_t>() : nullptr, d_out);
  *s->launches += 1;
  if (d_k_rows) {
    const uint64_t nk = (uint64_t)B * c.wl;
    k_diagk_scatter<<<(unsigned)((nk + 255) / 256), 256, 0, stream>>>(s->cols_k.as<uint32_t>(), c.wl, B,
                                                                     d_k_rows);
    *s->launches += 1;
  }
  QD_CUDA(cudaGetLastError());
  return 0;
}

// One chunk of B samples from host rows; results stay on the device in s->out / s->rows_k.
static int run_chunk(qb200_diagk* s, uint32_t B, const uint32_t* j, const int32_t* eta,
                     const long double* pivot, uint32_t delta_bound, bool want_k) {
  const DiagKConst& c = s->host.c;
  if (s->rows_j.reserve((size_t)B * c.wj * 4) || s->eta.reserve((size_t)B * 4) ||
      s->pivot.reserve((size_t)B * 16) || s->out.reserve((size_t)B * sizeof(DiagKOut)))
    return -100;
  if (want_k && s->rows_k.reserve((size_t)B * c.wl * 4)) return -100;
  QD_CUDA(cudaMemcpyAsync(s->rows_j.p, j, (size_t)B * c.wj * 4, cudaMemcpyHostToDevice, s->stream));
  QD_CUDA(cudaMemcpyAsync(s->eta.p, eta, (size_t)B * 4, cudaMemcpyHostToDevice, s->stream));
  QD_CUDA(cudaMemcpyAsync(s->pivot.p, pivot, (size_t)B * 16, cudaMemcpyHostToDevice, s->stream));
  return launch_chunk(s, B, s->rows_j.as<uint32_t>(), s->eta.as<int32_t>(), s->pivot.as<RawX87>(), delta_bound,
                      want_k ? s->rows_k.as<uint32_t>() : nullptr, s->out.as<DiagKOut>(), s->stream);
}

int qb200_diagk_sample_device(qb200_diagk* s, uint32_t n, const uint32_t* d_j, const int32_t* d_eta,
                              const long double* d_pivot, uint32_t delta_bound, uint32_t* d_k,
                              void* d_out, void* stream) {
  if (!s || !d_j || !d_eta || !d_pivot || !d_out) return set_error(-1, "null argument");
  QD_CUDA(cudaSetDevice(s->device));
  return launch_chunk(s, n, d_j, d_eta, (const RawX87*)d_pivot, delta_bound, d_k, (DiagKOut*)d_out,
                      stream ? (cudaStream_t)stream : s->stream);
}

int qb200_diagk_sample(qb200_diagk* s, uint32_t n, const uint32_t* j, const int32_t* eta,
                       const long double* pivot, uint32_t delta_bound, uint32_t* k, double* x_hi,
                       double* x_lo, int64_t* delta, int32_t* status) {
  if (!s || !j || !eta || !pivot) return set_error(-1, "null argument");
  QD_CUDA(cudaSetDevice(s->device));
  const DiagKConst& c = s->host.c;
  std::vector<DiagKOut> ho;
  for (uint32_t done = 0; done < n;) {
    const uint32_t B = n - done < s->chunk ? n - done : s->chunk;
    const int rc = run_chunk(s, B, j + (size_t)done * c.wj, eta + done, pivot + done, delta_bound, k != nullptr);
    if (rc) return rc;
    ho.resize(B);
    QD_CUDA(cudaMemcpyAsync(ho.data(), s->out.p, (size_t)B * sizeof(DiagKOut), cudaMemcpyDeviceToHost, s->stream));
    if (k)
      QD_CUDA(cudaMemcpyAsync(k + (size_t)done * c.wl, s->rows_k.p, (size_t)B * c.wl * 4, cudaMemcpyDeviceToHost,
                              s->stream));
    QD_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t i = 0; i < B; i++) {
      if (ho[i].status < 0) return set_error(-42, "The pivot is out of bounds.");
      if (x_hi) x_hi[done + i] = ho[i].x_hi;
      if (x_lo) x_lo[done + i] = ho[i].x_lo;
      if (delta) delta[done + i] = ho[i].delta;
      if (status) status[done + i] = ho[i].status;
    }
    done += B;
  }
  return 0;
}

int qb200_diagk_sample_drawn(qb200_diagk* s, qb200_exact* ex, uint32_t n, const qb200_exact_region* regions,
                             const uint32_t* t_r, const uint8_t* stream, uint64_t stream_len, const int32_t* eta,
                             const long double* pivot, uint32_t delta_bound, uint32_t* k, double* x_hi,
                             double* x_lo, int64_t* delta, int32_t* status, int32_t* exact_status) {
  if (!s || !ex || !regions || !stream || !eta || !pivot || !exact_status) return set_error(-1, "null argument");
  const DiagKConst& c = s->host.c;
  if (exact_j_limbs(ex) != c.wj || exact_device(ex) != s->device)
    return set_error(-2, "qb200_diagk_sample_drawn: the exact sampler is not a diagonal one of the same parameters "
                         "on the same device");
  QD_CUDA(cudaSetDevice(s->device));
  uint32_t dims[6];
  qb200_exact_dims(ex, dims);
  const uint32_t tl = dims[4] ? (dims[4] + 31) / 32 : 1;
  const uint8_t* d_stream = nullptr;
  int rc = exact_upload_stream(ex, stream, stream_len, &d_stream, s->stream);
  if (rc) return rc;
  const uint32_t chunk = s->chunk < exact_chunk(ex) ? s->chunk : exact_chunk(ex);
  std::vector<DiagKOut> ho;
  for (uint32_t done = 0; done < n;) {
    const uint32_t B = n - done < chunk ? n - done : chunk;
    const uint32_t* d_jT = nullptr;
    const int32_t* d_st = nullptr;
    rc = exact_draw_j_tiles(ex, B, regions + done, t_r ? t_r + (size_t)done * tl : nullptr, d_stream, stream_len,
                            &d_jT, &d_st, s->stream);
    if (rc) return rc;
    if (s->eta.reserve((size_t)B * 4) || s->pivot.reserve((size_t)B * 16) ||
        s->out.reserve((size_t)B * sizeof(DiagKOut)))
      return -100;
    if (k && s->rows_k.reserve((size_t)B * c.wl * 4)) return -100;
    QD_CUDA(cudaMemcpyAsync(s->eta.p, eta + done, (size_t)B * 4, cudaMemcpyHostToDevice, s->stream));
    QD_CUDA(cudaMemcpyAsync(s->pivot.p, pivot + done, (size_t)B * 16, cudaMemcpyHostToDevice, s->stream));
    rc = launch_chunk(s, B, nullptr, s->eta.as<int32_t>(), s->pivot.as<RawX87>(), delta_bound,
                      k ? s->rows_k.as<uint32_t>() : nullptr, s->out.as<DiagKOut>(), s->stream, d_jT);
    if (rc) return rc;
    ho.resize(B);
    QD_CUDA(cudaMemcpyAsync(ho.data(), s->out.p, (size_t)B * sizeof(DiagKOut), cudaMemcpyDeviceToHost, s->stream));
    QD_CUDA(cudaMemcpyAsync(exact_status + done, d_st, (size_t)B * 4, cudaMemcpyDeviceToHost, s->stream));
    if (k)
      QD_CUDA(cudaMemcpyAsync(k + (size_t)done * c.wl, s->rows_k.p, (size_t)B * c.wl * 4, cudaMemcpyDeviceToHost,
                              s->stream));
    QD_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t i = 0; i < B; i++) {
      if (ho[i].status < 0) return set_error(-42, "The pivot is out of bounds.");
      if (x_hi) x_hi[done + i] = ho[i].x_hi;
      if (x_lo) x_lo[done + i] = ho[i].x_lo;
      if (delta) delta[done + i] = ho[i].delta;
      if (status) status[done + i] = ho[i].status;
    }
    done += B;
  }
  return 0;
}

int qb200_diagk_tau_estimate(qb200_diagk* s, uint32_t n, uint32_t count, const uint32_t* j,
                             const int32_t* eta, const long double* pivot, uint32_t delta_bound,
                             uint32_t eta_bound, long double* tau, uint8_t* ok) {
  if (!s || !j || !eta || !pivot || !tau || !ok) return set_error(-1, "null argument");
  if (n == 0) return set_error(-2, "n must be positive");
  QD_CUDA(cudaSetDevice(s->device));
  const DiagKConst& c = s->host.c;
  uint32_t per = s->chunk / n;  // whole estimates per launch
  if (per == 0) per = 1;
  std::vector<double> hs;
  std::vector<int> hst;
  for (uint32_t done = 0; done < count;) {
    const uint32_t E = count - done < per ? count - done : per;
    const uint32_t B = E * n;
    const size_t first = (size_t)done * n;
    const int rc = run_chunk(s, B, j + first * c.wj, eta + first, pivot + first, delta_bound, false);
    if (rc) return rc;
    if (s->sums.reserve((size_t)E * 16) || s->status.reserve((size_t)E * 4)) return -100;
    k_diagk_tau<<<(E + 127) / 128, 128, 0, s->stream>>>(s->out.as<DiagKOut>(), s->eta.as<int32_t>(), c.l, n, E,
                                                        eta_bound, s->sums.as<double>(), s->status.as<int>());
    *s->launches += 1;
    QD_CUDA(cudaGetLastError());
    hs.resize((size_t)E * 2);
    hst.resize(E);
    QD_CUDA(cudaMemcpyAsync(hs.data(), s->sums.p, (size_t)E * 16, cudaMemcpyDeviceToHost, s->stream));
    QD_CUDA(cudaMemcpyAsync(hst.data(), s->status.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
    QD_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t t = 0; t < E; t++) {
      if (hst[t] < 0) return set_error(-42, "The pivot is out of bounds.");
      if (hst[t] == QB_DIAGK_OK_NEGATIVE_PHI || hst[t] == QB_DIAGK_GAVE_UP)
        return set_error(-43, hst[t] == QB_DIAGK_GAVE_UP
                                  ? "diagonal k sampler: gave up after 2^22 steps"
                                  : "diagonal k sampler: j below |eta| 2^(m + sigma) / r with l > 1000");
      if (hst[t] != 0) {
        // src/tau_estimate.cpp:190-193
        ok[done + t] = 0;
        tau[done + t] = DBL_MAX;
        continue;
      }
      // log2(sum alpha_phi^2 / n) / 2 - (m + sigma - l) with alpha_phi = 2^(m+sigma-l) x (:194-201)
      const long double a = ((long double)hs[2 * t] + (long double)hs[2 * t + 1]) / (long double)n;
      ok[done + t] = 1;
      tau[done + t] = log2l(a) / 2;
    }
    done += E;
  }
  return 0;
}

int qb200_diagk_h(qb200_diagk* s, uint32_t n, const double* x_hi, const double* x_lo, long double* h) {
  if (!s || !x_hi || !x_lo || !h) return set_error(-1, "null argument");
  QD_CUDA(cudaSetDevice(s->device));
  if (n == 0) return 0;
  if (s->xh.reserve((size_t)n * 8) || s->xl.reserve((size_t)n * 8) || s->hout.reserve((size_t)n * 16)) return -100;
  QD_CUDA(cudaMemcpyAsync(s->xh.p, x_hi, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  QD_CUDA(cudaMemcpyAsync(s->xl.p, x_lo, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  k_diagk_h<<<(n + 127) / 128, 128, 0, s->stream>>>(s->host.c.l, n, s->xh.as<double>(), s->xl.as<double>(),
                                                     s->hout.as<RawX87>());
  *s->launches += 1;
  QD_CUDA(cudaGetLastError());
  QD_CUDA(cudaMemcpyAsync(h, s->hout.p, (size_t)n * 16, cudaMemcpyDeviceToHost, s->stream));
  QD_CUDA(cudaStreamSynchronize(s->stream));
  return 0;
}

}  // extern "C"
