// qb200.cu -- the C ABI (include/qunundrum_b200.h) over the CUDA kernels.
//
// One qb200_context per worker process / GPU. A plan owns the device copies of
// a batch's descriptors and scratch; qb200_plan_run() only enqueues kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qunundrum_b200.h"
#include "ctx_access.hpp"
#include "kernels_client.cuh"
#include "kernels_fused1d.cuh"
#include "kernels_fused2d.cuh"
#include "kernels_plain.cuh"
#include "kernels_sigma_opt.cuh"
#include "plan.hpp"

using namespace qb200;

namespace {

thread_local std::string g_err = "";

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define QB_CUDA(call)                                                                   \
  do {                                                                                  \
    const cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess)                                                              \
      return fail(-100, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

// Device buffers come from a per-context pool: a plan returns its buffers to
// the pool when it is destroyed, so the synchronous API (one plan per call)
// does not pay cudaMalloc / cudaFree on every call.
struct Pool {
  std::vector<std::pair<void*, size_t>> free_;
  ~Pool() {
    for (auto& e : free_) cudaFree(e.first);
  }
  bool take(size_t n, void** p, size_t* got) {
    size_t best = free_.size();
    for (size_t i = 0; i < free_.size(); i++)
      if (free_[i].second >= n && free_[i].second <= 4 * n + 4096 &&
          (best == free_.size() || free_[i].second < free_[best].second))
        best = i;
    if (best == free_.size()) return false;
    *p = free_[best].first;
    *got = free_[best].second;
    free_.erase(free_.begin() + (long)best);
    return true;
  }
  void give(void* p, size_t n) { free_.emplace_back(p, n); }
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  Pool* pool = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (!p) return;
    if (pool)
      pool->give(p, bytes);
    else
      cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int reserve(size_t n) {
    if (n <= bytes) return 0;
    release();
    if (pool && pool->take(n, &p, &bytes)) return 0;
    QB_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return 0;
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

struct DevGeometry {
  DevBuf gx, gw;
};

}  // namespace

struct qb200_context {
  int device = 0;
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_stream = nullptr;  // device-to-host copies of the synchronous API
  // The three slice classes of the fused kernel are independent launches: the smaller two run on
  // side streams next to class 0, so that the partial last waves of the three overlap instead of
  // following each other (what is left of a GPU's share when one distribution is split over 8).
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
  bool overlap_classes = true;         // QB200_OVERLAP_CLASSES=0: one stream (A/B)
  // QB200_FUSED_LEAN=1: one prologue launch (axis tables + column records) and the slice summaries
  // written by the class kernels themselves (tickets) -- two dependent launches less per step than
  // k_axis2d -> k_fused_cols -> k_fused2d x 3 -> k_fused_final. Measured (B200, an eighth of the
  // bench distribution per step): 2.6 % faster when every step is enqueued eagerly, no gain once the
  // step is a CUDA graph (0.1538 vs 0.1537 ms), and 0.8 % slower on the whole distribution (the
  // ticket atomics and the closing warp sit in the hot kernel): off by default.
  bool fused_lean = false;
  bool use_graphs = true;
  std::vector<cudaEvent_t> events;
  int sm_count = 0;
  uint64_t launches = 0;
  std::map<int, std::unique_ptr<DevGeometry>> geo;
  // staging for the synchronous host API
  DevBuf out_cells, out_summary, out_scaled, out_status;
  void* h_summary = nullptr;
  size_t h_summary_bytes = 0;
  qb200::TextState* text = nullptr;  // text exporter / importer state (qb200_text.cu)
  Pool pool;  // declared last: destroyed first is fine, plans never outlive their context
};

struct qb200_plan {
  qb200_context* ctx = nullptr;
  Plan host;
  int algo = 0;        // resolved: 1 plain, 2 fused
  int fused_ok = 0;
  std::string fused_why;
  uint32_t n = 0;
  std::vector<DevSlice> h_slices;
  DevGeometry* geo = nullptr;
  DevBuf desc_a, desc_b, slices, tab_a, tab_b;
  // plain path scratch (per chunk of slices)
  uint32_t chunk = 0;
  bool plain_ready = false;
  DevBuf cells_c, cells_f, part_c, part_f, part_tp, values;
  // fused path
  FusedPlan2D fused;
  DevBuf fused_part, fused_cols, fused_slices, fused_tickets;
  // A plan that is run again with the same output buffers replays its step as ONE CUDA graph: at an
  // eighth of a distribution per GPU the step is 125 us of kernels and the host needs longer than
  // the first kernels run to enqueue the next ones (launch-bound). QB200_GRAPHS=0: always eager.
  cudaGraphExec_t graph_exec = nullptr;
  double* graph_cells = nullptr;
  double* graph_summary = nullptr;
  double* last_cells = nullptr;      // buffers of the previous eager run
  double* last_summary = nullptr;
  uint32_t graph_kernels = 0;
  bool graph_failed = false;
  // fused one-dimensional path: per-block partials and per-slice tickets
  DevBuf f1d_part, f1d_tickets;
  bool f1d_ready = false;
  // sigma-optimal, large l: the quick method's plan over the same slices (cells, mass and axis
  // tables for k_so_fast; see sigma_opt.cuh "the walk in closed form")
  qb200_plan* so_quick = nullptr;
  // sigma-optimal scratch
  DevBuf so_sigma, so_guess, so_norm, so_erra, so_sigma0, so_status, so_changed;
  void bind_pool(Pool* pool) {
    DevBuf* all[] = {&desc_a, &desc_b, &slices, &tab_a, &tab_b, &cells_c, &cells_f, &part_c,
                     &part_f, &part_tp, &values, &fused_part, &fused_cols, &fused_slices,
                     &so_sigma, &so_guess, &so_norm, &so_erra, &so_sigma0, &so_status,
                     &so_changed, &f1d_part, &f1d_tickets, &fused_tickets};
    for (DevBuf* b : all) b->pool = pool;
  }
};

namespace qb200 {
CtxView ctx_view(qb200_context* ctx) {
  return CtxView{ctx->device, ctx->sm_count, ctx->stream, &ctx->launches, &ctx->text};
}
int set_error(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace qb200

namespace {

int get_geometry(qb200_context* ctx, int D, DevGeometry** out) {
  auto it = ctx->geo.find(D);
  if (it != ctx->geo.end()) {
    *out = it->second.get();
    return 0;
  }
  const Geometry g = make_geometry(D);
  std::unique_ptr<DevGeometry> dg(new DevGeometry);
  if (int rc = dg->gx.reserve(g.gx.size() * sizeof(DD))) return rc;
  if (int rc = dg->gw.reserve(g.gw.size() * sizeof(double))) return rc;
  QB_CUDA(cudaMemcpy(dg->gx.p, g.gx.data(), g.gx.size() * sizeof(DD), cudaMemcpyHostToDevice));
  QB_CUDA(cudaMemcpy(dg->gw.p, g.gw.data(), g.gw.size() * sizeof(double),
                     cudaMemcpyHostToDevice));
  *out = dg.get();
  ctx->geo[D] = std::move(dg);
  return 0;
}

ParamsView view_of(const qb200_params* p) {
  ParamsView v;
  v.m = p->m;
  v.l = p->l;
  v.sigma = p->sigma;
  v.d_be = p->d_be;
  v.d_len = p->d_len;
  v.r_be = p->r_be;
  v.r_len = p->r_len;
  return v;
}

int upload_plan(qb200_plan* pl) {
  const Plan& h = pl->host;
  const uint32_t n = (uint32_t)h.slices.size();
  pl->n = n;
  const int D = h.D;
  const int NP = table_points(D);
  if (int rc = get_geometry(pl->ctx, D, &pl->geo)) return rc;
  std::vector<DevSlice>& ds = pl->h_slices;
  ds.resize(n);
  for (uint32_t i = 0; i < n; i++) {
    ds[i].tab_a = h.slices[i].tab_a;
    ds[i].tab_b = h.slices[i].tab_b;
    ds[i].scale_a = h.slices[i].scale_a;
    ds[i].scale_b = h.slices[i].scale_b;
    ds[i].eta_shift = h.slices[i].eta_shift;
  }
  if (int rc = pl->slices.reserve(std::max<size_t>(1, n) * sizeof(DevSlice))) return rc;
  if (n)
    QB_CUDA(cudaMemcpyAsync(pl->slices.p, ds.data(), n * sizeof(DevSlice), cudaMemcpyHostToDevice,
                            pl->ctx->stream));
  if (int rc = pl->desc_a.reserve(std::max<size_t>(1, h.tabs_a.size()) * sizeof(TabDesc))) return rc;
  if (!h.tabs_a.empty())
    QB_CUDA(cudaMemcpyAsync(pl->desc_a.p, h.tabs_a.data(), h.tabs_a.size() * sizeof(TabDesc),
                            cudaMemcpyHostToDevice, pl->ctx->stream));
  if (h.kind < 0) {
    if (int rc = pl->desc_b.reserve(std::max<size_t>(1, h.tabs_b.size()) * sizeof(TabDesc))) return rc;
    if (!h.tabs_b.empty())
      QB_CUDA(cudaMemcpyAsync(pl->desc_b.p, h.tabs_b.data(), h.tabs_b.size() * sizeof(TabDesc),
                              cudaMemcpyHostToDevice, pl->ctx->stream));
    if (int rc = pl->tab_a.reserve(std::max<size_t>(1, h.tabs_a.size()) * NP * sizeof(AxisD))) return rc;
    if (int rc = pl->tab_b.reserve(std::max<size_t>(1, h.tabs_b.size()) * NP * sizeof(AxisR))) return rc;
  }
  return 0;
}

// Scratch for the plain path: chunks of at most ~1 GiB of pass cells.
int reserve_plain(qb200_plan* pl) {
  const Plan& h = pl->host;
  const size_t D = (size_t)h.D;
  if (h.kind < 0) {
    const size_t per_slice = 5 * D * D * sizeof(double);
    size_t chunk = std::max<size_t>(1, (size_t(1) << 30) / per_slice);
    chunk = std::min<size_t>(chunk, std::max<uint32_t>(1, pl->n));
    chunk = std::min<size_t>(chunk, 65535);
    pl->chunk = (uint32_t)chunk;
    const size_t nb_c = (D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
    const size_t nb_f = (4 * D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
    if (int rc = pl->cells_c.reserve(chunk * D * D * sizeof(double))) return rc;
    if (h.richardson)
      if (int rc = pl->cells_f.reserve(chunk * 4 * D * D * sizeof(double))) return rc;
    if (int rc = pl->part_c.reserve(chunk * nb_c * 3 * sizeof(double))) return rc;
    if (int rc = pl->part_f.reserve(chunk * nb_f * 3 * sizeof(double))) return rc;
    if (int rc = pl->part_tp.reserve(chunk * nb_c * 2 * sizeof(double))) return rc;
  } else {
    const size_t NP = (size_t)table_points(h.D);
    size_t chunk = std::max<size_t>(1, (size_t(1) << 28) / (NP * sizeof(double)));
    chunk = std::min<size_t>(chunk, std::max<uint32_t>(1, pl->n));
    chunk = std::min<size_t>(chunk, 65535);
    pl->chunk = (uint32_t)chunk;
    const size_t nb = (D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
    if (int rc = pl->values.reserve(chunk * NP * sizeof(double))) return rc;
    if (int rc = pl->part_tp.reserve(chunk * nb * 2 * sizeof(double))) return rc;
  }
  return 0;
}

// The sigma-optimal method: fixed-point iteration of the reference's serial walk (see
// kernels_sigma_opt.cuh). Synchronises `st` between iterations (convergence flag).
int run_sigma_opt_2d(qb200_plan* pl, cudaStream_t st, double* d_cells, double* d_summary) {
  const Plan& h = pl->host;
  qb200_context* ctx = pl->ctx;
  if (pl->n == 0) return 0;
  const int D = h.D;
  SoLayout L;
  L.D = D;
  L.passes = h.richardson ? 2 : 1;
  L.n_c = (2 * D + 1) * (2 * D + 1);
  L.n_f = h.richardson ? (4 * D + 1) * (4 * D + 1) : 0;
  L.stride = L.n_c + L.n_f;
  if (pl->so_quick) {
    // Large l: cells and mass are the quick method's (fused kernel); the walk is a prefix minimum
    // in closed form (k_so_fast). A slice that leaves the proven range sends the batch through the
    // general iteration below.
    qb200_plan* q = pl->so_quick;
    if (int rc = pl->so_changed.reserve(sizeof(int))) return rc;
    int* d_fallback = pl->so_changed.as<int>();
    QB_CUDA(cudaMemsetAsync(d_fallback, 0, sizeof(int), st));
    if (int rc = qb200_plan_run(q, st, d_cells, d_summary)) return rc;
    const size_t wgt_bytes = (size_t)(4 * L.D + 1) * sizeof(double);  // the fine pass's abscissae
    k_so_fast<<<pl->n, QB_SOF_BLOCK, wgt_bytes, st>>>(q->host.c, L, q->slices.as<DevSlice>(), q->tab_a.as<AxisD>(),
                                             q->tab_b.as<AxisR>(), q->geo->gw.as<double>(), d_summary,
                                             d_fallback);
    ctx->launches++;
    QB_CUDA(cudaGetLastError());
    int fallback = 0;
    QB_CUDA(cudaMemcpyAsync(&fallback, d_fallback, sizeof(int), cudaMemcpyDeviceToHost, st));
    QB_CUDA(cudaStreamSynchronize(st));
    if (!fallback) return 0;
  }
  const size_t per_slice = (size_t)L.stride * 24 + (size_t)5 * D * D * 8;
  uint32_t chunk = (uint32_t)std::max<size_t>(1, (size_t(1) << 29) / per_slice);
  chunk = std::min<uint32_t>(std::min<uint32_t>(chunk, pl->n), 16384);
  const int nb_c = (D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
  const int nb_f = (4 * D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
  const size_t pts = (size_t)chunk * L.stride;
  if (int rc = pl->so_sigma.reserve(pts * sizeof(int))) return rc;
  if (int rc = pl->so_guess.reserve(pts * sizeof(int))) return rc;
  if (int rc = pl->so_norm.reserve(pts * sizeof(double))) return rc;
  if (int rc = pl->so_erra.reserve(pts * sizeof(double))) return rc;
  if (int rc = pl->so_sigma0.reserve((size_t)chunk * 2 * sizeof(int))) return rc;
  if (int rc = pl->so_status.reserve((size_t)chunk * 2 * sizeof(int))) return rc;
  if (int rc = pl->so_changed.reserve(sizeof(int))) return rc;
  if (int rc = pl->cells_c.reserve((size_t)chunk * D * D * sizeof(double))) return rc;
  if (h.richardson)
    if (int rc = pl->cells_f.reserve((size_t)chunk * 4 * D * D * sizeof(double))) return rc;
  if (int rc = pl->part_c.reserve((size_t)chunk * nb_c * 3 * sizeof(double))) return rc;
  if (int rc = pl->part_f.reserve((size_t)chunk * nb_f * 3 * sizeof(double))) return rc;
  if (int rc = pl->part_tp.reserve((size_t)chunk * nb_c * 2 * sizeof(double))) return rc;

  const int NP = table_points(D);
  const int n_a = (int)h.tabs_a.size(), n_b = (int)h.tabs_b.size();
  k_axis2d<<<dim3((NP + 127) / 128, n_a + n_b), 128, 0, st>>>(
      h.c, NP, n_a, pl->desc_a.as<TabDesc>(), pl->desc_b.as<TabDesc>(), pl->geo->gx.as<dd>(),
      pl->tab_a.as<AxisD>(), pl->tab_b.as<AxisR>());
  ctx->launches++;
  const TabDesc* da = pl->desc_a.as<TabDesc>();
  const TabDesc* db = pl->desc_b.as<TabDesc>();
  const dd* gx = pl->geo->gx.as<dd>();
  const AxisR* tb = pl->tab_b.as<AxisR>();
  int* sigma0 = pl->so_sigma0.as<int>();
  int* status = pl->so_status.as<int>();
  int* changed = pl->so_changed.as<int>();
  for (uint32_t s0 = 0; s0 < pl->n; s0 += chunk) {
    const uint32_t ns = std::min(chunk, pl->n - s0);
    const DevSlice* sl = pl->slices.as<DevSlice>() + s0;
    const unsigned pairs = ns * (unsigned)L.passes;
    QB_CUDA(cudaMemsetAsync(status, 0, pairs * sizeof(int), st));
    k_so_first<<<pairs, 256, 0, st>>>(h.c, h.so, L, sl, da, db, gx, tb, sigma0);
    const size_t npts = (size_t)ns * L.stride;
    k_so_fill<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(npts, L, sigma0, pl->so_guess.as<int>());
    ctx->launches += 2;
    const int max_pts = h.richardson ? L.n_f : L.n_c;
    int iter = 0, moved = 1;
    while (moved) {
      if (++iter > 4096) return fail(-30, "sigma-optimal walk did not converge");
      QB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
      k_so_step<<<dim3((max_pts + 127) / 128, pairs), 128, 0, st>>>(
          h.c, h.so, L, sl, da, db, gx, tb, sigma0, pl->so_guess.as<int>(), pl->so_sigma.as<int>(),
          pl->so_norm.as<double>(), pl->so_erra.as<double>(), status);
      k_so_scan<<<pairs, 1024, 0, st>>>(L, pl->so_sigma.as<int>(), pl->so_guess.as<int>(), changed);
      ctx->launches += 2;
      QB_CUDA(cudaMemcpyAsync(&moved, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
      QB_CUDA(cudaStreamSynchronize(st));
    }
    for (int pass = 0; pass < L.passes; pass++) {
      k_so_cells<<<dim3(pass ? nb_f : nb_c, ns), QB_PLAIN_BLOCK, 0, st>>>(
          h.c, L, pass, sl, pl->geo->gw.as<double>(), sigma0, pl->so_sigma.as<int>(),
          pl->so_norm.as<double>(), pl->so_erra.as<double>(),
          pass ? pl->cells_f.as<double>() : pl->cells_c.as<double>(),
          pass ? pl->part_f.as<double>() : pl->part_c.as<double>());
      ctx->launches++;
    }
    k_rich2d<<<dim3(nb_c, ns), QB_PLAIN_BLOCK, 0, st>>>(
        D, h.richardson, pl->cells_c.as<double>(), pl->cells_f.as<double>(),
        d_cells + (size_t)s0 * D * D, pl->part_tp.as<double>());
    k_so_final<<<(ns + 127) / 128, 128, 0, st>>>(
        (int)ns, L, nb_c, nb_f, nb_c, pl->part_c.as<double>(), pl->part_f.as<double>(),
        pl->part_tp.as<double>(), sigma0, status, d_summary + (size_t)s0 * QB200_SUMMARY_STRIDE);
    ctx->launches += 2;
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}

int run_plain_2d(qb200_plan* pl, cudaStream_t st, double* d_cells, double* d_summary) {
  if (pl->host.method == kMethodOptimalLocalSigma) return run_sigma_opt_2d(pl, st, d_cells, d_summary);
  if (!pl->plain_ready) {
    if (int rc = reserve_plain(pl)) return rc;
    pl->plain_ready = true;
  }
  const Plan& h = pl->host;
  qb200_context* ctx = pl->ctx;
  const int D = h.D;
  const int NP = table_points(D);
  const int n_a = (int)h.tabs_a.size(), n_b = (int)h.tabs_b.size();
  if (pl->n == 0) return 0;
  {
    dim3 grid((NP + 127) / 128, n_a + n_b);
    k_axis2d<<<grid, 128, 0, st>>>(h.c, NP, n_a, pl->desc_a.as<TabDesc>(), pl->desc_b.as<TabDesc>(),
                                   pl->geo->gx.as<dd>(), pl->tab_a.as<AxisD>(),
                                   pl->tab_b.as<AxisR>());
    ctx->launches++;
  }
  const int nb_c = (D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
  const int nb_f = (4 * D * D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
  for (uint32_t s0 = 0; s0 < pl->n; s0 += pl->chunk) {
    const uint32_t ns = std::min(pl->chunk, pl->n - s0);
    const DevSlice* sl = pl->slices.as<DevSlice>() + s0;
    k_pass2d<<<dim3(nb_c, ns), QB_PLAIN_BLOCK, 0, st>>>(
        h.c, D, 0, h.with_error ? 1 : 0, sl, pl->tab_a.as<AxisD>(), pl->tab_b.as<AxisR>(),
        pl->geo->gw.as<double>(), pl->cells_c.as<double>(), pl->part_c.as<double>());
    ctx->launches++;
    if (h.richardson) {
      k_pass2d<<<dim3(nb_f, ns), QB_PLAIN_BLOCK, 0, st>>>(
          h.c, D, 1, h.with_error ? 1 : 0, sl, pl->tab_a.as<AxisD>(), pl->tab_b.as<AxisR>(),
          pl->geo->gw.as<double>(), pl->cells_f.as<double>(), pl->part_f.as<double>());
      ctx->launches++;
    }
    k_rich2d<<<dim3(nb_c, ns), QB_PLAIN_BLOCK, 0, st>>>(
        D, h.richardson, pl->cells_c.as<double>(), pl->cells_f.as<double>(),
        d_cells + (size_t)s0 * D * D, pl->part_tp.as<double>());
    ctx->launches++;
    k_final2d<<<(ns + 127) / 128, 128, 0, st>>>(
        (int)ns, h.richardson, nb_c, nb_f, nb_c, pl->part_c.as<double>(), pl->part_f.as<double>(),
        pl->part_tp.as<double>(), d_summary + (size_t)s0 * QB200_SUMMARY_STRIDE);
    ctx->launches++;
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}

int run_plain_1d(qb200_plan* pl, cudaStream_t st, double* d_cells, double* d_summary) {
  if (!pl->plain_ready) {
    if (int rc = reserve_plain(pl)) return rc;
    pl->plain_ready = true;
  }
  const Plan& h = pl->host;
  qb200_context* ctx = pl->ctx;
  const int D = h.D;
  const int NP = table_points(D);
  const int nb = (D + QB_PLAIN_BLOCK - 1) / QB_PLAIN_BLOCK;
  for (uint32_t s0 = 0; s0 < pl->n; s0 += pl->chunk) {
    const uint32_t ns = std::min(pl->chunk, pl->n - s0);
    const DevSlice* sl = pl->slices.as<DevSlice>() + s0;
    k_vals1d<<<dim3((NP + 127) / 128, ns), 128, 0, st>>>(
        h.c, h.kind, NP, sl, pl->desc_a.as<TabDesc>(), pl->geo->gx.as<dd>(),
        pl->values.as<double>());
    k_cells1d<<<dim3(nb, ns), QB_PLAIN_BLOCK, 0, st>>>(
        D, h.richardson, sl, pl->values.as<double>(), pl->geo->gw.as<double>(),
        d_cells + (size_t)s0 * D, pl->part_tp.as<double>());
    k_final1d<<<(ns + 127) / 128, 128, 0, st>>>(
        (int)ns, nb, pl->part_tp.as<double>(), d_summary + (size_t)s0 * QB200_SUMMARY_STRIDE);
    ctx->launches += 3;
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}

// All slices of the batch in one launch (kernels_fused1d.cuh); grid.y carries the slices, so
// batches beyond 65535 slices take one launch per 65535.
int run_fused_1d(qb200_plan* pl, cudaStream_t st, double* d_cells, double* d_summary) {
  const Plan& h = pl->host;
  qb200_context* ctx = pl->ctx;
  if (pl->n == 0) return 0;
  const int D = h.D;
  const unsigned nb = (unsigned)((D + QB_F1D_BLOCK - 1) / QB_F1D_BLOCK);
  if (!pl->f1d_ready) {
    if (int rc = pl->f1d_part.reserve((size_t)pl->n * nb * 2 * sizeof(double))) return rc;
    if (int rc = pl->f1d_tickets.reserve((size_t)pl->n * sizeof(unsigned int))) return rc;
    // zeroed once: the kernel leaves every ticket at zero again
    QB_CUDA(cudaMemsetAsync(pl->f1d_tickets.p, 0, (size_t)pl->n * sizeof(unsigned int), st));
    pl->f1d_ready = true;
  }
  for (uint32_t s0 = 0; s0 < pl->n; s0 += 65535) {
    const uint32_t ns = std::min<uint32_t>(65535, pl->n - s0);
    const dim3 grid(nb, ns);
    const DevSlice* sl = pl->slices.as<DevSlice>() + s0;
    double* oc = d_cells + (size_t)s0 * D;
    double* pp = pl->f1d_part.as<double>() + (size_t)s0 * nb * 2;
    unsigned int* tk = pl->f1d_tickets.as<unsigned int>() + s0;
    double* sm = d_summary + (size_t)s0 * QB200_SUMMARY_STRIDE;
    const TabDesc* da = pl->desc_a.as<TabDesc>();
    const dd* gx = pl->geo->gx.as<dd>();
    const double* gw = pl->geo->gw.as<double>();
    if (h.kind == KIND_LINEAR_D)
      k_fused1d<KIND_LINEAR_D><<<grid, QB_F1D_BLOCK, 0, st>>>(h.c, D, h.richardson, sl, da, gx, gw, oc, pp, tk, sm);
    else if (h.kind == KIND_LINEAR_R)
      k_fused1d<KIND_LINEAR_R><<<grid, QB_F1D_BLOCK, 0, st>>>(h.c, D, h.richardson, sl, da, gx, gw, oc, pp, tk, sm);
    else
      k_fused1d<KIND_DIAGONAL><<<grid, QB_F1D_BLOCK, 0, st>>>(h.c, D, h.richardson, sl, da, gx, gw, oc, pp, tk, sm);
    ctx->launches++;
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}

uint32_t plain_chunk(const qb200_plan* pl) {
  const Plan& h = pl->host;
  const size_t D = (size_t)h.D;
  const size_t per_slice = h.kind < 0 ? 5 * D * D * sizeof(double)
                                      : (size_t)table_points(h.D) * sizeof(double);
  const size_t budget = h.kind < 0 ? (size_t(1) << 30) : (size_t(1) << 28);
  size_t chunk = std::max<size_t>(1, budget / per_slice);
  chunk = std::min<size_t>(chunk, std::max<uint32_t>(1, pl->n));
  return (uint32_t)std::min<size_t>(chunk, 65535);
}

uint32_t plain_launches(const qb200_plan* pl) {
  const uint32_t ch = plain_chunk(pl);
  const uint32_t chunks = pl->n ? (pl->n + ch - 1) / ch : 0;
  if (pl->host.kind < 0) return (pl->n ? 1 : 0) + chunks * (pl->host.richardson ? 4 : 3);
  return chunks * 3;
}

int setup_fused(qb200_plan* pl, unsigned n_chunks) {
  pl->fused_ok = 0;
  if (pl->host.kind >= 0) return 0;
  if (pl->host.method == kMethodOptimalLocalSigma) {
    pl->fused_why = "the sigma-optimal method runs on the point-array kernels";
    return 0;
  }
  if (!fused2d_prepare(pl->host, n_chunks, &pl->fused, &pl->fused_why)) return 0;
  if (int rc = pl->fused_part.reserve(std::max<size_t>(1, pl->fused.k.n_tiles) *
                                      QB_FUSED_PART_STRIDE * sizeof(double)))
    return rc;
  if (int rc = pl->fused_cols.reserve(pl->fused.cols_bytes)) return rc;
  if (int rc = pl->fused_tickets.reserve(std::max<size_t>(1, pl->host.slices.size()) * sizeof(unsigned int)))
    return rc;
  // zeroed once: the kernel leaves every ticket at zero again
  QB_CUDA(cudaMemsetAsync(pl->fused_tickets.p, 0, std::max<size_t>(1, pl->host.slices.size()) * sizeof(unsigned int),
                          pl->ctx->stream));
  const size_t fb = std::max<size_t>(1, pl->fused.fslices.size()) * sizeof(FusedSlice);
  if (int rc = pl->fused_slices.reserve(fb)) return rc;
  if (!pl->fused.fslices.empty())
    QB_CUDA(cudaMemcpyAsync(pl->fused_slices.p, pl->fused.fslices.data(),
                            pl->fused.fslices.size() * sizeof(FusedSlice), cudaMemcpyHostToDevice,
                            pl->ctx->stream));
  pl->fused_ok = 1;
  return 0;
}

int enqueue_fused_prologue(qb200_plan* plan, cudaStream_t st) {
  const Plan& h = plan->host;
  const int NP = table_points(h.D);
  const int n_a = (int)h.tabs_a.size(), n_b = (int)h.tabs_b.size();
  if (plan->ctx->fused_lean) {
    const int ncol = plan->fused.k.ncol;
    dim3 g((std::max(NP, ncol) + 127) / 128, n_a + 2 * n_b);
    k_fused_prologue<<<g, 128, 0, st>>>(h.c, h.D, NP, n_a, n_b, plan->desc_a.as<TabDesc>(),
                                        plan->desc_b.as<TabDesc>(), plan->geo->gx.as<dd>(),
                                        plan->geo->gw.as<double>(), plan->tab_a.as<AxisD>(),
                                        plan->tab_b.as<AxisR>(), plan->fused_cols.as<double>());
    plan->ctx->launches += 1;
    QB_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((NP + 127) / 128, n_a + n_b);
  k_axis2d<<<grid, 128, 0, st>>>(h.c, NP, n_a, plan->desc_a.as<TabDesc>(),
                                 plan->desc_b.as<TabDesc>(), plan->geo->gx.as<dd>(),
                                 plan->tab_a.as<AxisD>(), plan->tab_b.as<AxisR>());
  fused2d_launch_cols(plan->fused, h, st, plan->desc_b.as<TabDesc>(), plan->tab_b.as<AxisR>(),
                      plan->geo->gw.as<double>(), plan->fused_cols.as<double>());
  plan->ctx->launches += 2;
  QB_CUDA(cudaGetLastError());
  return 0;
}

FusedArgs fused_args_of(qb200_plan* plan, double* d_cells, double* d_summary) {
  const bool lean = plan->ctx->fused_lean;
  return fused2d_args(plan->fused, plan->fused_slices.as<FusedSlice>(),
                      plan->fused_cols.as<double>(), plan->tab_a.as<AxisD>(),
                      plan->geo->gw.as<double>(), plan->fused_part.as<double>(), d_cells,
                      lean ? plan->fused_tickets.as<unsigned int>() : nullptr, lean ? d_summary : nullptr);
}

int enqueue_fused_chunk(qb200_plan* plan, const FusedArgs& args, size_t c, cudaStream_t st,
                        bool overlap_classes = false) {
  qb200_context* ctx = plan->ctx;
  const FusedChunk& fc = plan->fused.chunks[c];
  int present = 0;
  for (int cl = 0; cl < 3; cl++) present += fc.class_tiles[cl + 1] > fc.class_tiles[cl] ? 1 : 0;
  if (overlap_classes && present > 1 && ctx->side[0]) {
    // fork: classes 1 and 2 on the side streams, behind everything st holds so far; join after
    QB_CUDA(cudaEventRecord(ctx->fork_ev, st));
    const cudaStream_t streams[3] = {st, ctx->side[0], ctx->side[1]};
    for (int k = 0; k < 2; k++) QB_CUDA(cudaStreamWaitEvent(ctx->side[k], ctx->fork_ev, 0));
    if (fused2d_launch_chunk(plan->fused, args, c, st, streams)) return fail(-100, "fused kernel launch failed");
    for (int k = 0; k < 2; k++) {
      QB_CUDA(cudaEventRecord(ctx->join_ev[k], ctx->side[k]));
      QB_CUDA(cudaStreamWaitEvent(st, ctx->join_ev[k], 0));
    }
  } else if (fused2d_launch_chunk(plan->fused, args, c, st)) {
    return fail(-100, "fused kernel launch failed");
  }
  for (int cl = 0; cl < 3; cl++)
    plan->ctx->launches += fc.class_tiles[cl + 1] > fc.class_tiles[cl] ? 1 : 0;
  return 0;
}

int enqueue_fused_epilogue(qb200_plan* plan, cudaStream_t st, double* d_summary) {
  if (plan->ctx->fused_lean) return 0;  // the class kernels wrote the summaries
  fused2d_launch_final(plan->fused, plan->host, st, plan->fused_part.as<double>(), d_summary);
  plan->ctx->launches += 1;
  QB_CUDA(cudaGetLastError());
  return 0;
}

// same_stream: the caller runs the plan on the context's own stream right away (the
// synchronous host API), so stream order already makes the uploads visible.
int finish_common(qb200_plan* pl, unsigned n_chunks = 1, bool same_stream = false) {
  if (int rc = upload_plan(pl)) return rc;
  if (int rc = setup_fused(pl, n_chunks)) return rc;
  // uploads above are asynchronous on the context stream from pageable vectors owned by
  // the plan (or already consumed): make them visible to any stream the caller runs on
  if (!same_stream) QB_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  pl->algo = (pl->fused_ok || pl->host.kind >= 0) ? 2 : 1;
  return 0;
}

}  // namespace

extern "C" {

int qb200_version(void) { return QB200_VERSION; }
const char* qb200_last_error(void) { return g_err.c_str(); }

int qb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int qb200_create(int device, qb200_context** out) {
  *out = nullptr;
  const int n = qb200_device_count();
  if (n <= 0)
    return fail(-101, "no CUDA device: the slice integrators have no CPU path");
  if (device < 0 || device >= n) return fail(-102, "bad device index");
  QB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(-103, std::string("built for sm_100a (B200); found ") + prop.name);
  qb200_context* ctx = new qb200_context;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->join_ev[k], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete ctx;
    return fail(-100, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  {
    const char* v = getenv("QB200_OVERLAP_CLASSES");
    ctx->overlap_classes = !(v && *v == '0');
    v = getenv("QB200_FUSED_LEAN");
    ctx->fused_lean = v && *v == '1';
    v = getenv("QB200_GRAPHS");
    ctx->use_graphs = !(v && *v == '0');
  }
  ctx->out_cells.pool = nullptr;
  *out = ctx;
  return 0;
}

void qb200_destroy(qb200_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int k = 0; k < 2; k++) {
    if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    if (ctx->join_ev[k]) cudaEventDestroy(ctx->join_ev[k]);
  }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  for (cudaEvent_t ev : ctx->events) cudaEventDestroy(ev);
  if (ctx->h_summary) cudaFreeHost(ctx->h_summary);
  if (ctx->text) qb200::text_state_destroy(ctx->text);
  delete ctx;
}

uint64_t qb200_launch_count(const qb200_context* ctx) { return ctx ? ctx->launches : 0; }

void* qb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void qb200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

static int create_plan2d(qb200_context* ctx, const qb200_params* params, int method,
                         int richardson, uint32_t dimension, uint32_t n, const int32_t* a_d,
                         const int32_t* a_r, unsigned n_chunks, qb200_plan** out,
                         bool same_stream = false) {
  std::unique_ptr<qb200_plan> pl(new qb200_plan);
  pl->ctx = ctx;
  pl->bind_pool(&ctx->pool);
  std::string err;
  if (int rc = plan_2d(view_of(params), method, richardson, dimension, n, a_d, a_r, &pl->host, &err))
    return fail(rc, err);
  if (int rc = finish_common(pl.get(), n_chunks, same_stream)) return rc;
  if (method == kMethodOptimalLocalSigma && richardson && dimension % 32 == 0 && params->l >= 256 && n > 0) {
    const char* off = getenv("QB200_SO_FAST");
    if (!(off && *off == '0')) {
      qb200_plan* q = nullptr;
      if (0 == create_plan2d(ctx, params, kMethodQuick, richardson, dimension, n, a_d, a_r, 1, &q, same_stream)) {
        if (q->fused_ok && q->fused.mode == 0)
          pl->so_quick = q;
        else
          qb200_plan_destroy(q);
      }
    }
  }
  *out = pl.release();
  return 0;
}

int qb200_plan2d_create(qb200_context* ctx, const qb200_params* params, int method,
                        int richardson, uint32_t dimension, uint32_t n,
                        const int32_t* a_d, const int32_t* a_r, qb200_plan** out) {
  *out = nullptr;
  if (!ctx || !params) return fail(-1, "null argument");
  QB_CUDA(cudaSetDevice(ctx->device));
  return create_plan2d(ctx, params, method, richardson, dimension, n, a_d, a_r, 1, out);
}

int qb200_plan1d_create(qb200_context* ctx, const qb200_params* params, int kind, int richardson,
                        uint32_t dimension, uint32_t n, const int32_t* a, const int32_t* eta,
                        qb200_plan** out) {
  *out = nullptr;
  if (!ctx || !params) return fail(-1, "null argument");
  QB_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<qb200_plan> pl(new qb200_plan);
  pl->ctx = ctx;
  pl->bind_pool(&ctx->pool);
  std::string err;
  if (int rc = plan_1d(view_of(params), kind, richardson, dimension, n, a, eta, &pl->host, &err))
    return fail(rc, err);
  if (int rc = finish_common(pl.get())) return rc;
  *out = pl.release();
  return 0;
}

void qb200_plan_destroy(qb200_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->ctx->device);
  if (plan->so_quick) qb200_plan_destroy(plan->so_quick);
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  delete plan;
}

uint64_t qb200_plan_cells(const qb200_plan* plan) {
  const uint64_t D = (uint64_t)plan->host.D;
  return (uint64_t)plan->n * (plan->host.kind < 0 ? D * D : D);
}

uint32_t qb200_plan_launches(const qb200_plan* plan) {
  if (plan->host.kind >= 0 && plan->algo == 2) return (plan->n + 65534) / 65535;
  if (plan->algo == 2) return fused2d_launches(plan->fused, plan->ctx->fused_lean);
  return plain_launches(plan);
}

int qb200_plan_set_algorithm(qb200_plan* plan, int algo) {
  if (plan->host.kind >= 0) {  // one-dimensional: 2 = the single-launch kernel (default)
    if (algo < 0 || algo > 2) return fail(-21, "unknown algorithm");
    plan->algo = algo == 1 ? 1 : 2;
    return 0;
  }
  if (algo == 0) {
    plan->algo = plan->fused_ok ? 2 : 1;
    return 0;
  }
  if (algo == 1) {
    plan->algo = 1;
    return 0;
  }
  if (algo == 2) {
    if (!plan->fused_ok) return fail(-20, "fused kernel not applicable: " + plan->fused_why);
    plan->algo = 2;
    return 0;
  }
  return fail(-21, "unknown algorithm");
}

int qb200_plan_algorithm(const qb200_plan* plan) { return plan->algo; }

int qb200_plan_run(qb200_plan* plan, void* stream, double* d_cells, double* d_summary) {
  qb200_context* ctx = plan->ctx;
  QB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  if (plan->host.kind >= 0)
    return plan->algo == 2 ? run_fused_1d(plan, st, d_cells, d_summary)
                           : run_plain_1d(plan, st, d_cells, d_summary);
  if (plan->algo == 2) {
    if (plan->n == 0) return 0;
    if (plan->graph_exec && plan->graph_cells == d_cells && plan->graph_summary == d_summary) {
      QB_CUDA(cudaGraphLaunch(plan->graph_exec, st));
      ctx->launches += plan->graph_kernels;
      return 0;
    }
    // second run into the same buffers: capture the step (the first one ran eagerly, so every
    // kernel attribute is set and the pool buffers exist)
    const bool capture = ctx->use_graphs && !plan->graph_failed && plan->last_cells == d_cells &&
                         plan->last_summary == d_summary;
    const uint64_t launches_before = ctx->launches;
    if (capture && cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      plan->graph_failed = true;
      return qb200_plan_run(plan, stream, d_cells, d_summary);
    }
    int rc = enqueue_fused_prologue(plan, st);
    const FusedArgs args = fused_args_of(plan, d_cells, d_summary);
    for (size_t c = 0; rc == 0 && c < plan->fused.chunks.size(); c++)
      rc = enqueue_fused_chunk(plan, args, c, st, /*overlap_classes=*/ctx->overlap_classes);
    if (rc == 0) rc = enqueue_fused_epilogue(plan, st, d_summary);
    if (capture) {
      cudaGraph_t graph = nullptr;
      const cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc == 0 && e == cudaSuccess && graph) {
        if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
        plan->graph_exec = nullptr;
        if (cudaGraphInstantiate(&plan->graph_exec, graph, 0) == cudaSuccess) {
          plan->graph_cells = d_cells;
          plan->graph_summary = d_summary;
          plan->graph_kernels = (uint32_t)(ctx->launches - launches_before);
        } else {
          plan->graph_exec = nullptr;
        }
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->launches = launches_before;  // nothing ran yet: the capture only recorded the step
      if (!plan->graph_exec) {          // could not capture: run eagerly from now on
        plan->graph_failed = true;
        if (rc) return rc;
      }
      return qb200_plan_run(plan, stream, d_cells, d_summary);
    }
    plan->last_cells = d_cells;
    plan->last_summary = d_summary;
    return rc;
  }
  return run_plain_2d(plan, st, d_cells, d_summary);
}

int qb200_plan_finish(const qb200_plan* plan, const double* hs, long double* tp, long double* te,
                      uint32_t* flags) {
  const Plan& h = plan->host;
  for (uint32_t i = 0; i < plan->n; i++) {
    const double* s = hs + (size_t)i * QB200_SUMMARY_STRIDE;
    if (tp) tp[i] = (long double)s[0] + (long double)s[1];
    const bool so = h.kind < 0 && h.method == kMethodOptimalLocalSigma;
    if (so && s[4] < 0.0)
      return fail(-31,
                  "sigma-optimal method: no admissible sigma at the first point, or the walk "
                  "reached sigma = 1 (increasing search); not supported on the GPU");
    if (te) te[i] = so ? total_error_sigma_opt(h, i, s)
                       : (h.kind < 0 ? total_error_2d(h, i, s[2], s[3]) : 0.0L);
    if (flags) {
      uint32_t f = kFlagMethodSimpson | (h.richardson ? kFlagMethodRichardson : 0u);
      if (h.kind < 0 && h.with_error) {
        // Coarse-pass points only decide the warning
        // (src/distribution_slice_compute_richardson.cpp:28-44, 69-72).
        if (s[4] == 0.0) f |= kFlagErrorBoundWarning;
        if (plan->algo == 2 && !plan->fused.has_bound && plan->fused.host_unbounded[i])
          f |= kFlagErrorBoundWarning;
      }
      flags[i] = f;
    }
  }
  return 0;
}

// Error paths: nothing of a failed call may still be in flight when its plan hands its
// buffers back to the pool (the sticky error, if any, is already in g_err).
static void drain(qb200_context* ctx) {
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaGetLastError();
}

static int run_sync(qb200_context* ctx, qb200_plan* pl, double* cells, long double* tp,
                    long double* te, uint32_t* flags) {
  const uint64_t ncells = qb200_plan_cells(pl);
  const size_t sum_bytes = (size_t)pl->n * QB200_SUMMARY_STRIDE * sizeof(double);
  if (int rc = ctx->out_cells.reserve(std::max<size_t>(8, ncells * sizeof(double)))) return rc;
  if (int rc = ctx->out_summary.reserve(std::max<size_t>(8, sum_bytes))) return rc;
  if (ctx->h_summary_bytes < sum_bytes) {
    if (ctx->h_summary) cudaFreeHost(ctx->h_summary);
    ctx->h_summary = nullptr;
    ctx->h_summary_bytes = 0;
    QB_CUDA(cudaHostAlloc(&ctx->h_summary, sum_bytes, cudaHostAllocDefault));
    ctx->h_summary_bytes = sum_bytes;
  }
  double* d_cells = ctx->out_cells.as<double>();
  double* d_summary = ctx->out_summary.as<double>();
  if (pl->host.kind < 0 && pl->algo == 2 && pl->n > 0 && pl->fused.chunks.size() > 1) {
    // Pipelined: the copy of chunk c runs on the copy stream while chunk c + 1 computes.
    const size_t nch = pl->fused.chunks.size();
    while (ctx->events.size() < nch) {
      cudaEvent_t ev;
      QB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      ctx->events.push_back(ev);
    }
    if (int rc = enqueue_fused_prologue(pl, ctx->stream)) return rc;
    const FusedArgs args = fused_args_of(pl, d_cells, d_summary);
    const size_t per = (size_t)pl->host.D * pl->host.D;
    for (size_t c = 0; c < nch; c++) {
      if (int rc = enqueue_fused_chunk(pl, args, c, ctx->stream, ctx->overlap_classes)) return rc;
      QB_CUDA(cudaEventRecord(ctx->events[c], ctx->stream));
      QB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->events[c], 0));
      const FusedChunk& fc = pl->fused.chunks[c];
      const size_t off = (size_t)fc.slot_begin * per, cnt = (size_t)(fc.slot_end - fc.slot_begin) * per;
      if (cnt)
        QB_CUDA(cudaMemcpyAsync(cells + off, d_cells + off, cnt * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    if (int rc = enqueue_fused_epilogue(pl, ctx->stream, d_summary)) return rc;
  } else {
    if (int rc = qb200_plan_run(pl, ctx->stream, d_cells, d_summary)) return rc;
    if (ncells)
      QB_CUDA(cudaMemcpyAsync(cells, d_cells, ncells * sizeof(double), cudaMemcpyDeviceToHost,
                              ctx->stream));
  }
  if (sum_bytes)
    QB_CUDA(cudaMemcpyAsync(ctx->h_summary, d_summary, sum_bytes, cudaMemcpyDeviceToHost,
                            ctx->stream));
  QB_CUDA(cudaStreamSynchronize(ctx->stream));
  QB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  return qb200_plan_finish(pl, (const double*)ctx->h_summary, tp, te, flags);
}

int qb200_slice2d_compute(qb200_context* ctx, const qb200_params* params, int method,
                          int richardson, uint32_t dimension, uint32_t n, const int32_t* a_d,
                          const int32_t* a_r, double* cells, long double* tp, long double* te,
                          uint32_t* flags) {
  if (!ctx || !params) return fail(-1, "null argument");
  QB_CUDA(cudaSetDevice(ctx->device));
  qb200_plan* pl = nullptr;
  // chunks of ~32 MB of results so that copies overlap compute
  const uint64_t bytes = (uint64_t)n * dimension * dimension * sizeof(double);
  const unsigned n_chunks = (unsigned)std::min<uint64_t>(16, std::max<uint64_t>(1, bytes >> 25));
  if (int rc = create_plan2d(ctx, params, method, richardson, dimension, n, a_d, a_r, n_chunks, &pl,
                             /*same_stream=*/true))
    return rc;
  const int rc = run_sync(ctx, pl, cells, tp, te, flags);
  if (rc) drain(ctx);  // enqueued work may still use the plan's pooled buffers
  qb200_plan_destroy(pl);
  return rc;
}

int qb200_slice2d_compute_scaled(qb200_context* ctx, const qb200_params* params, int method,
                                 int richardson, uint32_t dimension, uint32_t store_dimension,
                                 uint32_t n, const int32_t* a_d, const int32_t* a_r,
                                 long double* cells, long double* tp, long double* te,
                                 uint32_t* flags) {
  if (!ctx || !params) return fail(-1, "null argument");
  if (store_dimension == 0 || dimension % store_dimension != 0)
    return fail(-12, "distribution_slice_copy_scale(): Incompatible dimensions.");
  QB_CUDA(cudaSetDevice(ctx->device));
  qb200_plan* pl = nullptr;
  if (int rc = create_plan2d(ctx, params, method, richardson, dimension, n, a_d, a_r, 1, &pl,
                             /*same_stream=*/true))
    return rc;
  int rc = 0;
  do {
    const uint64_t ncells = qb200_plan_cells(pl);
    const size_t per = (size_t)store_dimension * store_dimension;
    const size_t sum_bytes = (size_t)n * QB200_SUMMARY_STRIDE * sizeof(double);
    if ((rc = ctx->out_cells.reserve(std::max<size_t>(8, ncells * sizeof(double))))) break;
    if ((rc = ctx->out_summary.reserve(std::max<size_t>(8, sum_bytes)))) break;
    if ((rc = ctx->out_scaled.reserve(std::max<size_t>(16, (size_t)n * per * 16)))) break;
    if ((rc = ctx->out_status.reserve(sizeof(int)))) break;
    std::vector<double> hs((size_t)n * QB200_SUMMARY_STRIDE);
    int status = 0;
    rc = -100;
    g_err = "qb200_slice2d_compute_scaled: CUDA error";
    if (cudaMemsetAsync(ctx->out_status.p, 0, sizeof(int), ctx->stream) != cudaSuccess) break;
    if (int r2 = qb200_plan_run(pl, ctx->stream, ctx->out_cells.as<double>(), ctx->out_summary.as<double>())) {
      rc = r2;
      break;
    }
    // distribution_slice_copy_scale on the device: the scaled cells leave as x87 long doubles
    for (uint32_t s0 = 0; s0 < n; s0 += 65535) {
      const uint32_t ns = std::min<uint32_t>(65535, n - s0);
      k_scale_x87<<<dim3((unsigned)((per + 255) / 256), ns), 256, 0, ctx->stream>>>(
          (int)dimension, (int)store_dimension,
          ctx->out_cells.as<double>() + (size_t)s0 * dimension * dimension,
          ctx->out_scaled.as<ulonglong2>() + (size_t)s0 * per, ctx->out_status.as<int>());
      ctx->launches++;
    }
    if (cudaGetLastError() != cudaSuccess) break;
    if (n && cudaMemcpyAsync(cells, ctx->out_scaled.p, (size_t)n * per * 16, cudaMemcpyDeviceToHost,
                             ctx->stream) != cudaSuccess)
      break;
    if (n && cudaMemcpyAsync(hs.data(), ctx->out_summary.p, sum_bytes, cudaMemcpyDeviceToHost,
                             ctx->stream) != cudaSuccess)
      break;
    if (cudaMemcpyAsync(&status, ctx->out_status.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) !=
        cudaSuccess)
      break;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) break;
    if (status) {
      rc = fail(-4, "copy_scale: a scaled cell is below the normal long double range");
      break;
    }
    rc = qb200_plan_finish(pl, hs.data(), tp, te, flags);
  } while (0);
  if (rc) drain(ctx);
  qb200_plan_destroy(pl);
  return rc;
}

int qb200_slice1d_compute(qb200_context* ctx, const qb200_params* params, int kind,
                          int richardson, uint32_t dimension, uint32_t n, const int32_t* a,
                          const int32_t* eta, double* cells, long double* tp, uint32_t* flags) {
  qb200_plan* pl = nullptr;
  if (int rc = qb200_plan1d_create(ctx, params, kind, richardson, dimension, n, a, eta, &pl))
    return rc;
  const int rc = run_sync(ctx, pl, cells, tp, nullptr, flags);
  if (rc) drain(ctx);
  qb200_plan_destroy(pl);
  return rc;
}

uint32_t qb200_heuristic_sigma(uint32_t l) { return heuristic_sigma(l); }

int qb200_host_constants(const qb200_params* p, double* out20) {
  HostConsts h;
  const int rc = host_consts_compute(p->m, p->l, p->sigma, p->d_be, p->d_len, p->r_be, p->r_len, &h);
  if (rc) return fail(rc, "bad parameters");
  const DD* v[10] = {&h.kappa, &h.kappa_q, &h.c_over_L, &h.n_over_L, &h.n1_over_L,
                     &h.beta_m, &h.rbeta_m, &h.r_m, &h.d_m, &h.rho};
  for (int i = 0; i < 10; i++) {
    out20[2 * i] = v[i]->hi;
    out20[2 * i + 1] = v[i]->lo;
  }
  return 0;
}

int qb200_measure_fp64_peak(qb200_context* ctx, double* flops) {
  QB_CUDA(cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  DevBuf out;
  if (int rc = out.reserve((size_t)blocks * threads * sizeof(double))) return rc;
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0));
  QB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    QB_CUDA(cudaEventRecord(e0, ctx->stream));
    k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(out.as<double>(), iters, 0.999999, 1e-9);
    ctx->launches++;
    QB_CUDA(cudaEventRecord(e1, ctx->stream));
    QB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3);
    if (rep > 0 && fl > best) best = fl;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *flops = best;
  return 0;
}

}  // extern "C"
