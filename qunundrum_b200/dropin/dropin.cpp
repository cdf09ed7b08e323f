// dropin.cpp -- the reference-side forwarding translation unit.
//
// This file is what a maintainer of ekera/qunundrum adds to the reference's
// src/ directory, IN PLACE OF the six translation units
//
//   distribution_slice_compute.cpp            distribution_slice_compute_richardson.cpp
//   linear_distribution_slice_compute.cpp     linear_distribution_slice_compute_richardson.cpp
//   diagonal_distribution_slice_compute.cpp   diagonal_distribution_slice_compute_richardson.cpp
//
// (and probability.cpp / linear_probability.cpp / diagonal_probability.cpp stop
// being on the generation path). It defines the same six functions, with the
// same C++ signatures, over the C ABI of libqunundrum_b200.so, so that the
// generator clients (src/main_generate_distribution.cpp:1115 etc.), the MPI
// protocol, the slice containers and the text format are untouched.
//
// It includes the REFERENCE'S OWN headers (never copied into this repository);
// here it is compiled against /root/reference/src only to prove that it builds
// and behaves (tests/test_dropin_gpu.py).
//
// Conventions kept:
//   * the caller owns and has *_slice_init()ed the slice; the callee fills
//     norm_matrix / norm_vector, total_probability, total_error, the coordinates
//     (and eta) and REPLACES the method bits of flags
//     (src/distribution_slice_compute.cpp:410-418, ..._richardson.cpp:69);
//   * total_probability is the sequential long double sum in the reference's
//     own loop order (src/distribution_slice_compute_richardson.cpp:47-64);
//   * errors are fatal: critical() -> exit(-1) (src/errors.c);
//   * one GPU per worker rank: device = (local rank - 1) mod #GPUs (rank 0 is the
//     server and never integrates), overridable with QB200_DEVICE.
#include "common.h"
#include "diagonal_distribution_slice.h"
#include "diagonal_parameters.h"
#include "distribution_slice.h"
#include "errors.h"
#include "linear_distribution_slice.h"
#include "parameters.h"

#include <gmp.h>

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <vector>

#include "qunundrum_b200.h"

namespace {

qb200_context* g_ctx = NULL;

// QB200_DROPIN_STATS=1: print, at exit, how many slices this process integrated and
// the wall time spent inside the C ABI calls (to tell GPU time from protocol time).
struct Stats {
  double seconds;
  unsigned long calls;
  bool on;
} g_stats = {0.0, 0, false};

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void print_stats() {
  if (g_stats.on && g_stats.calls)
    fprintf(stderr, "qunundrum_b200 drop-in: %lu slice calls, %.3f s inside the drop-in functions (%.1f us per call)\n",
            g_stats.calls, g_stats.seconds, 1e6 * g_stats.seconds / (double)g_stats.calls);
}

struct Timed {
  double t0;
  Timed() : t0(g_stats.on ? now_s() : 0.0) {}
  ~Timed() {
    if (g_stats.on) {
      g_stats.seconds += now_s() - t0;
      g_stats.calls++;
    }
  }
};

int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : fallback;
}

qb200_context* context() {
  if (g_ctx) return g_ctx;
  const int n = qb200_device_count();
  if (n <= 0) critical("qunundrum_b200: no CUDA device (there is no CPU path).");
  int device = env_int("QB200_DEVICE", -1);
  if (device < 0) {
    int local = env_int("OMPI_COMM_WORLD_LOCAL_RANK", -1);
    if (local < 0) local = env_int("MPI_LOCALRANKID", -1);
    if (local < 0) local = env_int("SLURM_LOCALID", -1);
    if (local < 0) local = env_int("QB200_MINIMPI_RANK", 1);
    device = ((local - 1) % n + n) % n;
  }
  if (0 != qb200_create(device, &g_ctx)) {
    critical("qunundrum_b200: %s", qb200_last_error());
  }
  const char* st = getenv("QB200_DROPIN_STATS");
  if (st && *st && *st != '0') {
    g_stats.on = true;
    atexit(print_stats);
  }
  return g_ctx;
}

struct Exported {
  std::vector<uint8_t> d, r;
  qb200_params p;
};

void export_z(const mpz_t z, std::vector<uint8_t>& out) {
  out.assign((mpz_sizeinbase(z, 2) + 7) / 8 + 1, 0);
  size_t count = 0;
  mpz_export(out.data(), &count, 1, 1, 1, 0, z);  // big-endian magnitude bytes
  out.resize(count ? count : 1);
}

void fill(Exported& e, uint32_t m, uint32_t l, uint32_t sigma, const mpz_t d, const mpz_t r) {
  export_z(d, e.d);
  export_z(r, e.r);
  e.p.m = m;
  e.p.l = l;
  e.p.sigma = sigma;
  e.p.d_be = e.d.data();
  e.p.d_len = e.d.size();
  e.p.r_be = e.r.data();
  e.p.r_len = e.r.size();
}

void compute_2d(Distribution_Slice* const slice, const Parameters* const parameters,
                const Distribution_Slice_Compute_Method method, const int32_t min_log_alpha_d,
                const int32_t min_log_alpha_r, const int richardson, const char* who) {
  const uint32_t dimension = slice->dimension;
  Exported e;
  fill(e, parameters->m, parameters->l, 0, parameters->d, parameters->r);
  // pinned result buffer, reused across calls (one integrating thread per rank)
  static double* cells = NULL;
  static size_t cells_cap = 0;
  if (cells_cap < (size_t)dimension * dimension) {
    qb200_host_free(cells);
    cells_cap = (size_t)dimension * dimension;
    cells = (double*)qb200_host_alloc(cells_cap * sizeof(double));
    if (NULL == cells) critical("%s(): Failed to allocate memory.", who);
  }
  long double total_probability = 0, total_error = 0;
  uint32_t flags = 0;
  qb200_context* const ctx = context();
  Timed timed;
  if (0 != qb200_slice2d_compute(ctx, &e.p, (int)method, richardson, dimension, 1,
                                 &min_log_alpha_d, &min_log_alpha_r, cells,
                                 &total_probability, &total_error, &flags)) {
    critical("%s(): %s", who, qb200_last_error());
  }
  slice->total_probability = 0;
  if (richardson) {
    for (uint32_t i = 0; i < dimension; i++) {
      for (uint32_t j = 0; j < dimension; j++) {
        slice->norm_matrix[dimension * j + i] = cells[(size_t)dimension * j + i];
        slice->total_probability += slice->norm_matrix[dimension * j + i];
      }
    }
  } else {
    for (uint32_t i = 0; i < dimension; i++) {
      for (uint32_t j = 0; j < dimension; j++) {
        slice->norm_matrix[i + dimension * j] = cells[i + (size_t)dimension * j];
        slice->total_probability += slice->norm_matrix[i + dimension * j];
      }
    }
  }
  slice->total_error = total_error;
  slice->min_log_alpha_d = min_log_alpha_d;
  slice->min_log_alpha_r = min_log_alpha_r;
  slice->flags &= ~(SLICE_FLAGS_MASK_METHOD);
  slice->flags |= flags;
}

void compute_1d(long double* const norm_vector, const uint32_t dimension, const qb200_params& p,
                const int kind, const int32_t min_log_alpha, const int32_t eta,
                const int richardson, long double* total_probability, uint32_t* flags,
                const char* who) {
  std::vector<double> cells(dimension);
  long double tp = 0;
  qb200_context* const ctx = context();
  Timed timed;
  if (0 != qb200_slice1d_compute(ctx, &p, kind, richardson, dimension, 1, &min_log_alpha,
                                 &eta, cells.data(), &tp, flags)) {
    critical("%s(): %s", who, qb200_last_error());
  }
  *total_probability = 0;
  for (uint32_t i = 0; i < dimension; i++) {
    norm_vector[i] = cells[i];
    *total_probability += norm_vector[i];
  }
}

void compute_linear(Linear_Distribution_Slice* const slice, const Parameters* const parameters,
                    const Linear_Distribution_Slice_Compute_Target target,
                    const int32_t min_log_alpha, const int richardson, const char* who) {
  if ((LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D != target) &&
      (LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_R != target)) {
    critical("%s(): Unknown target.", who);
  }
  Exported e;
  fill(e, parameters->m, parameters->l, 0, parameters->d, parameters->r);
  uint32_t flags = 0;
  compute_1d(slice->norm_vector, slice->dimension, e.p,
             (LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D == target) ? QB200_KIND_LINEAR_D
                                                                    : QB200_KIND_LINEAR_R,
             min_log_alpha, 0, richardson, &slice->total_probability, &flags, who);
  slice->total_error = 0;
  slice->min_log_alpha = min_log_alpha;
  slice->flags &= ~(SLICE_FLAGS_MASK_METHOD);
  slice->flags |= flags;
}

void compute_diagonal(Diagonal_Distribution_Slice* const slice,
                      const Diagonal_Parameters* const parameters, const int32_t min_log_alpha_r,
                      const int32_t eta, const int richardson, const char* who) {
  Exported e;
  fill(e, parameters->m, parameters->l, parameters->sigma, parameters->d, parameters->r);
  uint32_t flags = 0;
  compute_1d(slice->norm_vector, slice->dimension, e.p, QB200_KIND_DIAGONAL, min_log_alpha_r, eta,
             richardson, &slice->total_probability, &flags, who);
  slice->total_error = 0;
  slice->min_log_alpha_r = min_log_alpha_r;
  slice->eta = eta;
  slice->flags &= ~(SLICE_FLAGS_MASK_METHOD);
  slice->flags |= flags;
}

}  // namespace

void distribution_slice_compute(Distribution_Slice* const slice,
                                const Parameters* const parameters,
                                const Distribution_Slice_Compute_Method method,
                                const int32_t min_log_alpha_d, const int32_t min_log_alpha_r) {
  compute_2d(slice, parameters, method, min_log_alpha_d, min_log_alpha_r, 0,
             "distribution_slice_compute");
}

void distribution_slice_compute_richardson(Distribution_Slice* const slice,
                                           const Parameters* const parameters,
                                           const Distribution_Slice_Compute_Method method,
                                           const int32_t min_log_alpha_d,
                                           const int32_t min_log_alpha_r) {
  compute_2d(slice, parameters, method, min_log_alpha_d, min_log_alpha_r, 1,
             "distribution_slice_compute_richardson");
}

void linear_distribution_slice_compute(Linear_Distribution_Slice* const slice,
                                       const Parameters* const parameters,
                                       const Linear_Distribution_Slice_Compute_Target target,
                                       const int32_t min_log_alpha) {
  compute_linear(slice, parameters, target, min_log_alpha, 0, "linear_distribution_slice_compute");
}

void linear_distribution_slice_compute_richardson(
    Linear_Distribution_Slice* const slice, const Parameters* const parameters,
    const Linear_Distribution_Slice_Compute_Target target, const int32_t min_log_alpha) {
  compute_linear(slice, parameters, target, min_log_alpha, 1,
                 "linear_distribution_slice_compute_richardson");
}

void diagonal_distribution_slice_compute(Diagonal_Distribution_Slice* const slice,
                                         const Diagonal_Parameters* const parameters,
                                         const int32_t min_log_alpha_r, const int32_t eta) {
  compute_diagonal(slice, parameters, min_log_alpha_r, eta, 0,
                   "diagonal_distribution_slice_compute");
}

void diagonal_distribution_slice_compute_richardson(
    Diagonal_Distribution_Slice* const slice, const Diagonal_Parameters* const parameters,
    const int32_t min_log_alpha_r, const int32_t eta) {
  compute_diagonal(slice, parameters, min_log_alpha_r, eta, 1,
                   "diagonal_distribution_slice_compute_richardson");
}
