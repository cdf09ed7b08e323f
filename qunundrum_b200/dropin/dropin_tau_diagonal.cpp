// dropin_tau_diagonal.cpp -- reference-side forwarding TU for tau_estimate_diagonal
// (the diagonal half).
//
// What a maintainer of ekera/qunundrum adds to src/. It defines, with the reference's own
// signature (src/tau_estimate.h:111-117),
//
//   bool tau_estimate_diagonal(const Diagonal_Distribution *, Random_State *, uint32_t n,
//                              uint32_t delta_bound, uint32_t eta_bound, long double &tau)
//
// over qb200_diagk_sample (include/qunundrum_b200.h). The reference's tau_estimate.cpp stays in
// the build, compiled with -Dtau_estimate_diagonal=tau_estimate_diagonal_cpu_unused (a rename on
// the command line, no source change), so that the caller --
// estimate_runs_diagonal_distribution (src/main_estimate_runs_diagonal_distribution.cpp:413) --
// links against the function below.
//
// Division of labour for one estimate of n samples (src/tau_estimate.cpp:135-210):
//   * which region a sample falls in is decided here, on the host, by the reference's own
//     diagonal_distribution_sample_region (src/diagonal_distribution.cpp:306-352) on the caller's
//     Random_State: the number of bytes the sample then reads -- (bits(max - min) + 72) / 8,
//     src/random.c:163-164 -- depends on the region, so the position of sample i + 1 in the stream is
//     known only once sample i is placed; per Random_State this walk is serial wherever it runs.
//     What follows the region -- the bytes of alpha_r, those of t_r for an even r, the pivot of the
//     k sampling -- is only READ here (random_generate into the batch's byte buffer, the reference's
//     random_generate_pivot_inclusive), not computed with.
//   * alpha_r = min + (bytes mod (max - min)) (sample_alpha_from_region, src/sample.cpp:78-158: two
//     mpfr_exp2 at 3 m bits and an mpz_mod per sample in the reference), j =
//     sample_j_from_diagonal_alpha_r (src/sample.cpp:354-410: an mpz_invert and a product per sample),
//     k and alpha_phi given (j, eta, pivot) (sample_k_from_diagonal_j_eta_pivot, src/sample.cpp:412-646)
//     come from the GPU for all samples of a batch in ONE call, qb200_diagk_sample_drawn; j never
//     leaves the device. No GMP arithmetic is left in this file.
//   * the sum of alpha_phi^2, its log2 and tau are formed as the reference forms them, in MPFR at
//     PRECISION bits from the double-double alpha_phi the library returns.
//
// Semantics kept:
//   * batching. The caller asks for one estimate per call, 1000 calls per job with the same
//     arguments (TAU_CHUNK_SIZE, src/executables_estimate_runs_distribution.h:23;
//     src/main_estimate_runs_diagonal_distribution.cpp:410-422). The first call computes
//     QB200_TAU_BATCH (default 1000) consecutive estimates with one GPU call and the following
//     calls with the same (distribution, random state, n, bounds) return them in order; the
//     Random_State runs ahead to the end of the batch. With QB200_TAU_BATCH=1 the state after
//     every call is the reference's (what the tests compare); with a batch the states agree at
//     every batch boundary. If the caller changes its arguments in mid-batch the unused estimates
//     are dropped: the results remain correct samples, the stream position then differs from the
//     reference's (never the case in the reference's executable with the default batch).
//   * the random stream: the reference stops reading at the first sample that fails (no slice,
//     |eta| > eta_bound, k out of bounds; src/tau_estimate.cpp:163-188). The first two are seen
//     while drawing. k out of bounds is only known after the GPU call: if it happens before an
//     estimate's last sample, the Random_State is put back to where it was when that estimate
//     began, the draws up to the failing sample are repeated, and the later estimates of the batch
//     are drawn and computed anew. (A Random_State reading /dev/urandom cannot be put back and does
//     not need to be.)
//   * errors are fatal: critical() (src/errors.c).
#include "common.h"
#include "diagonal_distribution.h"
#include "diagonal_distribution_slice.h"
#include "diagonal_parameters.h"
#include "errors.h"
#include "math.h"
#include "random.h"
#include "sample.h"
#include "tau_estimate.h"

#include <gmp.h>
#include <mpfr.h>

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <map>
#include <vector>
#include <utility>
#include <vector>

#include "qunundrum_b200.h"

namespace {

qb200_context* g_ctx = NULL;

struct Stats {
  bool on = false;
  unsigned long calls = 0, samples = 0, replays = 0, regions = 0;
  double s_draw = 0, s_abi = 0, s_sum = 0;
} g_stats;

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void print_stats() {
  if (g_stats.on && g_stats.calls)
    fprintf(stderr,
            "qunundrum_b200 diagonal tau drop-in: %lu estimates, %lu samples; %.3f s choosing regions and reading the "
            "stream on the host (%lu distinct regions), %.3f s inside qb200_diagk_sample_drawn, %.3f s summing; "
            "%lu replays\n",
            g_stats.calls, g_stats.samples, g_stats.s_draw, g_stats.regions, g_stats.s_abi, g_stats.s_sum,
            g_stats.replays);
}

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
  if (0 != qb200_create(device, &g_ctx)) critical("qunundrum_b200: %s", qb200_last_error());
  const char* st = getenv("QB200_DROPIN_STATS");
  if (st && *st && *st != '0') {
    g_stats.on = true;
    atexit(print_stats);
  }
  return g_ctx;
}

// What is the same for every sample of a distribution.
struct Setup {
  bool valid = false;
  uint32_t m = 0, sigma = 0, l = 0, kappa_r = 0, dimension_max = 0, t_words = 1;
  mpz_t d, r;
  qb200_diagk* sampler = NULL;
  qb200_exact* exact = NULL;
  std::map<uint64_t, uint32_t> region_bytes;  // (sign, e, dimension, region) -> bytes of a sample
} g;

void setup_clear() {
  if (!g.valid) return;
  mpz_clear(g.d);
  mpz_clear(g.r);
  g.region_bytes.clear();
  if (g.sampler) qb200_diagk_destroy(g.sampler);
  if (g.exact) qb200_exact_destroy(g.exact);
  g.sampler = NULL;
  g.exact = NULL;
  g.valid = false;
}

void setup_for(const Diagonal_Parameters* p, uint32_t dimension_max) {
  if (g.valid && g.m == p->m && g.sigma == p->sigma && g.l == p->l && g.dimension_max >= dimension_max &&
      0 == mpz_cmp(g.d, p->d) && 0 == mpz_cmp(g.r, p->r))
    return;
  setup_clear();
  g.m = p->m;
  g.sigma = p->sigma;
  g.l = p->l;
  g.dimension_max = dimension_max;
  mpz_init_set(g.d, p->d);
  mpz_init_set(g.r, p->r);
  g.kappa_r = kappa(p->r);
  g.t_words = g.kappa_r ? (g.kappa_r + 31) / 32 : 1;
  if (dimension_max > 0 && g.kappa_r > 64) {
    critical("tau_estimate_diagonal(): r divisible by more than 2^64 is not supported.");
  }
  std::vector<uint8_t> db((mpz_sizeinbase(p->d, 2) + 7) / 8 + 1), rb((mpz_sizeinbase(p->r, 2) + 7) / 8 + 1);
  size_t dn = 0, rn = 0;
  mpz_export(db.data(), &dn, 1, 1, 1, 0, p->d);
  mpz_export(rb.data(), &rn, 1, 1, 1, 0, p->r);
  qb200_params q;
  q.m = p->m;
  q.l = p->l;
  q.sigma = p->sigma;
  q.d_be = db.data();
  q.d_len = dn;
  q.r_be = rb.data();
  q.r_len = rn;
  if (0 != qb200_diagk_create(context(), &q, &g.sampler)) {
    critical("tau_estimate_diagonal(): %s", qb200_last_error());
  }
  // dimension_max = 0: the one-sample entry points below, which take j from their caller
  if (dimension_max > 0 &&
      0 != qb200_exact_create(context(), &q, QB200_EXACT_DIAGONAL, dimension_max, 0, &g.exact)) {
    critical("tau_estimate_diagonal(): %s", qb200_last_error());
  }
  g.valid = true;
}

void setup_for(const Diagonal_Distribution* distribution) {
  uint32_t dimension_max = 1;
  for (uint32_t i = 0; i < distribution->count; i++) {
    const uint32_t dim = distribution->slices[i]->dimension;
    if (dim == 0 || (dim & (dim - 1)) != 0 || dim > 16384) {
      critical("tau_estimate_diagonal(): a slice dimension of %u (not a power of two up to 16384) is not "
               "supported by the GPU sampler.", dim);
    }
    if (dim > dimension_max) dimension_max = dim;
  }
  setup_for(&distribution->parameters, dimension_max);
}

// The rows of a batch: what the GPU call takes.
struct Rows {
  std::vector<qb200_exact_region> region;
  std::vector<uint32_t> t;       // t_r, g.t_words per sample (zero for an odd r)
  std::vector<int32_t> eta;
  std::vector<long double> pivot;
  std::vector<uint8_t> bytes;    // the samples' random bytes, one after the other
  void clear() {
    region.clear();
    t.clear();
    eta.clear();
    pivot.clear();
    bytes.clear();
  }
};

// The bytes sample_alpha_from_region reads for a region (random_generate_mpz, src/random.c:163-164):
// stream layout, once per region.
uint32_t bytes_of_region(int32_t min_log_alpha, uint32_t region, uint32_t dimension) {
  const uint64_t key = ((uint64_t)(uint32_t)(min_log_alpha + 65536) << 40) | ((uint64_t)dimension << 20) | region;
  auto it = g.region_bytes.find(key);
  if (it != g.region_bytes.end()) return it->second;
  uint32_t n = 0;
  if (0 != qb200_exact_region_bytes(g.exact, min_log_alpha, region, dimension, &n)) {
    critical("tau_estimate_diagonal(): %s", qb200_last_error());
  }
  g.region_bytes[key] = n;
  g_stats.regions++;
  return n;
}

// diagonal_distribution_sample_j_eta (src/diagonal_distribution.cpp:410-472) as far as the stream
// is concerned: the same reads from the same Random_State in the same order; the arithmetic on what
// is read happens on the device.
bool draw_row(const Diagonal_Distribution* distribution, Random_State* rs, Rows& rows, int32_t* eta) {
  double min_log_alpha_r, max_log_alpha_r;
  if (FALSE == diagonal_distribution_sample_region(distribution, rs, &min_log_alpha_r, &max_log_alpha_r, eta)) {
    return false;  // :366-383, :428-443
  }
  // the region as (slice coordinate, index, dimension): diagonal_distribution_slice_region_coordinates
  // (src/diagonal_distribution_slice.cpp:153-179) formed min = sgn (e + i / D), max = sgn (e + (i + 1) / D)
  // in doubles, exactly (D is a power of two)
  if (sgn_d(min_log_alpha_r) != sgn_d(max_log_alpha_r)) {
    critical("sample_alpha_from_region(): Incompatible signs for min_log_alpha and max_log_alpha.");
  }
  const double lo = abs_d(min_log_alpha_r), hi = abs_d(max_log_alpha_r);
  if (lo >= hi) {
    critical("sample_alpha_from_region(): Incompatible absolute values for min_log_alpha and max_log_alpha.");
  }
  const double e = floor(lo), dim = 1.0 / (hi - lo);
  qb200_exact_region q;
  q.min_log_alpha = (int32_t)(sgn_d(min_log_alpha_r) == -1 ? -e : e);
  q.dimension = (uint32_t)dim;
  q.region = (uint32_t)((lo - e) * dim);
  q.length = bytes_of_region(q.min_log_alpha, q.region, q.dimension);
  q.offset = rows.bytes.size();
  // sample_alpha_from_region -> random_generate_mpz(alpha, max - min, rs), src/sample.cpp:130
  rows.bytes.resize(rows.bytes.size() + q.length);
  random_generate(&rows.bytes[q.offset], q.length, rs);
  rows.region.push_back(q);
  // sample_j_from_diagonal_alpha_r -> random_generate_mpz(t_r, 2^kappa_r, rs), src/sample.cpp:370-375:
  // (kappa_r + 1 + 72) / 8 bytes, big-endian, reduced modulo 2^kappa_r -- its low kappa_r bits
  const size_t at = rows.t.size();
  rows.t.resize(at + g.t_words, 0u);
  if (g.kappa_r > 0) {
    uint8_t buf[32];
    const uint32_t len = (g.kappa_r + 1 + 64 + 8) / 8;  // <= 17
    random_generate(buf, len, rs);
    uint64_t v = 0;
    for (uint32_t i = 0; i < 8; i++) v = (v << 8) | buf[len - 8 + i];
    if (g.kappa_r < 64) v &= ((uint64_t)1 << g.kappa_r) - 1;
    rows.t[at] = (uint32_t)v;
    if (g.t_words > 1) rows.t[at + 1] = (uint32_t)(v >> 32);
  }
  return true;
}

// One estimate's draws (src/tau_estimate.cpp:158-188 without the arithmetic): one row per sample
// whose k is wanted. Returns the number of rows appended; *failed is set if the estimate fails on
// the host's side already -- no slice (diagonal_distribution_sample_j_eta returns FALSE) or
// |eta| > eta_bound -- with the stream where the reference leaves it in that case (the rows before
// the failing sample still go to the GPU: an earlier k out of bounds would have stopped the reference
// earlier). `limit` < n: replay of the first `limit` samples only.
uint32_t draw_estimate(const Diagonal_Distribution* distribution, Random_State* rs, uint32_t limit,
                       uint32_t eta_bound, Rows& rows, bool* failed) {
  *failed = false;
  for (uint32_t i = 0; i < limit; i++) {
    int32_t e = 0;
    const size_t n_region = rows.region.size(), n_t = rows.t.size(), n_bytes = rows.bytes.size();
    if (!draw_row(distribution, rs, rows, &e)) {
      *failed = true;
      return i;
    }
    // sample_k_from_diagonal_j_eta (src/sample.cpp:648-675) draws the pivot next; the check of
    // eta follows the k sampling (src/tau_estimate.cpp:175-181) and fails the estimate either way
    const long double p = random_generate_pivot_inclusive(rs);
    if (abs_i(e) > eta_bound) {
      rows.region.resize(n_region);  // the failing sample's k is not wanted
      rows.t.resize(n_t);
      rows.bytes.resize(n_bytes);
      *failed = true;
      return i;
    }
    rows.eta.push_back(e);
    rows.pivot.push_back(p);
  }
  return limit;
}

// tau from the n samples starting at row `first` (src/tau_estimate.cpp:185-201).
long double tau_of(uint32_t n, const double* x_hi, const double* x_lo, const int32_t* status,
                   const Diagonal_Parameters* p) {
  mpfr_t alpha, sum;
  mpfr_init2(alpha, PRECISION);
  mpfr_init2(sum, PRECISION);
  mpfr_set_ui(sum, 0, MPFR_RNDN);
  const long shift = (long)g.m + (long)g.sigma - (long)g.l;
  for (uint32_t i = 0; i < n; i++) {
    mpfr_set_d(alpha, x_hi[i], MPFR_RNDN);
    mpfr_add_d(alpha, alpha, x_lo[i], MPFR_RNDN);
    if (status[i] == QB200_DIAGK_OK_NEGATIVE_PHI) {  // alpha_phi = 2^(m+sigma-l) (x - 2^l)
      mpfr_t q;
      mpfr_init2(q, PRECISION);
      mpfr_set_ui_2exp(q, 1, (mpfr_exp_t)g.l, MPFR_RNDN);
      mpfr_sub(alpha, alpha, q, MPFR_RNDN);
      mpfr_clear(q);
    }
    mpfr_mul_2si(alpha, alpha, shift, MPFR_RNDN);
    mpfr_sqr(alpha, alpha, MPFR_RNDN);
    mpfr_add(sum, sum, alpha, MPFR_RNDN);
  }
  mpfr_div_ui(sum, sum, n, MPFR_RNDN);
  mpfr_log2(sum, sum, MPFR_RNDN);
  const long double tau = mpfr_get_ld(sum, MPFR_RNDN) / 2.0f - (p->m + p->sigma - p->l);
  mpfr_clear(alpha);
  mpfr_clear(sum);
  return tau;
}

// The results of a batch of consecutive estimates with the same arguments.
struct Batch {
  const Diagonal_Distribution* distribution = NULL;
  Random_State* rs = NULL;
  uint32_t n = 0, delta_bound = 0, eta_bound = 0;
  std::vector<long double> tau;
  std::vector<uint8_t> ok;
  size_t next = 0;
  Random_State state_after;  // the generator as the batch left it: a caller that re-seeds or draws from it
                             // between calls gets a new batch, not stale estimates
} g_batch;

void compute_batch(const Diagonal_Distribution* distribution, Random_State* rs, uint32_t n,
                   uint32_t delta_bound, uint32_t eta_bound, uint32_t B) {
  g_batch.tau.assign(B, DBL_MAX);
  g_batch.ok.assign(B, 0);
  const bool can_rewind = (NULL == rs->random_device);
  Rows rows, replay;
  std::vector<int32_t> status, exact_status;
  std::vector<double> x_hi, x_lo;
  std::vector<Random_State> entry;
  struct Est {
    size_t row;       // first row
    uint32_t count;   // rows
    bool failed;      // failed while drawing
  };
  std::vector<Est> est;
  uint32_t t0 = 0;
  while (t0 < B) {
    rows.clear();
    entry.clear();
    est.clear();
    double t = now_s();
    for (uint32_t e = t0; e < B; e++) {
      if (can_rewind) entry.push_back(*rs);
      Est s;
      s.row = rows.eta.size();
      s.count = draw_estimate(distribution, rs, n, eta_bound, rows, &s.failed);
      est.push_back(s);
    }
    g_stats.s_draw += now_s() - t;
    const uint32_t count = (uint32_t)rows.eta.size();
    g_stats.samples += count;
    x_hi.resize(count);
    x_lo.resize(count);
    status.resize(count);
    exact_status.resize(count);
    if (count) {
      t = now_s();
      if (rows.bytes.empty()) rows.bytes.push_back(0);
      if (0 != qb200_diagk_sample_drawn(g.sampler, g.exact, count, rows.region.data(),
                                        g.kappa_r ? rows.t.data() : NULL, rows.bytes.data(), rows.bytes.size(),
                                        rows.eta.data(), rows.pivot.data(), delta_bound, NULL, x_hi.data(),
                                        x_lo.data(), NULL, status.data(), exact_status.data())) {
        critical("tau_estimate_diagonal(): %s", qb200_last_error());
      }
      g_stats.s_abi += now_s() - t;
      for (uint32_t i = 0; i < count; i++) {
        if (exact_status[i] != QB200_EXACT_OK) {
          critical("tau_estimate_diagonal(): the GPU sampler declined a region (status %d).", exact_status[i]);
        }
      }
    }
    t = now_s();
    uint32_t restart = B;
    for (uint32_t e = t0; e < B; e++) {
      const Est& s = est[e - t0];
      uint32_t bad = s.count;
      for (uint32_t i = 0; i < s.count; i++) {
        if (status[s.row + i] == QB200_DIAGK_GAVE_UP) {
          critical("tau_estimate_diagonal(): sample_k_from_diagonal_j_eta_pivot(): gave up after 2^22 steps "
                   "(delta_bound = %u).", delta_bound);
        }
        if (status[s.row + i] == QB200_DIAGK_OUT_OF_BOUNDS) {
          bad = i;
          break;
        }
      }
      if (bad == s.count) {
        if (!s.failed) {
          g_batch.tau[e] = tau_of(n, &x_hi[s.row], &x_lo[s.row], &status[s.row], &distribution->parameters);
          g_batch.ok[e] = 1;
        }
        continue;  // (failed while drawing: DBL_MAX, FALSE, the stream is where the reference leaves it)
      }
      // k ran out of bounds at sample `bad`: the reference stops reading there
      // (src/tau_estimate.cpp:163-173). If the estimate's draws went on after that sample, what was
      // drawn is not what the reference draws: back to the estimate's entry state, its first
      // bad + 1 samples again, the later estimates of the batch anew.
      const bool drew_on = s.failed || bad + 1 < n;
      if (drew_on && can_rewind) {
        *rs = entry[e - t0];
        bool f2;
        replay.clear();
        draw_estimate(distribution, rs, bad + 1, eta_bound, replay, &f2);
        g_stats.replays++;
        restart = e + 1;
        break;
      }
    }
    g_stats.s_sum += now_s() - t;
    t0 = restart;
  }
}

}  // namespace

bool tau_estimate_diagonal(const Diagonal_Distribution* const distribution, Random_State* const random_state,
                           const uint32_t n, const uint32_t delta_bound, const uint32_t eta_bound,
                           long double& tau) {
  if (0 == n) {  // the reference's loop does not run: result stays FALSE, nothing drawn
    tau = DBL_MAX;
    return false;
  }
  g_stats.calls++;
  Batch& b = g_batch;
  if (!(b.next < b.ok.size() && b.distribution == distribution && b.rs == random_state && b.n == n &&
        b.delta_bound == delta_bound && b.eta_bound == eta_bound &&
        0 == memcmp(random_state, &b.state_after, sizeof(Random_State)))) {
    setup_for(distribution);
    const int want = env_int("QB200_TAU_BATCH", 1000);
    compute_batch(distribution, random_state, n, delta_bound, eta_bound, (uint32_t)(want > 0 ? want : 1));
    b.distribution = distribution;
    b.rs = random_state;
    b.n = n;
    b.delta_bound = delta_bound;
    b.eta_bound = eta_bound;
    b.next = 0;
    memcpy(&b.state_after, random_state, sizeof(Random_State));
  }
  tau = b.tau[b.next];
  return b.ok[b.next++] != 0;
}

#ifdef QB200_DROPIN_SAMPLE_K
// Optional second symbol: the reference's one-sample entry point itself (src/sample.h:184-191), for a
// build that compiles the reference's sample.cpp with
//   -Dsample_k_from_diagonal_j_eta_pivot=sample_k_from_diagonal_j_eta_pivot_cpu_unused
// (solve_diagonal_*; the reference's own known-answer test
// test_sample_k_from_diagonal_j_eta_pivot_kat(), src/test/test_sample.cpp:679-836, which
// integration/tools/sample_k_kat_check.cpp runs against this function). One GPU call per sample:
// a latency of tens of microseconds against the reference's 0.1 ... 10 ms.
bool sample_k_from_diagonal_j_eta_pivot(const Diagonal_Parameters* const parameters, long double pivot,
                                        const mpz_t j, const int32_t eta, const uint32_t delta_bound, mpz_t k,
                                        mpfr_t alpha_phi) {
  if ((pivot < 0) || (pivot > 1)) {
    critical("sample_k_from_diagonal_j_eta_pivot(): The pivot is out of bounds.");
  }
  setup_for(parameters, 0);
  // j + a 2^(m+sigma) gives the same k and alpha_phi as j (r j changes by a multiple of 2^(m+sigma) r):
  // the ABI takes j on [0, 2^(m+sigma))
  mpz_t jr;
  mpz_init(jr);
  mpz_fdiv_r_2exp(jr, j, g.m + g.sigma);
  std::vector<uint32_t> row(qb200_diagk_j_limbs(g.sampler), 0u), krow(qb200_diagk_k_limbs(g.sampler), 0u);
  size_t cnt = 0;
  mpz_export(row.data(), &cnt, -1, 4, 0, 0, jr);
  mpz_clear(jr);
  double x_hi = 0, x_lo = 0;
  int32_t status = 0;
  if (0 != qb200_diagk_sample(g.sampler, 1, row.data(), &eta, &pivot, delta_bound, krow.data(), &x_hi, &x_lo, NULL,
                              &status)) {
    critical("sample_k_from_diagonal_j_eta_pivot(): %s", qb200_last_error());
  }
  if (status == QB200_DIAGK_GAVE_UP) {
    critical("sample_k_from_diagonal_j_eta_pivot(): gave up after 2^22 steps (delta_bound = %u).", delta_bound);
  }
  if (status == QB200_DIAGK_OUT_OF_BOUNDS) {  // src/sample.cpp:622-636
    mpz_set_ui(k, 0);
    if (NULL != alpha_phi) mpfr_set_ui(alpha_phi, 0, MPFR_RNDN);
    return false;
  }
  mpz_import(k, krow.size(), -1, 4, 0, 0, krow.data());
  if (NULL != alpha_phi) {
    const bool negative_phi = (status == QB200_DIAGK_OK_NEGATIVE_PHI);  // alpha_phi = 2^(m+sigma-l) (x - 2^l)
    mpfr_t a;
    mpfr_init2(a, 128 + (negative_phi ? g.l : 0));
    mpfr_set_d(a, x_hi, MPFR_RNDN);
    mpfr_add_d(a, a, x_lo, MPFR_RNDN);  // exact
    if (negative_phi) {
      mpfr_t p;
      mpfr_init2(p, 64);
      mpfr_set_ui_2exp(p, 1, (mpfr_exp_t)g.l, MPFR_RNDN);
      mpfr_sub(a, a, p, MPFR_RNDN);       // exact
      mpfr_clear(p);
    }
    mpfr_mul_2si(a, a, (long)g.m + (long)g.sigma - (long)g.l, MPFR_RNDN);
    mpfr_set(alpha_phi, a, MPFR_RNDN);  // src/sample.cpp:577-579
    mpfr_clear(a);
  }
  return true;
}

// ... and its integrand (src/diagonal_probability.h), for a build that compiles diagonal_probability.cpp
// with -Ddiagonal_probability_approx_h=diagonal_probability_approx_h_cpu_unused: h at the angle
// phi = 2 pi x / 2^l, i.e. x = phi 2^l / (2 pi) handed to qb200_diagk_h as an unevaluated sum of two
// doubles. (The walk above does not come through here: the kernel forms x itself, exactly.)
void diagonal_probability_approx_h(mpfr_t norm, const mpfr_t phi, const Diagonal_Parameters* const parameters) {
  if (0 == mpfr_cmp_ui(phi, 0)) {  // src/diagonal_probability.cpp:104-108
    mpfr_set_ui(norm, 1, MPFR_RNDN);
    return;
  }
  setup_for(parameters, 0);
  const mpfr_prec_t prec = (mpfr_get_prec(phi) > 192 ? mpfr_get_prec(phi) : 192) + 64;
  mpfr_t x, two_pi;
  mpfr_init2(x, prec);
  mpfr_init2(two_pi, prec);
  mpfr_const_pi(two_pi, MPFR_RNDN);
  mpfr_mul_2si(two_pi, two_pi, 1, MPFR_RNDN);
  mpfr_div(x, phi, two_pi, MPFR_RNDN);
  mpfr_mul_2si(x, x, (long)g.l, MPFR_RNDN);
  const double x_hi = mpfr_get_d(x, MPFR_RNDN);
  mpfr_sub_d(x, x, x_hi, MPFR_RNDN);
  const double x_lo = mpfr_get_d(x, MPFR_RNDN);
  mpfr_clear(x);
  mpfr_clear(two_pi);
  long double h = 0;
  if (0 != qb200_diagk_h(g.sampler, 1, &x_hi, &x_lo, &h)) {
    critical("diagonal_probability_approx_h(): %s", qb200_last_error());
  }
  mpfr_set_ld(norm, h, MPFR_RNDN);
}
#endif
