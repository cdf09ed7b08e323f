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
//   * (j, eta) of every sample is drawn here, on the host, by the reference's own
//     diagonal_distribution_sample_region (src/diagonal_distribution.cpp:306-352), followed by the
//     arithmetic of sample_alpha_from_region (src/sample.cpp:77-157) and
//     sample_j_from_diagonal_alpha_r (src/sample.cpp:352-410) with GMP, step by step the same
//     calls on the same Random_State -- except that what these two functions recompute for
//     every sample is kept: the integer bounds round(2^|log alpha|) of a region (two mpfr_exp2
//     at 3 m bits per sample in the reference) and (r / 2^kappa_r)^-1 mod 2^(m + sigma) (one
//     mpz_invert per sample). Same integers, same draws, same j.
//   * k and alpha_phi given (j, eta, pivot) -- sample_k_from_diagonal_j_eta_pivot
//     (src/sample.cpp:412-646), the part that evaluates diagonal_probability_approx_h at
//     2 (m + sigma) bits -- come from the GPU for all n samples in one call.
//   * the sum of alpha_phi^2, its log2 and tau are formed as the reference forms them, in MPFR at
//     PRECISION bits from the double-double alpha_phi the library returns.
//
// Semantics kept:
//   * the random stream: the reference stops reading at the first sample that fails (no slice, k
//     out of bounds, |eta| > eta_bound; src/tau_estimate.cpp:163-188). All n samples are drawn
//     before the GPU is asked, so when sample i < n - 1 fails the Random_State is put back to where
//     it was on entry and the draws of samples 0 .. i are repeated; afterwards the state is
//     what the reference leaves behind. (A Random_State reading /dev/urandom cannot be put back and
//     does not need to be.)
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
#include <utility>
#include <vector>

#include "qunundrum_b200.h"

namespace {

qb200_context* g_ctx = NULL;

struct Stats {
  bool on = false;
  unsigned long calls = 0, samples = 0, replays = 0, bounds = 0;
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
            "qunundrum_b200 diagonal tau drop-in: %lu estimates, %lu samples; %.3f s drawing (j, eta) on the host "
            "(%lu bounds 2^|log alpha| computed), %.3f s inside qb200_diagk_sample, %.3f s summing; %lu replays\n",
            g_stats.calls, g_stats.samples, g_stats.s_draw, g_stats.bounds, g_stats.s_abi, g_stats.s_sum,
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
  uint32_t m = 0, sigma = 0, l = 0, kappa_r = 0;
  mpz_t d, r, inverse, pow2_n;  // inverse = (r / 2^kappa_r)^-1 mod 2^(m + sigma)
  qb200_diagk* sampler = NULL;
  uint32_t j_limbs = 0;
  struct Bounds {
    mpz_t min_alpha, modulus;  // round(2^|min_log|), round(2^|max_log|) - round(2^|min_log|)
  };
  std::map<std::pair<double, double>, Bounds*> regions;
  struct Z {
    mpz_t z;
  };
  std::map<std::pair<double, uint32_t>, Z*> bounds;
} g;

void setup_clear() {
  if (!g.valid) return;
  mpz_clear(g.d);
  mpz_clear(g.r);
  mpz_clear(g.inverse);
  mpz_clear(g.pow2_n);
  for (auto& e : g.regions) {
    mpz_clear(e.second->min_alpha);
    mpz_clear(e.second->modulus);
    delete e.second;
  }
  g.regions.clear();
  for (auto& e : g.bounds) {
    mpz_clear(e.second->z);
    delete e.second;
  }
  g.bounds.clear();
  if (g.sampler) qb200_diagk_destroy(g.sampler);
  g.sampler = NULL;
  g.valid = false;
}

void setup_for(const Diagonal_Parameters* p) {
  if (g.valid && g.m == p->m && g.sigma == p->sigma && g.l == p->l && 0 == mpz_cmp(g.d, p->d) &&
      0 == mpz_cmp(g.r, p->r))
    return;
  setup_clear();
  g.m = p->m;
  g.sigma = p->sigma;
  g.l = p->l;
  mpz_init_set(g.d, p->d);
  mpz_init_set(g.r, p->r);
  mpz_init(g.inverse);
  mpz_init(g.pow2_n);
  g.kappa_r = kappa(p->r);
  mpz_setbit(g.pow2_n, p->m + p->sigma);
  mpz_t t;
  mpz_init(t);
  mpz_fdiv_q_2exp(t, p->r, g.kappa_r);              // src/sample.cpp:377-379
  if (0 == mpz_invert(g.inverse, t, g.pow2_n)) {     // :384
    critical("tau_estimate_diagonal(): r / 2^kappa_r is not invertible modulo 2^(m + sigma).");
  }
  mpz_clear(t);
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
  g.j_limbs = qb200_diagk_j_limbs(g.sampler);
  g.valid = true;
}

// round(2^|log alpha|) as sample_alpha_from_region computes it (src/sample.cpp:97-124:
// mpfr_set_d, mpfr_exp2, mpfr_round at `precision` bits, mpfr_get_z), kept per (value, precision):
// the upper bound of one region is the lower bound of the next.
const mpz_t* bound_of(double abs_log_alpha, uint32_t precision) {
  const std::pair<double, uint32_t> key(abs_log_alpha, precision);
  auto it = g.bounds.find(key);
  if (it != g.bounds.end()) return &it->second->z;
  Setup::Z* b = new Setup::Z;
  mpz_init(b->z);
  mpfr_t x;
  mpfr_init2(x, precision);
  mpfr_set_d(x, abs_log_alpha, MPFR_RNDN);
  mpfr_exp2(x, x, MPFR_RNDN);
  mpfr_round(x, x);
  mpfr_get_z(b->z, x, MPFR_RNDN);
  mpfr_clear(x);
  g.bounds[key] = b;
  g_stats.bounds++;
  return &b->z;
}

// The integers sample_alpha_from_region (src/sample.cpp:77-128) derives from the region's
// bounds, once per region.
const Setup::Bounds* bounds_of(double min_log_alpha, double max_log_alpha) {
  const std::pair<double, double> key(min_log_alpha, max_log_alpha);
  auto it = g.regions.find(key);
  if (it != g.regions.end()) return it->second;
  if (sgn_d(min_log_alpha) != sgn_d(max_log_alpha)) {
    critical("sample_alpha_from_region(): Incompatible signs for min_log_alpha and max_log_alpha.");
  }
  if (abs_d(min_log_alpha) >= abs_d(max_log_alpha)) {
    critical("sample_alpha_from_region(): Incompatible absolute values for min_log_alpha and max_log_alpha.");
  }
  const uint32_t m = ceil(abs_d(max_log_alpha));   // :93
  const uint32_t precision = 3 * m;                // :95
  Setup::Bounds* b = new Setup::Bounds;
  mpz_init_set(b->min_alpha, *bound_of(abs_d(min_log_alpha), precision));             // :97-103, :118-120
  mpz_init(b->modulus);
  mpz_sub(b->modulus, *bound_of(abs_d(max_log_alpha), precision), b->min_alpha);      // :105-111, :122-128
  g.regions[key] = b;
  return b;
}

// diagonal_distribution_sample_j_eta (src/diagonal_distribution.cpp:410-472): the same draws
// from the same Random_State, the same j.
bool draw_j_eta(const Diagonal_Distribution* distribution, Random_State* rs, mpz_t j, int32_t* eta,
                mpz_t alpha_r, mpz_t t_r, mpz_t tmp) {
  double min_log_alpha_r, max_log_alpha_r;
  if (FALSE == diagonal_distribution_sample_region(distribution, rs, &min_log_alpha_r, &max_log_alpha_r, eta)) {
    return false;  // :366-383, :428-443
  }
  // sample_alpha_from_region(alpha_r, min, max, kappa_r, rs), src/sample.cpp:77-157
  const Setup::Bounds* b = bounds_of(min_log_alpha_r, max_log_alpha_r);
  random_generate_mpz(alpha_r, b->modulus, rs);     // :130
  mpz_add(alpha_r, b->min_alpha, alpha_r);          // :131
  if (g.kappa_r > 0) {                              // :133-144
    mpz_fdiv_r_2exp(tmp, alpha_r, g.kappa_r);
    mpz_sub(alpha_r, alpha_r, tmp);
  }
  if (sgn_d(min_log_alpha_r) == -1) mpz_neg(alpha_r, alpha_r);  // :147-149
  // sample_j_from_diagonal_alpha_r(j, alpha_r, parameters, rs), src/sample.cpp:352-410
  mpz_set_ui(t_r, 0);
  if (g.kappa_r > 0) {                              // :368-372
    mpz_set_ui(tmp, 0);
    mpz_setbit(tmp, g.kappa_r);
    random_generate_mpz(t_r, tmp, rs);
  }
  mpz_mul(j, g.inverse, alpha_r);                   // :386-387
  mpz_fdiv_q_2exp(j, j, g.kappa_r);                 // :389-393 (mpz_div floors)
  mpz_mul_2exp(tmp, t_r, g.m + g.sigma - g.kappa_r);  // :395-399
  mpz_add(j, j, tmp);                               // :400
  mpz_fdiv_r_2exp(j, j, g.m + g.sigma);             // :402-406 (mpz_mod: non-negative)
  return true;
}

struct Drawn {
  uint32_t count = 0;      // samples with (j, eta, pivot) drawn
  bool failed = false;     // the draw of sample `count` found no slice
};

void draw_all(const Diagonal_Distribution* distribution, Random_State* rs, uint32_t n, std::vector<uint32_t>& J,
              std::vector<int32_t>& eta, std::vector<long double>& pivot, Drawn* out) {
  mpz_t j, alpha_r, t_r, tmp;
  mpz_init(j);
  mpz_init(alpha_r);
  mpz_init(t_r);
  mpz_init(tmp);
  out->count = 0;
  out->failed = false;
  for (uint32_t i = 0; i < n; i++) {
    int32_t e = 0;
    if (!draw_j_eta(distribution, rs, j, &e, alpha_r, t_r, tmp)) {
      out->failed = true;
      break;
    }
    // sample_k_from_diagonal_j_eta (src/sample.cpp:648-675) draws the pivot next
    pivot[i] = random_generate_pivot_inclusive(rs);
    eta[i] = e;
    size_t cnt = 0;
    uint32_t* row = &J[(size_t)i * g.j_limbs];
    memset(row, 0, (size_t)g.j_limbs * 4);
    mpz_export(row, &cnt, -1, 4, 0, 0, j);
    out->count = i + 1;
  }
  mpz_clear(j);
  mpz_clear(alpha_r);
  mpz_clear(t_r);
  mpz_clear(tmp);
}

}  // namespace

bool tau_estimate_diagonal(const Diagonal_Distribution* const distribution, Random_State* const random_state,
                           const uint32_t n, const uint32_t delta_bound, const uint32_t eta_bound,
                           long double& tau) {
  if (0 == n) {  // the reference's loop does not run: result stays FALSE, nothing drawn
    tau = DBL_MAX;
    return false;
  }
  setup_for(&distribution->parameters);
  g_stats.calls++;
  const bool can_rewind = (NULL == random_state->random_device);
  Random_State entry;
  if (can_rewind) memcpy(&entry, random_state, sizeof entry);
  std::vector<uint32_t> J((size_t)n * g.j_limbs);
  std::vector<int32_t> eta(n);
  std::vector<long double> pivot(n);
  std::vector<double> x_hi(n), x_lo(n);
  std::vector<int32_t> status(n);
  Drawn drawn;
  double t0 = now_s();
  draw_all(distribution, random_state, n, J, eta, pivot, &drawn);
  g_stats.s_draw += now_s() - t0;
  g_stats.samples += drawn.count;
  if (drawn.count) {
    t0 = now_s();
    if (0 != qb200_diagk_sample(g.sampler, drawn.count, J.data(), eta.data(), pivot.data(), delta_bound, NULL,
                                x_hi.data(), x_lo.data(), NULL, status.data())) {
      critical("tau_estimate_diagonal(): %s", qb200_last_error());
    }
    g_stats.s_abi += now_s() - t0;
  }
  // the first sample at which the reference breaks (src/tau_estimate.cpp:163-188)
  uint32_t stop = drawn.count;
  for (uint32_t i = 0; i < drawn.count; i++) {
    if (status[i] == QB200_DIAGK_GAVE_UP) {
      critical("tau_estimate_diagonal(): sample_k_from_diagonal_j_eta_pivot(): gave up after 2^22 steps "
               "(delta_bound = %u).", delta_bound);
    }
    if (status[i] == QB200_DIAGK_OUT_OF_BOUNDS || abs_i(eta[i]) > eta_bound) {
      stop = i;
      break;
    }
  }
  if (stop < drawn.count) {
    if (can_rewind && (stop + 1 < drawn.count || drawn.failed)) {
      // the reference never drew samples stop + 1 ...: back to the entry state, samples 0 .. stop again
      memcpy(random_state, &entry, sizeof entry);
      Drawn again;
      draw_all(distribution, random_state, stop + 1, J, eta, pivot, &again);
      g_stats.replays++;
    }
    tau = DBL_MAX;
    return false;
  }
  if (drawn.failed) {
    tau = DBL_MAX;
    return false;
  }
  t0 = now_s();
  mpfr_t alpha, sum;
  mpfr_init2(alpha, PRECISION);
  mpfr_init2(sum, PRECISION);
  mpfr_set_ui(sum, 0, MPFR_RNDN);
  const long shift = (long)g.m + (long)g.sigma - (long)g.l;
  for (uint32_t i = 0; i < n; i++) {
    mpfr_set_d(alpha, x_hi[i], MPFR_RNDN);
    mpfr_add_d(alpha, alpha, x_lo[i], MPFR_RNDN);
    if (status[i] == QB200_DIAGK_OK_NEGATIVE_PHI) {  // alpha_phi = 2^(m+sigma-l) (x - 2^l)
      mpfr_t p;
      mpfr_init2(p, PRECISION);
      mpfr_set_ui_2exp(p, 1, (mpfr_exp_t)g.l, MPFR_RNDN);
      mpfr_sub(alpha, alpha, p, MPFR_RNDN);
      mpfr_clear(p);
    }
    mpfr_mul_2si(alpha, alpha, shift, MPFR_RNDN);
    mpfr_sqr(alpha, alpha, MPFR_RNDN);               // src/tau_estimate.cpp:185-186
    mpfr_add(sum, sum, alpha, MPFR_RNDN);
  }
  mpfr_div_ui(sum, sum, n, MPFR_RNDN);               // :194-201
  mpfr_log2(sum, sum, MPFR_RNDN);
  tau = mpfr_get_ld(sum, MPFR_RNDN) / 2.0f - (distribution->parameters.m + distribution->parameters.sigma -
                                              distribution->parameters.l);
  mpfr_clear(alpha);
  mpfr_clear(sum);
  g_stats.s_sum += now_s() - t0;
  return true;
}
