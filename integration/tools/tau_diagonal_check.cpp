// tau_diagonal_check.cpp -- TEST DRIVER: the drop-in tau_estimate_diagonal
// (qunundrum_b200/dropin/dropin_tau_diagonal.cpp) against the reference's own
// (src/tau_estimate.cpp:135-210, linked in under the name tau_estimate_diagonal_cpu_unused) in ONE
// process, on the same distribution, from identically seeded generators:
//
//   tau_diagonal_check <distribution> <n> <estimates> <delta_bound> <eta_bound> [<seed> [<stride>]]
//
// For every estimate: the same success flag, tau equal to within 2^-58, and -- the stream -- the
// two Random_States bit-identical after every <stride>-th call (default 1; run with
// QB200_TAU_BATCH=<stride>: the drop-in's generator runs ahead to the end of its batch). Prints one JSON line with the timings of both
// and exits 0 iff everything agrees. Built by integration/build.py (gpu flavour) and, against the
// CPU stand-in of the library, by tests/hostsim/shim_flavour.py.
#include "common.h"
#include "diagonal_distribution.h"
#include "keccak_random.h"
#include "random.h"
#include "tau_estimate.h"

#include <mpfr.h>

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

bool tau_estimate_diagonal_cpu_unused(const Diagonal_Distribution* const distribution,
                                      Random_State* const random_state, const uint32_t n,
                                      const uint32_t delta_bound, const uint32_t eta_bound, long double& tau);

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s <distribution> <n> <estimates> <delta_bound> <eta_bound> [<seed>]\n", argv[0]);
    return 2;
  }
  mpfr_set_default_prec(PRECISION);
  const uint32_t n = (uint32_t)atoi(argv[2]), count = (uint32_t)atoi(argv[3]);
  const uint32_t delta_bound = (uint32_t)strtoul(argv[4], NULL, 10), eta_bound = (uint32_t)atoi(argv[5]);
  const unsigned seed_id = argc > 6 ? (unsigned)atoi(argv[6]) : 1u;
  const uint32_t stride = argc > 7 ? (uint32_t)atoi(argv[7]) : 1u;
  FILE* f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  Diagonal_Distribution dist;
  diagonal_distribution_init_import(&dist, f);
  fclose(f);
  uint8_t seed[KECCAK_RANDOM_SEED_LENGTH];
  for (unsigned i = 0; i < KECCAK_RANDOM_SEED_LENGTH; i++) seed[i] = (uint8_t)(17 * i + 101 * seed_id + 3);
  Random_State a, b;
  random_init(&a);
  keccak_random_init_seed(&a.keccak_state, seed);
  random_init(&b);
  keccak_random_init_seed(&b.keccak_state, seed);
  // warm-up of the drop-in outside the timing (CUDA context, sampler set-up) on a third state
  {
    Random_State w;
    random_init(&w);
    keccak_random_init_seed(&w.keccak_state, seed);
    long double t;
    tau_estimate_diagonal(&dist, &w, n, delta_bound, eta_bound, t);
    // (with a batch, the rest of the warm-up batch is dropped when the state pointer changes)
    random_close(&w);
  }
  unsigned mismatched_flag = 0, mismatched_tau = 0, mismatched_state = 0, failed = 0;
  long double worst = 0, tau_sum = 0;
  double s_ref = 0, s_new = 0;
  for (uint32_t i = 0; i < count; i++) {
    long double ta = 0, tb = 0;
    double t0 = now_s();
    const bool ra = tau_estimate_diagonal_cpu_unused(&dist, &a, n, delta_bound, eta_bound, ta);
    s_ref += now_s() - t0;
    t0 = now_s();
    const bool rb = tau_estimate_diagonal(&dist, &b, n, delta_bound, eta_bound, tb);
    s_new += now_s() - t0;
    if (ra != rb) mismatched_flag++;
    if (!ra) failed++;
    if (ra && rb) {
      const long double e = fabsl(ta - tb);
      if (e > worst) worst = e;
      if (!(e <= ldexpl(1.0L, -58) * (1 + fabsl(ta)))) mismatched_tau++;
      tau_sum += ta;
    } else if (ta != tb) {
      mismatched_tau++;
    }
    if ((i + 1) % stride == 0 && 0 != memcmp(&a.keccak_state, &b.keccak_state, sizeof a.keccak_state)) {
      mismatched_state++;
    }
  }
  const bool ok = !mismatched_flag && !mismatched_tau && !mismatched_state;
  printf("{\"ok\": %s, \"estimates\": %u, \"n\": %u, \"failed_estimates\": %u, \"mismatched_flags\": %u, "
         "\"mismatched_taus\": %u, \"mismatched_states\": %u, \"worst_tau_difference\": %.3Le, "
         "\"mean_tau\": %.6Lf, \"reference_s\": %.4f, \"dropin_s\": %.4f, \"m\": %u, \"sigma\": %u, \"l\": %u, "
         "\"slices\": %u}\n",
         ok ? "true" : "false", count, n, failed, mismatched_flag, mismatched_tau, mismatched_state, worst,
         count > failed ? tau_sum / (count - failed) : 0.0L, s_ref, s_new, dist.parameters.m, dist.parameters.sigma,
         dist.parameters.l, dist.count);
  random_close(&a);
  random_close(&b);
  diagonal_distribution_clear(&dist);
  return ok ? 0 : 1;
}
