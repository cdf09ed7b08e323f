// dropin_tau.cpp -- reference-side forwarding TU for tau estimation.
//
// What a maintainer of ekera/qunundrum adds to src/. It defines, with the reference's own
// signatures (src/tau_estimate.h:51-117),
//
//   bool tau_estimate(const Distribution *, Random_State *, uint32_t n, long double &tau_d, long double &tau_r)
//   bool tau_estimate_linear(const Linear_Distribution *, Random_State *, uint32_t n, long double &tau)
//
// over qb200_sampler_tau_estimate (include/qunundrum_b200.h). The reference's tau_estimate.cpp
// stays in the build for tau_estimate_diagonal, compiled with
//   -Dtau_estimate=tau_estimate_cpu_unused -Dtau_estimate_linear=tau_estimate_linear_cpu_unused
// (two renames on the command line, no source change), so that the
// callers -- estimate_runs_distribution / estimate_runs_linear_distribution
// (src/main_estimate_runs_distribution.cpp:292-309) -- link against the functions below.
//
// Semantics kept:
//   * the random stream. The reference draws 8 bytes per pivot / fraction from the caller's
//     Random_State (src/random.c:116-156). The words are drawn here from the SAME Random_State
//     with the reference's own random_generate(), in the same order, and the library consumes
//     them exactly as the reference would (an estimate stops reading at its first out-of-bounds
//     sample). So for a given generator state the results are the reference's.
//   * batching. The callers ask for one estimate per call, 1000 calls per job with the same
//     arguments (TAU_CHUNK_SIZE, src/executables_estimate_runs_distribution.h:23). The first
//     call of a run computes QB200_TAU_BATCH (default 1000) estimates in one batch and the
//     following calls with the same (distribution, random state, n) return them in order; words
//     drawn but not consumed (estimates that failed early) stay queued in front of the
//     generator, so the logical stream -- queue first, then the Random_State -- is consumed
//     exactly as the reference consumes its generator (the Random_State itself runs ahead by the
//     length of the queue). If the caller changes n, the distribution or the generator in mid-batch, the
//     unused estimates are dropped: the results remain correct samples, but the stream position
//     then differs from the reference's (never the case in the reference's executables with the
//     default batch).
//   * the generator itself. The reference's Keccak sponge (src/keccak_random.c:96-125: one
//     keccak_f per 168 bytes, bytes taken one at a time, most significant first) delivers
//     48 MB/s, which would cap a client rank at 1.5e6 samples/s. draw_words() below produces the
//     SAME bytes from the SAME Keccak_Random_State (its lanes and offset are advanced exactly
//     as the reference would advance them) with an unrolled permutation and whole-lane
//     extraction; tests compare megabytes of both streams. QB200_TAU_REFERENCE_RNG=1 switches
//     back to random_generate(); a device-backed Random_State (random_init_device) and a state
//     left in mid-lane by other draws always go through random_generate().
//   * errors are fatal: critical() (src/errors.c).
#include "common.h"
#include "distribution.h"
#include "distribution_slice.h"
#include "errors.h"
#include "linear_distribution.h"
#include "linear_distribution_slice.h"
#include "keccak.h"
#include "keccak_random.h"
#include "random.h"
#include "tau_estimate.h"

#include <float.h>
#include <stdio.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "qunundrum_b200.h"

namespace {

qb200_context* g_ctx = NULL;

// QB200_DROPIN_STATS=1: at exit, where the time of this rank's estimates went.
struct Stats {
  bool on = false;
  unsigned long calls = 0, batches = 0, creates = 0;
  double s_draw = 0, s_abi = 0, s_create = 0;
} g_stats;

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void print_stats() {
  if (g_stats.on && g_stats.calls)
    fprintf(stderr,
            "qunundrum_b200 tau drop-in: %lu estimates in %lu batches; %.3f s drawing the random words, "
            "%.3f s inside qb200_sampler_tau_estimate, %lu sampler set-ups in %.3f s\n",
            g_stats.calls, g_stats.batches, g_stats.s_draw, g_stats.s_abi, g_stats.creates, g_stats.s_create);
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

uint64_t mix(uint64_t h, const void* p, size_t n) {  // FNV-1a
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

// One slice list as the sampler wants it, for either container.
struct SliceList {
  int dims;
  uint32_t m;
  uint32_t count;
  long double total;
  std::vector<uint32_t> dimension;
  std::vector<int32_t> c0, c1;
  std::vector<const long double*> cells;
  std::vector<long double> totals;
};

void describe(const Distribution* d, SliceList* out) {
  out->dims = QB200_SAMPLER_2D;
  out->m = d->parameters.m;
  out->count = d->count;
  out->total = d->total_probability;
  for (uint32_t i = 0; i < d->count; i++) {
    const Distribution_Slice* s = d->slices[i];
    out->dimension.push_back(s->dimension);
    out->c0.push_back(s->min_log_alpha_d);
    out->c1.push_back(s->min_log_alpha_r);
    out->cells.push_back(s->norm_matrix);
    out->totals.push_back(s->total_probability);
  }
}

void describe(const Linear_Distribution* d, SliceList* out) {
  out->dims = QB200_SAMPLER_LINEAR;
  out->m = d->parameters.m;
  out->count = d->count;
  out->total = d->total_probability;
  for (uint32_t i = 0; i < d->count; i++) {
    const Linear_Distribution_Slice* s = d->slices[i];
    out->dimension.push_back(s->dimension);
    out->c0.push_back(s->min_log_alpha);
    out->c1.push_back(0);
    out->cells.push_back(s->norm_vector);
    out->totals.push_back(s->total_probability);
  }
}

// Identity of a distribution's content as far as it can be told cheaply: the containers may
// be cleared and re-filled at the same address between two calls.
uint64_t fingerprint(const SliceList& l) {
  uint64_t h = 0xcbf29ce484222325ull;
  h = mix(h, &l.dims, sizeof l.dims);
  h = mix(h, &l.m, sizeof l.m);
  h = mix(h, &l.count, sizeof l.count);
  h = mix(h, &l.total, 10);
  for (uint32_t i = 0; i < l.count; i++) {
    h = mix(h, &l.dimension[i], 4);
    h = mix(h, &l.c0[i], 4);
    h = mix(h, &l.c1[i], 4);
    h = mix(h, &l.cells[i], sizeof(void*));
    h = mix(h, &l.totals[i], 10);
    const size_t nc = l.dims == 2 ? (size_t)l.dimension[i] * l.dimension[i] : l.dimension[i];
    if (nc) {
      h = mix(h, &l.cells[i][0], 10);
      h = mix(h, &l.cells[i][nc / 2], 10);
      h = mix(h, &l.cells[i][nc - 1], 10);
    }
  }
  return h;
}

// ---- the reference's random stream, faster ----------------------------------------------------

inline uint64_t rotl64(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }

// Keccak-f[1600] (FIPS 202 section 3), lanes[x + 5 y]; the same permutation as the reference's
// keccak_f (src/keccak.c:47-108), written with the rho offsets and pi destinations as tables
// so that the compiler unrolls each step.
void keccak_f1600(uint64_t* a) {
  static const uint64_t rc[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const unsigned rho[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                   25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};
  uint64_t b[25], c[5], d[5];
  for (int round = 0; round < 24; round++) {
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const int x = i % 5, y = i / 5;
      b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[i] ^ d[x], rho[i]);
    }
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const int x = i % 5, j = i - x;
      a[i] = b[i] ^ (~b[j + (x + 1) % 5] & b[j + (x + 2) % 5]);
    }
    a[0] ^= rc[round];
  }
}

// n consecutive 8-byte draws of random_generate() (src/random.c:88-114), as the little-endian
// words random_generate_pivot_*() would read them.
void draw_words(Random_State* rs, uint64_t* dst, size_t n, bool use_reference) {
  random_generate(dst, 0, rs);  // the reference's initialisation check (canary), no bytes
  Keccak_Random_State* const k = &rs->keccak_state;
  if (use_reference || NULL != rs->random_device || 0 != (k->offset % 8) ||
      k->offset > 8 * KECCAK_LANE_COUNT) {
    while (n) {  // random_generate() takes a 32-bit byte count
      const size_t t = n < (size_t)(1u << 27) ? n : (size_t)(1u << 27);
      random_generate(dst, (uint32_t)(8 * t), rs);
      dst += t;
      n -= t;
    }
    return;
  }
  uint32_t offset = k->offset;
  while (n) {
    if (offset >= 8 * KECCAK_LANE_COUNT) {  // src/keccak_random.c:108-111
      keccak_f1600(k->lanes);
      offset = KECCAK_RANDOM_SEED_LENGTH;
    }
    size_t take = (8 * KECCAK_LANE_COUNT - offset) / 8;
    if (take > n) take = n;
    const uint64_t* lane = k->lanes + offset / 8;
    for (size_t i = 0; i < take; i++) dst[i] = __builtin_bswap64(lane[i]);  // most significant byte first
    dst += take;
    n -= take;
    offset += (uint32_t)(8 * take);
  }
  k->offset = offset;
}

struct State {
  const void* distribution = NULL;
  const void* slices = NULL;
  uint32_t count = 0;
  long double total = 0;
  uint64_t print = 0;
  qb200_sampler* sampler = NULL;
  // results of the current batch
  Random_State* rs = NULL;
  uint32_t n = 0;
  std::vector<long double> tau0, tau1;
  std::vector<uint8_t> ok;
  size_t next = 0;
  // words drawn from rs and not consumed yet
  std::vector<uint64_t> fifo;
  Random_State* fifo_rs = NULL;
} g;

template <class Dist>
bool estimate(const Dist* distribution, Random_State* rs, uint32_t n, long double* tau0,
              long double* tau1, const char* who) {
  if (0 == n) {  // the reference's loop does not run: failure, nothing drawn
    *tau0 = DBL_MAX;
    if (tau1) *tau1 = DBL_MAX;
    return false;
  }
  g_stats.calls++;
  if (g.next < g.ok.size() && g.distribution == (const void*)distribution && g.rs == rs && g.n == n &&
      g.count == distribution->count && g.slices == (const void*)distribution->slices &&
      0 == memcmp(&g.total, &distribution->total_probability, 10)) {
    // (inside a batch only the container's identity is re-checked -- address, slice count, slice
    // array and total probability; the full content fingerprint is taken once per batch)
    *tau0 = g.tau0[g.next];
    if (tau1) *tau1 = g.tau1[g.next];
    return g.ok[g.next++] != 0;
  }
  SliceList l;
  describe(distribution, &l);
  if (0 == l.count) {  // the reference's walk selects nothing: one draw, failure
    uint64_t w;
    random_generate(&w, sizeof w, rs);
    *tau0 = DBL_MAX;
    if (tau1) *tau1 = DBL_MAX;
    return false;
  }
  const uint64_t fp = fingerprint(l);
  if (NULL == g.sampler || g.distribution != (const void*)distribution || g.print != fp) {
    if (g.sampler) qb200_sampler_destroy(g.sampler);
    g.sampler = NULL;
    const double t_create = now_s();
    if (0 != qb200_sampler_create(context(), l.dims, l.m, l.count, l.dimension.data(), l.c0.data(),
                                  l.c1.data(), l.cells.data(), l.totals.data(), l.total, &g.sampler)) {
      critical("%s(): %s", who, qb200_last_error());
    }
    g.distribution = (const void*)distribution;
    g.print = fp;
    g_stats.creates++;
    g_stats.s_create += now_s() - t_create;
  }
  if (g.fifo_rs != rs) {
    g.fifo.clear();
    g.fifo_rs = rs;
  }
  const uint32_t batch = (uint32_t)(env_int("QB200_TAU_BATCH", 1000) > 0 ? env_int("QB200_TAU_BATCH", 1000) : 1);
  const size_t wps = qb200_sampler_words_per_sample(g.sampler);
  const size_t need = (size_t)batch * n * wps;
  if (g.fifo.size() < need) {
    const size_t have = g.fifo.size();
    g.fifo.resize(need);
    const double t_draw = now_s();
    draw_words(rs, &g.fifo[have], need - have, env_int("QB200_TAU_REFERENCE_RNG", 0) != 0);
    g_stats.s_draw += now_s() - t_draw;
  }
  g.tau0.assign(batch, 0);
  g.tau1.assign(batch, 0);
  g.ok.assign(batch, 0);
  size_t used = 0;
  uint32_t done = 0;
  const double t_abi = now_s();
  if (0 != qb200_sampler_tau_estimate(g.sampler, n, batch, g.fifo.data(), g.fifo.size(), &used, &done,
                                      g.tau0.data(), g.tau1.data(), g.ok.data())) {
    critical("%s(): %s", who, qb200_last_error());
  }
  g_stats.s_abi += now_s() - t_abi;
  g_stats.batches++;
  if (done != batch) critical("%s(): internal error: the word queue ran dry.", who);
  g.fifo.erase(g.fifo.begin(), g.fifo.begin() + (long)used);
  g.rs = rs;
  g.n = n;
  g.next = 0;
  g.count = distribution->count;
  g.slices = (const void*)distribution->slices;
  g.total = distribution->total_probability;
  *tau0 = g.tau0[0];
  if (tau1) *tau1 = g.tau1[0];
  return g.ok[g.next++] != 0;
}

}  // namespace

// Test hook: n draws through either path (tests compare the two streams and the states they leave).
extern "C" void qb200_dropin_tau_draw(Random_State* random_state, uint64_t* dst, size_t n,
                                      int use_reference) {
  draw_words(random_state, dst, n, use_reference != 0);
}

bool tau_estimate(const Distribution* const distribution, Random_State* const random_state,
                  const uint32_t n, long double& tau_d, long double& tau_r) {
  return estimate(distribution, random_state, n, &tau_d, &tau_r, "tau_estimate");
}

bool tau_estimate_linear(const Linear_Distribution* const distribution,
                         Random_State* const random_state, const uint32_t n, long double& tau) {
  return estimate(distribution, random_state, n, &tau, (long double*)NULL, "tau_estimate_linear");
}
