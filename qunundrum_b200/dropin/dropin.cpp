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
//     server and never integrates), overridable with QB200_DEVICE. Unless the user set
//     CUDA_VISIBLE_DEVICES, the rank makes ONLY its own GPU visible before the CUDA runtime
//     starts: on an 8-GPU node a process that sees all GPUs spends seconds initialising
//     devices it never uses.
//
// Batching without touching the client. The generator clients ask for
// one slice per call (src/main_generate_distribution.cpp:1129-1360 and the linear / diagonal
// twins), which leaves a B200 > 99 % idle. The coordinates a client can be asked for are no
// secret: they are what the reference's own enumerators list for these parameters
// (src/distribution_enumerator.cpp, linear_..., diagonal_...). On the first call for a set of
// parameters the drop-in therefore integrates the WHOLE list in one C-ABI call (thousands of
// slices per launch) into a host cache (pageable: pinning 440 MB costs more than the copy saves) and
// serves this and every later call from it; cells
// do not depend on the batch they are computed in (bit-identical, tested), so the caller cannot
// tell. The two-dimensional client's dimension heuristic (:1222-1294) re-computes some slices at
// 512 / 1024: the first call at a new dimension integrates, in one batch, every coordinate the
// heuristic can send there (its geometric conditions; where the lower-dimension total is already
// cached, its thresholds with a decade of slack). A coordinate outside the speculated set is
// simply computed on its own. QB200_PREFETCH=0 restores one slice per call. (The six functions are
// called from the client's main thread only -- the reference's thread pool runs the server's
// exports, src/main_generate_distribution.cpp:1070 -- and are not re-entrant.)
#include "common.h"
#include "diagonal_distribution_enumerator.h"
#include "diagonal_distribution_slice.h"
#include "diagonal_parameters.h"
#include "distribution_enumerator.h"
#include "distribution_slice.h"
#include "errors.h"
#include "linear_distribution_enumerator.h"
#include "linear_distribution_slice.h"
#include "parameters.h"

#include <gmp.h>

#include <dirent.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <utility>
#include <vector>

#include "qunundrum_b200.h"

namespace {

qb200_context* g_ctx = NULL;

// QB200_DROPIN_STATS=1: print, at exit, how many slices this process integrated and
// the wall time spent inside the C ABI calls (to tell GPU time from protocol time).
struct Stats {
  double seconds;        // inside the drop-in functions
  double abi_seconds;    // of which inside the C ABI
  unsigned long calls;   // drop-in calls
  unsigned long abi_calls;
  unsigned long abi_slices;
  unsigned long hits;
  double context_seconds;  // qb200_create: CUDA start-up of this process
  bool on;
} g_stats = {0.0, 0.0, 0, 0, 0, 0, 0.0, false};

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void print_stats() {
  if (!(g_stats.on && g_stats.calls)) return;
  char line[768];   // one write(): the ranks of a farm share stderr
  const int len = snprintf(
      line, sizeof line,
      "qunundrum_b200 drop-in: %lu slice calls, %.3f s inside the drop-in functions (%.1f us per "
      "call); %lu C-ABI calls for %lu slices (%.1f slices per call, %.3f s), %lu calls served from "
      "prefetched batches; CUDA context created in %.3f s\n",
      g_stats.calls, g_stats.seconds, 1e6 * g_stats.seconds / (double)g_stats.calls, g_stats.abi_calls,
      g_stats.abi_slices, g_stats.abi_calls ? (double)g_stats.abi_slices / (double)g_stats.abi_calls : 0.0,
      g_stats.abi_seconds, g_stats.hits, g_stats.context_seconds);
  if (len > 0) (void)!write(2, line, (size_t)(len < (int)sizeof line ? len : (int)sizeof line - 1));
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

struct TimedAbi {
  double t0;
  explicit TimedAbi(size_t slices) : t0(g_stats.on ? now_s() : 0.0) {
    g_stats.abi_calls++;
    g_stats.abi_slices += slices;
  }
  ~TimedAbi() {
    if (g_stats.on) g_stats.abi_seconds += now_s() - t0;
  }
};

int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : fallback;
}

bool prefetch_enabled() {
  static int on = -1;
  if (on < 0) on = env_int("QB200_PREFETCH", 1) != 0;
  return on != 0;
}

// GPUs of this node without starting the CUDA runtime (one directory per GPU).
int count_gpus_procfs() {
  DIR* d = opendir("/proc/driver/nvidia/gpus");
  if (!d) return 0;
  int n = 0;
  while (struct dirent* e = readdir(d))
    if (e->d_name[0] != '.') n++;
  closedir(d);
  return n;
}

qb200_context* context() {
  if (g_ctx) return g_ctx;
  int device = env_int("QB200_DEVICE", -1);
  int local = env_int("OMPI_COMM_WORLD_LOCAL_RANK", -1);
  if (local < 0) local = env_int("MPI_LOCALRANKID", -1);
  if (local < 0) local = env_int("SLURM_LOCALID", -1);
  if (local < 0) local = env_int("QB200_MINIMPI_RANK", 1);
  // Make only this rank's GPU visible -- before anything initialises CUDA in this process.
  const char* vis = getenv("CUDA_VISIBLE_DEVICES");
  if ((!vis || !*vis) && env_int("QB200_PIN_VISIBLE", 1)) {
    const int total = count_gpus_procfs();
    if (total > 1) {
      const int want = device >= 0 ? device % total : ((local - 1) % total + total) % total;
      char buf[16];
      snprintf(buf, sizeof buf, "%d", want);
      setenv("CUDA_VISIBLE_DEVICES", buf, 1);
      device = 0;
    }
  }
  const int n = qb200_device_count();
  if (n <= 0) critical("qunundrum_b200: no CUDA device (there is no CPU path).");
  if (device < 0) device = ((local - 1) % n + n) % n;
  if (device >= n) device %= n;
  const double t0 = now_s();
  if (0 != qb200_create(device, &g_ctx)) {
    critical("qunundrum_b200: %s", qb200_last_error());
  }
  g_stats.context_seconds = now_s() - t0;
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

// ---- the prefetched batches -----------------------------------------------------------------

// What a batch was computed for: a new distribution (or method) drops every batch.
struct Identity {
  uint32_t m, l, sigma;
  int what;  // method (2D) or kind (1D), and the Richardson switch
  std::vector<uint8_t> d, r;
  bool operator==(const Identity& o) const {
    return m == o.m && l == o.l && sigma == o.sigma && what == o.what && d == o.d && r == o.r;
  }
};

typedef std::pair<int32_t, int32_t> Coord;

struct Batch {
  uint32_t dimension;
  size_t per;            // cells per slice
  double* cells;         // n * per (pageable: pinning 440 MB costs a second on a virtual machine)
  std::vector<long double> tp, te;
  std::vector<uint32_t> flags;
  std::map<Coord, uint32_t> index;
  Batch() : dimension(0), per(0), cells(NULL) {}
};

struct Cache {
  bool valid;
  Identity id;
  std::vector<Coord> list;            // everything the enumerator hands out, in its order
  std::map<uint32_t, std::vector<Batch*> > batches;  // by dimension
  Cache() : valid(false) {}
  void clear() {
    for (std::map<uint32_t, std::vector<Batch*> >::iterator it = batches.begin(); it != batches.end(); ++it) {
      for (size_t k = 0; k < it->second.size(); k++) {
        free(it->second[k]->cells);
        delete it->second[k];
      }
    }
    batches.clear();
    list.clear();
    valid = false;
  }
};

Cache g_cache2d, g_cache1d;

Identity identity_of(const Exported& e, int what) {
  Identity id;
  id.m = e.p.m;
  id.l = e.p.l;
  id.sigma = e.p.sigma;
  id.what = what;
  id.d = e.d;
  id.r = e.r;
  return id;
}

int32_t iabs32(int32_t v) { return v < 0 ? -v : v; }

// The dimension heuristic of the two-dimensional client, as a predicate for speculation only
// (src/main_generate_distribution.cpp:1222-1294; `dimension` is the slice dimension, half the
// client's required_dimension). tp_known: the total at the client's initial dimension, if a
// batch holds it. A wrong guess costs time, never correctness.
bool may_be_asked_at(uint32_t dimension, uint32_t base, const Coord& c, int32_t m, bool tp_known,
                     long double tp) {
  const int32_t max_alpha = iabs32(c.first) > iabs32(c.second) ? iabs32(c.first) : iabs32(c.second);
  if (max_alpha > m + 10) return false;  // the client skips these (:1196-1212)
  if (dimension <= base) return true;
  const bool tail = iabs32(c.first - c.second) <= 1 && ((c.first < 0) == (c.second < 0));
  if (dimension == 2 * base) {  // required_dimension 512
    if (max_alpha < m) return false;
    if (tail && max_alpha < m + 3) return true;             // initial requirement (:1233-1238)
    if (tail) return false;                                  // goes to 1024 right away
    return !tp_known || tp >= 1e-8L;                         // (:1262-1269, slack of 10)
  }
  if (dimension == 4 * base) {  // required_dimension 1024
    if (tail && max_alpha >= m + 3) return true;             // (:1229-1232)
    if (max_alpha < m + 10) return false;
    return !tp_known || tp >= 1e-11L;                        // (:1271-1278)
  }
  return false;
}

Batch* compute_batch_2d(const Exported& e, int method, int richardson, uint32_t dimension,
                        const std::vector<Coord>& coords, const char* who) {
  Batch* b = new Batch;
  b->dimension = dimension;
  b->per = (size_t)dimension * dimension;
  const size_t n = coords.size();
  b->cells = (double*)malloc(n * b->per * sizeof(double) + 8);
  if (NULL == b->cells) critical("%s(): Failed to allocate memory.", who);
  b->tp.assign(n, 0);
  b->te.assign(n, 0);
  b->flags.assign(n, 0);
  std::vector<int32_t> ad(n), ar(n);
  for (size_t i = 0; i < n; i++) {
    ad[i] = coords[i].first;
    ar[i] = coords[i].second;
    b->index[coords[i]] = (uint32_t)i;
  }
  TimedAbi timed(n);
  if (0 != qb200_slice2d_compute(context(), &e.p, method, richardson, dimension, (uint32_t)n, ad.data(),
                                 ar.data(), b->cells, b->tp.data(), b->te.data(), b->flags.data())) {
    critical("%s(): %s", who, qb200_last_error());
  }
  return b;
}

const Batch* find_batch(const std::vector<Batch*>& bs, const Coord& c, uint32_t* at) {
  for (size_t k = 0; k < bs.size(); k++) {
    std::map<Coord, uint32_t>::const_iterator f = bs[k]->index.find(c);
    if (f != bs[k]->index.end()) {
      *at = f->second;
      return bs[k];
    }
  }
  return NULL;
}

// Cells, total error and flag bits of one two-dimensional slice: from a prefetched batch of its
// dimension, computing one first if no batch holds the coordinate.
void lookup_2d(const Exported& e, const Parameters* const parameters, int method, int richardson,
               uint32_t dimension, const Coord& c, const char* who, const double** cells,
               long double* te, uint32_t* flags, std::vector<double>* scratch) {
  Cache& C = g_cache2d;
  if (prefetch_enabled()) {
    const Identity id = identity_of(e, method * 2 + richardson);
    if (!C.valid || !(C.id == id)) {
      C.clear();
      C.id = id;
      Distribution_Enumerator en;
      distribution_enumerator_init(&en, parameters, TRUE);
      int32_t a, b;
      while (distribution_enumerator_next(&a, &b, &en)) C.list.push_back(Coord(a, b));
      distribution_enumerator_clear(&en);
      C.valid = true;
    }
    std::vector<Batch*>& here = C.batches[dimension];
    uint32_t at = 0;
    const Batch* hit = find_batch(here, c, &at);
    if (NULL == hit) {
      // Everything the client may still ask for at this dimension, in one call. `base` is the
      // client's initial dimension: the smallest seen so far -- or, on the very first call, a
      // guess from the coordinate (a tail coordinate starts at 512 / 1024 under the heuristic,
      // src/main_generate_distribution.cpp:1226-1240); a wrong guess costs one more batch.
      const int32_t m = (int32_t)parameters->m;
      uint32_t base = dimension;
      bool first_call = true;
      for (std::map<uint32_t, std::vector<Batch*> >::const_iterator it = C.batches.begin();
           it != C.batches.end(); ++it) {
        if (!it->second.empty()) {
          first_call = false;
          if (it->first < base) base = it->first;
        }
      }
      // this dimension was speculated on before and missed c although no smaller dimension has
      // been seen: the first-call guess ("dimension heuristic") was wrong, the client runs at a
      // fixed dimension -- take everything that is left
      const bool again = !here.empty() && base == dimension;
      if (first_call) {
        if (dimension % 4 == 0 && may_be_asked_at(dimension, dimension / 4, c, m, false, 0)) base = dimension / 4;
        else if (dimension % 2 == 0 && may_be_asked_at(dimension, dimension / 2, c, m, false, 0)) base = dimension / 2;
      }
      const std::vector<Batch*>* lower = NULL;
      if (base < dimension) {
        std::map<uint32_t, std::vector<Batch*> >::const_iterator it = C.batches.find(base);
        if (it != C.batches.end()) lower = &it->second;
      }
      std::vector<Coord> want;
      bool listed = false;
      for (size_t i = 0; i < C.list.size(); i++) {
        const Coord& q = C.list[i];
        uint32_t k = 0;
        if (find_batch(here, q, &k)) continue;
        bool known = false;
        long double tp = 0;
        if (lower && !again) {
          const Batch* lb = find_batch(*lower, q, &k);
          if (lb) {
            known = true;
            tp = lb->tp[k];
          }
        }
        if (q == c || may_be_asked_at(dimension, again ? dimension : base, q, m, known, tp)) {
          want.push_back(q);
          listed = listed || q == c;
        }
      }
      if (!listed) want.push_back(c);
      here.push_back(compute_batch_2d(e, method, richardson, dimension, want, who));
      hit = find_batch(here, c, &at);
    }
    g_stats.hits++;
    *cells = hit->cells + (size_t)at * hit->per;
    *te = hit->te[at];
    *flags = hit->flags[at];
    return;
  }
  // one slice per call (QB200_PREFETCH=0)
  scratch->resize((size_t)dimension * dimension);
  long double tp = 0;
  TimedAbi timed(1);
  if (0 != qb200_slice2d_compute(context(), &e.p, method, richardson, dimension, 1, &c.first, &c.second,
                                 scratch->data(), &tp, te, flags)) {
    critical("%s(): %s", who, qb200_last_error());
  }
  *cells = scratch->data();
}

void compute_2d(Distribution_Slice* const slice, const Parameters* const parameters,
                const Distribution_Slice_Compute_Method method, const int32_t min_log_alpha_d,
                const int32_t min_log_alpha_r, const int richardson, const char* who) {
  const uint32_t dimension = slice->dimension;
  Exported e;
  fill(e, parameters->m, parameters->l, 0, parameters->d, parameters->r);
  context();
  Timed timed;
  const double* cells = NULL;
  long double total_error = 0;
  uint32_t flags = 0;
  std::vector<double> scratch;
  lookup_2d(e, parameters, (int)method, richardson, dimension, Coord(min_log_alpha_d, min_log_alpha_r),
            who, &cells, &total_error, &flags, &scratch);
  // widen, then the reference's own summation order (both entry points walk alpha_d outermost:
  // src/distribution_slice_compute.cpp:397-407, ..._richardson.cpp:47-64)
  const size_t count = (size_t)dimension * dimension;
  long double* const out = slice->norm_matrix;
  for (size_t k = 0; k < count; k++) out[k] = cells[k];
  long double total_probability = 0;
  for (uint32_t i = 0; i < dimension; i++) {
    const double* col = cells + i;
    for (uint32_t j = 0; j < dimension; j++) total_probability += (long double)col[(size_t)dimension * j];
  }
  slice->total_probability = total_probability;
  slice->total_error = total_error;
  slice->min_log_alpha_d = min_log_alpha_d;
  slice->min_log_alpha_r = min_log_alpha_r;
  slice->flags &= ~(SLICE_FLAGS_MASK_METHOD);
  slice->flags |= flags;
}

// One-dimensional slices. list_1d: the coordinates (min_log_alpha, eta) the enumerator of this
// distribution hands out; the whole list is one C-ABI call (one kernel launch).
typedef void (*Lister)(const void* parameters, const Coord& asked, std::vector<Coord>* list);

void compute_1d(long double* const norm_vector, const uint32_t dimension, const Exported& e,
                const int kind, Lister lister, const void* lister_parameters, const Coord& c,
                const int richardson, long double* total_probability, uint32_t* flags,
                const char* who) {
  context();
  Timed timed;
  Cache& C = g_cache1d;
  const double* cells = NULL;
  std::vector<double> scratch;
  if (prefetch_enabled()) {
    const Identity id = identity_of(e, kind * 2 + richardson);
    if (!C.valid || !(C.id == id)) {
      C.clear();
      C.id = id;
      lister(lister_parameters, c, &C.list);
      C.valid = true;
    }
    std::vector<Batch*>& here = C.batches[dimension];
    uint32_t at = 0;
    const Batch* hit = find_batch(here, c, &at);
    if (NULL == hit) {
      std::vector<Coord> want;
      bool listed = false;
      for (size_t i = 0; i < C.list.size(); i++) {
        uint32_t k = 0;
        if (find_batch(here, C.list[i], &k)) continue;
        want.push_back(C.list[i]);
        listed = listed || C.list[i] == c;
      }
      if (!listed) want.push_back(c);
      Batch* b = new Batch;
      b->dimension = dimension;
      b->per = dimension;
      const size_t n = want.size();
      b->cells = (double*)malloc(n * b->per * sizeof(double) + 8);
      if (NULL == b->cells) critical("%s(): Failed to allocate memory.", who);
      b->tp.assign(n, 0);
      b->flags.assign(n, 0);
      std::vector<int32_t> a(n), eta(n);
      for (size_t i = 0; i < n; i++) {
        a[i] = want[i].first;
        eta[i] = want[i].second;
        b->index[want[i]] = (uint32_t)i;
      }
      {
        TimedAbi timed_abi(n);
        if (0 != qb200_slice1d_compute(context(), &e.p, kind, richardson, dimension, (uint32_t)n, a.data(),
                                       eta.data(), b->cells, b->tp.data(), b->flags.data())) {
          critical("%s(): %s", who, qb200_last_error());
        }
      }
      here.push_back(b);
      hit = find_batch(here, c, &at);
    }
    g_stats.hits++;
    cells = hit->cells + (size_t)at * hit->per;
    *flags = hit->flags[at];
  }
  if (NULL == cells) {
    scratch.resize(dimension);
    long double tp = 0;
    TimedAbi timed_abi(1);
    if (0 != qb200_slice1d_compute(context(), &e.p, kind, richardson, dimension, 1, &c.first, &c.second,
                                   scratch.data(), &tp, flags)) {
      critical("%s(): %s", who, qb200_last_error());
    }
    cells = scratch.data();
  }
  *total_probability = 0;
  for (uint32_t i = 0; i < dimension; i++) {
    norm_vector[i] = cells[i];
    *total_probability += norm_vector[i];
  }
}

// What the generators' servers enumerate (src/main_generate_linear_distribution.cpp:669-672,
// src/main_generate_diagonal_distribution.cpp:681-688: mirrored = TRUE).
void list_linear(const void* parameters, const Coord&, std::vector<Coord>* list) {
  Linear_Distribution_Enumerator en;
  linear_distribution_enumerator_init(&en, (const Parameters*)parameters, TRUE);
  int32_t a;
  while (linear_distribution_enumerator_next(&a, &en)) list->push_back(Coord(a, 0));
  linear_distribution_enumerator_clear(&en);
}

void list_diagonal(const void* parameters, const Coord& asked, std::vector<Coord>* list) {
  const Diagonal_Parameters* const p = (const Diagonal_Parameters*)parameters;
  // every eta up to the larger of the distribution's bound and this call's |eta|
  uint32_t bound = p->eta_bound;
  if ((uint32_t)iabs32(asked.second) > bound) bound = (uint32_t)iabs32(asked.second);
  if (bound > 4096) return;
  Diagonal_Distribution_Enumerator en;
  diagonal_distribution_enumerator_init(&en, p, bound, TRUE);
  int32_t a, h;
  while (diagonal_distribution_enumerator_next(&a, &h, &en)) list->push_back(Coord(a, h));
  diagonal_distribution_enumerator_clear(&en);
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
  compute_1d(slice->norm_vector, slice->dimension, e,
             (LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D == target) ? QB200_KIND_LINEAR_D
                                                                    : QB200_KIND_LINEAR_R,
             list_linear, parameters, Coord(min_log_alpha, 0), richardson,
             &slice->total_probability, &flags, who);
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
  compute_1d(slice->norm_vector, slice->dimension, e, QB200_KIND_DIAGONAL, list_diagonal, parameters,
             Coord(min_log_alpha_r, eta), richardson, &slice->total_probability, &flags, who);
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
