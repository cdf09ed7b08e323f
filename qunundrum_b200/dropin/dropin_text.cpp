// dropin_text.cpp -- the reference-side forwarding translation unit for the slice
// TEXT FORMAT.
//
// This file is what a maintainer of ekera/qunundrum adds to the reference's src/
// directory IN PLACE OF the three translation units
//
//   distribution_slice_import_export.cpp
//   linear_distribution_slice_import_export.cpp
//   diagonal_distribution_slice_import_export.cpp
//
// It defines the same nine functions (*_slice_import, *_slice_init_import,
// *_slice_export) with the same C++ signatures. The header fields keep going
// through fscanf / fprintf exactly as in the reference; the cells and the total
// error -- one "%.24Lg\n" / "%Lg\n" per value, 65,537 per stored two-dimensional
// slice -- go through the C ABI of libqunundrum_b200.so (qb200_text_format_ld,
// qb200_text_parse_ld), which produces the very same bytes / bits on the GPU.
// distribution_export(), distribution_import(), the .txt layout, the MPI
// protocol and every caller are untouched.
//
// Conventions kept: callers own the slice and the FILE; the importers re-init
// the slice when the stored dimension differs; total_probability is the running
// long double sum in index order (src/distribution_slice_import_export.cpp:38-46);
// errors are fatal (critical(), src/errors.c). The importers need a seekable
// FILE (a regular file, as every caller in the reference passes): the numbers
// are read in one block and the position is then set to where fscanf would have
// left it.
#include "common.h"
#include "diagonal_distribution_slice.h"
#include "distribution_slice.h"
#include "errors.h"
#include "linear_distribution_slice.h"

#include <dirent.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <mutex>
#include <thread>
#include <vector>

#include "qunundrum_b200.h"

// dropin_collapse.cpp, when linked in: the text of a slice whose cells already lie in device
// memory (uploaded once for the collapse to the marginals), formatted ahead a few dozen slices per
// synchronisation. Weak: the text drop-in also works on its own.
bool qb200_dropin_resident_text(const long double* cells, size_t n, long double tail, const char** text,
                                size_t* len) __attribute__((weak));

namespace {

qb200_context* g_text_ctx = NULL;
std::mutex g_text_mutex;  // the server exports from worker threads (main_server_export_distribution)

// QB200_DROPIN_STATS=1: at exit, the time spent in the C ABI calls and in fwrite / fread.
struct TextStats {
  double abi_s, io_s;
  unsigned long exports, imports;
  bool on;
} g_tstats = {0.0, 0.0, 0, 0, false};

double tnow() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void print_text_stats() {
  if (g_tstats.on && (g_tstats.exports || g_tstats.imports))
    {
    char line[768];
    const int len = snprintf(line, sizeof line,
            "qunundrum_b200 text drop-in: %lu slice exports, %lu slice imports, %.3f s inside the C ABI, "
            "%.3f s in fwrite / fread\n",
            g_tstats.exports, g_tstats.imports, g_tstats.abi_s, g_tstats.io_s);
    if (len > 0) (void)!write(2, line, (size_t)(len < (int)sizeof line ? len : (int)sizeof line - 1));
  }
}

qb200_context* text_context_create();
int text_device();
std::once_flag g_text_once;

qb200_context* text_context() {
  std::call_once(g_text_once, [] { g_text_ctx = text_context_create(); });
  return g_text_ctx;
}

// The SERVER of a generator run (rank 0 of an MPI job) needs its CUDA context seconds after it
// started -- when the clients are done and the distribution is collapsed and exported -- and
// creating one takes 0.3 ... 2 s. Start it in the background when the process starts.
// QB200_EAGER_INIT=0 turns this off.
struct EagerServerContext {
  EagerServerContext() {
    const char* off = getenv("QB200_EAGER_INIT");
    if (off && *off == '0') return;
    const char* names[] = {"QB200_MINIMPI_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK"};
    const char* sizes[] = {"QB200_MINIMPI_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", NULL};
    for (int i = 0; i < 4; i++) {
      const char* r = getenv(names[i]);
      if (!r || !*r) continue;
      const char* n = sizes[i] ? getenv(sizes[i]) : NULL;
      if (atoi(r) == 0 && (!n || atoi(n) > 1)) {
        text_device();  // the environment is settled on this thread, before the other one starts
        // ... after the clients have created theirs (context creations of one node queue up in
        // the driver, and the clients' are on the critical path): QB200_EAGER_DELAY_MS, default 1500
        const char* dl = getenv("QB200_EAGER_DELAY_MS");
        const int delay_ms = (dl && *dl) ? atoi(dl) : 1500;
        std::thread([delay_ms] {
          if (delay_ms > 0) {
            struct timespec ts = {delay_ms / 1000, (long)(delay_ms % 1000) * 1000000L};
            nanosleep(&ts, NULL);
          }
          text_context();
        }).detach();
      }
      return;
    }
  }
} g_eager_server_context;

// Which device, and -- before anything starts CUDA in this process -- only that one visible (as in
// dropin.cpp: seconds of start-up per process on an 8-GPU node otherwise). Environment work:
// called from the main thread (static initialisation or the first use), once.
int text_device() {
  static int device = -1;
  if (device >= 0) return device;
  const char* v = getenv("QB200_TEXT_DEVICE");
  if (!v || !*v) v = getenv("QB200_DEVICE");
  device = (v && *v) ? atoi(v) : 0;
  if (device < 0) device = 0;
  const char* vis = getenv("CUDA_VISIBLE_DEVICES");
  const char* pin = getenv("QB200_PIN_VISIBLE");
  if ((!vis || !*vis) && !(pin && *pin == '0')) {
    int total = 0;
    if (DIR* d = opendir("/proc/driver/nvidia/gpus")) {
      while (struct dirent* e = readdir(d))
        if (e->d_name[0] != '.') total++;
      closedir(d);
    }
    if (total > 1) {
      char buf[16];
      snprintf(buf, sizeof buf, "%d", device % total);
      setenv("CUDA_VISIBLE_DEVICES", buf, 1);
      device = 0;
    }
  }
  return device;
}

qb200_context* text_context_create() {
  const char* st = getenv("QB200_DROPIN_STATS");
  if (st && *st && *st != '0') {
    g_tstats.on = true;
    atexit(print_text_stats);
  }
  int device = text_device();
  const int n = qb200_device_count();
  if (n <= 0) critical("qunundrum_b200: no CUDA device (there is no CPU path).");
  device = ((device % n) + n) % n;
  qb200_context* ctx = NULL;
  if (0 != qb200_create(device, &ctx)) critical("qunundrum_b200: %s", qb200_last_error());
  return ctx;
}

// The exporter's value loop: n cells, then the total error.
void export_values(FILE* const file, const long double* const values, const size_t n,
                   const long double total_error, const char* who) {
  std::lock_guard<std::mutex> lock(g_text_mutex);
  const char* text = NULL;
  size_t len = 0;
  qb200_context* const ctx = text_context();
  const double t0 = g_tstats.on ? tnow() : 0.0;
  if (!(qb200_dropin_resident_text && qb200_dropin_resident_text(values, n, total_error, &text, &len)) &&
      0 != qb200_text_format_ld(ctx, values, n, &total_error, &text, &len)) {
    critical("%s(): %s", who, qb200_last_error());
  }
  const double t1 = g_tstats.on ? tnow() : 0.0;
  if (len != fwrite(text, 1, len, file)) {
    critical("%s(): Failed to write to the file.", who);
  }
  if (g_tstats.on) {
    g_tstats.abi_s += t1 - t0;
    g_tstats.io_s += tnow() - t1;
    g_tstats.exports++;
  }
}

// The importer's value loop: n cells into values (summed into *total_probability
// in index order), then the total error.
void import_values(FILE* const file, long double* const values, const size_t n,
                   long double* const total_probability, long double* const total_error,
                   const char* who) {
  std::lock_guard<std::mutex> lock(g_text_mutex);
  const long pos = ftell(file);
  if (pos < 0) critical("%s(): The file is not seekable.", who);
  std::vector<char> buf;
  std::vector<long double> parsed(n + 1);
  size_t want = 40 * (n + 1) + 64, used = 0;
  int malformed_attempts = 0;
  qb200_context* const ctx = text_context();
  for (;;) {
    buf.resize(want);
    const double t0 = g_tstats.on ? tnow() : 0.0;
    const size_t got = fread(buf.data(), 1, want, file);
    const double t1 = g_tstats.on ? tnow() : 0.0;
    const int rc = qb200_text_parse_ld(ctx, buf.data(), got, n + 1, parsed.data(), &used);
    if (g_tstats.on) {
      g_tstats.io_s += t1 - t0;
      g_tstats.abi_s += tnow() - t1;
    }
    // A block that is not the rest of the file may end inside the last number: the result
    // only counts if something follows that number in the block (used < got) or the block
    // reached the end of the file (got < want).
    if (0 == rc && (used < got || got < want)) break;
    // Any failure on a block that is not the rest of the file may come from a number cut by the
    // block end (too few numbers, or a malformed stump such as "1.5e-"): read more. A file that
    // really is malformed fails once the block reaches the end of the file.
    // (A stump is cured by the next, twice as large block; two more attempts are allowed.)
    if (0 != rc && -20 != rc) malformed_attempts++;
    if (got == want && malformed_attempts <= 2) {
      if (0 != fseek(file, pos, SEEK_SET)) critical("%s(): The file is not seekable.", who);
      want *= 2;
      continue;
    }
    critical("%s(): Failed to import an element: %s", who, qb200_last_error());
  }
  if (0 != fseek(file, pos + (long)used, SEEK_SET)) critical("%s(): The file is not seekable.", who);
  *total_probability = 0;
  for (size_t i = 0; i < n; i++) {
    values[i] = parsed[i];
    *total_probability += values[i];
  }
  *total_error = parsed[n];
  if (g_tstats.on) g_tstats.imports++;
}

uint32_t import_dimension(FILE* const file, const char* who) {
  uint32_t dimension;
  if (1 != fscanf(file, "%u\n", &dimension)) critical("%s(): Failed to import the dimension.", who);
  return dimension;
}

void import_common_2d(Distribution_Slice* const slice, FILE* const file) {
  const char* who = "distribution_slice_import_common";
  if (1 != fscanf(file, "%d\n", &(slice->min_log_alpha_d))) critical("%s(): Failed to import min_log_alpha_d.", who);
  if (1 != fscanf(file, "%d\n", &(slice->min_log_alpha_r))) critical("%s(): Failed to import min_log_alpha_r.", who);
  if (1 != fscanf(file, "%x\n", &(slice->flags))) critical("%s(): Failed to import flags.", who);
  import_values(file, slice->norm_matrix, (size_t)slice->dimension * slice->dimension,
                &slice->total_probability, &slice->total_error, who);
}

void import_common_linear(Linear_Distribution_Slice* const slice, FILE* const file) {
  const char* who = "linear_distribution_slice_import";
  if (1 != fscanf(file, "%d\n", &(slice->min_log_alpha))) critical("%s(): Failed to import min_log_alpha.", who);
  if (1 != fscanf(file, "%x\n", &(slice->flags))) critical("%s(): Failed to import flags.", who);
  import_values(file, slice->norm_vector, slice->dimension, &slice->total_probability,
                &slice->total_error, who);
}

void import_common_diagonal(Diagonal_Distribution_Slice* const slice, FILE* const file) {
  const char* who = "diagonal_distribution_slice_import_common";
  if (1 != fscanf(file, "%d\n", &(slice->min_log_alpha_r))) critical("%s(): Failed to import min_log_alpha_r.", who);
  if (1 != fscanf(file, "%d\n", &(slice->eta))) critical("%s(): Failed to import eta.", who);
  if (1 != fscanf(file, "%x\n", &(slice->flags))) critical("%s(): Failed to import flags.", who);
  import_values(file, slice->norm_vector, slice->dimension, &slice->total_probability,
                &slice->total_error, who);
}

}  // namespace

// Shared with dropin_collapse.cpp: one context and one lock for everything the server does.
qb200_context* qb200_dropin_text_context() { return text_context(); }
std::mutex& qb200_dropin_text_mutex() { return g_text_mutex; }

/* ---- two-dimensional slices (src/distribution_slice_import_export.cpp) ---------- */

void distribution_slice_import(Distribution_Slice* const slice, FILE* const file) {
  const uint32_t dimension = import_dimension(file, "distribution_slice_import");
  if (dimension != slice->dimension) {
    distribution_slice_clear(slice);
    distribution_slice_init(slice, dimension);
  }
  import_common_2d(slice, file);
}

void distribution_slice_init_import(Distribution_Slice* const slice, FILE* const file) {
  const uint32_t dimension = import_dimension(file, "distribution_slice_init_import");
  distribution_slice_init(slice, dimension);
  import_common_2d(slice, file);
}

void distribution_slice_export(const Distribution_Slice* const slice, FILE* const file) {
  fprintf(file, "%u\n", slice->dimension);
  fprintf(file, "%d\n", slice->min_log_alpha_d);
  fprintf(file, "%d\n", slice->min_log_alpha_r);
  fprintf(file, "%.8x\n", slice->flags);
  export_values(file, slice->norm_matrix, (size_t)slice->dimension * slice->dimension,
                slice->total_error, "distribution_slice_export");
}

/* ---- linear slices (src/linear_distribution_slice_import_export.cpp) ------------ */

void linear_distribution_slice_import(Linear_Distribution_Slice* const slice, FILE* const file) {
  const uint32_t dimension = import_dimension(file, "linear_distribution_slice_import");
  if (dimension != slice->dimension) {
    linear_distribution_slice_clear(slice);
    linear_distribution_slice_init(slice, dimension);
  }
  import_common_linear(slice, file);
}

void linear_distribution_slice_init_import(Linear_Distribution_Slice* const slice,
                                           FILE* const file) {
  const uint32_t dimension = import_dimension(file, "linear_distribution_slice_init_import");
  linear_distribution_slice_init(slice, dimension);
  import_common_linear(slice, file);
}

void linear_distribution_slice_export(const Linear_Distribution_Slice* const slice,
                                      FILE* const file) {
  fprintf(file, "%u\n", slice->dimension);
  fprintf(file, "%d\n", slice->min_log_alpha);
  fprintf(file, "%.8x\n", slice->flags);
  export_values(file, slice->norm_vector, slice->dimension, slice->total_error,
                "linear_distribution_slice_export");
}

/* ---- diagonal slices (src/diagonal_distribution_slice_import_export.cpp) --------- */

void diagonal_distribution_slice_import(Diagonal_Distribution_Slice* const slice,
                                        FILE* const file) {
  const uint32_t dimension = import_dimension(file, "diagonal_distribution_slice_import");
  if (dimension != slice->dimension) {
    diagonal_distribution_slice_clear(slice);
    diagonal_distribution_slice_init(slice, dimension);
  }
  import_common_diagonal(slice, file);
}

void diagonal_distribution_slice_init_import(Diagonal_Distribution_Slice* const slice,
                                             FILE* const file) {
  const uint32_t dimension = import_dimension(file, "diagonal_distribution_slice_init_import");
  diagonal_distribution_slice_init(slice, dimension);
  import_common_diagonal(slice, file);
}

void diagonal_distribution_slice_export(const Diagonal_Distribution_Slice* const slice,
                                        FILE* const file) {
  fprintf(file, "%u\n", slice->dimension);
  fprintf(file, "%d\n", slice->min_log_alpha_r);
  fprintf(file, "%d\n", slice->eta);
  fprintf(file, "%.8x\n", slice->flags);
  export_values(file, slice->norm_vector, slice->dimension, slice->total_error,
                "diagonal_distribution_slice_export");
}
