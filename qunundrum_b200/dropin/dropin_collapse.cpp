// dropin_collapse.cpp -- the reference-side forwarding translation unit for the SERVER's work on a
// finished two-dimensional distribution.
//
// A maintainer of ekera/qunundrum deletes the two functions
//
//   linear_distribution_init_collapse_d   src/linear_distribution.cpp:152-237
//   linear_distribution_init_collapse_r   src/linear_distribution.cpp:239-324
//
// from src/linear_distribution.cpp and adds this file. (integration/build.py does the same without
// touching the reference: it compiles linear_distribution.cpp where it lies with the two names
// renamed by -D, as it does for tau_estimate.cpp.) Callers are unchanged:
// main_server_export_collapsed_distributions, src/main_generate_distribution.cpp:709-760, and
// filter_distribution / info_distribution.
//
// What the reference does there: for every slice of the distribution and every one of its D^2
// cells, `norm_vector[...] += probability / (long double)divisor` -- 2 x 10^8 long double divisions
// and additions for an m = 2048 distribution, 1.4 s of the server's wall clock once integration
// and export are on the GPU. Here the slices' cells are gathered ONCE into device memory
// (qb200_resident_create: pinned, double-buffered upload of the norm_matrix arrays as they lie),
// both marginals come from qb200_resident_collapse2d -- every element is the reference's own
// sequence of 64-bit-mantissa operations in the reference's order, so the collapsed slices are
// bit-identical to the reference's -- and the export that follows formats the slices from the
// same device copy without a second upload (qb200_dropin_resident_text below, used by
// dropin_text.cpp when this file is linked in).
//
// Conventions kept: dst is initialised with linear_distribution_init (capacity src->count), the
// destination slices are created in the order the source slices first name their coordinate,
// their total_probability / total_error are the running long double sums in source order, the
// distribution totals are copied from src, errors are fatal (critical(), same messages).
#include "common.h"
#include "distribution.h"
#include "distribution_slice.h"
#include "errors.h"
#include "linear_distribution.h"
#include "linear_distribution_slice.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <vector>

#include "qunundrum_b200.h"

// dropin_text.cpp: the server's context (shared so that the exporter formats from the same
// device copy) and the lock that serialises its users (the server exports from worker threads).
qb200_context* qb200_dropin_text_context();
std::mutex& qb200_dropin_text_mutex();

namespace {

double cnow() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct Mark {      // what identifies an uploaded slice: address, size and a few of its values
  uint32_t index;
  uint64_t cells;
  long double first, middle, last, tail;
};

struct Registry {
  qb200_resident* resident;
  std::map<const long double*, Mark> by_cells;
  std::vector<uint32_t> dimension;   // of resident slice i
  // the exporter's look-ahead: text of the resident slices [text_first, text_first + text_count)
  const char* text;
  std::vector<size_t> offsets, lengths;
  uint32_t text_first, text_count;
  double upload_s, collapse_s, format_s;
  unsigned long uploads, collapses, formats, text_hits;
  bool stats;
  Registry() : resident(NULL), text(NULL), text_first(0), text_count(0), upload_s(0), collapse_s(0),
               format_s(0), uploads(0), collapses(0), formats(0), text_hits(0), stats(false) {}
} g_reg;

void print_collapse_stats() {
  if (g_reg.stats)
    {
    char line[768];
    const int len = snprintf(line, sizeof line,
            "qunundrum_b200 collapse drop-in: %lu uploads (%.3f s), %lu collapses (%.3f s), %lu export "
            "batches (%.3f s) serving %lu slice exports from the device copy\n",
            g_reg.uploads, g_reg.upload_s, g_reg.collapses, g_reg.collapse_s, g_reg.formats,
            g_reg.format_s, g_reg.text_hits);
    if (len > 0) (void)!write(2, line, (size_t)(len < (int)sizeof line ? len : (int)sizeof line - 1));
  }
}

bool mark_matches(const Mark& k, const long double* cells, uint64_t n, long double tail) {
  return k.cells == n && (n == 0 || (k.first == cells[0] && k.middle == cells[n / 2] && k.last == cells[n - 1])) &&
         k.tail == tail;
}

void drop_resident() {
  if (g_reg.resident) qb200_resident_destroy(g_reg.resident);
  g_reg.resident = NULL;
  g_reg.by_cells.clear();
  g_reg.dimension.clear();
  g_reg.text = NULL;
  g_reg.text_count = 0;
}

// The device copy of src's slices: the one at hand if it still holds every slice of src (the
// filtered distribution is a subset of the unfiltered one), else a fresh upload.
void resident_for(const Distribution* const src, std::vector<uint32_t>* index_of, const char* who) {
  static bool once = false;
  if (!once) {
    once = true;
    const char* st = getenv("QB200_DROPIN_STATS");
    if (st && *st && *st != '0') {
      g_reg.stats = true;
      atexit(print_collapse_stats);
    }
  }
  index_of->assign(src->count, 0);
  bool reuse = g_reg.resident != NULL;
  for (uint32_t i = 0; reuse && i < src->count; i++) {
    const Distribution_Slice* s = src->slices[i];
    std::map<const long double*, Mark>::const_iterator it = g_reg.by_cells.find(s->norm_matrix);
    reuse = it != g_reg.by_cells.end() &&
            mark_matches(it->second, s->norm_matrix, (uint64_t)s->dimension * s->dimension, s->total_error);
    if (reuse) (*index_of)[i] = it->second.index;
  }
  if (reuse) return;
  drop_resident();
  const double t0 = cnow();
  std::vector<uint64_t> n_cells(src->count);
  std::vector<const long double*> cells(src->count);
  std::vector<long double> tails(src->count);
  for (uint32_t i = 0; i < src->count; i++) {
    const Distribution_Slice* s = src->slices[i];
    n_cells[i] = (uint64_t)s->dimension * s->dimension;
    cells[i] = s->norm_matrix;
    tails[i] = s->total_error;
  }
  if (0 != qb200_resident_create(qb200_dropin_text_context(), src->count, n_cells.data(), cells.data(),
                                 tails.data(), &g_reg.resident)) {
    critical("%s(): %s", who, qb200_last_error());
  }
  g_reg.dimension.resize(src->count);
  for (uint32_t i = 0; i < src->count; i++) {
    const Distribution_Slice* s = src->slices[i];
    Mark k;
    k.index = i;
    k.cells = n_cells[i];
    k.first = k.cells ? s->norm_matrix[0] : 0;
    k.middle = k.cells ? s->norm_matrix[k.cells / 2] : 0;
    k.last = k.cells ? s->norm_matrix[k.cells - 1] : 0;
    k.tail = s->total_error;
    g_reg.by_cells[s->norm_matrix] = k;
    g_reg.dimension[i] = s->dimension;
    (*index_of)[i] = i;
  }
  g_reg.uploads++;
  g_reg.upload_s += cnow() - t0;
}

void collapse(Linear_Distribution* const dst, const Distribution* const src, const int axis,
              const char* who) {
  uint32_t flags = (0 == axis) ? LINEAR_DISTRIBUTION_FLAG_D : LINEAR_DISTRIBUTION_FLAG_R;
  flags |= LINEAR_DISTRIBUTION_FLAG_COLLAPSED;
  linear_distribution_init(dst, &(src->parameters), flags, src->count);
  if (0 == src->count) return;

  uint32_t max_dimension = src->slices[0]->dimension;
  for (uint32_t i = 1; i < src->count; i++) {
    if (src->slices[i]->dimension > max_dimension) max_dimension = src->slices[i]->dimension;
  }
  for (uint32_t i = 0; i < src->count; i++) {
    if ((max_dimension % src->slices[i]->dimension) != 0) {
      critical("%s(): All slices in the source distribution must be of dimension that divides the "
               "maximum slice dimension.", who);
    }
  }

  // destination slices in first-appearance order; totals as the reference's running sums
  std::map<int32_t, uint32_t> where;
  std::vector<std::vector<uint32_t> > members;
  for (uint32_t i = 0; i < src->count; i++) {
    const Distribution_Slice* s = src->slices[i];
    const int32_t coordinate = (0 == axis) ? s->min_log_alpha_d : s->min_log_alpha_r;
    std::map<int32_t, uint32_t>::const_iterator it = where.find(coordinate);
    uint32_t j;
    if (it == where.end()) {
      Linear_Distribution_Slice* slice = linear_distribution_slice_alloc();
      linear_distribution_slice_init(slice, max_dimension);
      slice->min_log_alpha = coordinate;
      linear_distribution_insert_slice(dst, slice);
      j = dst->count - 1;
      where[coordinate] = j;
      members.push_back(std::vector<uint32_t>());
    } else {
      j = it->second;
    }
    dst->slices[j]->total_probability += s->total_probability;
    dst->slices[j]->total_error += s->total_error;
    members[j].push_back(i);
  }

  {
    std::lock_guard<std::mutex> lock(qb200_dropin_text_mutex());
    std::vector<uint32_t> index_of;
    resident_for(src, &index_of, who);
    const double t0 = cnow();
    const uint32_t n_dst = dst->count;
    std::vector<uint32_t> begin(n_dst + 1, 0), list;
    for (uint32_t j = 0; j < n_dst; j++) {
      for (size_t k = 0; k < members[j].size(); k++) list.push_back(index_of[members[j][k]]);
      begin[j + 1] = (uint32_t)list.size();
    }
    std::vector<long double> out((size_t)n_dst * max_dimension);
    if (0 != qb200_resident_collapse2d(g_reg.resident, axis, g_reg.dimension.data(), n_dst, begin.data(),
                                       list.data(), max_dimension, out.data())) {
      critical("%s(): %s", who, qb200_last_error());
    }
    for (uint32_t j = 0; j < n_dst; j++) {
      memcpy(dst->slices[j]->norm_vector, out.data() + (size_t)j * max_dimension,
             (size_t)max_dimension * sizeof(long double));
    }
    g_reg.collapses++;
    g_reg.collapse_s += cnow() - t0;
  }

  dst->total_probability = src->total_probability;
  dst->total_error = src->total_error;
}

// The exporter's look-ahead from resident slice `first`: up to 64 slices or ~32 MB of text (the
// library keeps two pinned buffers of the largest batch; pinning is not free).
uint32_t batch_from(uint32_t first) {
  const uint32_t total = (uint32_t)g_reg.dimension.size();
  uint32_t count = 0;
  size_t bytes = 0;
  while (first + count < total && count < 64 && bytes < (size_t(32) << 20)) {
    bytes += 30 * ((size_t)g_reg.dimension[first + count] * g_reg.dimension[first + count] + 1);
    count++;
  }
  return count;
}

}  // namespace

// For dropin_text.cpp (caller holds qb200_dropin_text_mutex()): the "%.24Lg\n" lines of a slice
// whose cells are on the device -- its n cells, then `tail` -- or false if they are not (any more).
// Slices are formatted ahead in resident order, a few dozen per synchronisation, which is the order
// distribution_export walks them in.
bool qb200_dropin_resident_text(const long double* cells, size_t n, long double tail, const char** text,
                                size_t* len) {
  if (NULL == g_reg.resident) return false;
  std::map<const long double*, Mark>::const_iterator it = g_reg.by_cells.find(cells);
  if (it == g_reg.by_cells.end() || !mark_matches(it->second, cells, n, tail)) return false;
  const uint32_t i = it->second.index;
  if (NULL == g_reg.text || i < g_reg.text_first || i >= g_reg.text_first + g_reg.text_count) {
    const double t0 = cnow();
    const uint32_t count = batch_from(i);
    g_reg.offsets.assign(count, 0);
    g_reg.lengths.assign(count, 0);
    if (0 != qb200_resident_format(g_reg.resident, i, count, &g_reg.text, g_reg.offsets.data(),
                                   g_reg.lengths.data())) {
      critical("distribution_slice_export(): %s", qb200_last_error());
    }
    g_reg.text_first = i;
    g_reg.text_count = count;
    // ... and the batch after it is formatted while the caller writes this one
    const uint32_t next = batch_from(i + count);
    if (next && 0 != qb200_resident_format_prefetch(g_reg.resident, i + count, next)) {
      critical("distribution_slice_export(): %s", qb200_last_error());
    }
    g_reg.formats++;
    g_reg.format_s += cnow() - t0;
  }
  const uint32_t k = i - g_reg.text_first;
  *text = g_reg.text + g_reg.offsets[k];
  *len = g_reg.lengths[k];
  g_reg.text_hits++;
  return true;
}

void linear_distribution_init_collapse_d(Linear_Distribution* const dst, const Distribution* const src) {
  collapse(dst, src, 0, "linear_distribution_init_collapse_d");
}

void linear_distribution_init_collapse_r(Linear_Distribution* const dst, const Distribution* const src) {
  collapse(dst, src, 1, "linear_distribution_init_collapse_r");
}
