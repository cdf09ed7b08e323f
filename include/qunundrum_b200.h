/* qunundrum_b200.h -- C ABI of the B200 slice integrators.
 *
 * Drop-in boundary for the slice-integration hot path of ekera/qunundrum.
 * Every entry point below replaces (a batch of calls to) one of the
 * reference's six C++ entry points; the reference-side forwarding TU a
 * maintainer adds is shown in INTEGRATION.md and built in
 * qunundrum_b200/csrc/dropin.cpp.
 *
 *   entry point here                         replaces (reference file:line)
 *   qb200_slice2d_compute                    distribution_slice_compute            src/distribution_slice_compute.cpp:38
 *                                            distribution_slice_compute_richardson src/distribution_slice_compute_richardson.cpp:17
 *   qb200_slice1d_compute (LINEAR_D / _R)    linear_distribution_slice_compute            src/linear_distribution_slice_compute.cpp:30
 *                                            linear_distribution_slice_compute_richardson src/linear_distribution_slice_compute_richardson.cpp:17
 *   qb200_slice1d_compute (DIAGONAL)         diagonal_distribution_slice_compute            src/diagonal_distribution_slice_compute.cpp:30
 *                                            diagonal_distribution_slice_compute_richardson src/diagonal_distribution_slice_compute_richardson.cpp:17
 *   qb200_slice2d_compute_scaled             ... followed by distribution_slice_copy_scale    src/distribution_slice.cpp:230
 *   qb200_resident_collapse2d                linear_distribution_init_collapse_d / _r       src/linear_distribution.cpp:152,239
 *   qb200_resident_format                    the value loops of distribution_export         src/distribution.cpp:305-357
 *   qb200_text_format_* / qb200_text_parse_* the value loops of *_slice_export / *_slice_import   (see "text export" below)
 *   qb200_sampler_tau_estimate               tau_estimate / tau_estimate_linear             src/tau_estimate.cpp:23,89
 *   qb200_diagk_sample                       sample_k_from_diagonal_j_eta_pivot             src/sample.cpp:412
 *   qb200_diagk_h                            diagonal_probability_approx_h                  src/diagonal_probability.cpp:99
 *   qb200_diagk_tau_estimate                 the sum and log of tau_estimate_diagonal       src/tau_estimate.cpp:135
 *   qb200_exact_alpha                        sample_alpha_from_region                       src/sample.cpp:78
 *   qb200_exact_j_k                          sample_j_from_alpha_r, sample_j_k_from_alpha_d, src/sample.cpp:160,210,275,354
 *                                            sample_j_k_from_alpha_d_r, sample_j_from_diagonal_alpha_r
 *   qb200_diagk_sample_drawn                 diagonal_distribution_sample_pair_j_k after    src/diagonal_distribution.cpp:474
 *                                            the region is chosen
 *
 * Conventions
 *  - plain pointers and sizes only; all buffers are caller-owned;
 *  - d and r (Parameters::d, Parameters::r, src/parameters.h:60-72) travel as
 *    big-endian magnitude bytes, i.e. what mpz_export(buf, &n, 1, 1, 1, 0, z)
 *    writes;
 *  - cells are IEEE doubles in the reference's own index order
 *    (2D: index = i_d + dimension * j_r, src/distribution_slice_compute.cpp:401);
 *    the binding widens them to the long double norm_matrix / norm_vector;
 *  - flags are the bits the reference ORs into slice->flags after clearing
 *    SLICE_FLAGS_MASK_METHOD (src/common.h:226-257): SIMPSON, RICHARDSON and the
 *    ERROR_BOUND_WARNING bit;
 *  - every function returns 0 on success and a negative code on failure, with
 *    qb200_last_error() giving the message; the reference convention (errors
 *    are fatal, critical() -> exit, src/errors.c) is applied by the binding;
 *  - there is no CPU path: without a CUDA device qb200_create() fails;
 *  - a context is not thread-safe: use it from one thread at a time (one context
 *    per MPI rank / per exporting thread, or a lock as in dropin_text.cpp).
 */
#ifndef QUNUNDRUM_B200_H
#define QUNUNDRUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_VERSION 2

/* Distribution_Slice_Compute_Method, src/distribution_slice.h:31-78. */
#define QB200_METHOD_HEURISTIC_SIGMA 0
#define QB200_METHOD_OPTIMAL_LOCAL_SIGMA 1
#define QB200_METHOD_QUICK 2

/* One-dimensional integrands. LINEAR_D / LINEAR_R are
 * Linear_Distribution_Slice_Compute_Target, src/linear_distribution_slice.h:30-40. */
#define QB200_KIND_LINEAR_D 0
#define QB200_KIND_LINEAR_R 1
#define QB200_KIND_DIAGONAL 2

/* Slice flag bits, src/common.h:226-257. */
#define QB200_FLAG_ERROR_BOUND_WARNING 0x00000001u
#define QB200_FLAG_METHOD_SIMPSON 0x00020000u
#define QB200_FLAG_METHOD_RICHARDSON 0x00080000u

/* The fields of Parameters (src/parameters.h:33-114) / Diagonal_Parameters
 * (src/diagonal_parameters.h:32-106) that the integrators read. */
typedef struct {
  uint32_t m;
  uint32_t l;
  uint32_t sigma; /* Diagonal_Parameters::sigma; ignored otherwise */
  const uint8_t *d_be;
  size_t d_len;
  const uint8_t *r_be;
  size_t r_len;
} qb200_params;

typedef struct qb200_context qb200_context;
typedef struct qb200_plan qb200_plan;

int qb200_version(void);

/* Message of the last failure on this thread (never NULL). */
const char *qb200_last_error(void);

/* Number of CUDA devices visible (0 if none / no driver). */
int qb200_device_count(void);

/* One context per MPI worker rank; device = (rank - 1) mod qb200_device_count(). */
int qb200_create(int device, qb200_context **ctx);
void qb200_destroy(qb200_context *ctx);

/* Kernel launches issued by this context so far (for accounting). */
uint64_t qb200_launch_count(const qb200_context *ctx);

/* Pinned host memory for result buffers (optional; plain malloc works too). */
void *qb200_host_alloc(size_t bytes);
void qb200_host_free(void *p);

/* ---- synchronous, host buffers in, host buffers out ----------------------- */

/* n two-dimensional slices of dimension `dimension` with coordinates
 * (min_log_alpha_d[i], min_log_alpha_r[i]).
 *   cells             n * dimension^2 doubles
 *   total_probability n long doubles (sum of the slice's cells)
 *   total_error       n long doubles (0 for the quick method)
 *   flags             n words
 * richardson = 0: distribution_slice_compute; 1: ..._compute_richardson. */
int qb200_slice2d_compute(qb200_context *ctx, const qb200_params *params, int method,
                          int richardson, uint32_t dimension, uint32_t n,
                          const int32_t *min_log_alpha_d, const int32_t *min_log_alpha_r,
                          double *cells, long double *total_probability,
                          long double *total_error, uint32_t *flags);

/* n one-dimensional slices. eta is read for QB200_KIND_DIAGONAL only (may be
 * NULL otherwise). cells: n * dimension doubles. */
int qb200_slice1d_compute(qb200_context *ctx, const qb200_params *params, int kind,
                          int richardson, uint32_t dimension, uint32_t n,
                          const int32_t *min_log_alpha, const int32_t *eta, double *cells,
                          long double *total_probability, uint32_t *flags);

/* The generator client's tail: a slice computed at 512 or 1024 is
 * scaled to MAX_SLICE_DIMENSION = 256 with distribution_slice_copy_scale before it is sent
 * (src/main_generate_distribution.cpp:1308-1343, src/distribution_slice.cpp:230-264). Here the
 * scaling happens on the device and only store_dimension^2 cells per slice cross the bus:
 *   cells  n * store_dimension^2 LONG DOUBLES (a scaled cell is the long double sum of
 *          (dimension / store_dimension)^2 cells in the reference's order, each partial sum
 *          rounded to 64 bits: bit-identical to copy_scale applied to the unscaled doubles;
 *          it does not fit a double)
 *   total_probability / total_error / flags as qb200_slice2d_compute (of the UNscaled slice, which
 *          is what copy_scale hands on; the binding ORs SLICE_FLAGS_SCALED).
 * dimension must be a multiple of store_dimension. */
int qb200_slice2d_compute_scaled(qb200_context *ctx, const qb200_params *params, int method,
                                 int richardson, uint32_t dimension, uint32_t store_dimension,
                                 uint32_t n, const int32_t *min_log_alpha_d,
                                 const int32_t *min_log_alpha_r, long double *cells,
                                 long double *total_probability, long double *total_error,
                                 uint32_t *flags);

/* ---- planned, device-resident execution ----------------------------------- */

/* A plan holds the batch's constants, coordinates and axis-table descriptors
 * in device memory. qb200_plan_run() enqueues all kernels of one pass over the
 * batch on `stream` (a cudaStream_t, NULL = the context's own stream) and
 * returns without synchronising:
 *   d_cells    device, n * dimension^(1|2) doubles
 *   d_summary  device, n * QB200_SUMMARY_STRIDE doubles (see below)
 * qb200_plan_finish() turns a host copy of the summary into the reference's
 * per-slice scalars. */
#define QB200_SUMMARY_STRIDE 8 /* tp_hi, tp_lo, m1, m2, bounded, 0, 0, 0 */

int qb200_plan2d_create(qb200_context *ctx, const qb200_params *params, int method,
                        int richardson, uint32_t dimension, uint32_t n,
                        const int32_t *min_log_alpha_d, const int32_t *min_log_alpha_r,
                        qb200_plan **plan);
int qb200_plan1d_create(qb200_context *ctx, const qb200_params *params, int kind,
                        int richardson, uint32_t dimension, uint32_t n,
                        const int32_t *min_log_alpha, const int32_t *eta,
                        qb200_plan **plan);
void qb200_plan_destroy(qb200_plan *plan);

/* Number of cells one run produces, and kernel launches per run. */
uint64_t qb200_plan_cells(const qb200_plan *plan);
uint32_t qb200_plan_launches(const qb200_plan *plan);

/* algo: 0 = automatic, 1 = plain kernels (one thread per cell), 2 = fused
 * kernel (fails if the plan does not meet its preconditions). */
int qb200_plan_set_algorithm(qb200_plan *plan, int algo);
int qb200_plan_algorithm(const qb200_plan *plan);

/* (A two-dimensional plan that is run again with the same d_cells / d_summary replays its step
 * as one CUDA graph, captured during the second such run; any other buffers run eagerly.) */
int qb200_plan_run(qb200_plan *plan, void *stream, double *d_cells, double *d_summary);
int qb200_plan_finish(const qb200_plan *plan, const double *h_summary,
                      long double *total_probability, long double *total_error,
                      uint32_t *flags);

/* ---- text export: "%.24Lg\n" lines ------------------------------------------
 *
 * SURVEY.md section 8(f) #1. Every cell of a stored slice is written by
 *   fprintf(file, "%.24Lg\n", slice->norm_matrix[i]);   (then total_error likewise)
 * in distribution_slice_export          src/distribution_slice_import_export.cpp:89-103
 *    linear_distribution_slice_export   src/linear_distribution_slice_import_export.cpp:82-97
 *    diagonal_distribution_slice_export src/diagonal_distribution_slice_import_export.cpp:87-103
 * -- 1.05e8 calls, 3.1 GB of text and 87 % of the wall clock of one m = 2048
 * distribution once the integration runs on the GPU. The entry points below
 * produce the same bytes (glibc semantics: exact value, round-half-even, %g
 * style selection, trailing zeros removed, inf / nan / signed zero) with one
 * kernel launch per call. */
#define QB200_TEXT_X87 0 /* x86-64 long double: 64-bit mantissa, 16-byte stride */
#define QB200_TEXT_F64 1 /* IEEE double, printed as (long double)x */

/* Largest possible text of n values (33 bytes each). */
size_t qb200_text_bound(size_t n);

/* n values followed by the optional `tail` value (a slice's total_error; NULL
 * for none), one line each. *text points into pinned host memory owned by the
 * context, valid until the next text call on it; *len is its length (no NUL). */
int qb200_text_format_ld(qb200_context *ctx, const long double *values, size_t n,
                         const long double *tail, const char **text, size_t *len);
int qb200_text_format_f64(qb200_context *ctx, const double *values, size_t n,
                          const double *tail, const char **text, size_t *len);

/* Device-resident form: enqueues one launch on `stream` (a cudaStream_t, NULL =
 * the context's own) and returns without synchronising. d_text has room for
 * cap bytes (qb200_text_bound(n) always suffices); *d_len (device) receives
 * the text length. kind = QB200_TEXT_X87 / QB200_TEXT_F64. */
int qb200_text_format_device(qb200_context *ctx, int kind, const void *d_values, size_t n,
                             char *d_text, size_t cap, uint64_t *d_len, void *stream);

/* ---- a stored distribution on the device ---------------------------------------
 *
 * what the generator's server does to a finished two-dimensional
 * distribution -- collapse it to its two marginals
 *   linear_distribution_init_collapse_d   src/linear_distribution.cpp:152-237
 *   linear_distribution_init_collapse_r   src/linear_distribution.cpp:239-324
 * (called from src/main_generate_distribution.cpp:709-760; ~2 x 10^8 long double divisions and
 * additions) and export every slice (distribution_export, src/distribution.cpp:305-357).
 * A qb200_resident holds the cells of all slices in device memory (uploaded ONCE from the
 * slices' own norm_matrix arrays by a pinned, double-buffered gather), each slice followed by one
 * spare value, its `tail` (total_error), so that a slice's export is one contiguous range. */
typedef struct qb200_resident qb200_resident;

/* cells[i]: n_cells[i] long doubles (read during the call only); tails may be NULL (zeros). */
int qb200_resident_create(qb200_context *ctx, uint32_t n_slices, const uint64_t *n_cells,
                          const long double *const *cells, const long double *tails,
                          qb200_resident **resident);
void qb200_resident_destroy(qb200_resident *resident);
uint64_t qb200_resident_cells(const qb200_resident *resident);

/* Collapse to a marginal: axis 0 = alpha_d (element x of a destination vector collects
 * norm_matrix[x + y * dimension] over y), axis 1 = alpha_r. dimension[i]: the dimension of
 * resident slice i (n_cells[i] = dimension[i]^2, a divisor of max_dimension). Destination k
 * sums the slices src_index[src_begin[k] .. src_begin[k + 1]) in that order -- the binding lists
 * them in the distribution's order, as the reference's loop meets them. out: n_dst *
 * max_dimension long doubles, bit-identical to the reference's: every element is the
 * reference's own sequence of `+= probability / (long double)divisor` in 64-bit arithmetic. */
int qb200_resident_collapse2d(qb200_resident *resident, int axis, const uint32_t *dimension,
                              uint32_t n_dst, const uint32_t *src_begin, const uint32_t *src_index,
                              uint32_t max_dimension, long double *out);

/* "%.24Lg\\n" lines of the slices [first, first + count): slice first + i occupies
 * text[offsets[i], offsets[i] + lengths[i]) -- its cells, then its tail. One exporter launch per
 * slice, enqueued back to back, no upload, one synchronisation; *text points into pinned memory
 * owned by the resident, valid until the call after the next one (two buffer sets).
 * qb200_resident_format_prefetch starts the same work for a batch and returns at once: a following
 * qb200_resident_format of exactly that batch only waits for it -- the caller writes batch k to its
 * file while batch k + 1 is formatted. offsets, lengths: count entries each. */
int qb200_resident_format(qb200_resident *resident, uint32_t first, uint32_t count,
                          const char **text, size_t *offsets, size_t *lengths);
int qb200_resident_format_prefetch(qb200_resident *resident, uint32_t first, uint32_t count);

/* ---- text import: "%Lg" numbers ------------------------------------------------
 *
 * The importers read every cell back with
 *   fscanf(file, "%Lg\n", &slice->norm_matrix[i])
 * (distribution_slice_import_common, src/distribution_slice_import_export.cpp:18-52;
 * linear: src/linear_distribution_slice_import_export.cpp:18-47; diagonal:
 * src/diagonal_distribution_slice_import_export.cpp:18-52). qb200_text_parse_ld
 * converts the first n white-space separated numbers of text[0, len) to x87 long
 * doubles, correctly rounded (nearest, ties to even -- what glibc's strtold
 * returns), including inf / nan, denormals, overflow and underflow. *consumed
 * (may be NULL) receives the offset of the byte after the n-th number and the
 * white space that follows it, i.e. where the reference's FILE position would be.
 * Errors: -20 fewer than n numbers, -21 a malformed number (or one longer than 255
 * characters), -22 an unsupported form (hexadecimal floats; more than 28
 * significant digits exactly on a rounding boundary). text needs no terminating
 * NUL. text[0, len) is taken to be complete: a number that ends exactly at
 * text + len is converted as it stands, so a caller that passes a block of a
 * larger file must make sure that something follows the last number it needs
 * (*consumed < len) or that the block reaches the end of the file. */
int qb200_text_parse_ld(qb200_context *ctx, const char *text, size_t len, size_t n,
                        long double *values, size_t *consumed);

/* Device-resident form. d_text must be readable up to the next multiple of 16
 * bytes past len (any content). d_values: n x 16 bytes. d_info: 5 uint64 on the
 * device: numbers found, status (0 ok, 1 malformed, 2 unsupported), index of the
 * first bad number, numbers that needed the exact decision, offset of number n
 * (len if there is none). Two launches on `stream`, no synchronisation. */
int qb200_text_parse_device(qb200_context *ctx, const char *d_text, size_t len, size_t n,
                            void *d_values, uint64_t *d_info, void *stream);

/* Introspection / test hooks (host logic; qb200_text_pow10 needs no GPU):
 * the 192-bit table entry of 10^k (little-endian 32-bit limbs, value =
 * T * 2^(e2 - 191), truncated; exact = 1 if nothing was cut off); a switch that
 * sends every value through the exact rounding decision (on = 1; on = 2 is a
 * profiling-only mode of the exporter that writes every tile at a fixed stride
 * instead of its chained offset -- the text is then NOT contiguous -- to time the
 * kernel without the look-back chain); and the number of values of the last call
 * that needed the exact decision. */
int qb200_text_pow10(int k, uint32_t w[6], int32_t *e2, uint32_t *exact);
int qb200_text_set_force_exact(qb200_context *ctx, int on);
uint64_t qb200_text_exact_count(qb200_context *ctx);

/* ---- sampling from a stored distribution: tau estimation ----------------------
 *
 * SURVEY.md section 8(f) #3. estimate_runs_* draw 10^6 estimates of n samples per tried n
 * (src/executables_estimate_runs_distribution.h:23-28) through
 *   tau_estimate          src/tau_estimate.cpp:23-87    -> distribution_sample_approximate_alpha_d_r
 *                                                          src/distribution.cpp:464-527
 *   tau_estimate_linear   src/tau_estimate.cpp:89-133   -> linear_distribution_sample_approximate_alpha
 *                                                          src/linear_distribution.cpp:618-666
 * and every sample is two linear long double walks (slices: src/distribution.cpp:384-401, then
 * the D^2 cells of the slice: src/distribution_slice.cpp:191-223). A qb200_sampler holds the
 * distribution in device memory; the entry points below take the random stream as the 64-bit
 * words the reference's random_generate_pivot_* (src/random.c:116-156) would have drawn, in
 * its order (slice pivot, region pivot, one fraction per axis), and return what the reference
 * returns for the same words: the same slices and regions (its x87 rounding included), the
 * same alphas to ~1e-30 relative, tau to the last place or two of a long double. */
#define QB200_SAMPLER_LINEAR 1 /* Linear_Distribution, src/linear_distribution.h */
#define QB200_SAMPLER_2D 2     /* Distribution, src/distribution.h:55-85 */

typedef struct qb200_sampler qb200_sampler;

/* The slices in the distribution's current (walk) order: dimension[i], coordinates
 * (c0 = min_log_alpha_d, c1 = min_log_alpha_r; linear: c0 = min_log_alpha, c1 may be NULL),
 * cells[i] = norm_matrix / norm_vector (dimension^dims long doubles, read during the call
 * only), slice_total[i] = total_probability of the slice, total_probability of the
 * distribution, m of its parameters. Dimensions must be powers of two. */
int qb200_sampler_create(qb200_context *ctx, int dims, uint32_t m, uint32_t n_slices,
                         const uint32_t *dimension, const int32_t *c0, const int32_t *c1,
                         const long double *const *cells, const long double *slice_total,
                         long double total_probability, qb200_sampler **sampler);
void qb200_sampler_destroy(qb200_sampler *sampler);

/* 64-bit words one successful sample consumes: 4 (two-dimensional) or 3 (linear). A sample
 * whose slice pivot runs past the last slice consumes one word and fails. */
uint32_t qb200_sampler_words_per_sample(const qb200_sampler *sampler);
uint64_t qb200_sampler_cells(const qb200_sampler *sampler);

/* k independent samples, sample i from words[i * words_per_sample ...]: index of the slice
 * and of the region (cell) the reference would select, alpha / 2^m per axis rounded to double
 * (signed), status 0 ok / 1 out of bounds (the reference returns FALSE) / 2 no region (the
 * reference calls critical()). Any output pointer may be NULL. */
int qb200_sampler_sample(qb200_sampler *sampler, uint32_t k, const uint64_t *words, int32_t *slice,
                         int32_t *cell, double *x0, double *x1, int32_t *status);

/* Up to `count` consecutive calls of tau_estimate(distribution, random_state, n, ...) on the
 * stream words[0, n_words): estimate t starts where estimate t - 1 stopped reading (an
 * estimate stops at its first out-of-bounds sample: ok = 0, tau = DBL_MAX, as the reference).
 * *done = estimates completed before the words ran out (count if n_words >= count * n *
 * words_per_sample), *words_used = words they consumed. tau1 is written for two-dimensional
 * distributions only (tau_d in tau0, tau_r in tau1) and may be NULL otherwise.
 * Error -41: a region walk ran off its slice (the reference: critical("Failed to sample a
 * region from the slice.")). */
int qb200_sampler_tau_estimate(qb200_sampler *sampler, uint32_t n, uint32_t count,
                               const uint64_t *words, size_t n_words, size_t *words_used,
                               uint32_t *done, long double *tau0, long double *tau1, uint8_t *ok);

/* Device-resident form: d_words holds count * n * words_per_sample words in the regular
 * layout; two launches on `stream` (NULL = the context's own), no synchronisation. d_sums:
 * count x 4 doubles (sum of (alpha_d / 2^m)^2 as hi, lo, then the same for alpha_r); d_status:
 * count ints (0, or the status of the estimate's first failing sample). */
int qb200_sampler_tau_device(qb200_sampler *sampler, uint32_t n, uint32_t count,
                             const uint64_t *d_words, double *d_sums, int32_t *d_status,
                             void *stream);

/* Test hooks: send every walk through the bit-exact x87 replay (on = 1) or every search through
 * the double-double path, skipping the quick pass in doubles (on = 2); the number of walks of the
 * last call that needed the replay; the smallest slice-pivot word that runs out of bounds
 * (return value 0: none does). */
int qb200_sampler_set_force_exact(qb200_sampler *sampler, int on);
uint64_t qb200_sampler_exact_count(const qb200_sampler *sampler);
int qb200_sampler_first_failing_word(const qb200_sampler *sampler, uint64_t *word);

/* ---- diagonal distribution: k given (j, eta) ------------------------------------------
 *
 * SURVEY.md section 8(f) #3, second half. diagonal_distribution_sample_pair_j_k
 * (src/diagonal_distribution.cpp:474-552) first draws (j, eta) from the stored distribution and
 * then calls
 *   sample_k_from_diagonal_j_eta_pivot   src/sample.cpp:412-646
 * which evaluates diagonal_probability_approx_h (src/diagonal_probability.cpp:99-162: two
 * mpfr_sin at 2 (m + sigma) bits) for k = k0, k0 + 1, k0 - 1, ... at 3 (m + sigma) bits until the
 * pivot is used up. Here the same k comes out of integer arithmetic on m-bit numbers (one
 * product r j, one product with d, Barrett divisions by r) and a double-double walk; see
 * qunundrum_b200/csrc/diagk.cuh. A qb200_diagk holds d, r and the reciprocal of r on the device.
 *
 * Integers travel as little-endian 32-bit words (mpz_export(buf, &count, -1, 4, 0, 0, z)), one
 * row per sample, zero padded: j in qb200_diagk_j_limbs() = ceil((m + sigma) / 32) words
 * (0 <= j < 2^(m + sigma)), k in qb200_diagk_k_limbs() = ceil(l / 32) words.
 * params: m, l, sigma, d, r of the Diagonal_Parameters (0 < d < r < 2^m, l <= m + sigma,
 * m + sigma <= 16384). */
typedef struct qb200_diagk qb200_diagk;

int qb200_diagk_create(qb200_context *ctx, const qb200_params *params, qb200_diagk **sampler);
void qb200_diagk_destroy(qb200_diagk *sampler);
uint32_t qb200_diagk_j_limbs(const qb200_diagk *sampler);
uint32_t qb200_diagk_k_limbs(const qb200_diagk *sampler);

#define QB200_DIAGK_OK 0
#define QB200_DIAGK_OUT_OF_BOUNDS 1 /* the reference returns FALSE (k = 0, alpha_phi = 0) */
#define QB200_DIAGK_OK_NEGATIVE_PHI 2 /* success with alpha_phi = 2^(m + sigma - l) (x - 2^l): the
                                       * reference's mpfr_fmod keeps the sign of a negative dividend
                                       * (src/sample.cpp:566-574), which happens for
                                       * j < |eta| 2^(m + sigma) / r only */
#define QB200_DIAGK_GAVE_UP 4       /* pivot not used up after 2^22 steps although delta_bound
                                     * allows more (a pivot within ~1e-7 of 1) */

/* n independent calls of sample_k_from_diagonal_j_eta_pivot(parameters, pivot[i], j[i], eta[i],
 * delta_bound, k, alpha_phi). Outputs (each may be NULL): k[i] (rows of k_limbs words);
 * alpha_phi[i] / 2^(m + sigma - l) as an unevaluated sum x_hi + x_lo of two doubles (the
 * binding scales it back: mpfr_set_d / mpfr_add_d / mpfr_mul_2si); delta[i] with
 * k = (k0 + delta) mod 2^l; status[i]. Error -42: a pivot outside [0, 1] (the reference:
 * critical("The pivot is out of bounds.")). */
int qb200_diagk_sample(qb200_diagk *sampler, uint32_t n, const uint32_t *j, const int32_t *eta,
                       const long double *pivot, uint32_t delta_bound, uint32_t *k, double *x_hi,
                       double *x_lo, int64_t *delta, int32_t *status);

/* Device-resident form: the rows of n samples in device memory (d_pivot: 16-byte long doubles),
 * launches on `stream` (NULL = the context's own), no synchronisation. d_out: n records of 32
 * bytes {double x_hi, x_lo; int64 delta; int32 status; int32 pad}; d_k (rows of k_limbs words)
 * may be NULL. Scratch of ~20 k bytes per sample (k = bytes of r) is held by the sampler. */
int qb200_diagk_sample_device(qb200_diagk *sampler, uint32_t n, const uint32_t *d_j,
                              const int32_t *d_eta, const long double *d_pivot,
                              uint32_t delta_bound, uint32_t *d_k, void *d_out, void *stream);

/* `count` estimates of n samples each from count * n rows (j, eta, pivot) in sample order:
 * tau[t] = log2(sum alpha_phi^2 / n) / 2 - (m + sigma - l) and ok[t] = 1, or DBL_MAX and 0 if
 * a sample of the estimate runs out of bounds or has |eta| > eta_bound
 * (src/tau_estimate.cpp:163-201). Error -43: a sample gave up, or has status 2 with l > 1000
 * (alpha_phi^2 ~ 2^(2 l) leaves the doubles). */
int qb200_diagk_tau_estimate(qb200_diagk *sampler, uint32_t n, uint32_t count, const uint32_t *j,
                             const int32_t *eta, const long double *pivot, uint32_t delta_bound,
                             uint32_t eta_bound, long double *tau, uint8_t *ok);

/* Test hook: on = 1 sends every walk through the exact path (h in double-double, rounded to and
 * subtracted in the x87 format) instead of deciding it in doubles with an error band first. */
int qb200_diagk_set_force_exact(qb200_diagk *sampler, int on);

/* diagonal_probability_approx_h at phi[i] = 2 pi (x_hi[i] + x_lo[i]) / 2^l, |x| <= 2^(l - 1),
 * rounded to long double as mpfr_get_ld does. */
int qb200_diagk_h(qb200_diagk *sampler, uint32_t n, const double *x_hi, const double *x_lo,
                  long double *h);

/* ---- exact samplers ---------------------------------------------------------
 * The reference draws an integer argument from a region of a slice and maps it to the pair (j, k)
 * with MPFR at 3 m bits and GMP, one sample per call:
 *   sample_alpha_from_region           src/sample.cpp:78-158   alpha = min + (v mod (max - min))
 *   sample_j_from_alpha_r              src/sample.cpp:160-208
 *   sample_j_k_from_alpha_d            src/sample.cpp:210-273
 *   sample_j_k_from_alpha_d_r          src/sample.cpp:275-352
 *   sample_j_from_diagonal_alpha_r     src/sample.cpp:354-410
 * (the second halves of linear_distribution_sample_alpha, src/linear_distribution.cpp:668-724,
 * diagonal_distribution_sample_alpha_r / _j_eta, src/diagonal_distribution.cpp:355-472, and
 * distribution_sample_pair_j_k, src/distribution.cpp:615-680). Here a batch of samples is one
 * call; the integers are the reference's bit for bit (qunundrum_b200/csrc/exact.cuh).
 *
 * A qb200_exact holds, for one set of parameters: (r / 2^kappa_r)^-1 and (d / 2^kappa_d)^-1 modulo
 * 2^n, d, and the table 2^(i / dimension_max) from which the bounds round(2^|log alpha|) of a region
 * are formed. kind QB200_EXACT_TWO_DIMENSIONAL: n = m + l, k has l bits (Parameters; also the
 * linear distributions); QB200_EXACT_DIAGONAL: n = m + sigma, no k (Diagonal_Parameters; k comes
 * from qb200_diagk). dimension_max: the largest slice dimension, a power of two <= 16384. emax:
 * bounds up to 2^emax (0: m + 64; |alpha| < 2^emax travels in alpha_limbs = ceil((emax + 1) / 32)
 * words). Limits: m >= 8, n <= 32768, kappa_d, kappa_r <= 64, regions with |min_log_alpha| >= 8.
 *
 * Integers travel as little-endian 32-bit words, one row per sample, zero padded: |alpha| in
 * alpha_limbs words with the sign apart (negative[i] = 1: alpha < 0), j in j_limbs = ceil(n / 32)
 * words, k in k_limbs = ceil(l / 32) words, t in max(1, ceil(kappa / 32)) words. */
typedef struct qb200_exact qb200_exact;

#define QB200_EXACT_TWO_DIMENSIONAL 0
#define QB200_EXACT_DIAGONAL 1

int qb200_exact_create(qb200_context *ctx, const qb200_params *params, int kind,
                       uint32_t dimension_max, uint32_t emax, qb200_exact **sampler);
void qb200_exact_destroy(qb200_exact *sampler);
/* out: alpha_limbs, j_limbs, k_limbs, kappa_d, kappa_r, emax. */
void qb200_exact_dims(const qb200_exact *sampler, uint32_t out[6]);

/* Duration of the LAST k_exact_alpha (out[0]) and k_exact_jk (out[1]) launch in milliseconds, from
 * CUDA events recorded around them on the launching stream (for benchmarks; a call of more than
 * ~300,000 samples is several launches and this is the last one's). */
int qb200_exact_kernel_ms(qb200_exact *sampler, float out[2]);

/* One region of a slice as distribution_slice_region_coordinates (src/distribution_slice.cpp:130-165)
 * and its linear / diagonal twins give it: |log alpha| on [e + region / dimension,
 * e + (region + 1) / dimension] with e = |min_log_alpha|, the slice's coordinate, whose sign is
 * alpha's; the sample's random bytes are stream[offset, offset + length). */
typedef struct qb200_exact_region {
  int32_t min_log_alpha;
  uint32_t region;
  uint32_t dimension;
  uint32_t length;
  uint64_t offset;
} qb200_exact_region;

#define QB200_EXACT_OK 0
#define QB200_EXACT_LENGTH 1      /* length is not what random_generate_mpz reads for this region */
#define QB200_EXACT_AMBIGUOUS 2   /* a bound within 2^-64 of a half-integer (never observed) */
#define QB200_EXACT_UNSUPPORTED 3 /* |min_log_alpha| < 8 or >= emax, dimension not a power of two
                                   * up to dimension_max, region >= dimension */

/* The bytes random_generate_mpz (src/random.c:158-181) reads for a sample of this region:
 * (bits(max - min) + 72) / 8. Stream layout, computed on the host (the caller cannot cut a
 * sample's bytes out of its random stream without it). Errors: -50 unsupported, -51 ambiguous. */
int qb200_exact_region_bytes(const qb200_exact *sampler, int32_t min_log_alpha, uint32_t region,
                             uint32_t dimension, uint32_t *bytes);

/* n independent calls of sample_alpha_from_region(alpha, min_log_alpha, max_log_alpha, kappa,
 * random_state) with the bytes the Random_State would deliver given in `stream`. status[i] != 0:
 * alpha[i] is not valid. */
int qb200_exact_alpha(qb200_exact *sampler, uint32_t n, const qb200_exact_region *regions,
                      uint32_t kappa, const uint8_t *stream, uint64_t stream_len, uint32_t *alpha,
                      int32_t *negative, int32_t *status);

#define QB200_EXACT_J_FROM_ALPHA_R 0      /* sample_j_from_alpha_r / sample_j_from_diagonal_alpha_r */
#define QB200_EXACT_J_K_FROM_ALPHA_D_R 1  /* sample_j_k_from_alpha_d_r */
#define QB200_EXACT_J_FROM_ALPHA_D_K 2    /* sample_j_k_from_alpha_d (k drawn by the caller) */

/* n independent calls of the (j, k) sampler `mode` selects. What the reference draws inside them
 * is drawn by the caller (it is stream layout: fixed sizes per parameter set) and passed in:
 * t = t_r (modes 0, 1; for mode 1 already multiplied by 2^kappa_t_r, src/sample.cpp:294-309) or
 * t_d (mode 2), NULL when the kappa in question is 0; k (mode 2, input). Outputs: j; k (mode 1).
 * Arguments a mode does not use may be NULL. */
int qb200_exact_j_k(qb200_exact *sampler, int mode, uint32_t n, const uint32_t *alpha_d,
                    const int32_t *negative_d, const uint32_t *alpha_r, const int32_t *negative_r,
                    const uint32_t *t, uint32_t *j, uint32_t *k);

/* The diagonal distribution's sample from its random bytes to k without leaving the device
 * (diagonal_distribution_sample_pair_j_k, src/diagonal_distribution.cpp:474-552, after the region
 * is chosen): alpha_r from regions[i] (kappa = kappa_r), j = sample_j_from_diagonal_alpha_r
 * (t_r: rows, NULL for odd r), then qb200_diagk_sample's outputs for (j, eta[i], pivot[i]).
 * `exact` must be a QB200_EXACT_DIAGONAL sampler of the same parameters. exact_status[i] != 0
 * (QB200_EXACT_*): the sample's other outputs are not valid. */
int qb200_diagk_sample_drawn(qb200_diagk *sampler, qb200_exact *exact, uint32_t n,
                             const qb200_exact_region *regions, const uint32_t *t_r,
                             const uint8_t *stream, uint64_t stream_len, const int32_t *eta,
                             const long double *pivot, uint32_t delta_bound, uint32_t *k,
                             double *x_hi, double *x_lo, int64_t *delta, int32_t *status,
                             int32_t *exact_status);

/* ---- introspection (host logic; usable without a GPU) --------------------- */

/* sigma chosen by QB200_METHOD_HEURISTIC_SIGMA for this l
 * (src/distribution_slice_compute.cpp:149-158). */
uint32_t qb200_heuristic_sigma(uint32_t l);

/* The double-double constants derived from (m, l, sigma, d, r), as 10 (hi, lo)
 * pairs: kappa, kappa_quick, C/L, N/L, (N+1)/L, beta/2^m, (r-beta)/2^m, r/2^m,
 * d/2^m, 2^m/r. */
int qb200_host_constants(const qb200_params *params, double *out20);

/* Measured FP64 FMA throughput of the device (flop/s, FMA = 2), from a
 * register-resident DFMA loop; used as the roofline denominator. */
int qb200_measure_fp64_peak(qb200_context *ctx, double *flops_per_second);

#ifdef __cplusplus
}
#endif

#endif /* QUNUNDRUM_B200_H */
