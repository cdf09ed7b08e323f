/* oracle/ref_capi.cpp -- plain-C handle API over the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load the library this
 * file is linked into (oracle/_ref/libqref.so).  Nothing under
 * qunundrum_b200/ links, loads or calls it.
 *
 * The library is built by oracle/Makefile from the reference's own sources
 * where they lie under /root/reference/src (never copied into this repo):
 * probability.cpp, linear_probability.cpp, diagonal_probability.cpp, the six
 * *_slice_compute*.cpp files, the *_slice.cpp containers, parameters*.cpp,
 * ...  This file only adds extern "C" entry points with plain pointers so
 * that Python (ctypes) can drive the reference:
 *
 *   reference function                                   (file:line)
 *   distribution_slice_compute[_richardson]              src/distribution_slice_compute.cpp:38,
 *                                                        src/distribution_slice_compute_richardson.cpp:17
 *   linear_distribution_slice_compute[_richardson]       src/linear_distribution_slice_compute.cpp:30,
 *                                                        src/linear_distribution_slice_compute_richardson.cpp:17
 *   diagonal_distribution_slice_compute[_richardson]     src/diagonal_distribution_slice_compute.cpp:30,
 *                                                        src/diagonal_distribution_slice_compute_richardson.cpp:17
 *   probability_approx / probability_approx_quick        src/probability.cpp:150,290
 *   linear_probability_d / linear_probability_r          src/linear_probability.cpp:21,170
 *   diagonal_probability_approx_f_eta                    src/diagonal_probability.cpp:18
 *   parameters_selection_deterministic_d_r               src/parameters_selection.cpp:21
 *   distribution_slice_export / _import                  src/distribution_slice_import_export.cpp:89,55
 *   linear_distribution_slice_export / _import           src/linear_distribution_slice_import_export.cpp:82,50
 *   diagonal_distribution_slice_export / _import         src/diagonal_distribution_slice_import_export.cpp:87,55
 */

#include "common.h"
#include "diagonal_distribution_slice.h"
#include "diagonal_parameters.h"
#include "diagonal_probability.h"
#include "distribution_slice.h"
#include "linear_distribution_slice.h"
#include "linear_probability.h"
#include "parameters.h"
#include "parameters_selection.h"
#include "probability.h"

#include <gmp.h>
#include <mpfr.h>
#include <mpi.h>

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- MPI is never used by the oracle: abort if anything reaches it. ------ */
extern "C" int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { abort(); }
extern "C" int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) {
  abort();
}
extern "C" int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm,
                        MPI_Status *) {
  abort();
}

static void ensure_precision() {
  /* As main() of every generator does (src/main_generate_distribution.cpp:1461). */
  mpfr_set_default_prec(PRECISION);
}

static void mpfr_to_str(char *out, size_t cap, const mpfr_t x) {
  /* 50 significant digits in scientific notation: d.ddddde[+-]xxx */
  if (mpfr_zero_p(x)) {
    snprintf(out, cap, "0");
    return;
  }
  mpfr_exp_t e;
  char *s = mpfr_get_str(NULL, &e, 10, 50, x, MPFR_RNDN);
  const char *digits = s;
  const char *sign = "";
  if (*digits == '-') {
    sign = "-";
    digits++;
  }
  snprintf(out, cap, "%s%c.%se%ld", sign, digits[0], digits + 1, (long)(e - 1));
  mpfr_free_str(s);
}

extern "C" {

const char *qref_version(void) { return mpfr_get_version(); }

/* Deterministic (Catalan-constant) d and r as decimal strings. */
int qref_deterministic_d_r(uint32_t m, char *d_out, char *r_out, size_t cap) {
  ensure_precision();
  mpz_t d, r;
  mpz_init(d);
  mpz_init(r);
  parameters_selection_deterministic_d_r(d, r, m);
  int rc = 0;
  if (mpz_sizeinbase(d, 10) + 2 > cap || mpz_sizeinbase(r, 10) + 2 > cap) {
    rc = -1;
  } else {
    mpz_get_str(d_out, 10, d);
    mpz_get_str(r_out, 10, r);
  }
  mpz_clear(d);
  mpz_clear(r);
  return rc;
}

/* ---- Parameters ----------------------------------------------------------- */

/* l == 0: parameters_explicit_m_s (l = ceil(m / s)); else _m_l. */
void *qref_parameters_new(uint32_t m, uint32_t s, uint32_t l, uint32_t t,
                          const char *d_dec, const char *r_dec) {
  ensure_precision();
  mpz_t d, r;
  mpz_init(d);
  mpz_init(r);
  if (0 != mpz_set_str(d, d_dec, 10) || 0 != mpz_set_str(r, r_dec, 10)) {
    mpz_clear(d);
    mpz_clear(r);
    return NULL;
  }
  Parameters *p = (Parameters *)malloc(sizeof(Parameters));
  parameters_init(p);
  if (0 == l) {
    parameters_explicit_m_s(p, d, r, m, s, t);
  } else {
    parameters_explicit_m_l(p, d, r, m, l, t);
  }
  mpz_clear(d);
  mpz_clear(r);
  return p;
}

void qref_parameters_free(void *h) {
  Parameters *p = (Parameters *)h;
  parameters_clear(p);
  free(p);
}

void qref_parameters_get(void *h, uint32_t *out8) {
  const Parameters *p = (const Parameters *)h;
  out8[0] = p->m;
  out8[1] = p->l;
  out8[2] = p->s;
  out8[3] = p->t;
  out8[4] = p->min_alpha_d;
  out8[5] = p->max_alpha_d;
  out8[6] = p->min_alpha_r;
  out8[7] = p->max_alpha_r;
}

void *qref_diagonal_parameters_new(uint32_t m, uint32_t sigma, uint32_t s,
                                   uint32_t l, uint32_t eta_bound, uint32_t t,
                                   const char *d_dec, const char *r_dec) {
  ensure_precision();
  mpz_t d, r;
  mpz_init(d);
  mpz_init(r);
  if (0 != mpz_set_str(d, d_dec, 10) || 0 != mpz_set_str(r, r_dec, 10)) {
    mpz_clear(d);
    mpz_clear(r);
    return NULL;
  }
  Diagonal_Parameters *p =
      (Diagonal_Parameters *)malloc(sizeof(Diagonal_Parameters));
  diagonal_parameters_init(p);
  if (0 == l) {
    diagonal_parameters_explicit_m_s(p, d, r, m, sigma, s, eta_bound, t);
  } else {
    diagonal_parameters_explicit_m_l(p, d, r, m, sigma, l, eta_bound, t);
  }
  mpz_clear(d);
  mpz_clear(r);
  return p;
}

void qref_diagonal_parameters_free(void *h) {
  Diagonal_Parameters *p = (Diagonal_Parameters *)h;
  diagonal_parameters_clear(p);
  free(p);
}

/* ---- Slices ---------------------------------------------------------------- */

/* cells: dimension^2 long doubles, index = i_d + dimension * j_r. */
void qref_distribution_slice_compute(void *params, int richardson, int method,
                                     uint32_t dimension, int32_t min_log_alpha_d,
                                     int32_t min_log_alpha_r, long double *cells,
                                     long double *total_probability,
                                     long double *total_error, uint32_t *flags) {
  ensure_precision();
  Distribution_Slice slice;
  distribution_slice_init(&slice, dimension);
  if (richardson) {
    distribution_slice_compute_richardson(
        &slice, (const Parameters *)params,
        (Distribution_Slice_Compute_Method)method, min_log_alpha_d,
        min_log_alpha_r);
  } else {
    distribution_slice_compute(&slice, (const Parameters *)params,
                               (Distribution_Slice_Compute_Method)method,
                               min_log_alpha_d, min_log_alpha_r);
  }
  memcpy(cells, slice.norm_matrix,
         sizeof(long double) * (size_t)dimension * dimension);
  *total_probability = slice.total_probability;
  *total_error = slice.total_error;
  *flags = slice.flags;
  distribution_slice_clear(&slice);
}

void qref_linear_distribution_slice_compute(void *params, int richardson,
                                            int target, uint32_t dimension,
                                            int32_t min_log_alpha,
                                            long double *cells,
                                            long double *total_probability,
                                            long double *total_error,
                                            uint32_t *flags) {
  ensure_precision();
  Linear_Distribution_Slice slice;
  linear_distribution_slice_init(&slice, dimension);
  if (richardson) {
    linear_distribution_slice_compute_richardson(
        &slice, (const Parameters *)params,
        (Linear_Distribution_Slice_Compute_Target)target, min_log_alpha);
  } else {
    linear_distribution_slice_compute(
        &slice, (const Parameters *)params,
        (Linear_Distribution_Slice_Compute_Target)target, min_log_alpha);
  }
  memcpy(cells, slice.norm_vector, sizeof(long double) * (size_t)dimension);
  *total_probability = slice.total_probability;
  *total_error = slice.total_error;
  *flags = slice.flags;
  linear_distribution_slice_clear(&slice);
}

void qref_diagonal_distribution_slice_compute(void *dparams, int richardson,
                                              uint32_t dimension,
                                              int32_t min_log_alpha_r,
                                              int32_t eta, long double *cells,
                                              long double *total_probability,
                                              long double *total_error,
                                              uint32_t *flags) {
  ensure_precision();
  Diagonal_Distribution_Slice slice;
  diagonal_distribution_slice_init(&slice, dimension);
  if (richardson) {
    diagonal_distribution_slice_compute_richardson(
        &slice, (const Diagonal_Parameters *)dparams, min_log_alpha_r, eta);
  } else {
    diagonal_distribution_slice_compute(
        &slice, (const Diagonal_Parameters *)dparams, min_log_alpha_r, eta);
  }
  memcpy(cells, slice.norm_vector, sizeof(long double) * (size_t)dimension);
  *total_probability = slice.total_probability;
  *total_error = slice.total_error;
  *flags = slice.flags;
  diagonal_distribution_slice_clear(&slice);
}

/* distribution_slice_copy_scale (src/distribution_slice.cpp:229-264): what a client does to
 * a slice above MAX_SLICE_DIMENSION before sending it. Returns the scaled flags. */
uint32_t qref_distribution_slice_copy_scale(uint32_t src_dimension, const long double *src_cells,
                                            uint32_t src_flags, uint32_t dst_dimension,
                                            long double *dst_cells) {
  Distribution_Slice src, dst;
  distribution_slice_init(&src, src_dimension);
  distribution_slice_init(&dst, dst_dimension);
  memcpy(src.norm_matrix, src_cells, sizeof(long double) * (size_t)src_dimension * src_dimension);
  src.flags = src_flags;
  distribution_slice_copy_scale(&dst, &src);
  memcpy(dst_cells, dst.norm_matrix, sizeof(long double) * (size_t)dst_dimension * dst_dimension);
  const uint32_t flags = dst.flags;
  distribution_slice_clear(&src);
  distribution_slice_clear(&dst);
  return flags;
}

/* ---- Slice text export / import (memory streams) ---------------------------- */

/* kind: 0 two-dimensional (c0 = min_log_alpha_d, c1 = min_log_alpha_r),
 *       1 linear (c0 = min_log_alpha), 2 diagonal (c0 = min_log_alpha_r, c1 = eta).
 * Returns the number of bytes the reference's exporter wrote (<= cap), or
 * (size_t)-1 if cap is too small. */
size_t qref_slice_export(int kind, uint32_t dimension, int32_t c0, int32_t c1,
                         uint32_t flags, const long double *cells,
                         long double total_error, char *out, size_t cap) {
  char *buf = NULL;
  size_t len = 0;
  FILE *f = open_memstream(&buf, &len);
  if (kind == 0) {
    Distribution_Slice slice;
    distribution_slice_init(&slice, dimension);
    memcpy(slice.norm_matrix, cells, sizeof(long double) * (size_t)dimension * dimension);
    slice.min_log_alpha_d = c0;
    slice.min_log_alpha_r = c1;
    slice.flags = flags;
    slice.total_error = total_error;
    distribution_slice_export(&slice, f);
    distribution_slice_clear(&slice);
  } else if (kind == 1) {
    Linear_Distribution_Slice slice;
    linear_distribution_slice_init(&slice, dimension);
    memcpy(slice.norm_vector, cells, sizeof(long double) * (size_t)dimension);
    slice.min_log_alpha = c0;
    slice.flags = flags;
    slice.total_error = total_error;
    linear_distribution_slice_export(&slice, f);
    linear_distribution_slice_clear(&slice);
  } else {
    Diagonal_Distribution_Slice slice;
    diagonal_distribution_slice_init(&slice, dimension);
    memcpy(slice.norm_vector, cells, sizeof(long double) * (size_t)dimension);
    slice.min_log_alpha_r = c0;
    slice.eta = c1;
    slice.flags = flags;
    slice.total_error = total_error;
    diagonal_distribution_slice_export(&slice, f);
    diagonal_distribution_slice_clear(&slice);
  }
  fclose(f);
  size_t ret = (size_t)-1;
  if (len <= cap) {
    memcpy(out, buf, len);
    ret = len;
  }
  free(buf);
  return ret;
}

/* The reference's importer on a text held in memory. cells must have room for
 * dimension^2 (kind 0) / dimension values; head receives dimension, c0, c1, flags.
 * Returns 0 on success. */
int qref_slice_import(int kind, const char *text, size_t len, uint32_t max_cells,
                      uint32_t *head4, long double *cells, long double *total_probability,
                      long double *total_error) {
  FILE *f = fmemopen((void *)text, len, "rb");
  if (!f) return -1;
  int rc = 0;
  if (kind == 0) {
    Distribution_Slice slice;
    distribution_slice_init_import(&slice, f);
    const size_t n = (size_t)slice.dimension * slice.dimension;
    if (n > max_cells) rc = -2;
    else memcpy(cells, slice.norm_matrix, sizeof(long double) * n);
    head4[0] = slice.dimension; head4[1] = (uint32_t)slice.min_log_alpha_d;
    head4[2] = (uint32_t)slice.min_log_alpha_r; head4[3] = slice.flags;
    *total_probability = slice.total_probability; *total_error = slice.total_error;
    distribution_slice_clear(&slice);
  } else if (kind == 1) {
    Linear_Distribution_Slice slice;
    linear_distribution_slice_init_import(&slice, f);
    if (slice.dimension > max_cells) rc = -2;
    else memcpy(cells, slice.norm_vector, sizeof(long double) * slice.dimension);
    head4[0] = slice.dimension; head4[1] = (uint32_t)slice.min_log_alpha;
    head4[2] = 0; head4[3] = slice.flags;
    *total_probability = slice.total_probability; *total_error = slice.total_error;
    linear_distribution_slice_clear(&slice);
  } else {
    Diagonal_Distribution_Slice slice;
    diagonal_distribution_slice_init_import(&slice, f);
    if (slice.dimension > max_cells) rc = -2;
    else memcpy(cells, slice.norm_vector, sizeof(long double) * slice.dimension);
    head4[0] = slice.dimension; head4[1] = (uint32_t)slice.min_log_alpha_r;
    head4[2] = (uint32_t)slice.eta; head4[3] = slice.flags;
    *total_probability = slice.total_probability; *total_error = slice.total_error;
    diagonal_distribution_slice_clear(&slice);
  }
  fclose(f);
  return rc;
}

/* ---- Point-wise integrands (known-answer tests) ---------------------------- */

/* theta_* are decimal strings parsed at PRECISION bits exactly as
 * test_mpfr_load() does in src/test/test_common.cpp; outputs are decimal
 * strings with 50 significant digits. Returns the bounded-error flag. */
int qref_probability_approx(void *params, uint32_t sigma, const char *theta_d,
                            const char *theta_r, char *norm_out, char *error_out,
                            size_t cap) {
  ensure_precision();
  mpfr_t td, tr, norm, error;
  mpfr_init2(td, PRECISION);
  mpfr_init2(tr, PRECISION);
  mpfr_init2(norm, PRECISION);
  mpfr_init2(error, PRECISION);
  mpfr_set_str(td, theta_d, 10, MPFR_RNDN);
  mpfr_set_str(tr, theta_r, 10, MPFR_RNDN);
  const bool bounded =
      probability_approx(norm, error, sigma, td, tr, (const Parameters *)params);
  mpfr_to_str(norm_out, cap, norm);
  mpfr_to_str(error_out, cap, error);
  mpfr_clear(td);
  mpfr_clear(tr);
  mpfr_clear(norm);
  mpfr_clear(error);
  return bounded ? 1 : 0;
}

void qref_probability_approx_quick(void *params, const char *theta_d,
                                   const char *theta_r, char *norm_out,
                                   size_t cap) {
  ensure_precision();
  mpfr_t td, tr, norm;
  mpfr_init2(td, PRECISION);
  mpfr_init2(tr, PRECISION);
  mpfr_init2(norm, PRECISION);
  mpfr_set_str(td, theta_d, 10, MPFR_RNDN);
  mpfr_set_str(tr, theta_r, 10, MPFR_RNDN);
  probability_approx_quick(norm, td, tr, (const Parameters *)params);
  mpfr_to_str(norm_out, cap, norm);
  mpfr_clear(td);
  mpfr_clear(tr);
  mpfr_clear(norm);
}

/* target: 0 = d (linear_probability_d), 1 = r (linear_probability_r). */
void qref_linear_probability(void *params, int target, const char *theta,
                             char *norm_out, size_t cap) {
  ensure_precision();
  mpfr_t th, norm;
  mpfr_init2(th, PRECISION);
  mpfr_init2(norm, PRECISION);
  mpfr_set_str(th, theta, 10, MPFR_RNDN);
  if (0 == target) {
    linear_probability_d(norm, th, (const Parameters *)params);
  } else {
    linear_probability_r(norm, th, (const Parameters *)params);
  }
  mpfr_to_str(norm_out, cap, norm);
  mpfr_clear(th);
  mpfr_clear(norm);
}

/* alpha_r is an integer given in decimal; theta_r = 2 pi alpha_r / 2^(m+sigma)
 * formed at 2 max(m + sigma, 192) bits as the slice code does
 * (src/diagonal_distribution_slice_compute.cpp:41-69), or at theta_precision
 * bits when given (the reference's KAT uses 192,
 * src/test/test_diagonal_probability.cpp:97-127). */
void qref_diagonal_probability_f_eta(void *dparams, const char *alpha_r_dec,
                                     int32_t eta, uint32_t theta_precision,
                                     char *norm_out, size_t cap) {
  ensure_precision();
  const Diagonal_Parameters *p = (const Diagonal_Parameters *)dparams;
  /* theta_precision == 0: as the slice code; the reference's KAT uses 192. */
  uint32_t precision = theta_precision;
  if (0 == precision) {
    precision = 2 * (p->m + p->sigma);
    if (precision < 2 * PRECISION) {
      precision = 2 * PRECISION;
    }
  }
  mpz_t alpha;
  mpz_init(alpha);
  mpz_set_str(alpha, alpha_r_dec, 10);

  mpfr_t theta, tmp, norm;
  mpfr_init2(theta, precision);
  mpfr_init2(tmp, precision);
  mpfr_init2(norm, PRECISION);

  mpfr_const_pi(theta, MPFR_RNDN);
  mpfr_mul_ui(theta, theta, 2, MPFR_RNDN);
  mpfr_set_ui_2exp(tmp, 1, (mpfr_exp_t)(p->m + p->sigma), MPFR_RNDN);
  mpfr_div(theta, theta, tmp, MPFR_RNDN);
  mpfr_mul_z(theta, theta, alpha, MPFR_RNDN);

  diagonal_probability_approx_f_eta(norm, theta, eta, p);
  mpfr_to_str(norm_out, cap, norm);

  mpfr_clear(theta);
  mpfr_clear(tmp);
  mpfr_clear(norm);
  mpz_clear(alpha);
}

} /* extern "C" */
