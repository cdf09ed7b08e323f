/* oracle/ref_capi_sample.cpp -- plain-C handle API over the UNMODIFIED reference's sampling
 * path.
 *
 * TEST INFRASTRUCTURE ONLY (see ref_capi.cpp): linked into oracle/_ref/libqref.so, which only
 * tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py load.
 *
 *   reference function                                   (file:line)
 *   distribution_init / _insert_slice / _sort_slices     src/distribution.cpp:40,159,258
 *   distribution_sample_slice / _region                  src/distribution.cpp:359,411
 *   distribution_sample_approximate_alpha_d_r            src/distribution.cpp:464
 *   distribution_slice_sample_region                     src/distribution_slice.cpp:167
 *   linear_distribution_init / _insert_slice             src/linear_distribution.cpp:47,434
 *   linear_distribution_sample_approximate_alpha         src/linear_distribution.cpp:618
 *   sample_approximate_alpha_from_region                 src/sample.cpp:24
 *   tau_estimate / tau_estimate_linear                   src/tau_estimate.cpp:23,89
 *   sample_k_from_diagonal_j_eta_pivot                   src/sample.cpp:412-646
 *   sample_alpha_from_region                             src/sample.cpp:78-158
 *   sample_j_from_alpha_r / sample_j_k_from_alpha_d[_r]  src/sample.cpp:160-352
 *   sample_j_from_diagonal_alpha_r                       src/sample.cpp:354-410
 *   diagonal_probability_approx_h                        src/diagonal_probability.cpp:99-162
 *   random_generate / keccak_random_init_seed            src/random.c:88, src/keccak_random.c:52
 *
 * fpLLL is not in the image: lattice_alpha_init / _clear (called by distribution_init, never by
 * the sampling functions above) come from integration/stubs/lattice_stub.cpp.
 */
#include "common.h"
#include "diagonal_parameters.h"
#include "diagonal_probability.h"
#include "distribution.h"
#include "distribution_slice.h"
#include "keccak_random.h"
#include "linear_distribution.h"
#include "linear_distribution_slice.h"
#include "parameters.h"
#include "random.h"
#include "sample.h"
#include "tau_estimate.h"

#include <gmp.h>
#include <mpfr.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace {

struct Dist {
  int dims;
  Distribution d2;
  Linear_Distribution d1;
};

void ensure_precision() { mpfr_set_default_prec(PRECISION); }

}  // namespace

extern "C" {

/* dims = 2: Distribution of n slices (c0 = min_log_alpha_d, c1 = min_log_alpha_r, dimension[i]^2
 * cells each, concatenated in `cells`); dims = 1: Linear_Distribution (c0 = min_log_alpha,
 * dimension[i] cells each). totals: the slices' total_probability (NULL: the sum of the cells in
 * index order, as the importers compute it). */
void *qref_dist_new(int dims, void *params, uint32_t n, const uint32_t *dimension, const int32_t *c0,
                    const int32_t *c1, const long double *cells, const long double *totals) {
  ensure_precision();
  Dist *h = (Dist *)calloc(1, sizeof(Dist));
  h->dims = dims;
  const Parameters *p = (const Parameters *)params;
  size_t off = 0;
  if (dims == 2) {
    distribution_init(&h->d2, p, n ? n : 1);
    for (uint32_t i = 0; i < n; i++) {
      Distribution_Slice *s = distribution_slice_alloc();
      distribution_slice_init(s, dimension[i]);
      const size_t nc = (size_t)dimension[i] * dimension[i];
      memcpy(s->norm_matrix, cells + off, nc * sizeof(long double));
      off += nc;
      s->min_log_alpha_d = c0[i];
      s->min_log_alpha_r = c1[i];
      long double t = 0;
      for (size_t k = 0; k < nc; k++) t += s->norm_matrix[k];
      s->total_probability = totals ? totals[i] : t;
      s->total_error = 0;
      distribution_insert_slice(&h->d2, s);
    }
  } else {
    linear_distribution_init(&h->d1, p, 0, n ? n : 1);
    for (uint32_t i = 0; i < n; i++) {
      Linear_Distribution_Slice *s = linear_distribution_slice_alloc();
      linear_distribution_slice_init(s, dimension[i]);
      const size_t nc = dimension[i];
      memcpy(s->norm_vector, cells + off, nc * sizeof(long double));
      off += nc;
      s->min_log_alpha = c0[i];
      long double t = 0;
      for (size_t k = 0; k < nc; k++) t += s->norm_vector[k];
      s->total_probability = totals ? totals[i] : t;
      s->total_error = 0;
      linear_distribution_insert_slice(&h->d1, s);
    }
  }
  return h;
}

void qref_dist_free(void *hh) {
  Dist *h = (Dist *)hh;
  if (!h) return;
  if (h->dims == 2)
    distribution_clear(&h->d2);
  else
    linear_distribution_clear(&h->d1);
  free(h);
}

/* The reference's own container (Distribution * or Linear_Distribution *), to hand to code
 * that takes the reference's types (the drop-in translation unit under test). */
void *qref_dist_ptr(void *hh) {
  Dist *h = (Dist *)hh;
  return h->dims == 2 ? (void *)&h->d2 : (void *)&h->d1;
}

void qref_dist_sort(void *hh) {
  Dist *h = (Dist *)hh;
  if (h->dims == 2)
    distribution_sort_slices(&h->d2);
  else
    linear_distribution_sort_slices(&h->d1);
}

/* The slices in their current order: coordinates, dimension and total_probability. */
void qref_dist_describe(void *hh, uint32_t *dimension, int32_t *c0, int32_t *c1, long double *totals,
                        long double *total) {
  Dist *h = (Dist *)hh;
  if (h->dims == 2) {
    for (uint32_t i = 0; i < h->d2.count; i++) {
      dimension[i] = h->d2.slices[i]->dimension;
      c0[i] = h->d2.slices[i]->min_log_alpha_d;
      c1[i] = h->d2.slices[i]->min_log_alpha_r;
      totals[i] = h->d2.slices[i]->total_probability;
    }
    *total = h->d2.total_probability;
  } else {
    for (uint32_t i = 0; i < h->d1.count; i++) {
      dimension[i] = h->d1.slices[i]->dimension;
      c0[i] = h->d1.slices[i]->min_log_alpha;
      c1[i] = 0;
      totals[i] = h->d1.slices[i]->total_probability;
    }
    *total = h->d1.total_probability;
  }
}

/* linear_distribution_init_collapse_d (axis 0) / _r (axis 1) of a dims = 2 handle
 * (src/linear_distribution.cpp:152-324): the destination slices in the order the reference
 * creates them. Returns their number; coords / totals: up to cap entries; vectors: cap *
 * *max_dimension long doubles (the first call may pass cap = 0 to learn the sizes). */
uint32_t qref_dist_collapse(void *hh, int axis, uint32_t cap, int32_t *coords, long double *vectors,
                            long double *totals, uint32_t *max_dimension) {
  Dist *h = (Dist *)hh;
  Linear_Distribution dst;
  if (axis == 0)
    linear_distribution_init_collapse_d(&dst, &h->d2);
  else
    linear_distribution_init_collapse_r(&dst, &h->d2);
  const uint32_t n = dst.count;
  const uint32_t md = n ? dst.slices[0]->dimension : 0;
  *max_dimension = md;
  for (uint32_t i = 0; i < n && i < cap; i++) {
    coords[i] = dst.slices[i]->min_log_alpha;
    totals[i] = dst.slices[i]->total_probability;
    memcpy(vectors + (size_t)i * md, dst.slices[i]->norm_vector, (size_t)md * sizeof(long double));
  }
  linear_distribution_clear(&dst);
  return n;
}

/* Overwrite the distribution's total_probability (to exercise the "> 1" branch of
 * distribution_sample_slice, src/distribution.cpp:373-381). */
void qref_dist_set_total(void *hh, long double total) {
  Dist *h = (Dist *)hh;
  if (h->dims == 2)
    h->d2.total_probability = total;
  else
    h->d1.total_probability = total;
}

/* A Random_State expanded from a 32-byte seed (keccak_random_init_seed). */
void *qref_random_new(const uint8_t *seed) {
  Random_State *rs = (Random_State *)calloc(1, sizeof(Random_State));
  random_init(rs);
  keccak_random_init_seed(&rs->keccak_state, seed);
  return rs;
}

void qref_random_free(void *rs) {
  if (!rs) return;
  random_close((Random_State *)rs);
  free(rs);
}

void qref_random_bytes(void *rs, uint8_t *out, uint32_t n) { random_generate(out, n, (Random_State *)rs); }

/* k calls of distribution_sample_region / linear_distribution_sample_region: out[4 * i ...] =
 * min_log_alpha_d, max_log_alpha_d, min_log_alpha_r, max_log_alpha_r (linear: the first two).
 * Returns the number of successful samples. */
uint32_t qref_dist_sample_region(void *hh, void *rs, uint32_t k, double *out, uint8_t *ok) {
  Dist *h = (Dist *)hh;
  uint32_t good = 0;
  for (uint32_t i = 0; i < k; i++) {
    double *o = out + 4 * (size_t)i;
    o[0] = o[1] = o[2] = o[3] = 0;
    bool r;
    if (h->dims == 2)
      r = distribution_sample_region(&h->d2, (Random_State *)rs, &o[0], &o[1], &o[2], &o[3]);
    else
      r = linear_distribution_sample_region(&h->d1, (Random_State *)rs, &o[0], &o[1]);
    ok[i] = r ? 1 : 0;
    good += r ? 1 : 0;
  }
  return good;
}

/* k calls of distribution_sample_approximate_alpha_d_r / linear_..._alpha; alpha / 2^m as long
 * double (mpfr_get_ld of the 192-bit value scaled by 2^-m; infinity on failure). */
uint32_t qref_dist_sample_alpha(void *hh, void *rs, uint32_t k, long double *a0, long double *a1,
                                uint8_t *ok) {
  ensure_precision();
  Dist *h = (Dist *)hh;
  mpfr_t x, y;
  mpfr_init2(x, PRECISION);
  mpfr_init2(y, PRECISION);
  const uint32_t m = h->dims == 2 ? h->d2.parameters.m : h->d1.parameters.m;
  uint32_t good = 0;
  for (uint32_t i = 0; i < k; i++) {
    bool r;
    if (h->dims == 2) {
      r = distribution_sample_approximate_alpha_d_r(&h->d2, (Random_State *)rs, x, y);
      mpfr_mul_2si(x, x, -(long)m, MPFR_RNDN);
      mpfr_mul_2si(y, y, -(long)m, MPFR_RNDN);
      a0[i] = mpfr_get_ld(x, MPFR_RNDN);
      a1[i] = mpfr_get_ld(y, MPFR_RNDN);
    } else {
      r = linear_distribution_sample_approximate_alpha(&h->d1, (Random_State *)rs, x);
      mpfr_mul_2si(x, x, -(long)m, MPFR_RNDN);
      a0[i] = mpfr_get_ld(x, MPFR_RNDN);
      a1[i] = 0;
    }
    ok[i] = r ? 1 : 0;
    good += r ? 1 : 0;
  }
  mpfr_clear(x);
  mpfr_clear(y);
  return good;
}

/* count calls of tau_estimate / tau_estimate_linear with n samples each. */
void qref_tau_estimate(void *hh, void *rs, uint32_t n, uint32_t count, long double *tau0,
                       long double *tau1, uint8_t *ok) {
  ensure_precision();
  Dist *h = (Dist *)hh;
  for (uint32_t i = 0; i < count; i++) {
    bool r;
    if (h->dims == 2) {
      r = tau_estimate(&h->d2, (Random_State *)rs, n, tau0[i], tau1[i]);
    } else {
      r = tau_estimate_linear(&h->d1, (Random_State *)rs, n, tau0[i]);
      tau1[i] = 0;
    }
    ok[i] = r ? 1 : 0;
  }
}

/* sample_k_from_diagonal_j_eta_pivot (src/sample.cpp:412) for one (j, eta, pivot): k in decimal,
 * alpha_phi as a long double after scaling by 2^-(m + sigma - l) at `precision` bits, and with
 * 60 significant digits (unscaled) in alpha_out. Returns 1 on success, 0 when the reference
 * returns FALSE, -1 if a buffer is too small. precision = 0: 2 l as in the reference's KAT test
 * (src/test/test_sample.cpp:717). */
int qref_sample_k_from_diagonal(void *params, long double pivot, const char *j_dec, int32_t eta,
                                uint32_t delta_bound, uint32_t precision, char *k_out, size_t k_cap,
                                long double *alpha_scaled, char *alpha_out, size_t alpha_cap) {
  ensure_precision();
  const Diagonal_Parameters *p = (const Diagonal_Parameters *)params;
  if (0 == precision) precision = 2 * p->l;
  if (precision < 64) precision = 64;
  mpz_t j, k;
  mpz_init(j);
  mpz_init(k);
  mpz_set_str(j, j_dec, 10);
  mpfr_t alpha;
  mpfr_init2(alpha, precision);
  const bool ok = sample_k_from_diagonal_j_eta_pivot(p, pivot, j, eta, delta_bound, k, alpha);
  int rc = ok ? 1 : 0;
  if (mpz_sizeinbase(k, 10) + 2 > k_cap) {
    rc = -1;
  } else {
    mpz_get_str(k_out, 10, k);
  }
  if (alpha_out) mpfr_snprintf(alpha_out, alpha_cap, "%.60Re", alpha);
  mpfr_mul_2si(alpha, alpha, -((long)p->m + (long)p->sigma - (long)p->l), MPFR_RNDN);
  *alpha_scaled = mpfr_get_ld(alpha, MPFR_RNDN);
  mpfr_clear(alpha);
  mpz_clear(j);
  mpz_clear(k);
  return rc;
}

/* diagonal_probability_approx_h (src/diagonal_probability.cpp:99) at phi given in decimal, read
 * at `precision` bits (the KAT test uses 3 l, src/test/test_diagonal_probability.cpp:212). */
long double qref_diagonal_probability_h(void *params, const char *phi_dec, uint32_t precision) {
  ensure_precision();
  const Diagonal_Parameters *p = (const Diagonal_Parameters *)params;
  if (0 == precision) precision = 3 * p->l;
  if (precision < 64) precision = 64;
  mpfr_t phi, norm;
  mpfr_init2(phi, precision);
  mpfr_init2(norm, precision);
  mpfr_set_str(phi, phi_dec, 10, MPFR_RNDN);
  diagonal_probability_approx_h(norm, phi, p);
  const long double out = mpfr_get_ld(norm, MPFR_RNDN);
  mpfr_clear(phi);
  mpfr_clear(norm);
  return out;
}

/* sample_alpha_from_region (src/sample.cpp:78-158): alpha as a signed hexadecimal string.
 * Returns 0, or -1 if the buffer is too small. */
int qref_sample_alpha_from_region(double min_log_alpha, double max_log_alpha, uint32_t kappa, void *rs,
                                  char *out, size_t cap) {
  ensure_precision();
  mpz_t alpha;
  mpz_init(alpha);
  sample_alpha_from_region(alpha, min_log_alpha, max_log_alpha, kappa, (Random_State *)rs);
  int rc = 0;
  if (mpz_sizeinbase(alpha, 16) + 3 > cap)
    rc = -1;
  else
    mpz_get_str(out, 16, alpha);
  mpz_clear(alpha);
  return rc;
}

/* The (j, k) samplers on given arguments (signed hexadecimal strings), drawing from rs what the
 * reference draws:
 *   mode 0  sample_j_from_alpha_r           (Parameters; alpha_r)              -> j
 *   mode 1  sample_j_k_from_alpha_d_r       (Parameters; alpha_d, alpha_r)     -> j, k
 *   mode 2  sample_j_k_from_alpha_d         (Parameters; alpha_d)              -> j, k (k drawn)
 *   mode 3  sample_j_from_diagonal_alpha_r  (Diagonal_Parameters; alpha_r)     -> j
 * Returns 0, or -1 if a buffer is too small. */
int qref_sample_j_k(int mode, void *params, const char *alpha_d_hex, const char *alpha_r_hex, void *rs,
                    char *j_out, char *k_out, size_t cap) {
  ensure_precision();
  mpz_t ad, ar, j, k;
  mpz_init(ad);
  mpz_init(ar);
  mpz_init(j);
  mpz_init(k);
  if (alpha_d_hex) mpz_set_str(ad, alpha_d_hex, 16);
  if (alpha_r_hex) mpz_set_str(ar, alpha_r_hex, 16);
  Random_State *r = (Random_State *)rs;
  switch (mode) {
    case 0:
      sample_j_from_alpha_r(j, ar, (const Parameters *)params, r);
      break;
    case 1:
      sample_j_k_from_alpha_d_r(j, k, ad, ar, (const Parameters *)params, r);
      break;
    case 2:
      sample_j_k_from_alpha_d(j, k, ad, (const Parameters *)params, r);
      break;
    default:
      sample_j_from_diagonal_alpha_r(j, ar, (const Diagonal_Parameters *)params, r);
      break;
  }
  int rc = 0;
  if (mpz_sizeinbase(j, 16) + 3 > cap || mpz_sizeinbase(k, 16) + 3 > cap) {
    rc = -1;
  } else {
    mpz_get_str(j_out, 16, j);
    mpz_get_str(k_out, 16, k);
  }
  mpz_clear(ad);
  mpz_clear(ar);
  mpz_clear(j);
  mpz_clear(k);
  return rc;
}

} /* extern "C" */
