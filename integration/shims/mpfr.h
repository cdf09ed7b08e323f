/* Declaration-only stand-in for <mpfr.h> (MPFR 4.x ABI, x86-64 / LP64).
 *
 * TEST INFRASTRUCTURE ONLY.  This image ships libmpfr.so.6 but not its
 * development header.  Only the subset of the documented MPFR API used by the
 * reference's hot-path sources (and by oracle/ref_capi.cpp) is declared.
 */
#ifndef QUNUNDRUM_B200_SHIM_MPFR_H
#define QUNUNDRUM_B200_SHIM_MPFR_H

#include <gmp.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long mpfr_prec_t;
typedef int mpfr_sign_t;
typedef long mpfr_exp_t;

typedef struct {
  mpfr_prec_t _mpfr_prec;
  mpfr_sign_t _mpfr_sign;
  mpfr_exp_t _mpfr_exp;
  mp_limb_t *_mpfr_d;
} __mpfr_struct;

typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct *mpfr_ptr;
typedef const __mpfr_struct *mpfr_srcptr;

typedef enum {
  MPFR_RNDN = 0,
  MPFR_RNDZ,
  MPFR_RNDU,
  MPFR_RNDD,
  MPFR_RNDA,
  MPFR_RNDF,
  MPFR_RNDNA = -1
} mpfr_rnd_t;

void mpfr_set_default_prec(mpfr_prec_t);
mpfr_prec_t mpfr_get_default_prec(void);
void mpfr_init2(mpfr_ptr, mpfr_prec_t);
void mpfr_init(mpfr_ptr);
void mpfr_clear(mpfr_ptr);
void mpfr_set_prec(mpfr_ptr, mpfr_prec_t);
mpfr_prec_t mpfr_get_prec(mpfr_srcptr);

/* mpfr_set is a macro over mpfr_set4 in the real header. */
int mpfr_set4(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t, int);
#define mpfr_set(a, b, r) mpfr_set4((a), (b), (r), (b)->_mpfr_sign)
#define mpfr_abs(a, b, r) mpfr_set4((a), (b), (r), 1)
int mpfr_set_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_set_ld(mpfr_ptr, long double, mpfr_rnd_t);
int mpfr_set_z(mpfr_ptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_set_ui_2exp(mpfr_ptr, unsigned long, mpfr_exp_t, mpfr_rnd_t);
int mpfr_set_str(mpfr_ptr, const char *, int, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);

double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
long double mpfr_get_ld(mpfr_srcptr, mpfr_rnd_t);
int mpfr_get_z(mpz_ptr, mpfr_srcptr, mpfr_rnd_t);
mpfr_exp_t mpfr_get_z_2exp(mpz_ptr, mpfr_srcptr);
char *mpfr_get_str(char *, mpfr_exp_t *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
void mpfr_free_str(char *);

int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_add_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_mul_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_sqr(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_div_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_div_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_add_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_sub_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_div_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_add_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_sub_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_fmod(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
void mpfr_set_inf(mpfr_ptr, int);
void mpfr_set_nan(mpfr_ptr);
void mpfr_set_zero(mpfr_ptr, int);
int mpfr_printf(const char *, ...);
#define mpfr_fprintf __gmpfr_fprintf /* as <mpfr.h> does: the library exports the prefixed name */
int mpfr_fprintf(FILE *, const char *, ...);
int mpfr_sprintf(char *, const char *, ...);
int mpfr_snprintf(char *, size_t, const char *, ...);

int mpfr_sin(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_cos(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_exp2(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_log2(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_gamma(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_pow(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_const_pi(mpfr_ptr, mpfr_rnd_t);
int mpfr_const_catalan(mpfr_ptr, mpfr_rnd_t);

int mpfr_cmp3(mpfr_srcptr, mpfr_srcptr, int);
#define mpfr_cmp(a, b) mpfr_cmp3((a), (b), 1)
int mpfr_cmp_ui_2exp(mpfr_srcptr, unsigned long, mpfr_exp_t);
#define mpfr_cmp_ui(a, u) mpfr_cmp_ui_2exp((a), (u), 0)
int mpfr_cmp_si_2exp(mpfr_srcptr, long, mpfr_exp_t);
#define mpfr_cmp_si(a, s) mpfr_cmp_si_2exp((a), (s), 0)
int mpfr_cmp_d(mpfr_srcptr, double);
int mpfr_cmp_ld(mpfr_srcptr, long double);
int mpfr_sgn(mpfr_srcptr);
int mpfr_zero_p(mpfr_srcptr);
int mpfr_nan_p(mpfr_srcptr);
int mpfr_inf_p(mpfr_srcptr);

int mpfr_round(mpfr_ptr, mpfr_srcptr);
int mpfr_ceil(mpfr_ptr, mpfr_srcptr);
int mpfr_floor(mpfr_ptr, mpfr_srcptr);
int mpfr_trunc(mpfr_ptr, mpfr_srcptr);

const char *mpfr_get_version(void);

#ifdef __cplusplus
}
#endif

#endif /* QUNUNDRUM_B200_SHIM_MPFR_H */
