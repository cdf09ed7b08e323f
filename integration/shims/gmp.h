/* Declaration-only stand-in for <gmp.h> (GMP 6.x ABI, x86-64 / LP64).
 *
 * TEST INFRASTRUCTURE ONLY.  This image ships libgmp.so.10 but not its
 * development header.  The declarations below cover exactly the subset of the
 * documented GMP API that the reference's hot-path sources and the drop-in
 * layer use, with the struct layout and the __gmpz_* link names of the
 * installed library.  On a machine with libgmp-dev, the real header is used
 * instead (put nothing from this directory on the include path).
 */
#ifndef QUNUNDRUM_B200_SHIM_GMP_H
#define QUNUNDRUM_B200_SHIM_GMP_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t;
typedef long mp_exp_t;

typedef struct {
  int _mp_alloc;
  int _mp_size;
  mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define mpz_init __gmpz_init
void __gmpz_init(mpz_ptr);
#define mpz_init_set __gmpz_init_set
void __gmpz_init_set(mpz_ptr, mpz_srcptr);
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
void __gmpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_fdiv_r_2exp __gmpz_fdiv_r_2exp
void __gmpz_fdiv_r_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_init_set_ui __gmpz_init_set_ui
void __gmpz_init_set_ui(mpz_ptr, unsigned long);
#define mpz_clear __gmpz_clear
void __gmpz_clear(mpz_ptr);
#define mpz_set __gmpz_set
void __gmpz_set(mpz_ptr, mpz_srcptr);
#define mpz_set_ui __gmpz_set_ui
void __gmpz_set_ui(mpz_ptr, unsigned long);
#define mpz_set_si __gmpz_set_si
void __gmpz_set_si(mpz_ptr, long);
#define mpz_set_str __gmpz_set_str
int __gmpz_set_str(mpz_ptr, const char *, int);
#define mpz_get_str __gmpz_get_str
char *__gmpz_get_str(char *, int, mpz_srcptr);
#define mpz_get_ui __gmpz_get_ui
unsigned long __gmpz_get_ui(mpz_srcptr);
#define mpz_get_d __gmpz_get_d
double __gmpz_get_d(mpz_srcptr);
#define mpz_setbit __gmpz_setbit
void __gmpz_setbit(mpz_ptr, mp_bitcnt_t);
#define mpz_clrbit __gmpz_clrbit
void __gmpz_clrbit(mpz_ptr, mp_bitcnt_t);
#define mpz_tstbit __gmpz_tstbit
int __gmpz_tstbit(mpz_srcptr, mp_bitcnt_t);
#define mpz_sizeinbase __gmpz_sizeinbase
size_t __gmpz_sizeinbase(mpz_srcptr, int);
#define mpz_add __gmpz_add
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_add_ui __gmpz_add_ui
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_sub __gmpz_sub
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_sub_ui __gmpz_sub_ui
void __gmpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul __gmpz_mul
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul_si __gmpz_mul_si
void __gmpz_mul_si(mpz_ptr, mpz_srcptr, long);
#define mpz_div __gmpz_fdiv_q
#define mpz_invert __gmpz_invert
int __gmpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul_ui __gmpz_mul_ui
void __gmpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul_2exp __gmpz_mul_2exp
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_neg __gmpz_neg
void __gmpz_neg(mpz_ptr, mpz_srcptr);
#define mpz_abs __gmpz_abs
void __gmpz_abs(mpz_ptr, mpz_srcptr);
#define mpz_mod __gmpz_mod
void __gmpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_div_ui __gmpz_fdiv_q_ui
#define mpz_fdiv_q_ui __gmpz_fdiv_q_ui
unsigned long __gmpz_fdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mod_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_r_ui __gmpz_fdiv_r_ui
unsigned long __gmpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_fdiv_q __gmpz_fdiv_q
void __gmpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_cmp __gmpz_cmp
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
#define mpz_cmp_ui __gmpz_cmp_ui
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
#define mpz_cmp_si __gmpz_cmp_si
int __gmpz_cmp_si(mpz_srcptr, long);
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)
#define mpz_export __gmpz_export
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
#define mpz_import __gmpz_import
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
#define mpz_probab_prime_p __gmpz_probab_prime_p
int __gmpz_probab_prime_p(mpz_srcptr, int);

#define gmp_printf __gmp_printf
int __gmp_printf(const char *, ...);
#define gmp_fprintf __gmp_fprintf
int __gmp_fprintf(FILE *, const char *, ...);
#define gmp_sprintf __gmp_sprintf
int __gmp_sprintf(char *, const char *, ...);
#define gmp_fscanf __gmp_fscanf
int __gmp_fscanf(FILE *, const char *, ...);

#ifdef __cplusplus
}
#endif

#endif /* QUNUNDRUM_B200_SHIM_GMP_H */
