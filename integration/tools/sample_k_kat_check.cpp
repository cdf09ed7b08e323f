// sample_k_kat_check.cpp -- TEST DRIVER: the reference's OWN known-answer tests of the diagonal k sampler
// and of its integrand h (test_diagonal_probability_h_approx_kat(), src/test/test_diagonal_probability.cpp:174-298,
// against the drop-in's diagonal_probability_approx_h):
// test_sample_k_from_diagonal_j_eta_pivot_kat() (src/test/test_sample.cpp:679-836: all 522 files of
// res/test-vectors, 25 records each, k compared with mpz_cmp and alpha_phi with test_cmp_ld), run
// against the drop-in's sample_k_from_diagonal_j_eta_pivot (qunundrum_b200/dropin/dropin_tau_diagonal.cpp
// built with -DQB200_DROPIN_SAMPLE_K) instead of the reference's (sample.cpp compiled with the rename
// -Dsample_k_from_diagonal_j_eta_pivot=sample_k_from_diagonal_j_eta_pivot_cpu_unused).
//
// Run from the reference's root directory (the test opens res/test-vectors/... relative to it). A
// mismatch ends in the reference's critical() (exit code != 0); success prints "ok".
//
// One thing had to be supplied: the reference's test_cmp_ld() (src/test/test_common.cpp:75-93) calls
// critical() when a value is negative, and alpha_phi is negative in half of the records -- the test
// function is not reachable from the reference's test_sample() (src/test/test_sample.cpp:838-846)
// and fails on the first file with the reference's OWN sampler, too (checked: this driver linked
// against the unmodified sample.cpp). test_common.cpp is therefore compiled with
// -Dtest_cmp_ld=test_cmp_ld_of_the_reference and the comparison below -- the same relative
// criterion on the magnitudes, plus equal signs -- takes its place.
#include "common.h"
#include "test/test_common.h"

#include <math.h>

void test_sample_k_from_diagonal_j_eta_pivot_kat();  // src/test/test_sample.cpp:679 (not in test_sample.h)
void test_diagonal_probability_h_approx_kat();       // src/test/test_diagonal_probability.cpp:174

bool test_cmp_ld(const long double a, const long double b, const long double tolerance) {
  if (a == b) return TRUE;
  if ((a < 0) != (b < 0)) return FALSE;
  const long double x = fabsl(a), y = fabsl(b);
  return (fabsl(x - y) / (x < y ? x : y) < tolerance) ? TRUE : FALSE;
}

#include <mpfr.h>

#include <stdio.h>

int main() {
  mpfr_set_default_prec(PRECISION);
  test_sample_k_from_diagonal_j_eta_pivot_kat();
  // ... and the one of its integrand: diagonal_probability_approx_h on the 522 files
  // diagonal-probabilities-h-det-*, test_cmp_ld on the long double values
  test_diagonal_probability_h_approx_kat();
  printf("ok\n");
  return 0;
}
