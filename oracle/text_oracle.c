/* oracle/text_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * The reference's text format IS libc's: every cell of a slice is written by
 *   fprintf(file, "%.24Lg\n", slice->norm_matrix[i]);
 * (src/distribution_slice_import_export.cpp:99-103,
 *  src/linear_distribution_slice_import_export.cpp:92-96,
 *  src/diagonal_distribution_slice_import_export.cpp:98-102)
 * and read back with fscanf(file, "%Lg\n", ...) (same files, :38-50 / :33-45 / :38-50).
 * The arithmetic therefore lives in a third-party dependency that is not under
 * /root/reference: GNU libc (this image: glibc 2.39, stdio-common/printf_fp.c and
 * stdlib/strtod_l.c) -- exact decimal expansion, round-half-even in the default
 * rounding mode. This file makes exactly those calls so that the CUDA formatter
 * and parser can be compared with them byte for byte / bit for bit.
 * Built by oracle/Makefile (target text) and by __graft_entry__.build().
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* The exporter's inner loop over n values; returns the number of bytes written
 * to out (capacity cap; 34 bytes per value always suffice). */
size_t text_oracle_format_ld(const long double *v, size_t n, char *out, size_t cap) {
  size_t pos = 0;
  for (size_t i = 0; i < n; i++) {
    if (cap - pos < 40) return (size_t)-1;
    pos += (size_t)snprintf(out + pos, cap - pos, "%.24Lg\n", v[i]);
  }
  return pos;
}

/* The importer's inner loop: n values by fscanf("%Lg\n") from a memory stream.
 * Returns the number of values parsed. */
size_t text_oracle_parse_ld(const char *text, size_t len, long double *v, size_t n) {
  FILE *f = fmemopen((void *)text, len, "rb");
  if (!f) return 0;
  size_t i = 0;
  for (; i < n; i++)
    if (1 != fscanf(f, "%Lg\n", &v[i])) break;
  fclose(f);
  return i;
}

const char *text_oracle_libc(void) {
#ifdef __GLIBC__
  static char buf[64];
  snprintf(buf, sizeof buf, "glibc %d.%d", __GLIBC__, __GLIBC_MINOR__);
  return buf;
#else
  return "unknown libc";
#endif
}
