// sampler_host.hpp -- the serial, device-independent part of qb200_sampler_tau_estimate
// (no CUDA in this file).
//
// What is inherently sequential in the reference's semantics and tiny in cost stays on the host:
//   * which word of the random stream every estimate starts at. tau_estimate()
//     (src/tau_estimate.cpp:23-87) stops at the first sample whose slice pivot runs past the
//     last slice (distribution_sample_slice() returns NULL after ONE draw,
//     src/distribution.cpp:359-409), so an estimate consumes 4 n (linear: 3 n) words or fewer;
//   * whether a pivot word does that: the walk's outcome is monotone in the word, so the smallest
//     failing word is found once per distribution by bisection with the reference's own loop in the
//     host's x87 long double arithmetic;
//   * tau = log2(mean alpha^2) / 2 - m in long double (src/tau_estimate.cpp:63-71).
// The CUDA library (qb200_sampler.cu) and the test-only CPU stand-in of tests/hostsim share it.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace qb200 {

#define QB_TAU_SKIP 0xffffffffffffffffull

struct FailureThreshold {
  bool any = false;      // some pivot word runs out of bounds
  uint64_t first = 0;    // the smallest such word
};

// distribution_sample_slice (src/distribution.cpp:359-409) for one pivot word, verbatim
// semantics: true if no slice is selected.
inline bool slice_walk_fails(const long double* totals, size_t n, long double total, uint64_t w) {
  long double pivot = (long double)w;
  pivot /= (long double)0xffffffffffffffffULL;
  if (total > 1) pivot *= total;
  for (size_t i = 0; i < n; i++) {
    pivot -= totals[i];
    if (pivot <= 0) return false;
  }
  return true;
}

inline FailureThreshold find_failure_threshold(const long double* totals, size_t n, long double total) {
  FailureThreshold t;
  if (!slice_walk_fails(totals, n, total, 0xffffffffffffffffULL)) return t;
  t.any = true;
  uint64_t lo = 0, hi = 0xffffffffffffffffULL;  // fails(hi); the smallest failing word is in [lo, hi]
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (slice_walk_fails(totals, n, total, mid))
      hi = mid;
    else
      lo = mid + 1;
  }
  t.first = lo;
  return t;
}

struct TauLayout {
  std::vector<uint64_t> off;  // first word of every estimate, QB_TAU_SKIP for one that fails
  size_t words_used = 0;
  uint32_t done = 0;          // estimates laid out before the words ran out
};

// Up to `count` estimates of n samples along words[0, n_words); wps = words per successful sample.
inline void tau_layout(const FailureThreshold& f, uint32_t wps, uint32_t n, uint32_t count,
                       const uint64_t* words, size_t n_words, TauLayout* out) {
  out->off.clear();
  out->off.reserve(count);
  size_t cur = 0;
  uint32_t t = 0;
  for (; t < count; t++) {
    size_t used = 0;
    bool fails = false, short_of_words = false;
    for (uint32_t i = 0; i < n; i++) {
      if (cur + used >= n_words) {
        short_of_words = true;
        break;
      }
      if (f.any && words[cur + used] >= f.first) {  // one draw, then FALSE
        used += 1;
        fails = true;
        break;
      }
      if (cur + used + wps > n_words) {
        short_of_words = true;
        break;
      }
      used += wps;
    }
    if (short_of_words) break;
    out->off.push_back(fails ? QB_TAU_SKIP : (uint64_t)cur);
    cur += used;
  }
  out->done = t;
  out->words_used = cur;
}

// sums: per estimate (sum x0^2 hi, lo, sum x1^2 hi, lo), x = alpha / 2^m; status: 0, or the
// status of the estimate's first failing sample (1 out of bounds, 2 no region). Returns 0, -40
// (the device found an out-of-bounds sample the layout did not predict) or -41 (the reference's
// "Failed to sample a region from the slice.").
inline int tau_finish(int dims, int m, uint32_t n, const TauLayout& lay, const double* sums,
                      const int* status, long double* tau0, long double* tau1, uint8_t* ok,
                      std::string* err) {
  const long double two_m = (long double)(2.0 * (double)m);
  for (uint32_t i = 0; i < lay.done; i++) {
    if (lay.off[i] == QB_TAU_SKIP) {
      ok[i] = 0;
      tau0[i] = DBL_MAX;
      if (tau1) tau1[i] = DBL_MAX;
      continue;
    }
    if (status[i] == 2) {
      *err = "Failed to sample a region from the slice.";
      return -41;
    }
    if (status[i] != 0) {
      *err = "internal error: the device and the host disagree on an out-of-bounds sample";
      return -40;
    }
    ok[i] = 1;
    const long double a = ((long double)sums[4 * i] + (long double)sums[4 * i + 1]) / (long double)n;
    tau0[i] = (two_m + log2l(a)) / 2 - (long double)m;
    if (dims == 2) {
      const long double b = ((long double)sums[4 * i + 2] + (long double)sums[4 * i + 3]) / (long double)n;
      tau1[i] = (two_m + log2l(b)) / 2 - (long double)m;
    } else if (tau1) {
      tau1[i] = 0;
    }
  }
  return 0;
}

}  // namespace qb200
