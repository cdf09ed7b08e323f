// kernels_diagk.cuh -- kernels of the diagonal k sampler (diagk.cuh has the per-sample code and
// the reference citations).
//
//   k_diagk_gather   j of a chunk from the caller's row-per-sample layout to tiles of 128 samples
//                    with the limbs interleaved (word i of the t-th sample of a tile at
//                    [i * 128 + t]) so that the threads of a warp, which all work on the same limb
//                    index at the same time, read consecutive words at a compile-time stride.
//   k_diagk          one thread per sample: r j, d (q + eta) mod r, divmod(2^l w, r) in 32-bit
//                    limbs (about 5 k^2 multiply-adds, k = limbs of r), then the walk over delta:
//                    the first steps per thread, the rest of a long walk (a pivot near 1 takes
//                    thousands of steps, up to 2 delta_bound + 1) thirty-two steps at a time
//                    across the warp, so that one such sample does not hold its CTA for
//                    milliseconds. r, d and the Barrett reciprocal are staged in shared
//                    memory (every thread reads the same limb: broadcast). Integer-pipe bound.
//   k_diagk_scatter  k of a chunk back to row-per-sample.
//   k_diagk_tau      one thread per estimate: sum of (alpha_phi / 2^(m+sigma-l))^2 in sample
//                    order (tau_estimate_diagonal, src/tau_estimate.cpp:135-210).
//   k_diagk_h        h at given x (diagonal_probability_approx_h), for the known-answer tests.
#pragma once

#include <cuda_runtime.h>

#include "diagk.cuh"
#include "sampler.cuh"

namespace qb200 {

// Samples are processed in tiles of QB_DIAGK_CTA (one CTA each); word i of the sample with index t
// inside tile b lies at [(b * w + i) * QB_DIAGK_CTA + t] (w = words per sample of the array).
#define QB_DIAGK_CTA 128

__global__ void k_diagk_gather(const uint32_t* __restrict__ rows, uint32_t w, uint32_t B,
                               uint32_t* __restrict__ tiles) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t tiles_n = (B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA;
  if (idx >= (uint64_t)tiles_n * w * QB_DIAGK_CTA) return;
  const uint32_t t = (uint32_t)(idx % QB_DIAGK_CTA);
  const uint64_t bi = idx / QB_DIAGK_CTA;
  const uint32_t i = (uint32_t)(bi % w), b = (uint32_t)(bi / w);
  const uint32_t g = b * QB_DIAGK_CTA + t;
  tiles[idx] = g < B ? rows[(size_t)g * w + i] : 0u;
}

__global__ void k_diagk_scatter(const uint32_t* __restrict__ tiles, uint32_t w, uint32_t B,
                                uint32_t* __restrict__ rows) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)w * B) return;
  const uint32_t g = (uint32_t)(idx / w), i = (uint32_t)(idx - (uint64_t)g * w);
  rows[idx] = tiles[((size_t)(g / QB_DIAGK_CTA) * w + i) * QB_DIAGK_CTA + g % QB_DIAGK_CTA];
}

struct DiagKOut {
  double x_hi, x_lo;   // alpha_phi / 2^(m + sigma - l)
  long long delta;
  int status;
  int pad;
};

// Steps of the pass in doubles that every thread takes on its own before the warp walks together.
#define QB_DIAGK_LOCAL_STEPS 7

// The rest of one lane's pass in doubles, thirty-two steps at a time across the warp: lane i
// evaluates h at step idx0 + i, an inclusive scan gives the pivot after every step, the first lane
// at or below the band decides. Called by all 32 lanes with the walking lane's state broadcast.
// Returns the outcome (QB_DIAGK_OK / _OUT_OF_BOUNDS / _GAVE_UP / _UNDECIDED) with delta and x.
__device__ __forceinline__ int diagk_warp_walk(uint32_t l, dd t, double Sd, dd p, uint64_t idx0,
                                               uint64_t delta_bound, int64_t* delta_out, dd* x_out) {
  const unsigned full = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const double two_l = l < 110 ? ldexp(1.0, (int)l) : 0.0, inv_two_l = l < 110 ? ldexp(1.0, -(int)l) : 0.0;
  const uint64_t last = 2 * delta_bound;
  for (;;) {
    const uint64_t idx = idx0 + lane;
    const bool valid = idx <= last && idx + 1 <= QB_DIAGK_MAX_STEPS;
    const int64_t delta = diagk_step_delta(idx);
    const dd x = diagk_walk_x(l, t, delta);
    double s = valid ? diagk_quick_h(l, two_l, inv_two_l, Sd, x) : 0.0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const double v = __shfl_up_sync(full, s, off);
      if (lane >= (unsigned)off) s += v;
    }
    const dd pld = dd_add_d(p, -s);
    const double pl = pld.hi + pld.lo;
    const double band = diagk_band(idx + 1);
    const unsigned hits = __ballot_sync(full, valid && pl <= band);
    const unsigned invalid = __ballot_sync(full, !valid);
    if (hits) {
      const int f = __ffs(hits) - 1;
      const double pf = __shfl_sync(full, pl, f), bf = __shfl_sync(full, band, f);
      *delta_out = (int64_t)__shfl_sync(full, (long long)delta, f);
      *x_out = make_dd(__shfl_sync(full, x.hi, f), __shfl_sync(full, x.lo, f));
      return pf < -bf ? QB_DIAGK_OK : QB_DIAGK_UNDECIDED;
    }
    if (invalid) {
      const uint64_t first_invalid = idx0 + (uint64_t)(__ffs(invalid) - 1);
      *delta_out = 0;
      *x_out = make_dd(0.0, 0.0);
      return first_invalid > last ? QB_DIAGK_OUT_OF_BOUNDS : QB_DIAGK_GAVE_UP;
    }
    p = make_dd(__shfl_sync(full, pld.hi, 31), __shfl_sync(full, pld.lo, 31));
    idx0 += 32;
  }
}

// One tile of 128 samples (tile index tb). `scratch`: this CTA's own 128 x (5 k + 9) words.
__device__ __forceinline__ void diagk_tile(const DiagKConst& c, uint32_t tb, const uint32_t* __restrict__ jT,
                                           const int32_t* __restrict__ eta, const RawX87* __restrict__ pivot,
                                           unsigned long long delta_bound, uint32_t B,
                                           uint32_t* __restrict__ scratch, uint32_t* __restrict__ kT,
                                           DiagKOut* __restrict__ out) {
  const uint32_t g = tb * QB_DIAGK_CTA + threadIdx.x;
  const size_t tile = (size_t)tb * QB_DIAGK_CTA;
  uint32_t* k_out = kT ? kT + tile * c.wl + threadIdx.x : nullptr;
  // every lane stays until the warp has walked together: `live` marks the ones with a sample
  bool live = g < B;
  X87 p = x87_zero();
  if (live) {
    bool ok = true;
    p = x87_load(pivot + g, &ok);
    if (!ok || p.neg || (p.mant != 0 && (p.exp > 0 || (p.exp == 0 && p.mant != 0x8000000000000000ull)))) {
      // the reference: critical("The pivot is out of bounds.") (src/sample.cpp:421-425)
      DiagKOut o;
      o.x_hi = o.x_lo = 0.0;
      o.delta = 0;
      o.status = -1;
      o.pad = 0;
      if (k_out)
        for (uint32_t i = 0; i < c.wl; i++) k_out[(size_t)i * QB_DIAGK_CTA] = 0;
      out[g] = o;
      live = false;
    }
  }
  DiagKFraction f;
  f.t = make_dd(0.0, 0.0);
  f.cflag = f.whole = false;
  dd S = make_dd(0.0, 0.0);
  int status = QB_DIAGK_OUT_OF_BOUNDS;
  int64_t delta = 0;
  dd x = make_dd(0.0, 0.0);
  DiagKQuick q;
  q.p = make_dd(0.0, 0.0);
  q.idx = 0;
  if (live) {
    diagk_fraction<QB_DIAGK_CTA>(c, jT + tile * c.wj + threadIdx.x, eta[g], scratch + threadIdx.x, k_out, &f);
    const dd st = sinpi_acc(f.t);
    S = dd_mul(st, st);
    if (c.force_exact) {
      status = QB_DIAGK_UNDECIDED;
    } else {
      q.p = x87_to_dd(p);  // exact
      status = diagk_quick_steps(c.l, f.t, S.hi, delta_bound, QB_DIAGK_LOCAL_STEPS, &q, &delta, &x);
    }
  }
  // the lanes still walking, one after the other, with the whole warp
  unsigned walking = __ballot_sync(0xffffffffu, live && status == QB_DIAGK_CONTINUE);
  while (walking) {
    const int L = __ffs(walking) - 1;
    const dd tL = make_dd(__shfl_sync(0xffffffffu, f.t.hi, L), __shfl_sync(0xffffffffu, f.t.lo, L));
    const double SL = __shfl_sync(0xffffffffu, S.hi, L);
    const dd pL = make_dd(__shfl_sync(0xffffffffu, q.p.hi, L), __shfl_sync(0xffffffffu, q.p.lo, L));
    const uint64_t iL = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)q.idx, L);
    int64_t dW;
    dd xW;
    const int sW = diagk_warp_walk(c.l, tL, SL, pL, iL, delta_bound, &dW, &xW);
    if ((int)(threadIdx.x & 31u) == L) {
      status = sW;
      delta = dW;
      x = xW;
    }
    walking &= walking - 1;
  }
  if (!live) return;
  if (status == QB_DIAGK_UNDECIDED) status = diagk_walk_exact(c.l, f.t, S, p, delta_bound, &delta, &x);
  DiagKOut o;
  o.pad = 0;
  dd xo;
  int64_t dout;
  o.status = diagk_finish<QB_DIAGK_CTA>(c, f, status, delta, x, k_out, &xo, &dout);
  o.x_hi = xo.hi;
  o.x_lo = xo.lo;
  o.delta = (long long)dout;
  out[g] = o;
}

// One CTA per tile, with its own scratch area of 128 x (5 k + 9) words (the loop below takes any
// grid). Measured and dropped (B200, m = 2048, 303,104 samples, profiles/r02_diagk_persistent_ab.txt):
// a persistent wave of 3 / 4 / 6 CTAs per SM that rewrites one scratch area per CTA -- 75 / 100 /
// 150 MB instead of 400 MB streamed once -- is SLOWER (2.12 / 1.96 / 1.90 ms against 1.76 ms) and
// does not lower the DRAM traffic either (1.23 GB per launch in ncu with 6 per SM: the in-flight
// scratch of a full wave does not fit the L2 next to the inputs, and fewer CTAs cost more than
// the traffic saves -- DRAM is at 8 % of its bandwidth here).
#ifndef QB_DIAGK_MIN_CTAS
#define QB_DIAGK_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(QB_DIAGK_CTA, QB_DIAGK_MIN_CTAS) k_diagk(DiagKConst c, const uint32_t* __restrict__ jT,
                                                const int32_t* __restrict__ eta,
                                                const RawX87* __restrict__ pivot,
                                                unsigned long long delta_bound, uint32_t B,
                                                uint32_t* __restrict__ scratch, uint32_t* __restrict__ kT,
                                                DiagKOut* __restrict__ out) {
  extern __shared__ uint32_t sh[];
  const uint32_t k = c.k;
  // r, d, mu, rho, dq, psi with their zero limbs: contiguous in global memory (qb200_diagk_create), c.r the first
  const uint32_t words = diagk_const_words(c);
  const uint32_t* src = c.r - QB_DIAGK_PAD;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sh[i] = src[i];
  __syncthreads();
  {
    uint32_t at = QB_DIAGK_PAD;
    c.r = sh + at;
    at += k + 2 * QB_DIAGK_PAD;
    c.d = sh + at;
    at += k + 2 * QB_DIAGK_PAD;
    c.mu = sh + at;
    at += k + 2 + 2 * QB_DIAGK_PAD;
    c.rho = sh + at;
    at += k + 2 * QB_DIAGK_PAD;
    c.dq = sh + at;
    at += c.wl + 2 * QB_DIAGK_PAD;
    c.psi = sh + at;
  }
  const uint32_t n_tiles = (B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA;
  uint32_t* mine = scratch + (size_t)blockIdx.x * QB_DIAGK_CTA * diagk_scratch_limbs(k);
  for (uint32_t tb = blockIdx.x; tb < n_tiles; tb += gridDim.x)
    diagk_tile(c, tb, jT, eta, pivot, delta_bound, B, mine, kT, out);
}

// status[t] = 0, 1 (a sample ran out of bounds, or |eta| > eta_bound: the reference breaks and
// returns FALSE), or the first other status of the estimate's samples (2: a sample with a negative
// unreduced phi and l > 1000, whose alpha_phi^2 ~ 2^(2 l) leaves the doubles).
__global__ void k_diagk_tau(const DiagKOut* __restrict__ out, const int32_t* __restrict__ eta, uint32_t l,
                            uint32_t n, uint32_t count, uint32_t eta_bound, double* __restrict__ sums,
                            int* __restrict__ status) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  dd s = make_dd(0.0, 0.0);
  int st = 0;
  for (uint32_t i = 0; i < n; i++) {
    const DiagKOut o = out[(size_t)t * n + i];
    if (o.status != 0 && !(o.status == QB_DIAGK_OK_NEGATIVE_PHI && l <= 1000)) {
      st = o.status;
      break;
    }
    const int32_t e = eta[(size_t)t * n + i];
    if ((uint32_t)(e < 0 ? -e : e) > eta_bound) {
      st = QB_DIAGK_OUT_OF_BOUNDS;
      break;
    }
    dd x = make_dd(o.x_hi, o.x_lo);
    if (o.status == QB_DIAGK_OK_NEGATIVE_PHI) x = dd_add_d(x, -ldexp(1.0, (int)l));
    s = dd_add(s, dd_mul(x, x));
  }
  sums[2 * (size_t)t] = s.hi;
  sums[2 * (size_t)t + 1] = s.lo;
  status[t] = st;
}

__global__ void k_diagk_h(uint32_t l, uint32_t n, const double* __restrict__ x_hi,
                          const double* __restrict__ x_lo, RawX87* __restrict__ out) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const dd x = quick_two_sum(x_hi[g], x_lo[g]);
  dd t = dd_add_d(x, -rint(x.hi));
  if (t.hi > 0.5) t = dd_add_d(t, -1.0);
  if (t.hi < -0.5) t = dd_add_d(t, 1.0);
  const dd st = sinpi_acc(t);
  const X87 h = x87_from_dd(diagk_h(l, dd_mul(st, st), x));
  RawX87 r;
  r.mant = h.mant;
  r.se = h.mant ? (uint64_t)(h.exp + 16383) : 0;
  out[g] = r;
}

}  // namespace qb200
