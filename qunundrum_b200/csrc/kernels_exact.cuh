// kernels_exact.cuh -- kernels of the exact samplers (exact.cuh has the per-sample code and the
// reference citations).
//
//   k_exact_alpha   one thread per sample: the region's bounds from the table of 2^(i/D), the
//                   modulus, the sample's bytes modulo the modulus (at most four quotient digits),
//                   alpha = min + remainder with the low kappa bits cleared. Every thread walks its
//                   OWN two table rows and its own bytes (32 sectors per load instruction of a warp):
//                   bound by those look-ups. Measured and dropped (B200, 303,104 samples at m = 2048,
//                   profiles/r02_exact_alpha_staged_ab.txt): the warp fetching each sample's words with
//                   one coalesced request and handing them over transposed through shared memory --
//                   0.55 ms against 0.39 ms (70 registers and 34 KB of shared memory per CTA leave 32 %
//                   occupancy, and the 32 load -> store hand-overs per chunk wait on each other:
//                   long-scoreboard 20 cycles per instruction at 21 % issue).
//   k_exact_jk      one thread per sample: j from alpha_r (one truncated product with the inverse
//                   of r / 2^kappa_r modulo 2^n, staged in shared memory: every thread reads the
//                   same limb), k from (alpha_d, j) (a truncated product with d), or j from
//                   (alpha_d, k). Integer-pipe bound like k_diagk: about (wa wn + wd wn) 32-bit
//                   multiply-adds per sample.
// Samples are processed in tiles of QB_DIAGK_CTA with the limbs interleaved exactly like the
// diagonal k sampler's (kernels_diagk.cuh: word i of the t-th sample of tile b of an array with w
// words per sample at [(b w + i) 128 + t]), so that the j tiles k_exact_jk writes are what k_diagk
// reads: a diagonal sample goes from its random bytes to k without leaving the device.
#pragma once

#include <cuda_runtime.h>

#include "exact.cuh"

// The tile size of kernels_diagk.cuh (not included here: its kernels belong to qb200_diagk.cu).
#ifndef QB_DIAGK_CTA
#define QB_DIAGK_CTA 128
#endif

namespace qb200 {

// Rows (w words per sample) to tiles and back: k_diagk_gather / k_diagk_scatter's layout.
__global__ void k_exact_gather(const uint32_t* __restrict__ rows, uint32_t w, uint32_t B,
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

__global__ void k_exact_scatter(const uint32_t* __restrict__ tiles, uint32_t w, uint32_t B,
                                uint32_t* __restrict__ rows) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)w * B) return;
  const uint32_t g = (uint32_t)(idx / w), i = (uint32_t)(idx - (uint64_t)g * w);
  rows[idx] = tiles[((size_t)(g / QB_DIAGK_CTA) * w + i) * QB_DIAGK_CTA + g % QB_DIAGK_CTA];
}

__global__ void __launch_bounds__(QB_DIAGK_CTA) k_exact_alpha(ExactConst c, const ExactRegion* __restrict__ regions,
                                                             uint32_t kappa, const uint8_t* __restrict__ stream,
                                                             unsigned long long stream_len, uint32_t B,
                                                             uint32_t* __restrict__ scratch,
                                                             uint32_t* __restrict__ alphaT,
                                                             int32_t* __restrict__ negative,
                                                             int32_t* __restrict__ status) {
  const uint32_t n_tiles = (B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA;
  uint32_t* mine = scratch + (size_t)blockIdx.x * QB_DIAGK_CTA * exact_alpha_scratch_limbs(c) + threadIdx.x;
  for (uint32_t tb = blockIdx.x; tb < n_tiles; tb += gridDim.x) {
    const uint32_t g = tb * QB_DIAGK_CTA + threadIdx.x;
    uint32_t* out = alphaT + (size_t)tb * QB_DIAGK_CTA * c.wa + threadIdx.x;
    if (g >= B) {
      for (uint32_t i = 0; i < c.wa; i++) out[(size_t)i * QB_DIAGK_CTA] = 0u;
      continue;
    }
    int neg = 0;
    const ExactRegion r = regions[g];
    const int st = exact_alpha<QB_DIAGK_CTA, QB_DIAGK_CTA>(c, r, kappa, stream, stream_len, mine, out, &neg);
    negative[g] = neg;
    status[g] = st;
  }
}

#define QB_EXACT_J_FROM_ALPHA_R 0
#define QB_EXACT_J_K_FROM_ALPHA_D_R 1
#define QB_EXACT_J_FROM_ALPHA_D_K 2

// tT: t tiles (tl words per sample; NULL when the kappa in question is 0). kT: k tiles, read in
// mode 2, written in mode 1.
template <int NC>
__global__ void __launch_bounds__(QB_DIAGK_CTA) k_exact_jk(ExactConst c, int mode, const uint32_t* __restrict__ adT,
                                                          const int32_t* __restrict__ neg_d,
                                                          const uint32_t* __restrict__ arT,
                                                          const int32_t* __restrict__ neg_r,
                                                          const uint32_t* __restrict__ tT, uint32_t tl,
                                                          uint32_t* kT, uint32_t B,
                                                          uint32_t* __restrict__ scratch,
                                                          uint32_t* jT) {
  extern __shared__ uint32_t sh[];
  // inv_r, inv_d, d with their zero limbs: contiguous in global memory (qb200_exact_create), inv_r first
  const uint32_t words = exact_const_words(c);
  const uint32_t* src = c.inv_r - QB_EXACT_PAD;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sh[i] = src[i];
  __syncthreads();
  c.inv_r = sh + QB_EXACT_PAD;
  c.inv_d = sh + (c.wn + 2 * QB_EXACT_PAD) + QB_EXACT_PAD;
  c.d = sh + 2 * (c.wn + 2 * QB_EXACT_PAD) + QB_EXACT_PAD;
  const uint32_t n_tiles = (B + QB_DIAGK_CTA - 1) / QB_DIAGK_CTA;
  uint32_t* mine = scratch + (size_t)blockIdx.x * QB_DIAGK_CTA * exact_jk_scratch_limbs(c) + threadIdx.x;
  constexpr int S = QB_DIAGK_CTA;
  for (uint32_t tb = blockIdx.x; tb < n_tiles; tb += gridDim.x) {
    const uint32_t g = tb * QB_DIAGK_CTA + threadIdx.x;
    const size_t tile = (size_t)tb * QB_DIAGK_CTA;
    uint32_t* j_out = jT + tile * c.wn + threadIdx.x;
    if (g >= B) {
      for (uint32_t i = 0; i < c.wn; i++) j_out[(size_t)i * S] = 0u;
      continue;
    }
    const uint32_t* t = tT ? tT + tile * tl + threadIdx.x : nullptr;
    if (mode == QB_EXACT_J_FROM_ALPHA_D_K) {
      exact_j_from_alpha_d_k<NC, S, S, S, S, S>(c, adT + tile * c.wa + threadIdx.x, neg_d[g],
                                            kT + tile * c.wk + threadIdx.x, t, mine, j_out);
      continue;
    }
    exact_j_from_alpha_r<NC, S, S, S, S>(c, arT + tile * c.wa + threadIdx.x, neg_r[g], t, mine, j_out);
    if (mode == QB_EXACT_J_K_FROM_ALPHA_D_R)
      exact_k_from_alpha_d_j<NC, S, S, S, S>(c, adT + tile * c.wa + threadIdx.x, neg_d[g], j_out, mine,
                                         kT + tile * c.wk + threadIdx.x);
  }
}

}  // namespace qb200
