// kernels_sampler.cuh -- kernels of the distribution sampler (sampler.cuh has the per-sample
// code and the reference citations).
//
//   k_seg_build   one CTA per segment (a slice's cells, or the list of slice totals): block
//                 summaries of QB_SEG_BLOCK (8) elements each in parallel, then the walk states before every
//                 block (prefix sum and running maximum, double-double). Runs once per
//                 distribution. HBM-bound: 16 B read per cell.
//   k_sample      one thread per sample: two searches (slices, then the cells of the slice: 10 + 11
//                 dependent 32-byte reads of the coarse index + one 128-byte block of cells each)
//                 and the two axis draws. Bound by L1 lookups of divergent addresses (82 % of the
//                 L1/TEX peak in ncu).
//   k_tau_reduce  one thread per estimate: the n squares summed in sample order.
#pragma once

#include <cuda_runtime.h>

#include "sampler.cuh"

namespace qb200 {

struct SegDesc {
  const RawX87* vals;
  double* vals_d;    // the elements as doubles (seg_block_doubles), or null
  SegCoarse* coarse;
  double* abs_out;
  uint32_t* guide;   // seg_guide_size(blocks) + 1 entries
  uint32_t n;
  uint32_t pad;
};

__global__ void __launch_bounds__(256) k_seg_build(const SegDesc* __restrict__ segs, int* __restrict__ bad) {
  const SegDesc s = segs[blockIdx.x];
  const uint32_t nb = (s.n + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
  __shared__ double sh_abs;
  if (threadIdx.x == 0) sh_abs = 0.0;
  __syncthreads();
  double ab = 0.0;
  bool ok = true;
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
    dd sum, maxp;
    double a;
    seg_block_summary(s.vals, s.vals_d, s.n, b, &sum, &maxp, &a, &ok);
    s.coarse[b + 1].c = sum;
    s.coarse[b + 1].m = maxp;
    ab += a;
  }
  atomicAdd(&sh_abs, ab);
  if (!ok) atomicOr(bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    seg_scan(s.coarse, nb);
    *s.abs_out = sh_abs;
  }
  __syncthreads();
  const uint32_t G = seg_guide_size(nb);
  for (uint32_t u = threadIdx.x; u <= G; u += blockDim.x) s.guide[u] = seg_guide_entry(s.coarse, nb, G, u);
}

#define QB_TAU_SKIP 0xffffffffffffffffull

#ifndef QB_SAMPLE_MIN_CTAS
#define QB_SAMPLE_MIN_CTAS 6
#endif

// Sample i of estimate t reads its words at off[t] + i * wps (off == QB_TAU_SKIP: the host
// already knows that the estimate fails, nothing to do). off == nullptr: regular layout.
//
// Bound by look-ups in L1: every load of a warp touches 32 different lines (ncu, round 2: 0.84 tag
// look-ups per cycle and SM). What helped: 16-byte loads (ld16, sampler.cuh; 44 -> 33 million
// look-ups per 2^20 samples, 4.4 -> 6.5e9 samples/s) and the elements of a block kept as doubles
// (seg_block_doubles: 4 loads per block instead of 8, 7.6e9). What did not: fetching the 32 blocks of a
// warp cooperatively (8 lanes x 16 bytes per block, values handed to their owners through shared
// memory): a quarter fewer look-ups again, but every lane then waits for the slowest search of its
// warp twice per sample -- 5.6e9 samples/s, removed (profiles/r02_sampler_lookups_ab.txt).
__global__ void __launch_bounds__(128, QB_SAMPLE_MIN_CTAS)
k_sample(SamplerView view, const uint64_t* __restrict__ words, const uint64_t* __restrict__ off, uint32_t n,
         uint64_t total, int force_exact, SampleOut* __restrict__ out) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const uint32_t wps = (uint32_t)view.dims + 2u;
  const uint64_t t = g / n, i = g - t * n;
  uint64_t base;
  if (off) {
    base = off[t];
    if (base == QB_TAU_SKIP) {
      out[g].status = kSampleOutOfBounds;
      out[g].exact = 0;
      return;
    }
    base += i * wps;
  } else {
    base = g * wps;
  }
  uint64_t w[4];  // (unrolled with a predicate: a loop to wps indexes w dynamically and puts it in local memory)
#pragma unroll
  for (uint32_t q = 0; q < 4; q++) w[q] = q < wps ? words[base + q] : 0ull;
  SampleOut o;
  sample_one(view, w, force_exact, &o);
  out[g] = o;
}

// sums[4 t ..] = sum (alpha_d / 2^m)^2 (hi, lo), sum (alpha_r / 2^m)^2 (hi, lo); status[t] = 0
// or the status of the first failing sample; exact_total += replayed walks.
__global__ void k_tau_reduce(const SampleOut* __restrict__ out, uint32_t n, uint32_t count,
                             double* __restrict__ sums, int* __restrict__ status,
                             unsigned long long* __restrict__ exact_total) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const SampleOut* o = out + (size_t)t * n;
  dd a, b;
  int st;
  tau_sums(o, n, &a, &b, &st);
  sums[4 * (size_t)t] = a.hi;
  sums[4 * (size_t)t + 1] = a.lo;
  sums[4 * (size_t)t + 2] = b.hi;
  sums[4 * (size_t)t + 3] = b.lo;
  status[t] = st;
  unsigned long long ex = 0;
  for (uint32_t i = 0; i < n; i++) ex += (unsigned long long)o[i].exact;
  if (ex) atomicAdd(exact_total, ex);
}

}  // namespace qb200
