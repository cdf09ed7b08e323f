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
  SegCoarse* coarse;
  double* abs_out;
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
    seg_block_summary(s.vals, s.n, b, &sum, &maxp, &a, &ok);
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
}

#define QB_TAU_SKIP 0xffffffffffffffffull

// Sample i of estimate t reads its words at off[t] + i * wps (off == QB_TAU_SKIP: the host
// already knows that the estimate fails, nothing to do). off == nullptr: regular layout.
__global__ void __launch_bounds__(128) k_sample(SamplerView view, const uint64_t* __restrict__ words,
                                                 const uint64_t* __restrict__ off, uint32_t n,
                                                 uint64_t total, int force_exact,
                                                 SampleOut* __restrict__ out) {
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
  uint64_t w[4];
  for (uint32_t q = 0; q < wps; q++) w[q] = words[base + q];
  SampleOut o;
  sample_one(view, w, force_exact, &o);
  out[g] = o;
}

// ---- the same in two phases, with the samples bucketed by slice in between -----------------------
// A sample's second search runs over the cells of ITS slice; with one thread per sample in draw
// order every lane of a warp walks another slice's index (15.8 of 32 lanes active per instruction,
// L1/TEX-bound on divergent lines in ncu, round 1). Here
//   k_sample_slices   does the first search only and counts the samples per slice (one atomic per
//                     distinct slice of a warp: __match_any_sync), remembering every sample's rank
//                     inside its slice's bucket;
//   k_sample_offsets  turns the counts into bucket starts (one block, exclusive scan);
//   k_sample_cells    thread q takes the q-th sample in bucket order -- the lanes of a warp then
//                     search the same slice (the probable slices fill whole warps), read the same
//                     upper levels of its index and finish after the same number of steps -- and
//                     writes the result at the sample's own position, so outputs and tau sums do
//                     not depend on the bucket order.
// Bucket n_slices holds the samples that ran out of bounds or were skipped.

__device__ __forceinline__ uint64_t sample_word_base(const uint64_t* __restrict__ off, uint32_t n,
                                                     uint32_t wps, uint64_t g) {
  if (!off) return g * wps;
  const uint64_t t = g / n, i = g - t * n;
  const uint64_t b = off[t];
  return b == QB_TAU_SKIP ? QB_TAU_SKIP : b + i * wps;
}

__global__ void __launch_bounds__(128) k_sample_slices(SamplerView view, const uint64_t* __restrict__ words,
                                                        const uint64_t* __restrict__ off, uint32_t n,
                                                        uint64_t total, int force_exact,
                                                        uint32_t* __restrict__ slice_of,
                                                        uint32_t* __restrict__ rank_of,
                                                        unsigned int* __restrict__ counts,
                                                        unsigned long long* __restrict__ exact_total) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = g < total;
  uint32_t sl = view.n_slices;
  int exact = 0;
  if (live) {
    const uint64_t base = sample_word_base(off, n, (uint32_t)view.dims + 2u, g);
    if (base != QB_TAU_SKIP) sl = sample_slice(view, words[base], force_exact, &exact);
    if (sl > view.n_slices) sl = view.n_slices;
  }
  const unsigned active = __ballot_sync(0xffffffffu, live);
  if (!live) return;
  const unsigned peers = __match_any_sync(active, sl);
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(peers) - 1;
  unsigned int start = 0;
  if ((int)lane == leader) start = atomicAdd(counts + sl, (unsigned int)__popc(peers));
  start = __shfl_sync(peers, start, leader);
  slice_of[g] = sl;
  rank_of[g] = start + (unsigned int)__popc(peers & ((1u << lane) - 1u));
  if (exact) atomicAdd(exact_total, (unsigned long long)exact);
}

// starts[b] = samples in the buckets before b (b = 0 .. n_buckets); one block.
__global__ void __launch_bounds__(1024) k_sample_offsets(const unsigned int* __restrict__ counts,
                                                          uint32_t n_buckets, unsigned int* __restrict__ starts) {
  __shared__ unsigned int part[1024];
  const uint32_t per = (n_buckets + 1023) / 1024;
  const uint32_t lo = threadIdx.x * per, hi = min(n_buckets, lo + per);
  unsigned int sum = 0;
  for (uint32_t b = lo; b < hi; b++) sum += counts[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned int v = threadIdx.x >= (unsigned)d ? part[threadIdx.x - d] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = part[threadIdx.x] - sum;
  for (uint32_t b = lo; b < hi; b++) {
    starts[b] = run;
    run += counts[b];
  }
  if (threadIdx.x == 1023) starts[n_buckets] = part[1023];
}

__global__ void __launch_bounds__(256) k_sample_order(const uint32_t* __restrict__ slice_of,
                                                       const uint32_t* __restrict__ rank_of,
                                                       const unsigned int* __restrict__ starts, uint64_t total,
                                                       uint32_t* __restrict__ order) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < total) order[starts[slice_of[g]] + rank_of[g]] = (uint32_t)g;
}

__global__ void __launch_bounds__(128) k_sample_cells(SamplerView view, const uint64_t* __restrict__ words,
                                                       const uint64_t* __restrict__ off, uint32_t n,
                                                       uint64_t total, int force_exact,
                                                       const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ slice_of,
                                                       SampleOut* __restrict__ out) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const uint64_t g = order[q];
  const uint32_t sl = slice_of[g];
  SampleOut o;
  sample_out_clear(&o);
  if (sl < view.n_slices) {
    const uint32_t wps = (uint32_t)view.dims + 2u;
    const uint64_t base = sample_word_base(off, n, wps, g);
    uint64_t w[4];
    w[0] = 0;
    for (uint32_t k = 1; k < wps; k++) w[k] = words[base + k];
    sample_in_slice(view, sl, w, force_exact, &o);
  }
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
