// kernels_client.cuh -- what the generator's client and server loops do to the slices AFTER the
// integration, on the device and bit for bit in the reference's x87
// long double arithmetic (x87soft.cuh):
//
//   distribution_slice_copy_scale            src/distribution_slice.cpp:230-264
//       (the client scales a 512 / 1024 slice to MAX_SLICE_DIMENSION = 256 before sending it,
//        src/main_generate_distribution.cpp:1308-1343)
//   linear_distribution_init_collapse_d/_r   src/linear_distribution.cpp:152-324
//       (the server collapses the finished distribution to its two marginals,
//        src/main_generate_distribution.cpp:709-760)
//
// Both are sums of long doubles in a fixed order; every partial sum is rounded to 64 bits, so the
// order is part of the result. The kernels keep the reference's order per output element and
// parallelise over the output elements.
#pragma once

#include <cuda_runtime.h>

#include "client_math.cuh"

namespace qb200 {

__device__ __forceinline__ ulonglong2 x87_store(X87 a, bool* ok) {
  uint64_t m = 0, se = 0;
  x87_encode(a, &m, &se, ok);
  ulonglong2 w;
  w.x = m;
  w.y = se;
  return w;
}

// ---- distribution_slice_copy_scale ---------------------------------------------------------------
// cells: n slices of D x D doubles (index i_d + D j_r); out: n slices of store x store x87 values.
// One thread per destination cell: 0 + the scale x scale block, alpha_d offset outermost
// (src/distribution_slice.cpp:249-261).
static __global__ void __launch_bounds__(256)
k_scale_x87(int D, int store, const double* __restrict__ cells, ulonglong2* __restrict__ out,
            int* __restrict__ status) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= store * store) return;
  const X87 acc = scale_cell_x87(cells + (size_t)blockIdx.y * D * D, D, store, idx);
  bool ok = true;
  const ulonglong2 w = x87_store(acc, &ok);
  if (!ok) atomicOr(status, 1);
  out[(size_t)blockIdx.y * store * store + idx] = w;
}

// ---- collapse to a marginal ----------------------------------------------------------------------
struct CollapseSrc {
  unsigned long long offset;  // first cell of the slice in the resident buffer (16-byte units)
  unsigned int dimension;
  unsigned int divisor;       // max_dimension / dimension
};

#define QB_COLLAPSE_WARPS 2

// Element e of destination slice blockIdx.y: for every source slice of that destination, in the
// distribution's order, += cell / divisor over the other axis in ascending order
// (src/linear_distribution.cpp:216-231 for alpha_d, :303-318 for alpha_r).
//   axis 0 (collapse to alpha_d): x = e / divisor, cells x + y D for y = 0 .. D - 1: a warp reads
//     consecutive x directly;
//   axis 1 (collapse to alpha_r): y = e / divisor, cells x + y D for x = 0 .. D - 1: a warp's lanes
//     own 32 different rows, so it stages 32 x 32 tiles through shared memory (coalesced row
//     loads, conflict-free column reads).
static __global__ void __launch_bounds__(32 * QB_COLLAPSE_WARPS)
k_collapse(int axis, unsigned max_dim, const unsigned* __restrict__ src_begin,
           const unsigned* __restrict__ src_index, const CollapseSrc* __restrict__ srcs,
           const ulonglong2* __restrict__ cells, ulonglong2* __restrict__ out,
           int* __restrict__ status) {
  __shared__ ulonglong2 tile[QB_COLLAPSE_WARPS][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned e0 = (blockIdx.x * QB_COLLAPSE_WARPS + warp) * 32;
  if (e0 >= max_dim) return;  // whole warp
  const unsigned e = e0 + lane;
  const bool live = e < max_dim;
  const unsigned dst = blockIdx.y;
  X87 acc = x87_zero();
  bool ok = true;
  for (unsigned s = src_begin[dst]; s < src_begin[dst + 1]; s++) {
    const CollapseSrc src = srcs[src_index[s]];
    const ulonglong2* c = cells + src.offset;
    const unsigned D = src.dimension, q = src.divisor;
    if (axis == 0) {
      if (live) {
        const unsigned x = e / q;
        for (unsigned y = 0; y < D; y++) {
          const ulonglong2 w = c[x + (size_t)y * D];
          acc = collapse_step(acc, w.x, w.y, q, &ok);
        }
      }
    } else {
      for (unsigned x0 = 0; x0 < D; x0 += 32) {
        __syncwarp();
        for (int rr = 0; rr < 32; rr++) {
          const unsigned er = e0 + rr < max_dim ? e0 + rr : max_dim - 1;
          const unsigned y = er / q;
          if (x0 + lane < D) tile[warp][rr][lane] = c[(x0 + lane) + (size_t)y * D];
        }
        __syncwarp();
        if (live) {
          const unsigned nx = D - x0 < 32 ? D - x0 : 32;
          for (unsigned k = 0; k < nx; k++) {
            const ulonglong2 w = tile[warp][lane][k];
            acc = collapse_step(acc, w.x, w.y, q, &ok);
          }
        }
      }
    }
  }
  if (live) {
    out[(size_t)dst * max_dim + e] = x87_store(acc, &ok);
  }
  if (!ok) atomicOr(status, 1);
}

}  // namespace qb200
