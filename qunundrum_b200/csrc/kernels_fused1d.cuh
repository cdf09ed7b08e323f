// kernels_fused1d.cuh -- all one-dimensional slices of a distribution in ONE launch.
//
// linear_distribution_slice_compute[_richardson]   src/linear_distribution_slice_compute.cpp:30-245
// diagonal_distribution_slice_compute[_richardson] src/diagonal_distribution_slice_compute.cpp:30-210
//
// A one-dimensional distribution is 41 (linear) to a few thousand (diagonal, (2 B_eta + 1) x 60)
// slices of D = 2048 cells: 5 10^5 ... 4 10^7 integrand evaluations, i.e. microseconds to a few
// hundred microseconds of FP64 work. The plain path (k_vals1d -> k_cells1d -> k_final1d) spends
// three launches and a round trip of the 6 D + 2 point values through memory per batch; here one
// launch does everything:
//
//   * grid = (blocks per slice, slices): a block of 128 threads owns 128 consecutive cells;
//   * thread I evaluates the five abscissae only its cell sees -- fine points 4 I .. 4 I + 3 and
//     the coarse mid-point 2 I + 1 -- and takes its right end (fine 4 I + 4 = coarse 2 I + 2) from
//     its neighbour through shared memory; the coarse end points ARE fine points (2^(i/D) =
//     2^(2i/2D), identical table entries), so a cell costs 5 evaluations instead of the
//     reference's 6.001;
//   * Simpson, Richardson (2 * fine - coarse) and the cell scaling with the operations of
//     pass1d_cell (slice_cells.cuh), so the cells equal the plain path's bit for bit;
//   * the block's double-double partial goes to part[slice][block]; the last block of a slice to
//     finish (ticket per slice) adds the partials in block order -- a fixed order, so the summary
//     does not depend on scheduling or on the batch a slice is computed in.
#pragma once

#include <cuda_runtime.h>

#include "kernels_plain.cuh"

namespace qb200 {

#define QB_F1D_BLOCK 128

__global__ void __launch_bounds__(QB_F1D_BLOCK)
k_fused1d(DevConsts c, int kind, int D, int richardson, const DevSlice* __restrict__ slices,
          const TabDesc* __restrict__ desc, const dd* __restrict__ gx,
          const double* __restrict__ gw, double* __restrict__ out, double* __restrict__ part,
          unsigned int* __restrict__ tickets, double* __restrict__ summary) {
  __shared__ double s_left[QB_F1D_BLOCK + 1];
  __shared__ double sh[QB_F1D_BLOCK], sl[QB_F1D_BLOCK];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  const unsigned slice = blockIdx.y;
  const int I = blockIdx.x * QB_F1D_BLOCK + tid;
  const DevSlice s = slices[slice];
  const TabDesc t = desc[s.tab_a];
  const dd* gxc = gx;                       // coarse pass, 2 D + 1 interleaved points
  const dd* gxf = gx + pass_offset(D, 1);   // fine pass, 4 D + 1
  const int step = richardson ? 4 : 2;      // points per cell in the pass that holds the ends
  const dd* ends = richardson ? gxf : gxc;
  double v0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0, cm = 0.0;
  if (I < D) {
    v0 = value_1d(c, kind, ends[step * I], t, s.eta_shift);
    cm = value_1d(c, kind, gxc[2 * I + 1], t, s.eta_shift);
    if (richardson) {
      f1 = value_1d(c, kind, gxf[4 * I + 1], t, s.eta_shift);
      f2 = value_1d(c, kind, gxf[4 * I + 2], t, s.eta_shift);
      f3 = value_1d(c, kind, gxf[4 * I + 3], t, s.eta_shift);
    }
  }
  s_left[tid] = v0;
  // the block's right end: the thread after the block's last active cell evaluates it, so the
  // extra evaluation does not lengthen a warp that already did five
  const int last = min(D, (int)(blockIdx.x + 1) * QB_F1D_BLOCK);  // first cell past the block
  if (I == last && tid < QB_F1D_BLOCK) s_left[tid] = value_1d(c, kind, ends[step * I], t, s.eta_shift);
  if (tid == QB_F1D_BLOCK - 1 && I == last - 1)
    s_left[QB_F1D_BLOCK] = value_1d(c, kind, ends[step * last], t, s.eta_shift);
  __syncthreads();
  double v = 0.0;
  if (I < D) {
    const double v4 = s_left[tid + 1];
    // pass1d_cell's operations, on registers
    v = (fma(4.0, cm, v0) + v4) / 6.0 * (gw[I] * s.scale_a);
    if (richardson) {
      const double* wf = gw + width_offset(D, 1);
      const double fa = (fma(4.0, f1, v0) + f2) / 6.0 * (wf[2 * I] * s.scale_a);
      const double fb = (fma(4.0, f3, f2) + v4) / 6.0 * (wf[2 * I + 1] * s.scale_a);
      v = 2.0 * (fa + fb) - v;
    }
    out[(size_t)slice * D + I] = v;
  }
  const dd tsum = block_sum_dd<QB_F1D_BLOCK>(make_dd(v, 0.0), sh, sl);
  const unsigned nb = gridDim.x;
  if (tid == 0) {
    double* p = part + ((size_t)slice * nb + blockIdx.x) * 2;
    p[0] = tsum.hi;
    p[1] = tsum.lo;
    __threadfence();
    s_last = (atomicAdd(tickets + slice, 1u) == nb - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    dd tp = make_dd(0.0, 0.0);
    const volatile double* p = part + (size_t)slice * nb * 2;
    for (unsigned b = 0; b < nb; b++) tp = dd_add(tp, make_dd(p[2 * b], p[2 * b + 1]));
    double* o = summary + (size_t)slice * 8;
    o[0] = tp.hi;
    o[1] = tp.lo;
    o[2] = o[3] = 0.0;
    o[4] = 1.0;
    o[5] = o[6] = o[7] = 0.0;
    tickets[slice] = 0;  // ready for the next run of the plan
  }
}

}  // namespace qb200
