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

// One instantiation per integrand and ONE call site of it (the six abscissae of a thread go
// through a loop that is not unrolled): the three integrands inlined at six call sites were 8000
// instructions, more than the instruction cache holds.
template <int KIND>
__global__ void __launch_bounds__(QB_F1D_BLOCK)
k_fused1d(DevConsts c, int D, int richardson, const DevSlice* __restrict__ slices,
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
  const int last = min(D, (int)(blockIdx.x + 1) * QB_F1D_BLOCK);  // first cell past the block
  // abscissa p of this thread: 0 left end, 1 coarse mid-point, 2..4 fine 4 I + 1 .. 3, 5 the
  // block's right end (evaluated by the thread after the block's last cell, or, in a full block,
  // by its last thread)
  double vals[6];
#pragma unroll 1
  for (int p = 0; p < 6; p++) {
    const dd* tab = ends;
    int at = step * I;
    bool need = I < D;
    if (p == 1) {
      tab = gxc;
      at = 2 * I + 1;
    } else if (p >= 2 && p <= 4) {
      tab = gxf;
      at = 4 * I + p - 1;
      need = need && richardson;
    } else if (p == 5) {
      at = step * last;
      need = (I == last) || (tid == QB_F1D_BLOCK - 1 && I == last - 1);
    }
    double v = 0.0;
    if (need) {
      const dd x = grid_x(tab[at], t.k_abs, t.sign, c.m);
      if (KIND == KIND_LINEAR_D) v = linear_d_value(x, c.d_m, c.omd_m, c.l);
      else if (KIND == KIND_LINEAR_R) v = linear_r_value(x, c);
      else v = diagonal_value(x, s.eta_shift, c.rho);
    }
    vals[p] = v;
  }
  const double v0 = vals[0], cm = vals[1], f1 = vals[2], f2 = vals[3], f3 = vals[4];
  s_left[tid] = v0;
  if (I == last) s_left[tid] = vals[5];
  if (tid == QB_F1D_BLOCK - 1 && I == last - 1) s_left[QB_F1D_BLOCK] = vals[5];
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
