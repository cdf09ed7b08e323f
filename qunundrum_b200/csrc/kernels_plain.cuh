// kernels_plain.cuh -- "plain" CUDA kernels: one thread per abscissa / per cell.
//
// These are the general path: any dimension, any (l, sigma), single pass or
// Richardson. The two-dimensional Richardson hot path for power-of-two
// dimensions runs through the fused kernel in kernels_fused2d.cuh instead; the
// plain kernels remain its on-device cross-check (tests compare the two) and
// serve the shapes the fused kernel does not take.
#pragma once

#include <cuda_runtime.h>

#include "slice_cells.cuh"

namespace qb200 {

struct DevSlice {   // device copy of SliceDesc
  int tab_a, tab_b;
  double scale_a, scale_b;
  double eta_shift;
};

// ---- axis tables ------------------------------------------------------------
// grid.x covers the NP abscissae, grid.y the tables (alpha_d tables first).
__global__ void k_axis2d(DevConsts c, int NP, int n_tab_a, const TabDesc* __restrict__ desc_a,
                         const TabDesc* __restrict__ desc_b, const dd* __restrict__ gx,
                         AxisD* __restrict__ tab_a, AxisR* __restrict__ tab_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP) return;
  const int t = blockIdx.y;
  const dd g = gx[i];
  if (t < n_tab_a) {
    AxisD o;
    axis_d_point(c, g, desc_a[t], &o);
    tab_a[(size_t)t * NP + i] = o;
  } else {
    AxisR o;
    axis_r_point(c, g, desc_b[t - n_tab_a], &o);
    tab_b[(size_t)(t - n_tab_a) * NP + i] = o;
  }
}

// ---- block reductions (fixed tree => deterministic) ------------------------
template <int BLOCK>
__device__ __forceinline__ void block_sum2_and(double& a, double& b, int& ok, double* sa,
                                               double* sb, int* so) {
  const int t = threadIdx.x;
  sa[t] = a;
  sb[t] = b;
  so[t] = ok;
  __syncthreads();
#pragma unroll
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (t < s) {
      sa[t] += sa[t + s];
      sb[t] += sb[t + s];
      so[t] &= so[t + s];
    }
    __syncthreads();
  }
  a = sa[0];
  b = sb[0];
  ok = so[0];
}

template <int BLOCK>
__device__ __forceinline__ dd block_sum_dd(dd v, double* sh, double* sl) {
  const int t = threadIdx.x;
  sh[t] = v.hi;
  sl[t] = v.lo;
  __syncthreads();
#pragma unroll
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (t < s) {
      const dd r = dd_add(make_dd(sh[t], sl[t]), make_dd(sh[t + s], sl[t + s]));
      sh[t] = r.hi;
      sl[t] = r.lo;
    }
    __syncthreads();
  }
  return make_dd(sh[0], sl[0]);
}

#define QB_PLAIN_BLOCK 256

// ---- one Simpson pass over a chunk of 2D slices ----------------------------
// grid.x: blocks over the Dp^2 cells of a slice, grid.y: slice within chunk.
// cells_pass: [slice][Dp^2]; partial: [slice][gridDim.x][3] = (m1, m2, bounded).
__global__ void __launch_bounds__(QB_PLAIN_BLOCK)
k_pass2d(DevConsts c, int D, int fine, int with_error, const DevSlice* __restrict__ slices,
         const AxisD* __restrict__ tab_a, const AxisR* __restrict__ tab_b,
         const double* __restrict__ gw, double* __restrict__ cells_pass,
         double* __restrict__ partial) {
  __shared__ double sa[QB_PLAIN_BLOCK], sb[QB_PLAIN_BLOCK];
  __shared__ int so[QB_PLAIN_BLOCK];
  const int Dp = fine ? 2 * D : D;
  const int NP = table_points(D);
  const DevSlice s = slices[blockIdx.y];
  const int cell = blockIdx.x * QB_PLAIN_BLOCK + threadIdx.x;
  double m1 = 0.0, m2 = 0.0;
  int ok = 1;
  if (cell < Dp * Dp) {
    const int I = cell % Dp, J = cell / Dp;
    const AxisD* td = tab_a + (size_t)s.tab_a * NP + pass_offset(D, fine);
    const AxisR* tr = tab_b + (size_t)s.tab_b * NP + pass_offset(D, fine);
    const double* w = gw + width_offset(D, fine);
    double mass;
    bool b;
    pass2d_cell(c, td, tr, w[I] * s.scale_a, w[J] * s.scale_b, I, J, with_error != 0, &mass,
                &m1, &m2, &b);
    cells_pass[(size_t)blockIdx.y * Dp * Dp + cell] = mass;
    ok = b ? 1 : 0;
  }
  block_sum2_and<QB_PLAIN_BLOCK>(m1, m2, ok, sa, sb, so);
  if (threadIdx.x == 0) {
    double* p = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
    p[0] = m1;
    p[1] = m2;
    p[2] = (double)ok;
  }
}

// ---- Richardson combination (src/distribution_slice_compute_richardson.cpp:47-64)
// out[i + D j] = 2 (four fine cells) - coarse, or the coarse cell alone.
// partial_tp: [slice][gridDim.x][2] double-double block sums of out.
__global__ void __launch_bounds__(QB_PLAIN_BLOCK)
k_rich2d(int D, int richardson, const double* __restrict__ coarse,
         const double* __restrict__ fine, double* __restrict__ out,
         double* __restrict__ partial_tp) {
  __shared__ double sh[QB_PLAIN_BLOCK], sl[QB_PLAIN_BLOCK];
  const int cell = blockIdx.x * QB_PLAIN_BLOCK + threadIdx.x;
  const size_t sl_c = (size_t)blockIdx.y * D * D;
  double v = 0.0;
  if (cell < D * D) {
    const int i = cell % D, j = cell / D;
    v = coarse[sl_c + cell];
    if (richardson) {
      const double* f = fine + (size_t)blockIdx.y * 4 * D * D;
      const size_t F = (size_t)2 * D;
      const double dp = ((f[F * (2 * j) + 2 * i] + f[F * (2 * j) + 2 * i + 1]) +
                         f[F * (2 * j + 1) + 2 * i]) +
                        f[F * (2 * j + 1) + 2 * i + 1];
      v = 2.0 * dp - v;
    }
    out[sl_c + cell] = v;
  }
  const dd t = block_sum_dd<QB_PLAIN_BLOCK>(make_dd(v, 0.0), sh, sl);
  if (threadIdx.x == 0) {
    double* p = partial_tp + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    p[0] = t.hi;
    p[1] = t.lo;
  }
}

// ---- per-slice summary -------------------------------------------------------
// One thread per slice sums the block partials in index order.
__global__ void k_final2d(int n, int richardson, int nb_c, int nb_f, int nb_o,
                          const double* __restrict__ part_c, const double* __restrict__ part_f,
                          const double* __restrict__ part_tp, double* __restrict__ summary) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double m1c = 0, m2c = 0, m1f = 0, m2f = 0;
  int ok = 1;
  for (int b = 0; b < nb_c; b++) {
    const double* p = part_c + ((size_t)s * nb_c + b) * 3;
    m1c += p[0];
    m2c += p[1];
    ok &= (p[2] != 0.0);
  }
  if (richardson)
    for (int b = 0; b < nb_f; b++) {
      const double* p = part_f + ((size_t)s * nb_f + b) * 3;
      m1f += p[0];
      m2f += p[1];
    }
  dd tp = make_dd(0.0, 0.0);
  for (int b = 0; b < nb_o; b++) {
    const double* p = part_tp + ((size_t)s * nb_o + b) * 2;
    tp = dd_add(tp, make_dd(p[0], p[1]));
  }
  double* o = summary + (size_t)s * 8;
  o[0] = tp.hi;
  o[1] = tp.lo;
  o[2] = richardson ? 2.0 * m1f - m1c : m1c;
  o[3] = richardson ? 2.0 * m2f - m2c : m2c;
  o[4] = (double)ok;
  o[5] = o[6] = o[7] = 0.0;
}

// ---- one-dimensional slices --------------------------------------------------
// values: [slice][NP] integrand values at the coarse + fine abscissae.
__global__ void k_vals1d(DevConsts c, int kind, int NP, const DevSlice* __restrict__ slices,
                         const TabDesc* __restrict__ desc, const dd* __restrict__ gx,
                         double* __restrict__ values) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP) return;
  const DevSlice s = slices[blockIdx.y];
  values[(size_t)blockIdx.y * NP + i] = value_1d(c, kind, gx[i], desc[s.tab_a], s.eta_shift);
}

__global__ void __launch_bounds__(QB_PLAIN_BLOCK)
k_cells1d(int D, int richardson, const DevSlice* __restrict__ slices,
          const double* __restrict__ values, const double* __restrict__ gw,
          double* __restrict__ out, double* __restrict__ partial_tp) {
  __shared__ double sh[QB_PLAIN_BLOCK], sl[QB_PLAIN_BLOCK];
  const int I = blockIdx.x * QB_PLAIN_BLOCK + threadIdx.x;
  const int NP = table_points(D);
  const DevSlice s = slices[blockIdx.y];
  double v = 0.0;
  if (I < D) {
    const double* vc = values + (size_t)blockIdx.y * NP;
    v = pass1d_cell(vc, gw[I] * s.scale_a, I);
    if (richardson) {
      const double* vf = vc + pass_offset(D, 1);
      const double* wf = gw + width_offset(D, 1);
      const double f = pass1d_cell(vf, wf[2 * I] * s.scale_a, 2 * I) +
                       pass1d_cell(vf, wf[2 * I + 1] * s.scale_a, 2 * I + 1);
      v = 2.0 * f - v;
    }
    out[(size_t)blockIdx.y * D + I] = v;
  }
  const dd t = block_sum_dd<QB_PLAIN_BLOCK>(make_dd(v, 0.0), sh, sl);
  if (threadIdx.x == 0) {
    double* p = partial_tp + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    p[0] = t.hi;
    p[1] = t.lo;
  }
}

__global__ void k_final1d(int n, int nb_o, const double* __restrict__ part_tp,
                          double* __restrict__ summary) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  dd tp = make_dd(0.0, 0.0);
  for (int b = 0; b < nb_o; b++) {
    const double* p = part_tp + ((size_t)s * nb_o + b) * 2;
    tp = dd_add(tp, make_dd(p[0], p[1]));
  }
  double* o = summary + (size_t)s * 8;
  o[0] = tp.hi;
  o[1] = tp.lo;
  o[2] = o[3] = 0.0;
  o[4] = 1.0;
  o[5] = o[6] = o[7] = 0.0;
}

// ---- FP64 peak microbenchmark -------------------------------------------------
// 8 independent DFMA chains per thread, register resident.
__global__ void k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace qb200
