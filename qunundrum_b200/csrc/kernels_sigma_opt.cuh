// kernels_sigma_opt.cuh -- CUDA kernels of the sigma-optimal method (see sigma_opt.cuh).
//
// Per (slice, pass) the reference's serial walk sigma_p = f_p(sigma_{p-1}) over the
// (2 D' + 1)^2 points is solved as a fixed point, in parallel over all points:
//
//   guess_p  = sigma_0 for all p                      (k_so_first: arg-min at the first point)
//   repeat   sigma_p = f_p(guess_{p-1})               (k_so_step : one thread per point)
//            guess   = inclusive prefix-min(sigma)    (k_so_scan : one block per (slice, pass))
//   until guess did not change.
//
// f_p(s) <= s and f_p is monotone, so every iterate is an upper bound of the true chain and the
// iteration stops exactly at it (proof in DESIGN.md); in practice 2-4 iterations. Points where
// the reference's increasing search would fire (sigma walked down to 1) break the f_p(s) <= s
// premise; they are detected and reported as unsupported instead of returning wrong cells.
#pragma once

#include <cuda_runtime.h>

#include "kernels_plain.cuh"
#include "sigma_opt.cuh"

namespace qb200 {

struct SoLayout {
  int D;         // coarse dimension
  int passes;    // 1 or 2
  int n_c, n_f;  // points per pass: (2 D + 1)^2, (4 D + 1)^2
  int stride;    // points per slice = n_c + n_f (n_f = 0 for a single pass)
};

__host__ __device__ inline int so_side(const SoLayout& L, int pass) {
  return pass ? 4 * L.D + 1 : 2 * L.D + 1;
}
__host__ __device__ inline int so_npts(const SoLayout& L, int pass) { return pass ? L.n_f : L.n_c; }
__host__ __device__ inline size_t so_base(const SoLayout& L, int slice, int pass) {
  return (size_t)slice * L.stride + (pass ? L.n_c : 0);
}

// The sigma-independent part of point p of (slice, pass).
__device__ __forceinline__ SoPoint so_point(const DevConsts& c, const SoLayout& L, const DevSlice& s,
                                            const TabDesc* desc_a, const TabDesc* desc_b,
                                            const dd* gx, const AxisR* tab_b, int pass, int p) {
  const int side = so_side(L, pass);
  const int i = p / side, j = p - i * side;  // alpha_d outer, alpha_r inner (reference loop order)
  const int off = pass_offset(L.D, pass);
  SoPoint pt;
  pt.xd_ = grid_x(gx[off + i], desc_a[s.tab_a].k_abs, desc_a[s.tab_a].sign, c.m);
  pt.xr_ = grid_x(gx[off + j], desc_b[s.tab_b].k_abs, desc_b[s.tab_b].sign, c.m);
  pt.t2 = tab_b[(size_t)s.tab_b * table_points(L.D) + off + j].t2;
  pt.h = fabs(pt.xd_.hi) + fabs(pt.xr_.hi);
  return pt;
}

// ---- first point: arg-min over sigma in [1, l - 2], smallest sigma among equals -------------
// grid.x = (slice, pass) pairs; sigma0[pair] = 0 when no sigma gives error < 1.
__global__ void __launch_bounds__(256)
k_so_first(DevConsts c, SigmaOptConsts q, SoLayout L, const DevSlice* __restrict__ slices,
           const TabDesc* __restrict__ desc_a, const TabDesc* __restrict__ desc_b,
           const dd* __restrict__ gx, const AxisR* __restrict__ tab_b, int* __restrict__ sigma0) {
  __shared__ double sf[256];
  __shared__ int se[256], ss[256];
  const int pair = blockIdx.x;
  const int slice = pair / L.passes, pass = pair % L.passes;
  const SoPoint pt = so_point(c, L, slices[slice], desc_a, desc_b, gx, tab_b, pass, 0);
  xd best = xd_make(1.0, c.m);  // error < 1  <=>  e < 2^m
  int bs = 0;
  for (int sigma = 1 + threadIdx.x; sigma < c.l - 1; sigma += blockDim.x) {
    double n;
    xd e;
    so_eval(c, q, pt, sigma, &n, &e);
    if (xd_less(e, best)) {  // strict: keeps the smallest sigma of this thread's equal values
      best = e;
      bs = sigma;
    }
  }
  sf[threadIdx.x] = best.f;
  se[threadIdx.x] = best.e;
  ss[threadIdx.x] = bs;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      xd a, b;
      a.f = sf[threadIdx.x];
      a.e = se[threadIdx.x];
      b.f = sf[threadIdx.x + st];
      b.e = se[threadIdx.x + st];
      const int sa = ss[threadIdx.x], sb = ss[threadIdx.x + st];
      const bool take_b = sb != 0 && (sa == 0 || xd_less(b, a) || (!xd_less(a, b) && sb < sa));
      if (take_b) {
        sf[threadIdx.x] = b.f;
        se[threadIdx.x] = b.e;
        ss[threadIdx.x] = sb;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) sigma0[pair] = ss[0];
}

// ---- one relaxation step: sigma_p = f_p(guess_{p-1}) ---------------------------------------------
// grid.x over points, grid.y = (slice, pass) pairs of the chunk.
// status[pair] bit 0: some point used the increasing search (unsupported).
__global__ void __launch_bounds__(128)
k_so_step(DevConsts c, SigmaOptConsts q, SoLayout L, const DevSlice* __restrict__ slices,
          const TabDesc* __restrict__ desc_a, const TabDesc* __restrict__ desc_b,
          const dd* __restrict__ gx, const AxisR* __restrict__ tab_b,
          const int* __restrict__ sigma0, const int* __restrict__ guess, int* __restrict__ sigma,
          double* __restrict__ norm, double* __restrict__ erra, int* __restrict__ status) {
  const int pair = blockIdx.y;
  const int slice = pair / L.passes, pass = pair % L.passes;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= so_npts(L, pass)) return;
  const int s0 = sigma0[pair];
  if (s0 == 0) return;
  const size_t base = so_base(L, slice, pass);
  const SoPoint pt = so_point(c, L, slices[slice], desc_a, desc_b, gx, tab_b, pass, p);
  double n;
  xd e;
  int sg;
  if (p == 0) {
    sg = s0;
    so_eval(c, q, pt, sg, &n, &e);
  } else {
    bool inc;
    sg = so_adjust(c, q, pt, guess[base + p - 1], &n, &e, &inc);
    if (inc) atomicOr(status + pair, 1);
  }
  // A_p = pi h n r/2^m (2 + s) 2^(sigma - sigma0); sign bit of the stored sigma = "not bounded"
  const int sl = sg - c.l;
  const double ph = 3.14159265358979323846 * pt.h;
  const double sv = sl > -1000 ? ldexp(ph, sl) : 0.0;
  erra[base + p] = ldexp(ph * (2.0 + sv) * n * c.r_m, max(sg - s0, -1000));
  norm[base + p] = n;
  sigma[base + p] = so_bounded(c, n, e) ? sg : -sg;
}

// ---- guess = inclusive prefix-min(|sigma|); *changed |= (guess moved) -------------------------
__global__ void __launch_bounds__(1024)
k_so_scan(SoLayout L, const int* __restrict__ sigma, int* __restrict__ guess,
          int* __restrict__ changed) {
  __shared__ int warp_min[32];
  __shared__ int carry_s;
  const int pair = blockIdx.x;
  const int slice = pair / L.passes, pass = pair % L.passes;
  const int n = so_npts(L, pass);
  const size_t base = so_base(L, slice, pass);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0x7fffffff;
  __syncthreads();
  int moved = 0;
  for (int start = 0; start < n; start += 1024) {
    const int p = start + threadIdx.x;
    int v = p < n ? abs(sigma[base + p]) : 0x7fffffff;
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v = min(v, o);
    }
    if (lane == 31) warp_min[w] = v;
    __syncthreads();
    if (w == 0) {
      int m = warp_min[lane];
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, m, off);
        if (lane >= off) m = min(m, o);
      }
      warp_min[lane] = m;
    }
    __syncthreads();
    int pre = carry_s;
    if (w > 0) pre = min(pre, warp_min[w - 1]);
    v = min(v, pre);
    if (p < n) {
      if (guess[base + p] != v) moved = 1;
      guess[base + p] = v;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = v;
    __syncthreads();
  }
  if (moved) atomicOr(changed, 1);
}

__global__ void k_so_fill(size_t n_total, SoLayout L, const int* __restrict__ sigma0,
                          int* __restrict__ guess) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_total) return;
  const int slice = (int)(k / L.stride);
  const int pass = (k - (size_t)slice * L.stride) >= (size_t)L.n_c ? 1 : 0;
  guess[k] = sigma0[slice * L.passes + pass];
}

// ---- Simpson cells of one pass from the point arrays ---------------------------------------------
// partial: [slice][gridDim.x][3] = (sum A, sum C, bounded)
__global__ void __launch_bounds__(QB_PLAIN_BLOCK)
k_so_cells(DevConsts c, SoLayout L, int pass, const DevSlice* __restrict__ slices,
           const double* __restrict__ gw, const int* __restrict__ sigma0,
           const int* __restrict__ sigma, const double* __restrict__ norm,
           const double* __restrict__ erra, double* __restrict__ cells_pass,
           double* __restrict__ partial) {
  __shared__ double sa[QB_PLAIN_BLOCK], sb[QB_PLAIN_BLOCK];
  __shared__ int so[QB_PLAIN_BLOCK];
  const int Dp = pass ? 2 * L.D : L.D;
  const int side = 2 * Dp + 1;
  const DevSlice s = slices[blockIdx.y];
  const int cell = blockIdx.x * QB_PLAIN_BLOCK + threadIdx.x;
  const size_t base = so_base(L, blockIdx.y, pass);
  const int s0 = sigma0[blockIdx.y * L.passes + pass];
  double A = 0.0, Cc = 0.0;
  int ok = 1;
  if (cell < Dp * Dp && s0 != 0) {
    const int I = cell % Dp, J = cell / Dp;
    const double* w = gw + width_offset(L.D, pass);
    const double w3[3] = {1.0, 4.0, 1.0};
    double acc = 0.0;
    for (int a = 0; a < 3; a++)      // alpha_d index 2 I + a is the OUTER index of the point arrays
      for (int b = 0; b < 3; b++) {
        const size_t p = base + (size_t)(2 * I + a) * side + (2 * J + b);
        const double wt = w3[a] * w3[b];
        const int sg = sigma[p];
        acc = fma(wt, norm[p], acc);
        A = fma(wt, erra[p], A);
        Cc = fma(wt, ldexp(1.0, min(s0 - abs(sg), 1000)), Cc);
        ok &= sg > 0;
      }
    const double f = (w[I] * s.scale_a) * (w[J] * s.scale_b) / 36.0;
    cells_pass[(size_t)blockIdx.y * Dp * Dp + cell] = acc * f * c.r_m;
    A *= f;
    Cc *= f;
  }
  block_sum2_and<QB_PLAIN_BLOCK>(A, Cc, ok, sa, sb, so);
  if (threadIdx.x == 0) {
    double* pp = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
    pp[0] = A;
    pp[1] = Cc;
    pp[2] = (double)ok;
  }
}

// summary slots: 0,1 mass (dd); 2,3 (A, C) coarse; 5,6 (A, C) fine; 7 sigma0_c + 65536 sigma0_f;
// 4 bounded (coarse pass), or -1 when the slice is unsupported / has no admissible sigma.
__global__ void k_so_final(int n, SoLayout L, int nb_c, int nb_f, int nb_o,
                           const double* __restrict__ part_c, const double* __restrict__ part_f,
                           const double* __restrict__ part_tp, const int* __restrict__ sigma0,
                           const int* __restrict__ status, double* __restrict__ summary) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double Ac = 0, Cc = 0, Af = 0, Cf = 0;
  int ok = 1;
  for (int b = 0; b < nb_c; b++) {
    const double* p = part_c + ((size_t)s * nb_c + b) * 3;
    Ac += p[0];
    Cc += p[1];
    ok &= (p[2] != 0.0);
  }
  if (L.passes == 2)
    for (int b = 0; b < nb_f; b++) {
      const double* p = part_f + ((size_t)s * nb_f + b) * 3;
      Af += p[0];
      Cf += p[1];
    }
  dd tp = make_dd(0.0, 0.0);
  for (int b = 0; b < nb_o; b++) {
    const double* p = part_tp + ((size_t)s * nb_o + b) * 2;
    tp = dd_add(tp, make_dd(p[0], p[1]));
  }
  const int s0c = sigma0[s * L.passes], s0f = L.passes == 2 ? sigma0[s * L.passes + 1] : 0;
  bool bad = s0c == 0 || (L.passes == 2 && s0f == 0);
  for (int k = 0; k < L.passes; k++) bad = bad || (status[s * L.passes + k] != 0);
  double* o = summary + (size_t)s * 8;
  o[0] = tp.hi;
  o[1] = tp.lo;
  o[2] = Ac;
  o[3] = Cc;
  o[4] = bad ? -1.0 : (double)ok;
  o[5] = Af;
  o[6] = Cf;
  o[7] = (double)(s0c + 65536 * s0f);
}

// ---- the large-l fast path: the walk as a prefix minimum (sigma_opt.cuh, "closed form") ------------
// One block per slice, both passes one after the other; the points of a pass are taken 256 at a
// time IN WALK ORDER (alpha_d outer, alpha_r inner), so the running minimum is a block scan with a
// carry. Per point: the norm by angle addition from the axis tables of the companion plan (the
// quick method's: kappa = -d/r; the fused kernel's separable sine), sigma*_p without a logarithm,
// the point's share of the two error sums with its separable Simpson weight, the bound test.
// Cells and mass come from the fused kernel of the companion plan; this kernel fills the rest of
// the summary (slots 2 .. 7 as k_so_final) and raises *fallback when a point leaves the range in
// which the closed form is proven (64 <= sigma_p <= l - 60).
#define QB_SOF_BLOCK 256
#define QB_SOF_NONE 0x3fffffff

__global__ void __launch_bounds__(QB_SOF_BLOCK)
k_so_fast(DevConsts c, SoLayout L, const DevSlice* __restrict__ slices, const AxisD* __restrict__ tab_a,
          const AxisR* __restrict__ tab_b, const double* __restrict__ gw, double* __restrict__ summary,
          int* __restrict__ fallback) {
  constexpr int NW = QB_SOF_BLOCK / 32;
  __shared__ int warp_min[2][NW];
  __shared__ int first_s;
  __shared__ double sa[QB_SOF_BLOCK], sb[QB_SOF_BLOCK];
  __shared__ int so[QB_SOF_BLOCK];
  extern __shared__ double s_wgt[];  // Simpson weight of every abscissa of an axis (4 D + 1 at most)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const DevSlice s = slices[blockIdx.x];
  const int NP = table_points(L.D);
  const double PI = 3.14159265358979323846;
  double out_a[2] = {0.0, 0.0}, out_c[2] = {0.0, 0.0};
  int sigma0[2] = {0, 0};
  int bounded_c = 1, bad = 0;
  for (int pass = 0; pass < L.passes; pass++) {
    const int Dp = pass ? 2 * L.D : L.D, side = 2 * Dp + 1, npts = side * side;
    const int off = pass_offset(L.D, pass);
    const AxisD* ta = tab_a + (size_t)s.tab_a * NP + off;
    const AxisR* tb = tab_b + (size_t)s.tab_b * NP + off;
    const double* wd = gw + width_offset(L.D, pass);
    for (int i = tid; i < side; i += QB_SOF_BLOCK) s_wgt[i] = so_axis_weight(wd, Dp, i);
    __syncthreads();
    double A = 0.0, Cc = 0.0;
    int ok = 1, s0 = 0, buf = 0;
    int carry = QB_SOF_NONE;  // running minimum of the chunks before this one (the same in every thread)
    int i = tid / side, j = tid - i * side;  // point p = start + tid of the walk: (i, j) = (alpha_d, alpha_r) index
    for (int start = 0; start < npts; start += QB_SOF_BLOCK) {
      const bool live = start + tid < npts;
      double n = 0.0, ph = 0.0, wgt = 0.0;
      int v = QB_SOF_NONE;
      if (live) {
        const AxisD d = ta[i];
        const AxisR r = tb[j];
        const double u = (d.xh + r.yh) + (d.xl + r.yl);
        // sin(pi u) / (pi u): the series next to the ridge, angle addition elsewhere
        const double t1 = fabs(u) < 0.0625 ? sincpi_small(u) : fma(d.sd, r.cr, d.cd * r.sr) / (QB_PI_HI * u);
        n = t1 * t1 * r.t2;
        ph = PI * (fabs(d.xh) + r.b);
        const double a = ph * n * c.r_m;
        v = so_fast_sigma_star(c.l, a);
        if (start + tid == 0) {
          v = so_fast_sigma_first(c.l, a);
          first_s = v;  // sigma_0 of the pass
        }
        wgt = s_wgt[i] * s_wgt[j];
      }
      // inclusive running minimum along the walk, carried from chunk to chunk: one barrier per chunk
      // (the warps' minima go to alternating buffers, every thread folds all of them into its carry)
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = min(v, t);
      }
      if (lane == 31) warp_min[buf][w] = v;
      __syncthreads();
      int pre = carry;
#pragma unroll
      for (int k = 0; k < NW; k++) {
        const int m = warp_min[buf][k];
        if (k < w) pre = min(pre, m);
        carry = min(carry, m);
      }
      buf ^= 1;
      v = min(v, pre);
      if (start == 0) s0 = first_s;
      if (live) {
        const int sg = v;
        if (sg < 64 || sg > c.l - 60) {
          bad = 1;
        } else {
          double era, ct;
          so_fast_terms(c, ph, n, sg, s0, &era, &ct);
          A = fma(wgt, era, A);
          Cc = fma(wgt, ct, Cc);
          if (pass == 0 && !so_fast_bounded(c, ph, n, sg)) ok = 0;
        }
      }
      j += QB_SOF_BLOCK;
      while (j >= side) {
        j -= side;
        i++;
      }
    }
    const double f = s.scale_a * s.scale_b / 36.0;
    A *= f;
    Cc *= f;
    block_sum2_and<QB_SOF_BLOCK>(A, Cc, ok, sa, sb, so);
    out_a[pass] = A;
    out_c[pass] = Cc;
    sigma0[pass] = s0;
    if (pass == 0) bounded_c = ok;
    __syncthreads();
  }
  bad = __syncthreads_or(bad);
  if (tid == 0) {
    double* o = summary + (size_t)blockIdx.x * 8;
    o[2] = out_a[0];
    o[3] = out_c[0];
    o[4] = bad ? -1.0 : (double)bounded_c;
    o[5] = out_a[1];
    o[6] = out_c[1];
    o[7] = (double)(sigma0[0] + 65536 * sigma0[1]);
    if (bad) atomicOr(fallback, 1);
  }
}

}  // namespace qb200
