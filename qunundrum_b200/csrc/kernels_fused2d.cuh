// kernels_fused2d.cuh -- the fused two-dimensional Richardson slice kernel.
//
// Replaces, for one batch of slices, what the reference does in
// distribution_slice_compute_richardson (src/distribution_slice_compute_richardson.cpp:17-73):
// two calls of distribution_slice_compute (src/distribution_slice_compute.cpp:38-453)
// at dimensions D and 2 D, each a (2 D' + 1)^2 grid of probability_approx
// (src/probability.cpp:150-288) evaluations followed by 3 x 3 Simpson sums, and
// the combination 2 * (four fine cells) - coarse cell.
//
// Work decomposition (B200: 148 SMs x 4 sub-partitions, FP64 pipe bound):
//   * a WARP owns a tile of 32 x 32 coarse cells of one slice; lane <-> coarse
//     row I (alpha_d, the fast index of norm_matrix, so stores coalesce), and the
//     warp marches over the tile's alpha_r columns;
//   * per lane, the alpha_d-only quantities of its five abscissae (four fine
//     points of the cell and the coarse mid point) live in registers; the
//     alpha_r-only quantities of the current column (including the whole second
//     factor T2 of the approximation, folded with the Simpson weight) are one
//     warp-uniform 80-byte record read through L1;
//   * sin(pi u), u = x_d + kappa x_r, is NOT evaluated per point: with
//     sin/cos(pi x_d) per row and sin/cos(pi kappa x_r) per column (double-double
//     reduced, in the axis tables) it is one multiply and one FMA,
//     sin(pi u) = sin(pi x_d) cos(pi y) + cos(pi x_d) sin(pi y);
//     near the ridge (|u| < 1/16), where that would lose relative accuracy, the
//     point falls back to a polynomial in u^2 (u itself is exact there);
//   * 1 / u^2 is MUFU.RCP64H plus one cubic Newton step (3 DFMA);
//   * every integrand value is computed once: the fine-grid row shared by
//     vertically adjacent cells comes from the neighbouring lane by shuffle, the
//     column shared by horizontally adjacent cells is applied twice with two
//     weights, and coarse main points reuse the fine values. 19 evaluations per
//     output cell instead of the reference's 20.09;
//   * the only redundancy is the one halo row below lane 31, evaluated in a
//     lanes<->columns pre-pass (1 % of the tile's work), which keeps warps fully
//     independent: no __syncthreads, no inter-warp traffic.
//
// Outputs per tile: 32 x 32 cells (coalesced 256-byte row segments) and a
// 4-double partial (mass, error moments, bound flag) summed per slice in a
// fixed order by k_fused_final => bitwise reproducible results.
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels_plain.cuh"
#include "plan.hpp"

namespace qb200 {

#define QB_FUSED_REC 10         // doubles per column record
#define QB_FUSED_PART_STRIDE 4  // doubles per tile partial
#define QB_FUSED_WARPS 4

struct FusedConst {
  double c1, c2, c3;   // 1/sinc^2(z) = 1 + c1 w + c2 w^2 + c3 w^3, w = u^2, z = pi u / Lambda
  double cs, e0s, r_m; // error bound pieces (src/probability.cpp:252-281)
  int D, nb, NP, ncol; // nb = D / 32, ncol = 5 D + 1 records per alpha_r table
  unsigned n_tiles;    // all classes
  unsigned tile_base, tile_end;  // the range this launch covers
};

// A slice as the fused kernel sees it. Slices are sorted by class so that each
// class is one contiguous tile range and one launch; `slot` is the position in
// the caller's order (where the cells and the summary go).
struct FusedSlice {
  int tab_a, tab_b;
  int slot, cls;
  double scale_a;
};

// A contiguous range of the caller's slices, integrated by at most three
// launches (one per class present). Device-resident runs use a single chunk;
// the synchronous host API uses several so that the device-to-host copy of
// chunk c overlaps the kernels of chunk c + 1.
struct FusedChunk {
  unsigned slot_begin, slot_end;
  unsigned class_tiles[4];  // tile_base of class 0, 1, 2 and the end, in sorted order
};

struct FusedPlan2D {
  FusedConst k;
  int mode = 0;        // 0: Lambda sin(pi u / Lambda) == pi u; 1: series correction
  bool has_err = false, has_m2 = false, has_bound = false;
  std::vector<unsigned char> host_unbounded;  // per slice, when the bound is decided on the host
  std::vector<FusedSlice> fslices;            // sorted by (chunk, class)
  std::vector<FusedChunk> chunks;
  size_t cols_bytes = 0;                      // size of the column-record buffer
};

// ---- column records ---------------------------------------------------------
// For alpha_r table t and record R = 5 J + k:
//   k = 0      fine main column h = 4 J  (also coarse main column)
//   k = 1,2,3  fine columns h = 4 J + k
//   k = 4      coarse mid column of cell J
//   R = 5 D    the last main column h = 4 D
// Fields: yh, yl, sr, cr, wF, wF2, wC, wC2, b, t2p.
//   wF  : fine Simpson weight of the column within cell J   (x T2 / pi^2)
//   wF2 : fine weight of the column as LAST column of cell J - 1 (k = 0); wF * b for k = 1, 2, 3
//   wC / wC2 : the same for the coarse pass (k = 4: wC2 = wC * b)
__global__ void k_fused_cols(int D, int m, const TabDesc* __restrict__ desc_b,
                             const AxisR* __restrict__ tab_b, const double* __restrict__ gw,
                             double* __restrict__ cols) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = 5 * D + 1;
  if (R >= ncol) return;
  const int t = blockIdx.y;
  const int NP = table_points(D);
  const AxisR* tc = tab_b + (size_t)t * NP;       // coarse interleaved
  const AxisR* tf = tc + pass_offset(D, 1);       // fine interleaved
  const double* wc = gw;                          // coarse widths [D]
  const double* wf = gw + width_offset(D, 1);     // fine widths [2 D]
  const double sb = pow2i(desc_b[t].k_abs - m) / 6.0;
  const double inv_pi2 = 0.101321183642337771443879463209;  // 1 / pi^2
  const int J = R / 5, k = R % 5;
  AxisR a;
  double wF = 0.0, wF2 = 0.0, wC = 0.0, wC2 = 0.0;
  if (k == 4) {
    a = tc[2 * J + 1];
    wC = 4.0 * wc[J];
  } else {
    a = tf[4 * J + k];
    if (k == 0) {
      if (J < D) {
        wF = wf[2 * J];
        wC = wc[J];
      }
      if (J >= 1) {
        wF2 = wf[2 * J - 1];
        wC2 = wc[J - 1];
      }
    } else if (k == 1) {
      wF = 4.0 * wf[2 * J];
    } else if (k == 2) {
      wF = wf[2 * J] + wf[2 * J + 1];
    } else {
      wF = 4.0 * wf[2 * J + 1];
    }
  }
  const double t2p = a.t2 * inv_pi2;
  const double f = sb * t2p;
  double* o = cols + ((size_t)t * ncol + R) * QB_FUSED_REC;
  o[0] = a.yh;
  o[1] = a.yl;
  o[2] = a.sr;
  o[3] = a.cr;
  o[4] = wF * f;
  o[5] = wF2 * f;
  o[6] = wC * f;
  o[7] = wC2 * f;
  // Interior fine columns (k = 1, 2, 3) have no "last column" weight and the coarse mid column
  // (k = 4) has no wC2: those slots carry the error-moment weights wF * b / wC * b, the very
  // products the march would otherwise form per column in every lane (bit-identical).
  if (k >= 1 && k <= 3) o[5] = o[4] * a.b;
  if (k == 4) o[7] = o[6] * a.b;
  o[8] = a.b;
  o[9] = t2p;
}

// Axis tables AND column records in one launch (grid.y: the alpha_d tables, the alpha_r tables,
// then one row of blocks per alpha_r table for its column records). A column record needs one
// abscissa of its table; it evaluates that abscissa itself (the same axis_r_point, the same bits)
// instead of waiting for the table kernel: one dependent launch less per step, which is what
// counts when a GPU holds an eighth of a distribution.
__global__ void k_fused_prologue(DevConsts c, int D, int NP, int n_tab_a, int n_tab_b,
                                 const TabDesc* __restrict__ desc_a, const TabDesc* __restrict__ desc_b,
                                 const dd* __restrict__ gx, const double* __restrict__ gw,
                                 AxisD* __restrict__ tab_a, AxisR* __restrict__ tab_b,
                                 double* __restrict__ cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (y < n_tab_a + n_tab_b) {
    if (i >= NP) return;
    const dd g = gx[i];
    if (y < n_tab_a) {
      AxisD o;
      axis_d_point(c, g, desc_a[y], &o);
      tab_a[(size_t)y * NP + i] = o;
    } else {
      AxisR o;
      axis_r_point(c, g, desc_b[y - n_tab_a], &o);
      tab_b[(size_t)(y - n_tab_a) * NP + i] = o;
    }
    return;
  }
  const int R = i;
  const int ncol = 5 * D + 1;
  if (R >= ncol) return;
  const int t = y - n_tab_a - n_tab_b;
  const double* wc = gw;                          // coarse widths [D]
  const double* wf = gw + width_offset(D, 1);     // fine widths [2 D]
  const double sb = pow2i(desc_b[t].k_abs - c.m) / 6.0;
  const double inv_pi2 = 0.101321183642337771443879463209;  // 1 / pi^2
  const int J = R / 5, k = R % 5;
  AxisR a;
  double wF = 0.0, wF2 = 0.0, wC = 0.0, wC2 = 0.0;
  if (k == 4) {
    axis_r_point(c, gx[2 * J + 1], desc_b[t], &a);                       // coarse interleaved 2 J + 1
    wC = 4.0 * wc[J];
  } else {
    axis_r_point(c, gx[pass_offset(D, 1) + 4 * J + k], desc_b[t], &a);   // fine interleaved 4 J + k
    if (k == 0) {
      if (J < D) {
        wF = wf[2 * J];
        wC = wc[J];
      }
      if (J >= 1) {
        wF2 = wf[2 * J - 1];
        wC2 = wc[J - 1];
      }
    } else if (k == 1) {
      wF = 4.0 * wf[2 * J];
    } else if (k == 2) {
      wF = wf[2 * J] + wf[2 * J + 1];
    } else {
      wF = 4.0 * wf[2 * J + 1];
    }
  }
  const double t2p = a.t2 * inv_pi2;
  const double f = sb * t2p;
  double* o = cols + ((size_t)t * ncol + R) * QB_FUSED_REC;
  o[0] = a.yh;
  o[1] = a.yl;
  o[2] = a.sr;
  o[3] = a.cr;
  o[4] = wF * f;
  o[5] = wF2 * f;
  o[6] = wC * f;
  o[7] = wC2 * f;
  if (k >= 1 && k <= 3) o[5] = o[4] * a.b;   // error-moment weights, as k_fused_cols
  if (k == 4) o[7] = o[6] * a.b;
  o[8] = a.b;
  o[9] = t2p;
}

// ---- device helpers ---------------------------------------------------------
struct ColRec {
  double yh, yl, sr, cr, wF, wF2, wC, wC2, b, t2p;
};

__device__ __forceinline__ ColRec load_col(const double* __restrict__ p) {
  const double2 a = __ldg((const double2*)p);
  const double2 b = __ldg((const double2*)p + 1);
  const double2 c = __ldg((const double2*)p + 2);
  const double2 d = __ldg((const double2*)p + 3);
  const double2 e = __ldg((const double2*)p + 4);
  ColRec r;
  r.yh = a.x; r.yl = a.y; r.sr = b.x; r.cr = b.y;
  r.wF = c.x; r.wF2 = c.y; r.wC = d.x; r.wC2 = d.y;
  r.b = e.x; r.t2p = e.y;
  return r;
}

struct RowReg {
  double xh, xl, sd, cd;
};

__device__ __forceinline__ RowReg load_row(const AxisD* __restrict__ p) {
  const double2 a = __ldg((const double2*)p);
  const double2 b = __ldg((const double2*)p + 1);
  RowReg r;
  r.xh = a.x; r.xl = a.y; r.sd = b.x; r.cd = b.y;
  return r;
}

__device__ __forceinline__ double rcp_seed(double w) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
  return r;
}

// pi^2 * T1 at one point: (sin(pi u) / u)^2 [* 1 / sinc^2(pi u / Lambda)].
//
// Slice classes (decided on the host from the slice coordinates):
//   CLS 0  general: branch-free main path (valid for |u| >= 1/16); the caller
//          patches the rare small-|u| points (the ridge) with eval_small(), one
//          test per column, so that the independent evaluations of a column
//          interleave in the FP64 pipe;
//   CLS 1  every |u| of the slice is below 1/16  -> 7-term polynomial only;
//   CLS 2  every |u| of the slice is below 2^-9  -> 3-term polynomial only.
// (Slices far below the diagonal, |alpha| << 2^m, are entirely in class 1 / 2:
// 40 % of the slices of a default m + 10 ... m - 30 distribution.)
template <int MODE>
__device__ __forceinline__ double series_corr(double w, const FusedConst& k) {
  return MODE == 1 ? fma(w, fma(w, fma(w, k.c3, k.c2), k.c1), 1.0) : 1.0;
}

// SAME: x_d and y = kappa x_r have the same sign over the whole tile. Then |u| = |x_d| + |y|,
// the low words change u by at most half an ulp, and u = fl(x_h + y_h) is as accurate (one
// rounding instead of two) for one DADD instead of three.
template <int MODE, bool SAME>
__device__ __forceinline__ double eval_main(const RowReg& r, const ColRec& c, const FusedConst& k,
                                            double& u_out) {
  const double S = fma(r.sd, c.cr, r.cd * c.sr);
  const double u = SAME ? (r.xh + c.yh) : (r.xh + c.yh) + (r.xl + c.yl);
  const double r0 = rcp_seed(u);
  const double e = fma(-u, r0, 1.0);
  const double rr = fma(r0, fma(e, e, e), r0);
  const double q = S * rr;
  double T = q * q;
  if (MODE == 1) T *= series_corr<MODE>(u * u, k);
  u_out = u;
  return T;
}

__device__ __forceinline__ unsigned abs_hi(double u) {
  return (unsigned)__double2hiint(u) & 0x7fffffffu;
}
#define QB_SMALL_U 0x3FB00000u  // |u| < 1/16

// pi sinc(pi u) = sum (-1)^k pi^(2k+1) / (2k+1)! w^k
template <int MODE>
__device__ __forceinline__ double eval_small(double u, const FusedConst& k) {  // |u| < 1/16
  const double w = u * u;
  double p = 4.6630280576761256442e-4;
  p = fma(p, w, -7.3704309457143507773e-3);
  p = fma(p, w, 8.2145886611128228799e-2);
  p = fma(p, w, -5.9926452932079207689e-1);
  p = fma(p, w, 2.5501640398773454439);
  p = fma(p, w, -5.1677127800499700292);
  p = fma(p, w, 3.1415926535897932385);
  return (p * p) * series_corr<MODE>(w, k);
}
template <int MODE>
__device__ __forceinline__ double eval_tiny(double u, const FusedConst& k) {  // |u| < 2^-9
  const double w = u * u;
  double p = 2.5501640398773454439;
  p = fma(p, w, -5.1677127800499700292);
  p = fma(p, w, 3.1415926535897932385);
  return (p * p) * series_corr<MODE>(w, k);
}

template <int MODE, int CLS, bool SAME>
__device__ __forceinline__ double eval_cls(const RowReg& r, const ColRec& c, const FusedConst& k) {
  const double u = SAME ? (r.xh + c.yh) : (r.xh + c.yh) + (r.xl + c.yl);
  return CLS == 1 ? eval_small<MODE>(u, k) : eval_tiny<MODE>(u, k);
}

// Evaluate one column against up to five rows.
#define QB_EVAL4(c_, T0, T1, T2, T3)                                                        \
  double T0, T1, T2, T3;                                                                    \
  if (CLS == 0) {                                                                           \
    double u0_, u1_, u2_, u3_;                                                              \
    T0 = eval_main<MODE, SAME>(r0, c_, k, u0_);                                                   \
    T1 = eval_main<MODE, SAME>(r1, c_, k, u1_);                                                   \
    T2 = eval_main<MODE, SAME>(r2, c_, k, u2_);                                                   \
    T3 = eval_main<MODE, SAME>(r3, c_, k, u3_);                                                   \
    if (RIDGE && min(min(abs_hi(u0_), abs_hi(u1_)), min(abs_hi(u2_), abs_hi(u3_))) < QB_SMALL_U) { \
      if (abs_hi(u0_) < QB_SMALL_U) T0 = eval_small<MODE>(u0_, k);                          \
      if (abs_hi(u1_) < QB_SMALL_U) T1 = eval_small<MODE>(u1_, k);                          \
      if (abs_hi(u2_) < QB_SMALL_U) T2 = eval_small<MODE>(u2_, k);                          \
      if (abs_hi(u3_) < QB_SMALL_U) T3 = eval_small<MODE>(u3_, k);                          \
    }                                                                                       \
  } else {                                                                                  \
    T0 = eval_cls<MODE, CLS, SAME>(r0, c_, k);                                                    \
    T1 = eval_cls<MODE, CLS, SAME>(r1, c_, k);                                                    \
    T2 = eval_cls<MODE, CLS, SAME>(r2, c_, k);                                                    \
    T3 = eval_cls<MODE, CLS, SAME>(r3, c_, k);                                                    \
  }
#define QB_EVAL1(row_, c_, T)                                          \
  double T;                                                            \
  if (CLS == 0) {                                                      \
    double u_;                                                         \
    T = eval_main<MODE, SAME>(row_, c_, k, u_);                              \
    if (RIDGE && abs_hi(u_) < QB_SMALL_U) T = eval_small<MODE>(u_, k); \
  } else {                                                             \
    T = eval_cls<MODE, CLS, SAME>(row_, c_, k);                              \
  }
#define QB_EVAL2(rowa_, rowb_, c_, Ta, Tb)                              \
  double Ta, Tb;                                                        \
  if (CLS == 0) {                                                       \
    double ua_, ub_;                                                    \
    Ta = eval_main<MODE, SAME>(rowa_, c_, k, ua_);                            \
    Tb = eval_main<MODE, SAME>(rowb_, c_, k, ub_);                            \
    if (RIDGE && min(abs_hi(ua_), abs_hi(ub_)) < QB_SMALL_U) {          \
      if (abs_hi(ua_) < QB_SMALL_U) Ta = eval_small<MODE>(ua_, k);      \
      if (abs_hi(ub_) < QB_SMALL_U) Tb = eval_small<MODE>(ub_, k);      \
    }                                                                   \
  } else {                                                              \
    Ta = eval_cls<MODE, CLS, SAME>(rowa_, c_, k);                             \
    Tb = eval_cls<MODE, CLS, SAME>(rowb_, c_, k);                             \
  }

struct FusedArgs {
  FusedConst k;
  const FusedSlice* slices;
  const AxisD* tab_a;
  const double* cols;
  const double* gw;
  double* out;
  double* part;
  // the slice's summary is written by the warp that finishes its last tile (ticket per slice,
  // self-resetting); nullptr: k_fused_final does it in a launch of its own
  unsigned int* tickets;
  double* summary;
};

// The tile's partial, and -- when tickets are in use -- the slice's summary by the warp that
// finishes its last tile: tile partials -> summary in tile order (the lanes fetch 32 partials at a
// time, lane 0 adds them in order: the sums of k_fused_final, bit for bit). Not inlined: it runs
// once per tile and must not take registers from the march.
__device__ __noinline__ void fused_close_tile(double* part, unsigned int* tickets, double* summary,
                                              unsigned slot, unsigned per_slice, unsigned rem, int lane,
                                              double tp, double err1, double err2, int ok) {
  unsigned last = 0;
  if (lane == 0) {
    double* p = part + ((size_t)slot * per_slice + rem) * QB_FUSED_PART_STRIDE;
    p[0] = tp;
    p[1] = err1;
    p[2] = err2;
    p[3] = (double)ok;
    if (tickets) {
      __threadfence();
      last = atomicAdd(tickets + slot, 1u) == per_slice - 1 ? 1u : 0u;
    }
  }
  if (tickets == nullptr) return;
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  const double* pbase = part + (size_t)slot * per_slice * QB_FUSED_PART_STRIDE;
  dd stp = make_dd(0.0, 0.0);
  double m1 = 0.0, m2 = 0.0;
  int sok = 1;
  for (unsigned t0 = 0; t0 < per_slice; t0 += 32) {
    const unsigned cnt = per_slice - t0 < 32u ? per_slice - t0 : 32u;
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 1.0;
    if ((unsigned)lane < cnt) {
      const double* p = pbase + (size_t)(t0 + lane) * QB_FUSED_PART_STRIDE;
      q0 = __ldcg(p);
      q1 = __ldcg(p + 1);
      q2 = __ldcg(p + 2);
      q3 = __ldcg(p + 3);
    }
    for (unsigned t = 0; t < cnt; t++) {
      const double v0 = __shfl_sync(0xffffffffu, q0, (int)t), v1 = __shfl_sync(0xffffffffu, q1, (int)t);
      const double v2 = __shfl_sync(0xffffffffu, q2, (int)t), v3 = __shfl_sync(0xffffffffu, q3, (int)t);
      stp = dd_add_d(stp, v0);
      m1 += v1;
      m2 += v2;
      sok &= (v3 != 0.0);
    }
  }
  if (lane == 0) {
    double* o = summary + (size_t)slot * 8;
    o[0] = stp.hi;
    o[1] = stp.lo;
    o[2] = m1;
    o[3] = m2;
    o[4] = (double)sok;
    o[5] = o[6] = o[7] = 0.0;
    tickets[slot] = 0;  // ready for the next run
  }
}

#ifndef QB_FUSED_MIN_CTAS
#define QB_FUSED_MIN_CTAS 4
#endif
#define QB_TILE_RECS 161  // 5 * 32 + 1 column records per tile
#define QB_FUSED_SMEM_BYTES \
  (QB_FUSED_WARPS * (QB_TILE_RECS * QB_FUSED_REC + 2 * 32) * (int)sizeof(double))

__device__ __forceinline__ ColRec load_col_s(const double* p) {
  const double2 a = *(const double2*)p;
  const double2 b = *((const double2*)p + 1);
  const double2 c = *((const double2*)p + 2);
  const double2 d = *((const double2*)p + 3);
  const double2 e = *((const double2*)p + 4);
  ColRec r;
  r.yh = a.x; r.yl = a.y; r.sr = b.x; r.cr = b.y;
  r.wF = c.x; r.wF2 = c.y; r.wC = d.x; r.wC2 = d.y;
  r.b = e.x; r.t2p = e.y;
  return r;
}

#define QB_BOUND_TEST(T_, arow_, col_)                                              \
  if (HAS_BOUND) {                                                                  \
    const double sv_ = k.cs * ((arow_) + (col_).b);                                 \
    const double room_ = QB_ERROR_BOUND - sv_ * (2.0 + sv_);                        \
    ok &= (room_ >= 0.0) && (k.e0s <= room_ * ((T_) * (col_).t2p) * k.r_m);         \
  }

// The march of one warp over the 161 columns of its tile. RIDGE = false is used when no point
// of the tile can have |u| < 1/16 (decided once per tile from the tile's corner values): the
// loop is then free of tests and branches. Class 1 / 2 tiles never test.
template <int MODE, int CLS, bool RIDGE, bool SAME, bool HAS_ERR, bool HAS_M2, bool HAS_BOUND>
__device__ __forceinline__ void fused_march(const FusedConst& k, const RowReg& r0, const RowReg& r1,
                                            const RowReg& r2, const RowReg& r3, const RowReg& rc,
                                            const double D0, const double D1, const double DC,
                                            const double fd, const double* s_cols,
                                            const double* s_halo, double* outp, const int D,
                                            const int I, const int lane, const double* gwc,
                                            const double* gwf, double& tp, double& err1,
                                            double& err2, int& ok) {
  // ---- main march over the tile's columns --------------------------------------
  double sF0, sF1, sF2, sF3, sC0, sCm;                          // per-cell row sums
  double tF0 = 0, tF1 = 0, tF2 = 0, tF3 = 0, tC0 = 0, tCm = 0;  // tile totals of the row sums
  double bF0 = 0, bF1 = 0, bF2 = 0, bF3 = 0, bC0 = 0, bCm = 0;  // ... weighted by b
  double qF0 = 0, qF1 = 0, qF2 = 0, qF3 = 0, qC0 = 0, qCm = 0;  // ... weighted by b^2

  const double* cp = s_cols;
  {
    // first column of the tile: starts cell J0
    const ColRec c = load_col_s(cp);
    QB_EVAL4(c, T0, T1, T2, T3)
    QB_EVAL1(rc, c, Tm)
    sF0 = c.wF * T0; sF1 = c.wF * T1; sF2 = c.wF * T2; sF3 = c.wF * T3;
    sC0 = c.wC * T0; sCm = c.wC * Tm;
    if (HAS_ERR) {
      const double wb = c.wF * c.b, wcb = c.wC * c.b;
      bF0 = wb * T0; bF1 = wb * T1; bF2 = wb * T2; bF3 = wb * T3;
      bC0 = wcb * T0; bCm = wcb * Tm;
      if (HAS_M2) {
        const double wbb = wb * c.b, wcbb = wcb * c.b;
        qF0 = wbb * T0; qF1 = wbb * T1; qF2 = wbb * T2; qF3 = wbb * T3;
        qC0 = wcbb * T0; qCm = wcbb * Tm;
      }
    }
    QB_BOUND_TEST(T0, fabs(r0.xh), c)
    QB_BOUND_TEST(Tm, fabs(rc.xh), c)
  }

  for (int jj = 0; jj < 32; jj++) {
    cp += QB_FUSED_REC;
#pragma unroll
    for (int q = 1; q <= 3; q++) {  // fine interior columns
      const ColRec c = load_col_s(cp);
      cp += QB_FUSED_REC;
      QB_EVAL4(c, T0, T1, T2, T3)
      sF0 = fma(c.wF, T0, sF0); sF1 = fma(c.wF, T1, sF1);
      sF2 = fma(c.wF, T2, sF2); sF3 = fma(c.wF, T3, sF3);
      if (HAS_ERR) {
        const double wb = c.wF2;  // = wF * b, prepared by k_fused_cols
        bF0 = fma(wb, T0, bF0); bF1 = fma(wb, T1, bF1);
        bF2 = fma(wb, T2, bF2); bF3 = fma(wb, T3, bF3);
        if (HAS_M2) {
          const double wbb = wb * c.b;
          qF0 = fma(wbb, T0, qF0); qF1 = fma(wbb, T1, qF1);
          qF2 = fma(wbb, T2, qF2); qF3 = fma(wbb, T3, qF3);
        }
      }
    }
    {  // coarse mid column
      const ColRec c = load_col_s(cp);
      cp += QB_FUSED_REC;
      QB_EVAL2(r0, rc, c, T0, Tm)
      sC0 = fma(c.wC, T0, sC0); sCm = fma(c.wC, Tm, sCm);
      if (HAS_ERR) {
        const double wcb = c.wC2;  // = wC * b, prepared by k_fused_cols
        bC0 = fma(wcb, T0, bC0); bCm = fma(wcb, Tm, bCm);
        if (HAS_M2) {
          const double wcbb = wcb * c.b;
          qC0 = fma(wcbb, T0, qC0); qCm = fma(wcbb, Tm, qCm);
        }
      }
      QB_BOUND_TEST(T0, fabs(r0.xh), c)
      QB_BOUND_TEST(Tm, fabs(rc.xh), c)
    }
    {  // boundary column: closes cell J0 + jj, opens the next one
      const ColRec c = load_col_s(cp);
      QB_EVAL4(c, T0, T1, T2, T3)
      QB_EVAL1(rc, c, Tm)
      sF0 = fma(c.wF2, T0, sF0); sF1 = fma(c.wF2, T1, sF1);
      sF2 = fma(c.wF2, T2, sF2); sF3 = fma(c.wF2, T3, sF3);
      sC0 = fma(c.wC2, T0, sC0); sCm = fma(c.wC2, Tm, sCm);
      QB_BOUND_TEST(T0, fabs(r0.xh), c)
      QB_BOUND_TEST(Tm, fabs(rc.xh), c)
      // rows 4 / c2 of this lane are rows 0 / c0 of the lane below
      double nF = __shfl_down_sync(0xffffffffu, sF0, 1);
      double nC = __shfl_down_sync(0xffffffffu, sC0, 1);
      if (lane == 31) {
        nF = s_halo[jj];
        nC = s_halo[32 + jj];
      }
      const double f0 = fma(4.0, sF1, sF0) + sF2;   // fine cell 2 I    (times D0)
      const double f1 = fma(4.0, sF3, sF2) + nF;    // fine cell 2 I + 1 (times D1)
      const double fine = fma(D1, f1, D0 * f0);
      const double coarse = DC * (fma(4.0, sCm, sC0) + nC);
      const double cell = fma(2.0, fine, -coarse);
      outp[(size_t)jj * D] = cell;
      tp += cell;
      if (HAS_ERR) {
        tF0 += sF0; tF1 += sF1; tF2 += sF2; tF3 += sF3;
        tC0 += sC0; tCm += sCm;
        // the boundary column carries both weights, except on the tile's last column
        const double wsum = (jj == 31) ? 0.0 : c.wF;
        const double wcsum = (jj == 31) ? 0.0 : c.wC;
        const double wb = (c.wF2 + wsum) * c.b, wcb = (c.wC2 + wcsum) * c.b;
        bF0 = fma(wb, T0, bF0); bF1 = fma(wb, T1, bF1);
        bF2 = fma(wb, T2, bF2); bF3 = fma(wb, T3, bF3);
        bC0 = fma(wcb, T0, bC0); bCm = fma(wcb, Tm, bCm);
        if (HAS_M2) {
          const double wbb = wb * c.b, wcbb = wcb * c.b;
          qF0 = fma(wbb, T0, qF0); qF1 = fma(wbb, T1, qF1);
          qF2 = fma(wbb, T2, qF2); qF3 = fma(wbb, T3, qF3);
          qC0 = fma(wcbb, T0, qC0); qCm = fma(wcbb, Tm, qCm);
        }
      }
      sF0 = c.wF * T0; sF1 = c.wF * T1; sF2 = c.wF * T2; sF3 = c.wF * T3;
      sC0 = c.wC * T0; sCm = c.wC * Tm;
    }
  }

  if (HAS_ERR) {
    // Composite alpha_d weights for the error totals: rows 0 / c0 are also rows
    // 4 / c2 of the lane above (the tile above accounts for lane 0's share in its
    // halo pre-pass).
    const double a0 = fabs(r0.xh), a1 = fabs(r1.xh), a2 = fabs(r2.xh), a3 = fabs(r3.xh);
    const double am = fabs(rc.xh);
    double EW0 = D0, EC0 = DC;
    if (lane > 0) {
      EW0 += fd * __ldg(gwf + 2 * I - 1);
      EC0 += fd * __ldg(gwc + I - 1);
    }
    const double W1 = 4.0 * D0, W2 = D0 + D1, W3 = 4.0 * D1, WC1 = 4.0 * DC;
    double e1 = EW0 * fma(a0, tF0, bF0);
    e1 = fma(W1, fma(a1, tF1, bF1), e1);
    e1 = fma(W2, fma(a2, tF2, bF2), e1);
    e1 = fma(W3, fma(a3, tF3, bF3), e1);
    double e1c = EC0 * fma(a0, tC0, bC0);
    e1c = fma(WC1, fma(am, tCm, bCm), e1c);
    err1 += fma(2.0, e1, -e1c);
    if (HAS_M2) {
      double e2 = EW0 * fma(a0 * a0, tF0, fma(2.0 * a0, bF0, qF0));
      e2 = fma(W1, fma(a1 * a1, tF1, fma(2.0 * a1, bF1, qF1)), e2);
      e2 = fma(W2, fma(a2 * a2, tF2, fma(2.0 * a2, bF2, qF2)), e2);
      e2 = fma(W3, fma(a3 * a3, tF3, fma(2.0 * a3, bF3, qF3)), e2);
      double e2c = EC0 * fma(a0 * a0, tC0, fma(2.0 * a0, bC0, qC0));
      e2c = fma(WC1, fma(am * am, tCm, fma(2.0 * am, bCm, qCm)), e2c);
      err2 += fma(2.0, e2, -e2c);
    }
  }

}

template <int MODE, int CLS, bool HAS_ERR, bool HAS_M2, bool HAS_BOUND>
__global__ void __launch_bounds__(QB_FUSED_WARPS * 32, QB_FUSED_MIN_CTAS) k_fused2d(FusedArgs a) {
  extern __shared__ __align__(16) double s_dyn[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const FusedConst& k = a.k;
  const unsigned tile = k.tile_base + blockIdx.x * QB_FUSED_WARPS + warp;
  if (tile >= k.tile_end) return;
  double* s_cols = s_dyn + warp * (QB_TILE_RECS * QB_FUSED_REC + 2 * 32);
  double* s_halo = s_cols + QB_TILE_RECS * QB_FUSED_REC;  // [2][32]
  const int D = k.D, nb = k.nb;
  const unsigned per_slice = (unsigned)(nb * nb);
  const unsigned sidx = tile / per_slice;
  const unsigned rem = tile - sidx * per_slice;
  const int jc = (int)(rem / (unsigned)nb), ib = (int)(rem - (unsigned)jc * nb);
  const int I0 = ib * 32, J0 = jc * 32, I = I0 + lane;
  const FusedSlice s = a.slices[sidx];
  const AxisD* tdc = a.tab_a + (size_t)s.tab_a * k.NP;
  const AxisD* tdf = tdc + (2 * D + 1);
  const double* gwc = a.gw;
  const double* gwf = a.gw + D;

  // ---- stage the tile's 161 column records in shared memory (per warp) ---------
  // (cp.async: all 26 16-byte copies of a lane are in flight at once and overlap the row loads)
  {
    const double2* g =
        (const double2*)(a.cols + ((size_t)s.tab_b * k.ncol + (size_t)5 * J0) * QB_FUSED_REC);
    const unsigned sh = (unsigned)__cvta_generic_to_shared(s_cols);
#pragma unroll 1
    for (int i = lane; i < QB_TILE_RECS * QB_FUSED_REC / 2; i += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sh + 16u * (unsigned)i),
                   "l"(g + i)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // alpha_d rows of this lane: fine h = 4 I .. 4 I + 3 and the coarse mid point.
  const RowReg r0 = load_row(tdf + 4 * I);
  const RowReg r1 = load_row(tdf + 4 * I + 1);
  const RowReg r2 = load_row(tdf + 4 * I + 2);
  const RowReg r3 = load_row(tdf + 4 * I + 3);
  const RowReg rc = load_row(tdc + 2 * I + 1);
  const double fd = s.scale_a * k.r_m / 6.0;
  // Simpson weights along alpha_d: fine rows (D0, 4 D0, D0 + D1, 4 D1, D1), coarse DC (1, 4, 1)
  const double D0 = fd * __ldg(gwf + 2 * I), D1 = fd * __ldg(gwf + 2 * I + 1);
  const double DC = fd * __ldg(gwc + I);

  double err1 = 0.0, err2 = 0.0;  // Richardson-combined error moments of this lane
  double tp = 0.0;
  int ok = 1;
  const bool RIDGE = true;  // the pre-pass below always tests (six evaluations per lane)
  constexpr bool SAME = false;  // ... with the full-precision sum
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- halo row h = 4 (I0 + 32): lanes <-> columns pre-pass -------------------
  {
    const RowReg rh = load_row(tdf + 4 * (I0 + 32));
    const double* cp = s_cols + (size_t)(5 * lane) * QB_FUSED_REC;
    const ColRec c0 = load_col_s(cp);
    const ColRec c1 = load_col_s(cp + QB_FUSED_REC);
    const ColRec c2 = load_col_s(cp + 2 * QB_FUSED_REC);
    const ColRec c3 = load_col_s(cp + 3 * QB_FUSED_REC);
    const ColRec cm = load_col_s(cp + 4 * QB_FUSED_REC);
    const ColRec c4 = load_col_s(cp + 5 * QB_FUSED_REC);
    QB_EVAL1(rh, c0, T0)
    QB_EVAL1(rh, c1, T1)
    QB_EVAL1(rh, c2, T2)
    QB_EVAL1(rh, c3, T3)
    QB_EVAL1(rh, cm, Tm)
    QB_EVAL1(rh, c4, T4)
    const double HF = fma(c4.wF2, T4, fma(c3.wF, T3, fma(c2.wF, T2, fma(c1.wF, T1, c0.wF * T0))));
    const double HC = fma(c4.wC2, T4, fma(cm.wC, Tm, c0.wC * T0));
    s_halo[lane] = HF;
    s_halo[32 + lane] = HC;
    if (HAS_ERR) {
      // weights of this row as p = 4 / c2 of lane 31's cell
      const double wf4 = fd * __ldg(gwf + 2 * (I0 + 31) + 1);
      const double wc2 = fd * __ldg(gwc + I0 + 31);
      const double ah = fabs(rh.xh);
      const double HFB = fma(c4.wF2 * c4.b, T4,
                             fma(c3.wF * c3.b, T3,
                                 fma(c2.wF * c2.b, T2, fma(c1.wF * c1.b, T1, (c0.wF * c0.b) * T0))));
      const double HCB = fma(c4.wC2 * c4.b, T4, fma(cm.wC * cm.b, Tm, (c0.wC * c0.b) * T0));
      err1 = 2.0 * wf4 * fma(ah, HF, HFB) - wc2 * fma(ah, HC, HCB);
      if (HAS_M2) {
        const double HFBB =
            fma(c4.wF2 * c4.b * c4.b, T4,
                fma(c3.wF * c3.b * c3.b, T3,
                    fma(c2.wF * c2.b * c2.b, T2,
                        fma(c1.wF * c1.b * c1.b, T1, (c0.wF * c0.b * c0.b) * T0))));
        const double HCBB =
            fma(c4.wC2 * c4.b * c4.b, T4, fma(cm.wC * cm.b * cm.b, Tm, (c0.wC * c0.b * c0.b) * T0));
        err2 = 2.0 * wf4 * fma(ah * ah, HF, fma(2.0 * ah, HFB, HFBB)) -
               wc2 * fma(ah * ah, HC, fma(2.0 * ah, HCB, HCBB));
      }
    }
    if (HAS_BOUND) {
      const double ah = fabs(rh.xh);
      QB_BOUND_TEST(T0, ah, c0)
      QB_BOUND_TEST(Tm, ah, cm)
      QB_BOUND_TEST(T4, ah, c4)
    }
    __syncwarp();
  }

  double* outp = a.out + (size_t)s.slot * D * D + (size_t)J0 * D + I;
  // u = x_d + y is monotone in both indices: its range over the tile (rows I0 .. I0 + 32,
  // columns J0 .. J0 + 32) is spanned by the corner sums. Warp-uniform.
  bool ridge = false;
  if (CLS == 0) {
    const double xa = __shfl_sync(0xffffffffu, r0.xh, 0);
    const double xb = __ldg(&tdf[4 * (I0 + 32)].xh);
    const double ya = s_cols[0], yb = s_cols[(QB_TILE_RECS - 1) * QB_FUSED_REC];
    const double lo = fmin(xa, xb) + fmin(ya, yb), hi = fmax(xa, xb) + fmax(ya, yb);
    ridge = !(lo > 0.0626 || hi < -0.0626);
  }
  // x_d (rows) and y (columns) keep their signs over the tile: same sign <=> no cancellation in u
  const bool same = (__shfl_sync(0xffffffffu, r0.xh, 0) > 0.0) == (s_cols[0] > 0.0);
  if (CLS == 0 && ridge)
    fused_march<MODE, CLS, true, false, HAS_ERR, HAS_M2, HAS_BOUND>(
        k, r0, r1, r2, r3, rc, D0, D1, DC, fd, s_cols, s_halo, outp, D, I, lane, gwc, gwf, tp, err1,
        err2, ok);
  else if (CLS == 0 && same)  // (measured: not worth a second instantiation for classes 1 / 2)
    fused_march<MODE, CLS, false, true, HAS_ERR, HAS_M2, HAS_BOUND>(
        k, r0, r1, r2, r3, rc, D0, D1, DC, fd, s_cols, s_halo, outp, D, I, lane, gwc, gwf, tp, err1,
        err2, ok);
  else
    fused_march<MODE, CLS, false, false, HAS_ERR, HAS_M2, HAS_BOUND>(
        k, r0, r1, r2, r3, rc, D0, D1, DC, fd, s_cols, s_halo, outp, D, I, lane, gwc, gwf, tp, err1,
        err2, ok);

  // warp reduction in a fixed order
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    tp += __shfl_down_sync(0xffffffffu, tp, off);
    err1 += __shfl_down_sync(0xffffffffu, err1, off);
    err2 += __shfl_down_sync(0xffffffffu, err2, off);
    ok &= __shfl_down_sync(0xffffffffu, ok, off);
  }
  fused_close_tile(a.part, a.tickets, a.summary, (unsigned)s.slot, per_slice, rem, lane, tp, err1, err2, ok);
}

// One WARP per slice: tile partials -> summary, in tile order. The lanes fetch 32 partials at a
// time (one coalesced 1 KB read instead of 32 dependent ones: the kernel was 9 us of latency for a
// 3362-slice batch, a fixed cost that matters once a GPU holds an eighth of a distribution); lane 0
// adds them in tile order, so the sums are those of the serial loop, bit for bit.
#define QB_FINAL_WARPS 4
__global__ void __launch_bounds__(32 * QB_FINAL_WARPS)
k_fused_final(unsigned n, unsigned per_slice, const double* __restrict__ part,
              double* __restrict__ summary) {
  __shared__ double stage[QB_FINAL_WARPS][32][QB_FUSED_PART_STRIDE];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned s = blockIdx.x * QB_FINAL_WARPS + warp;
  if (s >= n) return;  // whole warp
  dd tp = make_dd(0.0, 0.0);
  double m1 = 0.0, m2 = 0.0;
  int ok = 1;
  for (unsigned t0 = 0; t0 < per_slice; t0 += 32) {
    const unsigned cnt = per_slice - t0 < 32u ? per_slice - t0 : 32u;
    if (lane < cnt) {
      const double* p = part + ((size_t)s * per_slice + t0 + lane) * QB_FUSED_PART_STRIDE;
#pragma unroll
      for (int q = 0; q < QB_FUSED_PART_STRIDE; q++) stage[warp][lane][q] = p[q];
    }
    __syncwarp();
    if (lane == 0) {
      for (unsigned t = 0; t < cnt; t++) {
        tp = dd_add_d(tp, stage[warp][t][0]);
        m1 += stage[warp][t][1];
        m2 += stage[warp][t][2];
        ok &= (stage[warp][t][3] != 0.0);
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    double* o = summary + (size_t)s * 8;
    o[0] = tp.hi;
    o[1] = tp.lo;
    o[2] = m1;
    o[3] = m2;
    o[4] = (double)ok;
    o[5] = o[6] = o[7] = 0.0;
  }
}

// ---- host side ----------------------------------------------------------------

// Decide whether the fused kernel applies and which variants; sort the slices by class.
inline bool fused2d_prepare(const Plan& h, unsigned n_chunks, FusedPlan2D* f, std::string* why) {
  if (h.kind >= 0) {
    *why = "not a two-dimensional plan";
    return false;
  }
  if (!h.richardson) {
    *why = "single-pass (non-Richardson) request";
    return false;
  }
  if (h.D % 32 != 0) {
    *why = "dimension is not a multiple of 32";
    return false;
  }
  if ((size_t)h.slices.size() * (size_t)(h.D / 32) * (h.D / 32) >= (size_t(1) << 31)) {
    *why = "too many tiles";
    return false;
  }
  const DevConsts& c = h.c;
  const double akappa = std::fabs(c.kappa.hi) * (1.0 + 1e-12) + 1e-300;
  int rel_max = -100000;
  for (size_t i = 0; i < h.slices.size(); i++) {
    rel_max = std::max(rel_max, (int)std::labs((long)h.k_a[i]) - c.m);
    rel_max = std::max(rel_max, (int)std::labs((long)h.k_b[i]) - c.m);
  }
  if (h.slices.empty()) rel_max = 0;
  // MODE and the second error moment are decided per PLAN; a generator never asks for a coordinate
  // above m + 10 (src/main_generate_distribution.cpp:1196-1212), so every plan inside that range
  // takes the decision of the whole range: a slice then has the same bits whichever batch it is
  // computed in -- alone (one call per slice) or with the whole enumerator list (the prefetching
  // drop-in). Plans that reach further out decide for themselves.
  rel_max = std::max(rel_max, 10);
  // |u| <= |x_d| + |kappa| |x_r| < 2^(rel_max + 1) * (1 + |kappa|)
  const int log_u = rel_max + 1 + (int)std::ceil(std::log2(1.0 + akappa));
  if (c.lam_exp - log_u >= 29) {
    f->mode = 0;
  } else if (c.lam_exp - log_u >= 10) {
    f->mode = 1;
  } else {
    *why = "l - sigma too small for the sinc expansion of the inner sine";
    return false;
  }
  // 1 / sinc^2(z) = 1 + z^2/3 + z^4/15 + 2 z^6/189, z^2 = (pi/Lambda)^2 w
  const double q = std::ldexp(9.86960440108935861883, -2 * std::min(c.lam_exp, 500));
  f->k.c1 = q / 3.0;
  f->k.c2 = q * q / 15.0;
  f->k.c3 = q * q * q * (2.0 / 189.0);
  f->k.cs = c.cs;
  f->k.e0s = c.e0s;
  f->k.r_m = c.r_m;
  f->k.D = h.D;
  f->k.nb = h.D / 32;
  f->k.NP = table_points(h.D);
  f->k.ncol = 5 * h.D + 1;
  const unsigned per_slice = (unsigned)(f->k.nb * f->k.nb);
  f->k.n_tiles = (unsigned)h.slices.size() * per_slice;
  f->k.tile_base = 0;
  f->k.tile_end = f->k.n_tiles;
  f->has_err = h.with_error;
  // second moment: relative size cs * h_max / 2 versus the first
  const double csh = c.cs * std::ldexp(1.0, rel_max + 2);
  f->has_m2 = h.with_error && (csh > 1e-13);
  // The bound test needs the per-point norm only while e0s can matter:
  // T1 T2 r/2^m is never below ~1e-90 at double precision, so for e0s < 2^-330
  // the test is s (2 + s) <= bound, monotone in h => decided at the far corner.
  f->has_bound = h.with_error && (c.e0s >= std::ldexp(1.0, -330));
  f->host_unbounded.assign(h.slices.size(), 0);
  if (h.with_error && !f->has_bound) {
    for (size_t i = 0; i < h.slices.size(); i++) {
      const double hmax = 2.0 * (h.slices[i].scale_a + h.slices[i].scale_b);
      const double sv = c.cs * hmax;
      f->host_unbounded[i] = (QB_ERROR_BOUND - sv * (2.0 + sv) >= 0.0) ? 0 : 1;
    }
  }
  // classes: an upper bound of |u| over every abscissa of the slice (the last
  // main points sit at 2 * 2^(k - m))
  const unsigned n = (unsigned)h.slices.size();
  n_chunks = std::max(1u, std::min(n_chunks, std::max(1u, n)));
  f->fslices.clear();
  f->chunks.clear();
  unsigned base = 0;
  for (unsigned ch = 0; ch < n_chunks; ch++) {
    FusedChunk fc;
    // equal chunks, except that the first one is a quarter of a share (and the second makes up
    // for it): the first device-to-host copy of the synchronous API starts that much earlier
    const auto edge = [&](unsigned c) -> unsigned {
      if (c == 0) return 0u;
      if (c >= n_chunks) return n;
      const uint64_t quarters = c == 1 && n_chunks > 2 ? 1 : 4 * (uint64_t)c;
      return (unsigned)((uint64_t)n * quarters / (4 * (uint64_t)n_chunks));
    };
    fc.slot_begin = edge(ch);
    fc.slot_end = edge(ch + 1);
    std::vector<FusedSlice> by_cls[3];
    for (unsigned i = fc.slot_begin; i < fc.slot_end; i++) {
      const double umax = 2.0 * (h.slices[i].scale_a + akappa * h.slices[i].scale_b);
      FusedSlice fs;
      fs.tab_a = h.slices[i].tab_a;
      fs.tab_b = h.slices[i].tab_b;
      fs.slot = (int)i;
      fs.scale_a = h.slices[i].scale_a;
      fs.cls = umax < 0.001953124 ? 2 : (umax < 0.062499 ? 1 : 0);
      by_cls[fs.cls].push_back(fs);
    }
    for (int cl = 0; cl < 3; cl++) {
      fc.class_tiles[cl] = base;
      f->fslices.insert(f->fslices.end(), by_cls[cl].begin(), by_cls[cl].end());
      base += (unsigned)by_cls[cl].size() * per_slice;
    }
    fc.class_tiles[3] = base;
    f->chunks.push_back(fc);
  }
  f->cols_bytes = std::max<size_t>(1, h.tabs_b.size()) * (size_t)f->k.ncol * QB_FUSED_REC *
                  sizeof(double);
  return true;
}

inline uint32_t fused2d_launches(const FusedPlan2D& f, bool lean = true) {
  uint32_t n = lean ? 1 : 3;  // lean: one prologue launch, summaries inside the class kernels;
                              // else axis tables, column records and the final summary
  for (const FusedChunk& c : f.chunks)
    for (int cl = 0; cl < 3; cl++) n += c.class_tiles[cl + 1] > c.class_tiles[cl] ? 1 : 0;
  return n;
}

template <int MODE, int CLS>
inline cudaError_t fused2d_launch_variant(const FusedPlan2D& f, const FusedArgs& args,
                                          unsigned blocks, cudaStream_t st) {
  const int v = (f.has_err ? 1 : 0) | (f.has_m2 ? 2 : 0) | (f.has_bound ? 4 : 0);
  const dim3 g(blocks), b(QB_FUSED_WARPS * 32);
  const size_t sm = QB_FUSED_SMEM_BYTES;
#define QB_LAUNCH(E, M2, B)                                                                   \
  {                                                                                           \
    auto kern = k_fused2d<MODE, CLS, E, M2, B>;                                               \
    /* per device, not per process: set before every launch (a host-side table write) */      \
    const cudaError_t ea = cudaFuncSetAttribute(                                              \
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                          \
    if (ea != cudaSuccess) return ea;                                                         \
    kern<<<g, b, sm, st>>>(args);                                                             \
  }
  switch (v) {
    case 0: QB_LAUNCH(false, false, false) break;
    case 1: QB_LAUNCH(true, false, false) break;
    case 3: QB_LAUNCH(true, true, false) break;
    case 5: QB_LAUNCH(true, false, true) break;
    default: QB_LAUNCH(true, true, true) break;
  }
#undef QB_LAUNCH
  return cudaGetLastError();
}

// Enqueue the fused kernel for one chunk (one launch per slice class present).
// streams: where each class goes (nullptr: all three on st, one after the other).
inline int fused2d_launch_chunk(const FusedPlan2D& f, FusedArgs args, size_t chunk,
                                cudaStream_t st_all, const cudaStream_t* streams = nullptr) {
  const FusedChunk& c = f.chunks[chunk];
  for (int cl = 0; cl < 3; cl++) {
    const unsigned lo = c.class_tiles[cl], hi = c.class_tiles[cl + 1];
    if (lo >= hi) continue;
    const cudaStream_t st = streams ? streams[cl] : st_all;
    args.k.tile_base = lo;
    args.k.tile_end = hi;
    const unsigned blocks = (hi - lo + QB_FUSED_WARPS - 1) / QB_FUSED_WARPS;
    cudaError_t e;
    if (f.mode == 0) {
      e = cl == 0 ? fused2d_launch_variant<0, 0>(f, args, blocks, st)
                  : cl == 1 ? fused2d_launch_variant<0, 1>(f, args, blocks, st)
                            : fused2d_launch_variant<0, 2>(f, args, blocks, st);
    } else {
      e = cl == 0 ? fused2d_launch_variant<1, 0>(f, args, blocks, st)
                  : cl == 1 ? fused2d_launch_variant<1, 1>(f, args, blocks, st)
                            : fused2d_launch_variant<1, 2>(f, args, blocks, st);
    }
    if (e != cudaSuccess) return -100;
  }
  return 0;
}

inline FusedArgs fused2d_args(const FusedPlan2D& f, const FusedSlice* d_fslices,
                              const double* d_cols, const AxisD* tab_a, const double* gw,
                              double* part, double* d_cells, unsigned int* tickets = nullptr,
                              double* d_summary = nullptr) {
  FusedArgs args;
  args.tickets = tickets;
  args.summary = d_summary;
  args.k = f.k;
  args.slices = d_fslices;
  args.tab_a = tab_a;
  args.cols = d_cols;
  args.gw = gw;
  args.out = d_cells;
  args.part = part;
  return args;
}

inline void fused2d_launch_cols(const FusedPlan2D& f, const Plan& h, cudaStream_t st,
                                const TabDesc* desc_b, const AxisR* tab_b, const double* gw,
                                double* d_cols) {
  k_fused_cols<<<dim3((f.k.ncol + 127) / 128, (unsigned)h.tabs_b.size()), 128, 0, st>>>(
      h.D, h.c.m, desc_b, tab_b, gw, d_cols);
}

inline void fused2d_launch_final(const FusedPlan2D& f, const Plan& h, cudaStream_t st,
                                 const double* part, double* d_summary) {
  const unsigned n = (unsigned)h.slices.size();
  k_fused_final<<<(n + QB_FINAL_WARPS - 1) / QB_FINAL_WARPS, 32 * QB_FINAL_WARPS, 0, st>>>(
      n, (unsigned)(f.k.nb * f.k.nb), part, d_summary);
}

}  // namespace qb200
