// slice_cells.cuh -- per-abscissa and per-cell work items of the slice integrators.
//
// These are the bodies of the "plain" kernels in kernels.cu (one thread per
// abscissa / per cell). They are __host__ __device__ so that tests/hostsim can
// run the identical code on the CPU as a unit test of the mathematics; the
// shipped library calls them from device code only.
//
// Layout of one axis table (per distinct signed slice coordinate k and
// dimension D): NP = 6 D + 2 abscissae, first the coarse pass (interleaved main
// / mean points, 2 D + 1 entries: src/distribution_slice_compute.cpp:199-246 at
// dimension D) and then the fine pass used by the Richardson extrapolation
// (4 D + 1 entries: the same loop at dimension 2 D,
// src/distribution_slice_compute_richardson.cpp:28-44).
#pragma once

#include "integrands.cuh"

namespace qb200 {

struct AxisD {  // alpha_d axis
  double xh, xl;  // x_d = alpha_d / 2^m (double-double, signed)
  double sd, cd;  // sin(pi x_d), cos(pi x_d)
};

struct AxisR {  // alpha_r axis
  double yh, yl;  // y = kappa * x_r (double-double)
  double sr, cr;  // sin(pi y), cos(pi y)
  double t2;      // second factor of the approximation at x_r (theta_r only)
  double b;       // |x_r|
};

struct TabDesc {
  int k_abs;  // |min_log_alpha|
  int sign;   // +1 / -1
};

QHD int table_points(int D) { return 6 * D + 2; }
QHD int pass_offset(int D, int fine) { return fine ? 2 * D + 1 : 0; }
QHD int width_offset(int D, int fine) { return fine ? D : 0; }

QHD void axis_d_point(const DevConsts& c, dd g, TabDesc t, AxisD* out) {
  const dd x = grid_x(g, t.k_abs, t.sign, c.m);
  out->xh = x.hi;
  out->xl = x.lo;
  sincospi_dd(x, &out->sd, &out->cd);
}

QHD void axis_r_point(const DevConsts& c, dd g, TabDesc t, AxisR* out) {
  const dd x = grid_x(g, t.k_abs, t.sign, c.m);
  const dd y = dd_mul(c.kappa, x);
  out->yh = y.hi;
  out->yl = y.lo;
  sincospi_dd(y, &out->sr, &out->cr);
  out->t2 = t2_value(x, c.c_over_L, c.l);
  out->b = fabs(x.hi);
}

// One Simpson cell of one pass (src/distribution_slice_compute.cpp:334-407).
//   td, tr : the pass's interleaved tables (2 Dp + 1 entries each)
//   wd, wr : cell widths (Delta alpha / 2^m), already scaled by 2^(|k| - m)
// Outputs the cell mass and the two moments of h = |x_d| + |x_r| from which the
// error sum is rebuilt on the host:
//   sum error = 2 cs * m1 + cs^2 * m2 + e0 term   (src/probability.cpp:252-277)
// and whether all nine points satisfy the relative bound (:280-281).
QHD void pass2d_cell(const DevConsts& c, const AxisD* td, const AxisR* tr, double wd,
                     double wr, int I, int J, bool with_error, double* mass, double* m1,
                     double* m2, bool* bounded) {
  const double w3[3] = {1.0, 4.0, 1.0};
  double acc = 0.0, a1 = 0.0, a2 = 0.0;
  bool ok = true;
  for (int b = 0; b < 3; b++) {
    const AxisR r = tr[2 * J + b];
    for (int a = 0; a < 3; a++) {
      const AxisD d = td[2 * I + a];
      const double t1 = t1_value(make_dd(d.xh, d.xl), make_dd(r.yh, r.yl), c.lam_exp);
      const double n = t1 * r.t2;
      const double w = w3[a] * w3[b];
      acc = fma(w, n, acc);
      if (with_error) {
        const double h = fabs(d.xh) + r.b;
        a1 = fma(w * h, n, a1);
        a2 = fma(w * h * h, n, a2);
        const double s = c.cs * h;
        // error / norm <= bound  <=>  e0s <= (bound - s (2 + s)) * n * r/2^m
        const double room = QB_ERROR_BOUND - s * (2.0 + s);
        ok = ok && (room >= 0.0) && (c.e0s <= room * n * c.r_m);
      }
    }
  }
  const double f = (wd * wr) * c.r_m / 36.0;
  *mass = acc * f;
  *m1 = a1 * f;
  *m2 = a2 * f;
  *bounded = ok;
}

// ---- one-dimensional distributions ----------------------------------------

enum Kind1D { KIND_LINEAR_D = 0, KIND_LINEAR_R = 1, KIND_DIAGONAL = 2 };

QHD double value_1d(const DevConsts& c, int kind, dd g, TabDesc t, double eta_shift) {
  const dd x = grid_x(g, t.k_abs, t.sign, c.m);
  if (kind == KIND_LINEAR_D) return linear_d_value(x, c.d_m, c.omd_m, c.l);
  if (kind == KIND_LINEAR_R) return linear_r_value(x, c);
  return diagonal_value(x, eta_shift, c.rho);
}

// Simpson cell I of one pass from the pass's interleaved values
// (src/linear_distribution_slice_compute.cpp:176-187,
//  src/diagonal_distribution_slice_compute.cpp:176-185).
QHD double pass1d_cell(const double* v, double w, int I) {
  return (fma(4.0, v[2 * I + 1], v[2 * I]) + v[2 * I + 2]) / 6.0 * w;
}

}  // namespace qb200
