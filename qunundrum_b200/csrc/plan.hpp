// plan.hpp -- host-side planning of a batch of slices (no CUDA in this file).
//
// A batch is a list of slice coordinates of one distribution at one dimension.
// Planning (a) validates the request the way the reference's entry points
// would, (b) turns (m, l, sigma, d, r) into DevConsts once, (c) deduplicates
// the axis tables: slices that share a signed coordinate share the O(D)
// per-abscissa work (the reference recomputes it inside every one of the
// (2 D + 1)^2 integrand calls), and (d) builds the 2^(i/D) geometry tables.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "hostconst.hpp"
#include "sigma_opt.cuh"
#include "slice_cells.cuh"

namespace qb200 {

// Reference flag bits, src/common.h:226-257.
enum : uint32_t {
  kFlagErrorBoundWarning = 0x00000001u,
  kFlagMethodSimpson = 0x00020000u,
  kFlagMethodRichardson = 0x00080000u,
};

// Distribution_Slice_Compute_Method, src/distribution_slice.h:31-78.
enum { kMethodHeuristicSigma = 0, kMethodOptimalLocalSigma = 1, kMethodQuick = 2 };

struct ParamsView {
  uint32_t m, l, sigma;
  const uint8_t* d_be;
  size_t d_len;
  const uint8_t* r_be;
  size_t r_len;
};

// sigma = round((l + tau + 4 - 1.6515) / 2), tau = 11, in float arithmetic
// exactly as src/distribution_slice_compute.cpp:149-158.
inline uint32_t heuristic_sigma(uint32_t l) {
  const uint32_t tau = 11;
  return (uint32_t)round(((float)l + tau + 4 - 1.6515f) / 2.0f);
}

inline dd to_dd(DD v) { return make_dd(v.hi, v.lo); }

inline int make_dev_consts(const ParamsView& p, int method_2d, DevConsts* c, std::string* err,
                           SigmaOptConsts* so = nullptr) {
  uint32_t sigma = p.sigma;
  if (method_2d == kMethodOptimalLocalSigma) {
    sigma = 1;  // per-point sigma; the constants below that depend on sigma are not used
    if (p.l < 4) {
      *err = "l is too small for the sigma-optimal method";
      return -3;
    }
  }
  if (method_2d == kMethodHeuristicSigma) {
    sigma = heuristic_sigma(p.l);
    if (sigma > p.l) {
      // the reference forms 2^(l - sigma) with an unsigned difference (src/probability.cpp:178)
      // and returns NaN cells here; refuse instead
      *err = "heuristic sigma > l: l is too small for the error-bounded approximation";
      return -3;
    }
  }
  HostConsts h;
  const int rc = host_consts_compute(p.m, p.l, sigma, p.d_be, p.d_len, p.r_be, p.r_len, &h);
  if (rc != 0) {
    *err = rc == -1 ? "d and r must be non-zero"
                    : rc == -2 ? "d and r must be below 2^m" : "m, l or sigma out of range";
    return rc;
  }
  c->kappa = to_dd(method_2d == kMethodQuick ? h.kappa_q : h.kappa);
  c->c_over_L = to_dd(h.c_over_L);
  c->n_over_L = to_dd(h.n_over_L);
  c->n1_over_L = to_dd(h.n1_over_L);
  c->rho = to_dd(h.rho);
  c->r_m = h.r_m.hi;
  c->d_m = h.d_m.hi;
  c->omd_m = (1.0 - h.d_m.hi) - h.d_m.lo;
  c->beta_m = h.beta_m.hi;
  c->rbeta_m = h.rbeta_m.hi;
  c->m = (int)p.m;
  c->l = (int)p.l;
  c->sigma = (int)sigma;
  c->lam_exp = method_2d == kMethodQuick ? (int)p.l : (int)p.l - (int)sigma;
  c->cs = std::ldexp(3.14159265358979323846, (int)sigma - (int)p.l);
  c->e0s = std::ldexp(1.0, 4 - (int)sigma) + std::ldexp(1.0, 3 - (int)p.l);
  if (so) {
    for (int i = 0; i < 3; i++) so->q_mant[i] = h.q_mant[i];
    so->q_exp = h.q_exp;
  }
  return 0;
}

// Geometry of one dimension D: abscissa factors g in [1, 2] (double-double)
// for the coarse pass (2 D + 1 interleaved points) followed by the fine pass
// (4 D + 1), and the cell widths 2^((i+1)/D) - 2^(i/D) (D coarse, then 2 D fine)
// (src/distribution_slice_compute.cpp:196-212 and 331-347).
struct Geometry {
  int D = 0;
  std::vector<DD> gx;      // 6 D + 2
  std::vector<double> gw;  // 3 D
};

inline void fill_pass(int Dp, DD* gx, double* gw) {
  std::vector<DD> t((size_t)Dp + 1);
  exp2_table_dd((uint32_t)Dp, t.data());
  for (int i = 0; i <= Dp; i++) gx[2 * i] = t[i];
  for (int i = 0; i < Dp; i++) {
    const dd a = make_dd(t[i].hi, t[i].lo), b = make_dd(t[i + 1].hi, t[i + 1].lo);
    const dd mean = dd_mul_pow2(dd_add(a, b), 0.5);
    gx[2 * i + 1].hi = mean.hi;
    gx[2 * i + 1].lo = mean.lo;
    const dd w = dd_add(b, dd_neg(a));
    gw[i] = w.hi + w.lo;
  }
}

inline Geometry make_geometry(int D) {
  Geometry g;
  g.D = D;
  g.gx.resize((size_t)table_points(D));
  g.gw.resize((size_t)3 * D);
  fill_pass(D, g.gx.data(), g.gw.data());
  fill_pass(2 * D, g.gx.data() + pass_offset(D, 1), g.gw.data() + width_offset(D, 1));
  return g;
}

struct SliceDesc {
  int tab_a;      // 2D: alpha_d table, 1D: the table
  int tab_b;      // 2D: alpha_r table
  double scale_a; // 2^(|k_a| - m)
  double scale_b; // 2^(|k_b| - m)   (1 for 1D)
  double eta_shift;  // diagonal: eta * 2^sigma
};

struct Plan {
  DevConsts c;
  int D = 0;
  int richardson = 1;
  int kind = -1;       // -1: two-dimensional; else Kind1D
  int method = 0;      // 2D only
  bool with_error = false;  // 2D error-bounded approximation
  SigmaOptConsts so;        // method == kMethodOptimalLocalSigma only
  std::vector<TabDesc> tabs_a, tabs_b;
  std::vector<SliceDesc> slices;
  std::vector<int> k_a, k_b;  // signed coordinates as given
};

inline int coord_ok(int32_t k, uint32_t m, std::string* err) {
  const long ka = std::labs((long)k);
  const long rel = ka - (long)m;
  if (rel < -400 || rel > 59) {
    *err = "slice coordinate outside the supported range m - 400 <= |min_log_alpha| <= m + 59";
    return 0;
  }
  return 1;
}

// Distinct signed coordinates -> table index. Coordinates are within m - 400 .. m + 59
// (coord_ok), so the index is a flat array: planning a 3362-slice batch is on the critical path
// of the synchronous API (the GPU idles until the plan is uploaded).
struct TableIndex {
  int m;
  std::vector<int> slot;  // [(|k| - m + 400) * 2 + (k >= 0)] -> table, -1 = none yet
  explicit TableIndex(uint32_t m_) : m((int)m_), slot(2 * 460, -1) {}
};

inline int intern_table(TableIndex& index, std::vector<TabDesc>& tabs, int32_t k) {
  const int sign = k < 0 ? -1 : 1;  // sgn_d(): zero counts as positive (src/math.cpp)
  const int ka = (int)std::labs((long)k);
  int& at = index.slot[(size_t)(ka - index.m + 400) * 2 + (sign > 0 ? 1 : 0)];
  if (at >= 0) return at;
  TabDesc t;
  t.k_abs = ka;
  t.sign = sign;
  tabs.push_back(t);
  at = (int)tabs.size() - 1;
  return at;
}

inline int plan_2d(const ParamsView& p, int method, int richardson, uint32_t D, uint32_t n,
                   const int32_t* a_d, const int32_t* a_r, Plan* plan, std::string* err) {
  if (method != kMethodHeuristicSigma && method != kMethodQuick &&
      method != kMethodOptimalLocalSigma) {
    *err = "unknown method specified for computing the slice";
    return -11;
  }
  if (D == 0 || D > (1u << 20)) {
    *err = "bad slice dimension";
    return -12;
  }
  const int rc = make_dev_consts(p, method, &plan->c, err, &plan->so);
  if (rc != 0) return rc;
  plan->D = (int)D;
  plan->richardson = richardson ? 1 : 0;
  plan->kind = -1;
  plan->method = method;
  plan->with_error = (method != kMethodQuick);
  TableIndex ia(p.m), ib(p.m);
  plan->slices.resize(n);
  plan->k_a.assign(a_d, a_d + n);
  plan->k_b.assign(a_r, a_r + n);
  for (uint32_t i = 0; i < n; i++) {
    if (!coord_ok(a_d[i], p.m, err) || !coord_ok(a_r[i], p.m, err)) return -13;
    SliceDesc& s = plan->slices[i];
    s.tab_a = intern_table(ia, plan->tabs_a, a_d[i]);
    s.tab_b = intern_table(ib, plan->tabs_b, a_r[i]);
    s.scale_a = std::ldexp(1.0, (int)std::labs((long)a_d[i]) - (int)p.m);
    s.scale_b = std::ldexp(1.0, (int)std::labs((long)a_r[i]) - (int)p.m);
    s.eta_shift = 0.0;
  }
  return 0;
}

// kind: KIND_LINEAR_D / KIND_LINEAR_R (eta ignored) or KIND_DIAGONAL.
inline int plan_1d(const ParamsView& p, int kind, int richardson, uint32_t D, uint32_t n,
                   const int32_t* a, const int32_t* eta, Plan* plan, std::string* err) {
  if (kind != KIND_LINEAR_D && kind != KIND_LINEAR_R && kind != KIND_DIAGONAL) {
    *err = "unknown target";
    return -11;
  }
  if (D == 0 || D > (1u << 24)) {
    *err = "bad slice dimension";
    return -12;
  }
  ParamsView q = p;
  if (kind != KIND_DIAGONAL) q.sigma = 0;
  if (kind == KIND_DIAGONAL && p.sigma > 900) {
    *err = "sigma too large";
    return -3;
  }
  const int rc = make_dev_consts(q, /*method_2d=*/-1, &plan->c, err);
  if (rc != 0) return rc;
  plan->D = (int)D;
  plan->richardson = richardson ? 1 : 0;
  plan->kind = kind;
  plan->with_error = false;
  TableIndex ia(p.m);
  plan->slices.resize(n);
  plan->k_a.assign(a, a + n);
  plan->k_b.assign(n, 0);
  for (uint32_t i = 0; i < n; i++) {
    if (!coord_ok(a[i], p.m, err)) return -13;
    SliceDesc& s = plan->slices[i];
    s.tab_a = intern_table(ia, plan->tabs_a, a[i]);
    s.tab_b = 0;
    s.scale_a = std::ldexp(1.0, (int)std::labs((long)a[i]) - (int)p.m);
    s.scale_b = 1.0;
    s.eta_shift = 0.0;
    if (kind == KIND_DIAGONAL) {
      const int32_t e = eta ? eta[i] : 0;
      if (std::labs((long)e) > (1l << 20)) {
        *err = "eta out of range";
        return -14;
      }
      s.eta_shift = std::ldexp((double)e, (int)p.sigma);
      plan->k_b[i] = e;
    }
  }
  return 0;
}

// total_error of a 2D slice from the device moments (see pass2d_cell):
//   sum over cells of Simpson(error) * widths / 2^m, error as in
//   src/probability.cpp:252-277, Richardson-combined as
//   src/distribution_slice_compute_richardson.cpp:66.
// sigma-optimal: every point has its own sigma; the device returns, per pass,
//   A = sum w pi h n r/2^m (2 + s) 2^(sigma_p - sigma_0),  C = sum w 2^(sigma_0 - sigma_p)
// (weights include the cell widths), and sigma_0 of the pass (kernels_sigma_opt.cuh).
inline long double total_error_sigma_opt(const Plan& plan, size_t i, const double* s) {
  const DevConsts& c = plan.c;
  const int ka = (int)std::labs((long)plan.k_a[i]) - c.m;
  const int kb = (int)std::labs((long)plan.k_b[i]) - c.m;
  const int packed = (int)s[7];
  const int s0c = packed % 65536, s0f = packed / 65536;
  const long double konst = ldexpl(1.0L, 3 - c.l + ka + kb);
  const long double ec =
      ldexpl((long double)s[2], s0c - c.l) + ldexpl((long double)s[3], 4 - s0c) + konst;
  if (!plan.richardson) return ec;
  const long double ef =
      ldexpl((long double)s[5], s0f - c.l) + ldexpl((long double)s[6], 4 - s0f) + konst;
  return 2.0L * ef - ec;
}

inline long double total_error_2d(const Plan& plan, size_t i, double m1, double m2) {
  if (!plan.with_error) return 0.0L;
  const DevConsts& c = plan.c;
  const long double pi = 3.14159265358979323846264338327950288L;
  const int sl = c.sigma - c.l;
  long double te = ldexpl(2.0L * pi * (long double)m1, sl) +
                   ldexpl(pi * pi * (long double)m2, 2 * sl);
  const int ka = (int)std::labs((long)plan.k_a[i]) - c.m;
  const int kb = (int)std::labs((long)plan.k_b[i]) - c.m;
  te += ldexpl(1.0L, 4 - c.sigma + ka + kb) + ldexpl(1.0L, 3 - c.l + ka + kb);
  return te;
}

}  // namespace qb200
