// tests/hostsim/hostsim.cpp -- TEST-ONLY CPU twin of the device functions.
//
// Compiles the __host__ __device__ bodies of qunundrum_b200/csrc (qmath.cuh,
// integrands.cuh, slice_cells.cuh) and the host planner (plan.hpp) with g++ and
// drives them with plain loops, so the mathematics of the kernels can be
// checked against the oracle on a machine without a GPU (`pytest -m "not gpu"`).
// It is NOT part of the product: libqunundrum_b200.so contains no such loops,
// no CPU path, and nothing under qunundrum_b200/ references this file.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../qunundrum_b200/csrc/plan.hpp"
#include "../../qunundrum_b200/csrc/client_math.cuh"
#include "../../qunundrum_b200/csrc/sampler.cuh"

using namespace qb200;

static std::string g_err;

static ParamsView view(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d, size_t dn,
                       const uint8_t* r, size_t rn) {
  ParamsView p;
  p.m = m;
  p.l = l;
  p.sigma = sigma;
  p.d_be = d;
  p.d_len = dn;
  p.r_be = r;
  p.r_len = rn;
  return p;
}

extern "C" {

const char* hostsim_last_error() { return g_err.c_str(); }

int hostsim_host_consts(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d, size_t dn,
                        const uint8_t* r, size_t rn, double* out20) {
  HostConsts h;
  const int rc = host_consts_compute(m, l, sigma, d, dn, r, rn, &h);
  if (rc) return rc;
  const DD* v[10] = {&h.kappa, &h.kappa_q, &h.c_over_L, &h.n_over_L, &h.n1_over_L,
                     &h.beta_m, &h.rbeta_m, &h.r_m, &h.d_m, &h.rho};
  for (int i = 0; i < 10; i++) {
    out20[2 * i] = v[i]->hi;
    out20[2 * i + 1] = v[i]->lo;
  }
  return 0;
}

void hostsim_exp2_table(uint32_t n, double* out_hi_lo) {
  std::vector<DD> t((size_t)n + 1);
  exp2_table_dd(n, t.data());
  for (uint32_t i = 0; i <= n; i++) {
    out_hi_lo[2 * i] = t[i].hi;
    out_hi_lo[2 * i + 1] = t[i].lo;
  }
}

uint32_t hostsim_heuristic_sigma(uint32_t l) { return heuristic_sigma(l); }

int hostsim_slice2d(uint32_t m, uint32_t l, const uint8_t* d, size_t dn, const uint8_t* r,
                    size_t rn, int method, int richardson, uint32_t D, uint32_t n,
                    const int32_t* a_d, const int32_t* a_r, double* cells,
                    long double* total_probability, long double* total_error,
                    uint32_t* flags) {
  Plan plan;
  const int rc = plan_2d(view(m, l, 0, d, dn, r, rn), method, richardson, D, n, a_d, a_r,
                         &plan, &g_err);
  if (rc) return rc;
  const Geometry geo = make_geometry((int)D);
  const int NP = table_points((int)D);
  std::vector<AxisD> ta(plan.tabs_a.size() * (size_t)NP);
  std::vector<AxisR> tb(plan.tabs_b.size() * (size_t)NP);
  for (size_t t = 0; t < plan.tabs_a.size(); t++)
    for (int i = 0; i < NP; i++)
      axis_d_point(plan.c, make_dd(geo.gx[i].hi, geo.gx[i].lo), plan.tabs_a[t],
                   &ta[t * NP + i]);
  for (size_t t = 0; t < plan.tabs_b.size(); t++)
    for (int i = 0; i < NP; i++)
      axis_r_point(plan.c, make_dd(geo.gx[i].hi, geo.gx[i].lo), plan.tabs_b[t],
                   &tb[t * NP + i]);
  const int Dc = (int)D;
  std::vector<double> fine((size_t)4 * Dc * Dc);
  if (method == kMethodOptimalLocalSigma) {
    // Serial walk of the reference (src/distribution_slice_compute.cpp:265-287) with the same
    // per-point functions the kernels use (sigma_opt.cuh).
    for (uint32_t s = 0; s < n; s++) {
      const SliceDesc& sd = plan.slices[s];
      double* out = cells + (size_t)s * Dc * Dc;
      double summ[8] = {0, 0, 0, 0, 1, 0, 0, 0};
      int s0[2] = {0, 0};
      bool bounded = true;
      for (int pass = 0; pass <= plan.richardson; pass++) {
        const int Dp = pass ? 2 * Dc : Dc, side = 2 * Dp + 1, off = pass_offset(Dc, pass);
        std::vector<double> nrm((size_t)side * side), era((size_t)side * side);
        std::vector<int> sg((size_t)side * side);
        std::vector<char> okp((size_t)side * side);
        int sigma = 0;
        for (int i = 0; i < side; i++)
          for (int j = 0; j < side; j++) {
            SoPoint pt;
            pt.xd_ = grid_x(make_dd(geo.gx[off + i].hi, geo.gx[off + i].lo),
                            plan.tabs_a[sd.tab_a].k_abs, plan.tabs_a[sd.tab_a].sign, plan.c.m);
            pt.xr_ = grid_x(make_dd(geo.gx[off + j].hi, geo.gx[off + j].lo),
                            plan.tabs_b[sd.tab_b].k_abs, plan.tabs_b[sd.tab_b].sign, plan.c.m);
            pt.t2 = tb[(size_t)sd.tab_b * NP + off + j].t2;
            pt.h = fabs(pt.xd_.hi) + fabs(pt.xr_.hi);
            double nn;
            xd ee;
            if (i == 0 && j == 0) {
              xd best = xd_make(1.0, plan.c.m);
              for (int t = 1; t < plan.c.l - 1; t++) {
                so_eval(plan.c, plan.so, pt, t, &nn, &ee);
                if (xd_less(ee, best)) {
                  best = ee;
                  sigma = t;
                }
              }
              if (sigma == 0) {
                g_err = "no admissible sigma";
                return -31;
              }
              s0[pass] = sigma;
              so_eval(plan.c, plan.so, pt, sigma, &nn, &ee);
            } else {
              bool inc;
              sigma = so_adjust(plan.c, plan.so, pt, sigma, &nn, &ee, &inc);
            }
            const size_t p = (size_t)i * side + j;
            const int sl = sigma - plan.c.l;
            const double ph = 3.14159265358979323846 * pt.h;
            const double sv = sl > -1000 ? ldexp(ph, sl) : 0.0;
            nrm[p] = nn;
            era[p] = ldexp(ph * (2.0 + sv) * nn * plan.c.r_m, sigma - s0[pass]);
            sg[p] = sigma;
            okp[p] = so_bounded(plan.c, nn, ee);
          }
        const double* gw = geo.gw.data() + width_offset(Dc, pass);
        double* dst = pass ? fine.data() : out;
        double A = 0, Cc = 0;
        const double w3[3] = {1.0, 4.0, 1.0};
        for (int J = 0; J < Dp; J++)
          for (int I = 0; I < Dp; I++) {
            double acc = 0, a_ = 0, c_ = 0;
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++) {
                const size_t p = (size_t)(2 * I + a) * side + (2 * J + b);
                acc += w3[a] * w3[b] * nrm[p];
                a_ += w3[a] * w3[b] * era[p];
                c_ += w3[a] * w3[b] * ldexp(1.0, s0[pass] - sg[p]);
                if (pass == 0) bounded = bounded && okp[p];
              }
            const double f = (gw[I] * sd.scale_a) * (gw[J] * sd.scale_b) / 36.0;
            dst[I + (size_t)Dp * J] = acc * f * plan.c.r_m;
            A += a_ * f;
            Cc += c_ * f;
          }
        summ[pass ? 5 : 2] = A;
        summ[pass ? 6 : 3] = Cc;
      }
      summ[7] = (double)(s0[0] + 65536 * s0[1]);
      long double tp = 0;
      if (plan.richardson) {
        for (int i = 0; i < Dc; i++)
          for (int j = 0; j < Dc; j++) {
            const size_t F = (size_t)2 * Dc;
            const double f = fine[F * (2 * j) + 2 * i] + fine[F * (2 * j) + 2 * i + 1] +
                             fine[F * (2 * j + 1) + 2 * i] + fine[F * (2 * j + 1) + 2 * i + 1];
            out[(size_t)Dc * j + i] = 2.0 * f - out[(size_t)Dc * j + i];
            tp += out[(size_t)Dc * j + i];
          }
      } else {
        for (size_t i = 0; i < (size_t)Dc * Dc; i++) tp += out[i];
      }
      total_probability[s] = tp;
      total_error[s] = total_error_sigma_opt(plan, s, summ);
      flags[s] = kFlagMethodSimpson | (plan.richardson ? kFlagMethodRichardson : 0u) |
                 (!bounded ? kFlagErrorBoundWarning : 0u);
    }
    return 0;
  }
  for (uint32_t s = 0; s < n; s++) {
    const SliceDesc& sd = plan.slices[s];
    double* out = cells + (size_t)s * Dc * Dc;
    double M1[2] = {0, 0}, M2[2] = {0, 0};
    bool bounded = true;
    for (int pass = 0; pass <= plan.richardson; pass++) {
      const int Dp = pass ? 2 * Dc : Dc;
      const AxisD* td = &ta[(size_t)sd.tab_a * NP + pass_offset(Dc, pass)];
      const AxisR* tr = &tb[(size_t)sd.tab_b * NP + pass_offset(Dc, pass)];
      const double* gw = geo.gw.data() + width_offset(Dc, pass);
      double* dst = pass ? fine.data() : out;
      for (int J = 0; J < Dp; J++)
        for (int I = 0; I < Dp; I++) {
          double mass, m1, m2;
          bool ok;
          pass2d_cell(plan.c, td, tr, gw[I] * sd.scale_a, gw[J] * sd.scale_b, I, J,
                      plan.with_error, &mass, &m1, &m2, &ok);
          dst[I + (size_t)Dp * J] = mass;
          M1[pass] += m1;
          M2[pass] += m2;
          if (pass == 0) bounded = bounded && ok;
        }
    }
    long double tp = 0;
    if (plan.richardson) {
      for (int i = 0; i < Dc; i++)
        for (int j = 0; j < Dc; j++) {
          const size_t F = (size_t)2 * Dc;
          const double f = fine[F * (2 * j) + 2 * i] + fine[F * (2 * j) + 2 * i + 1] +
                           fine[F * (2 * j + 1) + 2 * i] + fine[F * (2 * j + 1) + 2 * i + 1];
          out[(size_t)Dc * j + i] = 2.0 * f - out[(size_t)Dc * j + i];
          tp += out[(size_t)Dc * j + i];
        }
      total_error[s] =
          total_error_2d(plan, s, 2.0 * M1[1] - M1[0], 2.0 * M2[1] - M2[0]);
    } else {
      for (size_t i = 0; i < (size_t)Dc * Dc; i++) tp += out[i];
      total_error[s] = total_error_2d(plan, s, M1[0], M2[0]);
    }
    total_probability[s] = tp;
    flags[s] = kFlagMethodSimpson | (plan.richardson ? kFlagMethodRichardson : 0u) |
               ((plan.with_error && !bounded) ? kFlagErrorBoundWarning : 0u);
  }
  return 0;
}

// The sigma-optimal method's error sums, bound flag and sigma_0 by the closed-form walk of
// sigma_opt.cuh (what k_so_fast computes; the norm here by the accurate quick-method evaluation):
// total_error and flags of n slices, to compare with the serial walk of hostsim_slice2d(method 1).
// Returns 1 when a point left the range in which the closed form is proven.
int hostsim_so_fast(uint32_t m, uint32_t l, const uint8_t* d, size_t dn, const uint8_t* r, size_t rn,
                    int richardson, uint32_t D, uint32_t n, const int32_t* a_d, const int32_t* a_r,
                    long double* total_error, uint32_t* flags, int32_t* sigma0_out) {
  Plan quick, opt;
  int rc = plan_2d(view(m, l, 0, d, dn, r, rn), kMethodQuick, richardson, D, n, a_d, a_r, &quick, &g_err);
  if (rc) return rc;
  rc = plan_2d(view(m, l, 0, d, dn, r, rn), kMethodOptimalLocalSigma, richardson, D, n, a_d, a_r, &opt, &g_err);
  if (rc) return rc;
  const Geometry geo = make_geometry((int)D);
  const DevConsts& c = quick.c;
  int fallback = 0;
  for (uint32_t s = 0; s < n; s++) {
    const SliceDesc& sd = quick.slices[s];
    double summ[8] = {0, 0, 0, 0, 1, 0, 0, 0};
    int s0[2] = {0, 0};
    bool bounded = true;
    for (int pass = 0; pass <= richardson; pass++) {
      const int Dp = pass ? 2 * (int)D : (int)D, side = 2 * Dp + 1, off = pass_offset((int)D, pass);
      const double* wd = geo.gw.data() + width_offset((int)D, pass);
      int run = 0x3fffffff;
      double A = 0, Cc = 0;
      for (int i = 0; i < side; i++)
        for (int j = 0; j < side; j++) {
          const dd xd_ = grid_x(make_dd(geo.gx[off + i].hi, geo.gx[off + i].lo), quick.tabs_a[sd.tab_a].k_abs,
                                quick.tabs_a[sd.tab_a].sign, c.m);
          const dd xr_ = grid_x(make_dd(geo.gx[off + j].hi, geo.gx[off + j].lo), quick.tabs_b[sd.tab_b].k_abs,
                                quick.tabs_b[sd.tab_b].sign, c.m);
          const double nn = t1_value(xd_, dd_mul(c.kappa, xr_), c.lam_exp) * t2_value(xr_, c.c_over_L, c.l);
          const double ph = 3.14159265358979323846 * (fabs(xd_.hi) + fabs(xr_.hi));
          const double a = ph * nn * c.r_m;
          const int star = (i == 0 && j == 0) ? so_fast_sigma_first(c.l, a) : so_fast_sigma_star(c.l, a);
          run = star < run ? star : run;
          if (i == 0 && j == 0) s0[pass] = run;
          if (run < 64 || run > c.l - 60) {
            fallback = 1;
            continue;
          }
          const double wgt = so_axis_weight(wd, Dp, i) * so_axis_weight(wd, Dp, j);
          double era, ct;
          so_fast_terms(c, ph, nn, run, s0[pass], &era, &ct);
          A += wgt * era;
          Cc += wgt * ct;
          const bool in_bound = so_fast_bounded(c, ph, nn, run);
          if (in_bound != so_bounded(c, nn, so_error_given_norm(c, ph, nn, run))) {
            g_err = "so_fast_bounded disagrees with so_bounded";
            return -77;
          }
          if (pass == 0 && !in_bound) bounded = false;
        }
      const double f = sd.scale_a * sd.scale_b / 36.0;
      summ[pass ? 5 : 2] = A * f;
      summ[pass ? 6 : 3] = Cc * f;
    }
    summ[7] = (double)(s0[0] + 65536 * s0[1]);
    total_error[s] = total_error_sigma_opt(opt, s, summ);
    flags[s] = kFlagMethodSimpson | (richardson ? kFlagMethodRichardson : 0u) |
               (!bounded ? kFlagErrorBoundWarning : 0u);
    sigma0_out[2 * s] = s0[0];
    sigma0_out[2 * s + 1] = s0[1];
  }
  return fallback;
}

int hostsim_slice1d(uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d, size_t dn,
                    const uint8_t* r, size_t rn, int kind, int richardson, uint32_t D,
                    uint32_t n, const int32_t* a, const int32_t* eta, double* cells,
                    long double* total_probability, uint32_t* flags) {
  Plan plan;
  const int rc = plan_1d(view(m, l, sigma, d, dn, r, rn), kind, richardson, D, n, a, eta,
                         &plan, &g_err);
  if (rc) return rc;
  const Geometry geo = make_geometry((int)D);
  const int NP = table_points((int)D);
  const int Dc = (int)D;
  std::vector<double> v((size_t)NP);
  for (uint32_t s = 0; s < n; s++) {
    const SliceDesc& sd = plan.slices[s];
    for (int i = 0; i < NP; i++)
      v[i] = value_1d(plan.c, kind, make_dd(geo.gx[i].hi, geo.gx[i].lo),
                      plan.tabs_a[sd.tab_a], sd.eta_shift);
    double* out = cells + (size_t)s * Dc;
    long double tp = 0;
    for (int I = 0; I < Dc; I++) {
      const double coarse = pass1d_cell(v.data(), geo.gw[I] * sd.scale_a, I);
      double res = coarse;
      if (plan.richardson) {
        const double* vf = v.data() + pass_offset(Dc, 1);
        const double* wf = geo.gw.data() + width_offset(Dc, 1);
        const double f = pass1d_cell(vf, wf[2 * I] * sd.scale_a, 2 * I) +
                         pass1d_cell(vf, wf[2 * I + 1] * sd.scale_a, 2 * I + 1);
        res = 2.0 * f - coarse;
      }
      out[I] = res;
      tp += res;
    }
    total_probability[s] = tp;
    flags[s] = kFlagMethodSimpson | (plan.richardson ? kFlagMethodRichardson : 0u);
  }
  return 0;
}

}  // extern "C"

// ---- text formatter (textfmt.cuh) -----------------------------------------------
#include "../../qunundrum_b200/csrc/text_tables.hpp"

namespace {
std::vector<text::Pow10Entry>& pow10_table() {
  static std::vector<text::Pow10Entry> t;
  if (t.empty()) text::build_pow10_table(t);
  return t;
}
}  // namespace

extern "C" {

int hostsim_pow10_entry(int k, uint32_t* w6, int32_t* e2, uint32_t* exact) {
  if (k < text::K_MIN || k > text::K_MAX) return -1;
  const text::Pow10Entry& e = pow10_table()[(size_t)(k - text::K_MIN)];
  for (int i = 0; i < 6; i++) w6[i] = e.w[i];
  *e2 = e.e2;
  *exact = e.exact;
  return 0;
}

int32_t hostsim_floor_log10_pow2(int32_t n) { return text::floor_log10_pow2(n); }

// The loop a slice exporter runs: "%.24Lg\n" per value. force_band = 1 sends every
// value through the exact rounding decision. Returns the text length; *n_exact
// counts the values that took the exact path.
size_t hostsim_text_format_ld(const long double* v, size_t n, char* out, int force_band,
                              uint64_t* n_exact) {
  const text::Pow10Entry* tab = pow10_table().data();
  std::vector<uint32_t> scratch(text::BIG_LIMBS);
  size_t pos = 0;
  uint64_t slow = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t mant;
    uint16_t se;
    memcpy(&mant, (const char*)&v[i], 8);
    memcpy(&se, (const char*)&v[i] + 8, 2);
    text::Piece p;
    uint64_t Mn;
    int qn;
    if (text::classify_x87(mant, se, &p, &Mn, &qn)) {
      text::Dec24 d;
      text::digits24(Mn, qn, tab, &d, force_band != 0);
      bool up = d.up != 0;
      if (d.undecided) {
        up = text::exact_round_up(Mn, qn, d.x, d.c0, d.c1, d.c2, scratch.data());
        slow++;
      }
      text::round_digits(&d, up);
      text::piece_from_digits(d, &p);
    }
    const int len = text::piece_length(p);
    text::piece_render(p, out + pos);
    pos += (size_t)len;
  }
  if (n_exact) *n_exact = slow;
  return pos;
}

}  // extern "C"

// ---- text parser (textparse.cuh) --------------------------------------------------
#include "../../qunundrum_b200/csrc/textparse.cuh"

extern "C" {

// The importer's loop: the first n white-space separated numbers of text. Returns 0,
// or a negative code as qb200_text_parse_ld does; *consumed as documented there.
int hostsim_text_parse_ld(const char* textp, size_t len, size_t n, long double* values,
                          size_t* consumed, int force_band, uint64_t* n_exact) {
  const text::Pow10Entry* tab = pow10_table().data();
  const unsigned char* s = (const unsigned char*)textp;
  std::vector<uint32_t> scratch(text::BIG_LIMBS);
  size_t pos = 0, found = 0;
  uint64_t slow = 0;
  {  // as the CUDA path: "fewer than n numbers" takes precedence over a malformed number
    size_t cnt = 0;
    bool in = false;
    for (size_t i = 0; i < len && cnt < n; i++) {
      const bool sp = text::is_space(s[i]);
      if (!sp && !in) cnt++;
      in = !sp;
    }
    if (cnt < n) return -20;
  }
  while (found < n) {
    while (pos < len && text::is_space(s[pos])) pos++;
    if (pos >= len) return -20;
    text::Decimal dec;
    int tlen = 0;
    const uint32_t st = text::parse_number(s + pos, (long)(len - pos), &dec, &tlen);
    const size_t end = pos + (size_t)tlen;
    if (tlen >= 256) return -21;
    if (st != text::PARSE_OK) return st == text::PARSE_MALFORMED ? -21 : -22;
    uint64_t mant = 0;
    uint32_t se = 0;
    int q = 0;
    if (dec.special) {
      mant = dec.special == 2 ? (1ULL << 63) : (3ULL << 62);
      se = 0x7fff;
    } else {
      const uint32_t r = text::decimal_to_x87(dec, tab, &mant, &se, &q, force_band != 0);
      if (r == 2) return -22;
      if (r == 1) {
        text::finish_exact(dec, mant, q, scratch.data(), &mant, &se);
        slow++;
      }
    }
    const uint16_t se16 = (uint16_t)(se | (dec.neg << 15));
    memset(&values[found], 0, 16);
    memcpy((char*)&values[found], &mant, 8);
    memcpy((char*)&values[found] + 8, &se16, 2);
    found++;
    pos = end;
  }
  while (pos < len && text::is_space(s[pos])) pos++;
  if (consumed) *consumed = pos;
  if (n_exact) *n_exact = slow;
  return 0;
}

// ---- sampler (sampler.cuh, x87soft.cuh) -----------------------------------------------------

// op: 0 add, 1 sub, 2 mul; returns 0 if an operand is not representable (denormal / inf / nan).
int hostsim_x87_op(int op, const long double* a, const long double* b, long double* out) {
  RawX87 ra, rb;
  memset(&ra, 0, 16);
  memset(&rb, 0, 16);
  memcpy(&ra, a, 10);
  memcpy(&rb, b, 10);
  bool ok = true;
  X87 x = x87_load(&ra, &ok), y = x87_load(&rb, &ok);
  if (!ok) return 0;
  X87 r = op == 0 ? x87_add(x, y) : op == 1 ? x87_add(x, x87_neg(y)) : x87_mul(x, y);
  memset(out, 0, 16);
  if (r.mant) {
    const int e = r.exp + 16383;
    if (e <= 0 || e >= 0x7fff) return 0;
    const uint16_t se = (uint16_t)(e | (r.neg << 15));
    memcpy(out, &r.mant, 8);
    memcpy((char*)out + 8, &se, 2);
  }
  return 1;
}

void hostsim_x87_pivot(uint64_t w, long double* out, double* dd_hi, double* dd_lo) {
  const X87 r = x87_pivot_inclusive(w);
  memset(out, 0, 16);
  if (r.mant) {
    const uint16_t se = (uint16_t)(r.exp + 16383);
    memcpy(out, &r.mant, 8);
    memcpy((char*)out + 8, &se, 2);
  }
  const dd v = x87_to_dd(r);
  *dd_hi = v.hi;
  *dd_lo = v.lo;
}

struct HostSampler {
  std::vector<RawX87> cells, totals;
  std::vector<SegCoarse> coarse;
  std::vector<uint32_t> guide;
  std::vector<double> cells_d, totals_d;
  std::vector<SamplerSlice> slices;
  std::vector<dd> geo;
  SamplerView view;
  bool bad = false;
};

static void build_segment(const RawX87* v, double* vd, uint32_t n, SegCoarse* coarse, double* abs_out, bool* bad) {
  const uint32_t nb = (n + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
  double ab = 0.0;
  bool ok = true;
  for (uint32_t b = 0; b < nb; b++) {
    dd sum, maxp;
    double a;
    seg_block_summary(v, vd, n, b, &sum, &maxp, &a, &ok);
    coarse[b + 1].c = sum;
    coarse[b + 1].m = maxp;
    ab += a;
  }
  seg_scan(coarse, nb);
  *abs_out = ab;
  if (!ok) *bad = true;
}

void* hostsim_sampler_new(int dims, uint32_t m, uint32_t n_slices, const uint32_t* dimension,
                          const int32_t* c0, const int32_t* c1, const long double* cells,
                          const long double* slice_total, const long double* total) {
  HostSampler* h = new HostSampler;
  std::vector<std::pair<uint32_t, uint32_t>> geo_off;
  uint64_t cell_off = 0, coarse_off = 0, guide_off = 0;
  h->slices.resize(n_slices);
  for (uint32_t i = 0; i < n_slices; i++) {
    const uint32_t D = dimension[i];
    uint32_t go = 0xffffffffu;
    for (auto& g : geo_off)
      if (g.first == D) go = g.second;
    if (go == 0xffffffffu) {
      go = (uint32_t)h->geo.size();
      geo_off.emplace_back(D, go);
      std::vector<DD> t((size_t)D + 1);
      exp2_table_dd(D, t.data());
      for (auto& e : t) h->geo.push_back(make_dd(e.hi, e.lo));
    }
    SamplerSlice& sl = h->slices[i];
    sl.cell_off = cell_off;
    sl.coarse_off = coarse_off;
    sl.n_cells = dims == 2 ? D * D : D;
    sl.D = D;
    sl.c0 = c0[i];
    sl.c1 = dims == 2 ? c1[i] : 0;
    sl.geo_off = go;
    sl.guide_off = (uint32_t)guide_off;
    sl.abs_sum = 0.0;
    cell_off += sl.n_cells;
    const uint32_t nb = (sl.n_cells + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
    coarse_off += nb + 1;
    guide_off += seg_guide_size(nb) + 1;
  }
  const uint64_t tco = coarse_off, tgo = guide_off;
  coarse_off += (n_slices + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK + 1;
  guide_off += seg_guide_size((n_slices + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK) + 1;
  h->cells.resize(cell_off);
  memcpy(h->cells.data(), cells, cell_off * 16);
  h->totals.resize(n_slices);
  memcpy(h->totals.data(), slice_total, (size_t)n_slices * 16);
  h->coarse.resize(coarse_off);
  h->cells_d.assign(cell_off + QB_SEG_BLOCK, 0.0);
  h->totals_d.assign((size_t)n_slices + QB_SEG_BLOCK, 0.0);
  for (uint32_t i = 0; i < n_slices; i++)
    build_segment(h->cells.data() + h->slices[i].cell_off, h->cells_d.data() + h->slices[i].cell_off,
                  h->slices[i].n_cells, h->coarse.data() + h->slices[i].coarse_off, &h->slices[i].abs_sum,
                  &h->bad);
  SamplerView& v = h->view;
  build_segment(h->totals.data(), h->totals_d.data(), n_slices, h->coarse.data() + tco, &v.totals_abs_sum,
                &h->bad);
  const bool shadow = !(getenv("QB200_SAMPLER_DOUBLES") && getenv("QB200_SAMPLER_DOUBLES")[0] == '0');
  v.cells_d = shadow ? h->cells_d.data() : nullptr;
  v.totals_d = shadow ? h->totals_d.data() : nullptr;
  h->guide.resize(guide_off);
  for (uint32_t i = 0; i <= n_slices; i++) {   // the guide tables, as k_seg_build fills them
    const uint32_t n = i < n_slices ? h->slices[i].n_cells : n_slices;
    const SegCoarse* co = h->coarse.data() + (i < n_slices ? h->slices[i].coarse_off : tco);
    uint32_t* g = h->guide.data() + (i < n_slices ? h->slices[i].guide_off : tgo);
    const uint32_t nb = (n + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK, G = seg_guide_size(nb);
    for (uint32_t u = 0; u <= G; u++) g[u] = seg_guide_entry(co, nb, G, u);
  }
  v.guide = getenv("QB200_SAMPLER_GUIDE") && getenv("QB200_SAMPLER_GUIDE")[0] == '0' ? nullptr : h->guide.data();
  v.totals_guide_off = (uint32_t)tgo;
  v.pad0 = 0;
  v.cells = h->cells.data();
  v.coarse = h->coarse.data();
  v.slices = h->slices.data();
  v.totals = h->totals.data();
  v.totals_coarse = h->coarse.data() + tco;
  v.geo = h->geo.data();
  memset(&v.dist_total, 0, 16);
  memcpy(&v.dist_total, total, 10);
  v.n_slices = n_slices;
  v.scale_by_total = *total > 1 ? 1 : 0;
  v.m = (int)m;
  v.dims = dims;
  return h;
}

void hostsim_sampler_free(void* h) { delete (HostSampler*)h; }

// How well the guide tables bracket: for k uniform pivots per segment kind (the slice totals; the
// cells of the slice each pivot selects), the number of searches whose bracket held and the sum of
// the bracket widths in blocks. out[0..5] = searches, brackets held, width sum (totals), then cells.
void hostsim_sampler_guide_stats(void* hh, uint32_t k, const uint64_t* words, double* out) {
  HostSampler* h = (HostSampler*)hh;
  const SamplerView& s = h->view;
  for (int i = 0; i < 6; i++) out[i] = 0.0;
  auto probe = [&](const SegCoarse* coarse, const uint32_t* guide, uint32_t n, X87 p, double* o) {
    const dd pd = x87_to_dd(p);
    const uint32_t nb = (n + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
    const double top = coarse[nb].m.hi;
    o[0] += 1.0;
    if (!(top > 0.0 && pd.hi > 0.0)) return;
    const uint32_t G = seg_guide_size(nb);
    const double f = pd.hi / top * (double)G;
    const uint32_t u = f >= (double)(G - 1) ? G - 1 : (uint32_t)f;
    const uint32_t L = guide[u ? u - 1 : 0], R = guide[u + 2 < G ? u + 2 : G];
    const bool below = L == 0 || !dd_ge(coarse[L].m, pd);
    const bool above = R >= nb || dd_ge(coarse[R + 1].m, pd);
    if (below && above && L <= R) {
      o[1] += 1.0;
      o[2] += (double)(R - L);
    }
  };
  for (uint32_t i = 0; i < k; i++) {
    bool ok = true;
    X87 p = x87_pivot_inclusive(words[2 * i]);
    if (s.scale_by_total) p = x87_mul(p, x87_load(&s.dist_total, &ok));
    probe(s.totals_coarse, h->guide.data() + s.totals_guide_off, s.n_slices, p, out);
    int exact = 0;
    const uint32_t sl = sample_slice(s, words[2 * i], 0, &exact);
    if (sl >= s.n_slices) continue;
    const SamplerSlice& S = s.slices[sl];
    const X87 p2 = x87_mul(x87_pivot_inclusive(words[2 * i + 1]), x87_load(s.totals + sl, &ok));
    probe(s.coarse + S.coarse_off, h->guide.data() + S.guide_off, S.n_cells, p2, out + 3);
  }
}
int hostsim_sampler_bad(void* h) { return ((HostSampler*)h)->bad ? 1 : 0; }

// k samples, sample i from words[i * (dims + 2) ...]; out: k x 8 doubles
// (sq0_hi, sq0_lo, sq1_hi, sq1_lo, x0, x1, slice, cell), status, exact.
void hostsim_sampler_sample(void* hh, uint32_t k, const uint64_t* words, int force_exact, double* out,
                            int32_t* status, int32_t* exact) {
  HostSampler* h = (HostSampler*)hh;
  const uint32_t wps = (uint32_t)h->view.dims + 2;
  for (uint32_t i = 0; i < k; i++) {
    SampleOut o;
    sample_one(h->view, words + (size_t)i * wps, force_exact, &o);
    double* d = out + 8 * (size_t)i;
    d[0] = o.sq0_hi; d[1] = o.sq0_lo; d[2] = o.sq1_hi; d[3] = o.sq1_lo;
    d[4] = o.x0; d[5] = o.x1; d[6] = o.slice; d[7] = o.cell;
    status[i] = o.status;
    exact[i] = o.exact;
  }
}

}  // extern "C"

// ---- diagonal k sampler (diagk.cuh): the device function in a plain loop ----------------------
#include "../../qunundrum_b200/csrc/diagk_host.hpp"

extern "C" {

void* hostsim_diagk_new(uint32_t m, uint32_t sigma, uint32_t l, const uint8_t* d, size_t dn,
                        const uint8_t* r, size_t rn) {
  DiagKHost* h = new DiagKHost;
  if (diagk_prepare(m, sigma, l, d, dn, r, rn, h, &g_err)) {
    delete h;
    return nullptr;
  }
  return h;
}
void hostsim_diagk_free(void* h) { delete (DiagKHost*)h; }
void hostsim_diagk_force_exact(void* h, int on) { ((DiagKHost*)h)->c.force_exact = on ? 1 : 0; }
void hostsim_diagk_dims(void* hh, uint32_t* out3) {
  const DiagKHost* h = (const DiagKHost*)hh;
  out3[0] = h->c.k;
  out3[1] = h->c.wj;
  out3[2] = h->c.wl;
}

// n samples; j: n x wj limbs, k_out: n x wl limbs (row per sample), x: n x (hi, lo).
int hostsim_diagk_sample(void* hh, uint32_t n, const uint32_t* j, const int32_t* eta,
                         const long double* pivot, uint64_t delta_bound, uint32_t* k_out, double* x,
                         int64_t* delta, int32_t* status) {
  const DiagKHost* h = (const DiagKHost*)hh;
  std::vector<uint32_t> scratch(diagk_scratch_limbs(h->c.k));
  for (uint32_t i = 0; i < n; i++) {
    RawX87 raw;
    memset(&raw, 0, 16);
    memcpy(&raw, &pivot[i], 10);
    bool ok = true;
    const X87 p = x87_load(&raw, &ok);
    if (!ok || p.neg) return -1;
    dd xo;
    status[i] = diagk_sample<1>(h->c, j + (size_t)i * h->c.wj, eta[i], p, delta_bound, scratch.data(),
                                k_out + (size_t)i * h->c.wl, &xo, &delta[i]);
    x[2 * i] = xo.hi;
    x[2 * i + 1] = xo.lo;
  }
  return 0;
}

void hostsim_sinpi_acc(double hi, double lo, double* out2) {
  const dd r = sinpi_acc(make_dd(hi, lo));
  out2[0] = r.hi;
  out2[1] = r.lo;
}

static void store_x87(X87 r, long double* out) {
  memset(out, 0, 16);
  if (r.mant) {
    const uint16_t se = (uint16_t)((r.exp + 16383) | (r.neg << 15));
    memcpy(out, &r.mant, 8);
    memcpy((char*)out + 8, &se, 2);
  }
}

void hostsim_x87_from_dd(double hi, double lo, long double* out) { store_x87(x87_from_dd(make_dd(hi, lo)), out); }

// h at x = (hi, lo) pairs, rounded to the x87 format.
void hostsim_diagk_h(uint32_t l, uint32_t n, const double* x, long double* out) {
  for (uint32_t i = 0; i < n; i++) {
    const dd xi = make_dd(x[2 * i], x[2 * i + 1]);
    // sin^2(pi x) from the fractional part of x
    const double nearest = rint(xi.hi);
    dd t = dd_add_d(xi, -nearest);
    if (t.hi > 0.5) t = dd_add_d(t, -1.0);
    if (t.hi < -0.5) t = dd_add_d(t, 1.0);
    const dd st = sinpi_acc(t);
    store_x87(x87_from_dd(diagk_h(l, dd_mul(st, st), xi)), &out[i]);
  }
}


// ---- generator client / server tail (client_math.cuh) -------------------------------------------

static void store_raw(X87 r, long double* out, bool* ok) {
  uint64_t m = 0, se = 0;
  x87_encode(r, &m, &se, ok);
  memset(out, 0, 16);
  memcpy(out, &m, 8);
  const uint16_t s16 = (uint16_t)se;
  memcpy((char*)out + 8, &s16, 2);
}

// (long double)a / (long double)q as the kernels compute it; returns 0 if not representable.
int hostsim_x87_div_u32(const long double* a, uint32_t q, long double* out) {
  uint64_t m = 0;
  uint16_t se = 0;
  memcpy(&m, a, 8);
  memcpy(&se, (const char*)a + 8, 2);
  bool ok = true;
  const X87 r = x87_div_u32(x87_load(m, se, &ok), q);
  store_raw(r, out, &ok);
  return ok ? 1 : 0;
}

int hostsim_x87_from_double(double v, long double* out) {
  bool ok = true;
  store_raw(x87_from_double(v), out, &ok);
  return ok ? 1 : 0;
}

// distribution_slice_copy_scale: src D x D doubles -> out store x store long doubles.
int hostsim_copy_scale(int D, int store, const double* src, long double* out) {
  bool ok = true;
  for (int idx = 0; idx < store * store; idx++) store_raw(scale_cell_x87(src, D, store, idx), out + idx, &ok);
  return ok ? 1 : 0;
}

// One destination vector of linear_distribution_init_collapse_d (axis 0) / _r (axis 1): the
// source slices in order, dims[i] x dims[i] long doubles each.
int hostsim_collapse(int axis, uint32_t max_dim, uint32_t n_src, const uint32_t* dims,
                     const long double* const* cells, long double* out) {
  bool ok = true;
  for (uint32_t e = 0; e < max_dim; e++) {
    X87 acc = x87_zero();
    for (uint32_t s = 0; s < n_src; s++) {
      const uint32_t D = dims[s], q = max_dim / D;
      for (uint32_t k = 0; k < D; k++) {
        const size_t at = axis == 0 ? (size_t)(e / q) + (size_t)k * D : (size_t)k + (size_t)(e / q) * D;
        uint64_t m = 0;
        uint16_t se = 0;
        memcpy(&m, cells[s] + at, 8);
        memcpy(&se, (const char*)(cells[s] + at) + 8, 2);
        acc = collapse_step(acc, m, se, q, &ok);
      }
    }
    store_raw(acc, out + e, &ok);
  }
  return ok ? 1 : 0;
}

}  // extern "C"

// ---- exact samplers (exact.cuh): the device functions in plain loops ---------------------------
#include "../../qunundrum_b200/csrc/exact_host.hpp"

extern "C" {

void* hostsim_exact_new(int kind, uint32_t m, uint32_t l, uint32_t sigma, const uint8_t* d, size_t dn,
                        const uint8_t* r, size_t rn, uint32_t dimension_max, uint32_t emax) {
  ExactHost* h = new ExactHost;
  if (exact_prepare(kind, m, l, sigma, d, dn, r, rn, dimension_max, emax, h, &g_err)) {
    delete h;
    return nullptr;
  }
  return h;
}
void hostsim_exact_free(void* h) { delete (ExactHost*)h; }
void hostsim_exact_dims(void* hh, uint32_t* out9) {
  const ExactConst& c = ((ExactHost*)hh)->c;
  const uint32_t v[9] = {c.wa, c.wn, c.wk, c.kappa_d, c.kappa_r, c.tw, c.P, c.table_dim, c.emax};
  for (int i = 0; i < 9; i++) out9[i] = v[i];
}
const uint32_t* hostsim_exact_table(void* hh) { return ((ExactHost*)hh)->table.data(); }
const uint32_t* hostsim_exact_inverse(void* hh, int which) {
  ExactHost* h = (ExactHost*)hh;
  return (which ? h->inv_d : h->inv_r).data() + QB_EXACT_PAD;
}

// The bytes random_generate_mpz reads for a region (0: see *status).
uint32_t hostsim_exact_region_bytes(void* hh, int32_t min_log_alpha, uint32_t region, uint32_t dimension,
                                    int32_t* status) {
  const ExactConst& c = ((ExactHost*)hh)->c;
  std::vector<uint32_t> lo(c.wa), M(c.wa + 1);
  ExactRegion g;
  g.min_log_alpha = min_log_alpha;
  g.region = region;
  g.dimension = dimension;
  g.length = 0;
  g.offset = 0;
  int st = 0;
  const uint32_t bits = exact_region_modulus<1, 1>(c, g, lo.data(), M.data(), &st);
  *status = st;
  return st == QB_EXACT_OK ? exact_bytes_for_bits(bits) : 0;
}

void hostsim_exact_alpha(void* hh, uint32_t n, const ExactRegion* regions, uint32_t kappa, const uint8_t* stream,
                         uint64_t stream_len, uint32_t* alpha, int32_t* negative, int32_t* status) {
  const ExactConst& c = ((ExactHost*)hh)->c;
  std::vector<uint32_t> scratch(exact_alpha_scratch_limbs(c));
  for (uint32_t i = 0; i < n; i++) {
    int neg = 0;
    status[i] = exact_alpha<1, 1>(c, regions[i], kappa, stream, stream_len, scratch.data(),
                                  alpha + (size_t)i * c.wa, &neg);
    negative[i] = neg;
  }
}

// mode 0 / 3: j from alpha_r (t: rows of max(1, ceil(kappa_r / 32)) words, NULL when kappa_r = 0);
// mode 1: (j, k) from (alpha_d, alpha_r); mode 2: j from (alpha_d, k) with k and t given.
void hostsim_exact_jk(void* hh, int mode, uint32_t n, const uint32_t* alpha_d, const int32_t* neg_d,
                      const uint32_t* alpha_r, const int32_t* neg_r, const uint32_t* t, uint32_t* j,
                      uint32_t* k) {
  const ExactConst& c = ((ExactHost*)hh)->c;
  std::vector<uint32_t> scratch(exact_jk_scratch_limbs(c));
  const uint32_t kap = mode == 2 ? c.kappa_d : c.kappa_r;
  const uint32_t tl = kap ? (kap + 31) / 32 : 1;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t* ti = t ? t + (size_t)i * tl : nullptr;
    uint32_t* ji = j + (size_t)i * c.wn;
    if (mode == 2) {
      exact_j_from_alpha_d_k<8, 1, 1, 1, 1, 1>(c, alpha_d + (size_t)i * c.wa, neg_d[i], k + (size_t)i * c.wk, ti,
                                            scratch.data(), ji);
      continue;
    }
    exact_j_from_alpha_r<8, 1, 1, 1, 1>(c, alpha_r + (size_t)i * c.wa, neg_r[i], ti, scratch.data(), ji);
    if (mode == 1)
      exact_k_from_alpha_d_j<8, 1, 1, 1, 1>(c, alpha_d + (size_t)i * c.wa, neg_d[i], ji, scratch.data(),
                                         k + (size_t)i * c.wk);
  }
}

// exact_mod on given numbers: V (nv limbs, room for one more) mod M (wm limbs, room for one more).
void hostsim_exact_mod(uint32_t* V, uint32_t nv, uint32_t* M, uint32_t wm) { exact_mod<1>(V, nv, M, wm); }

}  // extern "C"
