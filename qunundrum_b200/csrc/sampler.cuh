// sampler.cuh -- sampling (alpha_d, alpha_r) / alpha from a stored distribution, one sample
// per thread.
//
// Replaces, for a batch of samples, what the reference does per sample in
//   distribution_sample_approximate_alpha_d_r   src/distribution.cpp:464-527
//     distribution_sample_slice                 src/distribution.cpp:359-409   (walk over <= 6404 slices)
//     distribution_slice_sample_region          src/distribution_slice.cpp:167-228 (walk over D^2 cells)
//     distribution_slice_region_coordinates     src/distribution_slice.cpp:130-165
//     sample_approximate_alpha_from_region x 2  src/sample.cpp:24-77
//   linear_distribution_sample_approximate_alpha src/linear_distribution.cpp:618-666 (same, one axis)
// which is what tau_estimate / tau_estimate_linear (src/tau_estimate.cpp:23-133) spend their
// time in: 10^6 estimates of n samples per tried n in estimate_runs_*
// (src/executables_estimate_runs_distribution.h:23-28), every sample two linear long double
// walks of up to 6404 + 65,536 steps.
//
// Here a walk is a binary search over prefix sums kept per block of 8 elements (exact
// double-double images of the x87 numbers) followed by a scan of one block. The reference's
// walk rounds to 64 bits at every step, so its stopping index can differ from the exact one
// only when the pivot lies within the accumulated rounding error B of a prefix sum: the fast
// path accepts an index only if the pivot clears every earlier prefix (their running maximum --
// Richardson cells may be negative, so prefixes are not monotone) and the stopping prefix by
// more than B; otherwise the walk is replayed bit for bit with x87soft.cuh. The results are
// therefore the reference's for the same random words, not merely the same distribution.
//
// Everything is __host__ __device__: tests/hostsim runs the identical code on the CPU.
#pragma once

#include "integrands.cuh"
#include "x87soft.cuh"

namespace qb200 {

// Elements per block of the coarse index. Measured on the bench distribution (B200, 2^20 samples
// per call): 32 -> 2.03e9 samples/s, 16 -> 2.97e9, 8 -> 3.78e9, 4 -> 4.29e9; the index costs
// 32 / QB_SEG_BLOCK bytes per 16-byte cell. 8: +25 % memory for 1.9 x.
#ifndef QB_SEG_BLOCK
#define QB_SEG_BLOCK 8
#endif

// The structures below are read by threads whose addresses have nothing to do with each other:
// every load instruction of a warp is 32 look-ups in L1, and the kernel is bound by exactly those
// (ncu, round 2: 0.74 tag look-ups per cycle and SM, issue slots 25 % active). They are 16-byte
// aligned and read 16 bytes at a time (ld16 below): half the look-ups of the 8-byte loads the
// compiler emits for members it knows only to be 8-byte aligned.
struct alignas(16) RawX87 {  // the 16 bytes of an x86-64 long double
  uint64_t mant;
  uint64_t se;   // low 16 bits: sign and exponent; the rest is padding (ignored)
};

struct alignas(16) SegCoarse {  // state of the walk BEFORE a block
  dd c;             // prefix sum
  dd m;             // running maximum of the earlier prefix sums (-1e300 before the first)
};

// One 16-byte load from an array in GLOBAL memory (not for a member of a kernel parameter). As
// inline PTX: written as a plain 16-byte access the compiler narrows it to the 8 + 4 bytes that
// are used and is back to two look-ups.
QHD RawX87 ld16(const RawX87* p) {
#if defined(__CUDA_ARCH__)
  RawX87 r;
  asm("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.mant), "=l"(r.se) : "l"(__cvta_generic_to_global(p)));
  return r;
#else
  return *p;
#endif
}
QHD dd ld16(const dd* p) {
#if defined(__CUDA_ARCH__)
  dd r;
  asm("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.hi), "=d"(r.lo) : "l"(__cvta_generic_to_global(p)));
  return r;
#else
  return *p;
#endif
}

struct alignas(16) SamplerSlice {
  uint64_t cell_off;    // first cell in SamplerView::cells
  uint64_t coarse_off;  // first entry in SamplerView::coarse (n_blocks + 1 entries)
  uint32_t n_cells;
  uint32_t D;
  int32_t c0, c1;       // min_log_alpha_d, min_log_alpha_r (linear: min_log_alpha, 0)
  uint32_t geo_off;     // 2^(j / D), j = 0 .. D, in SamplerView::geo
  uint32_t guide_off;   // first entry of this slice's guide table in SamplerView::guide
  double abs_sum;       // sum of |cell|: scale of the rounding-error band
};

struct SamplerView {
  const RawX87* cells;
  const SegCoarse* coarse;
  const SamplerSlice* slices;
  const RawX87* totals;            // the slices' total_probability, in walk order
  const double* cells_d;           // the top 53 bits of every cell / total as a double (what the
  const double* totals_d;          // quick pass reads: 8 instead of 16 bytes per element); may be null
  const SegCoarse* totals_coarse;
  const dd* geo;
  const uint32_t* guide;           // guide tables (seg_guide_*): the slices', then the totals'
  uint32_t totals_guide_off;
  uint32_t pad0;
  double totals_abs_sum;
  RawX87 dist_total;               // distribution->total_probability
  uint32_t n_slices;
  int scale_by_total;              // total_probability > 1 (src/distribution.cpp:373)
  int m;
  int dims;                        // 2: Distribution, 1: Linear_Distribution
};

struct alignas(16) SampleOut {  // 64 bytes
  double sq0_hi, sq0_lo;  // (alpha_d / 2^m)^2   (linear: (alpha / 2^m)^2)
  double sq1_hi, sq1_lo;  // (alpha_r / 2^m)^2
  double x0, x1;          // alpha / 2^m rounded to double, signed
  int32_t slice, cell;
  int32_t status;         // 0 ok, 1 out of bounds (reference returns FALSE), 2 no region (reference: critical)
  int32_t exact;          // number of walks that needed the bit-exact replay
};

enum { kSampleOk = 0, kSampleOutOfBounds = 1, kSampleNoRegion = 2 };

QHD bool dd_ge(dd a, dd b) { return a.hi > b.hi || (a.hi == b.hi && a.lo >= b.lo); }
QHD dd dd_max(dd a, dd b) { return dd_ge(a, b) ? a : b; }

// The top 53 bits of an element as a double (truncated; zero for zeros and for magnitudes
// outside the double range -- the callers' error bands cover both).
QHD double x87_raw_to_double(const RawX87 r) {
  const uint32_t se = (uint32_t)r.se;
  const int e = (int)(se & 0x7fffu) - 16383 + 1023;
  if ((se & 0x7fffu) == 0 || e < 1 || e > 2046) return 0.0;
  return qb_bits_to_double(((uint64_t)(se & 0x8000u) << 48) | ((uint64_t)e << 52) |
                           ((r.mant >> 11) & 0xfffffffffffffull));
}

QHD X87 x87_load16(const RawX87* p, bool* ok) {  // p: an element of an array in global memory
  X87 v;
  const RawX87 r = ld16(p);
  if (!x87_decode(r.mant, (uint32_t)r.se & 0xffffu, &v)) *ok = false;
  return v;
}

QHD X87 x87_load(const RawX87* p, bool* ok) {
  X87 v;
  const RawX87 r = *p;
  if (!x87_decode(r.mant, (uint32_t)r.se & 0xffffu, &v)) *ok = false;
  return v;
}

// Summary of block b of a segment: its sum, the maximum of its own prefix sums (from 0), the
// sum of magnitudes, and whether every element decodes.
QHD void seg_block_summary(const RawX87* v, double* vd, uint32_t n, uint32_t b, dd* sum, dd* maxp,
                           double* abs_sum, bool* ok) {
  dd c = make_dd(0.0, 0.0), mx = make_dd(-1e300, 0.0);
  double ab = 0.0;
  const uint32_t lo = b * QB_SEG_BLOCK, hi = lo + QB_SEG_BLOCK < n ? lo + QB_SEG_BLOCK : n;
  for (uint32_t k = lo; k < hi; k++) {
    if (vd) vd[k] = x87_raw_to_double(v[k]);  // the copy the quick pass reads (seg_block_doubles)
    const dd x = x87_to_dd(x87_load(v + k, ok));
    c = dd_add(c, x);
    mx = dd_max(mx, c);
    ab += fabs(x.hi);
  }
  *sum = c;
  *maxp = mx;
  *abs_sum = ab;
}

// In place: entries 1 .. n_blocks of `coarse` hold the block summaries {sum, maxp} of blocks
// 0 .. n_blocks - 1; turn them into the walk states before each block (entry n_blocks: after
// the last one).
QHD void seg_scan(SegCoarse* coarse, uint32_t n_blocks) {
  dd c = make_dd(0.0, 0.0), m = make_dd(-1e300, 0.0);
  coarse[0].c = c;
  coarse[0].m = m;
  for (uint32_t b = 0; b < n_blocks; b++) {
    const dd sum = coarse[b + 1].c, maxp = coarse[b + 1].m;
    m = dd_max(m, dd_add(c, maxp));
    c = dd_add(c, sum);
    coarse[b + 1].c = c;
    coarse[b + 1].m = m;
  }
}

// ---- guide tables ------------------------------------------------------------------------------
// The search for the stopping block is a binary search over the running maxima: 10 + 11 DEPENDENT
// loads per sample for the bench distribution, each a trip to L1 / L2 -- the kernel's limit (ncu,
// round 1: L1/TEX-bound, issue slots 34 % active). A guide table per segment cuts the chain:
// guide[u] = first block whose running maximum reaches u / G of the segment's final maximum
// (G = a quarter to half of the number of blocks). The pivot's bucket u = floor(p / top * G)
// brackets the answer between guide[u - 1] and guide[u + 2] -- a bracket that is CHECKED against
// the running maxima themselves (two independent loads), so rounding in the table or in u can only
// cost a fall-back to the full search, never a different result -- and the binary search runs over
// a handful of blocks: 1 + 1 + ~2 dependent loads instead of 11.
QHD uint32_t seg_guide_size(uint32_t n_blocks) {
  uint32_t G = 1;
  while (G * 4 <= n_blocks) G <<= 1;
  return G;
}

// Entry u (0 <= u <= G) of the guide table of a segment with n_blocks blocks.
QHD uint32_t seg_guide_entry(const SegCoarse* coarse, uint32_t n_blocks, uint32_t G, uint32_t u) {
  if (u >= G) return n_blocks;
  const double top = coarse[n_blocks].m.hi;
  if (!(top > 0.0)) return 0;
  const double th = top * ((double)u / (double)G);
  uint32_t lo = 0, hi = n_blocks;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (coarse[mid + 1].m.hi >= th)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

// Rare paths are compiled out of line, so that the common path stays short.
#if defined(__CUDACC__)
#define QB_SEG_SLOW inline __host__ __device__ __noinline__
#else
#define QB_SEG_SLOW inline
#endif

// The reference's walk, replayed bit for bit: first k with pivot - v[0] - ... - v[k] <= 0.
QB_SEG_SLOW uint32_t seg_walk_exact(const RawX87* v, uint32_t n, X87 p) {
  bool ok = true;
  for (uint32_t k = 0; k < n; k++) {
    p = x87_add(p, x87_neg(x87_load(v + k, &ok)));
    if (x87_nonpositive(p)) return k;
  }
  return n;
}

// The block [k0, k1) in double-double, for the searches the pass in doubles leaves open (~1e-10 of
// them).
QB_SEG_SLOW uint32_t seg_find_slow(const RawX87* v, const SegCoarse* coarse, uint32_t lo, uint32_t k0,
                                   uint32_t k1, uint32_t n, double unit, X87 p, dd pd, int* exact) {
  dd c = coarse[lo].c, mprev = coarse[lo].m;
  bool ok = true;
  for (uint32_t k = k0; k < k1; k++) {
    c = dd_add(c, x87_to_dd(x87_load(v + k, &ok)));
    if (dd_ge(c, pd)) {
      const double band = (double)(k + 2) * unit;
      const dd over = dd_add(c, dd_neg(pd)), clear = dd_add(pd, dd_neg(mprev));
      if (over.hi > band && clear.hi > band) return k;
      *exact += 1;
      return seg_walk_exact(v, n, p);
    }
    mprev = dd_max(mprev, c);
  }
  // the block summary promised a hit inside this block; rounding of the summary itself
  *exact += 1;
  return seg_walk_exact(v, n, p);
}


// ---- the search, in three steps -------------------------------------------------------------
// seg_locate: the block the walk stops in (or the answer itself where no block is needed);
// seg_block_doubles: the block's elements as doubles; seg_decide: the stopping element.
// seg_find chains them. (The split exists because the middle step was tried with the whole warp --
// kernels_sampler.cuh -- and is kept: the pieces are easier to read and to test.)
// mode (test switch): 0 normal; 1 every walk through the bit-exact replay; 2 skip the quick pass in
// doubles (every search through the double-double path). *exact is incremented when the replay ran.

// 1 with *res = the first k at which the reference's walk stops (or n); 0 with *lo = the block to
// examine.
QHD int seg_locate(const RawX87* v, const SegCoarse* coarse, const uint32_t* guide, uint32_t n,
                   double abs_sum, X87 p, int mode, int* exact, uint32_t* lo_out, uint32_t* res) {
  *lo_out = 0;
  *res = 0;
  if (n == 0) return 1;
  if (mode == 1) {
    *exact += 1;
    *res = seg_walk_exact(v, n, p);
    return 1;
  }
  const dd pd = x87_to_dd(p);
  const uint32_t nb = (n + QB_SEG_BLOCK - 1) / QB_SEG_BLOCK;
  // smallest block whose running maximum at its end reaches the pivot. (Binary on purpose: the
  // kernel is bound by L1 lookups of divergent addresses, 82 % of peak in ncu; 4-, 8- and
  // 16-ary searches with independent probes per level were 11 %, 25 % and 44 % slower.)
  uint32_t lo = 0, hi = nb;  // answer in [lo, hi]; hi == nb: none
  if (guide) {
    const double top = coarse[nb].m.hi;
    if (top > 0.0 && pd.hi > 0.0) {
      const uint32_t G = seg_guide_size(nb);
      const double f = pd.hi / top * (double)G;
      const uint32_t u = f >= (double)(G - 1) ? G - 1 : (uint32_t)f;
      const uint32_t L = guide[u ? u - 1 : 0], R = guide[u + 2 < G ? u + 2 : G];
      // the bracket holds iff block L - 1 does not reach the pivot and block R does (or R == nb)
      const bool below = L == 0 || !dd_ge(ld16(&coarse[L].m), pd);
      const bool above = R >= nb || dd_ge(ld16(&coarse[R + 1].m), pd);
      if (below && above && L <= R) {
        lo = L;
        hi = R < nb ? R : nb;
      }
    }
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (dd_ge(ld16(&coarse[mid + 1].m), pd))
      hi = mid;
    else
      lo = mid + 1;
  }
  if (lo == nb) {
    // no prefix reaches the pivot: certain only if the largest one misses it by more than B
    const double unit = 1.0842021724855044e-19 * (fabs(pd.hi) + abs_sum);  // 2^-63 * scale
    const dd gap = dd_add(pd, dd_neg(coarse[nb].m));
    if (gap.hi > (double)(n + 2) * unit) {
      *res = n;
      return 1;
    }
    *exact += 1;
    *res = seg_walk_exact(v, n, p);
    return 1;
  }
  *lo_out = lo;
  return 0;
}

// The top 53 bits of the elements of block lo, zero past the end of the segment. vd (may be null):
// the same values kept as doubles by seg_block_summary -- four 16-byte loads instead of eight:
// 6.3 -> 7.6e9 samples/s on the bench distribution for half as much memory again. (The array
// behind vd is readable for a whole block past the last element of the distribution.)
QHD void seg_block_doubles(const RawX87* v, const double* vd, uint32_t n, uint32_t lo, double* x) {
  const uint32_t k0 = lo * QB_SEG_BLOCK;
  if (vd && (((size_t)(vd + k0)) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < QB_SEG_BLOCK; q += 2) {
      const dd pair = ld16(reinterpret_cast<const dd*>(vd + k0 + q));
      x[q] = k0 + q < n ? pair.hi : 0.0;
      x[q + 1] = k0 + q + 1 < n ? pair.lo : 0.0;
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < QB_SEG_BLOCK; q++)  // independent loads
    x[q] = k0 + q < n ? x87_raw_to_double(ld16(v + k0 + q)) : 0.0;
}

// Quick pass over block lo in plain doubles (x: seg_block_doubles): its prefix sums are within
// 2^-49 * scale of the exact ones, so a hit that clears the pivot -- and a pivot that clears every
// earlier prefix -- by 2^-47 * scale is the exact walk's and, a fortiori (2^-47 >> n 2^-63), the
// reference's. All but ~1e-10 of the searches end here.
// The pass has ONE exit: with a `return` per element the threads of a warp reached the code
// after the search at up to eight different times and stayed apart to the end of the kernel
// (ncu, round 2: 16 threads per instruction after the first search, 9 after the second).
QHD uint32_t seg_decide(const double* x, const RawX87* v, const SegCoarse* coarse, uint32_t lo, uint32_t n,
                        double abs_sum, X87 p, int mode, int* exact) {
  const dd pd = x87_to_dd(p);
  const double unit = 1.0842021724855044e-19 * (fabs(pd.hi) + abs_sum);  // 2^-63 * scale
  const uint32_t k0 = lo * QB_SEG_BLOCK, k1 = k0 + QB_SEG_BLOCK < n ? k0 + QB_SEG_BLOCK : n;
  if (mode != 2) {
    const double wide = 7.105427357601002e-15 * (fabs(pd.hi) + abs_sum);  // 2^-47 * scale
    double c = coarse[lo].c.hi, mprev = coarse[lo].m.hi;
    int hit = -1;
    bool clear = false;
#pragma unroll
    for (int q = 0; q < QB_SEG_BLOCK; q++) {
      c += x[q];
      if (hit < 0 && k0 + q < k1 && c >= pd.hi) {
        hit = q;
        clear = c - pd.hi > wide && pd.hi - mprev > wide;  // else: too close to call in doubles
      }
      mprev = hit < 0 ? fmax(mprev, c) : mprev;
    }
    if (clear) return k0 + (uint32_t)hit;
  }
  return seg_find_slow(v, coarse, lo, k0, k1, n, unit, p, pd, exact);
}

// First k at which the reference's walk stops, or n.
QHD uint32_t seg_find(const RawX87* v, const double* vd, const SegCoarse* coarse, const uint32_t* guide,
                      uint32_t n, double abs_sum, X87 p, int mode, int* exact) {
  uint32_t lo, res;
  if (seg_locate(v, coarse, guide, n, abs_sum, p, mode, exact, &lo, &res)) return res;
  double x[QB_SEG_BLOCK];
#pragma unroll
  for (int q = 0; q < QB_SEG_BLOCK; q++) x[q] = 0.0;
  if (mode != 2) seg_block_doubles(v, vd, n, lo, x);
  return seg_decide(x, v, coarse, lo, n, abs_sum, p, mode, exact);
}

// |alpha| / 2^m for region j of an axis with coordinate k and dimension D, and fraction word w
// (src/sample.cpp:42-61): alpha_min + (alpha_max - alpha_min) * (double)fraction, with
// alpha_min/max = round(2^(|k| + j / D)) and fraction = (w mod 2^63) / 2^63 (src/random.c:137-156).
QHD dd sample_axis(const SamplerView& s, const SamplerSlice& sl, int32_t k, uint32_t j, uint64_t w) {
  const int k_abs = k < 0 ? -k : k;
  const dd xmin = grid_x(ld16(s.geo + sl.geo_off + j), k_abs, 1, s.m);
  const dd xmax = grid_x(ld16(s.geo + sl.geo_off + j + 1), k_abs, 1, s.m);
  const double f = (double)(w & 0x7fffffffffffffffull) * 1.0842021724855044e-19;  // RN to 53 bits, / 2^63
  const dd x = dd_add(xmin, dd_mul_d(dd_add(xmax, dd_neg(xmin)), f));
  return x;
}

// First half of a sample: the slice the reference's walk over the slice totals stops at
// (src/distribution.cpp:359-409), or n_slices (out of bounds: the reference returns FALSE).
QHD X87 sample_slice_pivot(const SamplerView& s, uint64_t w0) {
  bool ok = true;
  X87 p = x87_pivot_inclusive(w0);
  if (s.scale_by_total) p = x87_mul(p, x87_load(&s.dist_total, &ok));
  return p;
}
QHD const uint32_t* sample_totals_guide(const SamplerView& s) {
  return s.guide ? s.guide + s.totals_guide_off : nullptr;
}
QHD uint32_t sample_slice(const SamplerView& s, uint64_t w0, int mode, int* exact) {
  return seg_find(s.totals, s.totals_d, s.totals_coarse, sample_totals_guide(s), s.n_slices, s.totals_abs_sum,
                  sample_slice_pivot(s, w0), mode, exact);
}

// Second half: the region inside slice i (src/distribution_slice.cpp:167-228) and the two axis
// draws (src/sample.cpp:24-77). w: the sample's words (w[0] is not read again).
QHD X87 sample_region_pivot(const SamplerView& s, uint32_t i, uint64_t w1) {
  bool ok = true;
  return x87_mul(x87_pivot_inclusive(w1), x87_load16(s.totals + i, &ok));
}
QHD const uint32_t* sample_slice_guide(const SamplerView& s, const SamplerSlice& sl) {
  return s.guide ? s.guide + sl.guide_off : nullptr;
}
// The sample once its cell c of slice i (= sl) is known.
QHD void sample_finish(const SamplerView& s, const SamplerSlice& sl, uint32_t i, uint32_t c, const uint64_t* w,
                       SampleOut* out) {
  out->slice = (int32_t)i;
  if (c >= sl.n_cells) {
    out->status = kSampleNoRegion;
    return;
  }
  out->cell = (int32_t)c;
  out->status = kSampleOk;
  if (s.dims == 2) {
    const dd xd = sample_axis(s, sl, sl.c0, c % sl.D, w[2]);
    const dd xr = sample_axis(s, sl, sl.c1, c / sl.D, w[3]);
    const dd qd = dd_mul(xd, xd), qr = dd_mul(xr, xr);
    out->sq0_hi = qd.hi; out->sq0_lo = qd.lo;
    out->sq1_hi = qr.hi; out->sq1_lo = qr.lo;
    out->x0 = sl.c0 < 0 ? -xd.hi : xd.hi;
    out->x1 = sl.c1 < 0 ? -xr.hi : xr.hi;
  } else {
    const dd x = sample_axis(s, sl, sl.c0, c, w[2]);
    const dd q = dd_mul(x, x);
    out->sq0_hi = q.hi; out->sq0_lo = q.lo;
    out->x0 = sl.c0 < 0 ? -x.hi : x.hi;
  }
}
QHD void sample_in_slice(const SamplerView& s, uint32_t i, const uint64_t* w, int mode, SampleOut* out, int* exact) {
  const SamplerSlice sl = s.slices[i];
  const uint32_t c = seg_find(s.cells + sl.cell_off, s.cells_d ? s.cells_d + sl.cell_off : nullptr,
                              s.coarse + sl.coarse_off, sample_slice_guide(s, sl), sl.n_cells, sl.abs_sum,
                              sample_region_pivot(s, i, w[1]), mode, exact);
  sample_finish(s, sl, i, c, w, out);
}

QHD void sample_out_clear(SampleOut* out) {
  out->sq0_hi = out->sq0_lo = out->sq1_hi = out->sq1_lo = 0.0;
  out->x0 = out->x1 = 0.0;
  out->slice = out->cell = -1;
  out->exact = 0;
  out->status = kSampleOutOfBounds;
}

// One sample from the words w[0 .. dims + 1] (the reference's draws, in its order: slice
// pivot, region pivot, one fraction per axis).
QHD void sample_one(const SamplerView& s, const uint64_t* w, int mode, SampleOut* out) {
  sample_out_clear(out);
  // (a counter of its own: the address goes to an out-of-line function, and with &out->exact the
  // whole result lived in local memory)
  int exact = 0;
  const uint32_t i = sample_slice(s, w[0], mode, &exact);
  if (i < s.n_slices) sample_in_slice(s, i, w, mode, out, &exact);  // else: status out of bounds
  out->exact = exact;
}

// Sum of the squares of the n samples of one estimate, in sample order (fixed => reproducible).
// Returns the index of the first failing sample, or n.
QHD uint32_t tau_sums(const SampleOut* o, uint32_t n, dd* s0, dd* s1, int* status) {
  dd a = make_dd(0.0, 0.0), b = make_dd(0.0, 0.0);
  *status = kSampleOk;
  for (uint32_t i = 0; i < n; i++) {
    if (o[i].status != kSampleOk) {
      *status = o[i].status;
      *s0 = a;
      *s1 = b;
      return i;
    }
    a = dd_add(a, make_dd(o[i].sq0_hi, o[i].sq0_lo));
    b = dd_add(b, make_dd(o[i].sq1_hi, o[i].sq1_lo));
  }
  *s0 = a;
  *s1 = b;
  return n;
}

}  // namespace qb200
