// kernels_text.cuh -- the text exporter's kernel: n values -> "%.24Lg\n" lines,
// contiguous, in order, in ONE pass.
//
// One thread per value (textfmt.cuh does the arithmetic), 1024 values per tile (four
// sub-tiles of 256, formatted and rendered one after the other).
// Line lengths vary (2..33 bytes), so tile t's text starts at the sum of all
// earlier tiles' lengths: tiles take tickets in launch order and chain their
// lengths through a decoupled look-back (one 64-bit status word per tile: flag
// in the top two bits, byte count below), so the text is written once, to its
// final place, with no second pass and no scratch copy. Algorithmic traffic:
// 16 B read + ~31 B written per value; the 319 KB table of powers of ten stays
// in L1/L2. In practice the kernel is bound by integer issue slots (~800 thread
// instructions per value), not by HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "textfmt.cuh"
#include "textparse.cuh"

namespace qb200 {
namespace text {

constexpr int TB = 256;                            // worker threads per block
#ifndef QB_TEXT_SUB
#define QB_TEXT_SUB 4
#endif
constexpr int SUB = QB_TEXT_SUB;                             // values per worker thread (sub-tiles per tile)
constexpr int TILE = TB * SUB;                     // values per tile
constexpr int STAGE_BYTES = TILE * MAX_TEXT + 32;  // tile text staged in shared memory

constexpr unsigned long long ST_FLAG_AGG = 1ULL << 62;     // tile length published
constexpr unsigned long long ST_FLAG_PREFIX = 2ULL << 62;  // inclusive prefix published
constexpr unsigned long long ST_VALUE = (1ULL << 62) - 1;

enum : int { SRC_X87 = 0, SRC_F64 = 1 };

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of `tile_len` over the tiles before `tile` (all lanes of one
// warp call this; every lane returns the sum). Tiles take their numbers from a
// ticket counter, so every earlier tile is already running and publishes at least
// its own length without waiting for anyone: no deadlock. Each lane keeps LB polls
// in flight, so one round trip to L2 covers a window of 32 * LB tiles: the prefix
// front advances that many tiles per round trip.
constexpr int LB = 2;

__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status,
                                                                 unsigned int tile,
                                                                 unsigned long long tile_len,
                                                                 int lane) {
  if (tile == 0) {
    if (lane == 0) st_status(status, ST_FLAG_PREFIX | tile_len);
    return 0;
  }
  if (lane == 0) st_status(status + tile, ST_FLAG_AGG | tile_len);
  unsigned long long excl = 0;
  long long j0 = (long long)tile - 1;  // poll k of lane l looks at tile j0 - 32 k - l
  while (true) {
    unsigned long long v[LB];
#pragma unroll
    for (int k = 0; k < LB; k++) {
      const long long j = j0 - 32 * k - lane;
      v[k] = j >= 0 ? ld_status(status + j) : ST_FLAG_PREFIX;  // before tile 0: empty prefix
    }
    bool found = false;
#pragma unroll
    for (int k = 0; k < LB; k++) {
      if (found) continue;  // uniform
      const long long j = j0 - 32 * k - lane;
      while ((v[k] >> 62) == 0) v[k] = ld_status(status + j);
      const unsigned has_prefix = __ballot_sync(0xffffffffu, (v[k] >> 62) == 2);
      const int first = has_prefix ? (__ffs(has_prefix) - 1) : 32;  // nearest tile with a prefix
      unsigned long long part = (lane <= first) ? (v[k] & ST_VALUE) : 0ULL;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      excl += part;
      found = has_prefix != 0;
    }
    if (found) break;
    j0 -= 32 * LB;
  }
  if (lane == 0) st_status(status + tile, ST_FLAG_PREFIX | (excl + tile_len));
  return excl;
}

// status: n_tiles words, zeroed before the launch; ticket: one zeroed word.
// out: the text (capacity cap bytes); total_out: total text length (written by the
// last tile even if it exceeds cap, in which case nothing beyond cap is stored).
//
// Block = 8 worker warps + 1 scan warp. Per sub-tile the workers format one value
// each, scan the line lengths and render their lines into shared memory behind the
// previous sub-tiles' text. As soon as the last sub-tile's lengths are known the scan
// warp chains the tile's length into the global prefix (look-back) while the workers
// render that sub-tile. Then everybody copies the staged text to out + prefix.
// Tiles must be this large: with 256-value tiles the prefix front (one L2 round trip
// per window of tiles) could not keep up with the rate at which tiles complete, and
// the kernel ran at 54 % of its chain-free speed.
constexpr int TEXT_THREADS = TB + 32;

template <int SRC>
__global__ void __launch_bounds__(TEXT_THREADS)
    k_text_format(const void* __restrict__ in, unsigned long long n,
                  const Pow10Entry* __restrict__ tab, unsigned char* __restrict__ out,
                  unsigned long long cap, unsigned long long* __restrict__ status,
                  unsigned int* __restrict__ ticket, unsigned long long* __restrict__ total_out,
                  unsigned long long* __restrict__ n_exact, int force_band) {
  __shared__ __align__(16) unsigned char stage[STAGE_BYTES];
  __shared__ uint32_t big[BIG_LIMBS];
  __shared__ uint32_t warp_sum[2][TB / 32];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool worker = warp < TB / 32;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  uint32_t running = 0;  // text bytes of the sub-tiles done so far

#pragma unroll 1
  for (int sub = 0; sub < SUB; sub++) {
    const unsigned long long i = ((unsigned long long)tile * SUB + sub) * TB + tid;
    const bool live = worker && i < n;
    Piece p;
    Dec24 d;
    uint64_t Mn = 0;
    int qn = 0;
    bool number = false;
    d.undecided = 0;
    d.up = 0;
    int len = 0;
    if (live) {
      if (SRC == SRC_X87) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(in) + i);
        number = classify_x87(v.x, (uint32_t)v.y & 0xffffu, &p, &Mn, &qn);
      } else {
        const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long*>(in) + i);
        number = classify_f64(v, &p, &Mn, &qn);
      }
      if (number) digits24(Mn, qn, tab, &d, force_band == 1);
      if (!d.undecided) {
        if (number) {
          round_digits(&d, d.up != 0);
          piece_from_digits(d, &p);
        }
        len = piece_length(p);
      }
    }
    // Undecided roundings (true ties of values >= 1, or a remainder inside the
    // 2^-107 error band): exact integer arithmetic, one thread at a time on the
    // block's scratch. Never taken for probabilities; kept for exactness.
    if (__syncthreads_or((int)d.undecided)) {
      for (int t = 0; t < TB; t++) {
        if (t == tid && d.undecided) {
          d.up = exact_round_up(Mn, qn, d.x, d.c0, d.c1, d.c2, big) ? 1u : 0u;
          round_digits(&d, d.up != 0);
          piece_from_digits(d, &p);
          len = piece_length(p);
          if (n_exact) atomicAdd(n_exact, 1ULL);
        }
        __syncthreads();
      }
    }
    // exclusive scan of the line lengths over the sub-tile
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (worker && lane == 31) warp_sum[sub & 1][warp] = (uint32_t)incl;
    __syncthreads();
    uint32_t before = 0, sub_len = 0;
#pragma unroll
    for (int w = 0; w < TB / 32; w++) {
      const uint32_t s = warp_sum[sub & 1][w];
      if (w < warp) before += s;
      sub_len += s;
    }
    if (worker) {
      if (live) piece_render(p, stage + running + before + (uint32_t)(incl - len));
    } else if (sub == SUB - 1) {
      // the scan warp: decoupled look-back over the tiles before this one
      // (force_band == 2: timing experiment without the chain, text layout is then wrong)
      const unsigned long long tile_len = running + sub_len;
      const unsigned long long excl = force_band == 2
                                          ? (unsigned long long)tile * (TILE * 30ULL)
                                          : lookback_exclusive(status, tile, tile_len, lane);
      if (lane == 0) {
        s_base = excl;
        if ((unsigned long long)(tile + 1) * TILE >= n) *total_out = excl + tile_len;
      }
    }
    running += sub_len;
  }
  __syncthreads();
  const uint32_t tile_len = running;

  // stage -> out + base (destination arbitrarily aligned): 4-byte words, coalesced
  const unsigned long long base = s_base;
  if (base + tile_len > cap) return;
  unsigned char* dst = out + base;
  const uint32_t head0 = (4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u;
  const uint32_t head = head0 < tile_len ? head0 : tile_len;
  if ((uint32_t)tid < head) dst[tid] = stage[tid];
  const uint32_t nwords = (tile_len - head) >> 2;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage);
  const uint32_t sel = 0x3210u + 0x1111u * head;  // bytes head .. head+3 of (hi:lo)
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  for (uint32_t w = tid; w < nwords; w += TEXT_THREADS)
    dw[w] = __byte_perm(sw[w], sw[w + 1], sel);
  const uint32_t done = head + 4u * nwords;
  if (done + tid < tile_len) dst[done + tid] = stage[done + tid];
}

// ---- importer: text -> values ------------------------------------------------------

constexpr int TOK_BYTES = 16;            // text bytes per thread in the tokenizer
constexpr int MAX_TOKEN = 256;           // longer tokens are reported as malformed

// info words written by the two parser kernels
enum : int { INFO_TOKENS = 0, INFO_STATUS = 1, INFO_FIRST_BAD = 2, INFO_EXACT = 3, INFO_WORDS = 4 };

// Pass 1: token starts. A token starts at a non-space byte that follows a space (or
// the beginning); fscanf("%Lg\n") skips any white space between numbers. starts[i]
// receives the byte offset of token i for i <= n (entry n, if present, is where the
// reference's file position would be after reading n numbers). text is padded to a
// multiple of 16 bytes with spaces. A tile is TOK_SUB x 4 KB of text (each thread
// takes one 16-byte vector of every 4 KB row), one look-back per tile.
constexpr int TOK_SUB = 8;
constexpr unsigned long long TOK_TILE_BYTES = (unsigned long long)TB * TOK_BYTES * TOK_SUB;

__global__ void __launch_bounds__(TB)
    k_text_tokenize(const unsigned char* __restrict__ text, unsigned long long len,
                    unsigned long long n, unsigned long long* __restrict__ starts,
                    unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket,
                    unsigned long long* __restrict__ info) {
  __shared__ uint32_t warp_sum[TOK_SUB][TB / 32];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const unsigned long long tile_pos = (unsigned long long)tile * TOK_TILE_BYTES;
  uint32_t mask[TOK_SUB];
  int incl[TOK_SUB];
#pragma unroll
  for (int r = 0; r < TOK_SUB; r++) {
    const unsigned long long pos = tile_pos + ((unsigned long long)r * TB + tid) * TOK_BYTES;
    uint32_t start_mask = 0;
    if (pos < len) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + pos));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t space_mask = 0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        if (is_space(c) || pos + j >= len) space_mask |= 1u << j;
      }
      const bool prev_space = pos == 0 ? true : is_space(__ldg(text + pos - 1));
      start_mask = ~space_mask & ((space_mask << 1) | (prev_space ? 1u : 0u)) & 0xffffu;
    }
    mask[r] = start_mask;
    int s = __popc(start_mask);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += v;
    }
    incl[r] = s;
    if (lane == 31) warp_sum[r][warp] = (uint32_t)s;
  }
  __syncthreads();
  // tokens are numbered by position: row r before row r + 1, within a row by thread
  uint32_t row_before[TOK_SUB];
  uint32_t tile_cnt = 0;
#pragma unroll
  for (int r = 0; r < TOK_SUB; r++) {
    uint32_t mine = tile_cnt;
#pragma unroll
    for (int w = 0; w < TB / 32; w++) {
      const uint32_t s = warp_sum[r][w];
      if (w < warp) mine += s;
      tile_cnt += s;
    }
    row_before[r] = mine;
  }
  if (warp == 0) {
    const unsigned long long excl = lookback_exclusive(status, tile, tile_cnt, lane);
    if (lane == 0) {
      s_base = excl;
      if (tile_pos + TOK_TILE_BYTES >= len) info[INFO_TOKENS] = excl + tile_cnt;
    }
  }
  __syncthreads();
  const unsigned long long base = s_base;
  if (base > n) return;  // only the first n + 1 starts are wanted
#pragma unroll
  for (int r = 0; r < TOK_SUB; r++) {
    uint32_t m = mask[r];
    unsigned long long idx = base + row_before[r] + (uint32_t)(incl[r] - __popc(m));
    const unsigned long long pos = tile_pos + ((unsigned long long)r * TB + tid) * TOK_BYTES;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      if (idx <= n) starts[idx] = pos + j;
      idx++;
    }
  }
}

// Pass 2: one thread per token.
__global__ void __launch_bounds__(TB)
    k_text_parse(const unsigned char* __restrict__ text, unsigned long long len,
                 unsigned long long n, const unsigned long long* __restrict__ starts,
                 const Pow10Entry* __restrict__ tab, ulonglong2* __restrict__ values,
                 unsigned long long* __restrict__ info, int force_band) {
  __shared__ uint32_t big[BIG_LIMBS];
  const int tid = threadIdx.x;
  const unsigned long long i = (unsigned long long)blockIdx.x * TB + tid;
  uint32_t undecided = 0;
  Decimal dec;
  uint64_t mant = 0;
  uint32_t se = 0;
  int q = 0;
  // starts[] is written for the tokens the tokenizer found only (same stream, earlier):
  // a truncated file leaves the tail of starts[] undefined, so those threads stay out
  const unsigned long long found = info[INFO_TOKENS];
  if (i < n && i < found && starts[i] < len) {
    const unsigned long long b = starts[i];
    int tlen = 0;
    uint32_t st = parse_number(text + b, (long)(len - b), &dec, &tlen);
    if (tlen >= MAX_TOKEN) st = PARSE_MALFORMED;
    if (st == PARSE_OK) {
      if (dec.special) {
        mant = dec.special == 2 ? (1ULL << 63) : (3ULL << 62);
        se = 0x7fffu;
      } else {
        const uint32_t r = decimal_to_x87(dec, tab, &mant, &se, &q, force_band != 0);
        if (r == 1) undecided = 1;
        if (r == 2) st = PARSE_UNSUPPORTED;
      }
    }
    if (st != PARSE_OK) {
      atomicMax(info + INFO_STATUS, (unsigned long long)st);
      atomicMin(info + INFO_FIRST_BAD, i);
    }
  }
  if (__syncthreads_or((int)undecided)) {
    for (int t = 0; t < TB; t++) {
      if (t == tid && undecided) {
        finish_exact(dec, mant, q, big, &mant, &se);
        atomicAdd(info + INFO_EXACT, 1ULL);
      }
      __syncthreads();
    }
  }
  if (i < n) values[i] = make_ulonglong2(mant, (unsigned long long)(se | (dec.neg << 15)));
}

}  // namespace text
}  // namespace qb200
