// Warp-cooperative exact order statistics over up to 32*R float64 values held in registers
// (R per lane) -- used for np.median in the null-ratio and within-sample normalisation kernels.
// No sort: a 32+32-bit bisection over order-preserving integer keys; padding entries are ~0.
#pragma once
#include <stdint.h>

namespace wcx {

__device__ __forceinline__ uint64_t dkey(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_d(uint64_t k) {
  uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// key of sorted rank `rank` (0-based) among the `count` real keys distributed over the warp
// (padding entries are ~0).  Bisection over the high 32 bits with two shortcuts: the bits shared
// by all keys are skipped, and as soon as the bracket holds a single key that key is fetched
// directly (about log2(count) + 2 iterations instead of 32 + 32).
// hmax_out (optional): largest high word among the real keys -- lets a caller that maps NaN to a key with high
// word 0xffffffff detect it without a pass of its own.
template <int R>
__device__ __forceinline__ uint64_t warp_select(const uint64_t (&key)[R], int rank, int count, uint32_t* hmax_out = nullptr) {
  uint32_t hmin = 0xffffffffu, hmax = 0u;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (key[r] != ~0ull) {
      const uint32_t h = (uint32_t)(key[r] >> 32);
      hmin = h < hmin ? h : hmin;
      hmax = h > hmax ? h : hmax;
    }
  }
  hmin = __reduce_min_sync(0xffffffffu, hmin);
  hmax = __reduce_max_sync(0xffffffffu, hmax);
  if (hmax_out) *hmax_out = hmax;
  const uint32_t diff = hmin ^ hmax;
  const int top = diff ? (31 - __clz(diff)) : -1;  // highest differing bit
  uint32_t hi = top >= 31 ? 0u : (hmin & ~((top >= 0 ? (2u << top) : 1u) - 1u));
  if (top < 0) hi = hmin;
  int below = 0;      // keys with high word < hi
  int inb = count;    // keys inside the current bracket [hi, hi + 2^(bit + 1))
  int bit = top;
#pragma unroll 1
  for (; bit >= 0 && inb > 1; bit--) {
    const uint32_t trial = hi | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      // compare + predicated add: two instructions per key (the C form compiles to three)
      asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "r"((uint32_t)(key[r] >> 32)), "r"(trial));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= rank) { hi = trial; inb = below + inb - c; below = c; }
    else { inb = c - below; }
  }
  if (inb == 1) {
    // the bracket [hi, hi + 2^(bit + 1)) holds exactly the wanted key: fetch it
    const int sh = bit + 1;  // number of still-unknown low bits of the high word
    uint32_t mh = 0, ml = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t h = (uint32_t)(key[r] >> 32);
      const bool in = key[r] != ~0ull && (sh >= 32 ? true : ((h ^ hi) >> sh) == 0u);
      if (in) { mh = h; ml = (uint32_t)key[r]; }
    }
    mh = __reduce_or_sync(0xffffffffu, mh);
    ml = __reduce_or_sync(0xffffffffu, ml);
    return ((uint64_t)mh << 32) | ml;
  }
  // several keys share the full high word: resolve on the low word among them
  uint32_t lo = 0;
  const int rank_in = rank - below;
#pragma unroll 1
  for (int b2 = 31; b2 >= 0; b2--) {
    const uint32_t trial = lo | (1u << b2);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += ((uint32_t)(key[r] >> 32) == hi && (uint32_t)key[r] < trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= rank_in) lo = trial;
  }
  return ((uint64_t)hi << 32) | lo;
}

// np.median of the `count` real keys (count >= 1, no NaN among them)
template <int R>
__device__ __forceinline__ double warp_median(const uint64_t (&key)[R], int count, uint32_t* hmax_out = nullptr) {
  const int hi_rank = count >> 1;
  const uint64_t up = warp_select<R>(key, hi_rank, count, hmax_out);
  const double upper = key_d(up);
  if (count & 1) return upper;
  int c_lt = 0;
  uint64_t best = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (key[r] < up) { c_lt++; best = key[r] > best ? key[r] : best; }
  }
  c_lt = __reduce_add_sync(0xffffffffu, c_lt);
  const uint32_t bh = __reduce_max_sync(0xffffffffu, (uint32_t)(best >> 32));
  const uint32_t bl = __reduce_max_sync(0xffffffffu, ((uint32_t)(best >> 32) == bh) ? (uint32_t)best : 0u);
  const double lower = (c_lt == hi_rank) ? key_d(((uint64_t)bh << 32) | bl) : upper;
  return (lower + upper) / 2.0;  // np.mean of the two middle values
}

}  // namespace wcx
