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

// key of sorted rank `rank` (0-based) among the warp-distributed keys
template <int R>
__device__ __forceinline__ uint64_t warp_select(const uint64_t (&key)[R], int rank) {
  // the high words of all real keys share a prefix (values of similar magnitude): skip those bits
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
  const uint32_t diff = hmin ^ hmax;
  const int top = diff ? (31 - __clz(diff)) : -1;  // highest differing bit
  uint32_t hi = top >= 31 ? 0u : (hmin & ~((top >= 0 ? (2u << top) : 1u) - 1u));
  if (top < 0) hi = hmin;
#pragma unroll 1
  for (int bit = top; bit >= 0; bit--) {
    uint32_t trial = hi | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += ((uint32_t)(key[r] >> 32) < trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= rank) hi = trial;
  }
  int below = 0, same = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    uint32_t h = (uint32_t)(key[r] >> 32);
    below += (h < hi) ? 1 : 0;
    same += (h == hi) ? 1 : 0;
  }
  below = __reduce_add_sync(0xffffffffu, below);
  same = __reduce_add_sync(0xffffffffu, same);
  uint32_t lo = 0;
  if (same == 1) {
    uint32_t mine = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
      if ((uint32_t)(key[r] >> 32) == hi) mine = (uint32_t)key[r];
    lo = __reduce_or_sync(0xffffffffu, mine);
  } else {
    const int rank_in = rank - below;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
      uint32_t trial = lo | (1u << bit);
      int c = 0;
#pragma unroll
      for (int r = 0; r < R; r++) c += ((uint32_t)(key[r] >> 32) == hi && (uint32_t)key[r] < trial) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c <= rank_in) lo = trial;
    }
  }
  return ((uint64_t)hi << 32) | lo;
}

// np.median of the `count` real keys (count >= 1, no NaN among them)
template <int R>
__device__ __forceinline__ double warp_median(const uint64_t (&key)[R], int count) {
  const int hi_rank = count >> 1;
  const uint64_t up = warp_select<R>(key, hi_rank);
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
