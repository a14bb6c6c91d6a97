// Null ratios of one target bin for one chunk of NR_CHUNK sample columns -- shared by the stand-alone kernel
// (null_ratios.cu) and the fused tail of the re-rank kernel (rerank.cu).  Reference loop: newref_tools.py:210-224.
#pragma once
#include "select.cuh"

namespace wcx {

constexpr int NR_CHUNK = 8;

// Key of a gathered sample value: the order-preserving integer of select.cuh, computed ONCE per (bin, column) by
// gather_cols_kernel instead of once per gather (300 x per value); NaN of either sign becomes NULL_NAN_KEY, whose
// high word is above that of every number (+inf: 0xfff00000), so the selection's own prefix scan reports it.
constexpr uint64_t NULL_NAN_KEY = 0xfffffffffffffffeull;  // ~0 is the padding sentinel of select.cuh
__device__ __forceinline__ uint64_t null_key(double v) { return v != v ? NULL_NAN_KEY : dkey(v); }

// xm: [n][NR_CHUNK] keys (null_key) of the sample values of this chunk; g: the bin's k indexes (lane-strided, -1
// padded beyond k, negative reference indexes already wrapped); out_row: &out[row][first column of the chunk].
// One warp.  np.median returns NaN when any value is NaN.
template <int R>
__device__ __forceinline__ void null_row_chunk(const uint64_t* __restrict__ xm, const int32_t (&g)[R], int k, int mc, int64_t b,
                                               int lane, double* __restrict__ out_row) {
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double mymed = nan;  // lane j keeps the median of column j of the chunk
  for (int mp = 0; mp < mc; mp += 2) {
    uint64_t key0[R], key1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (g[r] >= 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(xm + (int64_t)g[r] * NR_CHUNK + mp));
        key0[r] = ((uint64_t)v.y << 32) | v.x;
        key1[r] = ((uint64_t)v.w << 32) | v.z;
      } else {
        key0[r] = ~0ull;
        key1[r] = ~0ull;
      }
    }
    if (k > 0) {
      uint32_t hmax = 0;
      const double med0 = warp_median<R>(key0, k, &hmax);
      if (lane == mp) mymed = hmax == 0xffffffffu ? nan : med0;
      if (mp + 1 < mc) {
        const double med1 = warp_median<R>(key1, k, &hmax);
        if (lane == mp + 1) mymed = hmax == 0xffffffffu ? nan : med1;
      }
    }
  }
  // one division + log2 per column, all columns of the chunk at once (lane j = column j), coalesced store
  if (lane < mc) out_row[lane] = log2(key_d(__ldg(xm + b * NR_CHUNK + lane)) / mymed);
}

}  // namespace wcx
