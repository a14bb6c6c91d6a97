// Null ratios of one target bin for one chunk of NR_CHUNK sample columns -- shared by the stand-alone kernel
// (null_ratios.cu) and the fused tail of the re-rank kernel (rerank.cu).  Reference loop: newref_tools.py:210-224.
#pragma once
#include "select.cuh"

namespace wcx {

constexpr int NR_CHUNK = 8;

// xm: [n][NR_CHUNK] sample values of this chunk; g: the bin's k indexes (lane-strided, -1 padded beyond k, negative
// reference indexes already wrapped); out_row: &out[row][first column of the chunk].  One warp.
template <int R>
__device__ __forceinline__ void null_row_chunk(const double* __restrict__ xm, const int32_t (&g)[R], int k, int mc, int64_t b,
                                               int lane, double* __restrict__ out_row) {
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double mymed = nan;  // lane j keeps the median of column j of the chunk
  for (int mp = 0; mp < mc; mp += 2) {
    uint64_t key0[R], key1[R];
    bool nan0 = false, nan1 = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (g[r] >= 0) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(xm + (int64_t)g[r] * NR_CHUNK + mp));
        nan0 |= (v.x != v.x);
        nan1 |= (v.y != v.y);
        key0[r] = dkey(v.x);
        key1[r] = dkey(v.y);
      } else {
        key0[r] = ~0ull;
        key1[r] = ~0ull;
      }
    }
    nan0 = __any_sync(0xffffffffu, nan0);
    nan1 = __any_sync(0xffffffffu, nan1);
    const double med0 = (nan0 || k == 0) ? nan : warp_median<R>(key0, k);
    if (lane == mp) mymed = med0;
    if (mp + 1 < mc) {
      const double med1 = (nan1 || k == 0) ? nan : warp_median<R>(key1, k);
      if (lane == mp + 1) mymed = med1;
    }
  }
  // one division + log2 per column, all columns of the chunk at once (lane j = column j), coalesced store
  if (lane < mc) out_row[lane] = log2(__ldg(xm + b * NR_CHUNK + lane) / mymed);
}

}  // namespace wcx
