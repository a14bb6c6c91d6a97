// Null-ratio kernel ("K3").  Reference loop: newref_tools.py:210-224
//
//   for each chosen sample column `col` and each target bin b of the part:
//       r = log2(col[b] / np.median(col[indexes[b, :]]))
//
// Quirk reproduced: `indexes` are positions in the chromosome-EXCLUDED array but
// are applied to the FULL column; -1 wraps to the last bin (Python negative index).
//
// One warp per target bin: the warp keeps the bin's k indexes in registers (read once per
// m-chunk), gathers the k column values from the column-contiguous copy XT[m, :] (L2-resident:
// one column is N * 8 bytes), and selects the two middle order statistics with a 32+32-bit
// bisection over orderable keys -- no sort.  np.median returns NaN when any value is NaN.
#include "wcx_common.cuh"

namespace wcx {

namespace {
constexpr int NR_MAXK = 512;
constexpr int NR_R = NR_MAXK / 32;

__device__ __forceinline__ uint64_t dkey(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_d(uint64_t k) {
  uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// value of sorted rank `rank` (0-based) among the warp-distributed keys (padding = ~0)
__device__ __forceinline__ uint64_t warp_select(const uint64_t (&key)[NR_R], int rank) {
  // high 32 bits
  uint32_t hi = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; bit--) {
    uint32_t trial = hi | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < NR_R; r++) c += ((uint32_t)(key[r] >> 32) < trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= rank) hi = trial;
  }
  // hi = high word of the rank-th key.  Entries sharing it:
  int below = 0, same = 0;
#pragma unroll
  for (int r = 0; r < NR_R; r++) {
    uint32_t h = (uint32_t)(key[r] >> 32);
    below += (h < hi) ? 1 : 0;
    same += (h == hi) ? 1 : 0;
  }
  below = __reduce_add_sync(0xffffffffu, below);
  same = __reduce_add_sync(0xffffffffu, same);
  uint32_t lo = 0;
  if (same == 1) {
    // the unique holder broadcasts its low word
    uint32_t mine = 0;
#pragma unroll
    for (int r = 0; r < NR_R; r++)
      if ((uint32_t)(key[r] >> 32) == hi) mine = (uint32_t)key[r];
    lo = __reduce_or_sync(0xffffffffu, mine);
  } else {
    const int rank_in = rank - below;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
      uint32_t trial = lo | (1u << bit);
      int c = 0;
#pragma unroll
      for (int r = 0; r < NR_R; r++)
        c += ((uint32_t)(key[r] >> 32) == hi && (uint32_t)key[r] < trial) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c <= rank_in) lo = trial;
    }
  }
  return ((uint64_t)hi << 32) | lo;
}
}  // namespace

// xt: [m, n] column copies; idx: [rows, k]; out: [rows, m_total] written at columns m_off..m_off+m
__global__ void __launch_bounds__(256)
null_ratios_kernel(const double* __restrict__ xt, int64_t n, const int32_t* __restrict__ idx, int64_t row_begin,
                   int64_t rows, int k, int m, int m_off, int m_total, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t lrow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (lrow >= rows) return;
  int64_t g[NR_R];
#pragma unroll
  for (int r = 0; r < NR_R; r++) {
    int t = r * 32 + lane;
    int64_t v = -1;
    if (t < k) {
      v = idx[lrow * k + t];
      if (v < 0) v += n;  // Python negative index
    }
    g[r] = (t < k) ? v : -1;
  }
  const int64_t b = row_begin + lrow;
  const int hi_rank = k >> 1;  // upper middle (0-based); lower middle = hi_rank - 1 for even k
  for (int mm = 0; mm < m; mm++) {
    const double* col = xt + (int64_t)mm * n;
    uint64_t key[NR_R];
    bool has_nan = false;
#pragma unroll
    for (int r = 0; r < NR_R; r++) {
      if (g[r] >= 0) {
        double v = col[g[r]];
        has_nan |= (v != v);
        key[r] = dkey(v);
      } else {
        key[r] = ~0ull;
      }
    }
    has_nan = __any_sync(0xffffffffu, has_nan);
    double med;
    if (has_nan || k == 0) {
      med = __longlong_as_double(0x7ff8000000000000ll);
    } else {
      uint64_t up = warp_select(key, hi_rank);
      double upper = key_d(up);
      if (k & 1) {
        med = upper;
      } else {
        // lower middle: largest key strictly below `up` unless duplicates of `up` cover rank-1
        int c_lt = 0;
        uint64_t best = 0;
#pragma unroll
        for (int r = 0; r < NR_R; r++) {
          if (key[r] < up) { c_lt++; best = key[r] > best ? key[r] : best; }
        }
        c_lt = __reduce_add_sync(0xffffffffu, c_lt);
        uint32_t bh = __reduce_max_sync(0xffffffffu, (uint32_t)(best >> 32));
        uint32_t bl = __reduce_max_sync(0xffffffffu, ((uint32_t)(best >> 32) == bh) ? (uint32_t)best : 0u);
        double lower = (c_lt == hi_rank) ? key_d(((uint64_t)bh << 32) | bl) : upper;
        med = (lower + upper) / 2.0;  // np.mean of the two middle values
      }
    }
    if (lane == 0) out[lrow * m_total + m_off + mm] = log2(col[b] / med);
  }
}

int launch_null_ratios(const double* xt, int64_t n, const int32_t* idx, int64_t row_begin, int64_t row_end,
                       int32_t k, int32_t m, double* out, cudaStream_t st) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || m <= 0) return 0;
  if (k > NR_MAXK) { set_error("null_ratios: ref_size > 512 unsupported"); return 1; }
  const int warps = 8;
  unsigned grid = (unsigned)((rows + warps - 1) / warps);
  // chunks of 8 sample columns keep the gathered columns L2-resident across the grid
  const int chunk = 8;
  for (int m0 = 0; m0 < m; m0 += chunk) {
    int mc = m - m0 < chunk ? m - m0 : chunk;
    null_ratios_kernel<<<grid, warps * 32, 0, st>>>(xt + (int64_t)m0 * n, n, idx, row_begin, rows, k, mc, m0, m, out);
  }
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
