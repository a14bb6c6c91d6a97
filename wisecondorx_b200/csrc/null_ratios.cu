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
#include "select.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {
constexpr int NR_MAXK = 512;
constexpr int NR_R = NR_MAXK / 32;

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
      med = warp_median<NR_R>(key, k);
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
