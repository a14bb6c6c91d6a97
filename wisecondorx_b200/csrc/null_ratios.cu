// Null-ratio kernel ("K3").  Reference loop: newref_tools.py:210-224
//
//   for each chosen sample column `col` and each target bin b of the part:
//       r = log2(col[b] / np.median(col[indexes[b, :]]))
//
// Quirk reproduced: `indexes` are positions in the chromosome-EXCLUDED array but
// are applied to the FULL column; -1 wraps to the last bin (Python negative index).
//
// Layout: the chosen sample columns are copied chunk-wise into XM[chunk][n][8] (one 64-byte row of
// 8 sample values per bin), so one chunk (n * 64 bytes) stays L2-resident while the whole grid
// works on it and a gather of two adjacent samples is a single 16-byte load.
// One warp per target bin: the bin's k indexes live in registers (R = ceil(k / 32) per lane); per
// pair of samples the warp gathers k pairs of pre-computed keys (null_ratios.cuh) and selects the two middle order statistics
// of each with a bisection over order-preserving 64-bit keys (select.cuh) -- no sort.
// np.median returns NaN when any value is NaN.
#include "null_ratios.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {

// xm: [n][8] sample values of this chunk; idx: [rows, k]; out: [rows, m_total] at columns m_off..
template <int R>
__global__ void __launch_bounds__(256, R <= 10 ? 4 : 2)
null_ratios_kernel(const uint64_t* __restrict__ xm, int64_t n, const int32_t* __restrict__ idx, int64_t row_begin,
                   int64_t rows, int k, int mc, int m_off, int m_total, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t lrow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (lrow >= rows) return;
  int32_t g[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int t = r * 32 + lane;
    int64_t v = -1;
    if (t < k) {
      v = idx[lrow * k + t];
      if (v < 0) v += n;  // Python negative index
      v = v < 0 ? 0 : (v >= n ? n - 1 : v);  // device-resident lists are not validated by the host: never gather out of bounds
    }
    g[r] = (int32_t)v;
  }
  null_row_chunk<R>(xm, g, k, mc, row_begin + lrow, lane, out + lrow * m_total + m_off);
}

// XM[c][r][j] = X[r, ids[8c + j]] (zero padded past m); tile 32 rows per block, threads over (row, j)
__global__ void gather_cols_kernel(const double* __restrict__ x, int64_t n, int32_t s, const int32_t* __restrict__ ids,
                                   int32_t m, double* __restrict__ xm) {
  const int j = threadIdx.x & 7;
  const int64_t r = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int c = blockIdx.y;
  if (r >= n) return;
  const int mm = c * NR_CHUNK + j;
  reinterpret_cast<uint64_t*>(xm)[((int64_t)c * n + r) * NR_CHUNK + j] = null_key(mm < m ? x[r * s + ids[mm]] : 0.0);
}
}  // namespace

int launch_transpose_cols(const double* x, int64_t n, int32_t s, const int32_t* ids, int32_t m, double* xt,
                          cudaStream_t st) {
  if (n == 0 || m == 0) return 0;
  dim3 grid((unsigned)((n + 31) / 32), (m + NR_CHUNK - 1) / NR_CHUNK);
  gather_cols_kernel<<<grid, 256, 0, st>>>(x, n, s, ids, m, xt);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int64_t null_ratio_staging_doubles(int64_t n, int32_t m) { return (int64_t)((m + NR_CHUNK - 1) / NR_CHUNK) * n * NR_CHUNK; }

int launch_null_ratios(const double* xt, int64_t n, const int32_t* idx, int64_t row_begin, int64_t row_end,
                       int32_t k, int32_t m, double* out, cudaStream_t st) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || m <= 0) return 0;
  if (k > 512) { set_error("null_ratios: ref_size > 512 unsupported"); return 1; }
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  for (int m0 = 0; m0 < m; m0 += NR_CHUNK) {
    const int mc = m - m0 < NR_CHUNK ? m - m0 : NR_CHUNK;
    const uint64_t* xm = reinterpret_cast<const uint64_t*>(xt) + (int64_t)(m0 / NR_CHUNK) * n * NR_CHUNK;
    if (k <= 320)
      null_ratios_kernel<10><<<grid, warps * 32, 0, st>>>(xm, n, idx, row_begin, rows, k, mc, m0, m, out);
    else
      null_ratios_kernel<16><<<grid, warps * 32, 0, st>>>(xm, n, idx, row_begin, rows, k, mc, m0, m, out);
  }
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
