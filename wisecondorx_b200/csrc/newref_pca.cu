// newref preparation kernels: PCA fit / correction and the PCA-distance bin filter.
//
// Reference functions replaced (file:line under src/wisecondorx/):
//   train_pca                 newref_tools.py:138-147   PCA(n_components=5).fit(X^T), corrected = X^T / inverse(transform(X^T))
//   PCA-distance filter       newref_control.py:40-46   per-sample median profile, per-bin squared distance,
//                                                       median + 10 * MAD cutoff (floor 5.0)
//
// sklearn's PCA resolves to a *randomized* SVD for these shapes; what it
// approximates is the exact model computed here: per-bin mean over samples, Gram matrix
// G = Xc^T Xc (S x S, float64) of the centred matrix, eigen-decomposition of G on the host
// (S <= a few thousand: LAPACK through NumPy, control plane), components = Xc U / sigma.
//   gram_kernel          float64 SYRK tiled 64 x 64 per block over a chunk of bins; partial tiles are
//                        reduced in a fixed order (deterministic)
//   pca_apply_kernel     one warp per bin: components[c, b] and corrected[b, :] in one pass over X
//   col_count_kernel     exact column medians by 64-step key bisection, one launch per step
//   row_sqdist_kernel    d_b = sum_s (corrected[b, s] - med_s)^2
#include "select.cuh"
#include "wcx_common.cuh"
#include "newref_pca.cuh"

namespace wcx {

namespace {

// ---- per-bin mean over samples (the PCA mean_) -------------------------------------------------
__global__ void row_mean_kernel(const double* __restrict__ x, int64_t n, int s, double* __restrict__ mean) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  double acc = 0.0;
  for (int c = lane; c < s; c += 32) acc += x[row * s + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mean[row] = acc / (double)s;
}

// ---- Gram matrix: partial[chunk][S_pad][S_pad] tiles of 64 x 64, 256 threads, 4 x 4 per thread ----
constexpr int GT = 64;
constexpr int GK = 16;  // bins per shared-memory step

__global__ void __launch_bounds__(256)
gram_kernel(const double* __restrict__ x, const double* __restrict__ mean, int64_t n, int s, int64_t bins_per_chunk,
            double* __restrict__ partial, int s_pad) {
  __shared__ double As[GK][GT + 1];
  __shared__ double Bs[GK][GT + 1];
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;  // symmetric: lower triangle of tiles only
  const int64_t b0 = (int64_t)blockIdx.z * bins_per_chunk;
  int64_t b1 = b0 + bins_per_chunk;
  if (b1 > n) b1 = n;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  for (int64_t b = b0; b < b1; b += GK) {
    // load GK bins x 64 samples for the row tile and the column tile (centred)
    for (int e = threadIdx.x; e < GK * GT; e += 256) {
      const int kb = e / GT, c = e % GT;
      const int64_t bin = b + kb;
      const int sa = ti * GT + c, sb = tj * GT + c;
      double va = 0.0, vb = 0.0;
      if (bin < b1) {
        const double m = mean[bin];
        if (sa < s) va = x[bin * s + sa] - m;
        if (sb < s) vb = x[bin * s + sb] - m;
      }
      As[kb][c] = va;
      Bs[kb][c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kb = 0; kb < GK; kb++) {
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = As[kb][ty * 4 + i]; bb[i] = Bs[kb][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * bb[j];
    }
    __syncthreads();
  }
  double* out = partial + (int64_t)blockIdx.z * s_pad * s_pad;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) out[(int64_t)(ti * GT + ty * 4 + i) * s_pad + tj * GT + tx * 4 + j] = acc[i][j];
}

__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nchunks, int s, int s_pad, double* __restrict__ g) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s || j >= s) return;
  const int a = i >= j ? i : j, b = i >= j ? j : i;  // lower triangle holds the data
  double acc = 0.0;
  for (int c = 0; c < nchunks; c++) acc += partial[(int64_t)c * s_pad * s_pad + (int64_t)a * s_pad + b];
  g[(int64_t)i * s + j] = acc;
}

// ---- components + correction: one warp per bin ----------------------------------------------------
// u: [S, ncomp] eigenvectors, sigma[ncomp] singular values.  comps[c, b] = sum_s u[s,c] xc[b,s] / sigma_c,
// corrected[b, s] = x[b, s] / (sum_c (xc_b . comps_c-direction) ... ) evaluated like the reference:
// transformed = (x - mean) . comps^T over bins needs all bins, but equals sigma_c * u[s, c]; the
// reconstruction is sum_c sigma_c u[s,c] comps[c,b] + mean_b.
__global__ void __launch_bounds__(256)
pca_apply_kernel(const double* __restrict__ x, const double* __restrict__ mean, int64_t n, int s, const double* __restrict__ u,
                 const double* __restrict__ sigma, int ncomp, double* __restrict__ comps, double* __restrict__ corrected) {
  extern __shared__ double us[];  // [S * ncomp]
  for (int e = threadIdx.x; e < s * ncomp; e += blockDim.x) us[e] = u[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double m = mean[row];
  double dot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = lane; c < s; c += 32) {
    const double v = x[row * s + c] - m;
    for (int k = 0; k < ncomp; k++) dot[k] += v * us[c * ncomp + k];
  }
  double cb[8];
  for (int k = 0; k < ncomp; k++) {
    double d = dot[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    cb[k] = d / sigma[k];
    if (lane == 0) comps[(int64_t)k * n + row] = cb[k];
  }
  for (int c = lane; c < s; c += 32) {
    double rec = 0.0;
    for (int k = 0; k < ncomp; k++) rec += (sigma[k] * us[c * ncomp + k]) * cb[k];
    rec += m;
    corrected[row * s + c] = x[row * s + c] / rec;
  }
}

// ---- exact column medians (np.median(x, axis=0)) by key bisection ---------------------------------
// state per column: res (current prefix key), cnt[step] counters.  Launch `step` first folds the
// count of step - 1 into res, then counts keys below the next trial.
__global__ void __launch_bounds__(1024)
col_count_kernel(const double* __restrict__ x, int64_t n, int s, int64_t rows_per_block, int step, int64_t rank,
                 unsigned long long* __restrict__ res, unsigned long long* __restrict__ cnt /*[65][S]*/) {
  __shared__ unsigned long long s_cnt[32][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  unsigned long long r = 0;
  if (col < s) {
    // replay the decisions of the previous steps (identical in every block)
    for (int k = 0; k < step; k++) {
      const unsigned long long trial = r | (1ull << (63 - k));
      if (cnt[(int64_t)k * s + col] <= (unsigned long long)rank) r = trial;
    }
    if (blockIdx.y == 0 && threadIdx.y == 0) res[col] = r;
  }
  if (step >= 64) return;
  const unsigned long long trial = r | (1ull << (63 - step));
  unsigned long long c = 0;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > n) r1 = n;
  if (col < s)
    for (int64_t row = r0 + threadIdx.y; row < r1; row += 32) c += (dkey(x[row * s + col]) < trial) ? 1 : 0;
  s_cnt[threadIdx.y][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.y == 0 && col < s) {
    for (int y = 1; y < 32; y++) c += s_cnt[y][threadIdx.x];
    atomicAdd(&cnt[(int64_t)step * s + col], c);
  }
}

// value at sorted rank for each column from the bisection result; for even n the mean of ranks n/2-1, n/2
__global__ void col_median_finish_kernel(const unsigned long long* __restrict__ res_hi, const unsigned long long* __restrict__ res_lo,
                                         int s, int even, double* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= s) return;
  const double hi = key_d(res_hi[col]);
  out[col] = even ? (key_d(res_lo[col]) + hi) / 2.0 : hi;
}

__global__ void row_sqdist_kernel(const double* __restrict__ x, int64_t n, int s, const double* __restrict__ med, double* __restrict__ d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  double acc = 0.0;
  for (int c = lane; c < s; c += 32) { const double t = x[row * s + c] - med[c]; acc += t * t; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) d[row] = acc;
}

// ---- normalize_and_mask (newref_tools.py:110-129): exact integer column totals, then x / total ----
__global__ void col_sum_i32_kernel(const int32_t* __restrict__ counts, int64_t bins, int s, int64_t rows_per_block,
                                   unsigned long long* __restrict__ colsum) {
  __shared__ unsigned long long sh[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > bins) r1 = bins;
  long long acc = 0;
  if (col < s)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += counts[r * s + col];
  sh[threadIdx.y][threadIdx.x] = (unsigned long long)acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < s) {
    unsigned long long t = 0;
    for (int y = 0; y < 8; y++) t += sh[y][threadIdx.x];
    atomicAdd(&colsum[col], t);  // integer: exact and order independent
  }
}

__global__ void normalize_mask_kernel(const int32_t* __restrict__ counts, int s, const int32_t* __restrict__ mask_pos, int64_t n,
                                      const unsigned long long* __restrict__ colsum, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * s) return;
  const int64_t row = e / s;
  const int col = (int)(e % s);
  out[e] = (double)counts[(int64_t)mask_pos[row] * s + col] / (double)(long long)colsum[col];
}

}  // namespace

int launch_normalize_and_mask(const int32_t* counts, int64_t bins_total, int32_t s, const int32_t* mask_pos, int64_t n,
                              unsigned long long* colsum, double* out, cudaStream_t st) {
  WCX_CUDA_OK(cudaMemsetAsync(colsum, 0, sizeof(unsigned long long) * s, st));
  if (bins_total == 0 || s == 0) return 0;
  const int64_t rows_per_block = 512;
  dim3 grid((s + 31) / 32, (unsigned)((bins_total + rows_per_block - 1) / rows_per_block));
  col_sum_i32_kernel<<<grid, dim3(32, 8), 0, st>>>(counts, bins_total, s, rows_per_block, colsum);
  if (n > 0) normalize_mask_kernel<<<(unsigned)((n * s + 255) / 256), 256, 0, st>>>(counts, s, mask_pos, n, colsum, out);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_row_mean(const double* x, int64_t n, int32_t s, double* mean, cudaStream_t st) {
  if (n == 0) return 0;
  row_mean_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, s, mean);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int gram_chunks(int64_t n) {
  int c = (int)((n + 4095) / 4096);
  return c < 1 ? 1 : (c > 64 ? 64 : c);
}
int gram_s_pad(int32_t s) { return (s + GT - 1) / GT * GT; }

int launch_gram(const double* x, const double* mean, int64_t n, int32_t s, double* partial, double* g, cudaStream_t st) {
  const int nch = gram_chunks(n), s_pad = gram_s_pad(s), nt = s_pad / GT;
  const int64_t per = (n + nch - 1) / nch;
  const int64_t per16 = (per + GK - 1) / GK * GK;
  dim3 grid(nt, nt, nch);
  WCX_CUDA_OK(cudaMemsetAsync(partial, 0, sizeof(double) * (size_t)nch * s_pad * s_pad, st));
  gram_kernel<<<grid, 256, 0, st>>>(x, mean, n, s, per16, partial, s_pad);
  dim3 g2((s + 127) / 128, s);
  gram_reduce_kernel<<<g2, 128, 0, st>>>(partial, nch, s, s_pad, g);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_pca_apply(const double* x, const double* mean, int64_t n, int32_t s, const double* u, const double* sigma,
                     int32_t ncomp, double* comps, double* corrected, cudaStream_t st) {
  if (ncomp > 8) { set_error("pca: more than 8 components unsupported"); return 1; }
  const size_t smem = sizeof(double) * (size_t)s * ncomp;
  if (smem > 200 * 1024) { set_error("pca: too many samples for the shared-memory eigenvector table"); return 1; }
  static size_t attr = 0;
  if (smem > attr) {
    WCX_CUDA_OK(cudaFuncSetAttribute(pca_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  pca_apply_kernel<<<(unsigned)((n + 7) / 8), 256, smem, st>>>(x, mean, n, s, u, sigma, ncomp, comps, corrected);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// medians of every column of x [n, s]; work: res_hi/res_lo [s] u64, cnt [65 * s] u64
int launch_col_medians(const double* x, int64_t n, int32_t s, unsigned long long* res_hi, unsigned long long* res_lo,
                       unsigned long long* cnt, double* out, cudaStream_t st) {
  if (n == 0 || s == 0) return 0;
  const int64_t rows_per_block = 2048;
  dim3 grid((s + 31) / 32, (unsigned)((n + rows_per_block - 1) / rows_per_block));
  const int even = (n % 2 == 0) ? 1 : 0;
  for (int pass = 0; pass < (even ? 2 : 1); pass++) {
    const int64_t rank = pass == 0 ? n / 2 : n / 2 - 1;
    unsigned long long* res = pass == 0 ? res_hi : res_lo;
    WCX_CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 65 * (size_t)s, st));
    for (int step = 0; step <= 64; step++)
      col_count_kernel<<<grid, dim3(32, 32), 0, st>>>(x, n, s, rows_per_block, step, rank, res, cnt);
  }
  col_median_finish_kernel<<<(s + 127) / 128, 128, 0, st>>>(res_hi, even ? res_lo : res_hi, s, even, out);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_row_sqdist(const double* x, int64_t n, int32_t s, const double* med, double* d, cudaStream_t st) {
  if (n == 0) return 0;
  row_sqdist_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, s, med, d);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
