// Preparation kernels for the newref distance sweep (reference: newref_tools.py:192-203, the
// `chr_data - row` operand of get_ref_for_bins).  All bandwidth-bound single passes over X.
//
//  * col_stats       per-sample (column) mean over finite entries: the centring vector.  Squared
//                    distances are invariant under a per-column shift, and centring removes the
//                    catastrophic cancellation of |a|^2 + |b|^2 - 2ab for data around 1.0.
//  * center_round    Xc = tf32_round(float(X - mean)), zero padded to [n_pad, k_pad], plus the
//                    row norms of the ROUNDED values (so the tensor-core dot products of the
//                    rounded operands are exact products; only accumulation rounds).
//  * center_round_f16  the same for the f16 sweep: Xh = half((X - mean) * 2^e), e chosen so that every finite
//                    entry stays below 2^14; f16 and tf32 both carry 11 significant bits, so the error
//                    bound of the sweep is unchanged while the tensor pipe runs at twice the rate on
//                    half the operand bytes.
#include <cuda_fp16.h>

#include "wcx_common.cuh"

namespace wcx {

__global__ void col_stats_kernel(const double* __restrict__ x, int64_t n, int32_t s, int64_t rows_per_block,
                                 double* __restrict__ colsum, double* __restrict__ colcnt,
                                 unsigned long long* __restrict__ absmax) {
  __shared__ double ssum[8][33];
  __shared__ double scnt[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > n) r1 = n;
  double acc = 0.0, cnt = 0.0, amax = 0.0;
  if (col < s) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      double v = x[r * s + col];
      if (isfinite(v)) { acc += v; cnt += 1.0; amax = fmax(amax, fabs(v)); }
    }
  }
  // non-negative doubles order like their bit patterns
  amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
  amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
  amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
  amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  if (threadIdx.x == 0 && amax > 0.0) atomicMax(absmax, (unsigned long long)__double_as_longlong(amax));
  ssum[threadIdx.y][threadIdx.x] = acc;
  scnt[threadIdx.y][threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.y == 0 && col < s) {
    for (int y = 1; y < 8; y++) { acc += ssum[y][threadIdx.x]; cnt += scnt[y][threadIdx.x]; }
    atomicAdd(&colsum[col], acc);
    atomicAdd(&colcnt[col], cnt);
  }
}

int launch_col_stats(const double* x, int64_t n, int32_t s, double* colsum, double* colcnt, unsigned long long* absmax,
                     cudaStream_t st) {
  WCX_CUDA_OK(cudaMemsetAsync(absmax, 0, sizeof(unsigned long long), st));
  WCX_CUDA_OK(cudaMemsetAsync(colsum, 0, sizeof(double) * s, st));
  WCX_CUDA_OK(cudaMemsetAsync(colcnt, 0, sizeof(double) * s, st));
  if (n == 0 || s == 0) return 0;
  int64_t rows_per_block = 512;
  dim3 grid((s + 31) / 32, (unsigned)((n + rows_per_block - 1) / rows_per_block));
  col_stats_kernel<<<grid, dim3(32, 8), 0, st>>>(x, n, s, rows_per_block, colsum, colcnt, absmax);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// one warp per row
__global__ void center_round_kernel(const double* __restrict__ x, int64_t n, int32_t s,
                                    const double* __restrict__ colsum, const double* __restrict__ colcnt,
                                    float* __restrict__ xc, float* __restrict__ norm, int64_t n_pad, int32_t k_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float* out = xc + row * k_pad;
  double acc = 0.0;
  for (int c = lane; c < k_pad; c += 32) {
    float v = 0.f;
    if (row < n && c < s) {
      double cnt = colcnt[c];
      double mean = cnt > 0.0 ? colsum[c] / cnt : 0.0;
      v = round_tf32((float)(x[row * s + c] - mean));
    }
    out[c] = v;
    acc += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) norm[row] = (float)acc;
}

int launch_center_round(const double* x, int64_t n, int32_t s, const double* colsum, const double* colcnt,
                        float* xc, float* norm, int64_t n_pad, int32_t k_pad, cudaStream_t st) {
  if (n_pad == 0) return 0;
  const int warps = 8;
  unsigned grid = (unsigned)((n_pad + warps - 1) / warps);
  center_round_kernel<<<grid, warps * 32, 0, st>>>(x, n, s, colsum, colcnt, xc, norm, n_pad, k_pad);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// scale[0] = 2^e with (max|x| + max|mean|) * 2^e < 2^14 (|x - mean| cannot exceed that bound), scale[1] = 4^e
__global__ void prep_scale_kernel(const double* __restrict__ colsum, const double* __restrict__ colcnt, int32_t s,
                                  const unsigned long long* __restrict__ absmax, double* __restrict__ scale) {
  __shared__ double red[8];
  double m = 0.0;
  for (int c = threadIdx.x; c < s; c += blockDim.x) {
    const double cnt = colcnt[c];
    if (cnt > 0.0) m = fmax(m, fabs(colsum[c] / cnt));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmax(m, red[w]);
    const double bound = m + __longlong_as_double((long long)*absmax);
    int q = 0;
    if (bound > 0.0 && isfinite(bound)) frexp(bound, &q);  // bound = f * 2^q, f in [0.5, 1)
    int e = 14 - q;
    e = e > 400 ? 400 : (e < -400 ? -400 : e);
    scale[0] = ldexp(1.0, e);
    scale[1] = ldexp(1.0, 2 * e);
  }
}

// one warp per row
__global__ void center_round_f16_kernel(const double* __restrict__ x, int64_t n, int32_t s,
                                        const double* __restrict__ colsum, const double* __restrict__ colcnt,
                                        const double* __restrict__ scale, __half* __restrict__ xh,
                                        float* __restrict__ norm, float2* __restrict__ normres,
                                        unsigned int* __restrict__ rho_bits, double tau, int64_t n_pad, int32_t k_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  const double sc = scale[0];
  __half* out = xh + row * k_pad;
  double acc = 0.0, racc = 0.0;
  for (int c = lane; c < k_pad; c += 32) {
    __half h = __ushort_as_half((unsigned short)0);
    double t = 0.0;
    if (row < n && c < s) {
      const double cnt = colcnt[c];
      const double mean = cnt > 0.0 ? colsum[c] / cnt : 0.0;
      t = (x[row * s + c] - mean) * sc;
      h = __double2half(t);  // one rounding, to nearest even
    }
    out[c] = h;
    const double v = (double)__half2float(h);
    acc += v * v;
    racc += (t - v) * (t - v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    racc += __shfl_xor_sync(0xffffffffu, racc, o);
  }
  if (lane == 0) {
    norm[row] = (float)acc;
    // measured rounding residual |x - x^| of the row (rerank.cu: approx_eps_meas), rounded up; the largest ratio
    // (|x - x^| - tau) / |x^| over the finite rows bounds the residual of a candidate known only by its norm
    const double r = sqrt(racc) * (1.0 + 1e-9);
    const float rf = __double2float_ru(r);
    normres[row] = make_float2((float)acc, rf);
    if (row < n && acc > 0.0 && isfinite(acc) && isfinite(r)) {
      const double rho = fmax(r - tau, 0.0) / sqrt(acc) * (1.0 + 1e-6);
      atomicMax(rho_bits, __float_as_uint(__double2float_ru(rho)));  // non-negative floats order like their bit patterns
    }
  }
}

int launch_center_round_f16(const double* x, int64_t n, int32_t s, const double* colsum, const double* colcnt,
                            const unsigned long long* absmax, double* scale, void* xh, float* norm, float2* normres,
                            float* rho_max, double tau, int64_t n_pad, int32_t k_pad, cudaStream_t st) {
  if (n_pad == 0) return 0;
  WCX_CUDA_OK(cudaMemsetAsync(rho_max, 0, sizeof(float), st));
  prep_scale_kernel<<<1, 256, 0, st>>>(colsum, colcnt, s, absmax, scale);
  const int warps = 8;
  unsigned grid = (unsigned)((n_pad + warps - 1) / warps);
  center_round_f16_kernel<<<grid, warps * 32, 0, st>>>(x, n, s, colsum, colcnt, scale, reinterpret_cast<__half*>(xh), norm, normres,
                                                       reinterpret_cast<unsigned int*>(rho_max), tau, n_pad, k_pad);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
