// predict kernels: coverage normalisation, PCA projection, reference weights, optimal cutoff,
// the three within-sample normalisation passes and the between-sample segment z-score.
// All float64, all bandwidth-bound (HBM for the idx/dist rows, L2 for the gathers).
//
// Reference functions replaced (file:line under src/wisecondorx/):
//   coverage_normalize_and_mask  predict_tools.py:32-48
//   project_pc                   predict_tools.py:56-65
//   get_weights                  predict_tools.py:152-155
//   get_optimal_cutoff           predict_tools.py:74-82
//   normalize_repeat / _normalize_once   predict_tools.py:94-142
//   get_z_score                  overall_tools.py:88-119
#include "select.cuh"
#include "wcx_common.cuh"
#include "predict.cuh"

namespace wcx {

namespace {

constexpr int RED_BLOCKS = 592;  // 4 per SM; fixed so that reductions are run-to-run deterministic
constexpr int RED_THREADS = 256;
constexpr int PR_MAXK = 512;
constexpr int PR_R = PR_MAXK / 32;
constexpr double Z_MASK = 2.3263478740408408;  // scipy.stats.norm.ppf(0.99), predict_tools.py:104

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = l < (blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}

// ---- get_weights: 1 / mean(sqrt(row)) ------------------------------------------------------
__global__ void weights_kernel(const double* __restrict__ dist, int64_t n, int k, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double* d = dist + row * k;
  double acc = 0.0;
  for (int t = lane; t < k; t += 32) acc += sqrt(d[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = 1.0 / (acc / (double)k);
}

// ---- get_optimal_cutoff ---------------------------------------------------------------------
// state[0] = cutoff, state[1] = mean, state[2] = count
// phase 0: partial (count, sum) of d < cutoff ; phase 1: partial sum (d - mean)^2 of d < cutoff
__global__ void cutoff_partial_kernel(const double* __restrict__ dist, int64_t total, const double* __restrict__ state,
                                      int phase, double* __restrict__ partial) {
  __shared__ double sh[32];
  const double cutoff = state[0], mean = state[1];
  double a = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = dist[i];
    if (d < cutoff) {
      if (phase == 0) { a += d; c += 1.0; }
      else { const double t = d - mean; a += t * t; }
    }
  }
  double sa = block_sum(a, sh);
  double sc = block_sum(c, sh);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sa; partial[2 * blockIdx.x + 1] = sc; }
}

__global__ void cutoff_final_kernel(const double* __restrict__ partial, int nblocks, int phase, double* __restrict__ state) {
  __shared__ double sh[32];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += partial[2 * i]; c += partial[2 * i + 1]; }
  double sa = block_sum(a, sh);
  double sc = block_sum(c, sh);
  if (threadIdx.x == 0) {
    if (phase == 0) { state[2] = sc; state[1] = sa / sc; }
    else { state[0] = state[1] + 3.0 * sqrt(sa / state[2]); }
  }
}

// ---- coverage_normalize_and_mask -------------------------------------------------------------
__global__ void row_sum_partial_kernel(const double* __restrict__ raw, int64_t len, double* __restrict__ partial) {
  __shared__ double sh[32];
  const double* p = raw + (int64_t)blockIdx.y * len;
  double a = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) a += p[i];
  double s = block_sum(a, sh);
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}
__global__ void row_sum_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[(int64_t)blockIdx.x * nblocks + i];
  double s = block_sum(a, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}
__global__ void coverage_gather_kernel(const double* __restrict__ raw, int64_t bins_total, const int32_t* __restrict__ mask_pos,
                                       int64_t n, const double* __restrict__ totals, double* __restrict__ x) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[(int64_t)b * n + i] = raw[(int64_t)b * bins_total + mask_pos[i]] / totals[b];
}

// ---- project_pc --------------------------------------------------------------------------------
// t[b, c] = sum_i (x[b,i] - mu_i) * C[c,i]   (two-stage deterministic reduction), ncomp <= 8
__global__ void project_dots_partial_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ comps,
                                            const double* __restrict__ mean, int ncomp, double* __restrict__ partial) {
  __shared__ double sh[32];
  const int b = blockIdx.y;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[(int64_t)b * n + i] - mean[i];
    for (int c = 0; c < ncomp; c++) acc[c] += v * comps[(int64_t)c * n + i];
  }
  for (int c = 0; c < ncomp; c++) {
    double s = block_sum(acc[c], sh);
    if (threadIdx.x == 0) partial[((int64_t)b * gridDim.x + blockIdx.x) * 8 + c] = s;
  }
}
__global__ void project_dots_final_kernel(const double* __restrict__ partial, int nblocks, int ncomp, double* __restrict__ t) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  for (int c = 0; c < ncomp; c++) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[((int64_t)b * nblocks + i) * 8 + c];
    double s = block_sum(a, sh);
    if (threadIdx.x == 0) t[b * 8 + c] = s;
  }
}
__global__ void project_apply_kernel(double* __restrict__ x, int64_t n, const double* __restrict__ comps,
                                     const double* __restrict__ mean, int ncomp, const double* __restrict__ t) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rec = 0.0;
  for (int c = 0; c < ncomp; c++) rec += t[b * 8 + c] * comps[(int64_t)c * n + i];
  rec += mean[i];
  x[(int64_t)b * n + i] = x[(int64_t)b * n + i] / rec;
}

// ---- _normalize_once ------------------------------------------------------------------------------
// one warp per target bin i in [ct, n); loops over the B samples so idx/dist rows are read once.
__global__ void __launch_bounds__(256)
normalize_pass_kernel(const double* __restrict__ test_data, const double* __restrict__ copy_in, double* __restrict__ copy_out,
                      int B, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ dist, int k,
                      const double* __restrict__ cutoff_p, const int64_t* __restrict__ cum, int nchr, int64_t ct,
                      double* __restrict__ z_out, double* __restrict__ r_out, double* __restrict__ n_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ct + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const double cutoff = *cutoff_p;
  int c = 0;
  while (c < nchr && cum[c] <= i) c++;
  const int64_t cs = c == 0 ? 0 : cum[c - 1], ce = cum[c];
  const int64_t nc = ce - cs, nex = n - nc;
  int32_t g[PR_R];
#pragma unroll
  for (int r = 0; r < PR_R; r++) {
    const int t = r * 32 + lane;
    int32_t gg = -1;
    if (t < k && dist[i * k + t] < cutoff) {
      int64_t p = idx[i * k + t];
      if (p < 0) p += nex;                 // Python negative index into the chr-excluded array
      if (p >= 0 && p < nex) gg = (int32_t)(p < cs ? p : p + nc);
    }
    g[r] = gg;
  }
  const int64_t nout = n - ct;
  for (int b = 0; b < B; b++) {
    const double* cp = copy_in + (int64_t)b * n;
    uint64_t key[PR_R];
    double v[PR_R];
    int cnt = 0;
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < PR_R; r++) {
      double val = -1.0;
      if (g[r] >= 0) val = cp[g[r]];
      const bool keep = g[r] >= 0 && val >= 0.0;  // NaN and negatives (masked bins) dropped
      v[r] = keep ? val : 0.0;
      key[r] = keep ? dkey(val) : ~0ull;
      cnt += keep ? 1 : 0;
      sum += keep ? val : 0.0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double mean = nan, sd = nan, med = nan;
    if (cnt > 0) {
      mean = sum / (double)cnt;
      double ss = 0.0;
#pragma unroll
      for (int r = 0; r < PR_R; r++) {
        if (key[r] != ~0ull) { const double t = v[r] - mean; ss += t * t; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sd = sqrt(ss / (double)cnt);
      med = warp_median<PR_R>(key, cnt);
    }
    if (lane == 0) {
      const double x = test_data[(int64_t)b * n + i];
      const double z = (x - mean) / sd;
      z_out[(int64_t)b * nout + (i - ct)] = z;
      r_out[(int64_t)b * nout + (i - ct)] = x / med;
      n_out[(int64_t)b * nout + (i - ct)] = (double)cnt;
      if (copy_out) copy_out[(int64_t)b * n + i] = (fabs(z) >= Z_MASK) ? -1.0 : cp[i];
    }
  }
}

// ---- np.nanmedian over a row (optionally of log2) ---------------------------------------------------
// one block per row; bisection over order-preserving 64-bit keys, NaN excluded
__global__ void __launch_bounds__(1024)
nanmedian_kernel(const double* __restrict__ vals, int64_t len, int take_log2, double* __restrict__ out) {
  __shared__ unsigned long long s_cnt;
  __shared__ unsigned long long s_best;
  const double* p = vals + (int64_t)blockIdx.x * len;
  auto load = [&](int64_t i) -> double { double v = p[i]; return take_log2 ? log2(v) : v; };
  unsigned long long loc = 0;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) { double v = load(i); loc += (v == v) ? 1 : 0; }
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  atomicAdd(&s_cnt, loc);
  __syncthreads();
  const long long valid = (long long)s_cnt;
  if (valid == 0) {
    if (threadIdx.x == 0) out[blockIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const long long hi_rank = valid >> 1;
  uint64_t T = 0;
  for (int bit = 63; bit >= 0; bit--) {
    const uint64_t trial = T | (1ull << bit);
    loc = 0;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
      double v = load(i);
      if (v == v) loc += (dkey(v) < trial) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    atomicAdd(&s_cnt, loc);
    __syncthreads();
    if ((long long)s_cnt <= hi_rank) T = trial;
  }
  double upper = key_d(T);
  double med = upper;
  if ((valid & 1) == 0) {
    // lower middle
    unsigned long long best = 0;
    loc = 0;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
      double v = load(i);
      if (v == v) {
        uint64_t kk = dkey(v);
        if (kk < T) { loc++; best = kk > best ? kk : best; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_cnt = 0; s_best = 0; }
    __syncthreads();
    atomicAdd(&s_cnt, loc);
    atomicMax(&s_best, best);
    __syncthreads();
    double lower = ((long long)s_cnt == hi_rank) ? key_d((uint64_t)s_best) : upper;
    med = (lower + upper) / 2.0;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = med;
}

// ---- get_z_score ---------------------------------------------------------------------------------------
// one block per segment: 128 null columns x 2 bin lanes.  M <= 128 (the reference uses
// min(S, 100) null samples, newref_tools.py:211).
__global__ void __launch_bounds__(256)
segment_z_kernel(const double* __restrict__ nr, int m, const int32_t* __restrict__ inflate_pos, const double* __restrict__ r,
                 const double* __restrict__ w, const int64_t* __restrict__ seg_se, const double* __restrict__ seg_r,
                 double* __restrict__ z_out) {
  __shared__ double s_num[2][128];
  __shared__ double s_den[2][128];
  __shared__ int s_any[2][128];
  __shared__ double s_val[128];
  __shared__ double sh[32];
  __shared__ double s_mean, s_count;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int64_t s = seg_se[2 * blockIdx.x], e = seg_se[2 * blockIdx.x + 1];
  double num = 0.0, den = 0.0;
  int any = 0;
  if (tx < m) {
    for (int64_t b = s + ty; b < e; b += 2) {
      if (r[b] == 0.0) continue;  // bins without data are dropped (overall_tools.py:98-100,106)
      const int32_t p = inflate_pos[b];
      if (p < 0) continue;
      const double v = nr[(int64_t)p * m + tx];
      if (isfinite(v)) { num += w[b] * v; den += w[b]; any = 1; }  // non-finite entries are masked (:101-108)
    }
  }
  s_num[ty][tx] = num; s_den[ty][tx] = den; s_any[ty][tx] = any;
  __syncthreads();
  if (ty == 0) {
    double val = nan;
    if (tx < m && (s_any[0][tx] | s_any[1][tx])) val = (s_num[0][tx] + s_num[1][tx]) / (s_den[0][tx] + s_den[1][tx]);
    s_val[tx] = val;
  }
  __syncthreads();
  const bool ok = (ty == 0 && tx < m && isfinite(s_val[tx]));
  const double sum_v = block_sum(ok ? s_val[tx] : 0.0, sh);
  const double sum_c = block_sum(ok ? 1.0 : 0.0, sh);
  if (threadIdx.x == 0) { s_mean = sum_v / sum_c; s_count = sum_c; }
  __syncthreads();
  const double mean = s_mean, cnt = s_count;
  const double dv = ok ? (s_val[tx] - mean) : 0.0;
  const double ss = block_sum(dv * dv, sh);
  if (threadIdx.x == 0) {
    double z = nan;
    if (cnt > 0.0) {
      const double sd = sqrt(ss / cnt);
      z = (seg_r[blockIdx.x] - mean) / sd;
      if (z == z) { z = z > 1000.0 ? 1000.0 : z; z = z < -1000.0 ? -1000.0 : z; }  // min/max clip (:114-115)
    }
    z_out[blockIdx.x] = z;
  }
}

}  // namespace

// =================================================================================================
// host launchers
// =================================================================================================
int predict_red_blocks() { return RED_BLOCKS; }

int launch_weights(const double* dist, int64_t n, int32_t k, double* out, cudaStream_t st) {
  if (n == 0) return 0;
  weights_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(dist, n, k, out);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// state3 (device): [cutoff, mean, count]; partial (device): [RED_BLOCKS * 2]
int launch_optimal_cutoff(const double* dist, int64_t total, int32_t repeats, double* state3, double* partial,
                          cudaStream_t st) {
  static const double init[3] = {__builtin_inf(), 0.0, 0.0};
  WCX_CUDA_OK(cudaMemcpyAsync(state3, init, sizeof(init), cudaMemcpyHostToDevice, st));
  for (int i = 0; i < repeats; i++) {
    for (int phase = 0; phase < 2; phase++) {
      cutoff_partial_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(dist, total, state3, phase, partial);
      cutoff_final_kernel<<<1, 256, 0, st>>>(partial, RED_BLOCKS, phase, state3);
    }
  }
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// raw [B, bins_total] -> x [B, n] = project_pc(coverage_normalize_and_mask(raw))
// partial: [B * RED_BLOCKS * 8], totals: [B], tdots: [B * 8]
int launch_coverage_project(const double* raw, int32_t B, int64_t bins_total, const int32_t* mask_pos, int64_t n,
                            const double* comps, const double* mean, int32_t ncomp, double* x, double* partial,
                            double* totals, double* tdots, cudaStream_t st) {
  if (ncomp > 8) { set_error("project_pc: more than 8 components unsupported"); return 1; }
  dim3 g(RED_BLOCKS, B);
  row_sum_partial_kernel<<<g, RED_THREADS, 0, st>>>(raw, bins_total, partial);
  row_sum_final_kernel<<<B, 256, 0, st>>>(partial, RED_BLOCKS, totals);
  dim3 g2((unsigned)((n + 255) / 256), B);
  coverage_gather_kernel<<<g2, 256, 0, st>>>(raw, bins_total, mask_pos, n, totals, x);
  project_dots_partial_kernel<<<g, RED_THREADS, 0, st>>>(x, n, comps, mean, ncomp, partial);
  project_dots_final_kernel<<<B, 256, 0, st>>>(partial, RED_BLOCKS, ncomp, tdots);
  project_apply_kernel<<<g2, 256, 0, st>>>(x, n, comps, mean, ncomp, tdots);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// three passes of _normalize_once with the -1 masking in between (normalize_repeat), then the two
// nanmedians.  copy_a / copy_b: [B, n] ping-pong buffers.
int launch_normalize_repeat(const double* x, double* copy_a, double* copy_b, int32_t B, int64_t n, const int32_t* idx,
                            const double* dist, int32_t k, const double* cutoff_dev, const int64_t* cum_dev,
                            int32_t nchr, int64_t ct, double* z, double* r, double* nref, double* m_lr, double* m_z,
                            cudaStream_t st) {
  if (k > PR_MAXK) { set_error("normalize: ref_size > 512 unsupported"); return 1; }
  const int64_t nout = n - ct;
  if (nout <= 0 || B <= 0) return 0;
  WCX_CUDA_OK(cudaMemcpyAsync(copy_a, x, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToDevice, st));
  WCX_CUDA_OK(cudaMemcpyAsync(copy_b, x, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToDevice, st));
  const unsigned grid = (unsigned)((nout + 7) / 8);
  double* in = copy_a;
  double* out = copy_b;
  for (int pass = 0; pass < 3; pass++) {
    normalize_pass_kernel<<<grid, 256, 0, st>>>(x, in, pass < 2 ? out : nullptr, B, n, idx, dist, k, cutoff_dev, cum_dev,
                                                nchr, ct, z, r, nref);
    double* t = in; in = out; out = t;
  }
  nanmedian_kernel<<<B, 1024, 0, st>>>(r, nout, 1, m_lr);
  nanmedian_kernel<<<B, 1024, 0, st>>>(z, nout, 0, m_z);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_segment_z(const double* nr, int32_t m, const int32_t* inflate_pos, const double* r, const double* w,
                     const int64_t* seg_se, const double* seg_r, int32_t nseg, double* z_out, cudaStream_t st) {
  if (nseg <= 0) return 0;
  if (m > 128) { set_error("get_z_score: more than 128 null samples unsupported"); return 1; }
  segment_z_kernel<<<nseg, 256, 0, st>>>(nr, m, inflate_pos, r, w, seg_se, seg_r, z_out);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
