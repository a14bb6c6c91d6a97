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
#include <algorithm>
#include <cstdlib>

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
// Which reference bins a target bin uses depends only on the reference set and the cutoff (predict_tools.py:117-131:
// distances below the cutoff, chromosome-excluded position -> bin of the full vector), not on the sample or the pass:
// gather_list_kernel resolves it ONCE per (reference set, cutoff) into gl[i - ct][t] = global bin or -1, and the three
// passes of every sample read 4 bytes per reference bin instead of 12 (+ the position arithmetic per pass).
__global__ void __launch_bounds__(256)
gather_list_kernel(const int32_t* __restrict__ idx, const double* __restrict__ dist, int64_t n, int k,
                   const double* __restrict__ cutoff_p, const int64_t* __restrict__ cum, int nchr, int64_t ct,
                   int32_t* __restrict__ gl) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ct + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const double cutoff = *cutoff_p;
  int c = 0;
  while (c < nchr && cum[c] <= i) c++;
  const int64_t cs = c == 0 ? 0 : cum[c - 1], ce = cum[c];
  const int64_t nc = ce - cs, nex = n - nc;
  for (int t = lane; t < k; t += 32) {
    int32_t gg = -1;
    if (dist[i * k + t] < cutoff) {
      int64_t p = idx[i * k + t];
      if (p < 0) p += nex;                 // Python negative index into the chr-excluded array
      if (p >= 0 && p < nex) gg = (int32_t)(p < cs ? p : p + nc);
    }
    gl[(i - ct) * k + t] = gg;
  }
}

// one warp per target bin i in [ct, n); loops over the B samples so the gather list is read once per batch.
// (A thread-per-(bin, sample) variant on packed 15-bit codes -- the design of the null-ratio kernel -- was built and
// measured in round 2: 3.0 ms for the three passes of one sample against 1.1 ms here, 158 ms against 111 ms at batch 96;
// with one thread per bin the gather lists and the two value passes are uncoalesced.  Not kept.)
// R = register slots per lane (R / 4 quads of 4 consecutive list entries, 16-byte loads when k % 4 == 0).
// The kept values are >= 0, so their order-preserving key is the bit pattern with the sign bit set: the value is
// recovered from the key and needs no registers of its own.
template <int R>
__global__ void __launch_bounds__(256)
normalize_pass_kernel(const double* __restrict__ test_data, const double* __restrict__ copy_in, double* __restrict__ copy_out,
                      int B, int64_t n, const int32_t* __restrict__ gl, int k, int64_t ct,
                      double* __restrict__ z_out, double* __restrict__ r_out, double* __restrict__ n_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ct + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  int32_t g[R];
  const int32_t* grow = gl + (i - ct) * k;
  if ((k & 3) == 0) {
#pragma unroll
    for (int j = 0; j < R / 4; j++) {
      const int q = j * 32 + lane;
      int4 v = make_int4(-1, -1, -1, -1);
      if (4 * q < k) v = __ldg(reinterpret_cast<const int4*>(grow) + q);
      g[4 * j] = v.x; g[4 * j + 1] = v.y; g[4 * j + 2] = v.z; g[4 * j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int t = r * 32 + lane;
      g[r] = t < k ? __ldg(grow + t) : -1;
    }
  }
  const int64_t nout = n - ct;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (int b = 0; b < B; b++) {
    const double* cp = copy_in + (int64_t)b * n;
    uint64_t key[R];
    int cnt = 0;
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      double val = -1.0;
      if (g[r] >= 0) val = __ldg(cp + g[r]);
      const bool keep = val >= 0.0;  // NaN and negatives (masked bins) dropped
      key[r] = keep ? ((uint64_t)__double_as_longlong(val) | 0x8000000000000000ull) : ~0ull;
      cnt += keep ? 1 : 0;
      sum += keep ? val : 0.0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    double mean = nan, sd = nan, med = nan;
    if (cnt > 0) {
      mean = sum / (double)cnt;
      double ss = 0.0;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const double t = __longlong_as_double((long long)(key[r] & 0x7fffffffffffffffull)) - mean;
        ss += key[r] != ~0ull ? t * t : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sd = sqrt(ss / (double)cnt);
      med = warp_median<R>(key, cnt);
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

// ---- np.nanmedian over rows (optionally of log2): multi-CTA radix select -----------------------------------
// The two medians of normalize_repeat (predict_tools.py:106-107: nanmedian(log2(r)), nanmedian(z)) are order
// statistics of ~1.9e5 doubles per sample.  (The first version ran one CTA per sample with 64 bisection passes and
// log2 recomputed in each: 14.7 ms of the 18.3 ms of a batch-1 normalize, on 1 of 148 SMs.)  Here: the values become
// order-preserving 64-bit keys ONCE (log2 applied once, NaN counted out), then six passes of an 11-bit radix select
// (digits at bit 53, 42, 31, 20, 9, 0) over all SMs.  Pass p histograms digit p of the keys that match the prefix
// found so far; the next launch starts by scanning that histogram (every CTA does the same 2048-bin scan, CTA 0 of the
// row records the state).  Both middle order statistics ((cnt - 1) / 2 and cnt / 2) are tracked, so an even count
// needs no extra pass.  blockIdx.z selects the array (0: log2(r), 1: z): seven launches do both medians of a batch.
constexpr int RS_BINS = 2048;
constexpr int RS_PASSES = 6;
__constant__ int RS_SHIFT[RS_PASSES] = {53, 42, 31, 20, 9, 0};

struct RadixState {           // per (array, row, pass): state BEFORE pass p
  unsigned long long prefix[2];  // high bits fixed so far for the lower / upper middle rank
  long long rank[2];             // rank of the wanted key among the keys that match the prefix
};

__device__ __forceinline__ int rs_digit(unsigned long long key, int pass) {
  return (int)((key >> RS_SHIFT[pass]) & (pass == RS_PASSES - 1 ? 511ull : 2047ull));
}
// mask of the bits above digit `pass`
__device__ __forceinline__ unsigned long long rs_himask(int pass) {
  return pass == 0 ? 0ull : ~((1ull << (RS_SHIFT[pass - 1])) - 1ull);
}

// keys[a][row][i], cnt[a][row] = number of non-NaN values, hist of pass 0
__global__ void __launch_bounds__(256)
radix_keys_kernel(const double* __restrict__ r, const double* __restrict__ z, int64_t len, unsigned long long* __restrict__ keys,
                  unsigned long long* __restrict__ cnt, unsigned int* __restrict__ hist, int B) {
  __shared__ unsigned int sh[RS_BINS];
  __shared__ unsigned long long s_cnt;
  const int a = blockIdx.z, row = blockIdx.y;
  const double* src = (a == 0 ? r : z) + (int64_t)row * len;
  unsigned long long* kd = keys + ((int64_t)a * B + row) * len;
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned long long loc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    double v = src[i];
    if (a == 0) v = log2(v);
    unsigned long long k = ~0ull;  // NaN: above every number, never counted
    if (v == v) { k = dkey(v); loc++; atomicAdd(&sh[(int)(k >> 53)], 1u); }
    kd[i] = k;
  }
  loc = __reduce_add_sync(0xffffffffu, (unsigned)loc);
  if ((threadIdx.x & 31) == 0 && loc) atomicAdd(&s_cnt, loc);
  __syncthreads();
  unsigned int* h = hist + (((int64_t)a * B + row) * RS_PASSES + 0) * 2 * RS_BINS;
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
  if (threadIdx.x == 0 && s_cnt) atomicAdd(&cnt[(int64_t)a * B + row], s_cnt);
}

// scan of the histogram(s) of pass `pass - 1` -> state before `pass`; valid in every thread after the call
__device__ __forceinline__ RadixState rs_advance(const RadixState& prev, const unsigned int* __restrict__ h, int pass_done,
                                                 unsigned int* sh_scan /*[RS_BINS]*/) {
  RadixState out = prev;
  const bool same = prev.prefix[0] == prev.prefix[1] && prev.rank[0] == prev.rank[1];
  for (int sidx = 0; sidx < 2; sidx++) {
    if (sidx == 1 && same) { out.prefix[1] = out.prefix[0]; out.rank[1] = out.rank[0]; break; }
    // the two ranks share one histogram while their prefixes agree
    const unsigned int* hh = h + ((sidx == 1 && prev.prefix[0] != prev.prefix[1]) ? RS_BINS : 0);
    // block-wide inclusive scan of 2048 bins: 8 bins per thread (256 threads)
    unsigned int v[8], run = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = hh[threadIdx.x * 8 + j]; run += v[j]; }
    __syncthreads();
    sh_scan[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const unsigned int t = threadIdx.x >= o ? sh_scan[threadIdx.x - o] : 0u;
      __syncthreads();
      sh_scan[threadIdx.x] += t;
      __syncthreads();
    }
    const long long want = prev.rank[sidx];
    long long before = threadIdx.x ? (long long)sh_scan[threadIdx.x - 1] : 0ll;
    __syncthreads();
    // the owning thread publishes (digit, rank inside the bucket)
    if (want >= before && want < before + (long long)run) {
      int j = 0;
      while (want >= before + (long long)v[j]) { before += v[j]; j++; }
      sh_scan[256] = (unsigned)(threadIdx.x * 8 + j);
      sh_scan[257] = (unsigned)(want - before);
    }
    __syncthreads();
    out.prefix[sidx] = prev.prefix[sidx] | ((unsigned long long)sh_scan[256] << RS_SHIFT[pass_done]);
    out.rank[sidx] = (long long)sh_scan[257];
    __syncthreads();
  }
  return out;
}

// pass p in 1..RS_PASSES: advance the state from the histogram of pass p - 1, then histogram digit p of the matching
// keys.  p == RS_PASSES (launched with one CTA per row) only advances -- the prefixes are then the two keys -- and
// writes the median.  states: [2][B][RS_PASSES + 1].
__global__ void __launch_bounds__(256)
radix_pass_kernel(const unsigned long long* __restrict__ keys, int64_t len, const unsigned long long* __restrict__ cnt,
                  unsigned int* __restrict__ hist, RadixState* __restrict__ states, int pass, int B,
                  double* __restrict__ m_lr, double* __restrict__ m_z) {
  __shared__ unsigned int sh[2 * RS_BINS];
  __shared__ unsigned int sh_scan[258];
  const int a = blockIdx.z, row = blockIdx.y;
  const int64_t ar = (int64_t)a * B + row;
  const long long n = (long long)cnt[ar];
  if (n == 0) {
    if (pass == RS_PASSES && threadIdx.x == 0) (a == 0 ? m_lr : m_z)[row] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  RadixState prev;
  if (pass == 1) {
    prev.prefix[0] = prev.prefix[1] = 0ull;
    prev.rank[0] = (n - 1) >> 1;
    prev.rank[1] = n >> 1;
  } else {
    prev = states[ar * (RS_PASSES + 1) + (pass - 1)];
  }
  const RadixState st = rs_advance(prev, hist + (ar * RS_PASSES + (pass - 1)) * 2 * RS_BINS, pass - 1, sh_scan);
  if (pass == RS_PASSES) {
    // np.median / np.nanmedian: mean of the two middle order statistics (the same key twice for an odd count)
    if (threadIdx.x == 0) (a == 0 ? m_lr : m_z)[row] = (key_d(st.prefix[0]) + key_d(st.prefix[1])) / 2.0;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) states[ar * (RS_PASSES + 1) + pass] = st;
  for (int i = threadIdx.x; i < 2 * RS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned long long hm = rs_himask(pass);
  const bool split = st.prefix[0] != st.prefix[1];
  const unsigned long long* kd = keys + ar * len;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = kd[i];
    if (k == ~0ull) continue;
    const unsigned long long kh = k & hm;
    if (kh == st.prefix[0]) atomicAdd(&sh[rs_digit(k, pass)], 1u);
    else if (split && kh == st.prefix[1]) atomicAdd(&sh[RS_BINS + rs_digit(k, pass)], 1u);
  }
  __syncthreads();
  unsigned int* h = hist + (ar * RS_PASSES + pass) * 2 * RS_BINS;
  for (int i = threadIdx.x; i < 2 * RS_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// ---- get_z_score ---------------------------------------------------------------------------------------
// Per segment and null column: weighted mean of the column over the segment's bins with data, then
// z = (segment ratio - mean over columns) / population sd over columns (overall_tools.py:88-119).
// Two stages so that a whole-chromosome segment (16 598 bins at 15 kb) is not one CTA's serial loop: stage 1 cuts
// every segment into SZ_CHUNKS bin ranges (grid = chunks x segments), 128 null columns x 2 bin lanes per CTA, and
// writes per-chunk partial sums; stage 2 adds the chunks in a fixed order (deterministic) and finishes the segment.
// M <= 128 (the reference uses min(S, 100) null samples, newref_tools.py:211).
constexpr int SZ_CHUNKS = 32;

__global__ void __launch_bounds__(256)
segment_z_partial_kernel(const double* __restrict__ nr, int m, const int32_t* __restrict__ inflate_pos, const double* __restrict__ r,
                         const double* __restrict__ w, const int64_t* __restrict__ seg_se, double* __restrict__ partial) {
  __shared__ double s_num[128], s_den[128], s_any[128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int seg = blockIdx.y, ch = blockIdx.x;
  const int64_t s = seg_se[2 * seg], e = seg_se[2 * seg + 1], len = e - s;
  const int64_t cs = s + len * ch / SZ_CHUNKS, ce = s + len * (ch + 1) / SZ_CHUNKS;
  double num = 0.0, den = 0.0, any = 0.0;
  if (tx < m) {
    for (int64_t b = cs + ty; b < ce; b += 2) {
      if (r[b] == 0.0) continue;  // bins without data are dropped (overall_tools.py:98-100,106)
      const int32_t p = inflate_pos[b];
      if (p < 0) continue;
      const double v = nr[(int64_t)p * m + tx];
      if (isfinite(v)) { num += w[b] * v; den += w[b]; any = 1.0; }  // non-finite entries are masked (:101-108)
    }
  }
  if (ty == 1) { s_num[tx] = num; s_den[tx] = den; s_any[tx] = any; }
  __syncthreads();
  if (ty == 0) {
    double* o = partial + ((int64_t)seg * SZ_CHUNKS + ch) * 3 * 128;
    o[tx] = num + s_num[tx];
    o[128 + tx] = den + s_den[tx];
    o[256 + tx] = any + s_any[tx];
  }
}

__global__ void __launch_bounds__(128)
segment_z_final_kernel(const double* __restrict__ partial, int m, const double* __restrict__ seg_r, double* __restrict__ z_out) {
  __shared__ double sh[32];
  __shared__ double s_mean, s_count;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const int tx = threadIdx.x, seg = blockIdx.x;
  double num = 0.0, den = 0.0, any = 0.0;
  for (int ch = 0; ch < SZ_CHUNKS; ch++) {
    const double* o = partial + ((int64_t)seg * SZ_CHUNKS + ch) * 3 * 128;
    num += o[tx]; den += o[128 + tx]; any += o[256 + tx];
  }
  double val = nan;
  if (tx < m && any > 0.0) val = num / den;
  const bool ok = tx < m && isfinite(val);
  const double sum_v = block_sum(ok ? val : 0.0, sh);
  const double sum_c = block_sum(ok ? 1.0 : 0.0, sh);
  if (threadIdx.x == 0) { s_mean = sum_v / sum_c; s_count = sum_c; }
  __syncthreads();
  const double mean = s_mean, cnt = s_count;
  const double dv = ok ? (val - mean) : 0.0;
  const double ss = block_sum(dv * dv, sh);
  if (threadIdx.x == 0) {
    double z = nan;
    if (cnt > 0.0) {
      const double sd = sqrt(ss / cnt);
      z = (seg_r[seg] - mean) / sd;
      if (z == z) { z = z > 1000.0 ? 1000.0 : z; z = z < -1000.0 ? -1000.0 : z; }  // min/max clip (:114-115)
    }
    z_out[seg] = z;
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

// gl [n - ct, k] for one (reference set, cutoff): see gather_list_kernel
int launch_gather_list(const int32_t* idx, const double* dist, int64_t n, int32_t k, const double* cutoff_dev,
                       const int64_t* cum_dev, int32_t nchr, int64_t ct, int32_t* gl, cudaStream_t st) {
  const int64_t nout = n - ct;
  if (nout <= 0) return 0;
  gather_list_kernel<<<(unsigned)((nout + 7) / 8), 256, 0, st>>>(idx, dist, n, k, cutoff_dev, cum_dev, nchr, ct, gl);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t radix_scratch_bytes(int32_t B, int64_t len) {
  return sizeof(unsigned long long) * 2 * (size_t)B * (size_t)len       // keys
         + sizeof(unsigned long long) * 2 * (size_t)B                   // valid counts
         + sizeof(unsigned int) * 2 * (size_t)B * RS_PASSES * 2 * RS_BINS  // histograms
         + sizeof(RadixState) * 2 * (size_t)B * (RS_PASSES + 1);
}

// m_lr[b] = nanmedian(log2(r[b, :])), m_z[b] = nanmedian(z[b, :]) (predict_tools.py:106-107)
int launch_nanmedians(const double* r, const double* z, int32_t B, int64_t len, void* scratch, double* m_lr, double* m_z,
                      cudaStream_t st) {
  if (B <= 0) return 0;
  unsigned char* p = reinterpret_cast<unsigned char*>(scratch);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(p);
  p += sizeof(unsigned long long) * 2 * (size_t)B * (size_t)len;
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(p);
  p += sizeof(unsigned long long) * 2 * (size_t)B;
  unsigned int* hist = reinterpret_cast<unsigned int*>(p);
  const size_t hist_bytes = sizeof(unsigned int) * 2 * (size_t)B * RS_PASSES * 2 * RS_BINS;
  p += hist_bytes;
  RadixState* states = reinterpret_cast<RadixState*>(p);
  WCX_CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 2 * (size_t)B + hist_bytes, st));
  // about four CTAs per SM over the whole batch, at least 1024 keys per CTA
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((len + 1023) / 1024, (592 + 2 * B - 1) / (2 * B)));
  const dim3 grid(chunks, B, 2);
  radix_keys_kernel<<<grid, 256, 0, st>>>(r, z, len, keys, cnt, hist, B);
  for (int pass = 1; pass < RS_PASSES; pass++)
    radix_pass_kernel<<<grid, 256, 0, st>>>(keys, len, cnt, hist, states, pass, B, m_lr, m_z);
  radix_pass_kernel<<<dim3(1, B, 2), 256, 0, st>>>(keys, len, cnt, hist, states, RS_PASSES, B, m_lr, m_z);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// three passes of _normalize_once with the -1 masking in between (normalize_repeat; the two nanmedians that follow are
// launch_nanmedians).  copy_a / copy_b: [B, n] ping-pong buffers; gl: launch_gather_list.
int launch_normalize_repeat(const double* x, double* copy_a, double* copy_b, int32_t B, int64_t n, const int32_t* gl,
                            int32_t k, int64_t ct, double* z, double* r, double* nref, cudaStream_t st) {
  if (k > PR_MAXK) { set_error("normalize: ref_size > 512 unsupported"); return 1; }
  const int64_t nout = n - ct;
  if (nout <= 0 || B <= 0) return 0;
  WCX_CUDA_OK(cudaMemcpyAsync(copy_a, x, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToDevice, st));
  // the passes rewrite the target bins [ct, n) only: the second buffer needs the prefix [0, ct) of every sample once
  if (ct > 0)
    WCX_CUDA_OK(cudaMemcpy2DAsync(copy_b, sizeof(double) * (size_t)n, x, sizeof(double) * (size_t)n, sizeof(double) * (size_t)ct,
                                  (size_t)B, cudaMemcpyDeviceToDevice, st));
  const unsigned grid = (unsigned)((nout + 7) / 8);
  double* in = copy_a;
  double* out = copy_b;
  for (int pass = 0; pass < 3; pass++) {
    if (k <= 384)
      normalize_pass_kernel<12><<<grid, 256, 0, st>>>(x, in, pass < 2 ? out : nullptr, B, n, gl, k, ct, z, r, nref);
    else
      normalize_pass_kernel<16><<<grid, 256, 0, st>>>(x, in, pass < 2 ? out : nullptr, B, n, gl, k, ct, z, r, nref);
    double* t = in; in = out; out = t;
  }
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t segment_z_scratch_bytes(int32_t nseg) { return sizeof(double) * (size_t)(nseg > 0 ? nseg : 1) * SZ_CHUNKS * 3 * 128; }

int launch_segment_z(const double* nr, int32_t m, const int32_t* inflate_pos, const double* r, const double* w,
                     const int64_t* seg_se, const double* seg_r, int32_t nseg, double* z_out, double* partial, cudaStream_t st) {
  if (nseg <= 0) return 0;
  if (m > 128) { set_error("get_z_score: more than 128 null samples unsupported"); return 1; }
  if (nseg > 65535) { set_error("get_z_score: more than 65535 segments in one call unsupported"); return 1; }
  segment_z_partial_kernel<<<dim3(SZ_CHUNKS, nseg), 256, 0, st>>>(nr, m, inflate_pos, r, w, seg_se, partial);
  segment_z_final_kernel<<<nseg, 128, 0, st>>>(partial, m, seg_r, z_out);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
