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
#include <algorithm>
#include <cstdlib>

#include <cuda_fp16.h>

#include "null_ratios.cuh"
#include "packed_select.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {
struct NqCol;
}
// layout of the null-ratio staging buffer (`xt`): XM keys | codes of the fast path | column maps | histograms
struct NullStaging {
  uint16_t* xq;
  NqCol* cols;
  unsigned int* hist;
  size_t bytes;
};
NullStaging null_staging(double* xt, int64_t n, int32_t m);

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

// =================================================================================================================
// Fast path: one THREAD per (target bin, sample column).
//
// The warp-per-bin kernel above spends ~650 warp instructions per median of 300 (a 64-bit key bisection whose fixed
// cost per step -- warp reduction, bracket bookkeeping -- is paid per column): 18 ms for the 1.9e7 medians of config 3,
// issue bound at 0.5 % of its HBM traffic bound.  Here every sample value gets a 15-bit order-preserving CODE once per
// call (nq_* kernels: an equal-frequency map through a 65 536-bin histogram of the column, so the ~N values of a column
// share a code with ~N / 30 000 others), stored as the bit pattern of a positive normal fp16.  A thread keeps the codes
// of its bin's k reference bins packed two per register (150 registers for k = 300) and selects the upper-median code by
// a 15-step bisection in which one HSET2 + one HADD2 handle two keys: ~2 k instructions per step and thread, no
// shuffles, no shared memory, no divergence.  The two middle VALUES are then fetched in float64 through the positions of
// the selected codes.  A selected code that is not unique among the k keys (two reference bins inside one code cell,
// ~1 % of the medians; columns with NaN / inf / no spread; placeholder rows) is resolved by the exact warp-cooperative
// selection of select.cuh, one flagged median at a time -- results are identical to the kernel above in every case.
// =================================================================================================================
constexpr int NQ_STRIDE = 128;         // code row stride in columns (uint16)
constexpr int NQ_HBINS = 65536;        // histogram cells per column
constexpr uint32_t NQ_LO = PS_CODE_LO, NQ_HI = PS_CODE_HI, NQ_PAD = PS_CODE_PAD;

struct NqCol {
  double lo, scale;   // histogram cell of v: (v - lo) * scale
  double sum, sumsq;  // accumulators of nq_stats_kernel
  unsigned long long cnt, bad;
  int32_t exact;      // 1: every median of this column takes the exact path (NaN / inf present or zero spread)
  int32_t pad_;
};

__device__ __forceinline__ double nq_value(const uint64_t* __restrict__ xm, int64_t n, int64_t bin, int col) {
  return key_d(__ldg(xm + ((int64_t)(col >> 3) * n + bin) * NR_CHUNK + (col & 7)));
}
__device__ __forceinline__ uint64_t nq_key(const uint64_t* __restrict__ xm, int64_t n, int64_t bin, int col) {
  return __ldg(xm + ((int64_t)(col >> 3) * n + bin) * NR_CHUNK + (col & 7));
}

__global__ void __launch_bounds__(256)
nq_stats_kernel(const uint64_t* __restrict__ xm, int64_t n, int m, NqCol* __restrict__ cols) {
  __shared__ double sh[3][8];
  const int col = blockIdx.y;
  double s1 = 0.0, s2 = 0.0, bad = 0.0;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const double v = nq_value(xm, n, b, col);
    if (isfinite(v)) { s1 += v; s2 += v * v; } else bad += 1.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; sh[2][threadIdx.x >> 5] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0, c = 0.0;
    for (int w = 0; w < 8; w++) { a += sh[0][w]; b2 += sh[1][w]; c += sh[2][w]; }
    atomicAdd(&cols[col].sum, a);
    atomicAdd(&cols[col].sumsq, b2);
    if (c > 0.0) atomicAdd(&cols[col].bad, (unsigned long long)c);
  }
}

__global__ void nq_map_kernel(int64_t n, int m, NqCol* __restrict__ cols) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= m) return;
  NqCol c = cols[col];
  const double good = (double)n - (double)c.bad;
  const double mean = good > 0.0 ? c.sum / good : 0.0;
  double var = good > 0.0 ? c.sumsq / good - mean * mean : 0.0;
  if (!(var > 0.0)) var = 0.0;
  const double sd = sqrt(var);
  c.lo = mean - 8.0 * sd;
  c.scale = (double)NQ_HBINS / (16.0 * sd);
  c.exact = (c.bad != 0 || !(sd > 0.0) || !isfinite(c.scale) || !isfinite(c.lo)) ? 1 : 0;
  cols[col] = c;
}

__device__ __forceinline__ double nq_cell(double v, const NqCol& c) {
  double t = (v - c.lo) * c.scale;
  t = t < 0.0 ? 0.0 : t;
  const double top = (double)NQ_HBINS - 1.0 / 1024.0;
  return t > top ? top : t;  // NaN (exact columns only) falls through the comparisons: harmless
}

// hist[col][cell] += 1
__global__ void __launch_bounds__(256)
nq_hist_kernel(const uint64_t* __restrict__ xm, int64_t n, int m, const NqCol* __restrict__ cols, unsigned int* __restrict__ hist) {
  const int col = blockIdx.y;
  const NqCol c = cols[col];
  if (c.exact) return;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const double t = nq_cell(nq_value(xm, n, b, col), c);
    atomicAdd(&hist[(int64_t)col * (NQ_HBINS + 1) + (int)t], 1u);
  }
}

// in place: hist[col][0 .. HBINS] becomes the exclusive prefix sum (entry HBINS = total); one CTA per column
__global__ void __launch_bounds__(1024)
nq_prefix_kernel(unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[1024];
  unsigned int* h = hist + (int64_t)blockIdx.x * (NQ_HBINS + 1);
  constexpr int PER = NQ_HBINS / 1024;
  unsigned int v[PER], run = 0;
#pragma unroll
  for (int j = 0; j < PER; j++) { v[j] = h[threadIdx.x * PER + j]; run += v[j]; }
  sh[threadIdx.x] = run;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned int before = threadIdx.x ? sh[threadIdx.x - 1] : 0u;
#pragma unroll
  for (int j = 0; j < PER; j++) { h[threadIdx.x * PER + j] = before; before += v[j]; }
  if (threadIdx.x == 1023) h[NQ_HBINS] = before;
}

// xq[bin][col] = code: LO + floor((prefix[cell] + frac * count[cell]) * (HI - LO + 1) / total), monotone in the value
__global__ void __launch_bounds__(256)
nq_code_kernel(const uint64_t* __restrict__ xm, int64_t n, int m, const NqCol* __restrict__ cols, const unsigned int* __restrict__ hist,
               uint16_t* __restrict__ xq) {
  const int col = threadIdx.x & (NQ_STRIDE - 1);
  const int64_t b = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7);
  if (b >= n || col >= m) return;
  const NqCol c = cols[col];
  uint32_t code = NQ_LO;
  if (!c.exact) {
    const double t = nq_cell(nq_value(xm, n, b, col), c);
    const int cell = (int)t;
    const unsigned int* h = hist + (int64_t)col * (NQ_HBINS + 1);
    const double c0 = (double)h[cell], c1 = (double)h[cell + 1], tot = (double)h[NQ_HBINS];
    const double pos = c0 + (t - (double)cell) * (c1 - c0);
    const double q = pos * ((double)(NQ_HI - NQ_LO + 1) / tot);
    code = NQ_LO + (uint32_t)q;
    code = code > NQ_HI ? NQ_HI : code;
  }
  xq[b * NQ_STRIDE + col] = (uint16_t)code;
}

// reference position -> bin of the full column: -1 (missing entries) wraps to the last bin like a Python index.  The
// positions come from this library's own re-rank or were validated by the host (wcx_newref_null_ratios), so they lie
// in [-n, n): no clamp on the hot path.
__device__ __forceinline__ int32_t nq_wrap(int32_t v, int32_t n) { return v + ((v >> 31) & n); }

// NP packed registers (2 NP >= k keys; FULL: k == 2 NP exactly, no bounds checks in the gather), R = slots per lane of
// the exact warp-cooperative path (32 R >= k).
// Measured variants of this kernel at config 3 (stand-alone call incl. 0.6 ms of code building): 255 registers / 8 warps
// per SM 11.6 ms (this one); capped at 168 registers / 12 warps per SM 12.4 ms (the compiler then serialises the count
// into one accumulator chain); two threads per median with half the keys each (126 registers, four chains, 16 warps
// per SM, one shuffle per step) 14.0 ms; persistent blocks with the next chunk's positions prefetched by cp.async 14.0 ms.
// Neither more warps nor prefetching the positions bought throughput here.
template <int NP, int R, bool FULL>
__global__ void __launch_bounds__(128, NP > 128 ? 2 : (NP > 64 ? 3 : 4))
null_fast_kernel(const uint16_t* __restrict__ xq, const uint64_t* __restrict__ xm, const NqCol* __restrict__ cols, int64_t n,
                 const int32_t* __restrict__ idx, int64_t row_begin, int64_t rows, int k, int m, double* __restrict__ out) {
  // A block covers 128 consecutive (bin, column) pairs = at most 128 / m + 2 bins.  Their reference positions are turned
  // into element offsets of the code table once per block (wrap + multiply), shared by all columns of the bin.
  extern __shared__ int32_t s_off[];  // [bins of the block][k]
  const int lane = threadIdx.x & 31;
  const int64_t total = rows * m;
  const int64_t blk0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t row_first = blk0 / m;
  const int64_t blk_last = blk0 + blockDim.x - 1 < total - 1 ? blk0 + blockDim.x - 1 : total - 1;
  const int nrows_blk = (int)(blk_last / m - row_first) + 1;
  const int32_t n32 = (int32_t)n;
  for (int e = threadIdx.x; e < nrows_blk * k; e += blockDim.x) s_off[e] = nq_wrap(idx[row_first * k + e], n32) * NQ_STRIDE;
  __syncthreads();
  const int64_t p0 = blk0 + threadIdx.x;
  const bool active = p0 < total;
  const int64_t p = active ? p0 : total - 1;  // whole warps stay alive for the cooperative exact path
  const int64_t lrow = p / m;
  const int col = (int)(p - lrow * m);
  const int32_t* __restrict__ orow = s_off + (lrow - row_first) * k;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double med = nan;
  bool need_exact = cols[col].exact != 0;
  bool done = false;
  // rows whose k positions are all the same bin (placeholder rows of gonosomal references, newref_tools.py:186-191):
  // the median of k copies of a value is the value ((v + v) / 2 is exact)
  if (orow[0] == orow[k - 1] && orow[0] == orow[k >> 1]) {
    bool same = true;
    for (int t = 1; t < k; t++) same &= orow[t] == orow[0];
    if (same) {
      med = nq_value(xm, n, orow[0] / NQ_STRIDE, col);
      done = true;
    }
  }
  if (!done && !need_exact) {
    uint32_t k2[NP];
    const uint16_t* __restrict__ xqc = xq + col;
    if (FULL) {
      // k == 2 NP: four offsets per 128-bit shared-memory load (a broadcast: the threads of a bin read the same words)
      const int4* __restrict__ ov = reinterpret_cast<const int4*>(orow);
#pragma unroll
      for (int j = 0; j < NP / 2; j++) {
        const int4 q = ov[j];
        const uint32_t c0 = xqc[q.x], c1 = xqc[q.y], c2 = xqc[q.z], c3 = xqc[q.w];
        k2[2 * j] = c0 | (c1 << 16);
        k2[2 * j + 1] = c2 | (c3 << 16);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NP; j++) {
        uint32_t c0 = NQ_PAD, c1 = NQ_PAD;
        if (2 * j < k) c0 = xqc[orow[2 * j]];
        if (2 * j + 1 < k) c1 = xqc[orow[2 * j + 1]];
        k2[j] = c0 | (c1 << 16);
      }
    }
    const int t = k >> 1;  // rank of the upper middle key (the median itself for odd k)
    int below = 0;
    const uint32_t T = ps_select<NP>(k2, t, below);
    // T is the code of the rank-t key: count(< T) <= t < count(< T + 1).  It can stand for its VALUE only if no
    // other key shares the code; an even k also needs the rank t - 1 key: the largest key below T, provided rank t is
    // the first key with code T.
    int eq_hi, j_hi, eq_lo = 1, j_lo = 0;
    ps_find<NP>(k2, T, eq_hi, j_hi);
    bool ok = eq_hi == 1;
    if (ok && !(k & 1)) {
      ok = below == t;
      if (ok) {
        ps_find<NP>(k2, ps_max_below<NP>(k2, T), eq_lo, j_lo);
        ok = eq_lo == 1;
      }
    }
    if (ok) {
      const double v_hi = nq_value(xm, n, orow[j_hi] / NQ_STRIDE, col);
      if (k & 1) {
        med = v_hi;
      } else {
        const double v_lo = nq_value(xm, n, orow[j_lo] / NQ_STRIDE, col);
        med = (v_lo + v_hi) / 2.0;  // np.median: mean of the two middle values
      }
      done = true;
    } else {
      need_exact = true;
    }
  }
  // exact path, one flagged (bin, column) at a time by the whole warp
  uint32_t pending = __ballot_sync(0xffffffffu, need_exact && !done);
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    const int64_t srow = __shfl_sync(0xffffffffu, lrow, src);
    const int scol = __shfl_sync(0xffffffffu, col, src);
    const int32_t* __restrict__ ir = idx + srow * k;
    uint64_t key[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int t = r * 32 + lane;
      key[r] = t < k ? nq_key(xm, n, nq_wrap(ir[t], n32), scol) : ~0ull;
    }
    uint32_t hmax = 0;
    const double mm = warp_median<R>(key, k, &hmax);
    if (lane == src) med = hmax == 0xffffffffu ? nan : mm;  // np.median is NaN when any value is NaN
  }
  if (active) out[lrow * m + col] = log2(nq_value(xm, n, row_begin + lrow, col) / med);
}
}  // namespace

namespace {
__global__ void validate_positions_kernel(const int32_t* __restrict__ idx, int64_t count, int32_t n, int32_t* __restrict__ bad) {
  int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = idx[i];
    local |= (v < -n || v >= n) ? 1 : 0;
  }
  if (__any_sync(0xffffffffu, local) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}
}  // namespace

// bad[0] |= 1 if any position of a caller-supplied DEVICE index array lies outside [-n, n) (Python index semantics)
int launch_validate_positions(const int32_t* idx, int64_t count, int64_t n, int32_t* bad, cudaStream_t st) {
  if (count <= 0) return 0;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, 1184));
  validate_positions_kernel<<<grid, 256, 0, st>>>(idx, count, (int32_t)n, bad);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_transpose_cols(const double* x, int64_t n, int32_t s, const int32_t* ids, int32_t m, double* xt,
                          cudaStream_t st) {
  if (n == 0 || m == 0) return 0;
  dim3 grid((unsigned)((n + 31) / 32), (m + NR_CHUNK - 1) / NR_CHUNK);
  gather_cols_kernel<<<grid, 256, 0, st>>>(x, n, s, ids, m, xt);
  WCX_CUDA_OK(cudaGetLastError());
  if (m <= NQ_STRIDE) {
    // codes of the fast path (see above): column statistics, histogram, prefix, codes
    NullStaging ns = null_staging(xt, n, m);
    WCX_CUDA_OK(cudaMemsetAsync(ns.cols, 0, sizeof(NqCol) * NQ_STRIDE + sizeof(unsigned int) * (size_t)m * (NQ_HBINS + 1), st));
    const uint64_t* xm = reinterpret_cast<const uint64_t*>(xt);
    const unsigned gb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 592 / m + 1));
    nq_stats_kernel<<<dim3(gb, m), 256, 0, st>>>(xm, n, m, ns.cols);
    nq_map_kernel<<<1, NQ_STRIDE, 0, st>>>(n, m, ns.cols);
    nq_hist_kernel<<<dim3(gb, m), 256, 0, st>>>(xm, n, m, ns.cols, ns.hist);
    nq_prefix_kernel<<<m, 1024, 0, st>>>(ns.hist);
    nq_code_kernel<<<(unsigned)((n + 1) / 2), 256, 0, st>>>(xm, n, m, ns.cols, ns.hist, ns.xq);
    WCX_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

NullStaging null_staging(double* xt, int64_t n, int32_t m) {
  NullStaging ns;
  unsigned char* p = reinterpret_cast<unsigned char*>(xt);
  p += sizeof(double) * (size_t)((m + NR_CHUNK - 1) / NR_CHUNK) * n * NR_CHUNK;  // XM keys
  ns.xq = reinterpret_cast<uint16_t*>(p);
  p += ((sizeof(uint16_t) * (size_t)n * NQ_STRIDE + 255) / 256) * 256;
  ns.cols = reinterpret_cast<NqCol*>(p);
  p += sizeof(NqCol) * NQ_STRIDE;
  ns.hist = reinterpret_cast<unsigned int*>(p);
  p += sizeof(unsigned int) * (size_t)NQ_STRIDE * (NQ_HBINS + 1);
  ns.bytes = (size_t)(p - reinterpret_cast<unsigned char*>(xt));
  return ns;
}

int64_t null_ratio_staging_doubles(int64_t n, int32_t m) { return (int64_t)((null_staging(nullptr, n, m).bytes + 7) / 8); }

int launch_null_ratios(const double* xt, int64_t n, const int32_t* idx, int64_t row_begin, int64_t row_end,
                       int32_t k, int32_t m, double* out, cudaStream_t st) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || m <= 0) return 0;
  if (k > 512) { set_error("null_ratios: ref_size > 512 unsupported"); return 1; }
  const bool legacy = std::getenv("WCX_NULL_WARP") != nullptr;  // cross-check (tests): the warp-per-bin kernel
  if (!legacy && m <= NQ_STRIDE && k <= 400 && k >= 2 && sizeof(int32_t) * (size_t)(128 / m + 2) * k <= 40 * 1024) {
    NullStaging ns = null_staging(const_cast<double*>(xt), n, m);
    const uint64_t* xm = reinterpret_cast<const uint64_t*>(xt);
    const int64_t total = rows * m;
    const unsigned grid = (unsigned)((total + 127) / 128);
    // shared memory: reference offsets of the bins a block covers (128 / m + 2 of them, k each; rows 16-byte aligned)
    const size_t smem = sizeof(int32_t) * (size_t)(128 / m + 2) * k;
    const bool aligned = (k & 3) == 0;
#define WCX_NQ_LAUNCH(NP, R, FULL) null_fast_kernel<NP, R, FULL><<<grid, 128, smem, st>>>(ns.xq, xm, ns.cols, n, idx, row_begin, rows, k, m, out)
    if (k == 300 && aligned) WCX_NQ_LAUNCH(150, 10, true);
    else if (k <= 64) WCX_NQ_LAUNCH(32, 2, false);
    else if (k <= 128) WCX_NQ_LAUNCH(64, 4, false);
    else if (k <= 200) WCX_NQ_LAUNCH(100, 7, false);
    else if (k <= 300) WCX_NQ_LAUNCH(150, 10, false);
    else WCX_NQ_LAUNCH(200, 13, false);
#undef WCX_NQ_LAUNCH
    WCX_CUDA_OK(cudaGetLastError());
    return 0;
  }
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
