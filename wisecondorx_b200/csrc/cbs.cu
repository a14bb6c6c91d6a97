// CUDA circular binary segmentation -- replaces the R / DNAcopy call of the reference:
//   exec_cbs (predict_tools.py:242-263) -> exec_R (overall_tools.py:65-80) -> include/CBS.R:70-73
//   DNAcopy::segment(CNA(...), alpha = alpha, weights = w)      [DNAcopy 1.76, conda.yml:14]
//
// DNAcopy is not part of the reference tree; the algorithm restated here is the published one
// (Olshen et al. 2004; Venkatraman & Olshen 2007) with DNAcopy's documented defaults and decision
// flow (weighted `changepoints` / `wfindcpt`): maximal weighted t-statistic over all arcs, the
// |t| <= 0.1 / |t| >= 7 shortcuts, hybrid p-value (Siegmund tail approximation + permutation of the
// short-arc statistic) for segments longer than nmin, full permutation test otherwise, and the
// permutation t-tests that trim a two-change-point arc.  See oracle/cbs_oracle.py for the spec
// this file is compared with bit for bit (same Philox4x32-10 streams, same summation order).
//
// Mapping.  The recursion over segments is driven by the host in rounds; in every round ALL
// pending segments of ALL chromosome series (any number of samples x chromosomes) are processed
// together:
//   cbs_prepare_kernel   one thread per segment: weighted mean, centring, tss and the sequential
//                        prefix sums sx / cw (sequential on purpose: every rounding is specified)
//   cbs_maxarc_kernel    the O(n^2) all-arcs maximum, brute force, split into balanced chunks of
//                        start positions over the whole grid, register-tiled over end positions, division only
//                        for arcs that can still be the maximum; block reduce with the tie rule
//                        (largest bss, then smallest start, then smallest end)
//   cbs_tailp_kernel     Siegmund tail approximation (hybrid p-value, part 1) for every segment that needs it
//   cbs_perm_prep / _arcs / _count   permutation tests of ALL undecided segments of the round in one staged
//                        launch (256 permutations first, the remainder only for tests still undecided)
//   cbs_tprep / cbs_tperm  edge t-tests of two-change-point arcs (preparation batched per round)
// The scalar decisions (tail probability, thresholds) run on the host between the kernels.
// This file is compiled with -fmad=false: no FMA contraction, IEEE division.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "wcx_common.cuh"
#include "cbs.cuh"

namespace wcx {

namespace {

// ---------------- Philox4x32-10 ----------------
__host__ __device__ inline void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                            uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct PermStream {
  uint32_t k0, k1, b0, b1, b2, t;
  uint32_t buf[4];
  __device__ PermStream(uint32_t seed, uint32_t test, uint32_t lo, uint32_t hi, uint32_t perm)
      : k0(seed), k1(test), b0(perm), b1(lo), b2(hi), t(0) {}
  __device__ uint32_t below(uint32_t i) {
    if ((t & 3) == 0) philox4x32(t >> 2, b0, b1, b2, k0, k1, buf);
    const uint32_t v = buf[t & 3];
    t++;
    return (uint32_t)(((uint64_t)v * i) >> 32);
  }
};

// one segment of one series
struct Seg {
  int64_t lo, hi;   // absolute point range [lo, hi)
  int32_t series;   // chromosome series id
  int32_t pad;
};

struct SegPrep {  // written by cbs_prepare_kernel
  double tot_w, rtw, tss, tss_y;
  int32_t flat;   // 1: all values (numerically) equal -> no split
  int32_t pad;
};

struct ArcBest {
  double bss;
  int32_t i, j;
};

struct Chunk {
  int32_t seg;
  int32_t i0, i1;  // start positions [i0, i1)
  int32_t pad;
};

__device__ __forceinline__ bool arc_better(double b, int i, int j, double bb, int bi, int bj) {
  return (b > bb) || (b == bb && (i < bi || (i == bi && j < bj)));
}

// ------------------------------------------------------------------------------------------
// prepare: xc = x - weighted mean, sx[t] = sum_{u < t} w x (stored at t-1 for t = 1..n), cw likewise
// ------------------------------------------------------------------------------------------
// One CTA per segment.  Every sum keeps the left-to-right order the oracle specifies (np.cumsum / a sequential loop,
// as in DNAcopy's own Fortran loops): the element-wise terms of a tile of PREP_TILE points are computed by all threads
// from coalesced loads into shared memory, ONE thread adds them up in order (four independent chains), and all threads
// write the prefix arrays back.  (One thread per segment walking global memory took 5 ms for the 23 series of a
// sample -- 16 598 dependent round trips for chr1 at 15 kb.)
constexpr int PREP_TILE = 1024;
constexpr int PREP_THREADS = 128;

__global__ void __launch_bounds__(PREP_THREADS)
cbs_prepare_kernel(const double* __restrict__ y, const double* __restrict__ w, const Seg* __restrict__ segs,
                   int nseg, double* __restrict__ xc, double* __restrict__ sx, double* __restrict__ cw,
                   double* __restrict__ yy, SegPrep* __restrict__ prep) {
  __shared__ double s_a[PREP_TILE], s_b[PREP_TILE], s_c[PREP_TILE], s_d[PREP_TILE];
  __shared__ double s_mn[PREP_THREADS / 32], s_mx[PREP_THREADS / 32];
  __shared__ double s_avg, s_rtw;
  const int s = blockIdx.x;
  if (s >= nseg) return;
  const int tid = threadIdx.x;
  const int64_t lo = segs[s].lo, hi = segs[s].hi;
  double tw = 0.0, twx = 0.0, mn = y[lo], mx = mn;
  for (int64_t base = lo; base < hi; base += PREP_TILE) {
    const int nt = (int)(hi - base < PREP_TILE ? hi - base : PREP_TILE);
    for (int i = tid; i < nt; i += PREP_THREADS) {
      const double v = y[base + i], ww = w[base + i];
      s_a[i] = v * ww;
      s_b[i] = ww;
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < nt; i++) { twx += s_a[i]; tw += s_b[i]; }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  if ((tid & 31) == 0) { s_mn[tid >> 5] = mn; s_mx[tid >> 5] = mx; }
  if (tid == 0) { s_avg = twx / tw; s_rtw = sqrt(tw); }
  __syncthreads();
  const double avg = s_avg, rtw = s_rtw;
  double tss = 0.0, tssy = 0.0, asx = 0.0, acw = 0.0;
  for (int64_t base = lo; base < hi; base += PREP_TILE) {
    const int nt = (int)(hi - base < PREP_TILE ? hi - base : PREP_TILE);
    for (int i = tid; i < nt; i += PREP_THREADS) {
      const double ww = w[base + i];
      const double x = y[base + i] - avg;
      xc[base + i] = x;
      const double wx = ww * x;
      s_a[i] = wx;
      s_b[i] = ww;
      s_c[i] = wx * x;            // (ww * x) * x, as numpy evaluates ws * x * x
      const double yv = x * sqrt(ww);
      yy[base + i] = yv;
      s_d[i] = yv * yv;
    }
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < nt; i++) {
        tss += s_c[i];
        tssy += s_d[i];
        asx += s_a[i];
        acw += s_b[i];
        s_a[i] = asx;
        s_b[i] = acw;
      }
    __syncthreads();
    for (int i = tid; i < nt; i += PREP_THREADS) {
      sx[base + i] = s_a[i];
      cw[base + i] = s_b[i] / rtw;
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int q = 0; q < PREP_THREADS / 32; q++) { mn = s_mn[q] < mn ? s_mn[q] : mn; mx = s_mx[q] > mx ? s_mx[q] : mx; }
    SegPrep p;
    p.tot_w = tw; p.rtw = rtw; p.tss = tss; p.tss_y = tssy;
    p.flat = (fabs(mx - mn) < 1.5e-8) ? 1 : 0;
    p.pad = 0;
    prep[s] = p;
  }
}

// ------------------------------------------------------------------------------------------
// all-arcs maximum.  arcs 0 <= i < j <= n with al0 <= j - i <= n - al0.
// prefix arrays are stored shifted: P(t) = t == 0 ? 0 : arr[lo + t - 1]
// ------------------------------------------------------------------------------------------
// Register-tiled: a thread keeps MA_J end positions j (sx_j, cw_j) in registers and walks the chunk's start positions
// i staged in shared memory (one broadcast LDS.128 per i for MA_J arcs).  The division of the statistic is only
// executed for arcs that can still matter: with thr = best * (1 - 2^-50) an arc whose rounded quotient num / den
// reaches `best` always satisfies num >= fl(thr * den) (three roundings of 2^-53 each are covered by the 2^-50
// margin), so skipping the others cannot change the maximum or its tie rule; the survivors go through the exact
// division and the (largest bss, smallest i, smallest j) comparison.  7 FP64 operations per arc instead of a
// division, 0.5 shared-memory loads instead of 2 global ones.
constexpr int MA_J = 4;
constexpr int MA_ISTAGE = 512;

__global__ void __launch_bounds__(256)
cbs_maxarc_kernel(const double* __restrict__ sx, const double* __restrict__ cw, const Seg* __restrict__ segs,
                  const Chunk* __restrict__ chunks, int al0, ArcBest* __restrict__ partial) {
  __shared__ double2 s_iv[MA_ISTAGE];
  __shared__ double s_b[256];
  __shared__ int s_i[256], s_j[256];
  const int tid = threadIdx.x;
  const Chunk ch = chunks[blockIdx.x];
  const int64_t lo = segs[ch.seg].lo;
  const int n = (int)(segs[ch.seg].hi - lo);
  const double* psx = sx + lo - 1;  // psx[t] valid for t >= 1
  const double* pcw = cw + lo - 1;
  const double cwn = pcw[n];
  const unsigned span = (unsigned)(n - 2 * al0);  // arc widths d = j - i with al0 <= d <= n - al0
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double best = -1.0, thr = -1.0;
  int bi = 0, bj = 0;
  for (int ib = ch.i0; ib < ch.i1; ib += MA_ISTAGE) {
    const int ie = min(ch.i1, ib + MA_ISTAGE);
    __syncthreads();
    for (int t = tid; t < ie - ib; t += 256) {
      const int i = ib + t;
      s_iv[t] = make_double2(i == 0 ? 0.0 : psx[i], i == 0 ? 0.0 : pcw[i]);
    }
    __syncthreads();
    const int jend = min(n, ie - 1 + n - al0);
    for (int jt = ib + al0; jt <= jend; jt += 256 * MA_J) {
      double sxj[MA_J], cwj[MA_J];
#pragma unroll
      for (int u = 0; u < MA_J; u++) {
        const int j = jt + tid + 256 * u;
        const bool in = j <= jend;
        sxj[u] = in ? psx[j] : nan;  // NaN numerator: never passes the filter
        cwj[u] = in ? pcw[j] : 0.0;
      }
      // start positions that can pair with this tile: j - (n - al0) <= i <= j - al0
      const int i_lo = max(ib, jt - (n - al0));
      const int i_hi = min(ie, jt + 256 * MA_J - al0);
      int dbase = jt + tid - i_lo - al0;  // (j - i) - al0 of slot 0
      for (int i = i_lo; i < i_hi; i++, dbase--) {
        const double2 v = s_iv[i - ib];
#pragma unroll
        for (int u = 0; u < MA_J; u++) {
          const double sd = sxj[u] - v.x;
          const double dw = cwj[u] - v.y;
          const double num = sd * sd;
          const double den = dw * (cwn - dw);
          if ((unsigned)(dbase + 256 * u) <= span && num >= thr * den) {
            const double bss = num / den;
            const int j = jt + tid + 256 * u;
            if (arc_better(bss, i, j, best, bi, bj)) {
              best = bss; bi = i; bj = j;
              thr = best * (1.0 - 0x1p-50);
            }
          }
        }
      }
    }
  }
  s_b[tid] = best; s_i[tid] = bi; s_j[tid] = bj;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const int t = tid + o;
      if (arc_better(s_b[t], s_i[t], s_j[t], s_b[tid], s_i[tid], s_j[tid])) {
        s_b[tid] = s_b[t]; s_i[tid] = s_i[t]; s_j[tid] = s_j[t];
      }
    }
    __syncthreads();
  }
  if (tid == 0) { partial[blockIdx.x].bss = s_b[0]; partial[blockIdx.x].i = s_i[0]; partial[blockIdx.x].j = s_j[0]; }
}

// ------------------------------------------------------------------------------------------
// permutation statistic (hybrid: arcs of at most `max_width` points or their complements;
// max_width < 0: all arcs).  One thread per permutation; scratch layout [n][stride] so that
// the sequential index is coalesced across the permutations of a batch.
// ------------------------------------------------------------------------------------------
struct PermJob {
  int64_t lo;
  int64_t scratch_off;  // first double of this job's scratch: two planes of n * nperm doubles
  int32_t n, max_width;
  uint32_t seed, lo_id, hi_id;
  int32_t perm0;      // first permutation index of this batch
  int32_t nperm;      // permutations of this batch (= scratch stride)
  int32_t pad;
  double ostat, rtw, tot_w, tss_y;
};

// The permutation statistic in three kernels (the one-thread-per-permutation version spent ~70 ms per launch in its
// sequential arc scan whatever the number of permutations, profiles/r01g_launches_config3_with_predict.csv):
//   cbs_perm_prep_kernel   one thread per permutation: Fisher-Yates, prefix sums, re-centring (sequential by
//                          definition: every rounding and the Philox stream order are specified)
//   cbs_perm_arcs_kernel   the arc scan, parallel over (permutation, block of start positions); per-permutation
//                          maximum through atomicMax on the bit pattern (non-negative doubles order like integers,
//                          the initial -1.0 is below all of them) -- a maximum does not depend on the order
//   cbs_perm_count_kernel  statistic of each permutation against the observed one
// Per job the scratch holds two planes [n][nperm] (values, prefix sums; permutation index fastest so that the
// threads of a warp touch consecutive doubles) followed by best[nperm] and tss[nperm].
constexpr int PA_ICHUNK = 64;

__global__ void __launch_bounds__(128)
cbs_perm_prep_kernel(const double* __restrict__ yy, const double* __restrict__ w, const double* __restrict__ cw,
                     const PermJob* __restrict__ jobs, double* __restrict__ scratch_all) {
  const PermJob job = jobs[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= job.nperm) return;
  const int n = job.n;
  const int64_t stride = job.nperm;
  const double* y = yy + job.lo;
  const double* ws = w + job.lo;
  const double* pcw = cw + job.lo - 1;
  double* py = scratch_all + job.scratch_off + p;   // py[i * stride]
  double* sxp = py + (int64_t)n * stride;           // second plane: re-centred prefix sums, index t = 1..n at (t-1)
  double* bestp = scratch_all + job.scratch_off + 2 * (int64_t)n * stride;
  for (int i = 0; i < n; i++) py[(int64_t)i * stride] = y[i];
  PermStream st(job.seed, 0u, job.lo_id, job.hi_id, (uint32_t)(job.perm0 + p));
  // wxperm: Fisher-Yates from the top; px[i] = py[i] / rw[i]; accumulate sum(ws * px) on the fly is
  // not order-compatible with the oracle (which runs cumsum ascending), so store px first
  for (int i = n - 1; i >= 0; i--) {
    const uint32_t j = st.below((uint32_t)i + 1u);
    const double tmp = py[(int64_t)i * stride];
    const double vj = py[(int64_t)j * stride];
    py[(int64_t)j * stride] = tmp;
    py[(int64_t)i * stride] = vj / sqrt(ws[i]);  // position i is final: holds px[i] from now on
  }
  // prefix sums of ws * px (ascending), then re-centre
  double acc = 0.0;
  for (int i = 0; i < n; i++) {
    acc += ws[i] * py[(int64_t)i * stride];
    sxp[(int64_t)i * stride] = acc;
  }
  const double xbar = acc / job.tot_w;
  const double tss = job.tss_y - job.tot_w * xbar * xbar;
  for (int t = 1; t <= n; t++) sxp[(int64_t)(t - 1) * stride] = sxp[(int64_t)(t - 1) * stride] - xbar * (pcw[t] * job.rtw);
  bestp[p] = -1.0;
  bestp[stride + p] = tss;
}

// Same result as cbs_perm_prep_kernel with the Fisher-Yates shuffle in shared memory: one warp per permutation, the
// shuffle runs on a uint16 index array (n <= 65535) -- the lanes draw the 32 Philox numbers of a batch in parallel
// (the stream position of step i is n - 1 - i whatever the order they are computed in), lane 0 applies the 32
// swaps in order -- then px[i] = y[a[i]] / sqrt(w[i]) and the prefix sums in sequence (a warp-wide chain of
// additions, every lane keeping the sum of its own position).  The global-memory version above spends ~3 us per
// element in dependent random accesses (36-61 ms per launch at n = 12 k, profiles/r01h_launches_*).
__global__ void __launch_bounds__(256)
cbs_perm_fy_kernel(const double* __restrict__ yy, const double* __restrict__ w, const double* __restrict__ cw,
                   const PermJob* __restrict__ jobs, double* __restrict__ scratch_all, int warps, int n_stride) {
  extern __shared__ uint16_t fy_smem[];
  const PermJob job = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * warps + warp;
  if (warp >= warps || p >= job.nperm) return;  // warps are independent: no block-wide barrier below
  const int n = job.n;
  const int64_t stride = job.nperm;
  const double* y = yy + job.lo;
  const double* ws = w + job.lo;
  const double* pcw = cw + job.lo - 1;
  double* sxp = scratch_all + job.scratch_off + (int64_t)n * stride + p;  // prefix-sum plane, index t = 1..n at (t-1)
  double* bestp = scratch_all + job.scratch_off + 2 * (int64_t)n * stride;
  uint16_t* a = fy_smem + (size_t)warp * n_stride;
  for (int i = lane; i < n; i += 32) a[i] = (uint16_t)i;
  __syncwarp();
  for (int ib = n - 1; ib >= 0; ib -= 32) {
    const int i = ib - lane;
    uint32_t j = 0;
    if (i >= 0) {
      const uint32_t t = (uint32_t)(n - 1 - i);
      uint32_t buf[4];
      philox4x32(t >> 2, (uint32_t)(job.perm0 + p), job.lo_id, job.hi_id, job.seed, 0u, buf);
      j = (uint32_t)(((uint64_t)buf[t & 3] * (uint32_t)(i + 1)) >> 32);
    }
    const int cnt = min(32, ib + 1);
    // fully unrolled: the 32 shuffles are issued ahead of the swap chain instead of one per (dependent) swap
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const uint32_t jq = __shfl_sync(0xffffffffu, j, q);
      if (lane == 0 && q < cnt) {
        const int iq = ib - q;
        const uint16_t ti = a[iq], tj = a[jq];
        a[jq] = ti;
        a[iq] = tj;
      }
    }
    __syncwarp();
  }
  double acc = 0.0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    double wpx = 0.0;
    if (i < n) {
      const double wsi = ws[i];
      const double px = y[a[i]] / sqrt(wsi);
      wpx = wsi * px;
    }
    double mine = 0.0;
    // lanes past the end hold +0.0, which leaves the running sum bit-identical: always 32 steps, fully unrolled, so
    // that the shuffles run ahead of the chain of ordered additions
#pragma unroll
    for (int q = 0; q < 32; q++) {
      acc = acc + __shfl_sync(0xffffffffu, wpx, q);
      if (lane == q) mine = acc;
    }
    if (i < n) sxp[(int64_t)i * stride] = mine;
  }
  const double xbar = acc / job.tot_w;
  const double tss = job.tss_y - job.tot_w * xbar * xbar;
  for (int t = 1 + lane; t <= n; t += 32) sxp[(int64_t)(t - 1) * stride] = sxp[(int64_t)(t - 1) * stride] - xbar * (pcw[t] * job.rtw);
  if (lane == 0) {
    bestp[p] = -1.0;
    bestp[stride + p] = tss;
  }
}

// max over the arcs (i, j), j in [j0, j1], of the statistic; the division only runs for arcs that can raise `best`
// (same filter and margin as cbs_maxarc_kernel: the result equals the maximum of the rounded quotients)
__device__ __forceinline__ void perm_arcs(const double* __restrict__ sxp, int64_t stride, const double* __restrict__ pcw,
                                          double sxi, double cwi, double cwn, int j0, int j1, double& best, double& thr) {
  for (int j = j0; j <= j1; j++) {
    const double s = sxp[(int64_t)(j - 1) * stride] - sxi;
    const double dw = pcw[j] - cwi;
    const double num = s * s;
    const double den = dw * (cwn - dw);
    if (num >= thr * den) {
      const double bss = num / den;
      if (bss > best) { best = bss; thr = best * (1.0 - 0x1p-50); }
    }
  }
}

// grid (ceil(max nperm / 128), ceil(max n / PA_ICHUNK), jobs)
__global__ void __launch_bounds__(128)
cbs_perm_arcs_kernel(const double* __restrict__ cw, const PermJob* __restrict__ jobs, int al0, double* __restrict__ scratch_all) {
  const PermJob job = jobs[blockIdx.z];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = job.n;
  const int ia = blockIdx.y * PA_ICHUNK;
  if (p >= job.nperm || ia >= n) return;
  const int ib = min(n, ia + PA_ICHUNK);
  const int64_t stride = job.nperm;
  const double* pcw = cw + job.lo - 1;
  const double* sxp = scratch_all + job.scratch_off + (int64_t)n * stride + p;
  double* bestp = scratch_all + job.scratch_off + 2 * (int64_t)n * stride;
  const double cwn = pcw[n];
  double best = -1.0, thr = -1.0;
  const int mw = job.max_width;
  for (int i = ia; i < ib; i++) {
    const double sxi = i == 0 ? 0.0 : sxp[(int64_t)(i - 1) * stride];
    const double cwi = i == 0 ? 0.0 : pcw[i];
    const int jlo = i + al0, jhi = min(n, i + n - al0);
    if (mw < 0) {
      perm_arcs(sxp, stride, pcw, sxi, cwi, cwn, jlo, jhi, best, thr);
    } else {
      // short arcs and, through the complement, long ones: width <= mw or width >= n - mw
      const int j1 = min(jhi, i + mw);
      perm_arcs(sxp, stride, pcw, sxi, cwi, cwn, jlo, j1, best, thr);
      const int j2 = max(max(jlo, i + n - mw), j1 + 1);
      perm_arcs(sxp, stride, pcw, sxi, cwi, cwn, j2, jhi, best, thr);
    }
  }
  if (best >= 0.0) atomicMax(reinterpret_cast<long long*>(bestp + p), __double_as_longlong(best));
}

__global__ void __launch_bounds__(128)
cbs_perm_count_kernel(const PermJob* __restrict__ jobs, const double* __restrict__ scratch_all, int* __restrict__ nrej,
                      uint32_t* __restrict__ flags, int flag_words) {
  const PermJob job = jobs[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= job.nperm) return;
  const double* bestp = scratch_all + job.scratch_off + 2 * (int64_t)job.n * job.nperm;
  const double best = bestp[p], tss = bestp[job.nperm + p];
  const double pstat = best / ((tss - best) / ((double)job.n - 2.0));
  if (job.ostat <= pstat) {
    atomicAdd(nrej + blockIdx.y, 1);
    // WHICH permutations exceed: the sequential boundary (cbs_segment) depends on their order
    atomicOr(flags + (int64_t)blockIdx.y * flag_words + (p >> 5), 1u << (p & 31));
  }
}

// ------------------------------------------------------------------------------------------
// edge t-tests (DNAcopy wtpermp restated for weights)
// ------------------------------------------------------------------------------------------
struct TJob {
  int64_t lo;        // absolute start of the tested sub-range
  int32_t n1, n2;
  uint32_t seed, test, lo_id, hi_id;
  // filled by cbs_tprep_kernel
  double ostat, xbar, wp;
  int32_t m1, pos0, skip;  // skip: 1 -> p = 1 (n1 or n2 == 1), 2 -> p = 0 (|t| > 5 shortcut)
  int32_t nrej;
};

__global__ void cbs_tprep_kernel(const double* __restrict__ xc, const double* __restrict__ w, TJob* __restrict__ jobs, int njobs) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= njobs) return;
  TJob jb = jobs[q];
  const int n1 = jb.n1, n2 = jb.n2, n = n1 + n2;
  jb.nrej = 0;
  if (n1 == 1 || n2 == 1) { jb.skip = 1; jobs[q] = jb; return; }
  const double* x = xc + jb.lo;
  const double* ws = w + jb.lo;
  double w1 = 0, w2 = 0, s1 = 0, s2 = 0, q2 = 0;
  for (int i = 0; i < n1; i++) { w1 += ws[i]; s1 += ws[i] * x[i]; }
  for (int i = n1; i < n; i++) { w2 += ws[i]; s2 += ws[i] * x[i]; }
  for (int i = 0; i < n; i++) q2 += ws[i] * x[i] * x[i];
  const double wt = w1 + w2;
  const double xbar = (s1 + s2) / wt;
  const double tss = q2 - wt * xbar * xbar;
  double ostat, tstat;
  if (n1 <= n2) {
    jb.m1 = n1; jb.pos0 = 0; jb.wp = w1;
    ostat = 0.99999 * fabs(s1 / w1 - xbar);
    tstat = (ostat * ostat) * w1 * wt / w2;
  } else {
    jb.m1 = n2; jb.pos0 = n1; jb.wp = w2;
    ostat = 0.99999 * fabs(s2 / w2 - xbar);
    tstat = (ostat * ostat) * w2 * wt / w1;
  }
  tstat = tstat / ((tss - tstat) / ((double)n - 2.0));
  jb.ostat = ostat; jb.xbar = xbar;
  jb.skip = (tstat > 25.0 && jb.m1 >= 10) ? 2 : 0;
  jobs[q] = jb;
}

__global__ void __launch_bounds__(128)
cbs_tperm_kernel(const double* __restrict__ xc, const double* __restrict__ w, TJob* __restrict__ jobs, int q, int nperm,
                 double* __restrict__ scratch, int stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nperm) return;
  const TJob jb = jobs[q];
  const int n = jb.n1 + jb.n2;
  const double* x = xc + jb.lo;
  const double* ws = w + jb.lo;
  double* py = scratch + p;
  for (int i = 0; i < n; i++) py[(int64_t)i * stride] = x[i] * sqrt(ws[i]);
  PermStream st(jb.seed, jb.test, jb.lo_id, jb.hi_id, (uint32_t)p);
  double acc = 0.0;
  for (int t = 0; t < jb.m1; t++) {
    const int i = n - 1 - t;
    const uint32_t j = st.below((uint32_t)i + 1u);
    const double tmp = py[(int64_t)i * stride];
    const double vj = py[(int64_t)j * stride];
    py[(int64_t)j * stride] = tmp;
    py[(int64_t)i * stride] = vj;
    acc += sqrt(ws[jb.pos0 + t]) * vj;
  }
  const double pstat = fabs(acc / jb.wp - jb.xbar);
  if (jb.ostat <= pstat) atomicAdd(&jobs[q].nrej, 1);
}

// ------------------------------------------------------------------------------------------
// hybrid p-value, part 1: Siegmund's nu(x) on the integration grid of DNAcopy's `tailp`.
// nu(x) = (2 / x^2) exp(-2 sum_{k>=1} Phi(-x sqrt(k) / 2) / k); for the small x of long segments
// the series needs 1e4..1e6 terms per grid point (what makes this step seconds on a CPU), so it is
// summed here: one block per (grid point, job), the doubling stages of the reference loop kept
// (stage sums reduced in parallel), convergence test |delta / lnu| <= tol after every stage.
// out[job][i] = nu(x_i)^2 * integral_{tl_i}^{tl_i + dincr} dt / (t (1 - t))^2
// ------------------------------------------------------------------------------------------
struct TailJob {
  double b;      // sqrt of the observed statistic
  double delta;  // (kmax + 1) / n
  int32_t m;     // segment length
  int32_t pad;
};

__device__ __forceinline__ double dev_pnorm(double x) { return 0.5 * erfc(-x / 1.4142135623730951); }

__device__ __forceinline__ double dev_it1tsq(double x, double a) {
  double y = x + a - 0.5;
  double v = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
  y = x - 0.5;
  return v - (8.0 * y) / (1.0 - 4.0 * y * y) - 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
}

__global__ void __launch_bounds__(256)
cbs_tailp_kernel(const TailJob* __restrict__ jobs, int ngrid, double tol, double* __restrict__ out) {
  __shared__ double sh[8];
  __shared__ double s_stage;
  const TailJob jb = jobs[blockIdx.y];
  const int gi = blockIdx.x;
  const double dincr = (0.5 - jb.delta) / (double)ngrid;
  const double bsqrtm = jb.b / sqrt((double)jb.m);
  const double tl = 0.5 + (double)gi * dincr;            // tl after gi + 1 increments from 0.5 - dincr
  const double t = 0.5 - 0.5 * dincr + (double)(gi + 1) * dincr;
  const double x = bsqrtm / sqrt(t * (1.0 - t));
  double lnu1;
  if (x > 0.01) {
    lnu1 = log(2.0) - 2.0 * log(x);
    double lnu0 = lnu1;
    long long k0 = 0, kn = 2;  // stage covers k in (k0, k0 + kn]; reference stages: 2, then 2, 4, 8, ...
    bool first = true;
    for (;;) {
      double part = 0.0;
      for (long long k = k0 + 1 + threadIdx.x; k <= k0 + kn; k += blockDim.x) {
        const double dk = (double)k;
        part += 2.0 * dev_pnorm(-x * sqrt(dk) / 2.0) / dk;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; w++) tot += sh[w];
        s_stage = tot;
      }
      __syncthreads();
      if (!first) lnu0 = lnu1;
      lnu1 = lnu1 - s_stage;
      k0 += kn;
      if (!(fabs((lnu1 - lnu0) / lnu1) > tol)) break;  // `while (abs((lnu1 - lnu0) / lnu1) > tol)`
      if (first) first = false; else kn *= 2;          // 2 more terms, then the stage size doubles
      if (k0 > (1ll << 40)) break;
    }
  } else {
    lnu1 = -0.583 * x;
  }
  if (threadIdx.x == 0) {
    const double nux = exp(lnu1);
    out[(int64_t)blockIdx.y * ngrid + gi] = (nux * nux) * dev_it1tsq(tl, dincr);
  }
}

// ---------------- host-side scalar maths (DNAcopy tailp / nu / it1tsq) ----------------
double pnorm_(double x) { return 0.5 * std::erfc(-x / std::sqrt(2.0)); }

double nu_(double x, double tol) {
  double lnu1;
  if (x > 0.01) {
    lnu1 = std::log(2.0) - 2.0 * std::log(x);
    double lnu0 = lnu1;
    int k = 2;
    double dk = 0.0;
    for (int i = 0; i < k; i++) { dk += 1.0; lnu1 -= 2.0 * pnorm_(-x * std::sqrt(dk) / 2.0) / dk; }
    while (std::fabs((lnu1 - lnu0) / lnu1) > tol) {
      lnu0 = lnu1;
      for (int i = 0; i < k; i++) { dk += 1.0; lnu1 -= 2.0 * pnorm_(-x * std::sqrt(dk) / 2.0) / dk; }
      k *= 2;
    }
  } else {
    lnu1 = -0.583 * x;
  }
  return std::exp(lnu1);
}

double it1tsq_(double x, double a) {
  double y = x + a - 0.5;
  double v = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * std::log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
  y = x - 0.5;
  return v - (8.0 * y) / (1.0 - 4.0 * y * y) - 2.0 * std::log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
}

double tailp_(double b, double delta, int m, int ngrid, double tol) {
  const double dincr = (0.5 - delta) / ngrid;
  const double bsqrtm = b / std::sqrt((double)m);
  double tl = 0.5 - dincr, t = 0.5 - 0.5 * dincr, acc = 0.0;
  for (int i = 0; i < ngrid; i++) {
    tl += dincr;
    t += dincr;
    const double x = bsqrtm / std::sqrt(t * (1.0 - t));
    const double nux = nu_(x, tol);
    acc += (nux * nux) * it1tsq_(tl, dincr);
  }
  return 9.973557e-2 * (b * b * b) * std::exp(-b * b / 2.0) * acc;
}

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { set_error("cbs: cudaMalloc failed"); return 1; }
    cap = bytes;
    return 0;
  }
  ~DBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct CbsWorkspace {
  DBuf y, w, xc, sx, cw, yy, segs, prep, chunks, partial, scratch, nrej, tjobs, pjobs, flags;
  std::vector<int32_t> sbdry;  // sequential stopping boundary (cbs_set_boundary); empty: plain count over all permutations
};

// DNAcopy's sequential boundary (segment(): sbdry = getbdry(eta, nperm, max.ones), a triangular table: the row for a
// test that tolerates nrejc exceedances starts at nrejc (nrejc + 1) / 2 and holds nrejc + 1 permutation counts).  A
// permutation test is declared significant as soon as np >= sbdry[row + nrej] (fndcpt's inner loop).
void cbs_set_boundary(CbsWorkspace* ws, const int32_t* sbdry, int32_t n) {
  ws->sbdry.assign(sbdry, sbdry + (n > 0 ? n : 0));
}

CbsWorkspace* cbs_workspace_create() { return new CbsWorkspace(); }
void cbs_workspace_destroy(CbsWorkspace* ws) { delete ws; }

int cbs_segment(CbsWorkspace* ws, const double* y, const double* w, const int64_t* off, int32_t nseries,
                const int32_t* series_ids, double alpha, int32_t nperm, int32_t kmax, int32_t nmin, int32_t min_width,
                uint32_t seed, int32_t* ends_out, int32_t* nseg_out, CbsStats* stats, cudaStream_t st) {
  const int64_t total = off[nseries];
  for (int s = 0; s < nseries; s++) nseg_out[s] = 0;
  if (total == 0) return 0;
  if (ws->y.ensure(8 * total) || ws->w.ensure(8 * total) || ws->xc.ensure(8 * total) || ws->sx.ensure(8 * total) ||
      ws->cw.ensure(8 * total) || ws->yy.ensure(8 * total) || ws->nrej.ensure(64))
    return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(ws->y.p, y, 8 * total, cudaMemcpyHostToDevice, st));
  WCX_CUDA_OK(cudaMemcpyAsync(ws->w.p, w, 8 * total, cudaMemcpyHostToDevice, st));
  std::vector<std::vector<int32_t>> ends(nseries);
  std::vector<Seg> pending;
  for (int s = 0; s < nseries; s++)
    if (off[s + 1] > off[s]) pending.push_back(Seg{off[s], off[s + 1], s, 0});
  const int al0 = min_width;
  const int64_t CHUNK_ARCS = 1 << 20;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  while (!pending.empty()) {
    // segments too short to split are final
    std::vector<Seg> work;
    for (const Seg& sg : pending) {
      if (sg.hi - sg.lo >= 2 * (int64_t)min_width) work.push_back(sg);
      else ends[sg.series].push_back((int32_t)(sg.hi - off[sg.series]));
    }
    pending.clear();
    if (work.empty()) break;
    const int nseg = (int)work.size();
    if (stats) { stats->rounds++; stats->segments_tested += nseg; }
    if (ws->segs.ensure(sizeof(Seg) * nseg) || ws->prep.ensure(sizeof(SegPrep) * nseg)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(ws->segs.p, work.data(), sizeof(Seg) * nseg, cudaMemcpyHostToDevice, st));
    cbs_prepare_kernel<<<nseg, PREP_THREADS, 0, st>>>(ws->y.as<double>(), ws->w.as<double>(), ws->segs.as<Seg>(), nseg,
                                                       ws->xc.as<double>(), ws->sx.as<double>(), ws->cw.as<double>(),
                                                       ws->yy.as<double>(), ws->prep.as<SegPrep>());
    // balanced chunks of start positions: start i owns arcs(i) = min(n, i + n - al0) - (i + al0) + 1 end positions,
    // non-increasing in i, so a chunk of width CHUNK_ARCS / arcs(i0) never exceeds the budget
    std::vector<Chunk> chunks;
    for (int s = 0; s < nseg; s++) {
      const int n = (int)(work[s].hi - work[s].lo);
      for (int i0 = 0; i0 < n;) {
        const int64_t a0 = std::max<int64_t>(1, (int64_t)std::min(n, i0 + n - al0) - (i0 + al0) + 1);
        const int wdt = (int)std::max<int64_t>(1, std::min<int64_t>(n - i0, CHUNK_ARCS / a0));
        chunks.push_back(Chunk{s, i0, i0 + wdt, 0});
        i0 += wdt;
      }
    }
    const int nchunks = (int)chunks.size();
    if (ws->chunks.ensure(sizeof(Chunk) * nchunks) || ws->partial.ensure(sizeof(ArcBest) * nchunks)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(ws->chunks.p, chunks.data(), sizeof(Chunk) * nchunks, cudaMemcpyHostToDevice, st));
    cbs_maxarc_kernel<<<nchunks, 256, 0, st>>>(ws->sx.as<double>(), ws->cw.as<double>(), ws->segs.as<Seg>(),
                                               ws->chunks.as<Chunk>(), al0, ws->partial.as<ArcBest>());
    WCX_CUDA_OK(cudaGetLastError());
    std::vector<ArcBest> partial(nchunks);
    std::vector<SegPrep> prep(nseg);
    WCX_CUDA_OK(cudaMemcpyAsync(partial.data(), ws->partial.p, sizeof(ArcBest) * nchunks, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(prep.data(), ws->prep.p, sizeof(SegPrep) * nseg, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));
    if (stats) stats->launches += 2;
    std::vector<ArcBest> best(nseg, ArcBest{-1.0, 0, 0});
    for (int c = 0; c < nchunks; c++) {
      ArcBest& b = best[chunks[c].seg];
      const ArcBest& p = partial[c];
      if ((p.bss > b.bss) || (p.bss == b.bss && (p.i < b.i || (p.i == b.i && p.j < b.j)))) b = p;
    }
    // hybrid p-value, part 1 on the device for every segment that needs it this round
    std::vector<double> pval1_of(nseg, 2.0);
    {
      std::vector<TailJob> tjobs;
      std::vector<int> tseg;
      for (int s = 0; s < nseg; s++) {
        const int n = (int)(work[s].hi - work[s].lo);
        if (prep[s].flat || best[s].bss < 0.0 || n <= nmin) continue;
        const double ostat = best[s].bss / ((prep[s].tss - best[s].bss) / ((double)n - 2.0));
        const double ostat1 = ostat > 0 ? std::sqrt(ostat) : 0.0;
        const int width = best[s].j - best[s].i;
        const int l = std::min(width, n - width);
        if (ostat1 <= 0.1 || (ostat1 >= 7.0 && l >= 10)) continue;
        tjobs.push_back(TailJob{ostat1, ((double)kmax + 1.0) / (double)n, n, 0});
        tseg.push_back(s);
      }
      if (!tjobs.empty()) {
        const int ngrid = 100;
        const int nj = (int)tjobs.size();
        if (ws->tjobs.ensure(sizeof(TailJob) * nj) || ws->scratch.ensure(sizeof(double) * (size_t)nj * ngrid)) return 1;
        WCX_CUDA_OK(cudaMemcpyAsync(ws->tjobs.p, tjobs.data(), sizeof(TailJob) * nj, cudaMemcpyHostToDevice, st));
        cbs_tailp_kernel<<<dim3(ngrid, nj), 256, 0, st>>>(ws->tjobs.as<TailJob>(), ngrid, 1e-6, ws->scratch.as<double>());
        std::vector<double> terms((size_t)nj * ngrid);
        WCX_CUDA_OK(cudaMemcpyAsync(terms.data(), ws->scratch.p, sizeof(double) * terms.size(), cudaMemcpyDeviceToHost, st));
        WCX_CUDA_OK(cudaStreamSynchronize(st));
        if (stats) stats->launches++;
        for (int q = 0; q < nj; q++) {
          double acc = 0.0;
          for (int i = 0; i < ngrid; i++) acc += terms[(size_t)q * ngrid + i];
          const double b = tjobs[q].b;
          pval1_of[tseg[q]] = 9.973557e-2 * (b * b * b) * std::exp(-b * b / 2.0) * acc;
        }
      }
    }
    // decisions, part 1: gates and the hybrid p-value decide most segments; the rest need a permutation test
    std::vector<char> split_of(nseg, 0);
    std::vector<double> ostat_of(nseg, 0.0);
    struct PermTest {
      int seg, nrejc, mw, nrej, done;
      int cur = 0;           // permutations processed by the decision rule (1-based count)
      const int32_t* bd = nullptr;  // this test's row of the boundary table (nrejc + 1 entries) or nullptr
      int limit = 0;         // permutations after which the test is decided for certain
    };
    std::vector<PermTest> tests;
    for (int s = 0; s < nseg; s++) {
      const int n = (int)(work[s].hi - work[s].lo);
      const SegPrep& pr = prep[s];
      if (pr.flat || best[s].bss < 0.0) continue;
      double ostat = best[s].bss / ((pr.tss - best[s].bss) / ((double)n - 2.0));
      const double ostat1 = ostat > 0 ? std::sqrt(ostat) : 0.0;
      ostat *= 0.99999;
      ostat_of[s] = ostat;
      if (!(ostat1 > 0.1)) continue;
      const int width = best[s].j - best[s].i;
      const int l = std::min(width, n - width);
      if (ostat1 >= 7.0 && l >= 10) { split_of[s] = 1; continue; }
      if (n > nmin) {
        const double pval1 = pval1_of[s];
        if (pval1 > alpha) continue;
        tests.push_back(PermTest{s, (int)((alpha - pval1) * (double)nperm), kmax, 0, 0});
      } else {
        tests.push_back(PermTest{s, (int)(alpha * (double)nperm), -1, 0, 0});
      }
      PermTest& t = tests.back();
      const size_t row = (size_t)t.nrejc * ((size_t)t.nrejc + 1) / 2;
      t.limit = nperm;
      if (row + (size_t)t.nrejc < ws->sbdry.size()) {
        t.bd = ws->sbdry.data() + row;
        t.limit = std::max(1, std::min(nperm, (int)t.bd[t.nrejc]));  // with nrej <= nrejc the last boundary decides
      }
    }
    // permutation tests of the whole round in escalating batches (256, 1024, then 2048 at a time): all undecided
    // tests share one launch per stage; a test stops as soon as its rejection count exceeds the threshold, which
    // is the decision the full count would give (the permutation index is the Philox counter, not the batch)
    if (!tests.empty()) {
      if (stats) stats->perm_tests += (int)tests.size();
      std::vector<int> active(tests.size());
      for (size_t q = 0; q < tests.size(); q++) active[q] = (int)q;
      int stage = 0;
      while (!active.empty()) {
        const int want = stage == 0 ? 256 : nperm;  // a first look decides most tests; the rest run to the end
        stage++;
        // groups of tests whose scratch fits the budget
        size_t g0 = 0;
        std::vector<int> still;
        while (g0 < active.size()) {
          std::vector<PermJob> jobs;
          std::vector<int> jq;
          size_t doubles = 0;
          int maxnb = 0, maxn = 0;
          size_t g1 = g0;
          for (; g1 < active.size(); g1++) {
            PermTest& t = tests[active[g1]];
            const Seg& sg = work[t.seg];
            const int n = (int)(sg.hi - sg.lo);
            const int nb = std::min(want, t.limit - t.done);
            const size_t need = 2 * (size_t)n * nb + 2 * (size_t)nb;
            if (!jobs.empty() && (doubles + need) * sizeof(double) > ((size_t)8 << 30)) break;  // scratch budget per launch: 8 GB of the 180
            const int32_t sid = series_ids ? series_ids[sg.series] : sg.series;
            PermJob job;
            job.lo = sg.lo; job.scratch_off = (int64_t)doubles; job.n = n; job.max_width = t.mw;
            job.seed = (uint32_t)((uint64_t)seed * 1000003ull + (uint64_t)(uint32_t)sid);
            job.lo_id = (uint32_t)(sg.lo - off[sg.series]); job.hi_id = (uint32_t)(sg.hi - off[sg.series]);
            job.perm0 = t.done; job.nperm = nb; job.pad = 0;
            job.ostat = ostat_of[t.seg]; job.rtw = prep[t.seg].rtw; job.tot_w = prep[t.seg].tot_w; job.tss_y = prep[t.seg].tss_y;
            jobs.push_back(job);
            jq.push_back(active[g1]);
            doubles += need;
            maxnb = std::max(maxnb, nb);
            maxn = std::max(maxn, n);
          }
          const int nj = (int)jobs.size();
          const int flag_words = (maxnb + 31) / 32;
          if (ws->scratch.ensure(sizeof(double) * doubles) || ws->pjobs.ensure(sizeof(PermJob) * nj) || ws->nrej.ensure(sizeof(int) * nj) ||
              ws->flags.ensure(sizeof(uint32_t) * (size_t)nj * flag_words))
            return 1;
          WCX_CUDA_OK(cudaMemcpyAsync(ws->pjobs.p, jobs.data(), sizeof(PermJob) * nj, cudaMemcpyHostToDevice, st));
          WCX_CUDA_OK(cudaMemsetAsync(ws->nrej.p, 0, sizeof(int) * nj, st));
          WCX_CUDA_OK(cudaMemsetAsync(ws->flags.p, 0, sizeof(uint32_t) * (size_t)nj * flag_words, st));
          // shuffle in shared memory (one warp per permutation) when the index array fits, else in global memory
          const int n_stride = (maxn + 7) & ~7;
          const int fy_warps = maxn <= 65535 ? std::min(8, (200 * 1024) / (2 * n_stride)) : 0;
          if (fy_warps >= 1) {
            const int fy_smem = fy_warps * n_stride * 2;
            static int fy_attr = 0;
            if (fy_smem > fy_attr) {
              WCX_CUDA_OK(cudaFuncSetAttribute(cbs_perm_fy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
              fy_attr = 200 * 1024;
            }
            cbs_perm_fy_kernel<<<dim3((maxnb + fy_warps - 1) / fy_warps, nj), 256, fy_smem, st>>>(
                ws->yy.as<double>(), ws->w.as<double>(), ws->cw.as<double>(), ws->pjobs.as<PermJob>(), ws->scratch.as<double>(), fy_warps, n_stride);
          } else {
            cbs_perm_prep_kernel<<<dim3((maxnb + 127) / 128, nj), 128, 0, st>>>(ws->yy.as<double>(), ws->w.as<double>(), ws->cw.as<double>(),
                                                                              ws->pjobs.as<PermJob>(), ws->scratch.as<double>());
          }
          cbs_perm_arcs_kernel<<<dim3((maxnb + 127) / 128, (maxn + PA_ICHUNK - 1) / PA_ICHUNK, nj), 128, 0, st>>>(
              ws->cw.as<double>(), ws->pjobs.as<PermJob>(), al0, ws->scratch.as<double>());
          cbs_perm_count_kernel<<<dim3((maxnb + 127) / 128, nj), 128, 0, st>>>(ws->pjobs.as<PermJob>(), ws->scratch.as<double>(), ws->nrej.as<int>(),
                                                                               ws->flags.as<uint32_t>(), flag_words);
          WCX_CUDA_OK(cudaGetLastError());
          std::vector<int> h(nj);
          std::vector<uint32_t> hf((size_t)nj * flag_words);
          WCX_CUDA_OK(cudaMemcpyAsync(h.data(), ws->nrej.p, sizeof(int) * nj, cudaMemcpyDeviceToHost, st));
          WCX_CUDA_OK(cudaMemcpyAsync(hf.data(), ws->flags.p, sizeof(uint32_t) * hf.size(), cudaMemcpyDeviceToHost, st));
          WCX_CUDA_OK(cudaStreamSynchronize(st));
          if (stats) stats->launches += 3;
          for (int q = 0; q < nj; q++) {
            PermTest& t = tests[jq[q]];
            if (stats) stats->permutations += jobs[q].nperm;
            // fndcpt's loop over np = 1, 2, ... replayed on the exceedance positions of this batch: an exceedance beyond
            // nrejc ends the test (not significant); np >= bd[nrej] ends it the other way (significant)
            const int batch_end = t.done + jobs[q].nperm;
            int decided = 0;  // 1 significant, -1 not significant
            const uint32_t* fw = hf.data() + (size_t)q * flag_words;
            for (int wd = 0; wd < (jobs[q].nperm + 31) / 32 && !decided; wd++) {
              uint32_t bits = fw[wd];
              while (bits && !decided) {
                const int g = wd * 32 + __builtin_ctz(bits);
                bits &= bits - 1;
                const int e = t.done + g + 1;  // 1-based permutation count of this exceedance
                if (t.bd && e - 1 > t.cur && t.bd[t.nrej] <= e - 1) { decided = 1; break; }  // crossed before it
                t.nrej++;
                t.cur = e;
                if (t.nrej > t.nrejc) { decided = -1; break; }
                if (t.bd && t.bd[t.nrej] <= e) { decided = 1; break; }
              }
            }
            if (!decided && t.bd && batch_end > t.cur && t.bd[t.nrej] <= batch_end) decided = 1;
            t.cur = std::max(t.cur, batch_end);
            t.done = batch_end;
            if (decided < 0) continue;                                           // not significant: no split
            if (decided > 0 || t.done >= nperm) { split_of[t.seg] = 1; continue; }  // significant
            still.push_back(jq[q]);
          }
          g0 = g1;
        }
        active.swap(still);
      }
    }
    // decisions, part 2: change-points of the segments that split.  Arcs strictly inside a segment get two edge
    // t-tests; their preparation (sequential sums, one thread per test) runs once for the whole round
    std::vector<TJob> tj;
    std::vector<int> tj_of(nseg, -1);
    for (int s = 0; s < nseg; s++) {
      if (!split_of[s]) continue;
      const Seg& sg = work[s];
      const int n = (int)(sg.hi - sg.lo);
      const int i1 = best[s].i, i2 = best[s].j;
      if (i2 == n || i1 == 0) continue;
      const int32_t sid = series_ids ? series_ids[sg.series] : sg.series;
      TJob jobs[2];
      std::memset(jobs, 0, sizeof(jobs));
      jobs[0].lo = sg.lo; jobs[0].n1 = i1; jobs[0].n2 = i2 - i1; jobs[0].test = 1;
      jobs[1].lo = sg.lo + i1; jobs[1].n1 = i2 - i1; jobs[1].n2 = n - i2; jobs[1].test = 2;
      for (auto& jb : jobs) {
        jb.seed = (uint32_t)((uint64_t)seed * 1000003ull + (uint64_t)(uint32_t)sid);
        jb.lo_id = (uint32_t)(sg.lo - off[sg.series]);
        jb.hi_id = (uint32_t)(sg.hi - off[sg.series]);
      }
      tj_of[s] = (int)tj.size();
      tj.push_back(jobs[0]);
      tj.push_back(jobs[1]);
    }
    if (!tj.empty()) {
      const int ntj = (int)tj.size();
      if (ws->tjobs.ensure(sizeof(TJob) * ntj)) return 1;
      WCX_CUDA_OK(cudaMemcpyAsync(ws->tjobs.p, tj.data(), sizeof(TJob) * ntj, cudaMemcpyHostToDevice, st));
      cbs_tprep_kernel<<<(ntj + 31) / 32, 32, 0, st>>>(ws->xc.as<double>(), ws->w.as<double>(), ws->tjobs.as<TJob>(), ntj);
      WCX_CUDA_OK(cudaMemcpyAsync(tj.data(), ws->tjobs.p, sizeof(TJob) * ntj, cudaMemcpyDeviceToHost, st));
      WCX_CUDA_OK(cudaStreamSynchronize(st));
      if (stats) stats->launches++;
    }
    for (int s = 0; s < nseg; s++) {
      const Seg& sg = work[s];
      const int n = (int)(sg.hi - sg.lo);
      std::vector<int> cpts;
      const bool split = split_of[s] != 0;
      const int i1 = best[s].i, i2 = best[s].j;
      if (split) {
        if (i2 == n) cpts.push_back(i1);
        else if (i1 == 0) cpts.push_back(i2);
        else {
          for (int q = 0; q < 2; q++) {
            const int gq = tj_of[s] + q;
            const TJob& jb = tj[gq];
            double pval;
            if (jb.skip == 1) pval = 1.0;
            else if (jb.skip == 2) pval = 0.0;
            else {
              const int nn = jb.n1 + jb.n2;
              if (ws->scratch.ensure(sizeof(double) * (size_t)nn * nperm)) return 1;
              cbs_tperm_kernel<<<(nperm + 127) / 128, 128, 0, st>>>(ws->xc.as<double>(), ws->w.as<double>(), ws->tjobs.as<TJob>(), gq,
                                                                    nperm, ws->scratch.as<double>(), nperm);
              TJob back;
              WCX_CUDA_OK(cudaMemcpyAsync(&back, ws->tjobs.as<TJob>() + gq, sizeof(TJob), cudaMemcpyDeviceToHost, st));
              WCX_CUDA_OK(cudaStreamSynchronize(st));
              if (stats) { stats->launches++; stats->permutations += nperm; stats->t_tests++; }
              pval = (double)back.nrej / (double)nperm;
            }
            if (pval <= alpha) cpts.push_back(q == 0 ? i1 : i2);
          }
        }
      }
      if (cpts.empty()) {
        ends[sg.series].push_back((int32_t)(sg.hi - off[sg.series]));
      } else {
        int64_t a = sg.lo;
        for (int c : cpts) { pending.push_back(Seg{a, sg.lo + c, sg.series, 0}); a = sg.lo + c; }
        pending.push_back(Seg{a, sg.hi, sg.series, 0});
      }
    }
  }
  int64_t o = 0;
  for (int s = 0; s < nseries; s++) {
    std::sort(ends[s].begin(), ends[s].end());
    nseg_out[s] = (int32_t)ends[s].size();
    for (int32_t e : ends[s]) ends_out[o++] = e;
  }
  return 0;
}

}  // namespace wcx
