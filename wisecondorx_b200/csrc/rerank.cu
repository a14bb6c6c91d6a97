// Exact re-rank of the approximate candidate lists ("K2b") and the exact brute-force row path.
//
// Reference semantics reproduced bit for bit (newref_tools.py:255-278):
//   * distance d_j = np.sum(np.power(chr_data[j] - row, 2)) in float64 with three separate
//     roundings per term (subtract, square, add) and NumPy's pairwise summation order along the
//     contiguous sample axis -- realised here with __dsub_rn/__dmul_rn/__dadd_rn (no FMA
//     contraction) and a per-S "summation plan" that replays NumPy's recursion tree;
//   * selection = first ref_size entries of a stable ascending sort, i.e. ordered by
//     (distance, position), position = index in the chromosome-excluded array, values >= 1e10
//     never inserted, missing entries (-1, 1e10).
//
// Completeness proof obligation of the fast path.  The sweep kernels hand over, per row, lists
// of approximate values v~_j = |b^_j|^2 - 2<a^,b^_j> (tf32-rounded operands, fp32 accumulate)
// and a cut such that every candidate NOT listed has v~ >= cut.  With eps >= |d~_j - d_j| for
// every relevant j, the exact top-k is contained in { j : v~_j <= v~_(k) + 2 eps }.  The
// kernel evaluates exact distances for that prefix only, and flags the row for the brute-force
// path whenever the prefix is not strictly below the cut or the a-posteriori check
// (exact d_(k) + eps < bound) fails.  Flagged rows are recomputed by exact_rows_kernel, so the
// final indexes never depend on the approximation.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "wcx_common.cuh"

namespace wcx {

// ------------------------------------------------------------------------------------------
// NumPy pairwise-sum plan: postfix program of int32 triples (op, a, b):
//   op 0: LEAF  a = offset, b = length (<= 128)  -> push
//   op 1: ADD   pop r, pop l, push l + r
// ------------------------------------------------------------------------------------------
static void plan_rec(int32_t off, int32_t n, std::vector<int32_t>& p) {
  if (n <= 128) {
    p.push_back(0); p.push_back(off); p.push_back(n);
  } else {
    int32_t n2 = n / 2;
    n2 -= n2 % 8;
    plan_rec(off, n2, p);
    plan_rec(off + n2, n - n2, p);
    p.push_back(1); p.push_back(0); p.push_back(0);
  }
}

int build_sum_plan(int32_t s, int32_t* plan, int32_t cap) {
  std::vector<int32_t> p;
  plan_rec(0, s, p);
  if ((int32_t)p.size() > cap) return -1;
  for (size_t i = 0; i < p.size(); i++) plan[i] = p[i];
  return (int32_t)(p.size() / 3);
}

// (offset, length, #adds that follow) per leaf of the plan + the maximum depth of the combination stack
int plan_to_leaves(const int32_t* plan, int32_t plan_len, std::vector<int32_t>& leaves, int32_t* max_depth) {
  leaves.clear();
  int depth = 0, md = 0;
  for (int i = 0; i < plan_len; i++) {
    if (plan[3 * i] == 0) {
      leaves.push_back(plan[3 * i + 1]);
      leaves.push_back(plan[3 * i + 2]);
      leaves.push_back(0);
      depth++;
      md = depth > md ? depth : md;
    } else {
      if (leaves.empty() || depth < 2) return -1;
      leaves.back() += 1;
      depth--;
    }
  }
  *max_depth = md;
  return (int)(leaves.size() / 3);
}

// ------------------------------------------------------------------------------------------
// Leaf-major ("chain-interleaved") copy of X for the re-rank gather.
//
// The r01d profile of the LDG re-rank shows the L1 data pipe at 88 % (one warp-level LDG.128 of 8 quads touches 8
// cache lines = 8-10 wavefronts per 512 bytes, plus 4 wavefronts per LDS.128 of the target row) with L2 at 37 %.
// Here every row of X is stored once more in the order the summation consumes it: a leaf of NumPy's pairwise tree
// (<= 128 terms, accumulator chain c = element index mod 8) becomes `steps` units of 128 bytes; unit t holds, for
// chain c = 0..7, the chain's elements 2t and 2t+1 (16 bytes per chain).  Eight lanes (one per chain) then read one
// full 128-byte line per step (4 lines = 4 wavefronts per warp-level LDG.128), every lane adds its own chain in
// NumPy's order, and the target row's chain values of the leaf sit in registers while the lane group walks its
// candidates, so the target row costs 8 LDS.128 per leaf instead of one per candidate block.
// Padding (odd block count, tail padded to even) is 0.0 in every row: (0 - 0)^2 = +0 added to a non-negative
// accumulator leaves it bit-identical.  The < 8 tail terms of a leaf follow its units and are added in sequence.
// (A padding-free layout -- a pure permutation of the row, half units for odd block counts, 500 instead of 528 doubles
// at S = 500 -- was measured in round 2: 5 % fewer bytes but 41.8 instead of 38.2 ms, because units and rows no longer
// start on 128-byte lines and every 128-byte unit read touches two lines.  The alignment is worth more than the bytes.)
// desc: 4 int32 per leaf = (offset in doubles, steps, tail terms, 0).  Returns the permuted row length.
// ------------------------------------------------------------------------------------------
int build_leaf_layout(const int32_t* plan, int32_t plan_len, std::vector<int32_t>& perm, std::vector<int32_t>& desc) {
  perm.clear();
  desc.clear();
  for (int op = 0; op < plan_len; op++) {
    if (plan[3 * op] != 0) continue;
    const int off = plan[3 * op + 1], len = plan[3 * op + 2];
    const int nblk = len >> 3, steps = (nblk + 1) >> 1, tail = len - (nblk << 3);
    if (steps > 8) return -1;
    desc.push_back((int32_t)perm.size());
    desc.push_back(steps);
    desc.push_back(tail);
    desc.push_back(0);
    for (int t = 0; t < steps; t++)
      for (int c = 0; c < 8; c++) {
        perm.push_back(off + 8 * (2 * t) + c);
        perm.push_back(2 * t + 1 < nblk ? off + 8 * (2 * t + 1) + c : -1);
      }
    for (int i = 0; i < tail; i++) perm.push_back(off + 8 * nblk + i);
    while (perm.size() % 16) perm.push_back(-1);  // every leaf starts on a 128-byte line
  }
  return (int)perm.size();
}

namespace {

// xp[r][p] = perm[p] >= 0 ? x[r][perm[p]] : 0.0
__global__ void __launch_bounds__(256)
permute_rows_kernel(const double* __restrict__ x, int64_t n, int32_t s, const int32_t* __restrict__ perm, int32_t sp,
                    double* __restrict__ xp) {
  for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
    const double* xr = x + r * s;
    double* o = xp + r * sp;
    for (int p = threadIdx.x; p < sp; p += 256) {
      const int q = perm[p];
      o[p] = q >= 0 ? xr[q] : 0.0;
    }
  }
}

}  // namespace

int launch_permute_rows(const double* x, int64_t n, int32_t s, const int32_t* perm_dev, int32_t sp, double* xp, cudaStream_t st) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>(n, 148 * 16);
  permute_rows_kernel<<<grid, 256, 0, st>>>(x, n, s, perm_dev, sp, xp);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

namespace {

constexpr int RR_THREADS = 256;
constexpr int RR_MAXM = 1024;  // max candidates whose exact distance is evaluated per row

__device__ __forceinline__ uint64_t f64_key(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ uint32_t f32_key_(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32_(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Exact reference distance between the target row `a` (shared memory) and candidate row `b`
// (global), evaluated cooperatively by the 4 lanes of a quad (l = lane & 3); lane l owns the
// NumPy accumulator chains r[2l], r[2l+1].  All four lanes return the same value.
// VEC: rows are 16-byte aligned (S even) -> one 128-bit load per 8-element block.
template <bool VEC>
__device__ __forceinline__ void load_pair(const double* p, double& x0, double& x1) {
  if (VEC) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    x0 = v.x; x1 = v.y;
  } else {
    x0 = __ldg(p); x1 = __ldg(p + 1);
  }
}

// target row from shared memory: one 128-bit load per pair when the row is 16-byte aligned (half the wavefronts)
template <bool VEC>
__device__ __forceinline__ void lds_pair(const double* p, double& x0, double& x1) {
  if (VEC) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    x0 = v.x; x1 = v.y;
  } else {
    x0 = p[0]; x1 = p[1];
  }
}

template <bool VEC>
__device__ __forceinline__ double exact_sqdist_quad(const double* __restrict__ a, const double* __restrict__ b,
                                                    const int32_t* __restrict__ plan, int plan_len, int l) {
  // combination stack in registers (s0 = top; shifting instead of indexing keeps it out of local memory)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
  for (int op = 0; op < plan_len; op++) {
    const int code = plan[3 * op];
    if (code == 0) {
      const int off = plan[3 * op + 1], len = plan[3 * op + 2];
      double res;
      if (len < 8) {
        res = 0.0;
        for (int i = 0; i < len; i++) {
          double t = __dsub_rn(__ldg(b + off + i), a[off + i]);
          res = __dadd_rn(res, __dmul_rn(t, t));
        }
      } else {
        const int nblk = len >> 3;
        const double* bp = b + off + 2 * l;
        const double* ap = a + off + 2 * l;
        double x0, x1, a0, a1;
        load_pair<VEC>(bp, x0, x1);
        lds_pair<VEC>(ap, a0, a1);
        double t0 = __dsub_rn(x0, a0);
        double t1 = __dsub_rn(x1, a1);
        double r0 = __dmul_rn(t0, t0), r1 = __dmul_rn(t1, t1);
        int blk = 1;
        // batches of 5 blocks: all loads first (memory-level parallelism), then the ordered adds
        for (; blk + 5 <= nblk; blk += 5) {
          double y0[5], y1[5];
#pragma unroll
          for (int u = 0; u < 5; u++) load_pair<VEC>(bp + 8 * (blk + u), y0[u], y1[u]);
#pragma unroll
          for (int u = 0; u < 5; u++) {
            lds_pair<VEC>(ap + 8 * (blk + u), a0, a1);
            const double u0 = __dsub_rn(y0[u], a0);
            const double u1 = __dsub_rn(y1[u], a1);
            r0 = __dadd_rn(r0, __dmul_rn(u0, u0));
            r1 = __dadd_rn(r1, __dmul_rn(u1, u1));
          }
        }
        for (; blk < nblk; blk++) {
          double y0, y1;
          load_pair<VEC>(bp + 8 * blk, y0, y1);
          lds_pair<VEC>(ap + 8 * blk, a0, a1);
          const double u0 = __dsub_rn(y0, a0);
          const double u1 = __dsub_rn(y1, a1);
          r0 = __dadd_rn(r0, __dmul_rn(u0, u0));
          r1 = __dadd_rn(r1, __dmul_rn(u1, u1));
        }
        double s1 = __dadd_rn(r0, r1);
        double s2 = __dadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, 1));
        res = __dadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, 2));
        for (int i = nblk << 3; i < len; i++) {
          double t = __dsub_rn(__ldg(b + off + i), a[off + i]);
          res = __dadd_rn(res, __dmul_rn(t, t));
        }
      }
      s7 = s6; s6 = s5; s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = res;
    } else {
      s0 = __dadd_rn(s1, s0);  // left + right
      s1 = s2; s2 = s3; s3 = s4; s4 = s5; s5 = s6; s6 = s7;
    }
  }
  return s0;
}

// sort (d, pos) pairs ascending by (d, pos); arrays in shared memory, n_pow2 entries
__device__ void bitonic_sort_dpos(uint64_t* dkey, int32_t* pos, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = dkey[i], b = dkey[ixj];
          int32_t pa = pos[i], pb = pos[ixj];
          bool gt = (a > b) || (a == b && pa > pb);
          bool up = ((i & k) == 0);
          if (gt == up) { dkey[i] = b; dkey[ixj] = a; pos[i] = pb; pos[ixj] = pa; }
        }
      }
      __syncthreads();
    }
  }
}

// (Round 2 also measured ordering by rank counting -- every thread counts the pairs that order before its own with a
// four-instruction borrow chain per compare, no barriers, twice the instructions: 49.8 / 50.2 ms against 49.9 / 50.4 ms
// for re-rank + null ratios, no difference; this network stays.)
// Same sort with the elements in registers: element i = e * RR_THREADS + tid lives in slot e of thread tid.  A
// compare-exchange distance j < 32 is a warp shuffle, j >= RR_THREADS pairs two slots of one thread, only
// j = 32, 64, 128 go through shared memory (9 of the 45 steps at 512 entries) -- the r01d profile shows the
// barrier-per-step shared-memory version costing ~9k L1 wavefronts and 45 barriers per row.
__device__ __forceinline__ bool dpos_less(uint64_t ka, int32_t pa, uint64_t kb, int32_t pb) {
  return (ka < kb) || (ka == kb && pa < pb);
}
__device__ __forceinline__ void dpos_cswap(uint64_t& ka, int32_t& pa, uint64_t& kb, int32_t& pb, bool up) {
  const bool gt = dpos_less(kb, pb, ka, pa);
  if (gt == up) {
    const uint64_t tk = ka; ka = kb; kb = tk;
    const int32_t tp = pa; pa = pb; pb = tp;
  }
}
__device__ void bitonic_sort_dpos_regs(uint64_t* dkey, int32_t* pos, int n_pow2) {
  const int tid = threadIdx.x;
  const int E = n_pow2 > RR_THREADS ? n_pow2 / RR_THREADS : 1;  // 1, 2 or 4 (RR_MAXM = 4 * RR_THREADS)
  uint64_t kk[4];
  int32_t pp[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int i = e * RR_THREADS + tid;
    const bool valid = e < E && i < n_pow2;
    kk[e] = valid ? dkey[i] : ~0ull;
    pp[e] = valid ? pos[i] : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= RR_THREADS) {
        // both elements in this thread; the lower one (slot without bit j / RR_THREADS) decides the direction
        if (j == RR_THREADS) {
          dpos_cswap(kk[0], pp[0], kk[1], pp[1], ((tid) & k) == 0);
          if (E > 2) dpos_cswap(kk[2], pp[2], kk[3], pp[3], ((2 * RR_THREADS + tid) & k) == 0);
        } else {
          dpos_cswap(kk[0], pp[0], kk[2], pp[2], ((tid) & k) == 0);
          dpos_cswap(kk[1], pp[1], kk[3], pp[3], ((RR_THREADS + tid) & k) == 0);
        }
      } else {
        if (j >= 32) {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int i = e * RR_THREADS + tid;
            if (e < E && i < n_pow2) { dkey[i] = kk[e]; pos[i] = pp[e]; }
          }
          __syncthreads();
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
          if (e < E) {
            const int i = e * RR_THREADS + tid;
            uint64_t ok;
            int32_t op;
            if (j >= 32) {
              const bool valid = i < n_pow2;
              ok = valid ? dkey[i ^ j] : ~0ull;
              op = valid ? pos[i ^ j] : 0x7fffffff;
            } else {
              ok = __shfl_xor_sync(0xffffffffu, kk[e], j);
              op = __shfl_xor_sync(0xffffffffu, pp[e], j);
            }
            const bool keep_min = (((i & j) == 0) == ((i & k) == 0));
            const bool less_o = dpos_less(ok, op, kk[e], pp[e]);
            if (keep_min == less_o) { kk[e] = ok; pp[e] = op; }  // equal pairs are identical: either choice is the same
          }
        }
        if (j >= 32) __syncthreads();
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int i = e * RR_THREADS + tid;
    if (e < E && i < n_pow2) { dkey[i] = kk[e]; pos[i] = pp[e]; }
  }
  __syncthreads();
}

// One shared-memory atomic per warp instead of one per passing lane: all 32 lanes call, `pred` lanes get consecutive
// slots (the order inside a list does not matter to its readers).
__device__ __forceinline__ int warp_slot(int* counter, bool pred) {
  const uint32_t m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Rigorous bound on |d~ - d| (d~ = |a^|^2 + v~, the value the sweep formed; d the squared distance of the unrounded
// rows), all in the units of the list values.
//   an = |a^|^2, nb >= |b^|, e >= |a - a^| + |b - b^|, sD >= min(|a - b|, |a^ - b^|)
// | |a^-b^| - |a-b| | <= e gives | |a^-b^|^2 - |a-b|^2 | <= 2 sD e + e^2 whichever of the two distances sD bounds;
// the rest is arithmetic: fp32 accumulation of the K products (exact products of the rounded operands), the fp32
// roundings of the two norms and of v = |b^|^2 - 2 acc.  The factor 1.5 is slack on top of the worst case.
__device__ __forceinline__ double eps_from(double an, double nb, double e, double sD, int k_pad) {
  const double na = sqrt(an);
  const double rounding = 2.0 * sD * e + e * e;
  const double gamma = ((double)k_pad + 64.0) * 1.1920929e-7;  // fp32 accumulation, (k_pad + 64) * 2^-23
  const double accum = 2.0 * gamma * na * nb;
  const double misc = 4.77e-7 * (an + nb * nb + 2.0 * na * nb);  // fp32 norms + final fma rounding
  return 1.5 * (rounding + accum + misc) + 1e-300;
}

// One bound for every candidate whose squared distance (exact or approximate) is at most Dq, seen from a row with
// |a^|^2 = an: |b^| <= |a^| + sqrt(Dq).  Conversion residuals: worst case u |x| per row (ra < 0), or -- f16 operands --
// the row's own measured residual ra and rho |b^| + tau for the candidate (PrepView::rho_max: the largest measured
// ratio of the matrix; typically 0.4-0.5 u, which is what shrinks the evaluated prefix from ~1.29 k to ~1.15 k).
__device__ __forceinline__ double approx_eps(double an, double Dq, int k_pad, double abs_err, double ra, double rho) {
  const double u = 4.8852e-4;  // 2^-11 * (1 + 2^-10): f16 / tf32 round-to-nearest of an fp32-rounded value
  const double na = sqrt(an);
  const double sD = sqrt(Dq);
  const double nb = na + sD;
  const double tau = 2.0 * sqrt((double)k_pad) * abs_err;  // subnormal / flush-to-zero floor of both operands
  const double e = (ra >= 0.0 ? ra + fmin(rho, u) * nb + tau : u * (na + nb)) + tau;
  return eps_from(an, nb, e, sD, k_pad);
}

// Bound for ONE listed candidate from its own norm and measured residual (normres) and its approximate value v.
__device__ __forceinline__ double cand_eps(double an, double ra, float2 nr, double v, int k_pad, double abs_err) {
  const double tau = 2.0 * sqrt((double)k_pad) * abs_err;
  const double nb = sqrt((double)nr.x);
  const double e = ra + (double)nr.y + tau;
  const double sD = sqrt(fmax(v + an, 0.0) * 1.02 + 1e-300);  // |a^ - b^|^2 = v + an up to the arithmetic terms (2 % covers them: checked by `bound`)
  return eps_from(an, nb, e, sD, k_pad);
}

// ---- leaf-major evaluation (see build_leaf_layout) ------------------------------------------------------------

// volatile: the 2 x STEPS loads of a candidate pair are issued back to back, before the first dependent add
__device__ __forceinline__ double2 ldg_nc_d2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// One leaf (STEPS units of 128 bytes + `tail` sequential terms at row offset `off`) of the candidates
// sel[0..cn): 8 lanes (one per accumulator chain, c8 = lane & 7) per candidate, 32 candidates per pass; PAIR: two
// passes (A, B) in flight and the target row's unit (one LDS.128) shared by the pair.  Leaf sums go to
// leafres[candidate * nleaves + lf].
template <int STEPS, bool PAIR>
__device__ __forceinline__ void leaf_pass(const double* __restrict__ x, int sp, const double* __restrict__ a_s,
                                          const int32_t* __restrict__ sel, int cn, int off, int tail, int nleaves, int lf,
                                          double* __restrict__ leafres, int tid) {
  constexpr int NV = STEPS > 0 ? STEPS : 1;
  constexpr int G = RR_THREADS / 8;
  const int lane = tid & 31, c8 = tid & 7, grp = tid >> 3;
  const double2* ap = reinterpret_cast<const double2*>(a_s + off) + c8;  // unit t: ap[8 * t]
  const int toff = off + STEPS * 16 + c8;                                // this lane's tail term (c8 < tail)
  const double a_tail = c8 < tail ? a_s[toff] : 0.0;
  const int base = lane & ~7;
  for (int c0 = 0; c0 < cn; c0 += (PAIR ? 2 : 1) * G) {
    const int ciA = c0 + grp, ciB = ciA + G;
    // clamped: idle groups redo the last candidate, which keeps the warp converged for the shuffles
    const double* rowA = x + (int64_t)sel[ciA < cn ? ciA : cn - 1] * sp;
    const double* rowB = PAIR ? x + (int64_t)sel[ciB < cn ? ciB : cn - 1] * sp : rowA;
    const double2* pa = reinterpret_cast<const double2*>(rowA + off) + c8;
    const double2* pb = reinterpret_cast<const double2*>(rowB + off) + c8;
    double2 va[NV], vb[NV];
#pragma unroll
    for (int t = 0; t < STEPS; t++) va[t] = ldg_nc_d2(pa + 8 * t);
    if (PAIR) {
#pragma unroll
      for (int t = 0; t < STEPS; t++) vb[t] = ldg_nc_d2(pb + 8 * t);
    }
    double ta = 0.0, tb = 0.0;
    if (c8 < tail) {
      ta = __ldg(rowA + toff);
      if (PAIR) tb = __ldg(rowB + toff);
    }
    double rA = 0.0, rB = 0.0;
#pragma unroll
    for (int t = 0; t < STEPS; t++) {
      const double2 av = ap[8 * t];
      const double uA0 = __dsub_rn(va[t].x, av.x);
      rA = __dadd_rn(rA, __dmul_rn(uA0, uA0));
      if (PAIR) {
        const double uB0 = __dsub_rn(vb[t].x, av.x);
        rB = __dadd_rn(rB, __dmul_rn(uB0, uB0));
      }
      const double uA1 = __dsub_rn(va[t].y, av.y);
      rA = __dadd_rn(rA, __dmul_rn(uA1, uA1));
      if (PAIR) {
        const double uB1 = __dsub_rn(vb[t].y, av.y);
        rB = __dadd_rn(rB, __dmul_rn(uB1, uB1));
      }
    }
    // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), identical in all 8 lanes
    double sA = __dadd_rn(rA, __shfl_xor_sync(0xffffffffu, rA, 1));
    sA = __dadd_rn(sA, __shfl_xor_sync(0xffffffffu, sA, 2));
    sA = __dadd_rn(sA, __shfl_xor_sync(0xffffffffu, sA, 4));
    double sB = 0.0;
    if (PAIR) {
      sB = __dadd_rn(rB, __shfl_xor_sync(0xffffffffu, rB, 1));
      sB = __dadd_rn(sB, __shfl_xor_sync(0xffffffffu, sB, 2));
      sB = __dadd_rn(sB, __shfl_xor_sync(0xffffffffu, sB, 4));
    }
    if (tail > 0) {
      ta = __dsub_rn(ta, a_tail);
      tb = __dsub_rn(tb, a_tail);
      const double qa = __dmul_rn(ta, ta), qb = __dmul_rn(tb, tb);
      for (int i = 0; i < tail; i++) {
        sA = __dadd_rn(sA, __shfl_sync(0xffffffffu, qa, base + i));
        if (PAIR) sB = __dadd_rn(sB, __shfl_sync(0xffffffffu, qb, base + i));
      }
    }
    if (c8 == 0) {
      if (ciA < cn) leafres[ciA * nleaves + lf] = sA;
      if (PAIR && ciB < cn) leafres[ciB * nleaves + lf] = sB;
    }
  }
}

// combination of the leaf sums of one candidate in NumPy's order (register stack, as exact_sqdist_quad)
__device__ __forceinline__ double combine_leaves(const double* __restrict__ lr, const int32_t* __restrict__ plan, int plan_len) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
  int li = 0;
  for (int op = 0; op < plan_len; op++) {
    if (plan[3 * op] == 0) {
      s7 = s6; s6 = s5; s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = lr[li++];
    } else {
      s0 = __dadd_rn(s1, s0);
      s1 = s2; s2 = s3; s3 = s4; s4 = s5; s5 = s6; s6 = s7;
    }
  }
  return s0;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// rerank kernel: one CTA per target row
// ------------------------------------------------------------------------------------------
// shared memory: a[S] doubles | keys[maxc] u64 (later: exact distance keys) | sel[RR_MAXM] i32 | selv[RR_MAXM] f32 |
//                pos[RR_MAXM] i32 | plan
// LEAF: `x` is the leaf-major copy (row stride `sp` doubles, build_leaf_layout), `leaf_g` its descriptors.
// NULLS: the null ratios of the row (newref_tools.py:210-224) are computed by the same CTA right after its indexes
// are final -- the gather-latency-bound re-rank and the issue-bound median selection of different CTAs of an SM
// then overlap (stand-alone they ran back to back at ~50 % and ~75 % issue utilisation).  xm / m_total / null_out as
// in null_ratios.cu; rows that are not certified leave their null ratios to the caller.
template <bool VEC, bool LEAF>
__global__ void __launch_bounds__(RR_THREADS, 4)  // min 4 CTAs/SM keeps the batched loads in flight (64 regs)
rerank_kernel(const double* __restrict__ x, PrepView pv, CandView cv, int nlists, int maxc, const int64_t* __restrict__ cum,
              int nchr, int64_t row_begin, int k, int gonosomal, int32_t* __restrict__ idx_out,
              double* __restrict__ dist_out, int32_t* __restrict__ fail_flags, const int32_t* __restrict__ plan_g,
              int plan_len, int sp, const int32_t* __restrict__ leaf_g, int nleaves) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int row_len = LEAF ? sp : pv.s;  // doubles per row of `x`
  double* a_s = reinterpret_cast<double*>(rr_smem);
  uint32_t* vals = reinterpret_cast<uint32_t*>(a_s + ((row_len + 1) & ~1));  // [maxc] orderable keys of the approximate values
  uint32_t* jidx = vals + maxc;                                           // [maxc] candidate bins
  int32_t* sel = reinterpret_cast<int32_t*>(jidx + maxc);                 // [RR_MAXM]
  float* selv = reinterpret_cast<float*>(sel + RR_MAXM);                  // [RR_MAXM] approximate values of the selected
  int32_t* plan = sel + 2 * RR_MAXM;
  // LEAF: leaf descriptors and the per-candidate leaf sums [RR_MAXM][nleaves] follow the plan
  int32_t* leaf_s = plan + ((3 * plan_len + 3) & ~3);
  // vals / jidx are dead once `sel` is built: the exact (distance, position) pairs reuse the space
  uint64_t* keys = reinterpret_cast<uint64_t*>(vals);  // [RR_MAXM]
  int32_t* pos_s = reinterpret_cast<int32_t*>(keys + RR_MAXM);
  // LEAF: per-candidate leaf sums of a batch of candidates, in the rest of the dead list region
  double* leafres = reinterpret_cast<double*>(pos_s + RR_MAXM);
  const int leafres_bytes = maxc * 8 - RR_MAXM * 12;
  __shared__ int s_tot, s_m, s_fail, s_cnt;
  __shared__ int s_cs, s_ce;
  __shared__ float s_cut;
  __shared__ uint32_t s_vk, s_mn, s_mx;
  __shared__ unsigned long long s_ub;

  const int tid = threadIdx.x;
  const int64_t lrow = blockIdx.x;
  const int64_t row = row_begin + lrow;
  int32_t* oi = idx_out + lrow * k;
  double* od = dist_out + lrow * k;

  if (tid == 0) {
    int c = 0;
    while (c < nchr && cum[c] <= row) c++;
    s_cs = (int)(c == 0 ? 0 : cum[c - 1]);
    s_ce = (int)cum[c];
    s_fail = 0;
    s_tot = 0;
    s_mn = 0xffffffffu;
    s_mx = 0u;
    if (gonosomal && c != 22 && c != 23) s_cs = -1;
    float cut = __int_as_float(0x7f800000);
    for (int q = 0; q < nlists; q++) cut = fminf(cut, cv.cut[lrow * nlists + q]);
    s_cut = cut;
  }
  __syncthreads();
  if (s_cs < 0) {  // placeholder rows of gonosomal references (newref_tools.py:186-191)
    for (int t = tid; t < k; t += RR_THREADS) { oi[t] = 0; od[t] = 1.0; }
    return;
  }
  const int cs = s_cs, ce = s_ce;
  const float cut = s_cut;
  for (int i = tid; i < row_len; i += RR_THREADS) a_s[i] = x[row * row_len + i];
  for (int i = tid; i < 3 * plan_len; i += RR_THREADS) plan[i] = plan_g[i];
  if (LEAF)
    for (int i = tid; i < 4 * nleaves; i += RR_THREADS) leaf_s[i] = leaf_g[i];

  // gather the list entries below the common cut
  uint32_t mn = 0xffffffffu, mx = 0u;
  for (int q = 0; q < nlists; q++) {
    const int64_t slot = lrow * nlists + q;
    const int c = cv.cnt[slot];
    const uint2* le = cv.ent + slot * WCX_CAND_CAP;
    // four independent loads in flight per thread; the loop bound is uniform so that whole warps allocate their list
    // slots together (warp_slot)
    for (int b0 = 0; b0 < c; b0 += 4 * RR_THREADS) {
      uint2 e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = b0 + tid + u * RR_THREADS;
        e[u] = i < c ? le[i] : make_uint2(0x7f800000u, 0u);  // +inf is never below the cut
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float v = __uint_as_float(e[u].x);
        const bool pass = v < cut;
        const int p = warp_slot(&s_tot, pass);
        if (pass) {
          const uint32_t key = f32_key_(v);
          mn = min(mn, key);
          mx = max(mx, key);
          if (p < maxc) { vals[p] = key; jidx[p] = e[u].y; }
        }
      }
    }
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((tid & 31) == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
  __syncthreads();
  const int tot = s_tot;
  if (tot > maxc) {
    if (tid == 0) fail_flags[lrow] = 1;
    for (int t = tid; t < k; t += RR_THREADS) oi[t] = 0;  // valid until exact_rows rewrites the row
    return;
  }

  // k-th smallest approximate value: bisection over the key bits below the prefix shared by all keys; stops as
  // soon as the bracket holds a single key (about log2(tot) + 2 rounds instead of 32)
  if (tot > k && tid < 32) {
    // one warp, no block-wide barriers: ~12 rounds of tot / 32 compares per lane and a warp reduction
    const uint32_t diff = s_mn ^ s_mx;
    int bit = diff ? 31 - __clz(diff) : -1;  // highest differing bit
    uint32_t res = bit >= 31 ? 0u : (s_mn & ~((2u << (bit < 0 ? 0 : bit)) - 1u));
    if (bit < 0) res = s_mn;
    int below = 0, inb = tot;  // keys below res / inside the bracket [res, res + 2^(bit + 1))
    for (; bit >= 0 && inb > 1; bit--) {
      const uint32_t trial = res | (1u << bit);
      int c = 0;
      for (int i = tid; i < tot; i += 32) c += (vals[i] < trial) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c < k) { res = trial; inb = below + inb - c; below = c; }
      else inb = c - below;
    }
    if (bit >= 0) {
      // exactly one key left in the bracket and it is the k-th smallest: fetch it
      const uint32_t hi_excl = bit >= 31 ? 0xffffffffu : res + ((2u << bit) - 1u);  // inclusive upper end
      for (int i = tid; i < tot; i += 32) {
        const uint32_t v = vals[i];
        if (v >= res && v <= hi_excl) s_vk = v;
      }
    } else if (tid == 0) {
      s_vk = res;
    }
  }
  if (tid == 0) {
    s_m = 0;
    s_cnt = 0;
  }
  __syncthreads();
  double bound = 1e300, eps = 0.0;
  bool eps_ok = true;
  const bool meas = pv.normres != nullptr;
  const double an = (double)pv.norm[row];
  const double ra = meas ? (double)pv.normres[row].y : -1.0;
  if (tot > k) {
    const float vk = key_f32_(s_vk);
    const double D = fmax((double)vk + an, 0.0);
    const double rho = meas ? (double)pv.rho_max[0] : 0.0;
    // eps must hold for the approximate top-k (d~ <= D) AND for the exact top-k (d <= D + eps): fixed point from 1.02 D
    double Dq = D * 1.02 + 1e-300;
    for (int it = 0; it < 4; it++) {
      eps = approx_eps(an, Dq, pv.k_pad, (double)pv.abs_err, ra, rho);
      if (D + eps <= Dq) break;
      Dq = (D + eps) * 1.05;
    }
    eps_ok = D + eps <= Dq;  // false also for NaN
    bound = (double)vk + 2.0 * eps;
  }
  // every unlisted candidate has v >= cut: the prefix {v <= bound} must lie strictly below it
  if (!eps_ok || !((double)cut > bound)) {
    if (tid == 0) fail_flags[lrow] = 2;
    for (int t = tid; t < k; t += RR_THREADS) oi[t] = 0;  // valid until exact_rows rewrites the row
    return;
  }
  // select the candidates with v <= bound
  if (tid == 0) s_ub = 0ull;
  for (int b0 = 0; b0 < tot; b0 += RR_THREADS) {
    const int i = b0 + tid;
    const float v = i < tot ? key_f32_(vals[i]) : __int_as_float(0x7f800000);
    const bool pass = i < tot && (double)v <= bound;
    const int p = warp_slot(&s_m, pass);
    if (pass && p < RR_MAXM) { sel[p] = (int32_t)jidx[i]; selv[p] = v; }
  }
  __syncthreads();
  int m = s_m;
  if (m > RR_MAXM) {
    if (tid == 0) fail_flags[lrow] = 3;
    for (int t = tid; t < k; t += RR_THREADS) oi[t] = 0;  // valid until exact_rows rewrites the row
    return;
  }
  if (meas && tot > k) {
    // Per-candidate refinement.  With eps_j >= |v_j - (d_j - |a|^2)| for every selected j: the >= k candidates with
    // v_j <= v_(k) all have d_j - |a|^2 <= U = max(v_j + eps_j) over them, so d_(k) - |a|^2 <= U, and a member of the
    // exact top-k has v_j - eps_j <= d_j - |a|^2 <= U.  The exact top-k lies inside the selection (bound above), so
    // only the selected candidates with v_j - eps_j <= U need their exact distance.
    const float vk = key_f32_(s_vk);
    int jj[RR_MAXM / RR_THREADS];
    double lo[RR_MAXM / RR_THREADS];
    double umax = -1e300;
#pragma unroll
    for (int e = 0; e < RR_MAXM / RR_THREADS; e++) {
      const int i = tid + e * RR_THREADS;
      jj[e] = -1;
      lo[e] = 0.0;
      if (i < m) {
        const int j = sel[i];
        const float v = selv[i];
        const float2 nr = __ldg(pv.normres + j);
        const double ej = cand_eps(an, ra, nr, (double)v, pv.k_pad, (double)pv.abs_err);
        jj[e] = j;
        lo[e] = (double)v - ej;
        if (v <= vk) umax = fmax(umax, (double)v + ej);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
    if ((tid & 31) == 0) atomicMax(&s_ub, f64_key(umax));
    if (tid == 0) s_cnt = 0;
    __syncthreads();  // all reads of sel are done
    const uint64_t ubk = s_ub;
    const uint64_t ubu = (ubk & 0x8000000000000000ull) ? (ubk & 0x7fffffffffffffffull) : ~ubk;
    const double U = __longlong_as_double((long long)ubu);
#pragma unroll
    for (int e = 0; e < RR_MAXM / RR_THREADS; e++) {
      const bool keep = jj[e] >= 0 && !(lo[e] > U);
      const int p = warp_slot(&s_cnt, keep);
      if (keep) sel[p] = jj[e];
    }
    __syncthreads();
    m = s_cnt;
  }
  if (tid == 0 && cv.diag) { atomicAdd(cv.diag + 4, m); atomicAdd(cv.diag + 5, tot >> 4); }

  if (LEAF) {
    // exact distances, leaf-major (leaf_pass); candidates in batches whose leaf sums fit leafres
    const int cap = min(RR_MAXM, leafres_bytes / (8 * nleaves));
    for (int cb = 0; cb < m; cb += cap) {
      const int cn = min(cap, m - cb);
      for (int lf = 0; lf < nleaves; lf++) {
        const int off = leaf_s[4 * lf], steps = leaf_s[4 * lf + 1], tail = leaf_s[4 * lf + 2];
        switch (steps) {
          case 8: leaf_pass<8, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 7: leaf_pass<7, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 6: leaf_pass<6, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 5: leaf_pass<5, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 4: leaf_pass<4, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 3: leaf_pass<3, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 2: leaf_pass<2, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          case 1: leaf_pass<1, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
          default: leaf_pass<0, false>(x, sp, a_s, sel + cb, cn, off, tail, nleaves, lf, leafres, tid); break;
        }
      }
      __syncthreads();
      for (int ci = tid; ci < cn; ci += RR_THREADS) {
        const int j = sel[cb + ci];
        keys[cb + ci] = f64_key(combine_leaves(leafres + ci * nleaves, plan, plan_len));
        pos_s[cb + ci] = j < cs ? j : j - (ce - cs);  // position in the chromosome-excluded array
      }
      __syncthreads();
    }
  } else {
    // exact distances, one quad per candidate; results overwrite keys[0..m)
    const int quad = tid >> 2, l = tid & 3;
    for (int c0 = 0; c0 < m; c0 += RR_THREADS / 4) {
      const int ci = c0 + quad;
      const int cc = ci < m ? ci : (m - 1);  // keep the warp converged for the shuffles
      const int j = sel[cc];
      const double d = exact_sqdist_quad<VEC>(a_s, x + (int64_t)j * pv.s, plan, plan_len, l);
      if (l == 0 && ci < m) {
        keys[ci] = f64_key(d);
        pos_s[ci] = j < cs ? j : j - (ce - cs);  // position in the chromosome-excluded array
      }
    }
  }
  const uint64_t key_1e10 = f64_key(1e10);
  const int p2m = next_pow2(m < 2 ? 2 : m);
  for (int i = m + tid; i < p2m; i += RR_THREADS) { keys[i] = ~0ull; pos_s[i] = 0x7fffffff; }
  __syncthreads();
  if (LEAF) bitonic_sort_dpos_regs(keys, pos_s, p2m);
  else bitonic_sort_dpos(keys, pos_s, p2m);

  // a-posteriori completeness check: exact d_(k) + eps must stay below bound + |a|^2
  if (tid == 0 && tot > k) {
    const uint64_t kk = keys[k - 1];
    const uint64_t u = (kk & 0x8000000000000000ull) ? (kk & 0x7fffffffffffffffull) : ~kk;
    const double dk = __longlong_as_double((long long)u) * (pv.scale ? pv.scale[1] : 1.0);  // into the units of the list values
    if (!(dk - an + eps < bound)) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) fail_flags[lrow] = 4;
    for (int t = tid; t < k; t += RR_THREADS) oi[t] = 0;  // valid until exact_rows rewrites the row
    return;
  }
  for (int t = tid; t < k; t += RR_THREADS) {
    const bool have = t < m && keys[t] < key_1e10;
    if (have) {
      const uint64_t kk = keys[t];
      const uint64_t u = (kk & 0x8000000000000000ull) ? (kk & 0x7fffffffffffffffull) : ~kk;
      od[t] = __longlong_as_double((long long)u);
      oi[t] = pos_s[t];
    } else {
      od[t] = 1e10;
      oi[t] = -1;
    }
  }
}

int launch_rerank(const double* x, const PrepView& pv, CandView cv, int32_t nlists, const int64_t* cum_dev,
                  int32_t nchr, int64_t row_begin, int64_t row_end, int32_t k, int32_t gonosomal,
                  int32_t* idx_out, double* dist_out, int32_t* fail_flags, const int32_t* sum_plan,
                  int32_t plan_len, const double* xp, int32_t sp, const int32_t* leaf_dev, int32_t nleaves, cudaStream_t st) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0) return 0;
  if (nlists < 1 || nlists > 16) { set_error("rerank: bad number of candidate lists per row"); return 1; }
  const int maxc = nlists <= 2 ? 4096 : 8192;  // list entries below the common cut that fit in shared memory
  // the leaf-major gather needs the permuted copy (summation trees of up to 8 leaves, S <= 1024); otherwise the
  // row-major gather of quads
  const bool leaf = xp != nullptr && nleaves >= 1 && nleaves <= 8;
  const int row_len = leaf ? sp : pv.s;
  size_t smem = sizeof(double) * ((row_len + 1) & ~1) + (size_t)maxc * 8 + RR_MAXM * 8 + sizeof(int32_t) * ((3 * plan_len + 3) & ~3);
  if (leaf) smem += sizeof(int32_t) * 4 * nleaves;
  const bool vec = (pv.s % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  static size_t attr[3] = {0, 0, 0};
  const int which = leaf ? 2 : (vec ? 1 : 0);
  auto kern = which == 2 ? rerank_kernel<true, true> : which == 1 ? rerank_kernel<true, false> : rerank_kernel<false, false>;
  if (smem > attr[which]) {
    WCX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[which] = smem;
  }
  WCX_CUDA_OK(cudaMemsetAsync(fail_flags, 0, sizeof(int32_t) * rows, st));
  kern<<<(unsigned)rows, RR_THREADS, smem, st>>>(leaf ? xp : x, pv, cv, nlists, maxc, cum_dev, nchr, row_begin, k, gonosomal, idx_out,
                                               dist_out, fail_flags, sum_plan, plan_len, leaf ? sp : 0, leaf ? leaf_dev : nullptr,
                                               leaf ? nleaves : 0);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// brute-force exact rows (rows the fast path could not certify, WCX_KERNEL_EXACT): all N candidate
// distances in float64 into `scratch` (exact_dist_kernel, the candidate axis split over many CTAs
// per row so that a single failed row costs ~0.3 ms, not one CTA streaming all of X), then an
// exact (distance, position) top-k by 64-bit key bisection (exact_rows_kernel, one CTA per row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RR_THREADS)
exact_dist_kernel(const double* __restrict__ x, int64_t n, int s, const int64_t* __restrict__ cum, int nchr,
                  int64_t row_begin, const int32_t* __restrict__ rows_list, double* __restrict__ scratch,
                  const int32_t* __restrict__ plan_g, int plan_len, int cand_per_cta) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  double* a_s = reinterpret_cast<double*>(rr_smem);
  int32_t* plan = reinterpret_cast<int32_t*>(a_s + ((s + 1) & ~1));
  __shared__ int s_cs, s_ce;
  const int tid = threadIdx.x;
  const int64_t lrow = rows_list[blockIdx.x];
  const int64_t row = row_begin + lrow;
  uint64_t* dk = reinterpret_cast<uint64_t*>(scratch) + (int64_t)blockIdx.x * n;  // orderable keys of d
  if (tid == 0) {
    int c = 0;
    while (c < nchr && cum[c] <= row) c++;
    s_cs = (int)(c == 0 ? 0 : cum[c - 1]);
    s_ce = (int)cum[c];
  }
  for (int i = tid; i < s; i += RR_THREADS) a_s[i] = x[row * s + i];
  for (int i = tid; i < 3 * plan_len; i += RR_THREADS) plan[i] = plan_g[i];
  __syncthreads();
  const int cs = s_cs, ce = s_ce;
  const int quad = tid >> 2, l = tid & 3;
  const int64_t jb = (int64_t)blockIdx.y * cand_per_cta;
  int64_t je = jb + cand_per_cta;
  if (je > n) je = n;
  for (int64_t j0 = jb; j0 < je; j0 += RR_THREADS / 4) {
    int64_t j = j0 + quad;
    int64_t jj = j < n ? j : n - 1;
    double d = exact_sqdist_quad<false>(a_s, x + jj * s, plan, plan_len, l);
    if (l == 0 && j < je) {
      bool excluded = (j >= cs && j < ce) || !(d < 1e10);  // own chromosome, NaN, >= 1e10: never inserted
      dk[j] = excluded ? ~0ull : f64_key(d);
    }
  }
}

constexpr int EX_THREADS = 1024;  // selection kernel: one CTA per row, N keys per bisection pass

__global__ void __launch_bounds__(EX_THREADS)
exact_rows_kernel(int64_t n, const int64_t* __restrict__ cum, int nchr, int64_t row_begin,
                  const int32_t* __restrict__ rows_list, int k, int32_t* __restrict__ idx_out,
                  double* __restrict__ dist_out, const double* __restrict__ scratch) {
  __shared__ uint64_t keys[1024];
  __shared__ int32_t pos_s[1024];
  __shared__ int s_cs, s_ce, s_cnt;
  __shared__ unsigned long long s_count;
  __shared__ int s_scan[EX_THREADS];
  const int tid = threadIdx.x;
  const int64_t lrow = rows_list[blockIdx.x];
  const int64_t row = row_begin + lrow;
  const uint64_t* dk = reinterpret_cast<const uint64_t*>(scratch) + (int64_t)blockIdx.x * n;  // orderable keys of d

  if (tid == 0) {
    int c = 0;
    while (c < nchr && cum[c] <= row) c++;
    s_cs = (int)(c == 0 ? 0 : cum[c - 1]);
    s_ce = (int)cum[c];
  }
  __syncthreads();
  const int cs = s_cs, ce = s_ce;
  const uint64_t key_1e10 = f64_key(1e10);
  // number of valid candidates
  unsigned long long loc = 0;
#pragma unroll 8
  for (int64_t j = tid; j < n; j += EX_THREADS) loc += (dk[j] < key_1e10) ? 1 : 0;
  if (tid == 0) s_count = 0;
  __syncthreads();
  atomicAdd(&s_count, loc);
  __syncthreads();
  const long long valid = (long long)s_count;
  const int kk = valid < k ? (int)valid : k;  // entries to emit
  uint64_t T = 0;                             // kk-th smallest key
  if (kk > 0) {
    for (int bit = 63; bit >= 0; bit--) {
      uint64_t trial = T | (1ull << bit);
      loc = 0;
#pragma unroll 8
      for (int64_t j = tid; j < n; j += EX_THREADS) loc += (dk[j] < trial) ? 1 : 0;  // independent L2 loads, 8 in flight
      __syncthreads();
      if (tid == 0) s_count = 0;
      __syncthreads();
      atomicAdd(&s_count, loc);
      __syncthreads();
      if ((long long)s_count < kk) T = trial;
    }
  }
  // collect: strictly below T (any order), then ties in ascending position until kk entries
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (kk > 0) {
    for (int64_t j = tid; j < n; j += EX_THREADS) {
      if (dk[j] < T) {
        int p = atomicAdd(&s_cnt, 1);
        keys[p] = dk[j];
        pos_s[p] = (int)(j < cs ? j : j - (ce - cs));
      }
    }
    __syncthreads();
    int have = s_cnt;
    for (int64_t j0 = 0; j0 < n && have < kk; j0 += EX_THREADS) {
      int64_t j = j0 + tid;
      int f = (j < n && dk[j] == T) ? 1 : 0;
      s_scan[tid] = f;
      __syncthreads();
      // inclusive scan (Hillis-Steele)
      for (int o = 1; o < EX_THREADS; o <<= 1) {
        int v = tid >= o ? s_scan[tid - o] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
      }
      int p = have + s_scan[tid] - 1;
      if (f && p < kk) {
        keys[p] = T;
        pos_s[p] = (int)(j < cs ? j : j - (ce - cs));
      }
      have += s_scan[EX_THREADS - 1];
      __syncthreads();
    }
  }
  __syncthreads();
  const int p2 = next_pow2(kk < 2 ? 2 : kk);
  for (int i = kk + tid; i < p2; i += EX_THREADS) { keys[i] = ~0ull; pos_s[i] = 0x7fffffff; }
  __syncthreads();
  bitonic_sort_dpos(keys, pos_s, p2);
  int32_t* oi = idx_out + lrow * k;
  double* od = dist_out + lrow * k;
  for (int t = tid; t < k; t += EX_THREADS) {
    if (t < kk) {
      uint64_t q = keys[t];
      uint64_t u = (q & 0x8000000000000000ull) ? (q & 0x7fffffffffffffffull) : ~q;
      od[t] = __longlong_as_double((long long)u);
      oi[t] = pos_s[t];
    } else {
      od[t] = 1e10;
      oi[t] = -1;
    }
  }
}

int launch_exact_rows(const double* x, int64_t n, int32_t s, const int64_t* cum_dev, int32_t nchr,
                      int64_t row_begin, const int32_t* fail_rows, int32_t nfail, int32_t k, int32_t* idx_out,
                      double* dist_out, double* scratch, const int32_t* sum_plan, int32_t plan_len,
                      cudaStream_t st) {
  if (nfail <= 0) return 0;
  if (k > 1024) { set_error("exact_rows: k > 1024 unsupported"); return 1; }
  const size_t smem = sizeof(double) * ((s + 1) & ~1) + sizeof(int32_t) * 3 * plan_len;
  static size_t attr = 0;
  if (smem > attr) {
    WCX_CUDA_OK(cudaFuncSetAttribute(exact_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  // about four waves of CTAs over the whole batch, 64 candidates (one per quad) per pass
  int64_t chunks = (4 * 148 + nfail - 1) / nfail;
  if (chunks < 1) chunks = 1;
  int64_t per = (n + chunks - 1) / chunks;
  per = (per + 63) / 64 * 64;
  if (per < 64) per = 64;
  chunks = (n + per - 1) / per;
  if (chunks > 65535) { set_error("exact_rows: too many candidate chunks"); return 1; }
  exact_dist_kernel<<<dim3((unsigned)nfail, (unsigned)chunks), RR_THREADS, smem, st>>>(x, n, s, cum_dev, nchr, row_begin, fail_rows,
                                                                                  scratch, sum_plan, plan_len, (int)per);
  exact_rows_kernel<<<nfail, EX_THREADS, 0, st>>>(n, cum_dev, nchr, row_begin, fail_rows, k, idx_out, dist_out, scratch);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
