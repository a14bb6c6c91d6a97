// Exact re-rank, bulk-copy version (used when the sample count S is even, i.e. rows of X are
// 16-byte aligned; rerank.cu keeps the LDG version for odd S and documents the proof obligations
// of the fast path, which are identical here).
//
// Why: the exact float64 distance of ~390 nominated candidates per target bin is a gather of 4 KB
// rows of X (297 GB at config 3).  With per-thread LDG.128 one warp instruction touches 8 rows,
// i.e. 8 cache lines -> 10 L1 wavefronts per 512 bytes; the r01b profile shows the L1 data pipe
// at 86 % (39 M global + 31 M shared wavefronts per SM) while L2 runs at 35 % and the issue slots
// at 37 %.  Here the rows never pass through the LSU on their way in:
//
//   * every warp owns a private ring of stages in shared memory; a stage holds one leaf
//     (<= 128 terms = 1 KB) of NumPy's pairwise-summation tree for 8 candidates;
//   * lanes 0-7 of the warp issue one `cp.async.bulk` (global -> shared, completion on the stage's
//     mbarrier) per candidate, NST - 1 stages ahead of the arithmetic;
//   * the 8 quads of the warp then evaluate their candidate's leaf from shared memory in lock
//     step (LDS.128, rows padded to 1088 bytes: conflict free, 4 wavefronts per 512 bytes) with the
//     same 8 accumulator chains (2 per lane), no FMA contraction, and combine the leaves in
//     NumPy's order on a small register stack.
//
// No producer warp and no cross-warp synchronisation in the streaming phase: a first version
// with a shared ring and a producer warp let the quads of one warp wait on different slots, the
// warp diverged and ran at 91 ms against 50 ms for the LDG kernel.
// Reference semantics: newref_tools.py:260 (distance), :261-277 (selection); SURVEY.md A.1 / A.7.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wcx_common.cuh"

namespace wcx {

namespace {

constexpr int RT_MAXC = 4096;       // list entries below the cut kept for the k-th value search
constexpr int RT_MAXM = 1024;       // candidates evaluated exactly per row
constexpr int RT_MAXLEAVES = 128;
constexpr int RT_DEPTH = 8;         // register stack for the leaf combination
constexpr int RT_LEAF_STRIDE = 1088;  // 1 KB leaf + 64 B: the two quads of a quarter warp hit disjoint banks
constexpr int RT_STAGE_BYTES = 8 * RT_LEAF_STRIDE;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint64_t f64_key(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(uint64_t k) {
  uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ uint32_t f32_key_(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32_(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// same bound as rerank.cu
__device__ __forceinline__ double approx_eps(double an, double D, int k_pad, double abs_err) {
  const double u = 4.8852e-4;
  double na = sqrt(an);
  double sD = sqrt(D * 1.02 + 1e-300);
  double nb = na + sD;
  double e = u * (na + nb) + 2.0 * sqrt((double)k_pad) * abs_err;  // | |a^-b^| - |a-b| | <= |da| + |db|, |dx| <= u |x| + sqrt(K) abs_err
  double rounding = 2.0 * sD * e + e * e;
  double gamma = ((double)k_pad + 64.0) * 1.1920929e-7;
  double accum = 2.0 * gamma * na * nb;
  double misc = 4.77e-7 * (an + nb * nb + 2.0 * na * nb);
  return 1.5 * (rounding + accum + misc) + 1e-300;
}

// one leaf of the pairwise tree: a, b point at the first term of the leaf (both in shared memory, 16-byte
// aligned), evaluated by the 4 lanes of a quad; lane l owns NumPy's accumulators r[2l], r[2l+1].
__device__ __forceinline__ double leaf_sum_quad(const double* __restrict__ a, const double* __restrict__ b, int len, int l) {
  double res;
  if (len < 8) {
    res = 0.0;
    for (int i = 0; i < len; i++) {
      const double t = __dsub_rn(b[i], a[i]);
      res = __dadd_rn(res, __dmul_rn(t, t));
    }
    return res;
  }
  const int nblk = len >> 3;
  const double2* bp = reinterpret_cast<const double2*>(b) + l;
  const double2* ap = reinterpret_cast<const double2*>(a) + l;
  double2 bv = bp[0], av = ap[0];
  double t0 = __dsub_rn(bv.x, av.x), t1 = __dsub_rn(bv.y, av.y);
  double r0 = __dmul_rn(t0, t0), r1 = __dmul_rn(t1, t1);
#pragma unroll 5
  for (int blk = 1; blk < nblk; blk++) {
    bv = bp[4 * blk];
    av = ap[4 * blk];
    t0 = __dsub_rn(bv.x, av.x);
    t1 = __dsub_rn(bv.y, av.y);
    r0 = __dadd_rn(r0, __dmul_rn(t0, t0));
    r1 = __dadd_rn(r1, __dmul_rn(t1, t1));
  }
  const double s1 = __dadd_rn(r0, r1);
  const double s2 = __dadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, 1));
  res = __dadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, 2));
  for (int i = nblk << 3; i < len; i++) {
    const double t = __dsub_rn(b[i], a[i]);
    res = __dadd_rn(res, __dmul_rn(t, t));
  }
  return res;
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int NT>
__device__ __forceinline__ int block_sum(int c, int* s_red, int* s_total) {
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < NT / 32; w++) t += s_red[w];
    *s_total = t;
  }
  __syncthreads();
  return *s_total;
}

template <int NT>
__device__ void bitonic_sort_dpos(uint64_t* dkey, int32_t* pos, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = dkey[i], b = dkey[ixj];
          const int32_t pa = pos[i], pb = pos[ixj];
          const bool gt = (a > b) || (a == b && pa > pb);
          const bool up = ((i & k) == 0);
          if (gt == up) { dkey[i] = b; dkey[ixj] = a; pos[i] = pb; pos[ixj] = pa; }
        }
      }
      __syncthreads();
    }
  }
}

// shared memory: a[S] | ring[NW][nst][8][1088 B] | vals[RT_MAXC] u32 (later: dkey[RT_MAXM] u64, pos[RT_MAXM] i32) |
//                sel[RT_MAXM] i32 | leaves[3 L] i32 (offset, length, #adds after the leaf) | mbarriers[NW * nst]
template <int NW>
__global__ void __launch_bounds__(NW * 32)
rerank_bulk_kernel(const double* __restrict__ x, PrepView pv, CandView cv, int nlists, const int64_t* __restrict__ cum,
                   int nchr, int64_t row_begin, int k, int gonosomal, int32_t* __restrict__ idx_out,
                   double* __restrict__ dist_out, int32_t* __restrict__ fail_flags, const int32_t* __restrict__ leaves_g,
                   int nleaves, int nst) {
  constexpr int NT = NW * 32;
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const int S = pv.s;
  double* a_s = reinterpret_cast<double*>(rt_smem);
  unsigned char* ring = reinterpret_cast<unsigned char*>(a_s + S);
  uint32_t* vals = reinterpret_cast<uint32_t*>(ring + (size_t)NW * nst * RT_STAGE_BYTES);
  uint64_t* dkey = reinterpret_cast<uint64_t*>(vals);  // vals is dead once the k-th value is known
  int32_t* pos_s = reinterpret_cast<int32_t*>(dkey + RT_MAXM);
  int32_t* sel = reinterpret_cast<int32_t*>(vals + RT_MAXC);
  int32_t* leaves = sel + RT_MAXM;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(leaves + 3 * nleaves) + 7) & ~uintptr_t(7));
  __shared__ int s_tot, s_m, s_fail, s_cs, s_ce, s_total;
  __shared__ int s_red[NW];
  __shared__ float s_cut;

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
    s_m = 0;
    if (gonosomal && c != 22 && c != 23) s_cs = -1;
    float cut = __int_as_float(0x7f800000);
    for (int q = 0; q < nlists; q++) cut = fminf(cut, cv.cut[lrow * nlists + q]);
    s_cut = cut;
    for (int b = 0; b < NW * nst; b++) mbar_init(&bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (s_cs < 0) {  // placeholder rows of gonosomal references (newref_tools.py:186-191)
    for (int t = tid; t < k; t += NT) { oi[t] = 0; od[t] = 1.0; }
    return;
  }
  const int cs = s_cs, ce = s_ce;
  const float cut = s_cut;
  for (int i = tid; i < S; i += NT) a_s[i] = x[row * S + i];
  for (int i = tid; i < 3 * nleaves; i += NT) leaves[i] = leaves_g[i];
  // pass 1: values below the common cut (for the k-th smallest search)
  for (int q = 0; q < nlists; q++) {
    const int64_t slot = lrow * nlists + q;
    const int c = cv.cnt[slot];
    const uint2* le = cv.ent + slot * WCX_CAND_CAP;
    for (int i0 = tid; i0 < c; i0 += 4 * NT) {
      uint32_t v4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NT;
        v4[u] = i < c ? le[i].x : 0x7f800000u;  // +inf is never below the cut
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float v = __uint_as_float(v4[u]);
        if (v < cut) {
          const int p = atomicAdd(&s_tot, 1);
          if (p < RT_MAXC) vals[p] = f32_key_(v);
        }
      }
    }
  }
  __syncthreads();
  const int tot = s_tot;
  if (tot > RT_MAXC) {
    if (tid == 0) fail_flags[lrow] = 1;
    return;
  }
  double bound = 1e300, eps = 0.0;
  if (tot > k) {
    uint32_t res = 0;
    for (int bit = 31; bit >= 0; bit--) {
      const uint32_t trial = res | (1u << bit);
      int c = 0;
      for (int i = tid; i < tot; i += NT) c += (vals[i] < trial) ? 1 : 0;
      c = block_sum<NT>(c, s_red, &s_total);
      if (c < k) res = trial;
    }
    const float vk = key_f32_(res);
    const double an = (double)pv.norm[row];
    const double D = fmax((double)vk + an, 0.0);
    eps = approx_eps(an, D, pv.k_pad, (double)pv.abs_err);
    bound = (double)vk + 2.0 * eps;
  }
  if (!((double)cut > bound)) {  // the needed prefix must lie strictly below the cut
    if (tid == 0) fail_flags[lrow] = 1;
    return;
  }
  // pass 2: candidates with v <= bound
  for (int q = 0; q < nlists; q++) {
    const int64_t slot = lrow * nlists + q;
    const int c = cv.cnt[slot];
    const uint2* le = cv.ent + slot * WCX_CAND_CAP;
    for (int i0 = tid; i0 < c; i0 += 4 * NT) {
      uint2 e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NT;
        e[u] = i < c ? le[i] : make_uint2(0x7f800000u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float v = __uint_as_float(e[u].x);
        if (v < cut && (double)v <= bound) {
          const int p = atomicAdd(&s_m, 1);
          if (p < RT_MAXM) sel[p] = (int32_t)e[u].y;
        }
      }
    }
  }
  __syncthreads();
  const int m = s_m;
  if (m > RT_MAXM) {
    if (tid == 0) fail_flags[lrow] = 1;
    return;
  }
  if (tid == 0 && cv.diag) { atomicAdd(cv.diag + 4, m); atomicAdd(cv.diag + 5, tot >> 4); }

  // ---- exact distances: every warp streams groups of 8 candidates, leaf by leaf, through its own ring ----
  {
    const int warp = tid >> 5, lane = tid & 31, q = lane >> 2, l = lane & 3;
    const int ngroups = (m + 7) >> 3;
    const int my_groups = warp < ngroups ? (ngroups - warp + NW - 1) / NW : 0;
    const int ntasks = my_groups * nleaves;
    unsigned char* my_ring = ring + (size_t)warp * nst * RT_STAGE_BYTES;
    uint64_t* my_bar = bars + warp * nst;
    // issue cursor (task -> group, leaf, stage)
    int ig = warp, il = 0, ist = 0;
    auto issue = [&]() {
      const int off = leaves[3 * il], len = leaves[3 * il + 1];
      if (lane == 0) mbar_expect_tx(&my_bar[ist], 8u * (uint32_t)len * 8u);
      __syncwarp();
      if (lane < 8) {
        int ci = ig * 8 + lane;
        ci = ci < m ? ci : m - 1;  // short last group: re-read the last candidate, result unused
        bulk_g2s(my_ring + (size_t)ist * RT_STAGE_BYTES + lane * RT_LEAF_STRIDE, x + (int64_t)sel[ci] * S + off,
                 (uint32_t)len * 8u, &my_bar[ist]);
      }
      if (++il == nleaves) { il = 0; ig += NW; }
      if (++ist == nst) ist = 0;
    };
    int issued = 0;
    for (; issued < nst - 1 && issued < ntasks; issued++) issue();
    int g = warp, t = 0, st = 0;
    uint32_t parity = 0;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;  // combination stack, s0 = top
    for (int T = 0; T < ntasks; T++) {
      if (issued < ntasks) { issue(); issued++; }  // refills the stage consumed in the previous iteration
      mbar_wait(&my_bar[st], parity);
      const int off = leaves[3 * t], len = leaves[3 * t + 1], nadd = leaves[3 * t + 2];
      const double part = leaf_sum_quad(
          a_s + off, reinterpret_cast<const double*>(my_ring + (size_t)st * RT_STAGE_BYTES + q * RT_LEAF_STRIDE), len, l);
      s7 = s6; s6 = s5; s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = part;  // push
      for (int i = 0; i < nadd; i++) {  // pop right, pop left, push left + right
        s0 = __dadd_rn(s1, s0);
        s1 = s2; s2 = s3; s3 = s4; s4 = s5; s5 = s6; s6 = s7;
      }
      if (++t == nleaves) {
        const int ci = g * 8 + q;
        if (l == 0 && ci < m) {
          const int j = sel[ci];
          dkey[ci] = f64_key(s0);
          pos_s[ci] = j < cs ? j : j - (ce - cs);  // position in the chromosome-excluded array
        }
        t = 0;
        g += NW;
      }
      if (++st == nst) { st = 0; parity ^= 1u; }
      __syncwarp();  // all lanes are done with the stage before it is refilled
    }
  }
  const int p2m = next_pow2(m < 2 ? 2 : m);
  __syncthreads();
  for (int i = m + tid; i < p2m; i += NT) { dkey[i] = ~0ull; pos_s[i] = 0x7fffffff; }
  __syncthreads();
  bitonic_sort_dpos<NT>(dkey, pos_s, p2m);

  // a-posteriori completeness check: exact d_(k) + eps must stay below bound + |a|^2
  if (tid == 0 && tot > k) {
    const double dk = key_f64(dkey[k - 1]) * (pv.scale ? pv.scale[1] : 1.0);  // into the units of the list values
    const double an = (double)pv.norm[row];
    if (!(dk - an + eps < bound)) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) fail_flags[lrow] = 1;
    return;
  }
  const uint64_t key_1e10 = f64_key(1e10);
  for (int t = tid; t < k; t += NT) {
    const bool have = t < m && dkey[t] < key_1e10;
    od[t] = have ? key_f64(dkey[t]) : 1e10;
    oi[t] = have ? pos_s[t] : -1;
  }
}

}  // namespace

// leaves of the summation plan as (offset, length, number of ADD ops that follow the leaf); returns the
// number of leaves (or -1) and the maximum stack depth of the combination
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

// returns 1 if the bulk-copy variant cannot be used for this shape (caller falls back), -1 on error
int launch_rerank_bulk(const double* x, const PrepView& pv, CandView cv, int32_t nlists, const int64_t* cum_dev, int32_t nchr,
                       int64_t row_begin, int64_t row_end, int32_t k, int32_t gonosomal, int32_t* idx_out, double* dist_out,
                       int32_t* fail_flags, const int32_t* leaves_dev, int32_t nleaves, int32_t max_depth, cudaStream_t st) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0) return 0;
  const int S = pv.s;
  if ((S & 1) || (reinterpret_cast<uintptr_t>(x) & 15) || nleaves < 1 || nleaves > RT_MAXLEAVES || max_depth > RT_DEPTH ||
      k > RT_MAXM || nlists > 2)
    return 1;
  static const char* cfg = std::getenv("WCX_RERANK_BULK");  // "warps,stages" (tuning experiments)
  int nw = 4, nst = 2;
  if (cfg) {
    nw = std::atoi(cfg);
    const char* comma = std::strchr(cfg, ',');
    if (comma) nst = std::atoi(comma + 1);
  }
  if ((nw != 4 && nw != 8) || nst < 2 || nst > 8) { set_error("rerank_bulk: WCX_RERANK_BULK must be 4|8,2..8"); return -1; }
  const size_t smem = (size_t)S * 8 + (size_t)nw * nst * RT_STAGE_BYTES + RT_MAXC * 4 + RT_MAXM * 4 + (size_t)3 * nleaves * 4 + 8 +
                      (size_t)nw * nst * 8;
  if (smem > 227 * 1024) return 1;
  static size_t attr[2] = {0, 0};
  const int vi = nw == 8;
  if (smem > attr[vi]) {
    const cudaError_t e = nw == 8 ? cudaFuncSetAttribute(rerank_bulk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(rerank_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rerank_bulk: cannot set dynamic shared memory size"); return -1; }
    attr[vi] = smem;
  }
  if (cudaMemsetAsync(fail_flags, 0, sizeof(int32_t) * rows, st) != cudaSuccess) { set_error("rerank_bulk: memset failed"); return -1; }
  if (nw == 8)
    rerank_bulk_kernel<8><<<(unsigned)rows, 256, smem, st>>>(x, pv, cv, nlists, cum_dev, nchr, row_begin, k, gonosomal, idx_out, dist_out,
                                                           fail_flags, leaves_dev, nleaves, nst);
  else
    rerank_bulk_kernel<4><<<(unsigned)rows, 128, smem, st>>>(x, pv, cv, nlists, cum_dev, nchr, row_begin, k, gonosomal, idx_out, dist_out,
                                                           fail_flags, leaves_dev, nleaves, nst);
  if (cudaGetLastError() != cudaSuccess) { set_error("rerank_bulk: launch failed"); return -1; }
  return 0;
}

}  // namespace wcx
