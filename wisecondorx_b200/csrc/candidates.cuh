// Warp-cooperative maintenance of the per-row approximate-candidate lists shared by the
// CUDA-core and the tcgen05 distance kernels.
//
// A list holds (v, j) pairs with v = |b_j|^2 - 2 <a_i, b_j> (the row-constant |a_i|^2 is left
// out) for every candidate j whose v was below the row's threshold `thr` when it was seen.
// Invariant kept by every operation here (and relied upon by rerank.cu):
//
//     thr only decreases, and it is only ever lowered to a value t for which at least
//     WCX_CAND_KEEP list entries are known to be < t.
//
// Consequently every candidate that was rejected or dropped has v >= final thr ("cut"), and the
// list still contains every candidate below the cut.  Entries above the cut may linger in a list
// ("lazy deletion"); rerank.cu filters them.
#pragma once
#include "wcx_common.cuh"

namespace wcx {

__device__ __forceinline__ uint32_t f32_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Result of an exact compaction: new threshold, list minimum, exact counts below four probes.
struct CompactResult {
  float thr;
  float lo;
  int kept;
  float probe[4];  // ladder below thr: thr - (j + 1) * (thr - lo) / 8
  int below[4];    // exact number of surviving entries below each probe
};

__device__ __forceinline__ void make_probes(CompactResult& out) {
  const float d = (out.thr - out.lo) * 0.125f;
#pragma unroll
  for (int j = 0; j < 4; j++) { out.probe[j] = out.thr - (float)(j + 1) * d; out.below[j] = 0; }
}

// All 32 lanes call with the same arguments.  Requires KEEP < n <= 1024.
// Keeps the smallest entries: exactly KEEP of them, or -- early exit of the bisection -- any count
// in [KEEP, KEEP + KEEP/4] whose bound is a clean key prefix.  Rewrites the list in place.
template <int KEEP>
__device__ __forceinline__ CompactResult warp_compact(uint2* __restrict__ ent, int n) {
  constexpr int R = 32;
  constexpr int KEEP_HI = KEEP + KEEP / 4;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[R];
  uint2 e2[R];
  __syncwarp();
  // batched loads (32 x 64-bit requests in flight per lane)
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int e = r * 32 + lane;
    e2[r] = __ldcg(ent + (e < n ? e : 0));
  }
  uint32_t kmin = 0xffffffffu;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int e = r * 32 + lane;
    key[r] = (e < n) ? f32_key(__uint_as_float(e2[r].x)) : 0xffffffffu;
    kmin = key[r] < kmin ? key[r] : kmin;
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  uint32_t res = 0;
  uint32_t bound = 0;
  bool early = false;
#pragma unroll 1
  for (int bit = 31; bit >= 0; bit--) {
    const uint32_t trial = res | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += (key[r] < trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < KEEP) {
      res = trial;
    } else if (c <= KEEP_HI) {
      bound = trial;
      early = true;
      break;
    }
  }
  CompactResult out;
  out.lo = key_f32(kmin);
  int base = 0;
  const uint32_t lim = early ? bound : res;
#pragma unroll
  for (int r = 0; r < R; r++) {  // everything strictly below the limit
    const bool f = key[r] < lim;
    const uint32_t b = __ballot_sync(0xffffffffu, f);
    if (f) ent[base + __popc(b & lt_mask)] = e2[r];
    base += __popc(b);
  }
  if (early) {
    out.thr = key_f32(bound);
    out.kept = base;
  } else {
    // res == KEEP-th smallest key: ties fill up to KEEP
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool f = key[r] == res;
      const uint32_t b = __ballot_sync(0xffffffffu, f);
      if (f) {
        const int p = base + __popc(b & lt_mask);
        if (p < KEEP) ent[p] = e2[r];
      }
      base += __popc(b);
    }
    out.thr = key_f32(res);
    out.kept = KEEP;
  }
  // exact counts of the survivors below the ladder probes (the keys are still in registers)
  make_probes(out);
  {
    uint32_t pk[4];
    int c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) pk[j] = f32_key(out.probe[j]);
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int j = 0; j < 4; j++) c[j] += (key[r] < pk[j]) ? 1 : 0;  // probes lie below thr: all such entries survive
    }
#pragma unroll
    for (int j = 0; j < 4; j++) out.below[j] = __reduce_add_sync(0xffffffffu, c[j]);
  }
  __syncwarp();
  return out;
}

// Generic (slow, rare) compaction for lists of any length n <= WCX_CAND_CAP: bisection with the
// counts streamed from memory, then an in-place streaming filter that keeps the entries strictly
// below the KEEP-th smallest value `res` and sets thr = res.  Ties at `res` are dropped, which is
// safe: dropped entries have v >= thr, and if the boundary matters the re-rank's `cut > bound`
// check sends the row to the exact brute-force path.
template <int KEEP>
static __device__ __noinline__ CompactResult warp_compact_stream(uint2* __restrict__ ent, int n) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  __syncwarp();
  uint32_t res = 0, kmin = 0xffffffffu;
  for (int bit = 31; bit >= 0; bit--) {
    const uint32_t trial = res | (1u << bit);
    int c = 0;
    for (int e = lane; e < n; e += 32) {
      const uint32_t k = f32_key(__uint_as_float(__ldcg(ent + e).x));
      c += (k < trial) ? 1 : 0;
      kmin = k < kmin ? k : kmin;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < KEEP) res = trial;
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  int base = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {  // in-place left compaction: writes trail the reads
    const int e = e0 + lane;
    uint2 v = make_uint2(0u, 0u);
    bool f = false;
    if (e < n) { v = __ldcg(ent + e); f = f32_key(__uint_as_float(v.x)) < res; }
    const uint32_t b = __ballot_sync(0xffffffffu, f);
    if (f) ent[base + __popc(b & lt_mask)] = v;
    base += __popc(b);
    __syncwarp();
  }
  CompactResult out;
  out.thr = key_f32(res);
  out.lo = key_f32(kmin);
  out.kept = base;
  make_probes(out);  // counts start at 0: valid lower bounds
  return out;
}

}  // namespace wcx
