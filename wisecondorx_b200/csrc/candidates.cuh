// Warp-cooperative maintenance of the per-row approximate-candidate lists shared by the
// CUDA-core and the tcgen05 distance kernels.
//
// A list holds (v, j) pairs with v = |b_j|^2 - 2 <a_i, b_j> (the row-constant |a_i|^2 is left
// out) for every candidate j seen so far whose v was below the row's running threshold.  When a
// list is about to overflow, `warp_compact` keeps the WCX_CAND_KEEP smallest v and lowers the
// threshold to the KEEP-th smallest; everything ever dropped or rejected therefore has
// v >= final threshold ("cut"), which is what the exact re-rank needs for its proof of
// completeness (rerank.cu).
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

// All 32 lanes call with the same arguments.  Requires KEEP < n <= CAP.
// Rewrites val/idx in place with exactly KEEP entries and returns the KEEP-th smallest value.
__device__ __forceinline__ float warp_compact(float* __restrict__ val, int32_t* __restrict__ idx, int n) {
  constexpr int R = WCX_CAND_CAP / 32;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[R];
  int32_t id[R];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R; r++) {
    int e = r * 32 + lane;
    bool in = e < n;
    key[r] = in ? f32_key(val[e]) : 0xffffffffu;
    id[r] = in ? idx[e] : -1;
  }
  uint32_t res = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; bit--) {
    uint32_t trial = res | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += (key[r] < trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < WCX_CAND_KEEP) res = trial;
  }
  // res == KEEP-th smallest key.  First everything strictly below, then ties up to KEEP.
  int base = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    bool f = key[r] < res;
    uint32_t b = __ballot_sync(0xffffffffu, f);
    if (f) {
      int p = base + __popc(b & lt_mask);
      val[p] = key_f32(key[r]);
      idx[p] = id[r];
    }
    base += __popc(b);
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    bool f = key[r] == res;
    uint32_t b = __ballot_sync(0xffffffffu, f);
    if (f) {
      int p = base + __popc(b & lt_mask);
      if (p < WCX_CAND_KEEP) {
        val[p] = key_f32(key[r]);
        idx[p] = id[r];
      }
    }
    base += __popc(b);
  }
  __syncwarp();
  return key_f32(res);
}

}  // namespace wcx
