// Order statistics of up to 2 NP small keys held by ONE thread, packed two per 32-bit register.
//
// The keys are 15-bit codes stored as the bit patterns of positive normal fp16 numbers (0x0400 .. 0x7BFE; 0x7BFF is
// the padding key, above every code), so an fp16 compare orders them and one HSET2 + one HFMA2 handle two keys.  Used
// by the thread-per-median kernels: null ratios (null_ratios.cu, codes = equal-frequency map of a sample column) and
// the within-sample normalisation of predict (predict.cu, codes = linear map of the reference values of one bin).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace wcx {

constexpr uint32_t PS_CODE_LO = 0x0400u;   // smallest code: smallest positive normal fp16
constexpr uint32_t PS_CODE_HI = 0x7BFEu;   // largest code
constexpr uint32_t PS_CODE_PAD = 0x7BFFu;  // largest finite fp16: padding key

// ---- packed fp16 helpers: the codes are bit patterns of positive normal halves, so fp16 compares order them ----
__device__ __forceinline__ uint32_t h2_lt(uint32_t a, uint32_t b) {  // 1.0h per half where a < b
  uint32_t d;
  asm("set.lt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ int h2_total(uint32_t acc) {  // sum of the two (small, integral) halves
  return (int)(__half2float(__ushort_as_half((unsigned short)(acc & 0xffffu))) +
               __half2float(__ushort_as_half((unsigned short)(acc >> 16))));
}
template <int NP>
__device__ __forceinline__ int ps_count_lt(const uint32_t (&k2)[NP], uint32_t trial) {
  const uint32_t t2 = trial | (trial << 16);
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // four chains; each half counts at most NP / 4 + 1: exact in fp16
#pragma unroll
  for (int j = 0; j < NP; j += 4) {
    a0 = h2_add(a0, h2_lt(k2[j], t2));
    if (j + 1 < NP) a1 = h2_add(a1, h2_lt(k2[j + 1], t2));
    if (j + 2 < NP) a2 = h2_add(a2, h2_lt(k2[j + 2], t2));
    if (j + 3 < NP) a3 = h2_add(a3, h2_lt(k2[j + 3], t2));
  }
  return h2_total(a0) + h2_total(a1) + h2_total(a2) + h2_total(a3);
}
__device__ __forceinline__ uint32_t h2_eq(uint32_t a, uint32_t b) {  // 1.0h per half where a == b
  uint32_t d;
  asm("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// bit pattern of the fp16 pair (float(lo), float(hi)) for small non-negative integers (exact below 2048)
__host__ __device__ constexpr uint32_t h_bits(int v) {
  if (v == 0) return 0u;
  int e = 0;
  while ((v >> e) > 1) e++;
  return (uint32_t)(((e + 15) << 10) | (((v << (10 - e)) & 0x3ff)));
}
__host__ __device__ constexpr uint32_t h2_const(int lo, int hi) { return h_bits(lo) | (h_bits(hi) << 16); }

// largest key below `lim` as a code (0 if none): key * [key < lim] keeps the key or +0, fp16 max orders the patterns
template <int NP>
__device__ __forceinline__ uint32_t ps_max_below(const uint32_t (&k2)[NP], uint32_t lim) {
  const uint32_t l2 = lim | (lim << 16);
  uint32_t b0 = 0, b1 = 0;
#pragma unroll
  for (int j = 0; j < NP; j += 2) {
    b0 = h2_max(b0, h2_mul(k2[j], h2_lt(k2[j], l2)));
    if (j + 1 < NP) b1 = h2_max(b1, h2_mul(k2[j + 1], h2_lt(k2[j + 1], l2)));
  }
  const uint32_t b = h2_max(b0, b1);
  const uint32_t lo = b & 0xffffu, hi = b >> 16;
  return lo > hi ? lo : hi;
}
// number of keys equal to `code` and the sum of (position + 1) over them -- the position itself when there is one
template <int NP>
__device__ __forceinline__ void ps_find(const uint32_t (&k2)[NP], uint32_t code, int& count, int& pos) {
  const uint32_t c2 = code | (code << 16);
  uint32_t n0 = 0, n1 = 0, p0 = 0, p1 = 0;
#pragma unroll
  for (int j = 0; j < NP; j += 2) {
    const uint32_t e0 = h2_eq(k2[j], c2);
    n0 = h2_add(n0, e0);
    p0 = h2_fma(e0, h2_const(2 * j + 1, 2 * j + 2), p0);
    if (j + 1 < NP) {
      const uint32_t e1 = h2_eq(k2[j + 1], c2);
      n1 = h2_add(n1, e1);
      p1 = h2_fma(e1, h2_const(2 * j + 3, 2 * j + 4), p1);
    }
  }
  count = h2_total(n0) + h2_total(n1);
  pos = h2_total(p0) + h2_total(p1) - 1;  // meaningful only when count == 1 (sums of several positions may round)
}


// Code of the key of rank t (0-based) among the keys of k2: the largest T with count(key < T) <= t.  15 steps.
// below_out = count(key < T).
template <int NP>
__device__ __forceinline__ uint32_t ps_select(const uint32_t (&k2)[NP], int t, int& below_out) {
  uint32_t T = 0;
  int below = 0;
#pragma unroll 1
  for (int bit = 14; bit >= 0; bit--) {
    const uint32_t trial = T | (1u << bit);
    const int c = ps_count_lt<NP>(k2, trial);
    if (c <= t) { T = trial; below = c; }
  }
  below_out = below;
  return T;
}

}  // namespace wcx
