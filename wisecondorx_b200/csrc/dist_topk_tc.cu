// The distance sweep: fused bins x bins contraction + approximate top-k nomination on tcgen05 / TMEM / TMA (sm_100a).
//
// Reference loop replaced: get_ref_for_bins, newref_tools.py:255-278.  The bins x bins squared distance is the one
// dense contraction of WisecondorX:
//     d(i,j) = |a_i|^2 + |b_j|^2 - 2 <a_i, b_j>
// <a_i, b_j> runs on the 5th-generation tensor cores as a kind::f16 UMMA with fp32 accumulation in TMEM.  Operands are
// the centred matrix scaled by a power of two and rounded to f16 (newref_prep.cu: same 11-bit significand as tf32 at
// twice the tensor rate and half the bytes), K-major, staged by TMA into 128-byte-swizzled shared-memory tiles.  The
// sweep only NOMINATES candidates: rerank.cu recomputes the nominated distances exactly (float64, NumPy's summation
// order) and proves per row that no candidate was missed, so the f16 rounding never reaches the result.
//
// Shipping configuration (PAIR = true): clusters of two CTAs (one per SM) drive one cta_group::2 MMA of M = 256 target
// rows x N = 256 candidate columns per tile; each CTA stages its own 128 target rows and HALF of the candidate
// columns (the pair halves the L2 -> SM operand traffic of B).  The epilogue never writes the distance tile: each
// epilogue thread owns one target row (= one TMEM lane), pulls 32 accumulator columns at a time with tcgen05.ld, forms
// v = |b_j|^2 - 2 acc and appends (v, j) to the row's candidate list only if v is below the row's running threshold
// (1-2 % of the elements).
//
// Threshold maintenance (candidates.cuh states the invariant).  After the first 512 entries an exact warp-cooperative
// selection fixes thr and a ladder of four probe values below it; every append counts itself against the probes, and
// when the first probe has seen WCX_CAND_KEEP_TC entries below it thr drops to that probe -- no list traffic at all.
// Entries that end up above thr stay in the list (lazy deletion, filtered by rerank.cu); a physical compaction happens
// only if a list is about to overflow its 4096 slots (rare).
//
// Warp roles (384 threads, one CTA per SM, persistent over work items):
//   warp 0        TMA producer (one elected lane)
//   warp 1        MMA issuer   (one elected lane of the pair's leader CTA: tcgen05.mma cta_group::2 kind::f16, K = 16)
//   warp 2        TMEM allocator (512 columns = two 256-column fp32 accumulators)
//   warps 4..11   epilogue: BOTH warp groups drain every accumulator (group g takes the 32-column chunks 2 i + g), so a
//                 TMEM buffer is handed back after half a drain while the MMA warp fills the other one; each group keeps
//                 its own list per row and the two groups exchange thresholds through shared memory.
// Pipelines: smem full / empty mbarriers (TMA <-> MMA, 5 stages of 32 KB per CTA) and TMEM full / empty mbarriers
// (MMA <-> epilogue).  All CTA pairs walk the candidate axis in the same order and within one chromosome of each other
// (tile_own below), so every B tile is fetched from HBM once and served from L2 to the other pairs.
// The one-CTA-per-SM variant (PAIR = false) is kept as a cross-check path and for the single-tile test hook; the tf32
// operand variants of round 1 are gone.
#include <cuda.h>

#include <cstdlib>

#include "candidates.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {

constexpr int TM = WCX_TILE_M;          // 128
constexpr int TN = WCX_TILE_N_TC;       // 256
constexpr int BK = WCX_KBLOCK;          // 32 tf32 = 128 bytes per row and pipeline stage (f16: 64 elements, also 128 bytes)
constexpr int A_BYTES = TM * BK * 4;    // 16 KB
constexpr int B_BYTES = TN * BK * 4;    // 32 KB (1-CTA mode); a CTA of a pair stages half of it
constexpr int TC_THREADS = 384;
constexpr int NORM_BYTES = 8 * (TN / 2) * 4;  // per-epilogue-warp copy of the candidate norms of its half tile
constexpr int THR_BYTES = 2 * TM * 8;   // (item, thr) words exchanged between the two epilogue groups
constexpr int SPILL_BYTES = 256 * 32 * 4;  // 32 floats per epilogue thread (one chunk), [8][256] float4, see filter_chunk

template <bool PAIR> struct Cfg {
  static constexpr int STAGES = PAIR ? 5 : 3;  // 32 KB / 48 KB stages (K = 32 tf32 or 64 f16 elements each)
  static constexpr int B_STAGE = PAIR ? B_BYTES / 2 : B_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + NORM_BYTES + THR_BYTES + SPILL_BYTES;
};
constexpr int INIT_N = 512;             // entries collected before the first exact selection
constexpr int KEEP = WCX_CAND_KEEP_TC;  // per-list guarantee; the row's two lists together hold >= 2 * KEEP
constexpr uint32_t TMEM_COLS = 512;

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, K-major both,
// N = 256 (n_dim = N >> 3 at bit 17), M = 128 (m_dim = M >> 4 at bit 24)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// cta_group::2: the instruction spans both CTAs of the pair, M = 256 (128 rows from each CTA), N = 256
// (128 candidate columns staged by each CTA)
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
// kind::f16 with F16 operands (a_format = b_format = 0): same shapes, K = 16 per instruction
constexpr uint32_t IDESC_H = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC2_H = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// pair mode: the load lands in the issuing CTA's shared memory, the bytes are credited to CTA 0's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  if (F16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC2_H), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC2), "r"(accumulate)
        : "memory");
  }
}
// arrive on the barrier at the same offset in both CTAs of the pair once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on CTA 0's copy of a barrier from either CTA of the pair.  Relaxed: the barrier only hands a TMEM buffer back
// to the MMA warp, no generic-memory data travels with it, and the TMEM reads are complete (tcgen05.wait::ld) and fenced
// (tcgen05.fence::before_thread_sync) before the arrive.  The .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR,
// 13 % of the epilogue's samples in the r01d profile.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B
// between 8-row groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool F16>
__device__ __forceinline__ void umma_single(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  if (F16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC_H), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// A tile that lies entirely inside the target rows' own chromosome yields no candidates.  When that holds for all 256
// rows of a unit the tile is skipped by all three roles (5.7 % of the tiles on a whole genome; 29.98 vs 30.79 ms in
// the same run): the CTA pairs then run at most one chromosome (<= 16 MB of operands) apart on the candidate axis, so
// every B tile is still fetched from HBM once and served from L2 to the other pairs.  (The round-1 kernel was
// HBM-latency bound when it skipped tiles -- profiles/r01b, 209 GB of DRAM reads -- with a three-stage pipeline and
// no operand reuse inside a pair.)  A unit whose two halves belong to different chromosomes multiplies the tile and
// only the epilogue of the half that owns it skips the filter.
__device__ __forceinline__ bool tile_own(const WorkItem& w, int ct) {
  const int col0 = ct * TN;
  return col0 >= w.chr_s && col0 + TN <= w.chr_e;
}

}  // namespace

// per-row threshold state of one epilogue thread
struct RowState {
  float thr, lo;
  float p0, p1, p2, p3;  // probe ladder, p0 > p1 > p2 > p3, all below thr
  int c0, c1, c2, c3;    // lower bounds on the number of list entries below each probe
  int cnt;
  bool ladder;
};

// lower thr to the first probe while that probe has enough entries below it
__device__ __forceinline__ void ladder_advance(RowState& st) {
  while (st.ladder && st.c0 >= KEEP) {
    st.thr = st.p0;
    st.p0 = st.p1; st.p1 = st.p2; st.p2 = st.p3;
    st.c0 = st.c1; st.c1 = st.c2; st.c2 = st.c3;
    const float d = (st.thr - st.lo) * 0.125f;
    st.p3 = st.p2 - d;
    st.c3 = 0;
    if (!(d > 0.f)) break;
  }
}
// adopt a lower threshold published by the partner group (valid: it has >= KEEP entries below it)
__device__ __forceinline__ void ladder_adopt(RowState& st, float t) {
  if (t < st.thr) {
    st.thr = t;
    while (st.ladder && st.p0 >= st.thr) {  // probes at or above thr are useless: shift them out
      st.p0 = st.p1; st.p1 = st.p2; st.p2 = st.p3;
      st.c0 = st.c1; st.c1 = st.c2; st.c2 = st.c3;
      const float d = (st.thr - st.lo) * 0.125f;
      st.p3 = st.p2 - d;
      st.c3 = 0;
      if (!(d > 0.f)) break;
    }
  }
}

// explicit shared-window accesses (the compiler otherwise falls back to generic LD/ST for the carved-up dynamic buffer)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ unsigned long long lds_u64_volatile(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u64_volatile(uint32_t a, unsigned long long v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
// bits [0, k) set (k clamped to 0..32)
__device__ __forceinline__ uint32_t bits_below(int k) { return k <= 0 ? 0u : (k >= 32 ? 0xffffffffu : ((1u << k) - 1u)); }

// 32 accumulator columns of one row against the staged norms.
//
// Appends are rare per lane (about 2 % of the elements) but almost every group of columns has SOME
// lane of the warp that passes, so a per-element `if (v < thr) append()` makes the whole warp walk
// the append code for nearly every element.  Instead the hot loop is branch-free: it turns the
// accumulators into v in place and builds a 32-bit pass mask (compare + predicated OR, four
// partial masks for ILP); lanes with a non-zero mask park their 32 values in a private
// shared-memory strip and drain the set bits in a compact loop in which every lane works on its
// own entries at the same time.  The loop runs max-over-lanes(popcount) times, so one drain per 32
// columns (not two per 16) and a short body matter: the r01d profile had 31 instructions per
// iteration and 1.6 iterations per 16 columns, as much issue time as the arithmetic itself.
template <bool ALL_VALID>
__device__ __forceinline__ void filter_chunk(uint32_t (&r)[32], uint32_t snorm_a, int g0, const WorkItem& w, int64_t n,
                                             RowState& st, uint2* be, uint32_t spill_a) {
  uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
  for (int j4 = 0; j4 < 8; j4++) {
    const float4 nb = lds_f4(snorm_a + 16 * j4);
    r[4 * j4 + 0] = __float_as_uint(fmaf(-2.f, __uint_as_float(r[4 * j4 + 0]), nb.x));
    r[4 * j4 + 1] = __float_as_uint(fmaf(-2.f, __uint_as_float(r[4 * j4 + 1]), nb.y));
    r[4 * j4 + 2] = __float_as_uint(fmaf(-2.f, __uint_as_float(r[4 * j4 + 2]), nb.z));
    r[4 * j4 + 3] = __float_as_uint(fmaf(-2.f, __uint_as_float(r[4 * j4 + 3]), nb.w));
  }
#define WCX_PASS(M, J) \
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(M) : "f"(__uint_as_float(r[J])), "f"(st.thr), "r"(1u << (J)))
#pragma unroll
  for (int j4 = 0; j4 < 8; j4++) {
    WCX_PASS(m0, 4 * j4 + 0);
    WCX_PASS(m1, 4 * j4 + 1);
    WCX_PASS(m2, 4 * j4 + 2);
    WCX_PASS(m3, 4 * j4 + 3);
  }
#undef WCX_PASS
  uint32_t mask = (m0 | m1) | (m2 | m3);
  if (!ALL_VALID)  // columns past the end of the matrix or inside the rows' own chromosome
    mask &= bits_below((int)(n - g0)) & ~(bits_below(w.chr_e - g0) & ~bits_below(w.chr_s - g0));
  if (mask) {
#pragma unroll
    for (int q = 0; q < 8; q++)
      sts_f4(spill_a + q * 4096, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
             __uint_as_float(r[4 * q + 3]));
    uint2* wp = be + st.cnt;
    do {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const float v = lds_f(spill_a + (j >> 2) * 4096 + (j & 3) * 4);
      *wp++ = make_uint2(__float_as_uint(v), (uint32_t)(g0 + j));
      st.lo = fminf(st.lo, v);
#define WCX_COUNT(C, P) asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(C) : "f"(v), "f"(P))
      WCX_COUNT(st.c0, st.p0);
      WCX_COUNT(st.c1, st.p1);
      WCX_COUNT(st.c2, st.p2);
      WCX_COUNT(st.c3, st.p3);
#undef WCX_COUNT
    } while (mask);
    st.cnt = (int)(wp - be);
  }
}

// PAIR = false: one CTA per SM, cta_group::1, item i -> CTA (i mod grid).
// PAIR = true : clusters of two CTAs (one SM each) cooperate through cta_group::2: items 2p and 2p + 1
//               (same candidate-column range) form a 256-row tile; each CTA stages its own 128 target
//               rows and HALF of the 256 candidate columns, so the L2 -> SM operand traffic per tile
//               drops from 768 KB to 512 KB per SM; the leader CTA issues the MMAs for both.
// F16: operands are the scaled f16 matrix (newref_prep.cu) and the MMA is kind::f16 -- 64 elements per 128-byte
//       stage row, four K = 16 instructions per stage, i.e. half as many stages and MMAs per tile as TF32.
template <bool PAIR, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
dist_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, PrepView pv, const WorkItem* __restrict__ items,
                    int nitems, CandView cv, float* __restrict__ dbg_acc, int skip_own) {
  constexpr int STAGES = Cfg<PAIR>::STAGES;
  constexpr int STAGE_BYTES = Cfg<PAIR>::STAGE_BYTES;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first work unit (item or item pair)
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int nunits = PAIR ? (nitems >> 1) : nitems;
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  // (pointer arithmetic on the array itself: an integer round trip would demote every later access to generic LD/ST)
  unsigned char* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;       // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* s_norm = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [8 warps][TN / 2]
  // s_thr words are (item << 32 | float bits of thr): a threshold is only adopted from the same work item
  // [2 groups][TM] u64, then the spill strips [8][256] float4 (shared-window addresses)
  const uint32_t s_thr_a = smem_u32(smem + STAGES * STAGE_BYTES + 256 + NORM_BYTES);
  const uint32_t s_spill_a = s_thr_a + THR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int BKE = F16 ? 2 * BK : BK;  // elements per stage row
  const int kblocks = pv.k_pad / BKE;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], PAIR ? 16 : 8); }  // every epilogue warp (of both CTAs in pair mode) releases every buffer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < nunits; unit += unit_step) {
        const WorkItem w = items[PAIR ? 2 * unit + (int)crank : unit];
        const WorkItem wo = items[PAIR ? 2 * unit + (int)(crank ^ 1u) : unit];  // the other half of the 256-row tile
        for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
          if (skip_own && tile_own(w, ct) && tile_own(wo, ct)) continue;  // all three roles skip the same tiles
          const int col0 = ct * TN;
          for (int kb = 0; kb < kblocks; kb++) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* sa = smem + stage * STAGE_BYTES;
            unsigned char* sb = sa + A_BYTES;
            if (PAIR) {
              // CTA 0's barrier collects the bytes of both CTAs (2 x 32 KB per stage)
              if (crank == 0) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
              tma_load_2d_pair(sa, &tmap, &full[stage], kb * BKE, w.row0);
              tma_load_2d_pair(sb, &tmap, &full[stage], kb * BKE, col0 + (int)crank * (TN / 2));
            } else {
              mbar_expect_tx(&full[stage], STAGE_BYTES);
              tma_load_2d(sa, &tmap, &full[stage], kb * BKE, w.row0);
              tma_load_2d(sb, &tmap, &full[stage], kb * BKE, col0);
              tma_load_2d(sb + B_BYTES / 2, &tmap, &full[stage], kb * BKE, col0 + TN / 2);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair mode: leader CTA only) =====================
    if (lane == 0 && crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase = 0;
      for (int unit = unit0; unit < nunits; unit += unit_step) {
        const WorkItem w = items[PAIR ? 2 * unit : unit];
        const WorkItem wo = items[PAIR ? 2 * unit + 1 : unit];
        for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
          if (skip_own && tile_own(w, ct) && tile_own(wo, ct)) continue;
          mbar_wait(&tempty[buf], tphase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TN);
          for (int kb = 0; kb < kblocks; kb++) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t adesc = make_desc(sa);
            const uint64_t bdesc = make_desc(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; k++) {
              // advance 8 tf32 / 16 f16 = 32 bytes along K inside the 128B swizzle row: +2 in the >>4 address field
              if (PAIR) umma_pair<F16>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (kb | k) != 0 ? 1u : 0u);
              else umma_single<F16>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (kb | k) != 0 ? 1u : 0u);
            }
            // smem slot reusable once these MMAs have read it (pair mode: in both CTAs)
            if (PAIR) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          // accumulator complete (pair mode: each CTA's epilogue waits on its own copy)
          if (PAIR) umma_commit_pair(&tfull[buf]); else umma_commit(&tfull[buf]);
          if (++buf == 2) { buf = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread owns one target row and one half of the columns of every tile =====================
    // Both groups drain every accumulator (group g: columns 128 g .. 128 g + 127) while the MMA warp fills the other
    // TMEM buffer.  (Tile-alternating groups left each group idle for a whole MMA pass per tile: its buffer could only
    // be refilled after it had drained it -- 30 % of the epilogue's samples sat in that wait in the r01d profile.)
    const int grp = (warp - 4) >> 2;  // 0 / 1 -> column half
    const int q = warp & 3;
    const int row = q * 32 + lane;
    constexpr int TH = TN / 2;
    float* snorm = s_norm + (warp - 4) * TH;
    const uint32_t snorm_a = smem_u32(snorm);
    const uint32_t spill_a = s_spill_a + (threadIdx.x - 128) * 16;
    const uint32_t my_thr_a = s_thr_a + (uint32_t)(grp * TM + row) * 8;
    const uint32_t peer_thr_a = s_thr_a + (uint32_t)((grp ^ 1) * TM + row) * 8;
    uint32_t tphase = 0;
    int buf = 0;
    bool dbg_done = false;
    for (int unit = unit0; unit < nunits; unit += unit_step) {
      const int item = PAIR ? 2 * unit + (int)crank : unit;
      const WorkItem w = items[item];
      const WorkItem wo = items[PAIR ? (item ^ 1) : item];
      const bool row_ok = row < w.nrows;
      const int64_t slot = (int64_t)w.slot0 + (int64_t)row * w.slot_stride + grp;
      uint2* be = cv.ent + (row_ok ? slot : 0) * WCX_CAND_CAP;
      RowState st;
      st.thr = row_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);  // +inf / -inf
      st.lo = __int_as_float(0x7f800000);
      st.p0 = st.p1 = st.p2 = st.p3 = __int_as_float(0xff800000);
      st.c0 = st.c1 = st.c2 = st.c3 = 0;
      st.cnt = 0;
      st.ladder = false;
      sts_u64_volatile(my_thr_a, ((unsigned long long)(uint32_t)item << 32) | __float_as_uint(st.thr));
      for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
        if (skip_own && tile_own(w, ct) && tile_own(wo, ct)) continue;  // never multiplied
        const int my_buf = buf;
        const uint32_t my_phase = tphase;
        if (++buf == 2) { buf = 0; tphase ^= 1; }
        // this group's share of the tile: the 32-column chunks 2 i + grp (i = 0..3).  Interleaving at chunk granularity
        // (rather than halves of 128 columns) keeps the two lists of a row balanced when its nearest candidates come in
        // runs of adjacent bins -- a run shorter than a half tile used to land in one list only, whose threshold then
        // certified fewer than ref_size candidates (row 149352 of config 3 fell back to the exact path for that reason).
        const int col0 = ct * TN;
        if (tile_own(w, ct)) {  // nothing to filter: just hand the accumulator back
          mbar_wait(&tfull[my_buf], my_phase);
          tc_fence_after();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[my_buf]); else mbar_arrive(&tempty[my_buf]); }
          continue;
        }
        // stage the candidate norms of the four chunks (issued before waiting for the accumulator): lane l holds
        // floats 4 (l & 7) .. +3 of chunk l >> 3
        const float4 n0 = __ldg(reinterpret_cast<const float4*>(pv.norm + col0 + 64 * (lane >> 3) + 32 * grp) + (lane & 7));
        {
          const unsigned long long pw = lds_u64_volatile(peer_thr_a);
          if ((uint32_t)(pw >> 32) == (uint32_t)item && row_ok) ladder_adopt(st, __uint_as_float((uint32_t)pw));
        }
        __syncwarp();
        reinterpret_cast<float4*>(snorm)[lane] = n0;
        __syncwarp();
        mbar_wait(&tfull[my_buf], my_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(my_buf * TN + 32 * grp);
        const bool tile_valid = (col0 + TN <= pv.n) && (col0 + TN <= w.chr_s || col0 >= w.chr_e);
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr, ra);
#pragma unroll 1
        for (int i = 0; i < 4; i += 2) {  // chunk i in ra, chunk i + 1 in rb
          const int cc = 64 * i + 32 * grp;  // first tile column of chunk i; chunk i + 1 starts 64 columns later
          tmem_ld_wait();
          tmem_ld32(taddr + (uint32_t)(64 * (i + 1)), rb);
          if (dbg_acc != nullptr && !dbg_done && blockIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 32; j++) dbg_acc[row * TN + cc + j] = __uint_as_float(ra[j]);
          }
          if (tile_valid) filter_chunk<true>(ra, snorm_a + 128 * i, col0 + cc, w, pv.n, st, be, spill_a);
          else filter_chunk<false>(ra, snorm_a + 128 * i, col0 + cc, w, pv.n, st, be, spill_a);
          tmem_ld_wait();
          if (i + 2 < 4) tmem_ld32(taddr + (uint32_t)(64 * (i + 2)), ra);
          if (dbg_acc != nullptr && !dbg_done && blockIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 32; j++) dbg_acc[row * TN + cc + 64 + j] = __uint_as_float(rb[j]);
          }
          if (tile_valid) filter_chunk<true>(rb, snorm_a + 128 * (i + 1), col0 + cc + 64, w, pv.n, st, be, spill_a);
          else filter_chunk<false>(rb, snorm_a + 128 * (i + 1), col0 + cc + 64, w, pv.n, st, be, spill_a);
        }
        dbg_done = true;
        // this group's half of the accumulator is drained: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[my_buf]); else mbar_arrive(&tempty[my_buf]); }
        // threshold maintenance
        const float thr_before = st.thr;
        ladder_advance(st);
        if (cv.diag && st.thr < thr_before) atomicAdd(cv.diag + 2, 1);
        // first exact selection once INIT_N entries are in, and overflow protection afterwards
        uint32_t need = __ballot_sync(0xffffffffu, (!st.ladder && st.cnt >= INIT_N) || st.cnt > WCX_CAND_CAP - TH);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int64_t s_slot = (int64_t)w.slot0 + (int64_t)(q * 32 + src) * w.slot_stride + grp;
          const int s_cnt = __shfl_sync(0xffffffffu, st.cnt, src);
          CompactResult cr;
          if (s_cnt <= 1024) cr = warp_compact<KEEP>(cv.ent + s_slot * WCX_CAND_CAP, s_cnt);
          else cr = warp_compact_stream<KEEP>(cv.ent + s_slot * WCX_CAND_CAP, s_cnt);
          if (lane == src) {
            st.thr = fminf(st.thr, cr.thr);
            st.lo = cr.lo;
            st.cnt = cr.kept;
            st.p0 = cr.probe[0]; st.p1 = cr.probe[1]; st.p2 = cr.probe[2]; st.p3 = cr.probe[3];
            st.c0 = cr.below[0]; st.c1 = cr.below[1]; st.c2 = cr.below[2]; st.c3 = cr.below[3];
            st.ladder = true;
            if (cv.diag) atomicAdd(cv.diag + (s_cnt <= 1024 ? 0 : 1), 1);
          }
        }
        if (st.thr < thr_before) sts_u64_volatile(my_thr_a, ((unsigned long long)(uint32_t)item << 32) | __float_as_uint(st.thr));
      }
      if (row_ok) { cv.cnt[slot] = st.cnt; cv.cut[slot] = st.thr; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tc_encode_tensor_map(const PrepView& pv, void* tmap_storage_host) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    WCX_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  CUtensorMap* map = reinterpret_cast<CUtensorMap*>(tmap_storage_host);
  const size_t esz = pv.f16 ? 2 : 4;
  cuuint64_t gdim[2] = {(cuuint64_t)pv.k_pad, (cuuint64_t)pv.n_pad};
  cuuint64_t gstride[1] = {(cuuint64_t)pv.k_pad * esz};
  cuuint32_t box[2] = {(cuuint32_t)(pv.f16 ? 2 * BK : BK), (cuuint32_t)TM};  // 128 bytes x 128 rows
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, pv.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pv.xc), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return 1; }
  return 0;
}

// Candidate tiles that lie inside the own chromosome of all 256 rows of a unit produce no candidates and are not
// multiplied (5.7 % of the tiles on a whole genome).  WCX_SWEEP_MULTIPLY_OWN=1 multiplies them anyway (measurement).
static int sweep_skip_own() { return getenv("WCX_SWEEP_MULTIPLY_OWN") == nullptr ? 1 : 0; }

template <bool F16>
static int launch_single(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv, void* tmap_storage, float* dbg,
                         int grid_override, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    WCX_CUDA_OK(cudaFuncSetAttribute(dist_topk_tc_kernel<false, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const CUtensorMap* map = reinterpret_cast<const CUtensorMap*>(tmap_storage);
  const int grid = grid_override > 0 ? grid_override : (nitems < sms ? nitems : sms);
  dist_topk_tc_kernel<false, F16><<<grid, TC_THREADS, Cfg<false>::SMEM, st>>>(*map, pv, items, nitems, cv, dbg, sweep_skip_own());
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_dist_topk_tc(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                        int32_t* work_counter, void* tmap_storage, cudaStream_t st) {
  (void)work_counter;
  if (nitems == 0) return 0;
  if (!pv.f16) { set_error("dist_topk_tc: only the f16 operand set is supported"); return 1; }
  return launch_single<true>(pv, items, nitems, cv, tmap_storage, nullptr, 0, st);
}

template <bool F16>
static int launch_pair(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv, void* tmap_storage, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    WCX_CUDA_OK(cudaFuncSetAttribute(dist_topk_tc_kernel<true, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const CUtensorMap* map = reinterpret_cast<const CUtensorMap*>(tmap_storage);
  const int npairs = nitems / 2;
  const int pairs_grid = npairs < sms / 2 ? npairs : sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs_grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg<true>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float* dbg = nullptr;
  WCX_CUDA_OK(cudaLaunchKernelEx(&cfg, dist_topk_tc_kernel<true, F16>, *map, pv, items, nitems, cv, dbg, sweep_skip_own()));
  return 0;
}

// pair mode: `items` holds an even number of entries, (2p, 2p + 1) sharing one candidate-column range
int launch_dist_topk_tc_pair(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv, void* tmap_storage,
                             cudaStream_t st) {
  if (nitems == 0) return 0;
  if (nitems & 1) { set_error("pair sweep: odd number of work items"); return 1; }
  if (!pv.f16) { set_error("dist_topk_tc: only the f16 operand set is supported"); return 1; }
  return launch_pair<true>(pv, items, nitems, cv, tmap_storage, st);
}

int launch_dist_topk_tc_debug(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                              void* tmap_storage, float* dbg_acc, cudaStream_t st) {
  if (!pv.f16) { set_error("dist_topk_tc: only the f16 operand set is supported"); return 1; }
  return launch_single<true>(pv, items, nitems, cv, tmap_storage, dbg_acc, 1, st);
}

}  // namespace wcx
