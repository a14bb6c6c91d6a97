// tcgen05 / TMEM / TMA version of the fused distance + approximate top-k sweep (sm_100a).
//
// Reference loop replaced: get_ref_for_bins, newref_tools.py:255-278.  The bins x bins squared
// distance is the one dense contraction of WisecondorX:
//     d(i,j) = |a_i|^2 + |b_j|^2 - 2 <a_i, b_j>
// The <a_i, b_j> tile (128 target bins x 256 candidate bins, K = samples) runs on the 5th-gen
// tensor cores as a TF32 UMMA with fp32 accumulation in TMEM.  Operands are the centred,
// tf32-rounded matrix Xc (newref_prep.cu), K-major, staged by TMA into 128B-swizzled shared
// memory tiles.  The epilogue never writes the distance tile anywhere: each epilogue thread
// owns one target row (= one TMEM lane), pulls 32 accumulator columns at a time with
// tcgen05.ld, forms v = |b_j|^2 - 2 acc and appends (v, j) to the row's candidate list only if
// v is below the row's running threshold (about 1 % of the elements).  Lists are compacted
// warp-cooperatively (candidates.cuh) and finished exactly by rerank.cu.
//
// Warp roles (192 + 64 threads, one CTA per SM, persistent over work items):
//   warp 0      TMA producer (one elected lane)
//   warp 1      MMA issuer   (one elected lane, tcgen05.mma cta_group::1 kind::tf32, M128 N256 K8)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warps 4..7  epilogue, warp w owns TMEM lanes 32*(w%4) .. +31
// Pipelines: smem full/empty mbarriers (TMA <-> MMA, 4 stages of 48 KB) and TMEM full/empty
// mbarriers (MMA <-> epilogue, 2 accumulator buffers).
#include <cuda.h>

#include "candidates.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {

constexpr int TM = WCX_TILE_M;          // 128
constexpr int TN = WCX_TILE_N_TC;       // 256
constexpr int BK = WCX_KBLOCK;          // 32 tf32 = 128 bytes
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * BK * 4;    // 16 KB
constexpr int B_BYTES = TN * BK * 4;    // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TC_THREADS = 256;
constexpr int TC_SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, K-major both,
// N = 256 (n_dim = N >> 3 at bit 17), M = 128 (m_dim = M >> 4 at bit 24)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

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

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
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

__device__ __forceinline__ bool tile_skipped(const WorkItem& w, int ct) {
  const int col0 = ct * TN;
  return col0 >= w.chr_s && col0 + TN <= w.chr_e;
}

}  // namespace

__global__ void __launch_bounds__(TC_THREADS, 1)
dist_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, PrepView pv, const WorkItem* __restrict__ items,
                    int nitems, CandView cv, float* __restrict__ dbg_acc) {
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES]
  uint64_t* empty = bars + STAGES;         // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;     // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = pv.k_pad / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const WorkItem w = items[item];
        for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
          if (tile_skipped(w, ct)) continue;
          const int col0 = ct * TN;
          for (int kb = 0; kb < kblocks; kb++) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* sa = smem + stage * STAGE_BYTES;
            unsigned char* sb = sa + A_BYTES;
            mbar_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmap, &full[stage], kb * BK, w.row0);
            tma_load_2d(sb, &tmap, &full[stage], kb * BK, col0);
            tma_load_2d(sb + B_BYTES / 2, &tmap, &full[stage], kb * BK, col0 + TN / 2);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const WorkItem w = items[item];
        for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
          if (tile_skipped(w, ct)) continue;
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
              // advance 8 tf32 = 32 bytes along K inside the 128B swizzle row: +2 in the >>4 address field
              umma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[buf]);  // accumulator complete
          if (++buf == 2) { buf = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread owns one target row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int buf = 0;
    uint32_t tphase = 0;
    bool dbg_done = false;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const WorkItem w = items[item];
      const bool row_ok = row < w.nrows;
      const int64_t slot = (int64_t)w.slot0 + (int64_t)row * w.slot_stride;
      float* bv = cv.val + (row_ok ? slot : 0) * WCX_CAND_CAP;
      int32_t* bi = cv.idx + (row_ok ? slot : 0) * WCX_CAND_CAP;
      float thr = row_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);  // +inf / -inf
      int cnt = 0;
      for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
        if (tile_skipped(w, ct)) continue;
        const int col0 = ct * TN;
        mbar_wait(&tfull[buf], tphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TN);
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c0, r);
          tmem_ld_wait();
          const int g0 = col0 + c0;
          if (dbg_acc != nullptr && !dbg_done && blockIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 32; j++) dbg_acc[row * TN + c0 + j] = __uint_as_float(r[j]);
          }
          // column validity is warp-uniform: inside the matrix and outside the own chromosome
          const bool all_valid = (g0 + 32 <= pv.n) && (g0 + 32 <= w.chr_s || g0 >= w.chr_e);
          const float4* nrm4 = reinterpret_cast<const float4*>(pv.norm + g0);
          if (all_valid) {
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) {
              const float4 nb = __ldg(nrm4 + j4);
              const float nbs[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const float v = fmaf(-2.f, __uint_as_float(r[4 * j4 + e]), nbs[e]);
                if (v < thr) { bv[cnt] = v; bi[cnt] = g0 + 4 * j4 + e; cnt++; }
              }
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) {
              const float4 nb = __ldg(nrm4 + j4);
              const float nbs[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const int g = g0 + 4 * j4 + e;
                const bool ok = (g < pv.n) && !(g >= w.chr_s && g < w.chr_e);
                const float v = fmaf(-2.f, __uint_as_float(r[4 * j4 + e]), nbs[e]);
                if (ok && v < thr) { bv[cnt] = v; bi[cnt] = g; cnt++; }
              }
            }
          }
        }
        dbg_done = true;
        // accumulator buffer drained: hand it back to the MMA warp before the (rare) compaction
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (++buf == 2) { buf = 0; tphase ^= 1; }
        // lists that could overflow on the next tile get compacted, one row at a time, warp-cooperatively
        uint32_t need = __ballot_sync(0xffffffffu, cnt > WCX_CAND_CAP - TN);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int64_t s_slot = (int64_t)w.slot0 + (int64_t)(q * 32 + src) * w.slot_stride;
          const int s_cnt = __shfl_sync(0xffffffffu, cnt, src);
          const float t = warp_compact(cv.val + s_slot * WCX_CAND_CAP, cv.idx + s_slot * WCX_CAND_CAP, s_cnt);
          if (lane == src) { thr = t; cnt = WCX_CAND_KEEP; }
        }
      }
      // finalize this work item's lists
      uint32_t need = __ballot_sync(0xffffffffu, cnt > WCX_CAND_KEEP);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int64_t s_slot = (int64_t)w.slot0 + (int64_t)(q * 32 + src) * w.slot_stride;
        const int s_cnt = __shfl_sync(0xffffffffu, cnt, src);
        const float t = warp_compact(cv.val + s_slot * WCX_CAND_CAP, cv.idx + s_slot * WCX_CAND_CAP, s_cnt);
        if (lane == src) { thr = t; cnt = WCX_CAND_KEEP; }
      }
      if (row_ok) { cv.cnt[slot] = cnt; cv.cut[slot] = thr; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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
  cuuint64_t gdim[2] = {(cuuint64_t)pv.k_pad, (cuuint64_t)pv.n_pad};
  cuuint64_t gstride[1] = {(cuuint64_t)pv.k_pad * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)TM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pv.xc), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return 1; }
  return 0;
}

int launch_dist_topk_tc(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                        int32_t* work_counter, void* tmap_storage, cudaStream_t st) {
  (void)work_counter;
  if (nitems == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WCX_CUDA_OK(cudaFuncSetAttribute(dist_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const CUtensorMap* map = reinterpret_cast<const CUtensorMap*>(tmap_storage);
  int grid = nitems < sms ? nitems : sms;
  float* dbg = nullptr;
  dist_topk_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(*map, pv, items, nitems, cv, dbg);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_dist_topk_tc_debug(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                              void* tmap_storage, float* dbg_acc, cudaStream_t st) {
  WCX_CUDA_OK(cudaFuncSetAttribute(dist_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  const CUtensorMap* map = reinterpret_cast<const CUtensorMap*>(tmap_storage);
  dist_topk_tc_kernel<<<1, TC_THREADS, TC_SMEM, st>>>(*map, pv, items, nitems, cv, dbg_acc);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
