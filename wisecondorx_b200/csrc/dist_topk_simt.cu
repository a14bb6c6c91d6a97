// CUDA-core (fp32 FFMA) version of the fused distance + approximate top-k sweep.
//
// Reference loop replaced: get_ref_for_bins, newref_tools.py:255-278 (the `np.sum(np.power(
// chr_data - row, 2), 1)` distance at :260 and the sorted insert at :268-275).
//
// This kernel is the validated fallback/cross-check for the tcgen05 kernel (dist_topk_tc.cu): it
// produces candidate lists with identical semantics from the same tf32-rounded operands, so the
// two can be compared list-against-list on the GPU.  The N x N distance matrix never reaches
// HBM: each 128 x 128 tile lives in shared memory only long enough to be filtered against the
// per-row running thresholds.
#include "candidates.cuh"
#include "wcx_common.cuh"

namespace wcx {

namespace {
constexpr int TM = WCX_TILE_M;
constexpr int TN = WCX_TILE_N_SIMT;
constexpr int TK = 16;
constexpr int LDS_AB = TM + 4;
constexpr int LDD = TN + 1;
constexpr int SIMT_SMEM = (2 * TK * LDS_AB + TM * LDD) * 4;
}  // namespace

__global__ void __launch_bounds__(256)
dist_topk_simt_kernel(PrepView pv, const WorkItem* __restrict__ items, int nitems, CandView cv, int* work_counter) {
  extern __shared__ float smem[];
  float* As = smem;
  float* Bs = As + TK * LDS_AB;
  float* Dt = Bs + TK * LDS_AB;
  __shared__ float s_thr[TM];
  __shared__ int s_cnt[TM];
  __shared__ int s_item;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= nitems) break;
    const WorkItem w = items[item];
    if (tid < TM) { s_thr[tid] = __int_as_float(0x7f800000); s_cnt[tid] = 0; }
    __syncthreads();

    for (int ct = w.ct_begin; ct < w.ct_end; ct++) {
      const int col0 = ct * TN;
      if (col0 >= w.chr_s && col0 + TN <= w.chr_e) continue;  // tile entirely inside the own chromosome
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

      for (int k0 = 0; k0 < pv.k_pad; k0 += TK) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          int f = tid + 256 * h;
          int r = f >> 2, kq = f & 3;
          float4 a = *reinterpret_cast<const float4*>(pv.xc + (int64_t)(w.row0 + r) * pv.k_pad + k0 + 4 * kq);
          float4 b = *reinterpret_cast<const float4*>(pv.xc + (int64_t)(col0 + r) * pv.k_pad + k0 + 4 * kq);
          As[(4 * kq + 0) * LDS_AB + r] = a.x; As[(4 * kq + 1) * LDS_AB + r] = a.y;
          As[(4 * kq + 2) * LDS_AB + r] = a.z; As[(4 * kq + 3) * LDS_AB + r] = a.w;
          Bs[(4 * kq + 0) * LDS_AB + r] = b.x; Bs[(4 * kq + 1) * LDS_AB + r] = b.y;
          Bs[(4 * kq + 2) * LDS_AB + r] = b.z; Bs[(4 * kq + 3) * LDS_AB + r] = b.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; k++) {
          float a[8], b[8];
          float4 a0 = *reinterpret_cast<const float4*>(As + k * LDS_AB + ty * 8);
          float4 a1 = *reinterpret_cast<const float4*>(As + k * LDS_AB + ty * 8 + 4);
          float4 b0 = *reinterpret_cast<const float4*>(Bs + k * LDS_AB + tx * 8);
          float4 b1 = *reinterpret_cast<const float4*>(Bs + k * LDS_AB + tx * 8 + 4);
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) Dt[(ty * 8 + i) * LDD + tx * 8 + j] = acc[i][j];
      __syncthreads();

      // filter: warp `wid` owns rows 16*wid .. 16*wid+15 of the tile
      for (int rr = 0; rr < 16; rr++) {
        const int row = wid * 16 + rr;
        if (row >= w.nrows) break;
        float thr = s_thr[row];
        int cnt = s_cnt[row];
        const int64_t slot = (int64_t)w.slot0 + (int64_t)row * w.slot_stride;
        uint2* be = cv.ent + slot * WCX_CAND_CAP;
#pragma unroll
        for (int q = 0; q < TN / 32; q++) {
          const int c = q * 32 + lane;
          const int gcol = col0 + c;
          const float v = fmaf(-2.f, Dt[row * LDD + c], pv.norm[gcol]);
          const bool ok = (gcol < pv.n) && !(gcol >= w.chr_s && gcol < w.chr_e) && (v < thr);
          const uint32_t b = __ballot_sync(0xffffffffu, ok);
          if (ok) be[cnt + __popc(b & lt_mask)] = make_uint2(__float_as_uint(v), (uint32_t)gcol);
          cnt += __popc(b);
        }
        if (cnt > 1024 - TN) {  // warp_compact handles at most 1024 entries
          const CompactResult cr = warp_compact<WCX_CAND_KEEP>(be, cnt);
          thr = cr.thr;
          cnt = cr.kept;
        }
        if (lane == 0) { s_thr[row] = thr; s_cnt[row] = cnt; }
      }
      __syncthreads();
    }

    // finalize the lists of this work item
    for (int rr = 0; rr < 16; rr++) {
      const int row = wid * 16 + rr;
      if (row >= w.nrows) break;
      float thr = s_thr[row];
      int cnt = s_cnt[row];
      const int64_t slot = (int64_t)w.slot0 + (int64_t)row * w.slot_stride;
      if (lane == 0) { cv.cnt[slot] = cnt; cv.cut[slot] = thr; }
    }
  }
}

int launch_dist_topk_simt(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                          int32_t* work_counter, cudaStream_t st) {
  if (nitems == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WCX_CUDA_OK(cudaFuncSetAttribute(dist_topk_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SIMT_SMEM));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  WCX_CUDA_OK(cudaMemsetAsync(work_counter, 0, sizeof(int32_t), st));
  int grid = nitems < 2 * sms ? nitems : 2 * sms;
  dist_topk_simt_kernel<<<grid, 256, SIMT_SMEM, st>>>(pv, items, nitems, cv, work_counter);
  WCX_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wcx
