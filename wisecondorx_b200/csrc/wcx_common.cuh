// Shared declarations for the wisecondorx_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define WCX_CAND_CAP 4096   // per candidate-list capacity (entries)
#define WCX_CAND_KEEP 512   // SIMT kernel (one list per row and split): thr is only lowered to t when >= KEEP entries are known below t
#define WCX_CAND_KEEP_TC 256  // tcgen05 kernel: per list of one epilogue group (two lists per row and split share thresholds)
#define WCX_TILE_M 128      // target rows per work item
#define WCX_TILE_N_SIMT 128 // candidate columns per tile, CUDA-core kernel
#define WCX_TILE_N_TC 256   // candidate columns per tile, tcgen05 kernel
#define WCX_KBLOCK 32       // K elements (tf32) per pipeline stage = one 128-byte swizzle row

namespace wcx {

void set_error(const std::string& msg);

#define WCX_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      wcx::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +  \
                     __FILE__ + ":" + std::to_string(__LINE__));                           \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

// One work item of the distance/top-k sweep: TILE_M target rows of one chromosome against the
// candidate column tiles [ct_begin, ct_end) (tile width depends on the kernel).
struct WorkItem {
  int32_t row0;     // first target row (global bin index)
  int32_t nrows;    // <= WCX_TILE_M
  int32_t chr_s;    // [chr_s, chr_e) = the target rows' own chromosome: excluded candidates
  int32_t chr_e;
  int32_t ct_begin; // candidate column tile range of this split
  int32_t ct_end;
  int32_t slot0;    // candidate-list slot of row0 (slot = (row - row_begin) * nsplit + split)
  int32_t slot_stride;  // = nsplit
};

// Device-side view of the prepared (centred, tf32-rounded) matrix
struct PrepView {
  const float* xc;     // [n_pad, k_pad] row-major, zero padded
  const float* norm;   // [n_pad]  sum_s xc^2 (fp32)
  int64_t n;           // bins
  int64_t n_pad;
  int32_t s;           // samples
  int32_t k_pad;       // S rounded up to the K block of the operand type (32 tf32 / 64 f16 elements = 128 bytes)
  // f16 operands (dist_topk_tc.cu, F16 = true): xc points at __half data holding (X - mean) * scale[0]; norm and
  // every list value are in scaled units, scale[1] = scale[0]^2 converts an exact distance.  nullptr: unscaled.
  const double* scale;
  float abs_err;       // per-element absolute rounding error bound of the operand conversion (scaled units)
  int32_t f16;
  // f16 operands: measured conversion residuals.  normres[row] = (norm[row], |x_row - x^_row| rounded up); rho_max[0] =
  // max over the finite rows of (|x - x^| - tau) / |x^|, tau = 2 sqrt(k_pad) abs_err.  nullptr: worst-case unit round-off.
  const float2* normres;
  const float* rho_max;
};

struct CandView {
  uint2* ent;      // [slots, CAP]   .x = float bits of v = norm[j] - 2 * dot(i, j), .y = global candidate bin j
  int32_t* cnt;    // [slots]
  float* cut;      // [slots]        every non-listed candidate of the slot has v >= cut
  int32_t* diag;   // [8] counters: 0 exact compactions, 1 streamed compactions, 2 ladder steps, 3 adoptions
};

// ---- newref kernels (host launchers; all asynchronous on `st`) ---------------------------
int launch_col_stats(const double* x, int64_t n, int32_t s, double* colsum, double* colcnt, unsigned long long* absmax,
                     cudaStream_t st);
int launch_center_round(const double* x, int64_t n, int32_t s, const double* colsum, const double* colcnt,
                        float* xc, float* norm, int64_t n_pad, int32_t k_pad, cudaStream_t st);
// f16 operands: scale[0] = power of two that maps max|x| + max|mean| below 2^14, scale[1] = its square
int launch_center_round_f16(const double* x, int64_t n, int32_t s, const double* colsum, const double* colcnt,
                            const unsigned long long* absmax, double* scale, void* xh, float* norm, float2* normres,
                            float* rho_max, double tau, int64_t n_pad, int32_t k_pad, cudaStream_t st);
int launch_transpose_cols(const double* x, int64_t n, int32_t s, const int32_t* ids, int32_t m,
                          double* xt, cudaStream_t st);
int launch_dist_topk_simt(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                          int32_t* work_counter, cudaStream_t st);
int launch_dist_topk_tc(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                        int32_t* work_counter, void* tmap_storage, cudaStream_t st);
int launch_dist_topk_tc_pair(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv, void* tmap_storage,
                             cudaStream_t st);
int tc_encode_tensor_map(const PrepView& pv, void* tmap_storage_host);  // element type from pv.f16
int launch_rerank(const double* x, const PrepView& pv, CandView cv, int32_t nlists, const int64_t* cum_dev,
                  int32_t nchr, int64_t row_begin, int64_t row_end, int32_t k, int32_t gonosomal,
                  int32_t* idx_out, double* dist_out, int32_t* fail_flags, const int32_t* sum_plan,
                  int32_t plan_len, const double* xp, int32_t sp, const int32_t* leaf_dev, int32_t nleaves, cudaStream_t st);
// leaf-major copy of X for the re-rank gather (rerank.cu): layout from the summation plan, then the copy itself
int build_leaf_layout(const int32_t* plan, int32_t plan_len, std::vector<int32_t>& perm, std::vector<int32_t>& desc);
int launch_permute_rows(const double* x, int64_t n, int32_t s, const int32_t* perm_dev, int32_t sp, double* xp, cudaStream_t st);
int launch_exact_rows(const double* x, int64_t n, int32_t s, const int64_t* cum_dev, int32_t nchr,
                      int64_t row_begin, const int32_t* fail_rows, int32_t nfail, int32_t k,
                      int32_t* idx_out, double* dist_out, double* scratch, const int32_t* sum_plan,
                      int32_t plan_len, cudaStream_t st);
int64_t null_ratio_staging_doubles(int64_t n, int32_t m);
int launch_validate_positions(const int32_t* idx, int64_t count, int64_t n, int32_t* bad, cudaStream_t st);
int launch_null_ratios(const double* xt, int64_t n, const int32_t* idx, int64_t row_begin, int64_t row_end,
                       int32_t k, int32_t m, double* out, cudaStream_t st);

// NumPy pairwise-summation plan for a reduction of length s (see rerank.cu)
int build_sum_plan(int32_t s, int32_t* plan, int32_t cap);
// (offset, length, #adds) per leaf of the plan + maximum stack depth (rerank.cu)
int plan_to_leaves(const int32_t* plan, int32_t plan_len, std::vector<int32_t>& leaves, int32_t* max_depth);

}  // namespace wcx
