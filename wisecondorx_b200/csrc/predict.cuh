// Host launchers of the predict kernels (predict.cu).
#pragma once
#include "wcx_common.cuh"

namespace wcx {
int predict_red_blocks();
int launch_weights(const double* dist, int64_t n, int32_t k, double* out, cudaStream_t st);
int launch_optimal_cutoff(const double* dist, int64_t total, int32_t repeats, double* state3, double* partial,
                          cudaStream_t st);
int launch_coverage_project(const double* raw, int32_t B, int64_t bins_total, const int32_t* mask_pos, int64_t n,
                            const double* comps, const double* mean, int32_t ncomp, double* x, double* partial,
                            double* totals, double* tdots, cudaStream_t st);
int launch_gather_list(const int32_t* idx, const double* dist, int64_t n, int32_t k, const double* cutoff_dev,
                       const int64_t* cum_dev, int32_t nchr, int64_t ct, int32_t* gl, cudaStream_t st);
size_t radix_scratch_bytes(int32_t B, int64_t len);
int launch_nanmedians(const double* r, const double* z, int32_t B, int64_t len, void* scratch, double* m_lr, double* m_z,
                      cudaStream_t st);
int launch_normalize_repeat(const double* x, double* copy_a, double* copy_b, int32_t B, int64_t n, const int32_t* gl,
                            int32_t k, int64_t ct, double* z, double* r, double* nref, cudaStream_t st);
size_t segment_z_scratch_bytes(int32_t nseg);
int launch_segment_z(const double* nr, int32_t m, const int32_t* inflate_pos, const double* r, const double* w,
                     const int64_t* seg_se, const double* seg_r, int32_t nseg, double* z_out, double* partial, cudaStream_t st);
}  // namespace wcx
