// Host launchers of the newref preparation kernels (newref_pca.cu).
#pragma once
#include "wcx_common.cuh"

namespace wcx {
int launch_row_mean(const double* x, int64_t n, int32_t s, double* mean, cudaStream_t st);
int gram_chunks(int64_t n);
int gram_s_pad(int32_t s);
int launch_gram(const double* x, const double* mean, int64_t n, int32_t s, double* partial, double* g, cudaStream_t st);
int launch_pca_apply(const double* x, const double* mean, int64_t n, int32_t s, const double* u, const double* sigma,
                     int32_t ncomp, double* comps, double* corrected, cudaStream_t st);
int launch_col_medians(const double* x, int64_t n, int32_t s, unsigned long long* res_hi, unsigned long long* res_lo,
                       unsigned long long* cnt, double* out, cudaStream_t st);
int launch_row_sqdist(const double* x, int64_t n, int32_t s, const double* med, double* d, cudaStream_t st);
int launch_normalize_and_mask(const int32_t* counts, int64_t bins_total, int32_t s, const int32_t* mask_pos, int64_t n,
                              unsigned long long* colsum, double* out, cudaStream_t st);
}  // namespace wcx
