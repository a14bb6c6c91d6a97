/* wcx_b200 -- C-ABI of the B200-native WisecondorX numeric core (libwcx_b200.so).
 *
 * The reference (CenterForMedicalGeneticsGhent/WisecondorX v1.2.10) is pure Python + R and has
 * no FFI of its own; the seams this library replaces are ordinary Python call sites.  Each entry
 * point below names the reference function (file:line under src/wisecondorx/) whose arithmetic
 * it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; wcx_last_error() gives the text
 *     (thread-local).  There is NO CPU fallback: without a usable sm_100 device wcx_create fails.
 *   - the caller owns every buffer.  Pointers are C-contiguous.  `on_device` flags say whether a
 *     pointer is a host pointer (0) or a device pointer on the context's device (1).
 *   - all work of a context is issued on one CUDA stream (own stream by default, or the one given
 *     to wcx_set_stream, e.g. torch's current stream); calls with host outputs synchronise it.
 *   - a context is not re-entrant; use one context per host thread / per GPU.
 *   - entry points WITHOUT a wcx_ctx argument (wcx_host_*, wcx_predict_assemble, wcx_cbs_pack_count / _pack / _unpack)
 *     run on host threads and touch no device: they are the native form of Python glue the reference runs between
 *     the functions above (stacking the samples, result assembly, CBS.R around DNAcopy::segment), not substitutes
 *     for any kernel.  `threads` caps the host threads of a call.
 */
#ifndef WCX_B200_H
#define WCX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wcx_ctx wcx_ctx;

/* kernel selector for the distance sweep */
#define WCX_KERNEL_AUTO 0 /* = WCX_KERNEL_TC2H */
#define WCX_KERNEL_TC 1   /* (round 1: tf32, one CTA per SM; removed) now an alias of WCX_KERNEL_TCH */
#define WCX_KERNEL_SIMT 2 /* CUDA-core fp32 kernel (dist_topk_simt.cu), cross-check path */
#define WCX_KERNEL_EXACT 3 /* brute-force float64 rows only (exact_rows_kernel), slow; tests and ref_size > 400 */
#define WCX_KERNEL_TC2 4  /* (round 1: tf32, CTA pairs; removed) now an alias of WCX_KERNEL_TC2H */
#define WCX_KERNEL_TC2H 5 /* tcgen05 / TMEM / TMA kernel, 2-CTA pairs (cta_group::2), scaled f16 operands: the product path */
#define WCX_KERNEL_TCH 6  /* the same kernel with one CTA per SM (cta_group::1), cross-check path */

int wcx_version(void);
const char* wcx_last_error(void);

/* Creates a context on CUDA device `device` (must be compute capability 10.x). */
int wcx_create(int32_t device, wcx_ctx** out);
void wcx_destroy(wcx_ctx* ctx);
/* Issue subsequent work on `cuda_stream` (a cudaStream_t); NULL restores the context's own stream. */
int wcx_set_stream(wcx_ctx* ctx, void* cuda_stream);
int wcx_sync(wcx_ctx* ctx);
/* Page-locked host memory (cudaHostAlloc) for buffers that cross PCIe in every call: every entry point accepts
 * ordinary host pointers too, but copies from / to pageable memory are staged by the driver at a fraction of the
 * link rate.  The host mirror (wisecondorx_b200/_lib.py PinnedPool) hands these out as NumPy arrays. */
int wcx_host_alloc(uint64_t bytes, void** out);
int wcx_host_free(void* p);

/* ---- host-side preparation of newref (host threads, no device work, no context) -----------
 * wcx_host_stack_counts: the read counts of all samples as ONE matrix int32 [rows, samples] (the fill of `all_data`,
 * newref_tools.py:81-92 / :114-122): columns[s * nchr + c] -> the int32 counts of chromosome c of sample s
 * (lens[s * nchr + c] of them), chromosome c occupies rows offs[c] - offs[0] ... offs[c + 1] - offs[0], shorter
 * samples are zero padded.
 * wcx_host_bin_sums: out[b] = sum over the samples j of counts[b, cols[j]] / col_sum[j] (cols == NULL: all columns) --
 * np.sum(all_data / sum_per_sample, 1) of get_mask (newref_tools.py:94-97) with one division per element and the
 * additions in the order of NumPy's pairwise summation, i.e. the same float64 value. */
int wcx_host_stack_counts(const int32_t* const* columns, const int64_t* lens, int32_t samples, int32_t nchr,
                          const int64_t* offs, int32_t* out, int32_t threads);
int wcx_host_bin_sums(const int32_t* counts, int64_t bins, int32_t samples, const int32_t* cols, int32_t ncols,
                      const double* col_sum, double* out, int32_t threads);

/* ---- host side of the segmentation step (host threads, no device work, no context): include/CBS.R around
 * DNAcopy::segment, for a batch of samples.  r_rows[s] / w_rows[s] -> the log ratios and weights of sample s on the
 * concatenated bin axis of its chromosomes (0 = no data), offs[offs_at[s] ... offs_at[s + 1] - 1] its chromosome
 * offsets (nchr + 1 ascending values); (sample, chromosome) "slots" are numbered in that order.
 * wcx_cbs_pack_count: counts[slot] = bins with ratio != 0 (CBS.R:41).
 * wcx_cbs_pack: the NA-free series of all samples back to back, sample s from entry at[s]: y = ratios, w = weights with
 *   0 replaced by 1 (CBS.R:42), pos int32 = bin of every entry on the sample's axis; long_gaps[slot] = runs of more
 *   than na_thresh NA bins between two entries of the chromosome (what wcx_cbs_unpack may cut at).
 * wcx_cbs_unpack (CBS.R:80-129): series i = entries off[i] ... off[i + 1] of y / w / pos (the series handed to
 *   wcx_cbs_segment, which returned nseg[i] ascending exclusive ends per series, back to back in `ends`); chr_start[i] =
 *   first bin of its chromosome; output slots slot[i] ... slot[i + 1] (at least nseg[i] + long gaps of the series).
 *   Every segment is cut at the NA runs longer than na_thresh inside it, pieces of at most one bin are dropped, and each
 *   piece is written as (out_series = i, out_s 0-based start, out_e exclusive end on the chromosome, out_r = weighted
 *   mean of its entries, summed in NumPy's pairwise order); unused slots get out_series = -1. */
int wcx_cbs_pack_count(const double* const* r_rows, const int64_t* offs, const int64_t* offs_at, int32_t samples,
                       int64_t* counts, int32_t threads);
int wcx_cbs_pack(const double* const* r_rows, const double* const* w_rows, const int64_t* offs, const int64_t* offs_at,
                 int32_t samples, const int64_t* at, int64_t na_thresh, double* y, double* w, int32_t* pos,
                 int64_t* long_gaps, int32_t threads);
int wcx_cbs_unpack(const int32_t* pos, const double* y, const double* w, const int64_t* off, int32_t series,
                   const int32_t* ends, const int32_t* nseg, const int64_t* chr_start, const int64_t* slot,
                   int64_t na_thresh, int32_t* out_series, int64_t* out_s, int64_t* out_e, double* out_r,
                   int32_t threads);

/* ---- host-side text of the per-bin table of predict (no device work, no context) ------------
 * wcx_host_format_bins: the lines of <outid>_bins.bed for one chromosome (_generate_bins_bed, predict_output.py:59-84):
 * "chr \t start \t end \t chr:start-end \t ratio \t zscore \n" for bins i = 0 .. n - 1 with start = i * binsize + 1,
 * end = (i + 1) * binsize; a value of 0 prints as "nan", any other as Python's repr(float) (= str of the NumPy scalar
 * the reference prints).  out: at least n * (2 * strlen(chr_name) + 138) bytes; *len = bytes written.
 * wcx_host_format_repr: repr(x[i]) + "\n" for every value (26 bytes per value needed); test hook of the formatter. */
int wcx_host_format_bins(const char* chr_name, int64_t binsize, const double* r, const double* z, int64_t n, char* out,
                         int64_t cap, int64_t* len);
int wcx_host_format_repr(const double* x, int64_t n, char* out, int64_t cap, int64_t* len);

/* ---- newref ------------------------------------------------------------------------------
 * Replaces get_reference (newref_tools.py:155-224): get_ref_for_bins (:255-278) and the
 * null-ratio loop (:210-224).
 *
 * wcx_newref_load: X = pca_corrected_data [N, S] float64 row-major (newref_control.py:68 / :91),
 *   per/cum = masked_bins_per_chr(_cum) (C entries; C > 22 selects the gonosomal behaviour of
 *   newref_tools.py:186-191).  Copies (or borrows, if x_on_device) X and prepares the
 *   tensor-core operands.  A borrowed device X must stay alive until the next load/destroy.
 *   x_on_device: 0 host pointer, 1 device pointer on the context's device (borrowed), 2 the corrected matrix left in
 *   this context by wcx_pca_apply (x ignored), 3 device pointer on ANY device of the box (copied once, e.g. the
 *   matrix prepared on GPU 0 when the target-bin parts are spread over several GPUs).
 */
int wcx_newref_load(wcx_ctx* ctx, const double* x, int64_t n, int32_t s, const int64_t* per,
                    const int64_t* cum, int32_t c, int32_t x_on_device);

/* indexes int32 [rows, k] and distances float64 [rows, k] for target bins [row_begin, row_end)
 * == the part handled by get_reference(part, split_parts) (newref_tools.py:168, :244-247).
 * Bit-exact with the reference: same distances (NumPy summation order), same (distance,
 * position) order, -1 / 1e10 fillers, placeholder rows (0, 1.0) in gonosomal mode.
 * Either output may be NULL (the result stays on the device for wcx_newref_null_ratios). */
int wcx_newref_topk(wcx_ctx* ctx, int64_t row_begin, int64_t row_end, int32_t k, int32_t kernel,
                    int32_t* idx_out, double* dist_out, int32_t out_on_device);

/* null_ratios float64 [rows, M]: log2(col[b] / median(col[idx[b, :]])) for the M sample columns
 * `sample_ids` (host int32; the reference draws them with random.sample at newref_tools.py:215).
 * idx == NULL uses the indexes left on the device by the last wcx_newref_topk call for the same
 * row range and k. */
int wcx_newref_null_ratios(wcx_ctx* ctx, const int32_t* idx, int32_t idx_on_device, int64_t row_begin,
                           int64_t row_end, int32_t k, const int32_t* sample_ids, int32_t m,
                           double* out, int32_t out_on_device);

/* Top-k and null ratios of rows [row_begin, row_end) in one call -- the body of get_reference
 * (newref_tools.py:176-224) after the matrix is resident.  One sweep over the candidate axis nominates candidates for
 * every row; the rows then go through the exact re-rank and the null-ratio kernels.  With host outputs this happens in
 * row blocks -- the null-ratio kernels of a finished block run on a second stream next to the re-rank of the following
 * block and the D2H copy of a finished block overlaps the kernels of the next one; with device outputs every region of
 * rows is one re-rank launch followed by one null-ratio launch.  m == 0 skips the null ratios.  Outputs as wcx_newref_topk /
 * wcx_newref_null_ratios.  ref_size: 1..400 on the tensor-core path; 401..512 is served by the brute-force float64
 * row kernel (WCX_KERNEL_EXACT) -- same results, much slower. */
int wcx_newref_reference(wcx_ctx* ctx, int64_t row_begin, int64_t row_end, int32_t k, int32_t kernel,
                         const int32_t* sample_ids, int32_t m, int32_t* idx_out, double* dist_out,
                         double* null_out, int32_t out_on_device);

/* One-call host-to-host form of get_reference: wcx_newref_load + wcx_newref_reference. */
int wcx_get_reference(wcx_ctx* ctx, const double* x, int64_t n, int32_t s, const int64_t* per,
                      const int64_t* cum, int32_t c, int32_t k, int64_t row_begin, int64_t row_end,
                      const int32_t* sample_ids, int32_t m, int32_t kernel, int32_t* idx_out,
                      double* dist_out, double* null_out);

/* Counters of the last wcx_newref_topk call: out[0] = work items, out[1] = rows recomputed by the
 * exact brute-force path, out[2] = kernel launches issued since wcx_create, out[3] = rows swept by the main
 * sweep launch (the remaining rows: the partial last round, launched beside the re-rank), out[4] = sweep kernel used, out[5] = exact list compactions, out[6] = streamed
 * (overflow) compactions, out[7] = ladder threshold steps. */
int wcx_newref_stats(wcx_ctx* ctx, int64_t* out8);

/* Device time in milliseconds of the stages of the last wcx_newref_topk call, measured with CUDA
 * events on the context's stream: out[0] = sweep (distance + approximate top-k), out[1] = exact
 * re-rank, out[2] = brute-force rows, out[3] = last wcx_newref_null_ratios, out[4] = last
 * wcx_newref_load preparation kernels; out[5] = exact float64 distances evaluated by the re-rank,
 * out[6] = list entries below the cut gathered by the re-rank (counters, not times), out[7] = how long the sweep's
 * partial last round (launched beside the re-rank) runs past the main sweep launch; out[0] includes it. */
int wcx_newref_stage_ms(wcx_ctx* ctx, double* out8);

/* ---- newref preparation ---------------------------------------------------------------------
 * Replaces the numeric body of tool_newref_prep (newref_control.py:35-58).
 *
 * normalize_and_mask (newref_tools.py:110-129): counts int32 [bins_total, S] = per-sample read counts
 * stacked per chromosome (zero padded), mask_pos int32 [n] = np.flatnonzero(mask); out float64 [n, S] =
 * counts[mask_pos] / column totals.  Bit-exact (integer totals).  counts == NULL reuses the count matrix uploaded by the
 * previous call (same bins_total and S): the redo after the PCA-distance filter changes only the mask. */
int wcx_newref_normalize_and_mask(wcx_ctx* ctx, const int32_t* counts, int64_t bins_total, int32_t s,
                                  const int32_t* mask_pos, int64_t n, double* out, int32_t out_on_device);
/* train_pca (newref_tools.py:138-147), exact-PCA formulation in three steps:
 *   wcx_pca_gram : x [n, S] -> mean_out [n] (per-bin mean over samples = pca.mean_) and the S x S Gram
 *                  matrix of the centred data (host, row-major).  x stays on the device.
 *   (host)       : eigen-decomposition of the Gram matrix (S x S), u = top eigenvectors [S, ncomp],
 *                  sigma = sqrt(eigenvalues).
 *   wcx_pca_apply: comps_out [ncomp, n] (= pca.components_ up to sklearn's sign convention, applied
 *                  by the caller) and corrected_out [n, S] = x / (inverse_transform(transform(x))).
 *                  corrected_out may be NULL: the result then stays on the device for
 *                  wcx_pca_distance(corrected = NULL). */
int wcx_pca_gram(wcx_ctx* ctx, const double* x, int64_t n, int32_t s, int32_t x_on_device, double* mean_out,
                 double* gram_out);
int wcx_pca_apply(wcx_ctx* ctx, const double* u, const double* sigma, int32_t ncomp, double* comps_out,
                  double* corrected_out, int32_t corrected_on_device);
/* PCA-distance filter (newref_control.py:40-41): med_out [S] = np.median(corrected, axis=0),
 * d_out [n] = sum_s (corrected[b, s] - med[s])^2.  The MAD cutoff (:42-46) is a host scalar step. */
int wcx_pca_distance(wcx_ctx* ctx, const double* corrected, int64_t n, int32_t s, int32_t on_device,
                     double* med_out, double* d_out);
/* Device-resident chain (no [n, S] matrix crosses PCIe between the steps of tool_newref_prep and get_reference):
 *   wcx_newref_normalize_and_mask(out = NULL, out_on_device = 1)  keeps the matrix in the context,
 *   wcx_pca_gram(x = NULL, x_on_device = 1)                       reads it,
 *   wcx_pca_apply(corrected_out = NULL) / wcx_pca_distance(corrected = NULL)  keep / read the corrected matrix,
 *   wcx_newref_load(x = NULL, x_on_device = 2)                    loads the corrected matrix for get_reference.
 * wcx_prep_fetch copies a resident matrix to the host (which = 0: normalised + masked, 1: corrected). */
int wcx_prep_fetch(wcx_ctx* ctx, int32_t which, int64_t n, int32_t s, double* out);
/* Device pointer of a resident matrix (which as above) after the context's stream has drained: input of
 * wcx_newref_load(x_on_device = 3) on another context / GPU.  Valid until the next preparation call on `ctx`. */
int wcx_prep_device_ptr(wcx_ctx* ctx, int32_t which, int64_t n, int32_t s, const double** out);
/* Device milliseconds: out[0] = mean + Gram, out[1] = components + correction, out[2] = medians + distances. */
int wcx_newref_prep_stage_ms(wcx_ctx* ctx, double* out4);

/* ---- predict ------------------------------------------------------------------------------
 * Replaces normalize (predict_control.py:21-39) = coverage_normalize_and_mask (predict_tools.py:32),
 * project_pc (:56), get_weights (:152), get_optimal_cutoff (:74), normalize_repeat/_normalize_once
 * (:94-142); and get_z_score (overall_tools.py:88-119).  All float64.
 *
 * A context holds up to three device-resident reference sets (set_id 0 = autosomal keys "",
 * 1 = ".F", 2 = ".M" of the reference .npz): indexes int32 [n, k], distances float64 [n, k],
 * masked_bins_per_chr(_cum) (nchr), pca_components float64 [ncomp, n], pca_mean [n], and
 * mask_pos int32 [n] = positions of the True entries of the set's `mask` (np.flatnonzero) in the
 * concatenated unmasked bin axis of length bins_total = sum(bins_per_chr). */
int wcx_predict_load_ref(wcx_ctx* ctx, int32_t set_id, const int32_t* idx, const double* dist, int64_t n,
                         int32_t k, const int64_t* per, const int64_t* cum, int32_t nchr,
                         const double* pca_components, const double* pca_mean, int32_t ncomp,
                         const int32_t* mask_pos, int64_t bins_total);
/* get_weights: out[i] = 1 / mean(sqrt(distances[i, :])), out float64 [n] (host). */
int wcx_predict_weights(wcx_ctx* ctx, int32_t set_id, double* out);
/* get_optimal_cutoff over the set's distances (the reference always uses set 0). */
int wcx_predict_optimal_cutoff(wcx_ctx* ctx, int32_t set_id, int32_t repeats, double* cutoff_out);
/* coverage normalisation + PCA projection + the three within-sample normalisation passes for a
 * batch of B samples.  raw float64 [B, bins_total]: per-chromosome read counts padded/truncated to
 * the set's bins_per_chr and concatenated (predict_tools.py:37-44).  cp/ct as in
 * predict_control.py:22-29 (cp = 0, ct = 0 for autosomes; cp = 22, ct = cum[21] for gonosomes).
 * Outputs (host): z, r, nref float64 [B, n - ct]; m_lr, m_z float64 [B]. */
int wcx_predict_normalize(wcx_ctx* ctx, int32_t set_id, const double* raw, int32_t b, double cutoff,
                          int32_t cp, int64_t ct, double* z, double* r, double* nref, double* m_lr,
                          double* m_z);
/* Result assembly of a batch of samples that share one reference gender, on HOST threads (no device work, no
 * context): the stacking of the autosomal and gonosomal `normalize` results with the z-score shift (main.py:242-246),
 * get_post_processed_result (predict_control.py:49-63), inflate_results (predict_tools.py:163-170) and log_trans
 * (predict_tools.py:180-193) in one streaming pass per sample.  lr_aut float64 [samples, n_aut]: log2 of the ratios of
 * the autosomal call (taken by the caller; -inf / NaN where the ratio has no logarithm); z_aut / nref_aut float64
 * [*, n_aut]: z-scores and reference-bin counts of that call, sample s in row aut_row[s] (the call may have served
 * samples of the other gender too); lr_gon / z_gon / nref_gon float64 [samples, n_gon]: the same of the gonosomal
 * call; weights float64 [n_aut + n_gon] (main.py:246-247, shared by the samples); m_lr / m_z float64 [samples]: median
 * log ratio and median z-score of the autosomal call; mask uint8 [bins]: the reference's mask of this gender.
 * Outputs [samples, bins]: log ratios, z-scores, weights (0 = no data) and inflate int32 = row of the stacked null
 * ratios for every bin, -1 where the bin is masked or has fewer than minrefbins reference bins (the map
 * wcx_segment_zscore takes).  Returns 2 when there are fewer results than kept bins (the reference raises IndexError). */
int wcx_predict_assemble(const double* lr_aut, const double* z_aut, const double* nref_aut, int64_t n_aut,
                         const int32_t* aut_row, const double* lr_gon, const double* z_gon, const double* nref_gon,
                         int64_t n_gon, const double* weights, const double* m_lr, const double* m_z,
                         int32_t samples, double minrefbins, const uint8_t* mask, int64_t bins, double* out_r,
                         double* out_z, double* out_w, int32_t* out_inflate, int32_t threads);
/* get_z_score: nr = null ratios float64 [n_masked, m]; inflate_pos int32 [bins_total] = row of nr
 * for each unmasked bin or -1; r, w float64 [bins_total] = post-processed log2 ratios (0 = no
 * data) and weights; segments as [start, end) offsets into the concatenated bin axis with their
 * ratio seg_r.  z_out float64 [nseg]; NaN where the reference returns the string "nan".  nr == NULL reuses the null
 * ratios uploaded by the previous call (same n_masked and m): they belong to the reference, not to the sample. */
int wcx_segment_zscore(wcx_ctx* ctx, const double* nr, int64_t n_masked, int32_t m,
                       const int32_t* inflate_pos, const double* r, const double* w, int64_t bins_total,
                       const int64_t* seg_se, const double* seg_r, int32_t nseg, double* z_out);
/* Device milliseconds of the last predict calls: out[0] = coverage + projection, out[1] = gather lists + the
 * three normalisation passes + medians, out[2] = segment z-score, out[3] = last wcx_cbs_segment, out[4] = gather-list
 * build (only when the (reference set, cutoff) changed), out[5] = the three passes, out[6] = the two medians,
 * out[7] reserved. */
int wcx_predict_stage_ms(wcx_ctx* ctx, double* out8);

/* ---- CBS -----------------------------------------------------------------------------------
 * Replaces the R bridge: exec_cbs (predict_tools.py:242-263) -> exec_R (overall_tools.py:65-80) ->
 * include/CBS.R:70-73, DNAcopy::segment(CNA(y, chrom, x), alpha = alpha, weights = w) with DNAcopy
 * 1.76 defaults (nperm = 10000, hybrid p-value, kmax = 25, nmin = 200, min.width = 2, no undo).
 * Input: `nseries` NA-free series (one per chromosome [and sample]) concatenated in y / w (host,
 * weights > 0), off[nseries + 1] offsets, series_ids[nseries] (chromosome id mixed into the
 * permutation streams; NULL = 0..nseries-1).  Output: ends_out (capacity off[nseries]) receives, series
 * after series, the ascending exclusive end positions of the segments of each series (relative to the
 * series start); nseg_out[nseries] their counts.  The NA handling, the split of segments over long
 * NA runs and the weighted segment means of CBS.R:80-127 are host code (wisecondorx_b200/cbs.py). */
int wcx_cbs_segment(wcx_ctx* ctx, const double* y, const double* w, const int64_t* off, int32_t nseries,
                    const int32_t* series_ids, double alpha, int32_t nperm, uint32_t seed,
                    int32_t* ends_out, int32_t* nseg_out);
/* Sequential stopping boundary of the permutation tests: DNAcopy's segment() passes sbdry = getbdry(eta = 0.05, nperm,
 * max.ones = floor(nperm * alpha) + 1) to its change-point finder, which declares a test significant as soon as the
 * number of permutations done reaches sbdry[nrejc (nrejc + 1) / 2 + nrej] (nrej exceedances so far, nrejc tolerated).
 * The table is host int32 [n]; it stays in the context for the following wcx_cbs_segment calls.  n = 0 (the state of a
 * fresh context): every test runs all nperm permutations and counts. */
int wcx_cbs_set_boundary(wcx_ctx* ctx, const int32_t* sbdry, int32_t n);

/* Counters of the last wcx_cbs_segment: rounds, segments tested, permutation tests, edge t-tests
 * with permutations, permutations evaluated, kernel launches. */
int wcx_cbs_stats(wcx_ctx* ctx, int64_t* out6);

/* Test hook: raw tensor-core accumulators of one 128 x 256 tile of the kind::f16 MMA over the scaled f16 operands,
 * written to acc_out [128 * 256] (host). */
int wcx_debug_tc_tile_f16(wcx_ctx* ctx, int64_t row0, int64_t col0, float* acc_out);
/* Test hook: f16 operands (raw half bits) [n, k_pad_h], their norms [n], k_pad_h, {scale, scale^2}. */
int wcx_debug_prep_f16(wcx_ctx* ctx, uint16_t* xh_out, float* norm_out, int32_t* k_pad_out, double* scale_out);
/* Test hook: final length of the first `nslots` candidate lists of the last wcx_newref_topk call
 * (lists per row = out[3] of wcx_newref_stats x 2 for the tcgen05 kernel). */
int wcx_debug_list_counts(wcx_ctx* ctx, int32_t* cnt_out, int64_t nslots);
/* Test hook: fp32 operands of the CUDA-core cross-check kernel.  xc_out [n, k_pad] (host, may be NULL), norm_out [n] (host). */
int wcx_debug_prep(wcx_ctx* ctx, float* xc_out, float* norm_out, int32_t* k_pad_out);
/* Test hook, host only (no GPU needed): for S samples the NumPy pairwise-summation plan (int32 triples), the
 * leaf-major permutation of a row used by the re-rank gather (perm[p] = source column or -1 for zero padding) and
 * the leaf descriptors (offset, steps, tail terms, 0).  sizes3 = {permuted row length, leaves, plan triples}.
 * Output pointers may be NULL to query the sizes. */
int wcx_debug_leaf_layout(int32_t s, int32_t* perm_out, int32_t perm_cap, int32_t* desc_out, int32_t desc_cap,
                          int32_t* plan_out, int32_t plan_cap, int32_t* sizes3);

#ifdef __cplusplus
}
#endif
#endif /* WCX_B200_H */
