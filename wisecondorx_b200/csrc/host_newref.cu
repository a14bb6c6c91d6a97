// Host-side preparation of `newref` (no device work; host threads): the stacked count matrix of all samples and the
// per-bin coverage sums behind get_mask.
//
// Reference code replaced: the column-by-column fill of `all_data` (newref_tools.py:81-92, :114-122: one strided
// column per (chromosome, sample)) and `np.sum(all_data / sum_per_sample, 1)` (:94-97).  At 15 kb / 500 samples the
// matrix has 2.06e5 x 500 entries; the command line builds it once (int32, 0.41 GB) for the three masks (all samples,
// females, males) and the three passes, and these two steps were the largest host items left in `newref` after the GPU
// passes went to 0.1 s each.
//
// wcx_host_bin_sums returns exactly what NumPy returns for np.sum(counts[:, cols] / col_sum, 1): one division per
// element, and the additions in the order of NumPy's pairwise summation of a contiguous float64 row
// (host_sums.h).
// tests/test_host_pin.py compares the resulting masks with the live reference's get_mask.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host_sums.h"
#include "wcx_common.cuh"


extern "C" int wcx_host_stack_counts(const int32_t* const* columns, const int64_t* lens, int32_t samples, int32_t nchr,
                                     const int64_t* offs, int32_t* out, int32_t threads) {
  if (samples < 0 || nchr < 0) { wcx::set_error("wcx_host_stack_counts: negative size"); return 1; }
  if (samples == 0 || nchr == 0) return 0;
  if (!columns || !lens || !offs || !out) { wcx::set_error("wcx_host_stack_counts: null argument"); return 1; }
  for (int32_t c = 0; c < nchr; ++c) {
    if (offs[c + 1] < offs[c]) { wcx::set_error("wcx_host_stack_counts: offsets not ascending"); return 1; }
    for (int32_t s = 0; s < samples; ++s) {
      const int64_t len = lens[(int64_t)s * nchr + c];
      if (len < 0 || len > offs[c + 1] - offs[c] || (len > 0 && !columns[(int64_t)s * nchr + c])) {
        wcx::set_error("wcx_host_stack_counts: a sample is longer than its chromosome's rows (or has no data pointer)");
        return 1;
      }
    }
  }
  const int64_t total = offs[nchr] - offs[0];
  // rows [a, b) of the result, 16 samples at a time: 16 sequential input streams, one 64-byte line per output row piece
  wcx::host_parallel(total, 4096, threads, [=](int64_t a, int64_t b) {
    int32_t c = 0;
    while (c + 1 < nchr && offs[c + 1] - offs[0] <= a) ++c;
    for (int64_t row = a; row < b;) {
      while (c + 1 < nchr && offs[c + 1] - offs[0] <= row) ++c;
      const int64_t c0 = offs[c] - offs[0];
      const int64_t end = std::min(b, offs[c + 1] - offs[0]);
      for (int32_t s0 = 0; s0 < samples; s0 += 16) {
        const int32_t ns = std::min(16, samples - s0);
        const int32_t* src[16];
        int64_t len[16];
        for (int32_t j = 0; j < ns; ++j) {
          src[j] = columns[(int64_t)(s0 + j) * nchr + c];
          len[j] = lens[(int64_t)(s0 + j) * nchr + c];
        }
        for (int64_t r = row; r < end; ++r) {
          int32_t* o = out + r * samples + s0;
          const int64_t k = r - c0;
          for (int32_t j = 0; j < ns; ++j) o[j] = k < len[j] ? src[j][k] : 0;  // zero padded to the longest sample
        }
      }
      row = end;
    }
  });
  return 0;
}

extern "C" int wcx_host_bin_sums(const int32_t* counts, int64_t bins, int32_t samples, const int32_t* cols, int32_t ncols,
                                 const double* col_sum, double* out, int32_t threads) {
  if (bins < 0 || samples < 0 || ncols < 0) { wcx::set_error("wcx_host_bin_sums: negative size"); return 1; }
  if (bins == 0) return 0;
  if (!counts || !col_sum || !out) { wcx::set_error("wcx_host_bin_sums: null argument"); return 1; }
  const int32_t n = cols ? ncols : samples;
  if (cols)
    for (int32_t j = 0; j < ncols; ++j)
      if (cols[j] < 0 || cols[j] >= samples) { wcx::set_error("wcx_host_bin_sums: column index out of range"); return 1; }
  wcx::host_parallel(bins, 2048, threads, [=](int64_t a, int64_t b) {
    std::vector<double> row((size_t)std::max(n, 1));
    for (int64_t r = a; r < b; ++r) {
      const int32_t* src = counts + r * samples;
      if (cols)
        for (int32_t j = 0; j < n; ++j) row[j] = (double)src[cols[j]] / col_sum[j];
      else
        for (int32_t j = 0; j < n; ++j) row[j] = (double)src[j] / col_sum[j];
      out[r] = wcx::numpy_pairwise_sum(row.data(), n);
    }
  });
  return 0;
}
