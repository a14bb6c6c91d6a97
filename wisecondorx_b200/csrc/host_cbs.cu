// Host side of the segmentation step for a batch of samples (no device work; host threads): what include/CBS.R does
// around DNAcopy::segment, i.e. around wcx_cbs_segment here.
//
//   before (CBS.R:30-63)   ratio == 0 -> NA, weight == 0 -> 1 (1^-99 in R), every chromosome of every sample becomes an
//                          NA-free (ratio, weight) series; chromosomes without data are dropped
//   after  (CBS.R:80-129)  a segment is cut at every run of NA bins longer than int(2e6 / binsize) that lies inside it
//                          (:86-101), pieces of at most one bin are dropped (:103), the ratio of a piece is the weighted
//                          mean of its non-NA bins (:122-127), coordinates are 0-based half-open (:129)
//
// wcx_cbs_pack_count / wcx_cbs_pack build the two vectors of the ONE device call of a batch (all series back to back) and
// keep, for every NA-free entry, its bin position in the sample: a run of NA bins inside a segment is then simply a jump
// in the positions of two consecutive entries (a segment starts and ends on a non-NA bin and never leaves its
// chromosome), so wcx_cbs_unpack needs neither the ratios of the NA bins nor a per-segment scan of the chromosome.
// The weighted means are sums of the products y * w and of w over the entries of a piece in NumPy's pairwise order
// (host_sums.h) -- the value np.sum returns for the arrays CBS.R's R code would sum; oracle/cbs_oracle.py (test
// infrastructure) restates CBS.R with NumPy and tests/test_cbs_oracle.py, tests/test_predict_host_cpu.py compare the two
// bit for bit on random segmentations.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "host_sums.h"
#include "wcx_common.cuh"

namespace {

struct Rows {
  const double* const* r;
  const double* const* w;
  const int64_t* offs;     // the chromosome offsets of all samples back to back
  const int64_t* offs_at;  // [samples + 1]: sample s has offs[offs_at[s]] ... offs[offs_at[s + 1] - 1], i.e. offs_at[s + 1] - offs_at[s] - 1 chromosomes
};

inline int64_t series_of(const Rows& rows, int32_t s) { return rows.offs_at[s] - s; }  // first (sample, chromosome) slot of sample s

}  // namespace

extern "C" int wcx_cbs_pack_count(const double* const* r_rows, const int64_t* offs, const int64_t* offs_at, int32_t samples,
                                  int64_t* counts, int32_t threads) {
  if (samples < 0) { wcx::set_error("wcx_cbs_pack_count: negative size"); return 1; }
  if (samples == 0) return 0;
  if (!r_rows || !offs || !offs_at || !counts) { wcx::set_error("wcx_cbs_pack_count: null argument"); return 1; }
  for (int32_t s = 0; s < samples; ++s) {
    if (offs_at[s + 1] - offs_at[s] < 1 || !r_rows[s]) { wcx::set_error("wcx_cbs_pack_count: sample without offsets or data"); return 1; }
    for (int64_t i = offs_at[s]; i + 1 < offs_at[s + 1]; ++i)
      if (offs[i + 1] < offs[i]) { wcx::set_error("wcx_cbs_pack_count: offsets not ascending"); return 1; }
  }
  Rows rows{r_rows, nullptr, offs, offs_at};
  wcx::host_parallel(samples, 1, threads, [=](int64_t s0, int64_t s1) {
    for (int64_t s = s0; s < s1; ++s) {
      const double* r = rows.r[s];
      const int64_t* o = rows.offs + rows.offs_at[s];
      const int64_t nchr = rows.offs_at[s + 1] - rows.offs_at[s] - 1;
      int64_t* cnt = counts + series_of(rows, (int32_t)s);
      for (int64_t c = 0; c < nchr; ++c) {
        int64_t k = 0;
        for (int64_t b = o[c]; b < o[c + 1]; ++b) k += r[b] != 0.0;  // CBS.R:41 (NaN != 0: kept, like R's `== 0`)
        cnt[c] = k;
      }
    }
  });
  return 0;
}

extern "C" int wcx_cbs_pack(const double* const* r_rows, const double* const* w_rows, const int64_t* offs,
                            const int64_t* offs_at, int32_t samples, const int64_t* at, int64_t na_thresh, double* y,
                            double* w, int32_t* pos, int64_t* long_gaps, int32_t threads) {
  if (samples < 0) { wcx::set_error("wcx_cbs_pack: negative size"); return 1; }
  if (samples == 0) return 0;
  if (!r_rows || !w_rows || !offs || !offs_at || !at || !y || !w || !pos || !long_gaps) {
    wcx::set_error("wcx_cbs_pack: null argument");
    return 1;
  }
  Rows rows{r_rows, w_rows, offs, offs_at};
  wcx::host_parallel(samples, 1, threads, [=](int64_t s0, int64_t s1) {
    for (int64_t s = s0; s < s1; ++s) {
      const double* r = rows.r[s];
      const double* ww = rows.w[s];
      const int64_t* o = rows.offs + rows.offs_at[s];
      const int64_t nchr = rows.offs_at[s + 1] - rows.offs_at[s] - 1;
      int64_t* gaps = long_gaps + series_of(rows, (int32_t)s);
      int64_t k = at[s];
      for (int64_t c = 0; c < nchr; ++c) {
        int64_t last = -1, g = 0;
        for (int64_t b = o[c]; b < o[c + 1]; ++b) {
          if (r[b] == 0.0) continue;  // CBS.R:41
          y[k] = r[b];
          w[k] = ww[b] == 0.0 ? 1.0 : ww[b];  // CBS.R:42 -- 1^-99 is 1 in R
          pos[k] = (int32_t)b;
          if (last >= 0 && b - last - 1 > na_thresh) ++g;
          last = b;
          ++k;
        }
        gaps[c] = g;  // upper bound of the extra pieces wcx_cbs_unpack can cut out of this chromosome's segments
      }
    }
  });
  return 0;
}

extern "C" int wcx_cbs_unpack(const int32_t* pos, const double* y, const double* w, const int64_t* off, int32_t series,
                              const int32_t* ends, const int32_t* nseg, const int64_t* chr_start, const int64_t* slot,
                              int64_t na_thresh, int32_t* out_series, int64_t* out_s, int64_t* out_e, double* out_r,
                              int32_t threads) {
  if (series < 0) { wcx::set_error("wcx_cbs_unpack: negative size"); return 1; }
  if (series == 0) return 0;
  if (!pos || !y || !w || !off || !ends || !nseg || !chr_start || !slot || !out_series || !out_s || !out_e || !out_r) {
    wcx::set_error("wcx_cbs_unpack: null argument");
    return 1;
  }
  std::vector<int64_t> first_end((size_t)series + 1, 0);  // position of every series' first segment end in `ends`
  for (int32_t i = 0; i < series; ++i) {
    if (nseg[i] < 0 || off[i + 1] < off[i] || slot[i + 1] < slot[i]) { wcx::set_error("wcx_cbs_unpack: malformed offsets"); return 1; }
    first_end[i + 1] = first_end[i] + nseg[i];
  }
  std::atomic<int> bad{0};
  const int64_t* fe = first_end.data();
  wcx::host_parallel(series, 8, threads, [=, &bad](int64_t i0, int64_t i1) {
    std::vector<double> prod;
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t a0 = chr_start[i], n = off[i + 1] - off[i];
      const int32_t* p = pos + off[i];
      const double* yy = y + off[i];
      const double* wv = w + off[i];
      int64_t o = slot[i];
      const int64_t o_end = slot[i + 1];
      auto emit = [&](int64_t s1, int64_t e1, int64_t ka, int64_t kb) {  // 1-based first / last bin; entries [ka, kb)
        if (e1 - s1 <= 0) return;                                       // CBS.R:103
        if (o >= o_end) { bad = 1; return; }
        double r = std::numeric_limits<double>::quiet_NaN();
        if (kb > ka) {  // CBS.R:122-127
          prod.resize((size_t)(kb - ka));
          for (int64_t k = ka; k < kb; ++k) prod[(size_t)(k - ka)] = yy[k] * wv[k];
          r = wcx::numpy_pairwise_sum(prod.data(), kb - ka) / wcx::numpy_pairwise_sum(wv + ka, kb - ka);
        }
        out_series[o] = (int32_t)i;
        out_s[o] = s1 - 1 - a0;  // CBS.R:129, predict_tools.py:266-275
        out_e[o] = e1 - a0;
        out_r[o] = r;
        ++o;
      };
      int64_t a = 0;
      for (int64_t j = 0; j < nseg[i]; ++j) {
        const int64_t b = ends[fe[i] + j];
        if (b <= a || b > n) { bad = 1; break; }
        // DNAcopy's loc.start / loc.end of the segment: the bins of its first and last entry (1-based below)
        int64_t s1 = (int64_t)p[a] + 1, ka = a;
        for (int64_t k = a; k + 1 < b; ++k) {
          if ((int64_t)p[k + 1] - p[k] - 1 > na_thresh) {  // CBS.R:86-101: a long run of NA bins between two entries
            emit(s1, (int64_t)p[k] + 1, ka, k + 1);        // ... up to the bin before the run
            s1 = p[k + 1];                                 // CBS.R's end.pos: the last NA bin (1-based) opens the next piece
            ka = k + 1;
          }
        }
        emit(s1, (int64_t)p[b - 1] + 1, ka, b);
        a = b;
      }
      for (; o < o_end; ++o) out_series[o] = -1;  // unused slots
    }
  });
  if (bad.load()) { wcx::set_error("wcx_cbs_unpack: segment ends out of order or more pieces than slots"); return 1; }
  return 0;
}
