// Host-side result assembly of `predict` for a batch of samples (no device work; host threads).
//
// Reference code replaced, per sample: the stacking of the autosomal and gonosomal `normalize` results and the z-score
// shift (main.py:242-246), get_post_processed_result (predict_control.py:49-63: bins with fewer than minrefbins
// reference bins are blanked), inflate_results (predict_tools.py:163-170: back to the unmasked bin axis) and log_trans
// (predict_tools.py:180-193: non-finite log ratios blank r / z / w, the median log ratio is subtracted from every
// non-zero entry).  In the reference these are ~100 NumPy calls and Python loops per sample; in a batch of 96 samples
// at 15 kb they cost more wall-clock than every kernel of the batch together (predict).  Here it is
// one streaming pass per sample over the unmasked bin axis: the log2 itself is taken by the caller with NumPy (so the
// values are NumPy's on this machine, whichever SIMD log it dispatches to), everything else is exact elementwise
// arithmetic (one subtraction per value) and therefore bit-identical with the per-key, per-chromosome NumPy form of the
// reference (tests/test_predict_host_cpu.py compares this function with the pinned restatement of it on the CPU).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "wcx_common.cuh"

namespace {

struct AssembleArgs {
  const double* lr_a;           // [samples, n_aut]
  const double *z_a, *nref_a;  // [*, n_aut], sample s in row aut_row[s]
  const int32_t* aut_row;
  const double *lr_g, *z_g, *nref_g;  // [samples, n_gon]
  const double *weights, *m_lr, *m_z;
  int64_t n_aut, n_gon, bins;
  const uint8_t* mask;
  double minrefbins;
  double *out_r, *out_z, *out_w;
  int32_t* out_inflate;
};

void assemble_sample(const AssembleArgs& a, int32_t s) {
  const double* la = a.lr_a + (int64_t)s * a.n_aut;
  const double* za = a.z_a + (int64_t)a.aut_row[s] * a.n_aut;
  const double* na = a.nref_a + (int64_t)a.aut_row[s] * a.n_aut;
  const double* lg = a.lr_g + (int64_t)s * a.n_gon;
  const double* zg = a.z_g + (int64_t)s * a.n_gon;
  const double* ng = a.nref_g + (int64_t)s * a.n_gon;
  double* r = a.out_r + (int64_t)s * a.bins;
  double* z = a.out_z + (int64_t)s * a.bins;
  double* w = a.out_w + (int64_t)s * a.bins;
  int32_t* inf = a.out_inflate + (int64_t)s * a.bins;
  const double m_lr = a.m_lr[s], m_z = a.m_z[s];
  int64_t j = 0;  // position among the kept bins = position in the stacked autosomal + gonosomal results
  for (int64_t b = 0; b < a.bins; ++b) {
    if (!a.mask[b]) {
      r[b] = 0.0; z[b] = 0.0; w[b] = 0.0; inf[b] = -1;
      continue;
    }
    const bool aut = j < a.n_aut;
    const double lr = aut ? la[j] : lg[j - a.n_aut];
    const double zz = aut ? za[j] : zg[j - a.n_aut];
    const double nr = aut ? na[j] : ng[j - a.n_aut];
    const bool low = nr < a.minrefbins;          // predict_control.py:50-51 (a blanked ratio is 0, its logarithm -inf)
    const bool bad = low || !std::isfinite(lr);  // predict_tools.py:183-187
    r[b] = bad ? 0.0 : (lr != 0.0 ? lr - m_lr : 0.0);  // predict_tools.py:189-191
    z[b] = bad ? 0.0 : zz - m_z;                        // main.py:244
    w[b] = bad ? 0.0 : a.weights[j];
    inf[b] = low ? -1 : (int32_t)j;
    ++j;
  }
}

}  // namespace

extern "C" int wcx_predict_assemble(const double* lr_aut, const double* z_aut, const double* nref_aut, int64_t n_aut,
                                    const int32_t* aut_row, const double* lr_gon, const double* z_gon,
                                    const double* nref_gon, int64_t n_gon, const double* weights, const double* m_lr,
                                    const double* m_z, int32_t samples, double minrefbins, const uint8_t* mask,
                                    int64_t bins, double* out_r, double* out_z, double* out_w, int32_t* out_inflate,
                                    int32_t threads) {
  if (samples < 0 || n_aut < 0 || n_gon < 0 || bins < 0) { wcx::set_error("wcx_predict_assemble: negative size"); return 1; }
  if (samples == 0) return 0;
  if (!lr_aut || !z_aut || !nref_aut || !aut_row || !weights || !m_lr || !m_z || !mask || !out_r || !out_z || !out_w ||
      !out_inflate || (n_gon > 0 && (!lr_gon || !z_gon || !nref_gon))) {
    wcx::set_error("wcx_predict_assemble: null argument");
    return 1;
  }
  int64_t kept = 0;
  for (int64_t b = 0; b < bins; ++b) kept += mask[b] != 0;
  // the reference's inflate loop hands results[j] to the j-th kept bin: surplus results are ignored, missing ones raise
  if (n_aut + n_gon < kept) {
    wcx::set_error("wcx_predict_assemble: fewer results (" + std::to_string(n_aut + n_gon) + ") than kept bins (" +
                   std::to_string(kept) + ")");
    return 2;
  }
  for (int32_t s = 0; s < samples; ++s)
    if (aut_row[s] < 0) { wcx::set_error("wcx_predict_assemble: negative row index"); return 1; }
  AssembleArgs a{lr_aut, z_aut, nref_aut, aut_row, lr_gon, z_gon, nref_gon, weights, m_lr, m_z, n_aut, n_gon, bins, mask,
                 minrefbins, out_r, out_z, out_w, out_inflate};
  const int32_t nt = std::max(1, std::min(threads, samples));
  if (nt == 1) {
    for (int32_t s = 0; s < samples; ++s) assemble_sample(a, s);
    return 0;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (int32_t t = 0; t < nt; ++t)
    pool.emplace_back([&a, t, nt, samples]() {
      for (int32_t s = t; s < samples; s += nt) assemble_sample(a, s);
    });
  for (auto& th : pool) th.join();
  return 0;
}
