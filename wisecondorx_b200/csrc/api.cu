// C-ABI of libwcx_b200.so (see include/wcx_b200.h).  Host-side orchestration only: buffer
// management, work-item construction, stage timing; all arithmetic is in the kernels.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wcx_b200.h"
#include "wcx_common.cuh"

namespace wcx {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int launch_dist_topk_tc_debug(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                              void* tmap_storage, float* dbg_acc, cudaStream_t st);
}  // namespace wcx

using namespace wcx;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      set_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct wcx_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  // newref state
  const double* d_x = nullptr;  // owned (x_buf) or borrowed
  DevBuf x_buf, xc, norm, colsum, colcnt, cum_dev, items_dev, counter, cand_val, cand_idx, cand_cnt, cand_cut;
  DevBuf fail, fail_rows, plan_dev, scratch, idx_dev, dist_dev, xt, ids_dev, nr_dev, dbg;
  int64_t n = 0, n_pad = 0;
  int32_t s = 0, k_pad = 0, nchr = 0;
  std::vector<int64_t> per, cum;
  int32_t plan_len = 0;
  alignas(128) unsigned char tmap[128];
  bool loaded = false;
  // last topk
  int64_t last_rb = -1, last_re = -1;
  int32_t last_k = 0;
  int64_t stats[8] = {};
  double stage_ms[8] = {};
  int64_t launches = 0;
};

static PrepView prep_view(const wcx_ctx* c) {
  PrepView pv;
  pv.xc = c->xc.as<float>();
  pv.norm = c->norm.as<float>();
  pv.n = c->n;
  pv.n_pad = c->n_pad;
  pv.s = c->s;
  pv.k_pad = c->k_pad;
  return pv;
}

extern "C" {

int wcx_version(void) { return 100; }
const char* wcx_last_error(void) { return g_error.c_str(); }

int wcx_create(int32_t device, wcx_ctx** out) {
  if (!out) { set_error("wcx_create: out is NULL"); return 1; }
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_error("wcx_create: no CUDA device available (this library has no CPU fallback)");
    return 1;
  }
  if (device < 0 || device >= count) { set_error("wcx_create: bad device index"); return 1; }
  cudaDeviceProp prop;
  WCX_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("wcx_create: device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
              ", this library is built for sm_100a (B200) only");
    return 1;
  }
  WCX_CUDA_OK(cudaSetDevice(device));
  wcx_ctx* c = new wcx_ctx();
  c->device = device;
  WCX_CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  for (auto& e : c->ev) WCX_CUDA_OK(cudaEventCreate(&e));
  *out = c;
  return 0;
}

void wcx_destroy(wcx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->x_buf, &c->xc, &c->norm, &c->colsum, &c->colcnt, &c->cum_dev, &c->items_dev, &c->counter,
                    &c->cand_val, &c->cand_idx, &c->cand_cnt, &c->cand_cut, &c->fail, &c->fail_rows, &c->plan_dev,
                    &c->scratch, &c->idx_dev, &c->dist_dev, &c->xt, &c->ids_dev, &c->nr_dev, &c->dbg})
    b->release();
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int wcx_set_stream(wcx_ctx* c, void* cuda_stream) {
  if (!c) { set_error("null context"); return 1; }
  c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return 0;
}

int wcx_sync(wcx_ctx* c) {
  if (!c) { set_error("null context"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int wcx_newref_load(wcx_ctx* c, const double* x, int64_t n, int32_t s, const int64_t* per, const int64_t* cum,
                    int32_t nchr, int32_t x_on_device) {
  if (!c || !x || !per || !cum) { set_error("wcx_newref_load: null argument"); return 1; }
  if (n <= 0 || s <= 0 || nchr <= 0) { set_error("wcx_newref_load: empty matrix"); return 1; }
  if (n > 0x7fffff00ll) { set_error("wcx_newref_load: too many bins"); return 1; }
  if (cum[nchr - 1] != n) { set_error("wcx_newref_load: masked_bins_per_chr_cum[-1] != number of rows"); return 1; }
  for (int i = 0; i < nchr; i++) {
    int64_t prev = i ? cum[i - 1] : 0;
    if (per[i] < 0 || cum[i] - prev != per[i]) { set_error("wcx_newref_load: per/cum inconsistent"); return 1; }
  }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  c->loaded = false;
  c->n = n;
  c->s = s;
  c->nchr = nchr;
  c->per.assign(per, per + nchr);
  c->cum.assign(cum, cum + nchr);
  c->k_pad = (s + WCX_KBLOCK - 1) / WCX_KBLOCK * WCX_KBLOCK;
  c->n_pad = (n + 255) / 256 * 256 + 256;
  cudaStream_t st = c->stream;
  if (x_on_device) {
    c->d_x = x;
  } else {
    if (c->x_buf.ensure(sizeof(double) * (size_t)n * s)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->x_buf.p, x, sizeof(double) * (size_t)n * s, cudaMemcpyHostToDevice, st));
    c->d_x = c->x_buf.as<double>();
  }
  if (c->xc.ensure(sizeof(float) * (size_t)c->n_pad * c->k_pad)) return 1;
  if (c->norm.ensure(sizeof(float) * (size_t)c->n_pad)) return 1;
  if (c->colsum.ensure(sizeof(double) * s) || c->colcnt.ensure(sizeof(double) * s)) return 1;
  if (c->cum_dev.ensure(sizeof(int64_t) * nchr)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->cum_dev.p, cum, sizeof(int64_t) * nchr, cudaMemcpyHostToDevice, st));
  WCX_CUDA_OK(cudaEventRecord(c->ev[6], st));
  if (launch_col_stats(c->d_x, n, s, c->colsum.as<double>(), c->colcnt.as<double>(), st)) return 1;
  if (launch_center_round(c->d_x, n, s, c->colsum.as<double>(), c->colcnt.as<double>(), c->xc.as<float>(),
                          c->norm.as<float>(), c->n_pad, c->k_pad, st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[7], st));
  c->launches += 2;
  // NumPy pairwise-summation plan for length s
  std::vector<int32_t> plan(3 * 4096);
  int pl = build_sum_plan(s, plan.data(), (int32_t)plan.size());
  if (pl < 0) { set_error("wcx_newref_load: too many samples for the summation plan"); return 1; }
  c->plan_len = pl;
  if (c->plan_dev.ensure(sizeof(int32_t) * 3 * (size_t)pl)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->plan_dev.p, plan.data(), sizeof(int32_t) * 3 * (size_t)pl, cudaMemcpyHostToDevice, st));
  PrepView pv = prep_view(c);
  if (tc_encode_tensor_map(pv, c->tmap)) return 1;
  WCX_CUDA_OK(cudaStreamSynchronize(st));  // `plan` and caller's host buffers may go away
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]);
  c->stage_ms[4] = ms;
  c->loaded = true;
  c->last_rb = c->last_re = -1;
  return 0;
}

static void build_items(const wcx_ctx* c, int64_t rb, int64_t re, int tile_n, int nsplit, std::vector<WorkItem>& items) {
  const int64_t n = c->n;
  const int nct = (int)((n + tile_n - 1) / tile_n);
  const bool gon = c->nchr > 22;
  for (int ch = 0; ch < c->nchr; ch++) {
    if (gon && ch != 22 && ch != 23) continue;
    int64_t cs = c->cum[ch] - c->per[ch], ce = c->cum[ch];
    int64_t lo = std::max(cs, rb), hi = std::min(ce, re);
    for (int64_t r0 = lo; r0 < hi; r0 += WCX_TILE_M) {
      int nrows = (int)std::min<int64_t>(WCX_TILE_M, hi - r0);
      for (int q = 0; q < nsplit; q++) {
        WorkItem w;
        w.row0 = (int32_t)r0;
        w.nrows = nrows;
        w.chr_s = (int32_t)cs;
        w.chr_e = (int32_t)ce;
        w.ct_begin = (int)((int64_t)nct * q / nsplit);
        w.ct_end = (int)((int64_t)nct * (q + 1) / nsplit);
        w.slot0 = (int32_t)((r0 - rb) * nsplit + q);
        w.slot_stride = nsplit;
        items.push_back(w);
      }
    }
  }
}

int wcx_newref_topk(wcx_ctx* c, int64_t rb, int64_t re, int32_t k, int32_t kernel, int32_t* idx_out,
                    double* dist_out, int32_t out_on_device) {
  if (!c || !c->loaded) { set_error("wcx_newref_topk: call wcx_newref_load first"); return 1; }
  if (rb < 0 || re > c->n || rb > re) { set_error("wcx_newref_topk: bad row range"); return 1; }
  if (k <= 0 || k > 400) { set_error("wcx_newref_topk: ref_size must be in [1, 400]"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t rows = re - rb;
  std::memset(c->stats, 0, sizeof(c->stats));
  c->stage_ms[0] = c->stage_ms[1] = c->stage_ms[2] = 0.0;
  if (rows == 0) return 0;
  if (kernel == WCX_KERNEL_AUTO) kernel = WCX_KERNEL_TC;
  const int gon = c->nchr > 22 ? 1 : 0;
  if (c->idx_dev.ensure(sizeof(int32_t) * (size_t)rows * k) || c->dist_dev.ensure(sizeof(double) * (size_t)rows * k))
    return 1;
  if (c->fail.ensure(sizeof(int32_t) * (size_t)rows)) return 1;
  PrepView pv = prep_view(c);
  std::vector<int32_t> fail_list;

  if (kernel == WCX_KERNEL_EXACT) {
    // every real row through the brute-force path; placeholder rows via the rerank kernel's early exit
    for (int64_t r = rb; r < re; r++) {
      int ch = 0;
      while (ch < c->nchr && c->cum[ch] <= r) ch++;
      if (gon && ch != 22 && ch != 23) continue;
      fail_list.push_back((int32_t)(r - rb));
    }
    if (gon) {
      CandView cv0{nullptr, nullptr, nullptr, nullptr};
      // rerank kernel with nsplit = 0 would touch lists; fill placeholders on the host side instead
      std::vector<int32_t> hi((size_t)rows * k, 0);
      std::vector<double> hd((size_t)rows * k, 1.0);
      WCX_CUDA_OK(cudaMemcpyAsync(c->idx_dev.p, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, st));
      WCX_CUDA_OK(cudaMemcpyAsync(c->dist_dev.p, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice, st));
      WCX_CUDA_OK(cudaStreamSynchronize(st));
      (void)cv0;
    }
  } else {
    const int tile_n = kernel == WCX_KERNEL_SIMT ? WCX_TILE_N_SIMT : WCX_TILE_N_TC;
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device);
    // column splits: enough work items to balance a persistent grid when the part has few row tiles
    std::vector<WorkItem> items;
    build_items(c, rb, re, tile_n, 1, items);
    const int64_t row_tiles = (int64_t)items.size();
    int nsplit = 1;
    const int nct = (int)((c->n + tile_n - 1) / tile_n);
    while (nsplit < 4 && row_tiles * nsplit < 4 * dev_sms && nct / (nsplit * 2) >= 4) nsplit *= 2;
    if (nsplit > 1) {
      items.clear();
      build_items(c, rb, re, tile_n, nsplit, items);
    }
    const size_t slots = (size_t)rows * nsplit;
    if (c->cand_val.ensure(sizeof(float) * slots * WCX_CAND_CAP) || c->cand_idx.ensure(sizeof(int32_t) * slots * WCX_CAND_CAP) ||
        c->cand_cnt.ensure(sizeof(int32_t) * slots) || c->cand_cut.ensure(sizeof(float) * slots))
      return 1;
    if (c->items_dev.ensure(sizeof(WorkItem) * std::max<size_t>(items.size(), 1)) || c->counter.ensure(sizeof(int32_t))) return 1;
    if (!items.empty())
      WCX_CUDA_OK(cudaMemcpyAsync(c->items_dev.p, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, st));
    CandView cv{c->cand_val.as<float>(), c->cand_idx.as<int32_t>(), c->cand_cnt.as<int32_t>(), c->cand_cut.as<float>()};
    WCX_CUDA_OK(cudaMemsetAsync(c->cand_cnt.p, 0, sizeof(int32_t) * slots, st));
    WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
    if (kernel == WCX_KERNEL_SIMT) {
      if (launch_dist_topk_simt(pv, c->items_dev.as<WorkItem>(), (int)items.size(), cv, c->counter.as<int32_t>(), st)) return 1;
    } else {
      if (launch_dist_topk_tc(pv, c->items_dev.as<WorkItem>(), (int)items.size(), cv, c->counter.as<int32_t>(), c->tmap, st)) return 1;
    }
    WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
    if (launch_rerank(c->d_x, pv, cv, nsplit, c->cum_dev.as<int64_t>(), c->nchr, rb, re, k, gon, c->idx_dev.as<int32_t>(),
                      c->dist_dev.as<double>(), c->fail.as<int32_t>(), c->plan_dev.as<int32_t>(), c->plan_len, st))
      return 1;
    WCX_CUDA_OK(cudaEventRecord(c->ev[2], st));
    c->launches += 2;
    std::vector<int32_t> flags((size_t)rows);
    WCX_CUDA_OK(cudaMemcpyAsync(flags.data(), c->fail.p, sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < rows; i++)
      if (flags[(size_t)i]) fail_list.push_back((int32_t)i);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stage_ms[0] = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
    c->stage_ms[1] = ms;
    c->stats[0] = (int64_t)items.size();
    c->stats[3] = nsplit;
  }
  c->stats[4] = kernel;
  c->stats[1] = (int64_t)fail_list.size();
  if (!fail_list.empty()) {
    const size_t batch = std::max<size_t>(1, std::min<size_t>(fail_list.size(), (size_t)(1ull << 28) / (size_t)c->n));
    if (c->scratch.ensure(sizeof(double) * batch * (size_t)c->n) || c->fail_rows.ensure(sizeof(int32_t) * fail_list.size())) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->fail_rows.p, fail_list.data(), sizeof(int32_t) * fail_list.size(), cudaMemcpyHostToDevice, st));
    WCX_CUDA_OK(cudaEventRecord(c->ev[3], st));
    for (size_t off = 0; off < fail_list.size(); off += batch) {
      int nb = (int)std::min(batch, fail_list.size() - off);
      if (launch_exact_rows(c->d_x, c->n, c->s, c->cum_dev.as<int64_t>(), c->nchr, rb, c->fail_rows.as<int32_t>() + off, nb, k,
                            c->idx_dev.as<int32_t>(), c->dist_dev.as<double>(), c->scratch.as<double>(),
                            c->plan_dev.as<int32_t>(), c->plan_len, st))
        return 1;
      c->launches += 1;
    }
    WCX_CUDA_OK(cudaEventRecord(c->ev[4], st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]);
    c->stage_ms[2] = ms;
  }
  c->stats[2] = c->launches;
  c->last_rb = rb;
  c->last_re = re;
  c->last_k = k;
  const cudaMemcpyKind kind = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (idx_out) WCX_CUDA_OK(cudaMemcpyAsync(idx_out, c->idx_dev.p, sizeof(int32_t) * (size_t)rows * k, kind, st));
  if (dist_out) WCX_CUDA_OK(cudaMemcpyAsync(dist_out, c->dist_dev.p, sizeof(double) * (size_t)rows * k, kind, st));
  if (!out_on_device) WCX_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int wcx_newref_null_ratios(wcx_ctx* c, const int32_t* idx, int32_t idx_on_device, int64_t rb, int64_t re, int32_t k,
                           const int32_t* sample_ids, int32_t m, double* out, int32_t out_on_device) {
  if (!c || !c->loaded) { set_error("wcx_newref_null_ratios: call wcx_newref_load first"); return 1; }
  if (rb < 0 || re > c->n || rb > re || k <= 0 || m < 0 || !sample_ids || !out) { set_error("wcx_newref_null_ratios: bad argument"); return 1; }
  for (int i = 0; i < m; i++)
    if (sample_ids[i] < 0 || sample_ids[i] >= c->s) { set_error("wcx_newref_null_ratios: sample id out of range"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t rows = re - rb;
  if (rows == 0 || m == 0) return 0;
  const int32_t* d_idx = nullptr;
  if (!idx) {
    if (c->last_rb != rb || c->last_re != re || c->last_k != k) { set_error("wcx_newref_null_ratios: no matching device-resident indexes"); return 1; }
    d_idx = c->idx_dev.as<int32_t>();
  } else if (idx_on_device) {
    d_idx = idx;
  } else {
    if (c->idx_dev.ensure(sizeof(int32_t) * (size_t)rows * k)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->idx_dev.p, idx, sizeof(int32_t) * (size_t)rows * k, cudaMemcpyHostToDevice, st));
    d_idx = c->idx_dev.as<int32_t>();
    c->last_rb = rb; c->last_re = re; c->last_k = k;
  }
  if (c->xt.ensure(sizeof(double) * (size_t)m * c->n) || c->ids_dev.ensure(sizeof(int32_t) * m)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->ids_dev.p, sample_ids, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
  double* d_out = out;
  if (!out_on_device) {
    if (c->nr_dev.ensure(sizeof(double) * (size_t)rows * m)) return 1;
    d_out = c->nr_dev.as<double>();
  }
  WCX_CUDA_OK(cudaEventRecord(c->ev[5], st));
  if (launch_transpose_cols(c->d_x, c->n, c->s, c->ids_dev.as<int32_t>(), m, c->xt.as<double>(), st)) return 1;
  if (launch_null_ratios(c->xt.as<double>(), c->n, d_idx, rb, re, k, m, d_out, st)) return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[6], st));
  c->launches += 1 + (m + 7) / 8;
  if (!out_on_device) {
    WCX_CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)rows * m, cudaMemcpyDeviceToHost, st));
  }
  WCX_CUDA_OK(cudaStreamSynchronize(st));  // sample_ids staging + timing
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
  c->stage_ms[3] = ms;
  c->stats[2] = c->launches;
  return 0;
}

int wcx_get_reference(wcx_ctx* c, const double* x, int64_t n, int32_t s, const int64_t* per, const int64_t* cum, int32_t nchr,
                      int32_t k, int64_t rb, int64_t re, const int32_t* sample_ids, int32_t m, int32_t kernel,
                      int32_t* idx_out, double* dist_out, double* null_out) {
  if (wcx_newref_load(c, x, n, s, per, cum, nchr, 0)) return 1;
  if (wcx_newref_topk(c, rb, re, k, kernel, idx_out, dist_out, 0)) return 1;
  if (null_out && m > 0)
    if (wcx_newref_null_ratios(c, nullptr, 1, rb, re, k, sample_ids, m, null_out, 0)) return 1;
  return 0;
}

int wcx_newref_stats(wcx_ctx* c, int64_t* out8) {
  if (!c || !out8) { set_error("null argument"); return 1; }
  std::memcpy(out8, c->stats, sizeof(c->stats));
  return 0;
}

int wcx_newref_stage_ms(wcx_ctx* c, double* out8) {
  if (!c || !out8) { set_error("null argument"); return 1; }
  std::memcpy(out8, c->stage_ms, sizeof(c->stage_ms));
  return 0;
}

int wcx_debug_tc_tile(wcx_ctx* c, int64_t row0, int64_t col0, float* acc_out) {
  if (!c || !c->loaded || !acc_out) { set_error("wcx_debug_tc_tile: bad argument"); return 1; }
  if (row0 < 0 || row0 + WCX_TILE_M > c->n_pad || col0 < 0 || col0 % WCX_TILE_N_TC != 0 || col0 + WCX_TILE_N_TC > c->n_pad) {
    set_error("wcx_debug_tc_tile: tile out of range (col0 must be a multiple of 256)");
    return 1;
  }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  WorkItem w;
  w.row0 = (int32_t)row0; w.nrows = WCX_TILE_M; w.chr_s = -1; w.chr_e = -1;
  w.ct_begin = (int)(col0 / WCX_TILE_N_TC); w.ct_end = w.ct_begin + 1; w.slot0 = 0; w.slot_stride = 1;
  const size_t slots = WCX_TILE_M;
  if (c->cand_val.ensure(sizeof(float) * slots * WCX_CAND_CAP) || c->cand_idx.ensure(sizeof(int32_t) * slots * WCX_CAND_CAP) ||
      c->cand_cnt.ensure(sizeof(int32_t) * slots) || c->cand_cut.ensure(sizeof(float) * slots) ||
      c->items_dev.ensure(sizeof(WorkItem)) || c->dbg.ensure(sizeof(float) * WCX_TILE_M * WCX_TILE_N_TC))
    return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->items_dev.p, &w, sizeof(w), cudaMemcpyHostToDevice, st));
  CandView cv{c->cand_val.as<float>(), c->cand_idx.as<int32_t>(), c->cand_cnt.as<int32_t>(), c->cand_cut.as<float>()};
  PrepView pv = prep_view(c);
  if (launch_dist_topk_tc_debug(pv, c->items_dev.as<WorkItem>(), 1, cv, c->tmap, c->dbg.as<float>(), st)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(acc_out, c->dbg.p, sizeof(float) * WCX_TILE_M * WCX_TILE_N_TC, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  c->last_rb = c->last_re = -1;
  return 0;
}

int wcx_debug_prep(wcx_ctx* c, float* xc_out, float* norm_out, int32_t* k_pad_out) {
  if (!c || !c->loaded) { set_error("wcx_debug_prep: not loaded"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  if (k_pad_out) *k_pad_out = c->k_pad;
  if (xc_out) WCX_CUDA_OK(cudaMemcpyAsync(xc_out, c->xc.p, sizeof(float) * (size_t)c->n * c->k_pad, cudaMemcpyDeviceToHost, c->stream));
  if (norm_out) WCX_CUDA_OK(cudaMemcpyAsync(norm_out, c->norm.p, sizeof(float) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"
