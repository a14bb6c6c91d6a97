// C-ABI of libwcx_b200.so (see include/wcx_b200.h).  Host-side orchestration only: buffer
// management, work-item construction, stage timing; all arithmetic is in the kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wcx_b200.h"
#include "wcx_common.cuh"
#include "predict.cuh"
#include "cbs.cuh"
#include "newref_pca.cuh"

namespace wcx {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int launch_dist_topk_tc_debug(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv,
                              void* tmap_storage, float* dbg_acc, cudaStream_t st);
int launch_dist_topk_tc_pair(const PrepView& pv, const WorkItem* items, int32_t nitems, CandView cv, void* tmap_storage,
                             cudaStream_t st);
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

struct RefSet {
  DevBuf idx, dist, cum, comps, mean, mask_pos;
  DevBuf gl;                 // gather lists of the last (cutoff, ct): predict.cu gather_list_kernel
  double gl_cutoff = 0.0;
  int64_t gl_ct = -1;
  bool gl_valid = false;
  int64_t n = 0, bins_total = 0;
  int32_t k = 0, nchr = 0, ncomp = 0;
  std::vector<int64_t> cum_h;
  bool loaded = false;
};

struct wcx_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H of finished outputs overlapped with the next stage
  cudaStream_t null_stream = nullptr;  // high priority: null ratios of finished row blocks next to the re-rank of the next
  cudaStream_t tail_stream = nullptr;  // high priority: the sweep's partial last round (region B) next to the re-rank of region A
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev_blk[16] = {};         // per row block: re-rank done / null ratios done
  cudaEvent_t ev_null[16] = {};
  cudaEvent_t ev_tail[2] = {};         // timing of the exposed null-ratio tail
  cudaEvent_t ev_tailsweep[2] = {};    // the sweep's partial last round (region B) on the side stream
  cudaEvent_t ev[8] = {};
  // newref state
  const double* d_x = nullptr;  // owned (x_buf) or borrowed
  DevBuf x_buf, xc, norm, xh, norm_h, normres_h, rho_dev, scale_dev, absmax, colsum, colcnt, cum_dev, items_dev, counter, cand_ent, cand_cnt, cand_cut;
  DevBuf fail, fail_rows, plan_dev, leaves_dev, scratch, idx_dev, dist_dev, xt, ids_dev, nr_dev, dbg, diag;
  DevBuf xp, perm_dev, leafdesc_dev;  // leaf-major copy of X for the re-rank gather (rerank.cu)
  int32_t sp = 0, leaf_n = 0;
  int64_t n = 0, n_pad = 0;
  int32_t s = 0, k_pad = 0, k_pad_h = 0, nchr = 0;
  std::vector<int64_t> per, cum;
  int32_t plan_len = 0;
  int32_t plan_leaves = 0, plan_depth = 0;
  alignas(128) unsigned char tmap[128];
  alignas(128) unsigned char tmap_h[128];  // f16 operands
  bool loaded = false;
  bool tf32_ready = false;
  // last topk
  int64_t last_rb = -1, last_re = -1;
  int32_t last_k = 0;
  int64_t stats[8] = {};
  double stage_ms[8] = {};
  int64_t launches = 0;
  int64_t exact_evals = 0, gathered = 0;
  // predict state
  RefSet ref[3];
  DevBuf p_partial, p_totals, p_tdots, p_state, p_raw, p_x, p_copy_a, p_copy_b, p_z, p_r, p_n, p_mlr, p_mz, p_w;
  DevBuf z_nr, z_pos, z_r, z_w, z_se, z_segr, z_out, z_partial, p_radix;
  double predict_ms[8] = {};
  int64_t z_nr_rows = -1;  // null ratios resident for wcx_segment_zscore(nr = NULL)
  int32_t z_nr_m = 0;
  CbsWorkspace* cbs = nullptr;
  // newref prep state
  DevBuf q_counts, q_pos, q_colsum, q_x, q_mean, q_partial, q_gram, q_u, q_sigma, q_comps, q_corr, q_med, q_d, q_work;
  int64_t q_n = 0;
  int32_t q_s = 0;
  int64_t q_counts_rows = -1;  // count matrix resident for wcx_newref_normalize_and_mask(counts = NULL)
  int32_t q_counts_s = 0;
  const double* q_xptr = nullptr;
  double prep_ms[4] = {};
  CbsStats cbs_stats = {};
};

static constexpr float WCX_F16_ABS_ERR = 6.103515625e-05f;  // 2^-14, see prep_view_h

static PrepView prep_view(const wcx_ctx* c) {
  PrepView pv;
  pv.xc = c->xc.as<float>();
  pv.norm = c->norm.as<float>();
  pv.n = c->n;
  pv.n_pad = c->n_pad;
  pv.s = c->s;
  pv.k_pad = c->k_pad;
  pv.scale = nullptr;
  pv.abs_err = 0.f;
  pv.f16 = 0;
  pv.normres = nullptr;
  pv.rho_max = nullptr;
  return pv;
}

// the f16 operand set: scaled values, norms in scaled units
static PrepView prep_view_h(const wcx_ctx* c) {
  PrepView pv = prep_view(c);
  pv.xc = reinterpret_cast<const float*>(c->xh.p);
  pv.norm = c->norm_h.as<float>();
  pv.k_pad = c->k_pad_h;
  pv.scale = c->scale_dev.as<double>();
  // values below the smallest normal f16 (2^-14) may be rounded to a subnormal (spacing 2^-24) or, on a path that
  // flushes them, to zero: 2^-14 covers both
  pv.abs_err = WCX_F16_ABS_ERR;
  pv.f16 = 1;
  pv.normres = c->normres_h.as<float2>();
  pv.rho_max = c->rho_dev.as<float>();
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
  WCX_CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  WCX_CUDA_OK(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  {
    int lo = 0, hi = 0;
    WCX_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    WCX_CUDA_OK(cudaStreamCreateWithPriority(&c->null_stream, cudaStreamNonBlocking, hi));
    WCX_CUDA_OK(cudaStreamCreateWithPriority(&c->tail_stream, cudaStreamNonBlocking, hi));
  }
  for (auto& e : c->ev_blk) WCX_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : c->ev_null) WCX_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : c->ev_tail) WCX_CUDA_OK(cudaEventCreate(&e));
  for (auto& e : c->ev_tailsweep) WCX_CUDA_OK(cudaEventCreate(&e));
  for (auto& e : c->ev) WCX_CUDA_OK(cudaEventCreate(&e));
  *out = c;
  return 0;
}

void wcx_destroy(wcx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->x_buf, &c->xc, &c->norm, &c->xh, &c->norm_h, &c->normres_h, &c->rho_dev, &c->scale_dev, &c->absmax, &c->colsum, &c->colcnt, &c->cum_dev, &c->items_dev, &c->counter,
                    &c->cand_ent, &c->cand_cnt, &c->cand_cut, &c->fail, &c->fail_rows, &c->plan_dev, &c->leaves_dev,
                    &c->scratch, &c->idx_dev, &c->dist_dev, &c->xt, &c->ids_dev, &c->nr_dev, &c->dbg, &c->diag, &c->xp, &c->perm_dev, &c->leafdesc_dev, &c->p_partial,
                    &c->p_totals, &c->p_tdots, &c->p_state, &c->p_raw, &c->p_x, &c->p_copy_a, &c->p_copy_b, &c->p_z, &c->p_r,
                    &c->p_n, &c->p_mlr, &c->p_mz, &c->p_w, &c->z_nr, &c->z_pos, &c->z_r, &c->z_w, &c->z_se, &c->z_segr, &c->z_out, &c->z_partial, &c->p_radix})
    b->release();
  for (DevBuf* b : {&c->q_counts, &c->q_pos, &c->q_colsum, &c->q_x, &c->q_mean, &c->q_partial, &c->q_gram, &c->q_u, &c->q_sigma,
                    &c->q_comps, &c->q_corr, &c->q_med, &c->q_d, &c->q_work})
    b->release();
  for (auto& r : c->ref)
    for (DevBuf* b : {&r.idx, &r.dist, &r.cum, &r.comps, &r.mean, &r.mask_pos, &r.gl}) b->release();
  if (c->cbs) cbs_workspace_destroy(c->cbs);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_blk) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_null) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_tail) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_tailsweep) if (e) cudaEventDestroy(e);
  if (c->null_stream) cudaStreamDestroy(c->null_stream);
  if (c->tail_stream) cudaStreamDestroy(c->tail_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int wcx_set_stream(wcx_ctx* c, void* cuda_stream) {
  if (!c) { set_error("null context"); return 1; }
  c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return 0;
}

int wcx_host_alloc(uint64_t bytes, void** out) {
  if (!out) { set_error("wcx_host_alloc: out is NULL"); return 1; }
  *out = nullptr;
  WCX_CUDA_OK(cudaHostAlloc(out, bytes ? (size_t)bytes : 1, cudaHostAllocPortable));
  return 0;
}

int wcx_host_free(void* p) {
  if (p) WCX_CUDA_OK(cudaFreeHost(p));
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
  if (x_on_device == 2) {
    // the corrected matrix left in the context by wcx_pca_apply(corrected_out = NULL)
    if (!c || !c->q_corr.p || c->q_n != n || c->q_s != s) { set_error("wcx_newref_load: no device-resident corrected matrix of that shape"); return 1; }
    x = c->q_corr.as<double>();
  }
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
  c->k_pad_h = (s + 2 * WCX_KBLOCK - 1) / (2 * WCX_KBLOCK) * (2 * WCX_KBLOCK);
  c->n_pad = (n + 255) / 256 * 256 + 256;
  cudaStream_t st = c->stream;
  if (x_on_device == 3) {
    // a device pointer that may live on ANOTHER GPU of the box (multi-GPU newref: the corrected matrix is prepared on
    // one device): one copy over NVLink / PCIe into this context's own buffer (unified addressing resolves the source)
    if (c->x_buf.ensure(sizeof(double) * (size_t)n * s)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->x_buf.p, x, sizeof(double) * (size_t)n * s, cudaMemcpyDefault, st));
    c->d_x = c->x_buf.as<double>();
  } else if (x_on_device) {
    c->d_x = x;
  } else {
    if (c->x_buf.ensure(sizeof(double) * (size_t)n * s)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->x_buf.p, x, sizeof(double) * (size_t)n * s, cudaMemcpyHostToDevice, st));
    c->d_x = c->x_buf.as<double>();
  }
  if (c->xh.ensure(2 * (size_t)c->n_pad * c->k_pad_h) || c->norm_h.ensure(sizeof(float) * (size_t)c->n_pad)) return 1;
  if (c->normres_h.ensure(sizeof(float2) * (size_t)c->n_pad) || c->rho_dev.ensure(sizeof(float))) return 1;
  if (c->scale_dev.ensure(2 * sizeof(double)) || c->absmax.ensure(sizeof(unsigned long long))) return 1;
  if (c->colsum.ensure(sizeof(double) * s) || c->colcnt.ensure(sizeof(double) * s)) return 1;
  if (c->cum_dev.ensure(sizeof(int64_t) * nchr)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->cum_dev.p, cum, sizeof(int64_t) * nchr, cudaMemcpyHostToDevice, st));
  WCX_CUDA_OK(cudaEventRecord(c->ev[6], st));
  if (launch_col_stats(c->d_x, n, s, c->colsum.as<double>(), c->colcnt.as<double>(), c->absmax.as<unsigned long long>(), st)) return 1;
  c->tf32_ready = false;  // the fp32 / tf32 operand set is only built when a tf32 or CUDA-core sweep asks for it (ensure_tf32)
  if (launch_center_round_f16(c->d_x, n, s, c->colsum.as<double>(), c->colcnt.as<double>(), c->absmax.as<unsigned long long>(),
                              c->scale_dev.as<double>(), c->xh.p, c->norm_h.as<float>(), c->normres_h.as<float2>(),
                              c->rho_dev.as<float>(), 2.0 * sqrt((double)c->k_pad_h) * (double)WCX_F16_ABS_ERR, c->n_pad, c->k_pad_h, st))
    return 1;
  c->launches += 3;
  // NumPy pairwise-summation plan for length s
  std::vector<int32_t> plan(3 * 4096);
  int pl = build_sum_plan(s, plan.data(), (int32_t)plan.size());
  if (pl < 0) { set_error("wcx_newref_load: too many samples for the summation plan"); return 1; }
  c->plan_len = pl;
  {
    std::vector<int32_t> leaves;
    c->plan_leaves = plan_to_leaves(plan.data(), pl, leaves, &c->plan_depth);
    if (c->plan_leaves < 0) { set_error("wcx_newref_load: malformed summation plan"); return 1; }
    if (c->plan_depth > 8) { set_error("wcx_newref_load: too many samples (summation tree deeper than 8)"); return 1; }
    if (c->leaves_dev.ensure(sizeof(int32_t) * leaves.size())) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->leaves_dev.p, leaves.data(), sizeof(int32_t) * leaves.size(), cudaMemcpyHostToDevice, st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));  // `leaves` goes out of scope
  }
  if (c->plan_dev.ensure(sizeof(int32_t) * 3 * (size_t)pl)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->plan_dev.p, plan.data(), sizeof(int32_t) * 3 * (size_t)pl, cudaMemcpyHostToDevice, st));
  // leaf-major copy of X for the re-rank gather
  std::vector<int32_t> perm, leafdesc;
  c->sp = build_leaf_layout(plan.data(), pl, perm, leafdesc);
  c->leaf_n = c->sp > 0 ? (int32_t)(leafdesc.size() / 4) : 0;
  if (c->leaf_n >= 1 && c->leaf_n <= 8) {
    if (c->xp.ensure(sizeof(double) * (size_t)n * c->sp) || c->perm_dev.ensure(sizeof(int32_t) * perm.size()) ||
        c->leafdesc_dev.ensure(sizeof(int32_t) * leafdesc.size()))
      return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->perm_dev.p, perm.data(), sizeof(int32_t) * perm.size(), cudaMemcpyHostToDevice, st));
    WCX_CUDA_OK(cudaMemcpyAsync(c->leafdesc_dev.p, leafdesc.data(), sizeof(int32_t) * leafdesc.size(), cudaMemcpyHostToDevice, st));
    if (launch_permute_rows(c->d_x, n, s, c->perm_dev.as<int32_t>(), c->sp, c->xp.as<double>(), st)) return 1;
    c->launches += 1;
  } else {
    c->leaf_n = 0;
  }
  WCX_CUDA_OK(cudaEventRecord(c->ev[7], st));
  if (tc_encode_tensor_map(prep_view_h(c), c->tmap_h)) return 1;
  WCX_CUDA_OK(cudaStreamSynchronize(st));  // `plan` and caller's host buffers may go away
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]);
  c->stage_ms[4] = ms;
  c->loaded = true;
  c->last_rb = c->last_re = -1;
  return 0;
}

// fp32 (tf32-rounded) operands + norms + tensor map for the tf32 / CUDA-core sweeps; the default f16 sweep does not
// need them (1.3 ms and 0.39 GB at config 3)
static int ensure_tf32(wcx_ctx* c) {
  if (c->tf32_ready) return 0;
  if (c->xc.ensure(sizeof(float) * (size_t)c->n_pad * c->k_pad) || c->norm.ensure(sizeof(float) * (size_t)c->n_pad)) return 1;
  if (launch_center_round(c->d_x, c->n, c->s, c->colsum.as<double>(), c->colcnt.as<double>(), c->xc.as<float>(),
                          c->norm.as<float>(), c->n_pad, c->k_pad, c->stream))
    return 1;
  c->launches += 1;
  c->tf32_ready = true;
  return 0;
}

static void build_items(const wcx_ctx* c, int64_t rb, int64_t re, int tile_n, int nsplit, int lists_per_split,
                        std::vector<WorkItem>& items) {
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
        w.slot0 = (int32_t)((r0 - rb) * nsplit * lists_per_split + q * lists_per_split);
        w.slot_stride = nsplit * lists_per_split;
        items.push_back(w);
      }
    }
  }
}

}  // extern "C"

// null ratios requested together with the top-k (wcx_newref_reference): computed inside the re-rank kernel when
// the shape allows it, otherwise by the stand-alone kernels after it
// rows [r0, r1) of a call (relative to row_begin) whose candidate lists share one layout: nsplit column ranges x lists
// per range, the first list of the region at slot_base
struct Region {
  int64_t r0, r1;
  int nsplit;
  size_t slot_base;
};

struct NullPlan {
  const int32_t* sample_ids;
  int32_t m;
  double* out;
};

static int topk_impl(wcx_ctx* c, int64_t rb, int64_t re, int32_t k, int32_t kernel, int32_t* idx_out, double* dist_out,
                     int32_t out_on_device, const NullPlan* np) {
  if (!c || !c->loaded) { set_error("wcx_newref_topk: call wcx_newref_load first"); return 1; }
  if (rb < 0 || re > c->n || rb > re) { set_error("wcx_newref_topk: bad row range"); return 1; }
  if (k <= 0 || k > 512) { set_error("wcx_newref_topk: ref_size must be in [1, 512]"); return 1; }
  if (k > 400) kernel = WCX_KERNEL_EXACT;  // beyond the candidate-list guarantee of the sweep (2 x WCX_CAND_KEEP_TC per row)
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t rows = re - rb;
  std::memset(c->stats, 0, sizeof(c->stats));
  c->stage_ms[0] = c->stage_ms[1] = c->stage_ms[2] = 0.0;
  if (rows == 0) return 0;
  if (kernel == WCX_KERNEL_AUTO || kernel == WCX_KERNEL_TC2) kernel = WCX_KERNEL_TC2H;  // (ids of the removed tf32 variants
  if (kernel == WCX_KERNEL_TC) kernel = WCX_KERNEL_TCH;                                  //  select the f16 kernel of the same layout)
  if (kernel < WCX_KERNEL_TC || kernel > WCX_KERNEL_TCH) { set_error("wcx_newref_topk: unknown kernel id"); return 1; }
  const bool f16 = kernel == WCX_KERNEL_TC2H || kernel == WCX_KERNEL_TCH;
  const bool pair = kernel == WCX_KERNEL_TC2 || kernel == WCX_KERNEL_TC2H;
  const int gon = c->nchr > 22 ? 1 : 0;
  if (c->idx_dev.ensure(sizeof(int32_t) * (size_t)rows * k) || c->dist_dev.ensure(sizeof(double) * (size_t)rows * k))
    return 1;
  if (c->fail.ensure(sizeof(int32_t) * (size_t)rows)) return 1;
  if (kernel == WCX_KERNEL_SIMT && ensure_tf32(c)) return 1;  // the CUDA-core cross-check kernel reads fp32 operands
  PrepView pv = f16 ? prep_view_h(c) : prep_view(c);
  void* tmap = f16 ? c->tmap_h : c->tmap;
  std::vector<int32_t> fail_list;
  // null plan: gather the chosen sample columns first (needs only X), decide whether the re-rank kernel can fuse
  double* d_null = nullptr;
  c->stage_ms[3] = 0.0;
  if (np && np->m > 0) {
    for (int i = 0; i < np->m; i++)
      if (np->sample_ids[i] < 0 || np->sample_ids[i] >= c->s) { set_error("wcx_newref_reference: sample id out of range"); return 1; }
    if (c->xt.ensure(sizeof(double) * (size_t)null_ratio_staging_doubles(c->n, np->m)) || c->ids_dev.ensure(sizeof(int32_t) * np->m)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->ids_dev.p, np->sample_ids, sizeof(int32_t) * np->m, cudaMemcpyHostToDevice, st));
    if (launch_transpose_cols(c->d_x, c->n, c->s, c->ids_dev.as<int32_t>(), np->m, c->xt.as<double>(), st)) return 1;
    c->launches += 1;
    d_null = np->out;
    if (!out_on_device) {
      if (c->nr_dev.ensure(sizeof(double) * (size_t)rows * np->m)) return 1;
      d_null = c->nr_dev.as<double>();
    }
  }

  if (kernel == WCX_KERNEL_EXACT) {
    // every real row through the brute-force path; placeholder rows via the rerank kernel's early exit
    for (int64_t r = rb; r < re; r++) {
      int ch = 0;
      while (ch < c->nchr && c->cum[ch] <= r) ch++;
      if (gon && ch != 22 && ch != 23) continue;
      fail_list.push_back((int32_t)(r - rb));
    }
    if (gon) {
      // rerank kernel with nsplit = 0 would touch lists; fill placeholders on the host side instead
      std::vector<int32_t> hi((size_t)rows * k, 0);
      std::vector<double> hd((size_t)rows * k, 1.0);
      WCX_CUDA_OK(cudaMemcpyAsync(c->idx_dev.p, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, st));
      WCX_CUDA_OK(cudaMemcpyAsync(c->dist_dev.p, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice, st));
      WCX_CUDA_OK(cudaStreamSynchronize(st));
    }
  } else {
    const int tile_n = kernel == WCX_KERNEL_SIMT ? WCX_TILE_N_SIMT : WCX_TILE_N_TC;
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device);
    // Work items.  Every item is one tile of 128 target rows against the candidate column tiles; a persistent grid of
    // CTA pairs takes the units (two items = 256 rows) round-robin, so the sweep lasts ceil(units / pairs) unit times:
    // 10.12 rounds of work in 11 at config 3, 1.27 in 2 on 1/8 of the rows (8 GPUs).  Two regions of rows:
    //   A  the units that fill whole rounds of the grid -- the main launch;
    //   B  the units of the last, partial round -- a second launch on a high-priority stream that needs only 2 L of the
    //      SMs and runs WHILE region A is already in the re-rank (tensor-bound next to a gather-latency-bound kernel).
    // (Splitting region B into s candidate-column ranges instead, 2 s lists per row, was measured and does not pay: on 1/s of
    // the columns a list's threshold stays s times looser, the epilogue appends s times more entries and the first tiles of
    // every sub-unit append everything; the tail got no shorter on 8 GPUs.  The Region / nlists machinery below still
    // supports it.)
    std::vector<WorkItem> items;
    const int lps = kernel == WCX_KERNEL_SIMT ? 1 : 2;  // the tcgen05 kernel keeps one list per epilogue group
    build_items(c, rb, re, tile_n, 1, lps, items);
    const int nct = (int)((c->n + tile_n - 1) / tile_n);
    Region regA{0, rows, 1, 0}, regB{rows, rows, 1, 0};
    bool tail_overlap = false;
    size_t n_items_a = 0;
    if (pair) {
      const int n_units = ((int)items.size() + 1) / 2;
      const int P = std::max(1, dev_sms / 2);  // CTA pairs of the persistent grid
      static const bool no_overlap = std::getenv("WCX_TAIL_SERIAL") != nullptr;  // measurement switch: one launch, no overlap
      const int full_units = (n_units / P) * P;
      const int L = full_units > 0 ? n_units - full_units : 0;  // fewer units than CTA pairs: one launch, nothing to overlap
      const int sB = 1;
      tail_overlap = L > 0 && !no_overlap;
      std::vector<WorkItem> base;
      base.swap(items);
      const size_t nA = L > 0 ? std::min(base.size(), (size_t)2 * full_units) : base.size();
      const int64_t rowsA = nA < base.size() ? (int64_t)base[nA].row0 - rb : rows;
      regA = Region{0, rowsA, 1, 0};
      regB = Region{rowsA, rows, sB, (size_t)rowsA * lps};
      for (size_t i = 0; i < nA; i++) items.push_back(base[i]);  // slot0 / stride of build_items(nsplit = 1) are region A's
      if (items.size() & 1) {  // (only when there is no region B) pad the odd item count
        WorkItem d = items.back();
        d.row0 = 0; d.nrows = 0; d.slot0 = 0;
        items.push_back(d);
      }
      n_items_a = items.size();
      // region B: pair mode wants items (2p, 2p + 1) on the same column range -> split-major order, odd groups padded
      for (int q = 0; q < sB; q++) {
        int cnt = 0;
        WorkItem last{};
        for (size_t i = nA; i < base.size(); i++) {
          WorkItem w = base[i];
          w.ct_begin = (int)((int64_t)nct * q / sB);
          w.ct_end = (int)((int64_t)nct * (q + 1) / sB);
          w.slot0 = (int32_t)(regB.slot_base + ((int64_t)w.row0 - rb - rowsA) * sB * lps + q * lps);
          w.slot_stride = sB * lps;
          items.push_back(w);
          last = w;
          cnt++;
        }
        if (cnt & 1) {
          WorkItem d = last;
          d.row0 = 0; d.nrows = 0; d.slot0 = 0;
          items.push_back(d);
        }
      }
    } else {
      // one CTA per SM / CUDA-core kernels (cross-check paths): one uniform column split as before
      const int64_t row_tiles = (int64_t)items.size();
      int nsplit = 1;
      while (nsplit < 2 && row_tiles * nsplit < 4 * dev_sms && nct / (nsplit * 2) >= 8) nsplit *= 2;
      if (nsplit > 1) {
        items.clear();
        build_items(c, rb, re, tile_n, nsplit, lps, items);
      }
      regA = Region{0, rows, nsplit, 0};
    }
    const size_t slots = (size_t)(regA.r1 - regA.r0) * regA.nsplit * lps + (size_t)(regB.r1 - regB.r0) * regB.nsplit * lps;
    if (slots > 0x7fffffffull) { set_error("wcx_newref_topk: too many candidate lists for one call (split the row range)"); return 1; }
    if (c->cand_ent.ensure(sizeof(uint2) * slots * WCX_CAND_CAP) || c->cand_cnt.ensure(sizeof(int32_t) * slots) || c->cand_cut.ensure(sizeof(float) * slots))
      return 1;
    if (c->items_dev.ensure(sizeof(WorkItem) * std::max<size_t>(items.size(), 1)) || c->counter.ensure(sizeof(int32_t))) return 1;
    if (!items.empty())
      WCX_CUDA_OK(cudaMemcpyAsync(c->items_dev.p, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, st));
    if (c->diag.ensure(sizeof(int32_t) * 8)) return 1;
    WCX_CUDA_OK(cudaMemsetAsync(c->diag.p, 0, sizeof(int32_t) * 8, st));
    CandView cv{c->cand_ent.as<uint2>(), c->cand_cnt.as<int32_t>(), c->cand_cut.as<float>(), c->diag.as<int32_t>()};
    WCX_CUDA_OK(cudaMemsetAsync(c->cand_cnt.p, 0, sizeof(int32_t) * slots, st));
    WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
    if (kernel == WCX_KERNEL_SIMT) {
      if (launch_dist_topk_simt(pv, c->items_dev.as<WorkItem>(), (int)items.size(), cv, c->counter.as<int32_t>(), st)) return 1;
    } else if (pair && tail_overlap) {
      // region B on the tail stream (ready as soon as the lists are zeroed: its CTA pairs take the first SMs the main
      // launch frees), region A on the caller's stream; the re-rank of region A follows region A only
      WCX_CUDA_OK(cudaStreamWaitEvent(c->tail_stream, c->ev[0], 0));
      if (launch_dist_topk_tc_pair(pv, c->items_dev.as<WorkItem>(), (int)n_items_a, cv, tmap, st)) return 1;
      WCX_CUDA_OK(cudaEventRecord(c->ev_tailsweep[0], c->tail_stream));
      if (launch_dist_topk_tc_pair(pv, c->items_dev.as<WorkItem>() + n_items_a, (int)(items.size() - n_items_a), cv, tmap, c->tail_stream)) return 1;
      WCX_CUDA_OK(cudaEventRecord(c->ev_tailsweep[1], c->tail_stream));
      c->launches += 1;
    } else if (pair) {
      if (launch_dist_topk_tc_pair(pv, c->items_dev.as<WorkItem>(), (int)items.size(), cv, tmap, st)) return 1;
    } else {
      if (launch_dist_topk_tc(pv, c->items_dev.as<WorkItem>(), (int)items.size(), cv, c->counter.as<int32_t>(), tmap, st)) return 1;
    }
    WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
    {
      // The rows go through the re-rank in a few blocks.  The null ratios of a finished block run on a second,
      // high-priority stream next to the re-rank of the following block (the re-rank waits on gathers at ~50 %
      // issue utilisation, the median selection is issue bound), and with host outputs the D2H copy of a finished
      // block (copy stream) overlaps the kernels of the next one.
      static const bool serial_nulls = std::getenv("WCX_SERIAL_NULLS") != nullptr;
      // Device-resident outputs: nothing to copy while the kernels run, so every region is one re-rank launch followed by
      // one null-ratio launch on the same stream (measured r03o: 79.8 ms per step against 82.0 with eight row blocks and
      // the null ratios of a finished block on a side stream -- the two kernels take turns on the SMs anyway, and every
      // extra launch has its own tail).  Host outputs keep the row blocks: the D2H copy of a block hides behind the
      // kernels of the next one.
      const bool side = d_null && !serial_nulls && !out_on_device;   // null ratios on the side stream
      int bq = 0;       // block counter over both regions (events)
      int last_blk = -1;
      for (const Region& rg : {regA, regB}) {
        const int64_t rrows = rg.r1 - rg.r0;
        if (rrows <= 0) continue;
        if (tail_overlap && rg.r0 == regB.r0 && rg.r1 == regB.r1) WCX_CUDA_OK(cudaStreamWaitEvent(st, c->ev_tailsweep[1], 0));
        const int nlists = rg.nsplit * lps;
        // >= 4 k rows per block (7 waves of re-rank CTAs); with host outputs the copy of a block hides behind the kernels of
        // the next one, so a part of 24 k rows (8 GPUs) wants more than two blocks
        const int nblk = out_on_device ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(rg.nsplit > 1 && pair ? 3 : 8, rrows / 4000));
        for (int bi = 0; bi < nblk; bi++, bq++) {
          const int64_t r0 = rg.r0 + rrows * bi / nblk, r1 = rg.r0 + rrows * (bi + 1) / nblk;
          if (r1 <= r0) continue;
          const size_t s0 = rg.slot_base + (size_t)(r0 - rg.r0) * nlists;
          CandView cvb{cv.ent + s0 * WCX_CAND_CAP, cv.cnt + s0, cv.cut + s0, cv.diag};
          if (launch_rerank(c->d_x, pv, cvb, nlists, c->cum_dev.as<int64_t>(), c->nchr, rb + r0, rb + r1, k, gon,
                            c->idx_dev.as<int32_t>() + r0 * k, c->dist_dev.as<double>() + r0 * k, c->fail.as<int32_t>() + r0,
                            c->plan_dev.as<int32_t>(), c->plan_len, c->leaf_n ? c->xp.as<double>() : nullptr, c->sp,
                            c->leafdesc_dev.as<int32_t>(), c->leaf_n, st))
            return 1;
          c->launches += 1;
          WCX_CUDA_OK(cudaEventRecord(c->ev_blk[bq], st));
          cudaEvent_t ready = c->ev_blk[bq];
          if (d_null) {
            cudaStream_t ns = side ? c->null_stream : st;
            if (side) WCX_CUDA_OK(cudaStreamWaitEvent(ns, c->ev_blk[bq], 0));
            if (launch_null_ratios(c->xt.as<double>(), c->n, c->idx_dev.as<int32_t>() + r0 * k, rb + r0, rb + r1, k, np->m,
                                   d_null + r0 * np->m, ns))
              return 1;
            c->launches += (np->m + 7) / 8;
            WCX_CUDA_OK(cudaEventRecord(c->ev_null[bq], ns));
            ready = c->ev_null[bq];
            last_blk = bq;
          }
          if (!out_on_device) {
            WCX_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, ready, 0));
            if (idx_out) WCX_CUDA_OK(cudaMemcpyAsync(idx_out + r0 * k, c->idx_dev.as<int32_t>() + r0 * k, sizeof(int32_t) * (size_t)(r1 - r0) * k, cudaMemcpyDeviceToHost, c->copy_stream));
            if (dist_out) WCX_CUDA_OK(cudaMemcpyAsync(dist_out + r0 * k, c->dist_dev.as<double>() + r0 * k, sizeof(double) * (size_t)(r1 - r0) * k, cudaMemcpyDeviceToHost, c->copy_stream));
            if (d_null) WCX_CUDA_OK(cudaMemcpyAsync(np->out + r0 * np->m, d_null + r0 * np->m, sizeof(double) * (size_t)(r1 - r0) * np->m, cudaMemcpyDeviceToHost, c->copy_stream));
          }
        }
      }
      // the caller's stream owns the results again once the side stream has drained
      WCX_CUDA_OK(cudaEventRecord(c->ev_tail[0], st));
      if (side && last_blk >= 0) WCX_CUDA_OK(cudaStreamWaitEvent(st, c->ev_null[last_blk], 0));
      WCX_CUDA_OK(cudaEventRecord(c->ev_tail[1], st));
    }
    WCX_CUDA_OK(cudaEventRecord(c->ev[2], st));
    c->launches += 1;
    std::vector<int32_t> flags((size_t)rows);
    uint32_t diag_h[8] = {};
    WCX_CUDA_OK(cudaMemcpyAsync(diag_h, c->diag.p, sizeof(diag_h), cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(flags.data(), c->fail.p, sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));
    static const bool debug_fail = std::getenv("WCX_DEBUG_FAIL") != nullptr;
    for (int64_t i = 0; i < rows; i++)
      if (flags[(size_t)i]) {
        fail_list.push_back((int32_t)i);
        if (debug_fail) fprintf(stderr, "[wcx] row %lld not certified by the fast path (reason %d) -> exact_rows\n", (long long)(rb + i), flags[(size_t)i]);
      }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stage_ms[0] = ms;
    if (tail_overlap) {
      // + the time the partial last round needs on top of the main launch (it starts on the first freed SMs and ends
      // inside the re-rank of region A): sweep = first list write to last list write
      float ms_b = 0.f;
      cudaEventElapsedTime(&ms_b, c->ev[0], c->ev_tailsweep[1]);
      c->stage_ms[7] = ms_b > ms ? ms_b - ms : 0.f;
      c->stage_ms[0] = ms_b > ms ? ms_b : ms;
    } else {
      c->stage_ms[7] = 0.0;
    }
    cudaEventElapsedTime(&ms, c->ev[1], c->ev_tail[0]);
    c->stage_ms[1] = ms;
    c->stats[0] = (int64_t)items.size();
    c->stats[3] = regA.r1 - regA.r0;  // rows swept by the main launch (the rest: the overlapped partial round)
    c->stats[5] = diag_h[0];
    c->stats[6] = diag_h[1];
    c->stats[7] = diag_h[2];
    c->exact_evals = diag_h[4];
    c->gathered = (int64_t)diag_h[5] * 16;
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
  if (d_null) {
    WCX_CUDA_OK(cudaEventRecord(c->ev[5], st));
    if (kernel != WCX_KERNEL_EXACT) {
      // rows the fast path did not certify got their indexes from exact_rows: their null ratios follow here
      for (int32_t i : fail_list) {
        if (launch_null_ratios(c->xt.as<double>(), c->n, c->idx_dev.as<int32_t>() + (size_t)i * k, rb + i, rb + i + 1, k, np->m,
                               d_null + (size_t)i * np->m, st))
          return 1;
        c->launches += (np->m + 7) / 8;
      }
    } else {
      if (launch_null_ratios(c->xt.as<double>(), c->n, c->idx_dev.as<int32_t>(), rb, re, k, np->m, d_null, st)) return 1;
      c->launches += (np->m + 7) / 8;
    }
    WCX_CUDA_OK(cudaEventRecord(c->ev[6], st));
  }
  c->stats[2] = c->launches;
  c->last_rb = rb;
  c->last_re = re;
  c->last_k = k;
  if (out_on_device) {
    if (idx_out) WCX_CUDA_OK(cudaMemcpyAsync(idx_out, c->idx_dev.p, sizeof(int32_t) * (size_t)rows * k, cudaMemcpyDeviceToDevice, st));
    if (dist_out) WCX_CUDA_OK(cudaMemcpyAsync(dist_out, c->dist_dev.p, sizeof(double) * (size_t)rows * k, cudaMemcpyDeviceToDevice, st));
  } else {
    const bool chunked = kernel != WCX_KERNEL_EXACT;  // the re-rank loop above already copied the certified rows
    if (!chunked) {
      if (idx_out) WCX_CUDA_OK(cudaMemcpyAsync(idx_out, c->idx_dev.p, sizeof(int32_t) * (size_t)rows * k, cudaMemcpyDeviceToHost, st));
      if (dist_out) WCX_CUDA_OK(cudaMemcpyAsync(dist_out, c->dist_dev.p, sizeof(double) * (size_t)rows * k, cudaMemcpyDeviceToHost, st));
    } else {
      WCX_CUDA_OK(cudaStreamSynchronize(c->copy_stream));  // the block copies must land before the patches below
      for (int32_t i : fail_list) {
        if (idx_out) WCX_CUDA_OK(cudaMemcpyAsync(idx_out + (size_t)i * k, c->idx_dev.as<int32_t>() + (size_t)i * k, sizeof(int32_t) * k, cudaMemcpyDeviceToHost, st));
        if (dist_out) WCX_CUDA_OK(cudaMemcpyAsync(dist_out + (size_t)i * k, c->dist_dev.as<double>() + (size_t)i * k, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
      }
    }
    if (d_null) {
      if (chunked) {
        for (int32_t i : fail_list)
          WCX_CUDA_OK(cudaMemcpyAsync(np->out + (size_t)i * np->m, d_null + (size_t)i * np->m, sizeof(double) * np->m, cudaMemcpyDeviceToHost, st));
      } else {
        WCX_CUDA_OK(cudaMemcpyAsync(np->out, d_null, sizeof(double) * (size_t)rows * np->m, cudaMemcpyDeviceToHost, st));
      }
    }
    WCX_CUDA_OK(cudaStreamSynchronize(st));
  }
  if (d_null) {
    if (out_on_device) WCX_CUDA_OK(cudaStreamSynchronize(st));
    // what the null ratios add after the last re-rank block (they overlap the re-rank otherwise) + the failed rows
    float ms = 0.f, ms2 = 0.f;
    cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
    if (kernel != WCX_KERNEL_EXACT) cudaEventElapsedTime(&ms2, c->ev_tail[0], c->ev_tail[1]);
    c->stage_ms[3] = ms + ms2;
  }
  return 0;
}

extern "C" {

int wcx_newref_topk(wcx_ctx* c, int64_t rb, int64_t re, int32_t k, int32_t kernel, int32_t* idx_out,
                    double* dist_out, int32_t out_on_device) {
  return topk_impl(c, rb, re, k, kernel, idx_out, dist_out, out_on_device, nullptr);
}

int wcx_newref_reference(wcx_ctx* c, int64_t rb, int64_t re, int32_t k, int32_t kernel, const int32_t* sample_ids, int32_t m,
                         int32_t* idx_out, double* dist_out, double* null_out, int32_t out_on_device) {
  if (m < 0 || (m > 0 && (!sample_ids || !null_out))) { set_error("wcx_newref_reference: bad argument"); return 1; }
  NullPlan np{sample_ids, m, null_out};
  return topk_impl(c, rb, re, k, kernel, idx_out, dist_out, out_on_device, m > 0 ? &np : nullptr);
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
    // caller-supplied device positions are checked on the device before anything gathers through them
    int32_t bad = 0;
    if (c->diag.ensure(sizeof(int32_t) * 8)) return 1;
    WCX_CUDA_OK(cudaMemsetAsync(c->diag.p, 0, sizeof(int32_t), st));
    if (launch_validate_positions(idx, rows * (int64_t)k, c->n, c->diag.as<int32_t>(), st)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(&bad, c->diag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaStreamSynchronize(st));
    if (bad) { set_error("wcx_newref_null_ratios: index out of range [-n, n)"); return 1; }
    d_idx = idx;
  } else {
    // caller-supplied positions: Python semantics, -n <= v < n (negative wraps); anything else would gather out of bounds
    for (size_t i = 0; i < (size_t)rows * k; i++)
      if (idx[i] < -c->n || idx[i] >= c->n) { set_error("wcx_newref_null_ratios: index out of range [-n, n)"); return 1; }
    if (c->idx_dev.ensure(sizeof(int32_t) * (size_t)rows * k)) return 1;
    WCX_CUDA_OK(cudaMemcpyAsync(c->idx_dev.p, idx, sizeof(int32_t) * (size_t)rows * k, cudaMemcpyHostToDevice, st));
    d_idx = c->idx_dev.as<int32_t>();
    c->last_rb = rb; c->last_re = re; c->last_k = k;
  }
  if (c->xt.ensure(sizeof(double) * (size_t)null_ratio_staging_doubles(c->n, m)) || c->ids_dev.ensure(sizeof(int32_t) * m)) return 1;
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
  // one pass: sweep, then re-rank + null ratios block by block with the D2H copies of finished blocks overlapped
  return wcx_newref_reference(c, rb, re, k, kernel, sample_ids, null_out ? m : 0, idx_out, dist_out, null_out, 0);
}

int wcx_newref_stats(wcx_ctx* c, int64_t* out8) {
  if (!c || !out8) { set_error("null argument"); return 1; }
  std::memcpy(out8, c->stats, sizeof(c->stats));
  return 0;
}

int wcx_newref_stage_ms(wcx_ctx* c, double* out8) {
  if (!c || !out8) { set_error("null argument"); return 1; }
  c->stage_ms[5] = (double)c->exact_evals;
  c->stage_ms[6] = (double)c->gathered;
  std::memcpy(out8, c->stage_ms, sizeof(c->stage_ms));
  return 0;
}

static int debug_tile(wcx_ctx* c, int64_t row0, int64_t col0, float* acc_out, bool f16) {
  if (!c || !c->loaded || !acc_out) { set_error("wcx_debug_tc_tile: bad argument"); return 1; }
  if (row0 < 0 || row0 + WCX_TILE_M > c->n_pad || col0 < 0 || col0 % WCX_TILE_N_TC != 0 || col0 + WCX_TILE_N_TC > c->n_pad) {
    set_error("wcx_debug_tc_tile: tile out of range (col0 must be a multiple of 256)");
    return 1;
  }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  WorkItem w;
  w.row0 = (int32_t)row0; w.nrows = WCX_TILE_M; w.chr_s = -1; w.chr_e = -1;
  w.ct_begin = (int)(col0 / WCX_TILE_N_TC); w.ct_end = w.ct_begin + 1; w.slot0 = 0; w.slot_stride = 2;
  const size_t slots = 2 * WCX_TILE_M;
  if (c->cand_ent.ensure(sizeof(uint2) * slots * WCX_CAND_CAP) || c->cand_cnt.ensure(sizeof(int32_t) * slots) || c->cand_cut.ensure(sizeof(float) * slots) ||
      c->items_dev.ensure(sizeof(WorkItem)) || c->dbg.ensure(sizeof(float) * WCX_TILE_M * WCX_TILE_N_TC))
    return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->items_dev.p, &w, sizeof(w), cudaMemcpyHostToDevice, st));
  CandView cv{c->cand_ent.as<uint2>(), c->cand_cnt.as<int32_t>(), c->cand_cut.as<float>(), nullptr};
  if (!f16) { set_error("wcx_debug_tc_tile: only the f16 operand set is supported"); return 1; }
  PrepView pv = f16 ? prep_view_h(c) : prep_view(c);
  if (launch_dist_topk_tc_debug(pv, c->items_dev.as<WorkItem>(), 1, cv, f16 ? c->tmap_h : c->tmap, c->dbg.as<float>(), st)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(acc_out, c->dbg.p, sizeof(float) * WCX_TILE_M * WCX_TILE_N_TC, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  c->last_rb = c->last_re = -1;
  return 0;
}

int wcx_debug_tc_tile_f16(wcx_ctx* c, int64_t row0, int64_t col0, float* acc_out) { return debug_tile(c, row0, col0, acc_out, true); }

int wcx_debug_prep_f16(wcx_ctx* c, uint16_t* xh_out, float* norm_out, int32_t* k_pad_out, double* scale_out) {
  if (!c || !c->loaded) { set_error("wcx_debug_prep_f16: not loaded"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  if (k_pad_out) *k_pad_out = c->k_pad_h;
  if (xh_out) WCX_CUDA_OK(cudaMemcpyAsync(xh_out, c->xh.p, 2 * (size_t)c->n * c->k_pad_h, cudaMemcpyDeviceToHost, c->stream));
  if (norm_out) WCX_CUDA_OK(cudaMemcpyAsync(norm_out, c->norm_h.p, sizeof(float) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
  if (scale_out) WCX_CUDA_OK(cudaMemcpyAsync(scale_out, c->scale_dev.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int wcx_debug_list_counts(wcx_ctx* c, int32_t* cnt_out, int64_t nslots) {
  if (!c || !cnt_out) { set_error("wcx_debug_list_counts: bad argument"); return 1; }
  if ((size_t)nslots * sizeof(int32_t) > c->cand_cnt.cap) { set_error("wcx_debug_list_counts: more slots than allocated"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  WCX_CUDA_OK(cudaMemcpyAsync(cnt_out, c->cand_cnt.p, sizeof(int32_t) * (size_t)nslots, cudaMemcpyDeviceToHost, c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// host only (no device needed): the leaf-major layout of a row of S samples (rerank.cu build_leaf_layout)
int wcx_debug_leaf_layout(int32_t s, int32_t* perm_out, int32_t perm_cap, int32_t* desc_out, int32_t desc_cap, int32_t* plan_out,
                          int32_t plan_cap, int32_t* sizes3) {
  if (s <= 0 || !sizes3) { set_error("wcx_debug_leaf_layout: bad argument"); return 1; }
  std::vector<int32_t> plan(3 * 4096);
  const int pl = build_sum_plan(s, plan.data(), (int32_t)plan.size());
  if (pl < 0) { set_error("wcx_debug_leaf_layout: too many samples"); return 1; }
  std::vector<int32_t> perm, desc;
  const int sp = build_leaf_layout(plan.data(), pl, perm, desc);
  if (sp < 0) { set_error("wcx_debug_leaf_layout: leaf too long"); return 1; }
  sizes3[0] = sp; sizes3[1] = (int32_t)desc.size() / 4; sizes3[2] = pl;
  if (perm_out) { if (perm_cap < sp) { set_error("perm_cap too small"); return 1; } std::memcpy(perm_out, perm.data(), sizeof(int32_t) * sp); }
  if (desc_out) { if (desc_cap < (int32_t)desc.size()) { set_error("desc_cap too small"); return 1; } std::memcpy(desc_out, desc.data(), sizeof(int32_t) * desc.size()); }
  if (plan_out) { if (plan_cap < 3 * pl) { set_error("plan_cap too small"); return 1; } std::memcpy(plan_out, plan.data(), sizeof(int32_t) * 3 * pl); }
  return 0;
}

int wcx_debug_prep(wcx_ctx* c, float* xc_out, float* norm_out, int32_t* k_pad_out) {
  if (!c || !c->loaded) { set_error("wcx_debug_prep: not loaded"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  if (ensure_tf32(c)) return 1;
  if (k_pad_out) *k_pad_out = c->k_pad;
  if (xc_out) WCX_CUDA_OK(cudaMemcpyAsync(xc_out, c->xc.p, sizeof(float) * (size_t)c->n * c->k_pad, cudaMemcpyDeviceToHost, c->stream));
  if (norm_out) WCX_CUDA_OK(cudaMemcpyAsync(norm_out, c->norm.p, sizeof(float) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ================================================================================================
// predict
// ================================================================================================
static int h2d(DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
  if (b.ensure(bytes ? bytes : 1)) return 1;
  if (bytes) WCX_CUDA_OK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
  return 0;
}

int wcx_predict_load_ref(wcx_ctx* c, int32_t set_id, const int32_t* idx, const double* dist, int64_t n, int32_t k,
                         const int64_t* per, const int64_t* cum, int32_t nchr, const double* comps, const double* mean,
                         int32_t ncomp, const int32_t* mask_pos, int64_t bins_total) {
  if (!c || set_id < 0 || set_id > 2 || !idx || !dist || !per || !cum || !comps || !mean || !mask_pos) {
    set_error("wcx_predict_load_ref: bad argument");
    return 1;
  }
  if (n <= 0 || k <= 0 || k > 512 || nchr <= 0 || ncomp <= 0 || ncomp > 8 || cum[nchr - 1] != n) {
    set_error("wcx_predict_load_ref: inconsistent shapes (need 0 < k <= 512, ncomp <= 8, cum[-1] == n)");
    return 1;
  }
  for (int i = 0; i < nchr; i++)
    if (cum[i] - (i ? cum[i - 1] : 0) != per[i]) { set_error("wcx_predict_load_ref: per/cum inconsistent"); return 1; }
  for (int64_t i = 0; i < n; i++)
    if (mask_pos[i] < 0 || mask_pos[i] >= bins_total) { set_error("wcx_predict_load_ref: mask position out of range"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  RefSet& r = c->ref[set_id];
  r.loaded = false;
  r.gl_valid = false;
  if (h2d(r.idx, idx, sizeof(int32_t) * (size_t)n * k, st) || h2d(r.dist, dist, sizeof(double) * (size_t)n * k, st) ||
      h2d(r.cum, cum, sizeof(int64_t) * nchr, st) || h2d(r.comps, comps, sizeof(double) * (size_t)ncomp * n, st) ||
      h2d(r.mean, mean, sizeof(double) * (size_t)n, st) || h2d(r.mask_pos, mask_pos, sizeof(int32_t) * (size_t)n, st))
    return 1;
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  r.n = n; r.k = k; r.nchr = nchr; r.ncomp = ncomp; r.bins_total = bins_total;
  r.cum_h.assign(cum, cum + nchr);
  r.loaded = true;
  return 0;
}

static RefSet* get_ref(wcx_ctx* c, int32_t set_id) {
  if (!c || set_id < 0 || set_id > 2 || !c->ref[set_id].loaded) {
    set_error("predict: reference set not loaded (wcx_predict_load_ref)");
    return nullptr;
  }
  return &c->ref[set_id];
}

int wcx_predict_weights(wcx_ctx* c, int32_t set_id, double* out) {
  RefSet* r = get_ref(c, set_id);
  if (!r || !out) { if (r) set_error("wcx_predict_weights: null output"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (c->p_w.ensure(sizeof(double) * (size_t)r->n)) return 1;
  if (launch_weights(r->dist.as<double>(), r->n, r->k, c->p_w.as<double>(), st)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(out, c->p_w.p, sizeof(double) * (size_t)r->n, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  c->launches += 1;
  return 0;
}

int wcx_predict_optimal_cutoff(wcx_ctx* c, int32_t set_id, int32_t repeats, double* cutoff_out) {
  RefSet* r = get_ref(c, set_id);
  if (!r || !cutoff_out || repeats < 0) { if (r) set_error("wcx_predict_optimal_cutoff: bad argument"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (c->p_state.ensure(sizeof(double) * 4) || c->p_partial.ensure(sizeof(double) * (size_t)predict_red_blocks() * 8 * 128)) return 1;
  if (launch_optimal_cutoff(r->dist.as<double>(), r->n * (int64_t)r->k, repeats, c->p_state.as<double>(), c->p_partial.as<double>(), st)) return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(cutoff_out, c->p_state.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  c->launches += 4 * repeats;
  return 0;
}

int wcx_predict_normalize(wcx_ctx* c, int32_t set_id, const double* raw, int32_t B, double cutoff, int32_t cp, int64_t ct,
                          double* z, double* rr, double* nref, double* m_lr, double* m_z) {
  RefSet* r = get_ref(c, set_id);
  if (!r) return 1;
  if (!raw || !z || !rr || !nref || !m_lr || !m_z || B <= 0 || B > 128) { set_error("wcx_predict_normalize: bad argument (1 <= B <= 128)"); return 1; }
  if (cp < 0 || cp > r->nchr || ct != (cp == 0 ? 0 : r->cum_h[cp - 1])) {
    set_error("wcx_predict_normalize: ct must equal masked_bins_per_chr_cum[cp - 1]");
    return 1;
  }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = r->n, nout = n - ct;
  const size_t bn = sizeof(double) * (size_t)B * n, bo = sizeof(double) * (size_t)B * (nout > 0 ? nout : 1);
  if (c->p_raw.ensure(sizeof(double) * (size_t)B * r->bins_total) || c->p_x.ensure(bn) || c->p_copy_a.ensure(bn) || c->p_copy_b.ensure(bn) ||
      c->p_z.ensure(bo) || c->p_r.ensure(bo) || c->p_n.ensure(bo) || c->p_mlr.ensure(sizeof(double) * B) || c->p_mz.ensure(sizeof(double) * B) ||
      c->p_state.ensure(sizeof(double) * 4) || c->p_partial.ensure(sizeof(double) * (size_t)predict_red_blocks() * 8 * 128) ||
      c->p_totals.ensure(sizeof(double) * 128) || c->p_tdots.ensure(sizeof(double) * 128 * 8) ||
      c->p_radix.ensure(radix_scratch_bytes(B, nout > 0 ? nout : 1)))
    return 1;
  WCX_CUDA_OK(cudaMemcpyAsync(c->p_raw.p, raw, sizeof(double) * (size_t)B * r->bins_total, cudaMemcpyHostToDevice, st));
  // the cutoff travels through device memory (state[3]) so the kernels read one source of truth
  WCX_CUDA_OK(cudaMemcpyAsync(c->p_state.as<double>() + 3, &cutoff, sizeof(double), cudaMemcpyHostToDevice, st));
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (launch_coverage_project(c->p_raw.as<double>(), B, r->bins_total, r->mask_pos.as<int32_t>(), n, r->comps.as<double>(),
                              r->mean.as<double>(), r->ncomp, c->p_x.as<double>(), c->p_partial.as<double>(),
                              c->p_totals.as<double>(), c->p_tdots.as<double>(), st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
  if (nout > 0 && !(r->gl_valid && r->gl_cutoff == cutoff && r->gl_ct == ct)) {
    r->gl_valid = false;
    if (r->gl.ensure(sizeof(int32_t) * (size_t)nout * r->k)) return 1;
    if (launch_gather_list(r->idx.as<int32_t>(), r->dist.as<double>(), n, r->k, c->p_state.as<double>() + 3, r->cum.as<int64_t>(),
                           r->nchr, ct, r->gl.as<int32_t>(), st))
      return 1;
    r->gl_cutoff = cutoff; r->gl_ct = ct; r->gl_valid = true;
    c->launches += 1;
  }
  WCX_CUDA_OK(cudaEventRecord(c->ev[2], st));
  if (launch_normalize_repeat(c->p_x.as<double>(), c->p_copy_a.as<double>(), c->p_copy_b.as<double>(), B, n, r->gl.as<int32_t>(),
                              r->k, ct, c->p_z.as<double>(), c->p_r.as<double>(), c->p_n.as<double>(), st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[3], st));
  if (nout > 0 && launch_nanmedians(c->p_r.as<double>(), c->p_z.as<double>(), B, nout, c->p_radix.p, c->p_mlr.as<double>(),
                                    c->p_mz.as<double>(), st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[4], st));
  if (nout > 0) {
    WCX_CUDA_OK(cudaMemcpyAsync(z, c->p_z.p, sizeof(double) * (size_t)B * nout, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(rr, c->p_r.p, sizeof(double) * (size_t)B * nout, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(nref, c->p_n.p, sizeof(double) * (size_t)B * nout, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(m_lr, c->p_mlr.p, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    WCX_CUDA_OK(cudaMemcpyAsync(m_z, c->p_mz.p, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  }
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->predict_ms[0] = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[4]);
  c->predict_ms[1] = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
  c->predict_ms[4] = ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
  c->predict_ms[5] = ms;
  cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]);
  c->predict_ms[6] = ms;
  c->launches += 6 + 3 + 1 + 6;  // coverage + projection, three passes, key pass + six radix passes
  return 0;
}

int wcx_segment_zscore(wcx_ctx* c, const double* nr, int64_t n_masked, int32_t m, const int32_t* inflate_pos, const double* r,
                       const double* w, int64_t bins_total, const int64_t* seg_se, const double* seg_r, int32_t nseg, double* z_out) {
  if (!c || !inflate_pos || !r || !w || !seg_se || !seg_r || !z_out || m <= 0 || m > 128 || nseg < 0) {
    set_error("wcx_segment_zscore: bad argument (1 <= null samples <= 128)");
    return 1;
  }
  if (!nr && (c->z_nr_rows != n_masked || c->z_nr_m != m || !c->z_nr.p)) {
    set_error("wcx_segment_zscore: nr == NULL but no null ratios of that shape are resident");
    return 1;
  }
  for (int i = 0; i < nseg; i++)
    if (seg_se[2 * i] < 0 || seg_se[2 * i + 1] > bins_total || seg_se[2 * i] > seg_se[2 * i + 1]) { set_error("wcx_segment_zscore: segment out of range"); return 1; }
  for (int64_t i = 0; i < bins_total; i++)
    if (inflate_pos[i] >= n_masked) { set_error("wcx_segment_zscore: inflate position out of range"); return 1; }
  if (nseg == 0) return 0;
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (nr) {
    // 0.16 GB at 15 kb: uploaded once per reference, later calls pass NULL
    c->z_nr_rows = -1;
    if (h2d(c->z_nr, nr, sizeof(double) * (size_t)n_masked * m, st)) return 1;
    c->z_nr_rows = n_masked;
    c->z_nr_m = m;
  }
  if (h2d(c->z_pos, inflate_pos, sizeof(int32_t) * (size_t)bins_total, st) ||
      h2d(c->z_r, r, sizeof(double) * (size_t)bins_total, st) || h2d(c->z_w, w, sizeof(double) * (size_t)bins_total, st) ||
      h2d(c->z_se, seg_se, sizeof(int64_t) * 2 * (size_t)nseg, st) || h2d(c->z_segr, seg_r, sizeof(double) * (size_t)nseg, st) ||
      c->z_out.ensure(sizeof(double) * (size_t)nseg) || c->z_partial.ensure(segment_z_scratch_bytes(nseg)))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (launch_segment_z(c->z_nr.as<double>(), m, c->z_pos.as<int32_t>(), c->z_r.as<double>(), c->z_w.as<double>(),
                       c->z_se.as<int64_t>(), c->z_segr.as<double>(), nseg, c->z_out.as<double>(), c->z_partial.as<double>(), st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
  WCX_CUDA_OK(cudaMemcpyAsync(z_out, c->z_out.p, sizeof(double) * (size_t)nseg, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->predict_ms[2] = ms;
  c->launches += 2;
  return 0;
}

int wcx_predict_stage_ms(wcx_ctx* c, double* out8) {
  if (!c || !out8) { set_error("null argument"); return 1; }
  std::memcpy(out8, c->predict_ms, sizeof(c->predict_ms));
  return 0;
}

// ================================================================================================
// CBS
// ================================================================================================
int wcx_cbs_segment(wcx_ctx* c, const double* y, const double* w, const int64_t* off, int32_t nseries,
                    const int32_t* series_ids, double alpha, int32_t nperm, uint32_t seed, int32_t* ends_out,
                    int32_t* nseg_out) {
  if (!c || !y || !w || !off || !ends_out || !nseg_out || nseries < 0) { set_error("wcx_cbs_segment: bad argument"); return 1; }
  if (!(alpha > 0.0) || alpha > 1.0 || nperm < 1) { set_error("wcx_cbs_segment: alpha must be in (0, 1], nperm >= 1"); return 1; }
  for (int s = 0; s < nseries; s++)
    if (off[s + 1] < off[s]) { set_error("wcx_cbs_segment: offsets must be ascending"); return 1; }
  for (int64_t i = 0; i < off[nseries]; i++)
    if (!(w[i] > 0.0) || !std::isfinite(y[i])) { set_error("wcx_cbs_segment: weights must be positive and values finite"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  if (!c->cbs) c->cbs = cbs_workspace_create();
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
  // DNAcopy 1.76 defaults of segment(): kmax = 25, nmin = 200, min.width = 2 (CBS.R:73 overrides only alpha / weights)
  if (cbs_segment(c->cbs, y, w, off, nseries, series_ids, alpha, nperm, 25, 200, 2, seed, ends_out, nseg_out, &c->cbs_stats, c->stream))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->predict_ms[3] = ms;
  c->launches += c->cbs_stats.launches;
  return 0;
}

int wcx_cbs_set_boundary(wcx_ctx* c, const int32_t* sbdry, int32_t n) {
  if (!c || n < 0 || (n > 0 && !sbdry)) { set_error("wcx_cbs_set_boundary: bad argument"); return 1; }
  for (int i = 0; i < n; i++)
    if (sbdry[i] < 1) { set_error("wcx_cbs_set_boundary: boundary values are permutation counts >= 1"); return 1; }
  if (!c->cbs) c->cbs = cbs_workspace_create();
  cbs_set_boundary(c->cbs, sbdry, n);
  return 0;
}

int wcx_cbs_stats(wcx_ctx* c, int64_t* out6) {
  if (!c || !out6) { set_error("null argument"); return 1; }
  out6[0] = c->cbs_stats.rounds; out6[1] = c->cbs_stats.segments_tested; out6[2] = c->cbs_stats.perm_tests;
  out6[3] = c->cbs_stats.t_tests; out6[4] = c->cbs_stats.permutations; out6[5] = c->cbs_stats.launches;
  return 0;
}

// ================================================================================================
// newref preparation: normalize_and_mask, train_pca, PCA-distance filter
// ================================================================================================
int wcx_newref_normalize_and_mask(wcx_ctx* c, const int32_t* counts, int64_t bins_total, int32_t s, const int32_t* mask_pos,
                                  int64_t n, double* out, int32_t out_on_device) {
  if (!c || !mask_pos || (!out && !out_on_device) || bins_total <= 0 || s <= 0 || n < 0) { set_error("wcx_newref_normalize_and_mask: bad argument"); return 1; }
  if (!counts && (c->q_counts_rows != bins_total || c->q_counts_s != s || !c->q_counts.p)) {
    set_error("wcx_newref_normalize_and_mask: counts == NULL but no count matrix of that shape is resident");
    return 1;
  }
  for (int64_t i = 0; i < n; i++)
    if (mask_pos[i] < 0 || mask_pos[i] >= bins_total) { set_error("wcx_newref_normalize_and_mask: mask position out of range"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (counts) {
    c->q_counts_rows = -1;
    if (h2d(c->q_counts, counts, sizeof(int32_t) * (size_t)bins_total * s, st)) return 1;
    c->q_counts_rows = bins_total;
    c->q_counts_s = s;
  }
  if (h2d(c->q_pos, mask_pos, sizeof(int32_t) * (size_t)n, st) || c->q_colsum.ensure(sizeof(unsigned long long) * s)) return 1;
  double* d_out = out;
  if (!out_on_device || !out) {
    // out == NULL with out_on_device: the matrix stays in the context (input of wcx_pca_gram(x = NULL))
    if (c->q_x.ensure(sizeof(double) * (size_t)std::max<int64_t>(n, 1) * s)) return 1;
    d_out = c->q_x.as<double>();
    c->q_xptr = d_out;
    c->q_n = n;
    c->q_s = s;
  }
  if (launch_normalize_and_mask(c->q_counts.as<int32_t>(), bins_total, s, c->q_pos.as<int32_t>(), n,
                                c->q_colsum.as<unsigned long long>(), d_out, st))
    return 1;
  if (!out_on_device && n > 0) WCX_CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n * s, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  c->launches += 2;
  return 0;
}

int wcx_pca_gram(wcx_ctx* c, const double* x, int64_t n, int32_t s, int32_t x_on_device, double* mean_out, double* gram_out) {
  if (!c || (!x && !x_on_device) || !mean_out || !gram_out || n <= 0 || s <= 0) { set_error("wcx_pca_gram: bad argument"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (x_on_device && !x) {
    // the matrix left in the context by wcx_newref_normalize_and_mask(out = NULL)
    if (c->q_xptr != c->q_x.as<double>() || !c->q_x.p || c->q_n != n || c->q_s != s) { set_error("wcx_pca_gram: no device-resident matrix of that shape"); return 1; }
  } else if (x_on_device) {
    c->q_xptr = x;
  } else {
    if (h2d(c->q_x, x, sizeof(double) * (size_t)n * s, st)) return 1;
    c->q_xptr = c->q_x.as<double>();
  }
  c->q_n = n; c->q_s = s;
  const int s_pad = gram_s_pad(s), nch = gram_chunks(n);
  if (c->q_mean.ensure(sizeof(double) * (size_t)n) || c->q_partial.ensure(sizeof(double) * (size_t)nch * s_pad * s_pad) ||
      c->q_gram.ensure(sizeof(double) * (size_t)s * s))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (launch_row_mean(c->q_xptr, n, s, c->q_mean.as<double>(), st)) return 1;
  if (launch_gram(c->q_xptr, c->q_mean.as<double>(), n, s, c->q_partial.as<double>(), c->q_gram.as<double>(), st)) return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
  WCX_CUDA_OK(cudaMemcpyAsync(mean_out, c->q_mean.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaMemcpyAsync(gram_out, c->q_gram.p, sizeof(double) * (size_t)s * s, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->prep_ms[0] = ms;
  c->launches += 3;
  return 0;
}

int wcx_pca_apply(wcx_ctx* c, const double* u, const double* sigma, int32_t ncomp, double* comps_out, double* corrected_out,
                  int32_t corrected_on_device) {
  if (!c || !u || !sigma || !comps_out || ncomp <= 0 || ncomp > 8 || !c->q_xptr) { set_error("wcx_pca_apply: bad argument or wcx_pca_gram not called"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = c->q_n;
  const int32_t s = c->q_s;
  if (h2d(c->q_u, u, sizeof(double) * (size_t)s * ncomp, st) || h2d(c->q_sigma, sigma, sizeof(double) * ncomp, st) ||
      c->q_comps.ensure(sizeof(double) * (size_t)ncomp * n))
    return 1;
  double* d_corr = corrected_out;
  if (!corrected_on_device || !corrected_out) {
    if (c->q_corr.ensure(sizeof(double) * (size_t)n * s)) return 1;
    d_corr = c->q_corr.as<double>();
  }
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (launch_pca_apply(c->q_xptr, c->q_mean.as<double>(), n, s, c->q_u.as<double>(), c->q_sigma.as<double>(), ncomp,
                       c->q_comps.as<double>(), d_corr, st))
    return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
  WCX_CUDA_OK(cudaMemcpyAsync(comps_out, c->q_comps.p, sizeof(double) * (size_t)ncomp * n, cudaMemcpyDeviceToHost, st));
  if (corrected_out && !corrected_on_device)
    WCX_CUDA_OK(cudaMemcpyAsync(corrected_out, d_corr, sizeof(double) * (size_t)n * s, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->prep_ms[1] = ms;
  c->launches += 1;
  return 0;
}

int wcx_pca_distance(wcx_ctx* c, const double* corrected, int64_t n, int32_t s, int32_t on_device, double* med_out, double* d_out) {
  if (!c || !med_out || !d_out || n <= 0 || s <= 0) { set_error("wcx_pca_distance: bad argument"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const double* d_x = nullptr;
  if (!corrected) {
    if (c->q_corr.cap < sizeof(double) * (size_t)n * s || c->q_n != n || c->q_s != s) { set_error("wcx_pca_distance: no device-resident corrected matrix"); return 1; }
    d_x = c->q_corr.as<double>();
  } else if (on_device) {
    d_x = corrected;
  } else {
    if (h2d(c->q_corr, corrected, sizeof(double) * (size_t)n * s, st)) return 1;
    d_x = c->q_corr.as<double>();
  }
  if (c->q_med.ensure(sizeof(double) * s) || c->q_d.ensure(sizeof(double) * (size_t)n) ||
      c->q_work.ensure(sizeof(unsigned long long) * (size_t)(67 * s)))
    return 1;
  unsigned long long* wk = c->q_work.as<unsigned long long>();
  WCX_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (launch_col_medians(d_x, n, s, wk, wk + s, wk + 2 * s, c->q_med.as<double>(), st)) return 1;
  if (launch_row_sqdist(d_x, n, s, c->q_med.as<double>(), c->q_d.as<double>(), st)) return 1;
  WCX_CUDA_OK(cudaEventRecord(c->ev[1], st));
  WCX_CUDA_OK(cudaMemcpyAsync(med_out, c->q_med.p, sizeof(double) * s, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaMemcpyAsync(d_out, c->q_d.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
  WCX_CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->prep_ms[2] = ms;
  c->launches += 132;
  return 0;
}

int wcx_prep_fetch(wcx_ctx* c, int32_t which, int64_t n, int32_t s, double* out) {
  if (!c || !out || n <= 0 || s <= 0 || c->q_n != n || c->q_s != s) { set_error("wcx_prep_fetch: bad argument or shape"); return 1; }
  const DevBuf& b = which == 0 ? c->q_x : c->q_corr;
  if (!b.p || b.cap < sizeof(double) * (size_t)n * s) { set_error("wcx_prep_fetch: that matrix is not resident"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  WCX_CUDA_OK(cudaMemcpyAsync(out, b.p, sizeof(double) * (size_t)n * s, cudaMemcpyDeviceToHost, c->stream));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int wcx_prep_device_ptr(wcx_ctx* c, int32_t which, int64_t n, int32_t s, const double** out) {
  if (!c || !out || c->q_n != n || c->q_s != s) { set_error("wcx_prep_device_ptr: bad argument or shape"); return 1; }
  const DevBuf& b = which == 0 ? c->q_x : c->q_corr;
  if (!b.p || b.cap < sizeof(double) * (size_t)n * s) { set_error("wcx_prep_device_ptr: that matrix is not resident"); return 1; }
  WCX_CUDA_OK(cudaSetDevice(c->device));
  WCX_CUDA_OK(cudaStreamSynchronize(c->stream));  // the caller may hand the pointer to another context / device
  *out = b.as<double>();
  return 0;
}

int wcx_newref_prep_stage_ms(wcx_ctx* c, double* out4) {
  if (!c || !out4) { set_error("null argument"); return 1; }
  std::memcpy(out4, c->prep_ms, sizeof(c->prep_ms));
  return 0;
}

}  // extern "C"
