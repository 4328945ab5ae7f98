// engine.cu -- engine lifetime, workspaces, acoustic-model upload/tiling, graph upload.
#include <mutex>
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cuda_internal.cuh"

namespace mfa {
static thread_local std::string g_last_error;
int set_error(int code, const std::string &msg) { g_last_error = msg; return code; }
}  // namespace mfa
using namespace mfa;

extern "C" const char *mfa_last_error(void) { return g_last_error.c_str(); }
extern "C" int mfa_abi_version(void) { return 2; }

int mfa_engine::get(int id, size_t bytes, void **out) {
  Buf &b = dev[id];
  if (bytes > b.cap) {
    if (b.p) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }   // every stream: a K3 in flight may use it
    size_t cap = bytes + bytes / 8 + 256;
    cudaError_t err = cudaMalloc(&b.p, cap);
    if (err != cudaSuccess) { b.p = nullptr; return set_error(MFA_ERR_NOMEM, std::string("cudaMalloc of ") + std::to_string(cap) + " bytes: " + cudaGetErrorString(err)); }
    b.cap = cap;
  }
  *out = b.p;
  return MFA_OK;
}

int mfa_engine::get_pinned(int id, size_t bytes, void **out) {
  Buf &b = pin[id];
  if (bytes > b.cap) {
    if (b.p) { CUDA_TRY(cudaStreamSynchronize(stream)); CUDA_TRY(cudaFreeHost(b.p)); b.p = nullptr; b.cap = 0; }
    size_t cap = bytes + bytes / 8 + 256;
    CUDA_TRY(cudaMallocHost(&b.p, cap));
    b.cap = cap;
  }
  *out = b.p;
  return MFA_OK;
}

namespace {
struct OptionDesc { const char *name; int mfa_engine_cfg::*field; };
const OptionDesc kOptions[] = {
    {"vit_band", &mfa_engine_cfg::vit_band}, {"vit_maxgroups", &mfa_engine_cfg::vit_maxgroups}, {"vit_wide", &mfa_engine_cfg::vit_wide}, {"vit_wide_poll", &mfa_engine_cfg::vit_wide_poll}, {"vit_wide_ctas", &mfa_engine_cfg::vit_wide_ctas}, {"vit_graph_smem", &mfa_engine_cfg::vit_graph_smem},
    {"vit_nw2_kb", &mfa_engine_cfg::vit_nw2_kb}, {"vit_carveout", &mfa_engine_cfg::vit_carveout}, {"vit_carveout_band", &mfa_engine_cfg::vit_carveout_band},
    {"vit_prio", &mfa_engine_cfg::vit_prio}, {"k3_overlap", &mfa_engine_cfg::k3_overlap}, {"pipeline_split", &mfa_engine_cfg::pipeline_split}, {"acc_impl", &mfa_engine_cfg::acc_impl},
    {"tc_k96", &mfa_engine_cfg::tc_k96}, {"tc_poly", &mfa_engine_cfg::tc_poly}, {"mfcc_generic", &mfa_engine_cfg::mfcc_generic},
    {"trace", &mfa_engine_cfg::trace}};
}  // namespace

// An engine drives ~17 streams (main, copy, Viterbi size classes, join, gather ...) and MFA runs several jobs per GPU.  CUDA maps streams
// onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8): streams sharing a queue serialise, and a long-running kernel of one job
// (the polling Viterbi fallback level) then holds back unrelated work of another.  Measured, 2 jobs end to end: 1.16 -> 1.35 M x RT with
// 32 queues.  The variable is read when the CUDA context is created, so it is set when the library is loaded (unless the user set it).
__attribute__((constructor)) static void mfa_b200_default_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" int mfa_engine_set_option(mfa_engine *e, const char *name, int value) {
  if (!e || !name) return set_error(MFA_ERR_INVALID, "null argument");
  for (const auto &o : kOptions)
    if (!strcmp(o.name, name)) { e->cfg.*(o.field) = value; return MFA_OK; }
  return set_error(MFA_ERR_INVALID, std::string("unknown engine option '") + name + "'");
}
extern "C" int mfa_engine_get_option(mfa_engine *e, const char *name, int *value) {
  if (!e || !name || !value) return set_error(MFA_ERR_INVALID, "null argument");
  for (const auto &o : kOptions)
    if (!strcmp(o.name, name)) { *value = e->cfg.*(o.field); return MFA_OK; }
  return set_error(MFA_ERR_INVALID, std::string("unknown engine option '") + name + "'");
}

extern "C" int mfa_engine_create(int device, mfa_engine **out) {
  if (!out) return set_error(MFA_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0)
    return set_error(MFA_ERR_CUDA, std::string("no usable CUDA device (this engine has no CPU fallback): ") + cudaGetErrorString(err));
  if (device < 0 || device >= n) return set_error(MFA_ERR_INVALID, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return set_error(MFA_ERR_UNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; this build targets sm_100a (B200) only");
  auto *e = new mfa_engine();
  for (const auto &o : kOptions) {   // the one place the environment is consulted
    std::string name = "MFA_";
    for (const char *c = o.name; *c; c++) name += (char)toupper((unsigned char)*c);
    if (const char *v = getenv(name.c_str())) e->cfg.*(o.field) = atoi(v);
  }
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  // side streams carry the Viterbi size classes (class k on side[k]); larger classes = longer utterances = the launch's critical path,
  // so they get the higher stream priorities and their CTAs are placed first (MFA_VIT_PRIO=0: all equal)
  int prio_least = 0, prio_greatest = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  const bool use_prio = e->cfg.vit_prio != 0;
  for (int k = 0; k < mfa_engine::kSide; k++) {
    const int prio = use_prio ? std::max(prio_greatest, prio_least - k) : prio_least;
    CUDA_TRY(cudaStreamCreateWithPriority(&e->side[k], cudaStreamNonBlocking, prio));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_bimg, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_k3_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fb, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithPriority(&e->sj, cudaStreamNonBlocking, prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&e->sw, cudaStreamNonBlocking, prio_greatest));
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_wide, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&e->sg, cudaStreamNonBlocking));
  CUDA_TRY(cudaHostAlloc((void **)&e->h_fb_ring, (mfa_engine::kFbRing + 1) * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
  memset(e->h_fb_ring, 0, (mfa_engine::kFbRing + 1) * sizeof(int32_t));
  for (int k = 0; k < 2; k++) {
    CUDA_TRY(cudaMallocHost(&e->stage_mem[k], mfa_engine::kStageBytes));
    CUDA_TRY(cudaEventCreateWithFlags(&e->stage_ev[k], cudaEventDisableTiming));
  }
  *out = e;
  return MFA_OK;
}

extern "C" int mfa_engine_destroy(mfa_engine *e) {
  if (!e) return MFA_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();   // main stream, K3's side / join streams, the gather stream
  for (auto &b : e->dev) if (b.p) cudaFree(b.p);
  for (auto &b : e->pin) if (b.p) cudaFreeHost(b.p);
  for (auto ev : e->gmm_ev) cudaEventDestroy(ev);
  for (auto ev : e->ev_piece) cudaEventDestroy(ev);
  for (auto ev : e->st_ev) cudaEventDestroy(ev);
  for (int k = 0; k < mfa_engine::kSide; k++) { if (e->side[k]) cudaStreamDestroy(e->side[k]); if (e->ev_join[k]) cudaEventDestroy(e->ev_join[k]); }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_bimg) cudaEventDestroy(e->ev_bimg);
  if (e->ev_k3_done) cudaEventDestroy(e->ev_k3_done);
  if (e->ev_fb) cudaEventDestroy(e->ev_fb);
  if (e->ev_wide) cudaEventDestroy(e->ev_wide);
  if (e->sj) cudaStreamDestroy(e->sj);
  if (e->sw) cudaStreamDestroy(e->sw);
  if (e->sg) cudaStreamDestroy(e->sg);
  if (e->h_fb_ring) cudaFreeHost(e->h_fb_ring);
  for (int k = 0; k < 2; k++) { if (e->stage_mem[k]) cudaFreeHost(e->stage_mem[k]); if (e->stage_ev[k]) cudaEventDestroy(e->stage_ev[k]); }
  cudaStreamDestroy(e->stream);
  delete e;
  return MFA_OK;
}

int mfa_engine::join_k3() {
  if (!k3_pending) return MFA_OK;
  CUDA_TRY(cudaStreamWaitEvent(stream, ev_k3_done, 0));
  k3_pending = false;
  return MFA_OK;
}

extern "C" int mfa_engine_sync(mfa_engine *e) {
  if (!e) return set_error(MFA_ERR_INVALID, "null engine");
  MFA_TRY(e->join_k3());
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  CUDA_TRY(cudaGetLastError());
  e->harvest_fallbacks();
  return MFA_OK;
}
extern "C" void *mfa_engine_stream(mfa_engine *e) { return e ? (void *)e->stream : nullptr; }
extern "C" int mfa_engine_sm_count(mfa_engine *e) { return e ? e->sm_count : 0; }
extern "C" int64_t mfa_engine_launch_count(mfa_engine *e) { return e ? e->launches : 0; }
extern "C" int64_t mfa_engine_band_fallbacks(mfa_engine *e) {
  if (!e) return 0;
  if (e->fb_pending) { cudaSetDevice(e->device); e->join_k3(); cudaStreamSynchronize(e->stream); e->harvest_fallbacks(); }   // counts of launches still in flight
  return e->band_fallbacks;
}
namespace {
__global__ void stage_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace
int mfa_engine::stage_copy(void *dst, const void *src_pinned, size_t bytes) {
  // arena chunks are 256-byte aligned and device buffers carry >= 256 bytes of slack, so whole 16-byte words may be moved
  const size_t n16 = (bytes + 15) / 16;
  const unsigned blocks = (unsigned)std::min<size_t>((n16 + 255) / 256, 4 * (size_t)sm_count);
  stage_copy_kernel<<<blocks, 256, 0, stream>>>((const uint4 *)src_pinned, (uint4 *)dst, n16);
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}
int mfa_engine::begin_call() {
  stage_cur ^= 1;
  if (stage_busy[stage_cur]) { CUDA_TRY(cudaEventSynchronize(stage_ev[stage_cur])); stage_busy[stage_cur] = false; }
  stage_used = 0;
  return MFA_OK;
}
int mfa_engine::end_call() {
  if (stage_used > 0) { CUDA_TRY(cudaEventRecord(stage_ev[stage_cur], stream)); stage_busy[stage_cur] = true; }
  stage_used = kStageBytes;   // uploads outside a scope take the synchronous path
  return MFA_OK;
}
int mfa_engine::gmm_timing_begin() {
  if ((size_t)gmm_ev_used + 2 > gmm_ev.size()) {
    for (int k = 0; k < 2; k++) { cudaEvent_t ev; CUDA_TRY(cudaEventCreate(&ev)); gmm_ev.push_back(ev); }
  }
  CUDA_TRY(cudaEventRecord(gmm_ev[gmm_ev_used], stream));
  return MFA_OK;
}
int mfa_engine::gmm_timing_end(int64_t rows) {
  CUDA_TRY(cudaEventRecord(gmm_ev[gmm_ev_used + 1], stream));
  gmm_ev_used += 2; gmm_rows += rows;
  return MFA_OK;
}
extern "C" int mfa_engine_gmm_timing(mfa_engine *e, float *total_ms, int64_t *n_launches, int64_t *n_rows) {
  if (!e) return set_error(MFA_ERR_INVALID, "null engine");
  float tot = 0.0f;
  for (int k = 0; k < e->gmm_ev_used; k += 2) {
    CUDA_TRY(cudaEventSynchronize(e->gmm_ev[k + 1]));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->gmm_ev[k], e->gmm_ev[k + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = e->gmm_ev_used / 2;
  if (n_rows) *n_rows = e->gmm_rows;
  return MFA_OK;
}

int mfa_engine::stage_begin(int stage) {
  if (cfg.trace) {
    static thread_local std::chrono::steady_clock::time_point t0;
    auto now = std::chrono::steady_clock::now();
    if (st_stage.empty()) t0 = now;
    fprintf(stderr, "[trace] host %.3f ms: enqueue stage %d\n", std::chrono::duration<double, std::milli>(now - t0).count(), stage);
  }
  const size_t k = st_stage.size();
  while (st_ev.size() < 2 * (k + 1)) { cudaEvent_t ev; CUDA_TRY(cudaEventCreate(&ev)); st_ev.push_back(ev); }
  st_stage.push_back(stage);
  CUDA_TRY(cudaEventRecord(st_ev[2 * k], stream));
  return MFA_OK;
}
int mfa_engine::stage_end(cudaStream_t on) {
  CUDA_TRY(cudaEventRecord(st_ev[2 * (st_stage.size() - 1) + 1], on ? on : stream));
  return MFA_OK;
}
extern "C" int mfa_engine_stage_timing(mfa_engine *e, float *ms4) {
  if (!e || !ms4) return set_error(MFA_ERR_INVALID, "null argument");
  for (int i = 0; i < mfa_engine::ST_N; i++) ms4[i] = 0.0f;
  for (size_t k = 0; k < e->st_stage.size(); k++) {
    CUDA_TRY(cudaEventSynchronize(e->st_ev[2 * k + 1]));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->st_ev[2 * k], e->st_ev[2 * k + 1]));
    ms4[e->st_stage[k]] += ms;
  }
  return MFA_OK;
}

extern "C" int mfa_engine_gmm_flops(mfa_engine *e, double *useful_flops) {
  if (!e || !useful_flops) return set_error(MFA_ERR_INVALID, "null argument");
  *useful_flops = e->gmm_flops;
  return MFA_OK;
}
extern "C" int mfa_engine_gmm_issued_flops(mfa_engine *e, double *issued_flops) {
  if (!e || !issued_flops) return set_error(MFA_ERR_INVALID, "null argument");
  *issued_flops = e->gmm_issued;
  return MFA_OK;
}

// ------------------------------------------------------------------------------------------------ model
mfa_model::~mfa_model() {
  cudaSetDevice(device);
  for (void *p : {(void *)d_pdf_off, (void *)d_tid2pdf, (void *)d_gconsts, (void *)d_miv, (void *)d_iv, (void *)d_weights, (void *)d_tile_pdf0,
                  (void *)d_tile_seg, (void *)d_W, (void *)d_G, (void *)d_gauss_row, d_tc_w, (void *)d_tc_colscale, (void *)d_acc, d_tc_rows, (void *)d_tc_g,
                  (void *)d_first_tid, (void *)d_self_loop_tid, (void *)d_log_probs, (void *)d_tid_cost, (void *)d_tc_flag,
                  (void *)sp_gconsts, (void *)sp_miv, (void *)sp_iv, (void *)sp_weights, (void *)sp_pdf_off, (void *)acc_spare})
    if (p) cudaFree(p);
}

int mfa_model::acc_take(size_t bytes) {
  if (d_acc && acc_cap_bytes >= bytes) return MFA_OK;
  if (!d_acc && acc_spare && acc_cap_bytes >= bytes) { d_acc = acc_spare; acc_spare = nullptr; return MFA_OK; }
  for (double **p : {&d_acc, &acc_spare}) if (*p) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaFree(*p)); *p = nullptr; }
  acc_cap_bytes = bytes + bytes / 4 + 4096;   // head room: the number of Gaussians moves by a few per cent per iteration
  CUDA_TRY(cudaMalloc((void **)&d_acc, acc_cap_bytes));
  return MFA_OK;
}

// greedy packing of whole pdfs into tiles of MFA_TILE_N Gaussian rows; every pdf's column range is padded to a multiple of
// MFA_SEG_ALIGN columns (padding = gconst -1e30 / zero weights) so the tensor-core epilogue can work on 4-column groups
int mfa_model::layout_tiles() {
  h_tile_pdf0.clear();
  auto padded = [](int ng) { return (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN; };
  int cur = 0, t = -1;
  for (int p = 0; p < num_pdfs; p++) {
    int ng = h_pdf_off[p + 1] - h_pdf_off[p];
    if (ng <= 0) return set_error(MFA_ERR_INVALID, "pdf " + std::to_string(p) + " has no Gaussians");
    if (padded(ng) > MFA_TILE_N) return set_error(MFA_ERR_UNSUPPORTED, "pdf " + std::to_string(p) + " has " + std::to_string(ng) + " Gaussians (> " + std::to_string(MFA_TILE_N) + ")");
    if (t < 0 || cur + padded(ng) > MFA_TILE_N) { t++; cur = 0; h_tile_pdf0.push_back(p); }
    cur += padded(ng);
  }
  n_tiles = t + 1;
  h_tile_pdf0.push_back(num_pdfs);
  kdim = 2 * dim;
  return MFA_OK;
}

int mfa_model::ensure_host() {
  if (!host_stale) return MFA_OK;
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = eng->stream;
  const size_t G = (size_t)num_gauss, D = (size_t)dim;
  h_gconsts.resize(G); h_miv.resize(G * D); h_iv.resize(G * D);
  CUDA_TRY(cudaMemcpyAsync(h_gconsts.data(), d_gconsts, G * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_miv.data(), d_miv, G * D * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_iv.data(), d_iv, G * D * 4, cudaMemcpyDeviceToHost, s));
  if (d_weights) { h_weights.resize(G); CUDA_TRY(cudaMemcpyAsync(h_weights.data(), d_weights, G * 4, cudaMemcpyDeviceToHost, s)); }
  CUDA_TRY(cudaStreamSynchronize(s));
  host_stale = false;
  return MFA_OK;
}

// the fp32 CUDA-core kernel's layout ([tile][k][128] weights, per-tile gconsts and segment starts): only the cross-check kernel and
// models the tcgen05 kernel cannot take need it, so it is built on first use from the host mirrors
int mfa_model::ensure_ffma() {
  if (ffma_ready) return MFA_OK;
  MFA_TRY(ensure_host());
  auto padded = [](int ng) { return (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN; };
  std::vector<float> W((size_t)n_tiles * kdim * MFA_TILE_N, 0.0f), G((size_t)n_tiles * MFA_TILE_N, -1.0e30f);
  h_tile_seg.assign((size_t)n_tiles * (MFA_TILE_N + 1), MFA_TILE_N);
  h_gauss_col.assign(num_gauss, 0);
  std::vector<int32_t> grow(num_gauss);
  for (int tl = 0; tl < n_tiles; tl++) {
    int col = 0;
    int32_t *seg = &h_tile_seg[(size_t)tl * (MFA_TILE_N + 1)];
    for (int p = h_tile_pdf0[tl]; p < h_tile_pdf0[tl + 1]; p++) {
      seg[p - h_tile_pdf0[tl]] = col;
      int c = col;
      for (int m = h_pdf_off[p]; m < h_pdf_off[p + 1]; m++, c++) {
        G[(size_t)tl * MFA_TILE_N + c] = h_gconsts[m];
        grow[m] = tl * MFA_TILE_N + c;
        h_gauss_col[m] = tl * MFA_TILE_N + c;
        for (int d = 0; d < dim; d++) {
          W[((size_t)tl * kdim + d) * MFA_TILE_N + c] = h_miv[(size_t)m * dim + d];
          W[((size_t)tl * kdim + dim + d) * MFA_TILE_N + c] = -0.5f * h_iv[(size_t)m * dim + d];
        }
      }
      col += padded(h_pdf_off[p + 1] - h_pdf_off[p]);
    }
    seg[h_tile_pdf0[tl + 1] - h_tile_pdf0[tl]] = col;  // end of the last pdf (padding included); remaining entries stay TILE_N
  }
  cudaStream_t s = eng->stream;
  auto up = [&](auto **dp, const auto &v) -> int {
    using T = typename std::remove_reference<decltype(v[0])>::type;
    if (*dp) { CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree((void *)*dp)); *dp = nullptr; }
    CUDA_TRY(cudaMalloc((void **)dp, std::max<size_t>(1, v.size()) * sizeof(T)));
    CUDA_TRY(cudaMemcpyAsync((void *)*dp, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return MFA_OK;
  };
  MFA_TRY(up(&d_W, W)); MFA_TRY(up(&d_G, G)); MFA_TRY(up(&d_tile_pdf0, h_tile_pdf0)); MFA_TRY(up(&d_tile_seg, h_tile_seg));
  MFA_TRY(up(&d_gauss_row, grow));
  CUDA_TRY(cudaStreamSynchronize(s));
  ffma_ready = true;
  return MFA_OK;
}

extern "C" int mfa_model_create(mfa_engine *e, const mfa_model_desc *d, mfa_model **out) {
  if (!e || !d || !out) return set_error(MFA_ERR_INVALID, "null argument");
  if (d->dim <= 0 || d->dim > 64) return set_error(MFA_ERR_UNSUPPORTED, "feature dim must be in 1..64");
  if (d->num_pdfs <= 0 || d->pdf_off[d->num_pdfs] != d->num_gauss) return set_error(MFA_ERR_INVALID, "pdf_off inconsistent with num_gauss");
  CUDA_TRY(cudaSetDevice(e->device));
  auto *m = new mfa_model();
  m->eng = e; m->device = e->device; m->dim = d->dim; m->num_pdfs = d->num_pdfs; m->num_gauss = d->num_gauss; m->num_tids = d->num_tids;
  m->h_pdf_off.assign(d->pdf_off, d->pdf_off + d->num_pdfs + 1);
  m->h_gconsts.assign(d->gconsts, d->gconsts + d->num_gauss);
  m->h_miv.assign(d->means_invvars, d->means_invvars + (size_t)d->num_gauss * d->dim);
  m->h_iv.assign(d->inv_vars, d->inv_vars + (size_t)d->num_gauss * d->dim);
  if (d->weights) m->h_weights.assign(d->weights, d->weights + d->num_gauss);
  m->h_tid2pdf.assign(d->tid2pdf, d->tid2pdf + d->num_tids + 1);
  for (int t = 1; t <= d->num_tids; t++)
    if (m->h_tid2pdf[t] < 0 || m->h_tid2pdf[t] >= d->num_pdfs) { delete m; return set_error(MFA_ERR_INVALID, "tid2pdf out of range"); }
  auto alloc_up = [&](auto **dp, const auto &v) -> int {
    using T = typename std::remove_reference<decltype(v[0])>::type;
    CUDA_TRY(cudaMalloc((void **)dp, std::max<size_t>(1, v.size()) * sizeof(T)));
    CUDA_TRY(cudaMemcpyAsync((void *)*dp, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    return MFA_OK;
  };
  int r = alloc_up(&m->d_pdf_off, m->h_pdf_off);
  if (!r) r = alloc_up(&m->d_tid2pdf, m->h_tid2pdf);
  if (!r) r = alloc_up(&m->d_miv, m->h_miv);
  if (!r) r = alloc_up(&m->d_iv, m->h_iv);
  if (!r) r = alloc_up(&m->d_gconsts, m->h_gconsts);
  if (!r && !m->h_weights.empty()) r = alloc_up(&m->d_weights, m->h_weights);
  if (!r) r = m->layout_tiles();
  if (!r) { cudaError_t ce = cudaStreamSynchronize(e->stream); if (ce != cudaSuccess) r = set_error(MFA_ERR_CUDA, cudaGetErrorString(ce)); }
  if (r) { delete m; return r; }
  *out = m;
  return MFA_OK;
}

extern "C" int mfa_model_destroy(mfa_model *m) {
  // the engine may already be gone (Python tears objects down in any order): never touch m->eng here
  if (m) { cudaSetDevice(m->device); cudaDeviceSynchronize(); }
  delete m;
  return MFA_OK;
}

// GmmAligner.boost_silence -> Kaldi gmm-boost-silence: the weights of the given pdfs are scaled WITHOUT renormalisation, i.e.
// gconst += log(factor)
extern "C" int mfa_model_boost_pdfs(mfa_model *m, float factor, const int32_t *pdfs, int32_t n) {
  if (!m || (n > 0 && !pdfs) || !(factor > 0.0f)) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(m->eng->device));
  MFA_TRY(m->ensure_host());
  float lb = logf(factor);
  for (int i = 0; i < n; i++) {
    int p = pdfs[i];
    if (p < 0 || p >= m->num_pdfs) return set_error(MFA_ERR_INVALID, "pdf id out of range");
    for (int g = m->h_pdf_off[p]; g < m->h_pdf_off[p + 1]; g++) { m->h_gconsts[g] += lb; if (!m->h_weights.empty()) m->h_weights[g] *= factor; }
  }
  cudaStream_t s = m->eng->stream;
  CUDA_TRY(cudaMemcpyAsync(m->d_gconsts, m->h_gconsts.data(), m->h_gconsts.size() * 4, cudaMemcpyHostToDevice, s));
  if (m->d_weights) CUDA_TRY(cudaMemcpyAsync(m->d_weights, m->h_weights.data(), m->h_weights.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  m->tc_ready = false;
  m->ffma_ready = false;
  return MFA_OK;
}

// ------------------------------------------------------------------------------------------------ graphs
namespace mfa {
struct CachedBlock { int device; void *p; size_t cap; };
static std::mutex g_cache_mu;
static std::vector<CachedBlock> g_cache;
static constexpr size_t kCacheBlocks = 6;
void *dev_cache_take(int device, size_t bytes, size_t *cap) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  int best = -1;
  for (int i = 0; i < (int)g_cache.size(); i++)
    if (g_cache[i].device == device && g_cache[i].cap >= bytes && g_cache[i].cap <= 2 * bytes + (1 << 20) && (best < 0 || g_cache[i].cap < g_cache[best].cap)) best = i;
  if (best < 0) return nullptr;
  void *p = g_cache[best].p;
  *cap = g_cache[best].cap;
  g_cache.erase(g_cache.begin() + best);
  return p;
}
void dev_cache_give(int device, void *p, size_t cap) {
  if (!p) return;
  void *victim = nullptr; int vdev = 0;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.push_back({device, p, cap});
    if (g_cache.size() > kCacheBlocks) {   // drop the smallest
      int k = 0;
      for (int i = 1; i < (int)g_cache.size(); i++) if (g_cache[i].cap < g_cache[k].cap) k = i;
      victim = g_cache[k].p; vdev = g_cache[k].device;
      g_cache.erase(g_cache.begin() + k);
    }
  }
  if (victim) { cudaSetDevice(vdev); cudaFree(victim); }
}
}  // namespace mfa

mfa_graphs::~mfa_graphs() {
  // no cudaFree (it synchronises the device and has a long latency tail): the blocks are parked for the next batch's graphs
  if (d_blob) mfa::dev_cache_give(device, d_blob, d_blob_cap);
  if (d_rag) mfa::dev_cache_give(device, d_rag, rag_meta_bytes);
}

namespace mfa {
int upload_graphs(mfa_engine *e, mfa_graphs *g) {
  if (g->d_blob && g->device == e->device) return MFA_OK;
  if (g->host_w_stale) return set_error(MFA_ERR_UNSUPPORTED, "these graphs had their transition costs re-folded on another device; pack them again for this one");
  if (g->d_blob) { dev_cache_give(g->device, g->d_blob, g->d_blob_cap); g->d_blob = nullptr; }
  size_t A = g->a_src.size();
  const auto &barc = g->h_barc; const auto &pack = g->h_apack;   // built by mfa_graphs_pack on its worker threads
  if (barc.size() != 2 * A || pack.size() != A) return set_error(MFA_ERR_INVALID, "graphs object without device images");
  for (int u = 0; u < g->n_utts; u++)
    if (g->lp_off[u + 1] - g->lp_off[u] >= 0xFFFF) return set_error(MFA_ERR_UNSUPPORTED, "utterance graph references >= 65535 pdfs");
  struct Item { const void *h; size_t bytes; void **d; };
  std::vector<Item> items = {
      {g->st_off.data(), g->st_off.size() * 8, (void **)&g->d_st_off}, {g->arc_off.data(), g->arc_off.size() * 8, (void **)&g->d_arc_off},
      {g->lp_off.data(), g->lp_off.size() * 8, (void **)&g->d_lp_off}, {g->inb_off.data(), g->inb_off.size() * 8, (void **)&g->d_inb_off},
      {g->start.data(), g->start.size() * 4, (void **)&g->d_start}, {g->n_eps.data(), g->n_eps.size() * 4, (void **)&g->d_n_eps},
      {g->in_begin.data(), g->in_begin.size() * 4, (void **)&g->d_in_begin}, {g->a_tid.data(), A * 4, (void **)&g->d_a_tid},
      {g->a_olabel.data(), A * 4, (void **)&g->d_a_olabel}, {g->lp2pdf.data(), g->lp2pdf.size() * 4, (void **)&g->d_lp2pdf},
      {pack.data(), A * 4, (void **)&g->d_a_pack}, {g->a_w.data(), A * 4, (void **)&g->d_a_w},
      {g->final_w.data(), g->final_w.size() * 4, (void **)&g->d_final_w}, {g->a_src.data(), A * 4, (void **)&g->d_a_src},
      {g->b_start.data(), g->b_start.size() * 4, (void **)&g->d_b_start}, {g->b_maxback.data(), g->b_maxback.size() * 4, (void **)&g->d_b_maxback},
      {g->b_stw.data(), g->b_stw.size() * 4, (void **)&g->d_b_stw}, {barc.data(), barc.size() * 4, (void **)&g->d_b_arc},
      {g->b_fin.data(), g->b_fin.size() * 4, (void **)&g->d_b_fin},
      {g->b_arcid.data(), g->b_arcid.size() * 2, (void **)&g->d_b_arcid}, {g->b_orig.data(), g->b_orig.size() * 2, (void **)&g->d_b_orig},
      {g->a_w0.data(), g->a_w0.size() * 4, (void **)&g->d_a_w0}};
  size_t total = 0;
  for (auto &it : items) total += (it.bytes + 255) / 256 * 256;
  const size_t want = std::max<size_t>(total, 256);
  g->d_blob = dev_cache_take(e->device, want, &g->d_blob_cap);
  if (g->d_blob) CUDA_TRY(cudaDeviceSynchronize());   // the block's previous user may still have kernels in flight
  else { g->d_blob_cap = want + want / 8; CUDA_TRY(cudaMalloc(&g->d_blob, g->d_blob_cap)); }
  g->d_bytes = total; g->device = e->device;
  size_t off = 0;
  for (auto &it : items) {
    *it.d = (char *)g->d_blob + off;
    if (it.bytes) CUDA_TRY(cudaMemcpyAsync(*it.d, it.h, it.bytes, cudaMemcpyHostToDevice, e->stream));
    off += (it.bytes + 255) / 256 * 256;
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));  // the host arrays are pageable: the copies are staged, not asynchronous
  return MFA_OK;
}
}  // namespace mfa

namespace {
// one CTA per utterance: by-source arc weights from the unfolded weights + the per-tid cost, then the band copy (arcs grouped by
// destination; b_arcid = utterance-local index of the same arc in by-source order)
__global__ void refold_kernel(const int64_t *__restrict__ arc_off, const int32_t *__restrict__ a_tid, const float *__restrict__ a_w0,
                              const float *__restrict__ tid_cost, const uint16_t *__restrict__ b_arcid, float *__restrict__ a_w, uint2 *__restrict__ b_arc) {
  const int64_t a0 = arc_off[blockIdx.x], A = arc_off[blockIdx.x + 1] - a0;
  for (int64_t k = threadIdx.x; k < A; k += blockDim.x) {
    const int tid = a_tid[a0 + k];
    a_w[a0 + k] = tid > 0 ? a_w0[a0 + k] + tid_cost[tid] : a_w0[a0 + k];
  }
  __syncthreads();
  for (int64_t j = threadIdx.x; j < A; j += blockDim.x) b_arc[a0 + j].y = __float_as_uint(a_w[a0 + b_arcid[a0 + j]]);
}
}  // namespace

namespace mfa {
int refold_graphs(mfa_engine *e, mfa_graphs *g, const float *d_tid_cost) {
  if (g->n_utts == 0) return MFA_OK;
  refold_kernel<<<(unsigned)g->n_utts, 256, 0, e->stream>>>(g->d_arc_off, g->d_a_tid, g->d_a_w0, d_tid_cost, g->d_b_arcid, g->d_a_w, (uint2 *)g->d_b_arc);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  g->host_w_stale = true;
  return MFA_OK;
}
}  // namespace mfa

extern "C" int mfa_graphs_destroy(mfa_graphs *g) { delete g; return MFA_OK; }
