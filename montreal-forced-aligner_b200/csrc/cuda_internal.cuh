// cuda_internal.cuh -- engine / model structures for the CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "internal.h"

#define CUDA_TRY(x)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (x);                                                                            \
    if (e__ != cudaSuccess) return mfa::set_error(MFA_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define MFA_TRY(x)            \
  do {                        \
    int r__ = (x);            \
    if (r__ != MFA_OK) return r__; \
  } while (0)

enum DevBuf {
  DB_PCM = 0, DB_SAMPLE_OFF, DB_FRAME_OFF, DB_ROW_OFF, DB_UTT2SPK, DB_MFCC, DB_FEATS, DB_CMVN_PART, DB_CMVN_STATS,
  DB_SPK_UTT_OFF, DB_SPK_UTTS, DB_LDA, DB_FMLLR, DB_LL, DB_BP, DB_ALI, DB_PERFRAME, DB_WORDS, DB_WORD_OFF, DB_NUM_WORDS,
  DB_TOTAL_LIKE, DB_STATUS, DB_COL_OFF, DB_TILE_OFF, DB_BP_OFF, DB_UTT_ORDER, DB_IO_FEATS, DB_IO_LL, DB_IO_ALI, DB_XSPLIT,
  DB_CHUNK_FRAME_OFF, DB_SCRATCH, DB_TC_ITEMS, DB_BIMG, DB_TILE_ROW0, DB_TILE_ROWS, DB_LL_OFF, DB_LD_U, DB_MFCC_TAB, DB_BBP, DB_BBP_OFF, DB_BORDER, DB_FALLBACK, DB_ACC_INT, DB_ACC_ORDER, DB_FM_STATS, DB_FM_AUX, DB_FM_INVG, DB_FM_W, DB_FM_OUT, DB_FB_BP, DB_FB_BIG, DB_MLE, DB_RAG_CNT, DB_WIDE_BP, DB_FALLBACK2, DB_K3_OFFS, DB_FB_CTL, DB_TC_CTR, DB_N
};
enum PinBuf { PB_A = 0, PB_B, PB_C, PB_D, PB_E, PB_N };

// Experiment / test switches.  Read from the environment ONCE, when the engine is created (MFA_<UPPER-CASE NAME>), and changed
// afterwards only through mfa_engine_set_option; no kernel launcher consults the environment.
struct mfa_engine_cfg {
  int vit_band = 1;            // 0: every utterance on the sparse Viterbi kernel
  int vit_maxgroups = 8;       // band width in groups of 32 states (1..8); tests narrow it to exercise the fallback
  int vit_wide = 1;            // 1: utterances that outgrow the band run on the 32-group band kernel before the sparse kernel is asked
  int vit_wide_poll = -1;      // wide level: 1 = its CTAs poll the overflow list next to the primary launches, 0 = it runs after their join; -1 = poll in device-buffer calls only
  int vit_wide_ctas = 8;       // CTAs of the wide level
  int vit_graph_smem = 0;      // 1: band kernel copies each graph to shared memory (default: read through L1)
  int vit_nw2_kb = -1;         // size classes up to this many KB of shared memory run 2 warps per utterance (-1: 20, or 44 with graph_smem)
  int vit_carveout = 100;      // shared-memory carve-out (percent) of the sparse kernel
  int vit_carveout_band = -1;  // ... of the band kernel (-1: 70, or 100 with graph_smem)
  int k3_overlap = 1;          // 1: device-buffer mfa_align_pcm leaves its Viterbi launch un-joined so that the next call's K1 / features overlap its tail
  int vit_prio = 1;            // side-stream priorities rise with the size class (creation time only)
  int pipeline_split = -1;     // host-PCM path: -1 auto, 0 whole, 1 every stage per segment, 2 stream (K1..K2 per segment, one K3)
  int acc_impl = 0;            // K4: 0 counting sort + register accumulation, 1 first version (f64 atomics)
  int tc_k96 = 0;              // K2: 1 forces the K = 96 operand geometry
  int tc_poly = 0;             // K2 epilogue: 1 = every fourth exp2 runs as an FMA-pipe polynomial instead of MUFU.EX2 (A/B switch)
  int mfcc_generic = 0;        // K1: 1 forces the generic (shared-memory FFT) kernel
  int trace = 0;               // print host enqueue times per stage
};

struct mfa_engine {
  mfa_engine_cfg cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  size_t smem_optin = 0;
  int64_t launches = 0;
  int64_t band_fallbacks = 0;          // utterances the band Viterbi kernel handed to the sparse kernel (cumulative, harvested counts)
  // every Viterbi launch's fallback count lands in its own slot of a pinned, device-mapped ring (written by viterbi_fallback_kernel);
  // the slots are summed into band_fallbacks whenever the stream has been synchronised anyway -- never inside a step
  static constexpr int kFbRing = 256;
  int32_t *h_fb_ring = nullptr;
  int fb_pending = 0;
  void harvest_fallbacks() { for (int i = 0; i < fb_pending; i++) band_fallbacks += h_fb_ring[i]; fb_pending = 0; }
  // K2 timing: event pairs recorded around each K2 launch of the current API call
  std::vector<cudaEvent_t> gmm_ev;
  int gmm_ev_used = 0;
  int64_t gmm_rows = 0;
  double gmm_flops = 0.0;              // useful FLOPs (2*(2D+1) per frame x Gaussian actually scored) of the K2 launches timed
  double gmm_issued = 0.0;             // FLOPs those launches issued on the tensor pipe (padding included)
  void gmm_timing_reset() { gmm_ev_used = 0; gmm_rows = 0; gmm_flops = 0.0; gmm_issued = 0.0; }
  int gmm_timing_begin();
  int gmm_timing_end(int64_t rows);
  // side streams + events: the Viterbi size classes run concurrently (fork/join around the main stream)
  static constexpr int kSide = 12;
  cudaStream_t side[kSide] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {};
  // per-utterance B images gathered ahead of time on a side stream (gmm_tc.cu prefetch_b_images): valid for utterances [pf_u0, pf_u1)
  // of graphs pf_g under model pf_m until the next ragged launch consumes or replaces them
  // K3 runs on the side streams and is joined on `sj`, NOT on the main stream: the next call's K1 / feature kernels (which share no
  // buffer with it) start while the Viterbi tail is still running; whoever touches something K3 reads or writes calls join_k3() first
  // (every entry point does at its start, except the fused alignment, which defers it to just before its own K2).
  cudaStream_t sj = nullptr, sg = nullptr, sw = nullptr;      // join stream of K3; gather stream of the B-image prefetch
  cudaEvent_t ev_k3_done = nullptr, ev_fb = nullptr, ev_wide = nullptr;
  bool k3_pending = false;
  const char *pend_out[6] = {}; size_t pend_bytes[6] = {};   // output buffers of the K3 in flight
  int join_k3();
  bool k3_writes(const void *p, size_t bytes) const {
    if (!k3_pending) return false;
    for (int i = 0; i < 6; i++)
      if (pend_out[i] && (const char *)p < pend_out[i] + pend_bytes[i] && pend_out[i] < (const char *)p + bytes) return true;
    return false;
  }
  cudaEvent_t ev_bimg = nullptr;
  const void *pf_g = nullptr, *pf_m = nullptr;
  int pf_u0 = 0, pf_u1 = 0;
  std::vector<cudaEvent_t> ev_piece;   // H2D pieces of the PCM upload (end-to-end path)
  // per-stage CUDA-event timing of the last fused call: intervals (begin event, end event, stage) on the main stream
  enum Stage { ST_MFCC = 0, ST_FEAT = 1, ST_GMM = 2, ST_VITERBI = 3, ST_N = 4 };
  std::vector<cudaEvent_t> st_ev;
  std::vector<int> st_stage;           // stage of interval k (events 2k, 2k+1)
  int stage_begin(int stage);
  int stage_end(cudaStream_t on = nullptr);   // `on`: record the end of the interval on another stream (K3: its join stream)
  void stage_reset() { st_stage.clear(); }
  // cached MFCC tables (device blob in DB_MFCC_TAB) for the last option set
  bool mfcc_tab_valid = false;
  mfa_mfcc_opts mfcc_tab_opts{};
  alignas(8) unsigned char mfcc_tab_desc[128] = {};
  struct Buf { void *p = nullptr; size_t cap = 0; };
  Buf dev[DB_N];
  Buf pin[PB_N];
  // grow-only device buffer; contents are NOT preserved on growth
  int get(int id, size_t bytes, void **out);
  int get_pinned(int id, size_t bytes, void **out);
  template <typename T> int getT(int id, size_t n, T **out) { void *p; int r = get(id, n * sizeof(T), &p); *out = (T *)p; return r; }
  // Pinned staging for the small host->device uploads of an API call (offsets, plans, work lists): the bytes are copied into an
  // engine-owned pinned arena first, so the cudaMemcpyAsync neither blocks the host nor drains the stream (a copy from pageable
  // memory does both) and the caller's vector may die immediately.  Two arenas serve alternate calls; an arena is reused once
  // the call that filled it has drained (event).  Entry points open a scope with CallScope.
  static constexpr size_t kStageBytes = (size_t)32 << 20;
  void *stage_mem[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  bool stage_busy[2] = {false, false};
  size_t stage_used = kStageBytes;     // full until the first begin_call
  int stage_cur = 0;
  int begin_call();
  int end_call();
  // moves staged bytes with a small SM kernel that reads the pinned arena directly: a cudaMemcpyAsync would queue on the copy
  // engine BEHIND bulk transfers already enqueued there (the PCM pieces of the end-to-end path: ~20 ms), stalling the stream
  int stage_copy(void *dst, const void *src_pinned, size_t bytes);
  void *stage_alloc(size_t bytes) {
    const size_t off = (stage_used + 255) & ~(size_t)255;
    if (!stage_mem[stage_cur] || off + bytes > kStageBytes) return nullptr;
    stage_used = off + bytes;
    return (char *)stage_mem[stage_cur] + off;
  }
  // upload a host array into a device buffer slot (stream-ordered; `h` may be freed on return)
  template <typename T> int upload(int id, const T *h, size_t n, T **out) {
    int r = getT<T>(id, n ? n : 1, out);
    if (r) return r;
    if (n) {
      if (void *st = stage_alloc(n * sizeof(T))) {
        memcpy(st, h, n * sizeof(T));
        int r2 = stage_copy(*out, st, n * sizeof(T));
        if (r2) return r2;
      } else {
        CUDA_TRY(cudaMemcpyAsync(*out, h, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
      }
    }
    return MFA_OK;
  }
};

struct CallScope {   // one per C-ABI entry point that uploads: recycles the staging arena; joins a K3 still in flight unless told to defer
  mfa_engine *e;
  explicit CallScope(mfa_engine *e_, bool defer_join = false) : e(e_) { if (!defer_join) e->join_k3(); e->begin_call(); }
  ~CallScope() { e->end_call(); }
};

// Tiled acoustic model on the device.  Gaussians are regrouped into tiles of TILE_N rows that never split a pdf
// (padding rows have gconst = -1e30 and zero weights), so a tile's per-pdf log-sum-exp is local to the tile.
#define MFA_TILE_N 128
#define MFA_SEG_ALIGN 4

struct mfa_model {
  mfa_engine *eng = nullptr;
  int device = 0;                      // copied from the engine at creation: the destructor must not touch `eng` (it may be gone)
  int dim = 0, num_pdfs = 0, num_gauss = 0, num_tids = 0;
  std::vector<int32_t> h_pdf_off, h_tid2pdf;   // h_pdf_off is ALWAYS current (the host plans K2 tiles from it)
  std::vector<float> h_gconsts, h_miv, h_iv, h_weights;   // host mirrors of the parameters; stale after a device M-step until ensure_host()
  bool host_stale = false;
  int ensure_host();                   // refresh the host mirrors from the device arrays
  int32_t *d_pdf_off = nullptr, *d_tid2pdf = nullptr;
  float *d_gconsts = nullptr, *d_miv = nullptr, *d_iv = nullptr;  // natural layout (K4, K5, the M-step, source of the K2 operand images)
  float *d_weights = nullptr;          // [num_gauss] mixture weights (nullptr when the model was created without them)
  // transition model tables for on-device training (mfa_model_set_transitions)
  int num_tstates = 0;
  int32_t *d_first_tid = nullptr, *d_self_loop_tid = nullptr;
  float *d_log_probs = nullptr, *d_tid_cost = nullptr;
  // tiled layout (K2)
  int n_tiles = 0, kdim = 0;           // kdim = 2*dim
  std::vector<int32_t> h_tile_pdf0;    // [n_tiles+1] first pdf of each tile
  std::vector<int32_t> h_tile_seg;     // [n_tiles][TILE_N+1] column where local pdf k starts; padded with TILE_N..
  std::vector<int32_t> h_gauss_col;    // [num_gauss] tile*TILE_N + column of natural Gaussian m
  int32_t *d_tile_pdf0 = nullptr, *d_tile_seg = nullptr;
  float *d_W = nullptr;                // [n_tiles][kdim][TILE_N]  (k-major: coalesced / conflict-free tile loads)
  float *d_G = nullptr;                // [n_tiles][TILE_N] gconsts (-1e30 padding)
  int32_t *d_gauss_row = nullptr;      // [num_gauss] tile row (tile*TILE_N + col) of natural Gaussian m
  bool ffma_ready = false;             // d_W / d_G / d_tile_* (the fp32 CUDA-core kernel's layout) are built on first use
  int layout_tiles();                  // host: n_tiles / h_tile_pdf0 from h_pdf_off
  int ensure_ffma();
  // tensor-core operand images (gmm_tc.cu); built lazily
  void *d_tc_w = nullptr;
  size_t tc_w_bytes = 0;
  float tc_scale_shift = 0.0f;
  std::vector<float> h_tc_colscale;    // per-dimension power-of-two feature scaling folded into the weights
  float *d_tc_colscale = nullptr;
  bool tc_ready = false;
  int tc_n_tiles = 0;                  // tiles of the dense width-class tiling (gmm_tc.cu)
  bool tc_unsupported = false;         // weights beyond the fp16 range / dim too large: the fp32 kernel scores this model
  size_t tc_cap_gauss = 0;             // Gaussians d_tc_rows / d_tc_g were allocated for
  int32_t *d_tc_flag = nullptr;        // device flag: a weight exceeded the fp16 range while the operand rows were written
  void *d_tc_rows = nullptr;           // fp16 hi/lo weight rows [2][G][tc_k] (row-major; source of the per-utterance tile gather)
  int tc_k = 96;                       // K extent of the operand images: 80 (gconst added by the epilogue) or 96 (gconst as fp16 columns)
  float *d_tc_g = nullptr;             // gconst * log2(e): [G] per Gaussian, then [n_tiles][128] per column of the dense tiling
  uint64_t tc_version = 0;             // hash of the pdf -> Gaussian layout and the operand geometry: the key of the cached ragged plans
  double *d_acc = nullptr;
  // Steady-state training must not call cudaMalloc / cudaFree: on this platform a tenth of those calls take 30-400 ms (tools/
  // mstep_latency.py).  The M-step writes into a SPARE set of parameter arrays and swaps; the accumulator block and the K2 images keep
  // their allocations; everything grows by capacity only.
  size_t par_cap = 0;                  // Gaussians the CURRENT d_gconsts / d_miv / d_iv / d_weights were allocated for
  float *sp_gconsts = nullptr, *sp_miv = nullptr, *sp_iv = nullptr, *sp_weights = nullptr; int32_t *sp_pdf_off = nullptr;
  size_t sp_cap = 0;                   // Gaussians the spare set holds
  double *acc_spare = nullptr;         // a consumed accumulator block waiting for the next mfa_acc_zero / mfa_acc_write
  size_t acc_cap_bytes = 0;            // bytes of the block d_acc or acc_spare points to
  size_t tc_w_cap = 0;                 // bytes d_tc_w was allocated with
  int acc_take(size_t bytes);          // make d_acc a block of at least `bytes` (reusing the spare)
  ~mfa_model();
};

namespace mfa {
int upload_graphs(mfa_engine *e, mfa_graphs *g);
// kernels' host launchers (device pointers only)
int launch_mfcc(mfa_engine *e, const mfa_mfcc_opts *o, const int16_t *d_pcm, const int64_t *d_sample_off, int32_t n_utts,
                const int64_t *d_frame_off, int64_t n_frames, float *d_out, int64_t frame_base = 0);
int launch_cmvn_stats(mfa_engine *e, const float *d_feats, int dim, const int64_t *d_frame_off, const int32_t *h_utt2spk,
                      int32_t n_utts, int32_t n_spk, double *d_stats);
int launch_features(mfa_engine *e, const mfa_feat_opts *o, const float *d_in, const int64_t *d_frame_off, const int64_t *h_frame_off,
                    const int64_t *d_row_off, const int32_t *d_utt2spk, int32_t n_utts, const double *d_cmvn_stats, float *d_out,
                    int out_ld);
int launch_gmm_ffma(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld);
int launch_gmm_tc(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld);
bool gmm_tc_supported(mfa_model *m);
int build_tc_device(mfa_model *m, bool layout_changed);        // K2 operand images from the device-resident natural-layout parameters
int refold_graphs(mfa_engine *e, mfa_graphs *g, const float *d_tid_cost);   // a_w = a_w0 + tid_cost[a_tid] on the device, band copy included
int prefetch_b_images(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts);   // gather_b on a side stream, ahead of the features
int launch_gmm_tc_ragged(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts, const float *d_feats, const int64_t *h_row_off,
                         const int64_t *h_frame_off, float *d_out, const int64_t *h_ll_off, const int64_t *h_ld);
int launch_transpose(mfa_engine *e, const float *d_in, int64_t rows, int64_t cols, int64_t in_ld, float *d_out, int64_t out_ld);
struct ViterbiArgs {
  const mfa_graphs *g; int32_t utt0, n_utts;  // utterances [utt0, utt0+n_utts) of the graph batch
  const float *d_llT; int64_t ld;              // pdf-major log-likelihoods [num_pdfs][ld]
  const int64_t *d_col_off;                    // [n_utts] column of each utterance's frame 0 in d_llT (multiple of 8)
  const int64_t *d_ll_off, *d_ld_u;            // ragged mode (both non-null): per-utterance block offset / leading dim; rows = local pdfs
  const int64_t *d_frame_off;                  // [n_utts+1] output frame offsets (relative to d_ali / d_per_frame)
  const int64_t *h_frame_off; const int64_t *h_col_off;
  int32_t *d_ali; float *d_per_frame; int32_t *d_words; const int64_t *d_word_off; int32_t *d_num_words;
  float *d_total_like; int32_t *d_status;
  mfa_align_opts opts;
  bool host_call = false;                      // the caller's buffers are host memory (the call ends with a join and device -> host copies)
};
int launch_viterbi(mfa_engine *e, const ViterbiArgs &a);
size_t viterbi_band_smem(int64_t S, int64_t A, int64_t P, bool graph_in_smem, bool wide = false);   // shared memory the band kernel needs for one utterance
int launch_viterbi_band_wide(mfa_engine *e, const ViterbiArgs &a, const std::vector<int32_t> &subset, const int32_t *d_fb, int32_t *d_fb2,
                             int32_t *h_count, int32_t *d_ctl);
int launch_viterbi_band(mfa_engine *e, const ViterbiArgs &a, const std::vector<int32_t> &subset, int max_groups, int32_t *d_fallback, int32_t *d_ctl);
int launch_acc_stats(mfa_engine *e, mfa_model *m, const float *d_feats, const int32_t *d_ali, int64_t n_frames);
int launch_fmllr_acc(mfa_engine *e, mfa_model *mp, mfa_model *ms, const float *d_feats, const int32_t *d_ali, const float *d_tid_weight,
                     const int64_t *d_frame_off, const int64_t *h_frame_off, const int32_t *h_utt2spk, int32_t n_utts, int32_t n_spk,
                     double *d_stats);
int launch_fmllr_update(mfa_engine *e, const double *d_stats, int dim, int32_t n_spk, int num_iters, double min_count, float *d_W,
                        double *d_impr, double *d_count);
}  // namespace mfa
