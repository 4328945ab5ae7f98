// pipeline.cu -- C-ABI entry points (host/device buffer handling) and the fused PCM -> alignment pipeline.
//
// mfa_align_pcm is AlignFunction._run's per-job loop (reference: montreal_forced_aligner/alignment/
// multiprocessing.py:791-863) plus the feature stages feeding it (corpus/features.py:193-251, 287-376) for one GPU:
// MFCC -> per-speaker CMVN -> deltas | splice+LDA (+fMLLR) -> all-pdf log-likelihoods -> beam Viterbi, in utterance
// chunks sized so the pdf-major log-likelihood block and the back-pointers fit the HBM workspace.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// copy caller data to the device if it is host-resident; returns the device pointer to use
template <typename T>
int to_device(mfa_engine *e, int slot, const T *p, size_t n, int where, const T **out) {
  if (where == MFA_DEVICE) { *out = p; return MFA_OK; }
  T *d;
  MFA_TRY(e->getT<T>(slot, n ? n : 1, &d));
  if (n) CUDA_TRY(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
  *out = d;
  return MFA_OK;
}
template <typename T>
int out_buffer(mfa_engine *e, int slot, T *p, size_t n, int where, T **out) {
  if (where == MFA_DEVICE) { *out = p; return MFA_OK; }
  return e->getT<T>(slot, n ? n : 1, out);
}
template <typename T>
int from_device(mfa_engine *e, const T *d, T *p, size_t n, int where) {
  if (where == MFA_DEVICE || n == 0) return MFA_OK;
  CUDA_TRY(cudaMemcpyAsync(p, d, n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
  return MFA_OK;
}

int check_offsets(const int64_t *off, int n, const char *what) {
  if (!off) return set_error(MFA_ERR_INVALID, std::string(what) + " is null");
  for (int i = 0; i < n; i++) if (off[i + 1] < off[i]) return set_error(MFA_ERR_INVALID, std::string(what) + " is not non-decreasing");
  return MFA_OK;
}

int run_gmm(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t rows, float *d_llT, int64_t ld, int impl) {
  MFA_TRY(e->gmm_timing_begin());
  int r;
  if (impl == 1) r = launch_gmm_ffma(e, m, d_feats, rows, d_llT, ld);
  else r = launch_gmm_tc(e, m, d_feats, rows, d_llT, ld);
  if (r) return r;
  return e->gmm_timing_end(rows);
}

}  // namespace

extern "C" {

int mfa_mfcc(mfa_engine *e, const mfa_mfcc_opts *o, const int16_t *pcm, const int64_t *sample_off, int32_t n_utts, const int64_t *frame_off,
             float *out, int where) {
  if (!e || !o || !sample_off || !frame_off || n_utts < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(sample_off, n_utts, "sample_off"));
  for (int u = 0; u < n_utts; u++)
    if (frame_off[u + 1] - frame_off[u] != mfa_mfcc_num_frames(o, sample_off[u + 1] - sample_off[u]))
      return set_error(MFA_ERR_INVALID, "frame_off does not match mfa_mfcc_num_frames for utterance " + std::to_string(u));
  int64_t ns = sample_off[n_utts] - sample_off[0], nf = frame_off[n_utts];
  int64_t *d_so, *d_fo;
  MFA_TRY(e->upload(DB_SAMPLE_OFF, sample_off, (size_t)n_utts + 1, &d_so));
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  const int16_t *d_pcm; float *d_out;
  MFA_TRY(to_device(e, DB_PCM, pcm, (size_t)(sample_off[0] + ns), where, &d_pcm));
  MFA_TRY(out_buffer(e, DB_MFCC, out, (size_t)nf * o->num_ceps, where, &d_out));
  MFA_TRY(launch_mfcc(e, o, d_pcm, d_so, n_utts, d_fo, nf, d_out));
  MFA_TRY(from_device(e, d_out, out, (size_t)nf * o->num_ceps, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_cmvn_stats(mfa_engine *e, const float *feats, int32_t dim, const int64_t *frame_off, const int32_t *utt2spk, int32_t n_utts,
                   int32_t n_spk, double *stats, int where) {
  if (!e || !frame_off || !utt2spk || !stats) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(frame_off, n_utts, "frame_off"));
  int64_t *d_fo; const float *d_feats; double *d_stats;
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  MFA_TRY(to_device(e, DB_IO_FEATS, feats, (size_t)frame_off[n_utts] * dim, where, &d_feats));
  size_t ns = (size_t)n_spk * 2 * (dim + 1);
  MFA_TRY(out_buffer(e, DB_CMVN_STATS, stats, ns, where, &d_stats));
  CUDA_TRY(cudaMemsetAsync(d_stats, 0, ns * sizeof(double), e->stream));
  MFA_TRY(launch_cmvn_stats(e, d_feats, dim, d_fo, utt2spk, n_utts, n_spk, d_stats));
  MFA_TRY(from_device(e, d_stats, stats, ns, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_features(mfa_engine *e, const mfa_feat_opts *o, const float *in, const int64_t *frame_off, const int32_t *utt2spk, int32_t n_utts,
                 float *out, int where) {
  if (!e || !o || !frame_off) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(frame_off, n_utts, "frame_off"));
  if ((o->fmllr || o->cmvn_stats) && !utt2spk) return set_error(MFA_ERR_INVALID, "utt2spk required with fmllr / cmvn_stats");
  int64_t nf = frame_off[n_utts];
  int od = mfa_feat_out_dim(o);
  int64_t *d_fo; int32_t *d_u2s = nullptr; const float *d_in; float *d_out; double *d_stats = nullptr;
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  if (utt2spk) MFA_TRY(e->upload(DB_UTT2SPK, utt2spk, (size_t)n_utts, &d_u2s));
  if (o->cmvn_stats) MFA_TRY(e->upload(DB_CMVN_STATS, o->cmvn_stats, (size_t)o->n_spk * 2 * (o->in_dim + 1), &d_stats));
  MFA_TRY(to_device(e, DB_IO_FEATS, in, (size_t)nf * o->in_dim, where, &d_in));
  MFA_TRY(out_buffer(e, DB_FEATS, out, (size_t)nf * od, where, &d_out));
  if ((const void *)d_in == (const void *)d_out && !(o->mode == 0 && !o->fmllr)) return set_error(MFA_ERR_INVALID, "in-place only for mode 0 without fMLLR");
  MFA_TRY(launch_features(e, o, d_in, d_fo, frame_off, d_fo, d_u2s, n_utts, d_stats, d_out, od));
  MFA_TRY(from_device(e, d_out, out, (size_t)nf * od, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_cmvn_apply(mfa_engine *e, float *feats, int32_t dim, const int64_t *frame_off, const int32_t *utt2spk, int32_t n_utts, int32_t n_spk,
                   const double *stats, int where) {
  mfa_feat_opts o{};
  o.mode = 0; o.in_dim = dim; o.n_spk = n_spk; o.cmvn_stats = stats;
  if (where == MFA_DEVICE) return mfa_features(e, &o, feats, frame_off, utt2spk, n_utts, feats, where);
  return mfa_features(e, &o, feats, frame_off, utt2spk, n_utts, feats, MFA_HOST);
}

int mfa_gmm_loglikes(mfa_engine *e, mfa_model *m, const float *feats, int64_t n_frames, float *out, int where, int impl) {
  if (!e || !m || n_frames < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  e->gmm_timing_reset();
  const float *d_feats; float *d_out;
  MFA_TRY(to_device(e, DB_IO_FEATS, feats, (size_t)n_frames * m->dim, where, &d_feats));
  MFA_TRY(out_buffer(e, DB_IO_LL, out, (size_t)n_frames * m->num_pdfs, where, &d_out));
  const int64_t chunk = 65536;
  float *d_llT;
  MFA_TRY(e->getT<float>(DB_LL, (size_t)m->num_pdfs * std::min<int64_t>(chunk, round_up(std::max<int64_t>(n_frames, 1), 128)), &d_llT));
  for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
    int64_t n = std::min(chunk, n_frames - f0), ld = round_up(n, 128);
    MFA_TRY(run_gmm(e, m, d_feats + f0 * m->dim, n, d_llT, ld, impl));
    MFA_TRY(launch_transpose(e, d_llT, m->num_pdfs, n, ld, d_out + f0 * m->num_pdfs, m->num_pdfs));
  }
  MFA_TRY(from_device(e, d_out, out, (size_t)n_frames * m->num_pdfs, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

}  // extern "C"

namespace {

struct ChunkPlan { int u0, n; int64_t ld; std::vector<int64_t> col_off, ll_off, ld_u; int64_t ll_floats = 0; };

// split utterances into chunks bounded by the log-likelihood block (+ back-pointers) budget
int plan_chunks(const mfa_graphs *g, const int64_t *frame_off, int n_utts, int num_pdfs, int dim, int64_t budget, std::vector<ChunkPlan> &plans,
                bool ragged = false, int u_begin = 0) {
  int u = u_begin;
  while (u < n_utts) {
    ChunkPlan c; c.u0 = u; c.n = 0;
    int64_t cols = 0, bp = 0, llf = 0;
    while (u < n_utts) {
      int64_t T = frame_off[u + 1] - frame_off[u];
      int64_t S = g->st_off[u + 1] - g->st_off[u], P = g->lp_off[u + 1] - g->lp_off[u];
      int64_t ncols = cols + round_up(T, 8);
      int64_t nbp = bp + (T + 1) * S * 2;
      int64_t nllf = llf + P * round_up(T, 8);
      // ragged: per-utterance [P_u][ld_u] blocks + A images (384 B per frame) + B images (~P_u*11 Gaussian rows * 384 B)
      int64_t bytes = ragged ? nllf * 4 + round_up(ncols, 128) * ((int64_t)dim * 4 + 800) + nbp + (P * 12 + 128) * 400
                             : round_up(ncols, 128) * ((int64_t)num_pdfs * 4 + (int64_t)dim * 4 + 400) + nbp;
      if (c.n > 0 && bytes > budget) break;
      c.col_off.push_back(cols); c.ll_off.push_back(llf); c.ld_u.push_back(round_up(T, 8));
      cols = ncols; bp = nbp; llf = nllf; c.n++; u++;
    }
    c.ld = round_up(std::max<int64_t>(cols, 1), 128);
    c.ll_floats = llf;
    plans.push_back(std::move(c));
  }
  return MFA_OK;
}

// shared tail of mfa_align / mfa_align_pcm: per chunk, (features ->) log-likelihoods -> Viterbi
struct AlignIO {
  int32_t *d_ali; float *d_pf; int32_t *d_words; int64_t *d_word_off; int32_t *d_num_words; float *d_total; int32_t *d_status;
};

}  // namespace

extern "C" {

int mfa_align(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_align_opts *o, const float *loglikes, const int64_t *frame_off,
              int32_t n_utts, int32_t *ali, float *per_frame, int32_t *words, const int64_t *word_off, int32_t *num_words,
              float *total_like, int32_t *status, int where) {
  if (!e || !m || !g || !o || !frame_off || !word_off) return set_error(MFA_ERR_INVALID, "bad argument");
  if (n_utts != g->n_utts) return set_error(MFA_ERR_INVALID, "n_utts does not match the graph batch");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(frame_off, n_utts, "frame_off"));
  MFA_TRY(check_offsets(word_off, n_utts, "word_off"));
  MFA_TRY(upload_graphs(e, g));
  const int P = m->num_pdfs;
  int64_t nf = frame_off[n_utts], nw = word_off[n_utts];
  const float *d_ll;
  MFA_TRY(to_device(e, DB_IO_LL, loglikes, (size_t)nf * P, where, &d_ll));
  AlignIO io;
  MFA_TRY(out_buffer(e, DB_ALI, ali, (size_t)nf, where, &io.d_ali));
  MFA_TRY(out_buffer(e, DB_PERFRAME, per_frame, (size_t)nf, where, &io.d_pf));
  MFA_TRY(out_buffer(e, DB_WORDS, words, (size_t)nw, where, &io.d_words));
  MFA_TRY(out_buffer(e, DB_NUM_WORDS, num_words, (size_t)n_utts, where, &io.d_num_words));
  MFA_TRY(out_buffer(e, DB_TOTAL_LIKE, total_like, (size_t)n_utts, where, &io.d_total));
  MFA_TRY(out_buffer(e, DB_STATUS, status, (size_t)n_utts, where, &io.d_status));
  int64_t *d_fo;
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  MFA_TRY(e->upload(DB_WORD_OFF, word_off, (size_t)n_utts + 1, &io.d_word_off));
  CUDA_TRY(cudaMemsetAsync(io.d_ali, 0, (size_t)nf * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(io.d_pf, 0, (size_t)nf * 4, e->stream));
  if (nw) CUDA_TRY(cudaMemsetAsync(io.d_words, 0, (size_t)nw * 4, e->stream));
  std::vector<ChunkPlan> plans;
  MFA_TRY(plan_chunks(g, frame_off, n_utts, P, 0, (int64_t)4 << 30, plans));
  for (auto &c : plans) {
    float *d_llT; int64_t *d_col;
    MFA_TRY(e->getT<float>(DB_LL, (size_t)P * c.ld, &d_llT));
    MFA_TRY(e->upload(DB_COL_OFF, c.col_off.data(), c.col_off.size(), &d_col));
    // frame-major [T][P] rows of each utterance -> pdf-major columns at col_off (padding columns are never read
    // beyond the 8-frame block of the last frame, whose values are unused)
    CUDA_TRY(cudaMemsetAsync(d_llT, 0, (size_t)P * c.ld * 4, e->stream));
    for (int k = 0; k < c.n; k++) {
      int u = c.u0 + k;
      int64_t T = frame_off[u + 1] - frame_off[u];
      MFA_TRY(launch_transpose(e, d_ll + frame_off[u] * P, T, P, P, d_llT + c.col_off[k], c.ld));
    }
    ViterbiArgs a{};
    a.g = g; a.utt0 = c.u0; a.n_utts = c.n; a.d_llT = d_llT; a.ld = c.ld; a.d_col_off = d_col; a.d_frame_off = d_fo + c.u0;
    a.h_frame_off = frame_off + c.u0; a.h_col_off = c.col_off.data();
    a.d_ali = io.d_ali; a.d_per_frame = io.d_pf; a.d_words = io.d_words; a.d_word_off = io.d_word_off + c.u0;
    a.d_num_words = io.d_num_words + c.u0; a.d_total_like = io.d_total + c.u0; a.d_status = io.d_status + c.u0; a.opts = *o;
    MFA_TRY(launch_viterbi(e, a));
    MFA_TRY(e->join_k3());
  }
  MFA_TRY(from_device(e, io.d_ali, ali, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_pf, per_frame, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_words, words, (size_t)nw, where));
  MFA_TRY(from_device(e, io.d_num_words, num_words, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_total, total_like, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_status, status, (size_t)n_utts, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_align_pcm(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_pipeline_opts *o, const int16_t *pcm, const int64_t *sample_off,
                  const int32_t *utt2spk, int32_t n_utts, int32_t n_spk, const int64_t *frame_off, int32_t *ali, float *per_frame,
                  int32_t *words, const int64_t *word_off, int32_t *num_words, float *total_like, int32_t *status, int where) {
  if (!e || !m || !g || !o || !sample_off || !frame_off || !word_off || !utt2spk) return set_error(MFA_ERR_INVALID, "bad argument");
  if (n_utts != g->n_utts) return set_error(MFA_ERR_INVALID, "n_utts does not match the graph batch");
  CUDA_TRY(cudaSetDevice(e->device));
  // The Viterbi launch of the PREVIOUS call may still be running (it is joined on its own stream): K1, the CMVN statistics and the
  // feature kernel of this call share no buffer with it and start right away; the join happens before this call's K2 (which
  // overwrites the log-likelihoods K3 reads) -- or here, if this call writes into the output buffers that launch is still filling.
  CallScope scope(e, /*defer_join=*/true);
  MFA_TRY(check_offsets(sample_off, n_utts, "sample_off"));
  MFA_TRY(check_offsets(word_off, n_utts, "word_off"));
  for (int u = 0; u < n_utts; u++)
    if (frame_off[u + 1] - frame_off[u] != mfa_mfcc_num_frames(&o->mfcc, sample_off[u + 1] - sample_off[u]))
      return set_error(MFA_ERR_INVALID, "frame_off does not match mfa_mfcc_num_frames for utterance " + std::to_string(u));
  mfa_feat_opts fo = o->feat;
  fo.in_dim = o->mfcc.num_ceps;
  if (fo.fmllr && fo.n_spk != n_spk) return set_error(MFA_ERR_INVALID, "feat.n_spk must equal n_spk when fMLLR transforms are given");
  const int D = mfa_feat_out_dim(&fo);
  if (D != m->dim) return set_error(MFA_ERR_INVALID, "feature pipeline yields dim " + std::to_string(D) + " but the model expects " + std::to_string(m->dim));
  e->gmm_timing_reset();
  e->stage_reset();
  MFA_TRY(upload_graphs(e, g));
  const int P = m->num_pdfs, C = o->mfcc.num_ceps;
  const int64_t nf = frame_off[n_utts], nw = word_off[n_utts], ns = sample_off[n_utts];
  // ---- inputs
  int64_t *d_so, *d_fo; int32_t *d_u2s; const int16_t *d_pcm;
  MFA_TRY(e->upload(DB_SAMPLE_OFF, sample_off, (size_t)n_utts + 1, &d_so));
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  MFA_TRY(e->upload(DB_UTT2SPK, utt2spk, (size_t)n_utts, &d_u2s));
  // host PCM is uploaded in pieces on a copy stream; each piece's MFCC launch waits only for its own bytes (below)
  if (where == MFA_DEVICE) d_pcm = pcm;
  else { int16_t *d; MFA_TRY(e->getT<int16_t>(DB_PCM, (size_t)std::max<int64_t>(ns, 1), &d)); d_pcm = d; }
  AlignIO io;
  MFA_TRY(out_buffer(e, DB_ALI, ali, (size_t)nf, where, &io.d_ali));
  MFA_TRY(out_buffer(e, DB_PERFRAME, per_frame, (size_t)nf, where, &io.d_pf));
  MFA_TRY(out_buffer(e, DB_WORDS, words, (size_t)nw, where, &io.d_words));
  MFA_TRY(out_buffer(e, DB_NUM_WORDS, num_words, (size_t)n_utts, where, &io.d_num_words));
  MFA_TRY(out_buffer(e, DB_TOTAL_LIKE, total_like, (size_t)n_utts, where, &io.d_total));
  MFA_TRY(out_buffer(e, DB_STATUS, status, (size_t)n_utts, where, &io.d_status));
  MFA_TRY(e->upload(DB_WORD_OFF, word_off, (size_t)n_utts + 1, &io.d_word_off));
  if (where == MFA_HOST || e->cfg.k3_overlap == 0 || e->k3_writes(io.d_ali, (size_t)nf * 4) || e->k3_writes(io.d_pf, (size_t)nf * 4) ||
      e->k3_writes(io.d_words, (size_t)nw * 4) || e->k3_writes(io.d_num_words, (size_t)n_utts * 4) || e->k3_writes(io.d_total, (size_t)n_utts * 4) ||
      e->k3_writes(io.d_status, (size_t)n_utts * 4))
    MFA_TRY(e->join_k3());
  CUDA_TRY(cudaMemsetAsync(io.d_ali, 0, (size_t)nf * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(io.d_pf, 0, (size_t)nf * 4, e->stream));
  if (nw) CUDA_TRY(cudaMemsetAsync(io.d_words, 0, (size_t)nw * 4, e->stream));
  auto mark_outputs = [&]() {   // the Viterbi launch just enqueued writes these; it stays un-joined until someone needs them
    const void *ps[6] = {io.d_ali, io.d_pf, io.d_words, io.d_num_words, io.d_total, io.d_status};
    const size_t bs[6] = {(size_t)nf * 4, (size_t)nf * 4, (size_t)nw * 4, (size_t)n_utts * 4, (size_t)n_utts * 4, (size_t)n_utts * 4};
    for (int i = 0; i < 6; i++) { e->pend_out[i] = (const char *)ps[i]; e->pend_bytes[i] = bs[i]; }
  };
  // ---- segments.  Device-resident PCM: one segment.  Host PCM: the batch is cut after a speaker boundary near the middle so
  // that the second half is still crossing PCIe while the first half is scored and aligned; CMVN statistics are per speaker,
  // so a cut is only legal where every speaker's utterances are contiguous (MFA orders jobs by speaker-utterance key) and the
  // speaker changes.  All uploads are queued on the copy stream before any compute, each MFCC launch waits for its own bytes.
  float *d_mfcc; double *d_stats = nullptr;
  MFA_TRY(e->getT<float>(DB_MFCC, (size_t)nf * C, &d_mfcc));
  const size_t nst = (size_t)n_spk * 2 * (C + 1);
  if (o->apply_cmvn) {
    if (fo.cmvn_stats) MFA_TRY(e->upload(DB_CMVN_STATS, fo.cmvn_stats, nst, &d_stats));
    else { MFA_TRY(e->getT<double>(DB_CMVN_STATS, nst, &d_stats)); CUDA_TRY(cudaMemsetAsync(d_stats, 0, nst * sizeof(double), e->stream)); }
  }
  std::vector<int> seg_begin{0};
  // Two ways to use the cuts (engine option pipeline_split): 2 (default for host PCM when the batch fits one chunk) = STREAM: MFCC, CMVN,
  // features and log-likelihoods run per segment as its bytes arrive, the Viterbi stage runs once over the whole batch at the
  // end -- it is bounded by its longest utterance's sequential recursion, not by throughput, so it must not be cut;
  // 1 = every stage per segment (measured on the 10 h config-2 workload: end to end 45 -> 55 ms, two Viterbi tails); 0 = no cuts.
  const bool split_forced = e->cfg.pipeline_split >= 0;
  const int split_mode = split_forced ? e->cfg.pipeline_split : 2;
  const int64_t budget = o->workspace_bytes > 0 ? o->workspace_bytes : ((int64_t)8 << 30);
  const bool ragged = o->gmm_impl == 0 && gmm_tc_supported(m);
  std::vector<ChunkPlan> whole;
  bool stream = false;
  if (where == MFA_HOST && split_mode == 2 && (split_forced || (n_utts >= 64 && ns >= ((int64_t)64 << 20)))) {
    MFA_TRY(plan_chunks(g, frame_off, n_utts, P, D, budget, whole, ragged));
    stream = whole.size() == 1 && ragged;   // (dense scoring keeps the chunk loop: its tiles want 128-column alignment)
  }
  if (where == MFA_HOST && n_utts >= 2 && (stream || split_mode == 1)) {
    bool contiguous = true;
    if (o->apply_cmvn && !fo.cmvn_stats) {
      std::vector<char> seen((size_t)std::max(n_spk, 1), 0);
      for (int u = 0; u < n_utts && contiguous; u++) {
        const int sp = utt2spk[u];
        if (sp < 0 || sp >= n_spk) return set_error(MFA_ERR_INVALID, "utt2spk out of range");
        if (u > 0 && sp != utt2spk[u - 1] && seen[sp]) contiguous = false;
        seen[sp] = 1;
      }
    }
    if (contiguous) {
      const int n_cut = stream ? 4 : 2;                      // segments
      for (int k = 1, u = 0; k < n_cut; k++) {
        while (u < n_utts && sample_off[u] < ns * k / n_cut) u++;
        if (o->apply_cmvn && !fo.cmvn_stats) while (u < n_utts && u > 0 && utt2spk[u] == utt2spk[u - 1]) u++;
        if (u > seg_begin.back() && u < n_utts) seg_begin.push_back(u);
      }
    } else stream = false;
  }
  if (seg_begin.size() == 1) stream = false;
  seg_begin.push_back(n_utts);
  if (ragged) {
    // B images of the first chunk (the whole batch when it fits the workspace) start building now, under K1
    int n_first = n_utts;
    if (!stream) {
      std::vector<ChunkPlan> first;
      MFA_TRY(plan_chunks(g, frame_off, seg_begin[1], P, D, budget, first, true, 0));
      n_first = first.empty() ? 0 : first[0].n;
    }
    MFA_TRY(prefetch_b_images(e, m, g, 0, n_first));
  }
  // ---- uploads (host PCM): pieces of >= 32 MB that never straddle a segment
  struct Piece { int u0, u1, ev; };
  std::vector<Piece> pieces;
  if (where == MFA_HOST) {
    const int64_t piece = std::max<int64_t>((int64_t)16 << 20, ns / 12 + 1);   // samples per H2D piece
    cudaStream_t cs = e->side[0];
    CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
    CUDA_TRY(cudaStreamWaitEvent(cs, e->ev_fork, 0));   // earlier work on the main stream may still read DB_PCM
    for (size_t sg = 0; sg + 1 < seg_begin.size(); sg++) {
      int u0 = seg_begin[sg];
      while (u0 < seg_begin[sg + 1]) {
        int u1 = u0;
        while (u1 < seg_begin[sg + 1] && sample_off[u1 + 1] - sample_off[u0] <= piece) u1++;
        if (u1 == u0) u1 = u0 + 1;
        const int64_t s0 = sample_off[u0], s1 = sample_off[u1];
        const int k = (int)pieces.size();
        if ((int)e->ev_piece.size() <= k) { cudaEvent_t ev; CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e->ev_piece.push_back(ev); }
        if (s1 > s0) CUDA_TRY(cudaMemcpyAsync((int16_t *)d_pcm + s0, pcm + s0, (size_t)(s1 - s0) * sizeof(int16_t), cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaEventRecord(e->ev_piece[k], cs));
        pieces.push_back({u0, u1, k});
        u0 = u1;
      }
    }
  }
  size_t next_piece = 0;
  // STREAM mode: one chunk plan for the whole batch; its buffers are filled segment by segment
  float *w_feats = nullptr, *w_llT = nullptr; int64_t *w_col = nullptr, *w_ll_off = nullptr, *w_ld_u = nullptr;
  if (stream) {
    ChunkPlan &c = whole[0];
    MFA_TRY(e->join_k3());
    MFA_TRY(e->getT<float>(DB_FEATS, (size_t)c.ld * D, &w_feats));
    MFA_TRY(e->getT<float>(DB_LL, ragged ? (size_t)c.ll_floats + 8 : (size_t)P * c.ld, &w_llT));
    MFA_TRY(e->upload(DB_COL_OFF, c.col_off.data(), c.col_off.size(), &w_col));
    if (ragged) {
      MFA_TRY(e->upload(DB_LL_OFF, c.ll_off.data(), c.ll_off.size(), &w_ll_off));
      MFA_TRY(e->upload(DB_LD_U, c.ld_u.data(), c.ld_u.size(), &w_ld_u));
    }
    CUDA_TRY(cudaMemsetAsync(w_feats, 0, (size_t)c.ld * D * 4, e->stream));
  }
  for (size_t sg = 0; sg + 1 < seg_begin.size(); sg++) {
    const int sb = seg_begin[sg], se = seg_begin[sg + 1];
    // ---- K1 + CMVN statistics of the segment
    MFA_TRY(e->stage_begin(mfa_engine::ST_MFCC));
    if (where == MFA_DEVICE) {
      MFA_TRY(launch_mfcc(e, &o->mfcc, d_pcm, d_so, n_utts, d_fo, nf, d_mfcc));
    } else {
      for (; next_piece < pieces.size() && pieces[next_piece].u1 <= se; next_piece++) {
        const Piece &pc = pieces[next_piece];
        CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_piece[pc.ev], 0));
        MFA_TRY(launch_mfcc(e, &o->mfcc, d_pcm, d_so, n_utts, d_fo, frame_off[pc.u1] - frame_off[pc.u0], d_mfcc, frame_off[pc.u0]));
      }
    }
    if (o->apply_cmvn && !fo.cmvn_stats) MFA_TRY(launch_cmvn_stats(e, d_mfcc, C, d_fo + sb, utt2spk + sb, se - sb, n_spk, d_stats));
    MFA_TRY(e->stage_end());
    if (stream) {
      // features and log-likelihoods of the segment's utterances, written into the whole-batch buffers
      ChunkPlan &c = whole[0];
      const int sn = se - sb;
      MFA_TRY(e->stage_begin(mfa_engine::ST_FEAT));
      MFA_TRY(launch_features(e, &fo, d_mfcc, d_fo + sb, frame_off + sb, w_col + sb, d_u2s + sb, sn, d_stats, w_feats, D));
      MFA_TRY(e->stage_end());
      MFA_TRY(e->stage_begin(mfa_engine::ST_GMM));
      MFA_TRY(e->gmm_timing_begin());
      MFA_TRY(launch_gmm_tc_ragged(e, m, g, sb, sn, w_feats, c.col_off.data() + sb, frame_off + sb, w_llT, c.ll_off.data() + sb, c.ld_u.data() + sb));
      MFA_TRY(e->gmm_timing_end(frame_off[se] - frame_off[sb]));
      MFA_TRY(e->stage_end());
      if (se == n_utts) {
        ViterbiArgs a{};
        a.d_ll_off = w_ll_off; a.d_ld_u = w_ld_u;
        a.g = g; a.utt0 = 0; a.n_utts = n_utts; a.d_llT = w_llT; a.ld = c.ld; a.d_col_off = w_col; a.d_frame_off = d_fo;
        a.h_frame_off = frame_off; a.h_col_off = c.col_off.data();
        a.d_ali = io.d_ali; a.d_per_frame = io.d_pf; a.d_words = io.d_words; a.d_word_off = io.d_word_off;
        a.d_num_words = io.d_num_words; a.d_total_like = io.d_total; a.d_status = io.d_status; a.opts = o->align; a.host_call = where == MFA_HOST;
        MFA_TRY(e->stage_begin(mfa_engine::ST_VITERBI));
        MFA_TRY(launch_viterbi(e, a));
        MFA_TRY(e->stage_end(e->sj));
        mark_outputs();
      }
      continue;
    }
    // ---- chunks of the segment
    std::vector<ChunkPlan> plans;
    MFA_TRY(plan_chunks(g, frame_off, se, P, D, budget, plans, ragged, sb));
    for (auto &c : plans) {
      float *d_feats, *d_llT; int64_t *d_col, *d_ll_off = nullptr, *d_ld_u = nullptr;
      MFA_TRY(e->getT<float>(DB_FEATS, (size_t)c.ld * D, &d_feats));
      MFA_TRY(e->upload(DB_COL_OFF, c.col_off.data(), c.col_off.size(), &d_col));
      MFA_TRY(e->stage_begin(mfa_engine::ST_FEAT));
      CUDA_TRY(cudaMemsetAsync(d_feats, 0, (size_t)c.ld * D * 4, e->stream));
      MFA_TRY(launch_features(e, &fo, d_mfcc, d_fo + c.u0, frame_off + c.u0, d_col, d_u2s + c.u0, c.n, d_stats, d_feats, D));
      MFA_TRY(e->stage_end());
      MFA_TRY(e->join_k3());   // from here on this chunk overwrites what a Viterbi launch still in flight reads (log-likelihoods, offsets)
      MFA_TRY(e->getT<float>(DB_LL, ragged ? (size_t)c.ll_floats + 8 : (size_t)P * c.ld, &d_llT));
      MFA_TRY(e->stage_begin(mfa_engine::ST_GMM));
      if (ragged) {
        MFA_TRY(e->upload(DB_LL_OFF, c.ll_off.data(), c.ll_off.size(), &d_ll_off));
        MFA_TRY(e->upload(DB_LD_U, c.ld_u.data(), c.ld_u.size(), &d_ld_u));
        MFA_TRY(e->gmm_timing_begin());
        MFA_TRY(launch_gmm_tc_ragged(e, m, g, c.u0, c.n, d_feats, c.col_off.data(), frame_off + c.u0, d_llT, c.ll_off.data(), c.ld_u.data()));
        MFA_TRY(e->gmm_timing_end(frame_off[c.u0 + c.n] - frame_off[c.u0]));
      } else {
        MFA_TRY(run_gmm(e, m, d_feats, c.ld, d_llT, c.ld, o->gmm_impl));
      }
      MFA_TRY(e->stage_end());
      ViterbiArgs a{};
      a.d_ll_off = d_ll_off; a.d_ld_u = d_ld_u;
      a.g = g; a.utt0 = c.u0; a.n_utts = c.n; a.d_llT = d_llT; a.ld = c.ld; a.d_col_off = d_col; a.d_frame_off = d_fo + c.u0;
      a.h_frame_off = frame_off + c.u0; a.h_col_off = c.col_off.data();
      a.d_ali = io.d_ali; a.d_per_frame = io.d_pf; a.d_words = io.d_words; a.d_word_off = io.d_word_off + c.u0;
      a.d_num_words = io.d_num_words + c.u0; a.d_total_like = io.d_total + c.u0; a.d_status = io.d_status + c.u0; a.opts = o->align; a.host_call = where == MFA_HOST;
      MFA_TRY(e->stage_begin(mfa_engine::ST_VITERBI));
      MFA_TRY(launch_viterbi(e, a));
      MFA_TRY(e->stage_end(e->sj));
      mark_outputs();
    }
  }
  if (where == MFA_HOST) MFA_TRY(e->join_k3());   // the copies below read what K3 writes
  MFA_TRY(from_device(e, io.d_ali, ali, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_pf, per_frame, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_words, words, (size_t)nw, where));
  MFA_TRY(from_device(e, io.d_num_words, num_words, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_total, total_like, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_status, status, (size_t)n_utts, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_align_feats(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_align_opts *o, const float *feats, const int64_t *frame_off,
                    int32_t n_utts, int32_t gmm_impl, int64_t workspace_bytes, int32_t *ali, float *per_frame, int32_t *words,
                    const int64_t *word_off, int32_t *num_words, float *total_like, int32_t *status, int where) {
  if (!e || !m || !g || !o || !frame_off || !word_off) return set_error(MFA_ERR_INVALID, "bad argument");
  if (n_utts != g->n_utts) return set_error(MFA_ERR_INVALID, "n_utts does not match the graph batch");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(frame_off, n_utts, "frame_off"));
  MFA_TRY(check_offsets(word_off, n_utts, "word_off"));
  e->gmm_timing_reset();
  e->stage_reset();
  MFA_TRY(upload_graphs(e, g));
  const int P = m->num_pdfs, D = m->dim;
  const int64_t nf = frame_off[n_utts], nw = word_off[n_utts];
  int64_t *d_fo; const float *d_in;
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  MFA_TRY(to_device(e, DB_IO_FEATS, feats, (size_t)nf * D, where, &d_in));
  AlignIO io;
  MFA_TRY(out_buffer(e, DB_ALI, ali, (size_t)nf, where, &io.d_ali));
  MFA_TRY(out_buffer(e, DB_PERFRAME, per_frame, (size_t)nf, where, &io.d_pf));
  MFA_TRY(out_buffer(e, DB_WORDS, words, (size_t)nw, where, &io.d_words));
  MFA_TRY(out_buffer(e, DB_NUM_WORDS, num_words, (size_t)n_utts, where, &io.d_num_words));
  MFA_TRY(out_buffer(e, DB_TOTAL_LIKE, total_like, (size_t)n_utts, where, &io.d_total));
  MFA_TRY(out_buffer(e, DB_STATUS, status, (size_t)n_utts, where, &io.d_status));
  MFA_TRY(e->upload(DB_WORD_OFF, word_off, (size_t)n_utts + 1, &io.d_word_off));
  CUDA_TRY(cudaMemsetAsync(io.d_ali, 0, (size_t)nf * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(io.d_pf, 0, (size_t)nf * 4, e->stream));
  if (nw) CUDA_TRY(cudaMemsetAsync(io.d_words, 0, (size_t)nw * 4, e->stream));
  const int64_t budget = workspace_bytes > 0 ? workspace_bytes : ((int64_t)8 << 30);
  const bool ragged = gmm_impl == 0 && gmm_tc_supported(m);
  mfa_feat_opts copy{};   // mode 0: the features are final; the kernel only moves them into the 8-aligned per-utterance columns K2 reads
  copy.mode = 0; copy.in_dim = D;
  std::vector<ChunkPlan> plans;
  MFA_TRY(plan_chunks(g, frame_off, n_utts, P, D, budget, plans, ragged));
  for (auto &c : plans) {
    float *d_feats, *d_llT; int64_t *d_col, *d_ll_off = nullptr, *d_ld_u = nullptr;
    MFA_TRY(e->getT<float>(DB_FEATS, (size_t)c.ld * D, &d_feats));
    MFA_TRY(e->getT<float>(DB_LL, ragged ? (size_t)c.ll_floats + 8 : (size_t)P * c.ld, &d_llT));
    MFA_TRY(e->upload(DB_COL_OFF, c.col_off.data(), c.col_off.size(), &d_col));
    MFA_TRY(e->stage_begin(mfa_engine::ST_FEAT));
    CUDA_TRY(cudaMemsetAsync(d_feats, 0, (size_t)c.ld * D * 4, e->stream));
    MFA_TRY(launch_features(e, &copy, d_in, d_fo + c.u0, frame_off + c.u0, d_col, nullptr, c.n, nullptr, d_feats, D));
    MFA_TRY(e->stage_end());
    MFA_TRY(e->stage_begin(mfa_engine::ST_GMM));
    if (ragged) {
      MFA_TRY(e->upload(DB_LL_OFF, c.ll_off.data(), c.ll_off.size(), &d_ll_off));
      MFA_TRY(e->upload(DB_LD_U, c.ld_u.data(), c.ld_u.size(), &d_ld_u));
      MFA_TRY(e->gmm_timing_begin());
      MFA_TRY(launch_gmm_tc_ragged(e, m, g, c.u0, c.n, d_feats, c.col_off.data(), frame_off + c.u0, d_llT, c.ll_off.data(), c.ld_u.data()));
      MFA_TRY(e->gmm_timing_end(frame_off[c.u0 + c.n] - frame_off[c.u0]));
    } else {
      MFA_TRY(run_gmm(e, m, d_feats, c.ld, d_llT, c.ld, gmm_impl));
    }
    MFA_TRY(e->stage_end());
    ViterbiArgs a{};
    a.d_ll_off = d_ll_off; a.d_ld_u = d_ld_u;
    a.g = g; a.utt0 = c.u0; a.n_utts = c.n; a.d_llT = d_llT; a.ld = c.ld; a.d_col_off = d_col; a.d_frame_off = d_fo + c.u0;
    a.h_frame_off = frame_off + c.u0; a.h_col_off = c.col_off.data();
    a.d_ali = io.d_ali; a.d_per_frame = io.d_pf; a.d_words = io.d_words; a.d_word_off = io.d_word_off + c.u0;
    a.d_num_words = io.d_num_words + c.u0; a.d_total_like = io.d_total + c.u0; a.d_status = io.d_status + c.u0; a.opts = *o;
    MFA_TRY(e->stage_begin(mfa_engine::ST_VITERBI));
    MFA_TRY(launch_viterbi(e, a));
    MFA_TRY(e->stage_end(e->sj));
    MFA_TRY(e->join_k3());
  }
  MFA_TRY(from_device(e, io.d_ali, ali, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_pf, per_frame, (size_t)nf, where));
  MFA_TRY(from_device(e, io.d_words, words, (size_t)nw, where));
  MFA_TRY(from_device(e, io.d_num_words, num_words, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_total, total_like, (size_t)n_utts, where));
  MFA_TRY(from_device(e, io.d_status, status, (size_t)n_utts, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_acc_stats(mfa_engine *e, mfa_model *m, const float *feats, const int32_t *ali, int64_t n_frames, int where) {
  if (!e || !m || n_frames < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  const float *d_feats; const int32_t *d_ali;
  MFA_TRY(to_device(e, DB_IO_FEATS, feats, (size_t)n_frames * m->dim, where, &d_feats));
  MFA_TRY(to_device(e, DB_IO_ALI, ali, (size_t)n_frames, where, &d_ali));
  MFA_TRY(launch_acc_stats(e, m, d_feats, d_ali, n_frames));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_fmllr_acc(mfa_engine *e, mfa_model *post_model, mfa_model *m, const float *feats, const int32_t *ali, const float *tid_weight,
                  const int64_t *frame_off, const int32_t *utt2spk, int32_t n_utts, int32_t n_spk, double *stats, int where) {
  if (!e || !m || !frame_off || !utt2spk || !stats || n_utts < 0 || n_spk < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  if (!post_model) post_model = m;
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  MFA_TRY(check_offsets(frame_off, n_utts, "frame_off"));
  const int64_t nf = frame_off[n_utts];
  int64_t *d_fo; const float *d_feats; const int32_t *d_ali; float *d_tw = nullptr; double *d_stats;
  MFA_TRY(e->upload(DB_FRAME_OFF, frame_off, (size_t)n_utts + 1, &d_fo));
  if (tid_weight) MFA_TRY(e->upload(DB_SCRATCH, tid_weight, (size_t)m->num_tids + 1, &d_tw));
  MFA_TRY(to_device(e, DB_IO_FEATS, feats, (size_t)nf * m->dim, where, &d_feats));
  MFA_TRY(to_device(e, DB_IO_ALI, ali, (size_t)nf, where, &d_ali));
  const size_t ns = (size_t)n_spk * (size_t)mfa_fmllr_stats_size(m->dim);
  MFA_TRY(out_buffer(e, DB_FM_STATS, stats, ns, where, &d_stats));
  CUDA_TRY(cudaMemsetAsync(d_stats, 0, ns * sizeof(double), e->stream));
  MFA_TRY(launch_fmllr_acc(e, post_model, m, d_feats, d_ali, d_tw, d_fo, frame_off, utt2spk, n_utts, n_spk, d_stats));
  MFA_TRY(from_device(e, d_stats, stats, ns, where));
  if (where == MFA_HOST) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

int mfa_fmllr_update(mfa_engine *e, const double *stats, int32_t dim, int32_t n_spk, int32_t num_iters, double min_count, float *transforms,
                     double *objf_impr, double *count, int where) {
  if (!e || !stats || !transforms || n_spk < 0 || dim < 1) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  MFA_TRY(e->join_k3());
  if (num_iters <= 0) num_iters = 40;
  const size_t ns = (size_t)n_spk * (size_t)mfa_fmllr_stats_size(dim), nw = (size_t)n_spk * dim * (dim + 1);
  const double *d_stats; float *d_W; double *d_out;
  MFA_TRY(to_device(e, DB_FM_STATS, stats, ns, where, &d_stats));
  MFA_TRY(out_buffer(e, DB_FM_W, transforms, nw, where, &d_W));
  MFA_TRY(e->getT<double>(DB_FM_OUT, (size_t)2 * std::max(n_spk, 1), &d_out));
  MFA_TRY(launch_fmllr_update(e, d_stats, dim, n_spk, num_iters, min_count, d_W, d_out, d_out + n_spk));
  MFA_TRY(from_device(e, d_W, transforms, nw, where));
  // the two small per-speaker vectors always land in host memory
  if (objf_impr) CUDA_TRY(cudaMemcpyAsync(objf_impr, d_out, sizeof(double) * n_spk, cudaMemcpyDeviceToHost, e->stream));
  if (count) CUDA_TRY(cudaMemcpyAsync(count, d_out + n_spk, sizeof(double) * n_spk, cudaMemcpyDeviceToHost, e->stream));
  if (where == MFA_HOST || objf_impr || count) CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

}  // extern "C"
