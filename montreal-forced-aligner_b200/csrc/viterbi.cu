// viterbi.cu -- K3: beam Viterbi over many utterances' training graphs at once, with AlignUtteranceWrapper's
// retry-with-wider-beam inside the kernel.
//
// Replaces GmmAligner.align_utterance / export_alignments -> Kaldi AlignUtteranceWrapper + FasterDecoder
// (reference call sites: montreal_forced_aligner/alignment/multiprocessing.py:846-853, online/alignment.py:97-107).
// Semantics per SURVEY.md A.7 (decoder/faster-decoder.cc, decoder/decoder-wrappers.cc):
//   cost = graph cost (+ AddTransitionProbs, folded in at pack time) - acoustic_scale * loglike
//   per frame: GetCutoff (beam, min_active widening with beam_delta) -> expand tokens under the cutoff over
//   emitting arcs, keep new tokens under best_new + adaptive_beam -> epsilon closure under the same cutoff;
//   success iff a live token sits in a final state after the last frame, else rerun with retry_beam.
//
// Formulation: a dense "pull" dynamic programme -- every state takes the min over its in-arcs (arcs are grouped by
// destination, so no atomics), tokens outside the beam are masked to +inf.  One CTA per utterance; the graph's arcs,
// both token arrays and an 8-frame block of the utterance's acoustic costs live in shared memory; back-pointers
// (uint16 arc index per frame x state) stream to HBM and are walked back by one thread at the end.
// Token costs are kept relative to the frame's best token (fp32) with the running offset in fp64, which is
// as accurate as FasterDecoder's double-precision token costs at the magnitudes that matter for comparisons.
//
// Differences from FasterDecoder that cannot change a surviving best path: (1) new tokens are pruned against the
// final per-frame cutoff rather than the cutoff as it tightens in hash-list order (a superset of tokens can exist in
// Kaldi for one frame; they are pruned by the next GetCutoff unless fewer than min_active tokens are in the beam);
// (2) ties between equal-cost in-arcs resolve to the first arc in the state's arc order.
#include <algorithm>
#include <numeric>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {
constexpr int VT = 128;  // threads per CTA
constexpr unsigned kNoArc = 0xFFFFu;
constexpr unsigned kEps = 0xFFFFu;

struct VitParams {
  const int64_t *st_off, *arc_off, *lp_off, *inb_off;
  const int32_t *start, *n_eps, *in_begin, *a_tid, *a_olabel, *lp2pdf;
  const uint32_t *a_pack;
  const float *a_w, *final_w;
  int utt0;
  const int32_t *order;
  const float *llT;
  int64_t ld;
  const int64_t *col_off, *frame_off, *bp_off, *word_off;
  uint16_t *bp;
  int32_t *ali, *num_words, *words, *status;
  float *per_frame, *total_like;
  float acwt, beam, retry_beam, beam_delta;
  int min_active;
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool SMEM_ARCS>
__global__ void __launch_bounds__(VT)
viterbi_kernel(VitParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ float red_f[2][VT / 32];
  __shared__ int red_i[2][2 * (VT / 32)];
  __shared__ int sh_flag, sh_cnt, sh_best_state;
  __shared__ float sh_sel;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ul = p.order[blockIdx.x];
  const int ug = p.utt0 + ul;
  const int S = (int)(p.st_off[ug + 1] - p.st_off[ug]);
  const int A = (int)(p.arc_off[ug + 1] - p.arc_off[ug]);
  const int P = (int)(p.lp_off[ug + 1] - p.lp_off[ug]);
  const int64_t T = p.frame_off[ul + 1] - p.frame_off[ul];
  const int start = p.start[ug];
  const bool has_eps = p.n_eps[ug] > 0;
  if (start < 0 || S == 0) { if (tid == 0) { p.status[ul] = MFA_ALIGN_EMPTY_GRAPH; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }
  if (T == 0) { if (tid == 0) { p.status[ul] = MFA_ALIGN_ZERO_FRAMES; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }

  float *cost_a = (float *)smraw;
  float *cost_b = cost_a + S;
  float *ac = cost_b + S;                       // [8][P]
  uint16_t *inb = (uint16_t *)(ac + 8 * P);     // [S+1]
  const int inb_words = (S + 2) / 2;
  uint32_t *s_pack = (uint32_t *)inb + inb_words;
  float *s_w = (float *)(s_pack + A);
  const int32_t *g_inb = p.in_begin + p.inb_off[ug];
  const uint32_t *g_pack = p.a_pack + p.arc_off[ug];
  const float *g_w = p.a_w + p.arc_off[ug];
  const float *fin = p.final_w + p.st_off[ug];
  const int32_t *lp2pdf = p.lp2pdf + p.lp_off[ug];
  for (int i = tid; i <= S; i += VT) inb[i] = (uint16_t)g_inb[i];
  if (SMEM_ARCS) for (int i = tid; i < A; i += VT) { s_pack[i] = g_pack[i]; s_w[i] = g_w[i]; }
  const uint32_t *pack = SMEM_ARCS ? s_pack : g_pack;
  const float *aw = SMEM_ARCS ? s_w : g_w;
  uint16_t *bp = p.bp + p.bp_off[ul];           // rows 0..T-1 (+ row T: initial epsilon closure)
  const float *ll = p.llT + p.col_off[ul];
  const float inf = INFINITY;

  int result = MFA_ALIGN_NO_FINAL;
  double offset = 0.0;
  float *cur = cost_a, *nxt = cost_b;
  int par = 0;

  for (int attempt = 0; attempt < 2; attempt++) {
    const float beam = attempt == 0 ? p.beam : p.retry_beam;
    if (attempt == 1 && !(p.retry_beam > 0.0f)) break;
    __syncthreads();
    cur = cost_a; nxt = cost_b; offset = 0.0;
    for (int s = tid; s < S; s += VT) cur[s] = (s == start) ? 0.0f : inf;
    if (has_eps) for (int s = tid; s < S; s += VT) bp[(size_t)T * S + s] = (uint16_t)kNoArc;
    __syncthreads();
    float cutoff = inf, adaptive = inf;  // GetCutoff of the initial token list
    int n_tot = 1, n_beam = 1;
    if (has_eps) {
      // ProcessNonemitting(+inf) from the start state
      for (;;) {
        if (tid == 0) sh_flag = 0;
        __syncthreads();
        int ch = 0;
        for (int s = tid; s < S; s += VT)
          for (int a = inb[s]; a < inb[s + 1]; a++) {
            uint32_t pk = pack[a];
            if ((pk >> 16) != kEps) continue;
            float v = cur[pk & 0xFFFF] + aw[a];
            if (v < cur[s]) { cur[s] = v; bp[(size_t)T * S + s] = (uint16_t)a; ch = 1; }
          }
        if (ch) sh_flag = 1;
        __syncthreads();
        if (!sh_flag) break;
        __syncthreads();
      }
      int ct = 0, cb = 0;
      for (int s = tid; s < S; s += VT) { float v = cur[s]; if (v < inf) { ct++; if (v <= beam) cb++; } }
      ct = warp_sum(ct); cb = warp_sum(cb);
      if (lane == 0) { red_i[par][warp] = ct; red_i[par][VT / 32 + warp] = cb; }
      __syncthreads();
      n_tot = 0; n_beam = 0;
      for (int w = 0; w < VT / 32; w++) { n_tot += red_i[par][w]; n_beam += red_i[par][VT / 32 + w]; }
      par ^= 1;
    }
    bool dead = false;
    for (int64_t t = 0; t < T; t++) {
      // ---- GetCutoff for the tokens in `cur` (normalised: best == 0)
      if (n_tot <= p.min_active) { cutoff = inf; adaptive = inf; }
      else if (n_beam > p.min_active) { cutoff = beam; adaptive = beam; }
      else {
        // min_active-th order statistic (0-based) of the finite costs: compact into `nxt`, radix-select in warp 0
        if (tid == 0) sh_cnt = 0;
        __syncthreads();
        for (int s = tid; s < S; s += VT) { float v = cur[s]; if (v < inf) nxt[atomicAdd(&sh_cnt, 1)] = v; }
        __syncthreads();
        if (warp == 0) {
          const int n = sh_cnt;
          unsigned prefix = 0, mask = 0;
          int want = p.min_active;
          for (int bit = 31; bit >= 0; bit--) {
            unsigned b = 1u << bit;
            int c0 = 0;
            for (int i = lane; i < n; i += 32) { unsigned x = __float_as_uint(nxt[i]); if ((x & mask) == prefix && !(x & b)) c0++; }
            c0 = warp_sum(c0);
            if (want >= c0) { prefix |= b; want -= c0; }
            mask |= b;
          }
          if (lane == 0) sh_sel = __uint_as_float(prefix);
        }
        __syncthreads();
        cutoff = sh_sel; adaptive = cutoff + p.beam_delta;
        __syncthreads();
      }
      // ---- acoustic costs for frames t..t+7
      if ((t & 7) == 0) {
        __syncthreads();
        for (int lp = tid; lp < P; lp += VT) {
          const float4 *src = (const float4 *)(ll + (size_t)lp2pdf[lp] * p.ld + t);
          float4 v0 = src[0], v1 = src[1];
          ac[0 * P + lp] = -p.acwt * v0.x; ac[1 * P + lp] = -p.acwt * v0.y; ac[2 * P + lp] = -p.acwt * v0.z; ac[3 * P + lp] = -p.acwt * v0.w;
          ac[4 * P + lp] = -p.acwt * v1.x; ac[5 * P + lp] = -p.acwt * v1.y; ac[6 * P + lp] = -p.acwt * v1.z; ac[7 * P + lp] = -p.acwt * v1.w;
        }
        __syncthreads();
      }
      const float *acf = ac + (int)(t & 7) * P;
      uint16_t *bprow = bp + (size_t)t * S;
      // ---- ProcessEmitting: pull over in-arcs
      float lmin = inf;
      for (int s = tid; s < S; s += VT) {
        float best = inf; unsigned barc = kNoArc;
        for (int a = inb[s], a1 = inb[s + 1]; a < a1; a++) {
          uint32_t pk = pack[a];
          unsigned lp = pk >> 16;
          if (lp == kEps) continue;
          float c = cur[pk & 0xFFFF];
          if (c < cutoff) { float v = (c + aw[a]) + acf[lp]; if (v < best) { best = v; barc = a; } }
        }
        nxt[s] = best; bprow[s] = (uint16_t)barc;
        lmin = fminf(lmin, best);
      }
      lmin = warp_min(lmin);
      if (lane == 0) red_f[par][warp] = lmin;
      __syncthreads();
      float best_new = red_f[par][0];
#pragma unroll
      for (int w = 1; w < VT / 32; w++) best_new = fminf(best_new, red_f[par][w]);
      if (!(best_new < inf)) { dead = true; break; }
      const float next_cutoff = best_new + adaptive;  // inf stays inf
      int ct = 0, cb = 0;
      for (int s = tid; s < S; s += VT) {
        float v = nxt[s];
        if (v < next_cutoff) { v -= best_new; ct++; if (v <= beam) cb++; } else v = inf;
        nxt[s] = v;
      }
      if (has_eps) {
        // ProcessNonemitting(next_cutoff), normalised cutoff = adaptive
        for (;;) {
          __syncthreads();
          if (tid == 0) sh_flag = 0;
          __syncthreads();
          int ch = 0;
          for (int s = tid; s < S; s += VT)
            for (int a = inb[s]; a < inb[s + 1]; a++) {
              uint32_t pk = pack[a];
              if ((pk >> 16) != kEps) continue;
              float c = nxt[pk & 0xFFFF];
              if (!(c <= adaptive)) continue;
              float v = c + aw[a];
              if (v <= adaptive && v < nxt[s]) { nxt[s] = v; bprow[s] = (uint16_t)a; ch = 1; }
            }
          if (ch) sh_flag = 1;
          __syncthreads();
          if (!sh_flag) break;
        }
        ct = 0; cb = 0;
        for (int s = tid; s < S; s += VT) { float v = nxt[s]; if (v < inf) { ct++; if (v <= beam) cb++; } }
      }
      ct = warp_sum(ct); cb = warp_sum(cb);
      if (lane == 0) { red_i[par][warp] = ct; red_i[par][VT / 32 + warp] = cb; }
      __syncthreads();
      n_tot = 0; n_beam = 0;
#pragma unroll
      for (int w = 0; w < VT / 32; w++) { n_tot += red_i[par][w]; n_beam += red_i[par][VT / 32 + w]; }
      par ^= 1;
      offset += (double)best_new;
      float *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (dead) continue;
    // ---- ReachedFinal / best final token
    float lbest = inf;
    for (int s = tid; s < S; s += VT) { float v = cur[s]; if (v < inf) lbest = fminf(lbest, v + fin[s]); }
    lbest = warp_min(lbest);
    if (lane == 0) red_f[par][warp] = lbest;
    if (tid == 0) sh_best_state = 0x7fffffff;
    __syncthreads();
    float fbest = red_f[par][0];
    for (int w = 1; w < VT / 32; w++) fbest = fminf(fbest, red_f[par][w]);
    par ^= 1;
    if (fbest < inf) {
      for (int s = tid; s < S; s += VT) { float v = cur[s]; if (v < inf && v + fin[s] == fbest) atomicMin(&sh_best_state, s); }
      __syncthreads();
      result = attempt == 0 ? MFA_ALIGN_OK : MFA_ALIGN_RETRIED;
      if (tid == 0) p.total_like[ul] = (float)(-(offset + (double)fbest) / (double)p.acwt);
      break;
    }
    __syncthreads();
  }
  __syncthreads();  // all back-pointer writes of this CTA are visible to thread 0 below (same CTA, global memory)
  if (tid != 0) return;
  p.status[ul] = result;
  if (result == MFA_ALIGN_NO_FINAL) { p.num_words[ul] = 0; p.total_like[ul] = 0.0f; return; }
  // ---- back-trace (one thread): frames T-1..0, then the initial epsilon closure
  const int32_t *a_tid = p.a_tid + p.arc_off[ug], *a_ol = p.a_olabel + p.arc_off[ug];
  int32_t *ali = p.ali + p.frame_off[ul];
  float *pf = p.per_frame + p.frame_off[ul];
  int32_t *words = p.words + p.word_off[ul];
  const int wcap = (int)(p.word_off[ul + 1] - p.word_off[ul]);
  int nw = 0;
  int s = sh_best_state;
  int64_t t = T - 1;
  int guard = 0;
  while (t >= 0) {
    unsigned a = bp[(size_t)t * S + s];
    if (a == kNoArc) { p.status[ul] = MFA_ALIGN_NO_FINAL; p.num_words[ul] = 0; return; }  // cannot happen
    uint32_t pk = g_pack[a];
    int ol = a_ol[a];
    if (ol != 0) { if (nw < wcap) words[nw] = ol; nw++; }
    s = pk & 0xFFFF;
    if ((pk >> 16) == kEps) { if (++guard > S) { p.status[ul] = MFA_ALIGN_NO_FINAL; p.num_words[ul] = 0; return; } continue; }
    guard = 0;
    ali[t] = a_tid[a];
    pf[t] = ll[(size_t)lp2pdf[pk >> 16] * p.ld + t];
    t--;
  }
  if (has_eps) {
    guard = 0;
    while (s != start && guard++ <= S) {
      unsigned a = bp[(size_t)T * S + s];
      if (a == kNoArc) break;
      int ol = a_ol[a];
      if (ol != 0) { if (nw < wcap) words[nw] = ol; nw++; }
      s = g_pack[a] & 0xFFFF;
    }
  }
  int n = nw < wcap ? nw : wcap;
  for (int i = 0; i < n / 2; i++) { int32_t x = words[i]; words[i] = words[n - 1 - i]; words[n - 1 - i] = x; }
  p.num_words[ul] = nw;
}

}  // namespace

namespace mfa {

int launch_viterbi(mfa_engine *e, const ViterbiArgs &a) {
  const mfa_graphs *g = a.g;
  const int n = a.n_utts;
  if (n == 0) return MFA_OK;
  // shared-memory need and back-pointer offsets per utterance
  std::vector<int64_t> bp_off(n + 1, 0);
  std::vector<size_t> need(n);
  std::vector<double> work(n);
  for (int u = 0; u < n; u++) {
    int ug = a.utt0 + u;
    int64_t S = g->st_off[ug + 1] - g->st_off[ug], A = g->arc_off[ug + 1] - g->arc_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug];
    int64_t T = a.h_frame_off[u + 1] - a.h_frame_off[u];
    bp_off[u + 1] = bp_off[u] + (T + (g->n_eps[ug] > 0 ? 1 : 0)) * S;
    need[u] = (size_t)S * 8 + (size_t)P * 32 + (size_t)((S + 2) / 2) * 4 + (size_t)A * 8;
    work[u] = (double)T * (double)(A + S);
    if (a.h_col_off[u] % 8 != 0) return set_error(MFA_ERR_INVALID, "col_off must be a multiple of 8");
    if (a.h_col_off[u] + ((T + 7) / 8) * 8 > a.ld) return set_error(MFA_ERR_INVALID, "log-likelihood leading dimension too small for 8-frame blocks");
  }
  uint16_t *d_bp; int64_t *d_bp_off; int32_t *d_order;
  MFA_TRY(e->getT<uint16_t>(DB_BP, (size_t)bp_off[n] + 8, &d_bp));
  MFA_TRY(e->upload(DB_BP_OFF, bp_off.data(), bp_off.size(), &d_bp_off));
  // classes by shared-memory need; inside a class, longest work first
  const size_t limit = e->smem_optin - 1024;
  const size_t bounds[4] = {40 * 1024, 72 * 1024, 110 * 1024, limit};
  std::vector<int> cls(n);
  for (int u = 0; u < n; u++) {
    int c = 0;
    while (c < 4 && need[u] > bounds[c]) c++;
    cls[u] = c;  // 4 = arcs stay in global memory
    if (c == 4) {
      size_t A = (size_t)(g->arc_off[a.utt0 + u + 1] - g->arc_off[a.utt0 + u]);
      if (need[u] - A * 8 > limit) return set_error(MFA_ERR_UNSUPPORTED, "utterance graph too large for the Viterbi kernel's shared memory");
    }
  }
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cls[x] != cls[y] ? cls[x] < cls[y] : work[x] > work[y]; });
  MFA_TRY(e->upload(DB_UTT_ORDER, order.data(), order.size(), &d_order));
  CUDA_TRY(cudaStreamSynchronize(e->stream));  // locals uploaded
  VitParams p;
  p.st_off = g->d_st_off; p.arc_off = g->d_arc_off; p.lp_off = g->d_lp_off; p.inb_off = g->d_inb_off;
  p.start = g->d_start; p.n_eps = g->d_n_eps; p.in_begin = g->d_in_begin; p.a_tid = g->d_a_tid; p.a_olabel = g->d_a_olabel; p.lp2pdf = g->d_lp2pdf;
  p.a_pack = g->d_a_pack; p.a_w = g->d_a_w; p.final_w = g->d_final_w;
  p.utt0 = a.utt0; p.llT = a.d_llT; p.ld = a.ld; p.col_off = a.d_col_off; p.frame_off = a.d_frame_off; p.bp_off = d_bp_off; p.word_off = a.d_word_off;
  p.bp = d_bp; p.ali = a.d_ali; p.num_words = a.d_num_words; p.words = a.d_words; p.status = a.d_status; p.per_frame = a.d_per_frame;
  p.total_like = a.d_total_like;
  p.acwt = a.opts.acoustic_scale; p.beam = a.opts.beam; p.retry_beam = a.opts.retry_beam; p.beam_delta = a.opts.beam_delta; p.min_active = a.opts.min_active;
  int pos = 0;
  CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
  // largest-need classes first: they hold the longest utterances (work ~ T * arcs), the small ones fill in around them
  for (int c = 4; c >= 0; c--) {
    int cnt = 0;
    pos = 0;
    for (int k = 0; k < n; k++) { if (cls[order[k]] < c) pos++; }
    size_t mx = 0;
    while (pos + cnt < n && cls[order[pos + cnt]] == c) {
      int u = order[pos + cnt];
      size_t nd = need[u];
      if (c == 4) nd -= (size_t)(g->arc_off[a.utt0 + u + 1] - g->arc_off[a.utt0 + u]) * 8;
      mx = std::max(mx, nd); cnt++;
    }
    if (cnt == 0) continue;
    p.order = d_order + pos;
    size_t smem = (mx + 15) / 16 * 16;
    cudaStream_t st = e->side[c];
    CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fork, 0));
    if (c < 4) {
      CUDA_TRY(cudaFuncSetAttribute(viterbi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
      viterbi_kernel<true><<<cnt, VT, smem, st>>>(p);
    } else {
      CUDA_TRY(cudaFuncSetAttribute(viterbi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
      viterbi_kernel<false><<<cnt, VT, smem, st>>>(p);
    }
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e->ev_join[c], st));
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join[c], 0));
  }
  return MFA_OK;
}

}  // namespace mfa
