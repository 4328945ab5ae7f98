// viterbi_band.cu -- K3 (primary path): beam Viterbi as a dense band recursion, one warp per utterance, the utterance's
// whole graph in shared memory.
//
// Replaces GmmAligner.align_utterance / export_alignments -> Kaldi AlignUtteranceWrapper + FasterDecoder (reference call
// sites: montreal_forced_aligner/alignment/multiprocessing.py:846-853, online/alignment.py:97-107); semantics per
// SURVEY.md A.7, identical to viterbi.cu (the sparse token-passing kernel, which stays as the path for graphs with
// input-epsilon arcs and as the fallback when a token set outgrows the band).
//
// Formulation.  mfa_graphs_pack renumbers each graph's states in a topological order of its strongly-connected-component
// DAG (graph.cc: build_band) and groups arcs by destination.  Every arc then goes forward by at most 255 positions or
// back by at most `maxback` (inside a silence model), so
//   * the lowest live state index never decreases by more than maxback, and
//   * the tokens alive under the beam sit in a narrow index window (median 29 states, maximum 170 on LibriSpeech-shaped
//     graphs) that slides through the graph as the utterance is consumed.
// Per frame the warp evaluates, for every state d of the 32-aligned window [lo - maxback, max(s + fwd(s))], the PULL
//   new(d) = min over in-arcs (s -> d) with cost(s) < cutoff of  (cost(s) + w) + (-acoustic_scale * loglike(pdf))
// lane = d mod 32, results in registers; one REDUX gives the frame's best cost; survivors (< best + adaptive beam) are
// renormalised and written to the other cost ring.  No atomics, no token lists, no hash: the only per-frame
// synchronisation is two __syncwarp.  Token costs live in two 512-entry shared-memory rings indexed by (state & 511).
// Back-pointers are ONE BYTE per window slot (index of the winning in-arc of that state, 0xFF = dead): a 256-byte row per
// frame plus the row's first group, instead of 2 bytes x all states of the graph; they are walked back in 32-frame batches
// staged through shared memory.  Acoustic costs of 4 frames x all pdfs of the utterance are prefetched two blocks ahead
// with cp.async (three stages), so the HBM latency of the log-likelihood matrix never sits on the recursion.
//
// GetCutoff (beam / min_active widening with beam_delta), the retry with retry_beam and tie-breaking (lowest by-source arc
// index; lowest original state id among equal final costs) are the same as the sparse kernel's, so both produce the same
// alignments.  A frame whose window would exceed GMAX groups hands the utterance to the sparse kernel (fallback list).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {
// Band geometry, a template parameter of everything below: GM groups of 32 band states a frame may span.  The primary kernels run
// GM = 8 (window of 256 states: median live window 29, maximum 170 at beam 10); the first fallback level re-runs an utterance whose
// window outgrew that -- in practice a retry-beam pass -- with GM = 32 (1 024 states) before the sparse kernel is asked.
template <int GM> struct BandK {
  static constexpr int GMAX = GM;            // groups of 32 band states a frame may span
  static constexpr int RS = GM * 32;         // window slots per frame
  static constexpr int ROWB = RS * 2;        // back-pointer row stride in bytes (2 bytes per slot)
  static constexpr int WRING = GM * 64;      // cost ring entries; >= RS + largest maxback (96) + 32
  static constexpr unsigned MASK = WRING - 1;
};
constexpr int GM_MAIN = 8, GM_WIDE = 32;
constexpr int BIAS = 16;             // back-pointer source-delta code = (dst - src) + BIAS, in 0..255
constexpr int NST = 2;               // acoustic-cost stages of 4 frames (a block is prefetched 4 frames ahead: several microseconds)
constexpr int BT_ROWS = 8;           // frames per back-trace batch (two staging buffers)
constexpr unsigned FULL = 0xffffffffu;
// Warps per utterance (template parameter NW = 4 or 2): warp w pulls window groups w, w + NW, ...  Four warps shorten the
// per-frame dependency chain (used where shared memory limits an SM to a few utterances); two warps issue ~35 % fewer
// instructions per frame (used where many utterances share an SM and the issue slots are the limit).

struct BandParams {
  const int64_t *st_off, *arc_off, *lp_off;
  const int32_t *b_start, *b_maxback, *a_tid, *a_olabel, *lp2pdf;
  const uint32_t *b_stw;
  const uint2 *b_arc;
  const float *b_fin;
  const uint16_t *b_arcid, *b_orig;
  int utt0;
  const int32_t *order;
  const float *llT;
  int64_t ld;
  const int64_t *col_off, *frame_off, *word_off, *ll_off, *ld_u, *bp_off;
  uint8_t *bp;   // per utterance: [T][RS] uint16 {in-arc choice, source-delta code}, then [T] uint16 first group of each row
  int32_t *ali, *num_words, *words, *status, *fallback;   // fallback[0] = count, fallback[1..] = chunk-local utterance ids | attempt at overflow << 30
  int32_t *done;   // primary launches: counter of finished CTAs (the wide level polls it); nullptr elsewhere
  float *per_frame, *total_like;
  float acwt, beam, retry_beam, beam_delta;
  int min_active, max_groups;
};

__device__ __forceinline__ uint32_t f2key(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// min_active-th (0-based) smallest live cost (Kaldi: nth_element over the token costs).  The decoder sits at this branch on
// most frames when the beam alone keeps fewer than min_active tokens, so it must be short AND shallow: the survivors' costs
// are compact per-warp segments (written while pruning the previous frame); every lane ranks its own value against all
// others with independent broadcast reads (no dependent exchange network), and the lane whose rank is `want` holds the
// answer.  More than 64 survivors: bitwise radix select.
template <int NW, int GM> struct LiveList { const float *seg; int n[NW]; };     // seg[w * SEG + j], j < n[w]; SEG = RS / NW
template <int NW, int GM> __device__ __forceinline__ float live_at(const LiveList<NW, GM> &L, int i) {
  constexpr int SEG = BandK<GM>::RS / NW;
  int w = 0;
#pragma unroll
  for (int k = 0; k < NW - 1; k++) { if (i >= L.n[k] && w == k) { i -= L.n[k]; w = k + 1; } }
  return L.seg[w * SEG + i];
}
template <int NW, int GM> __device__ __forceinline__ float select_rank(const LiveList<NW, GM> &L, int n, int want, int lane) {
  constexpr int SEG = BandK<GM>::RS / NW;
  if (n <= 32 && n - want <= 12) {
    // the usual case: a few more survivors than min_active.  The (n - want)-th largest value is the answer: peel maxima off with
    // one REDUX each (costs are >= +0, so their bit patterns order like the values; a peeled or empty lane holds 0)
    uint32_t xb = lane < n ? __float_as_uint(live_at(L, lane)) : 0u, m = 0u;
    for (int r = n - want; r > 0; r--) {
      m = __reduce_max_sync(FULL, xb);
      if (r > 1) { const unsigned who = __ballot_sync(FULL, xb == m); if (lane == __ffs(who) - 1) xb = 0u; }
    }
    return __uint_as_float(m);
  }
  if (n <= 64) {
    const float x0 = lane < n ? live_at(L, lane) : INFINITY, x1 = (n > 32 && lane + 32 < n) ? live_at(L, lane + 32) : INFINITY;
    int r0 = 0, r1 = 0, jj = 0;
    if (n <= 32) {
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const float *sg = L.seg + w * SEG;
#pragma unroll 4
        for (int j = 0; j < L.n[w]; j++, jj++) { const float y = sg[j]; r0 += (y < x0) || (y == x0 && jj < lane); }
      }
    } else {
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const float *sg = L.seg + w * SEG;
#pragma unroll 2
        for (int j = 0; j < L.n[w]; j++, jj++) {
          const float y = sg[j];
          r0 += (y < x0) || (y == x0 && jj < lane);
          r1 += (y < x1) || (y == x1 && jj < lane + 32);
        }
      }
    }
    const unsigned m0 = __ballot_sync(FULL, r0 == want && lane < n), m1 = __ballot_sync(FULL, r1 == want && lane + 32 < n);
    return m0 ? __shfl_sync(FULL, x0, __ffs(m0) - 1) : __shfl_sync(FULL, x1, __ffs(m1) - 1);
  }
  // costs are >= +0, so the raw bit patterns order like the values
  unsigned prefix = 0, mask = 0;
  for (int bit = 31; bit >= 0; bit--) {
    const unsigned b = 1u << bit;
    int c0 = 0;
    for (int i = lane; i < ((n + 31) & ~31); i += 32) {
      const unsigned x = i < n ? __float_as_uint(live_at(L, i)) : 0xFFFFFFFFu;
      c0 += __popc(__ballot_sync(FULL, (x & mask) == prefix && !(x & b)));
    }
    if (want >= c0) { prefix |= b; want -= c0; }
    mask |= b;
  }
  return __uint_as_float(prefix);
}

// pull for one band state in two halves, so that the shared-memory reads are in flight while the cutoff is being ranked:
// gather the first four in-arcs' (source cost, candidate cost) -- in-degrees are 2..5 in training graphs -- then reduce them
// under the cutoff; a loop takes any further arcs.
struct Pull { uint32_t st; uint32_t s[4]; float c[4], x[4]; };
template <int GM, class LdArc>
__device__ __forceinline__ void pull_gather(Pull &q, const uint32_t st, LdArc ld_arc, const float *__restrict__ cur,
                                            const float *__restrict__ acf, const float nacwt) {
  q.st = st;
  const int a = st & 0xFFFF, cnt = (st >> 16) & 0xFF;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    q.c[j] = INFINITY; q.x[j] = INFINITY; q.s[j] = 0;
    if (j < cnt) {
      const uint2 ar = ld_arc(a + j);
      q.s[j] = ar.x;
      q.c[j] = cur[ar.x & BandK<GM>::MASK];
      q.x[j] = __fadd_rn(__fadd_rn(q.c[j], __uint_as_float(ar.y)), __fmul_rn(nacwt, acf[(ar.x >> 16) * 4]));
    }
  }
}
template <int GM, class LdArc>
__device__ __forceinline__ void pull_reduce(const Pull &q, LdArc ld_arc, const float *__restrict__ cur, const float *__restrict__ acf,
                                            const float cutoff, const float nacwt, float &v, uint32_t &arg) {
  v = INFINITY; arg = 0xFFu;
#pragma unroll
  for (int j = 0; j < 4; j++) if (q.c[j] < cutoff && q.x[j] < v) { v = q.x[j]; arg = (uint32_t)j | ((q.s[j] & 0xFFFFu) << 8); }   // choice | source state << 8
  const int a = q.st & 0xFFFF, cnt = (q.st >> 16) & 0xFF;
  for (int j = 4; j < cnt; j++) {
    const uint2 ar = ld_arc(a + j);
    const float c = cur[ar.x & BandK<GM>::MASK];
    const float x = __fadd_rn(__fadd_rn(c, __uint_as_float(ar.y)), __fmul_rn(nacwt, acf[(ar.x >> 16) * 4]));
    if (c < cutoff && x < v) { v = x; arg = (uint32_t)j | ((ar.x & 0xFFFFu) << 8); }
  }
}

// One utterance (chunk-local id `ul`) on the calling CTA of NW warps; `bp` = its back-pointer rows; attempts start at `first_attempt`
// (the fallback level skips the beam that is already known to fail).
template <int NW, bool GS, int GM>
__device__ __forceinline__ void band_utt(const BandParams &p, const int ul, uint8_t *const bp, unsigned char *smraw, const int first_attempt,
                                         int32_t *overflow_list) {
  constexpr int GMAX = BandK<GM>::GMAX, RS = BandK<GM>::RS, ROWB = BandK<GM>::ROWB, WRING = BandK<GM>::WRING;
  constexpr unsigned MASK = BandK<GM>::MASK;
  constexpr int NT = NW * 32, GPW = GMAX / NW, SEG = GPW * 32;
  __shared__ uint32_t s_min[NW];                   // per-warp best new cost (ordered key) of the current frame
  __shared__ __align__(16) int s_stat[NW][4];      // per-warp {lowest live state, highest live state, highest reachable state, n_tot | n_beam << 16}
  __shared__ float s_live[NW * SEG];               // per-warp compact lists of the survivors' costs (for GetCutoff's rank)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ug = p.utt0 + ul;
  const int S = (int)(p.st_off[ug + 1] - p.st_off[ug]);
  const int A = (int)(p.arc_off[ug + 1] - p.arc_off[ug]);
  const int P = (int)(p.lp_off[ug + 1] - p.lp_off[ug]);
  const int64_t T = p.frame_off[ul + 1] - p.frame_off[ul];
  if (T == 0) { if (tid == 0) { p.status[ul] = MFA_ALIGN_ZERO_FRAMES; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }
  const int start = p.b_start[ug], maxback = p.b_maxback[ug];

  // GS: the utterance's graph is copied to shared memory; !GS: it stays in global memory and the (small, slowly sliding) window
  // of it that a frame touches is read through L1 -- the CTA then needs only the cost rings and the acoustic stages.
  const int Spad = (S + 32) & ~1;                               // states beyond S read as 0 (no in-arcs) up to the next group
  const uint32_t *stw;                                          // [S+]  first in-arc | in-degree << 16 | forward reach << 24
  const uint2 *arcs;                                            // [A]  {source state | local pdf << 16, weight bits}
  float *ring;                                                  // [2][WRING], 16-byte aligned
  if (GS) {
    uint32_t *sw = (uint32_t *)smraw;
    uint2 *sa = (uint2 *)(sw + Spad);
    ring = (float *)smraw + ((Spad + 2 * A + 3) & ~3);
    const uint32_t *gs = p.b_stw + p.st_off[ug];
    const uint2 *ga = p.b_arc + p.arc_off[ug];
    for (int i = tid; i < Spad; i += NT) sw[i] = i < S ? gs[i] : 0u;
    for (int i = tid; i < A; i += NT) sa[i] = ga[i];
    stw = sw; arcs = sa;
  } else {
    stw = p.b_stw + p.st_off[ug]; arcs = p.b_arc + p.arc_off[ug];
    ring = (float *)smraw;
  }
  float *ac = ring + 2 * WRING;                                 // [NST][P][4] raw log-likelihoods; later the back-trace staging area
  auto ld_st = [&](int d) -> uint32_t { return GS ? stw[d] : (d < S ? __ldg(stw + d) : 0u); };
  auto ld_arc = [&](int a) -> uint2 { return GS ? arcs[a] : __ldg(arcs + a); };
  const bool rag = p.ll_off != nullptr;
  const float *ll = p.llT + (rag ? p.ll_off[ul] : p.col_off[ul]);
  const int64_t ldu = rag ? p.ld_u[ul] : p.ld;
  const int32_t *lp2pdf = p.lp2pdf + p.lp_off[ug];
  uint16_t *bpg = (uint16_t *)(bp + (size_t)T * ROWB);      // [T] first group of each row
  const float inf = INFINITY, nacwt = -p.acwt;
  const int NB = (int)((T + 3) >> 2);
  const int max_groups = min(p.max_groups, GMAX);
  const int gend = (S - 1) >> 5;                            // last group holding a state

  auto issue_block = [&](int b, int stage) {
    if (b < NB) {
      float *dst = ac + (size_t)stage * 4 * P;
      for (int lp = tid; lp < P; lp += NT) cp_async16(dst + 4 * lp, ll + (size_t)(rag ? lp : lp2pdf[lp]) * ldu + 4 * (size_t)b);
    }
    cp_async_commit();
  };

  int result = MFA_ALIGN_NO_FINAL, over_attempt = 0;
  bool overflow = false;
  double offset = 0.0;
  int lo = 0, hi = 0;
  float *cur = ring, *nxt = ring + WRING;

  for (int attempt = first_attempt; attempt < 2 && !overflow; attempt++) {
    const float beam = attempt == 0 ? p.beam : p.retry_beam;
    if (attempt == 1 && !(p.retry_beam > 0.0f)) break;
    cp_async_wait<0>();
    __syncthreads();
    for (int i = tid; i < 2 * WRING; i += NT) ring[i] = inf;
    cur = ring; nxt = ring + WRING;
    __syncthreads();
    if (tid == 0) { cur[start & MASK] = 0.0f; s_live[0] = 0.0f; }
    lo = hi = start;
    int hib = start + (int)(ld_st(start) >> 24);
    int n_tot = 1, n_beam = 1;
    LiveList<NW, GM> L;
    L.seg = s_live; L.n[0] = 1;
#pragma unroll
    for (int w = 1; w < NW; w++) L.n[w] = 0;
    float cutoff = inf, adaptive = inf;
    offset = 0.0;
    issue_block(0, 0); issue_block(1, 1);
    cp_async_wait<1>();
    __syncthreads();
    int stage = 0;                                            // stage holding the current 4-frame block
    bool dead = false;
    for (int t = 0; t < (int)T; t++) {
      const float *acf = ac + (size_t)stage * 4 * P + (int)(t & 3);
      const int glo = max(lo - maxback, 0) >> 5;
      const int ng = min(hib >> 5, gend) - glo + 1;
      if (ng > max_groups) { overflow = true; over_attempt = attempt; break; }
      // ---- pull, first half: warp w owns groups w and w + NW of the window; the gather does not need the cutoff
      Pull q;
      pull_gather<GM>(q, warp < ng ? ld_st((glo + warp) * 32 + lane) : 0u, ld_arc, cur, acf, nacwt);
      // ---- GetCutoff on the live tokens (normalised: best == 0); every warp computes the same values
      // (a warp without a group this frame needs neither value)
      if (n_tot <= p.min_active) { cutoff = inf; adaptive = inf; }
      else if (n_beam > p.min_active) { cutoff = beam; adaptive = beam; }
      else if (warp < ng) { cutoff = select_rank(L, n_tot, p.min_active, lane); adaptive = cutoff + p.beam_delta; }
      // ---- pull, second half
      float nv[GPW];
      uint32_t na[GPW];
      pull_reduce<GM>(q, ld_arc, cur, acf, cutoff, nacwt, nv[0], na[0]);
      na[0] |= q.st & 0xFF000000u;                     // forward reach in the top byte
      uint32_t kmin = f2key(nv[0]);
#pragma unroll
      for (int k = 1; k < GPW; k++) {
        nv[k] = inf; na[k] = 0xFFu;
        if (warp + k * NW < ng) {
          Pull q2;
          pull_gather<GM>(q2, ld_st((glo + warp + k * NW) * 32 + lane), ld_arc, cur, acf, nacwt);
          pull_reduce<GM>(q2, ld_arc, cur, acf, cutoff, nacwt, nv[k], na[k]);
          na[k] |= q2.st & 0xFF000000u;
          kmin = min(kmin, f2key(nv[k]));
        }
      }
      kmin = __reduce_min_sync(FULL, kmin);
      if (lane == 0) s_min[warp] = kmin;
      __syncthreads();                                 // A: every warp has finished reading `cur` and the survivor lists; s_min complete
      {
        uint32_t m = s_min[0];
#pragma unroll
        for (int w = 1; w < NW; w++) m = min(m, s_min[w]);
        kmin = m;
      }
      const float best_new = key2f(kmin);
      if (!(best_new < inf)) { dead = true; break; }
      const float next_cutoff = best_new + adaptive;   // inf stays inf
      // ---- prune, renormalise, back-pointers, this warp's survivor list and window bounds
      uint16_t *bprow = (uint16_t *)(bp + (size_t)t * ROWB);
      int nt = 0, nb = 0, myhib = -1, flo = 0x7fffffff, fhi = -1;
#pragma unroll
      for (int k = 0; k < GPW; k++) {
        const int i = warp + k * NW;
        if (i < ng) {
          const int d = (glo + i) * 32 + lane;
          float v = nv[k];
          const bool keep = v < next_cutoff;
          v = keep ? v - best_new : inf;
          nxt[d & MASK] = v;
          cur[d & MASK] = inf;
          __stcs(bprow + i * 32 + lane, (uint16_t)(keep ? ((na[k] & 0xFFu) | ((uint32_t)(d - (int)((na[k] >> 8) & 0xFFFF) + BIAS) << 8)) : 0xFFu));
          const unsigned km = __ballot_sync(FULL, keep);
          if (keep) { s_live[warp * SEG + nt + __popc(km & ((1u << lane) - 1u))] = v; myhib = max(myhib, d + (int)(na[k] >> 24)); }
          nt += __popc(km);
          nb += __popc(__ballot_sync(FULL, keep && v <= beam));
          if (km) { flo = min(flo, (glo + i) * 32 + __ffs(km) - 1); fhi = (glo + i) * 32 + 31 - __clz(km); }
        }
      }
      myhib = __reduce_max_sync(FULL, myhib);
      if (lane == 0) *(int4 *)s_stat[warp] = make_int4(flo, fhi, myhib, nt | (nb << 16));
      if (tid == 0) bpg[t] = (uint16_t)glo;
      if ((t & 3) == 3) {                              // the block just finished frees its stage for block b + 2; nobody reads `ac` between A and B
        issue_block((t >> 2) + NST, stage);
        stage = stage == NST - 1 ? 0 : stage + 1;
        cp_async_wait<NST - 1>();                      // everything but the block just issued has landed: the next block is complete
      }
      __syncthreads();                                 // B: rings, lists, statistics and the next frame's acoustic block are visible
      {
        const int4 s0 = *(const int4 *)s_stat[0];
        lo = s0.x; hi = s0.y; hib = s0.z; n_tot = s0.w & 0xFFFF; n_beam = s0.w >> 16; L.n[0] = s0.w & 0xFFFF;
#pragma unroll
        for (int w = 1; w < NW; w++) {
          const int4 sw = *(const int4 *)s_stat[w];
          lo = min(lo, sw.x); hi = max(hi, sw.y); hib = max(hib, sw.z);
          n_tot += sw.w & 0xFFFF; n_beam += sw.w >> 16; L.n[w] = sw.w & 0xFFFF;
        }
      }
      offset += (double)best_new;
      float *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (overflow || dead) continue;
    // ---- ReachedFinal / best final token (ties: lowest original state id, like the sparse kernel); every warp redundantly
    const float *fin = p.b_fin + p.st_off[ug];
    const uint16_t *orig = p.b_orig + p.st_off[ug];
    float fv[GMAX];
    uint32_t kf = 0xFFFFFFFFu;
    const int g0 = lo >> 5;
#pragma unroll
    for (int i = 0; i < GMAX; i++) {
      const int d = (g0 + i) * 32 + lane;
      fv[i] = inf;
      if (d <= hi && d < S) { const float c = cur[d & MASK]; if (c < inf) fv[i] = c + fin[d]; }
      kf = min(kf, f2key(fv[i]));
    }
    kf = __reduce_min_sync(FULL, kf);
    const float fbest = key2f(kf);
    if (fbest < inf) {
      uint32_t sel = 0xFFFFFFFFu;
#pragma unroll
      for (int i = 0; i < GMAX; i++) {
        const int d = (g0 + i) * 32 + lane;
        if (fv[i] == fbest) sel = min(sel, ((uint32_t)orig[d] << 16) | (uint32_t)d);
      }
      sel = __reduce_min_sync(FULL, sel);
      hi = (int)(sel & 0xFFFF);    // reuse: the best final band state
      result = attempt == 0 ? MFA_ALIGN_OK : MFA_ALIGN_RETRIED;
      if (tid == 0) p.total_like[ul] = (float)(-(offset + (double)fbest) / (double)p.acwt);
      break;
    }
  }
  cp_async_wait<0>();
  __syncthreads();                 // all back-pointer rows written; nobody touches `ac` any more
  if (warp != 0) return;
  if (overflow) {
    if (lane == 0) { const int k = atomicAdd(overflow_list, 1); overflow_list[1 + k] = ul | (over_attempt << 30); }
    return;
  }
  if (result == MFA_ALIGN_NO_FINAL) {
    if (lane == 0) { p.status[ul] = result; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; }
    return;
  }
  // ---- back-trace (warp 0).  Pass 1, BT_ROWS frames at a time: a batch's rows (and their first groups) are copied to shared
  // memory with cp.async one batch ahead of use and lane 0 follows the chain using only the staged rows (state -= source
  // delta), parking (state, in-arc choice) of every frame in the `ali` output array.  Pass 2: the lanes resolve 4 x 32 frames
  // at a time -- graph, transition-id, word label and log-likelihood look-ups of different frames are independent, so their
  // latencies overlap -- and emit the outputs in frame order.
  uint8_t *stg = (uint8_t *)ac;                                    // [2][BT_ROWS][ROWB]
  uint16_t *g0s = (uint16_t *)(stg + 2 * BT_ROWS * ROWB);          // [2][BT_ROWS] (16-byte aligned)
  const int32_t *a_tid = p.a_tid + p.arc_off[ug], *a_ol = p.a_olabel + p.arc_off[ug];
  const uint16_t *arcid = p.b_arcid + p.arc_off[ug];
  int32_t *ali = p.ali + p.frame_off[ul];
  float *pf = p.per_frame + p.frame_off[ul];
  int32_t *words = p.words + p.word_off[ul];
  const int wcap = (int)(p.word_off[ul + 1] - p.word_off[ul]);
  auto stage_batch = [&](int b, int buf) {                         // frames [b * BT_ROWS, ...): whole rows, BT_ROWS * ROWB contiguous bytes
    if (b >= 0) {
      const uint8_t *src = bp + (size_t)b * BT_ROWS * ROWB;
      uint8_t *dst = stg + buf * BT_ROWS * ROWB;
      for (int i = lane; i < BT_ROWS * ROWB / 16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
      if (lane == 0) cp_async16(g0s + buf * BT_ROWS, bpg + (size_t)b * BT_ROWS);
    }
    cp_async_commit();
  };
  int s = hi, bad = 0;
  const int nbat = (int)((T + BT_ROWS - 1) / BT_ROWS);
  stage_batch(nbat - 1, (nbat - 1) & 1);
  for (int b = nbat - 1; b >= 0; b--) {
    const int buf = b & 1, r0 = b * BT_ROWS, n = min(BT_ROWS, (int)T - r0);
    stage_batch(b - 1, buf ^ 1);
    cp_async_wait<1>();
    __syncwarp();
    if (lane == 0) {
      const uint16_t *rows = (const uint16_t *)(stg + buf * BT_ROWS * ROWB);
      for (int k = n - 1; k >= 0; k--) {
        const int slot = s - 32 * (int)g0s[buf * BT_ROWS + k];
        const unsigned w = (slot >= 0 && slot < RS) ? rows[k * RS + slot] : 0xFFu;
        if ((w & 0xFFu) == 0xFFu) { bad = 1; break; }    // cannot happen
        ali[r0 + k] = s | (int)((w & 0xFFu) << 16);
        s -= (int)(w >> 8) - BIAS;
      }
    }
    bad = __shfl_sync(FULL, bad, 0);
    if (bad) break;
    __syncwarp();
  }
  cp_async_wait<0>();
  if (bad) { if (lane == 0) { p.status[ul] = MFA_ALIGN_NO_FINAL; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }
  __syncwarp();                    // lane 0's parked codes are visible to the warp (same CTA, global memory)
  int nw = 0;
  for (int t0 = 0; t0 < (int)T; t0 += 128) {
    int code[4], tid_[4], ol[4];
    float lk[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int t = t0 + 32 * u + lane; code[u] = t < (int)T ? __ldcg(ali + t) : -1; }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      tid_[u] = 0; ol[u] = 0; lk[u] = 0.0f;
      if (code[u] >= 0) {
        const int t = t0 + 32 * u + lane;
        const int j = (int)(ld_st(code[u] & 0xFFFF) & 0xFFFF) + (code[u] >> 16);
        const int arc = arcid[j];
        const int lp = (int)(ld_arc(j).x >> 16);
        tid_[u] = a_tid[arc]; ol[u] = a_ol[arc];
        lk[u] = ll[(size_t)(rag ? lp : lp2pdf[lp]) * ldu + t];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int t = t0 + 32 * u + lane;
      if (code[u] >= 0) { ali[t] = tid_[u]; pf[t] = lk[u]; }
      const unsigned m = __ballot_sync(FULL, ol[u] != 0);
      if (ol[u] != 0) { const int idx = nw + __popc(m & ((1u << lane) - 1u)); if (idx < wcap) words[idx] = ol[u]; }
      nw += __popc(m);
    }
  }
  if (lane == 0) { p.status[ul] = result; p.num_words[ul] = nw; }
}

#ifndef MFA_BAND_MINB2
#define MFA_BAND_MINB2 10
#endif
template <int NW, bool GS>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 4 : MFA_BAND_MINB2)
viterbi_band_kernel(BandParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int ul = p.order[blockIdx.x];
  band_utt<NW, GS, GM_MAIN>(p, ul, p.bp + p.bp_off[ul], smraw, 0, p.fallback);
  // (thread 0 sits in warp 0, the last to leave band_utt: its list entry, if any, is written)
  if (threadIdx.x == 0 && p.done) { __threadfence(); atomicAdd(p.done, 1); }
}

// First fallback level, always enqueued with the primary launches and running NEXT to them: the utterances on the device-side list
// `fb` ({count, ids | attempt << 30}: their live window outgrew 8 groups) run again on the same recursion with a 32-group window, four
// warps, starting at the beam that overflowed.  The CTAs poll the list while the primary CTAs are still running (ctl[0] counts the
// finished ones, ctl[1] is the consumption cursor), so an overflowing utterance starts its second pass the moment it overflows, not
// after the join of all classes: it is by construction a long utterance with a wide beam, i.e. the launch's critical path.  What
// outgrows even 32 groups goes on `fb2` for the sparse kernel.  The count is stored to a host-mapped slot -- the host never
// synchronises inside a step.  (The kernel is enqueued AFTER the primary launches: tools that serialise kernels then still terminate.)
__device__ __forceinline__ int ld_vol(const int32_t *p) { return *(const volatile int32_t *)p; }
__global__ void __launch_bounds__(128, 2)
viterbi_band_wide_kernel(BandParams p, const int32_t *fb, int32_t *__restrict__ fb2, int64_t slab, int32_t *h_count, int32_t *ctl, int total) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ int s_entry;
  for (;;) {
    if (threadIdx.x == 0) {
      int entry = -1;
      unsigned long long t0 = 0;
      for (;;) {
        const int n = ld_vol(fb), c = ld_vol(ctl + 1);
        if (c < n) {
          if (atomicCAS(ctl + 1, c, c + 1) != c) continue;
          while ((entry = ld_vol(fb + 1 + c)) < 0) __nanosleep(100);   // the count is bumped before the entry is stored
          break;
        }
        if (ld_vol(ctl) >= total) {            // every primary CTA has finished (fence + atomic on their side): the count is final
          __threadfence();
          if (ld_vol(ctl + 1) >= ld_vol(fb)) break;
          continue;
        }
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > 20000000000ull) __trap();   // 20 s without a primary CTA finishing: fail loudly rather than spin
        __nanosleep(2000);
      }
      if (entry < 0) { *(volatile int32_t *)h_count = ld_vol(fb); __threadfence_system(); }
      s_entry = entry;
    }
    __syncthreads();
    const int e = s_entry;
    __syncthreads();
    if (e < 0) return;
    band_utt<4, false, GM_WIDE>(p, e & 0x3FFFFFFF, p.bp + (size_t)blockIdx.x * slab, smraw, (e >> 30) & 1, fb2);
    __syncthreads();
  }
}

}  // namespace

namespace mfa {

size_t viterbi_band_smem(int64_t S, int64_t A, int64_t P, bool graph_in_smem, bool wide) {
  const int64_t graph = graph_in_smem ? ((((S + 32) & ~(int64_t)1) + 2 * A + 3) & ~(int64_t)3) * 4 : 0;
  const int64_t wring = wide ? BandK<GM_WIDE>::WRING : BandK<GM_MAIN>::WRING, rowb = wide ? BandK<GM_WIDE>::ROWB : BandK<GM_MAIN>::ROWB;
  return (size_t)(graph + 2 * wring * 4 + std::max<int64_t>(NST * 16 * P, 2 * BT_ROWS * rowb + 4 * BT_ROWS) + 16);
}

// Launches the band kernel for the utterances in `subset` (chunk-local ids, all band_ok).  d_fallback: [1 + n_utts] ints, count
// first; cleared here.
int launch_viterbi_band(mfa_engine *e, const ViterbiArgs &a, const std::vector<int32_t> &subset, int max_groups, int32_t *d_fallback, int32_t *d_ctl) {
  const mfa_graphs *g = a.g;
  const int n = a.n_utts, ns = (int)subset.size();
  CUDA_TRY(cudaMemsetAsync(d_fallback, 0xFF, ((size_t)n + 1) * sizeof(int32_t), e->stream));   // entries: -1 = not written yet
  CUDA_TRY(cudaMemsetAsync(d_fallback, 0, sizeof(int32_t), e->stream));
  CUDA_TRY(cudaMemsetAsync(d_ctl, 0, 2 * sizeof(int32_t), e->stream));
  if (ns == 0) return MFA_OK;
  const size_t limit = e->smem_optin - 4096;   // the kernel also has ~2 KB of static shared memory
  const bool graph_smem = e->cfg.vit_graph_smem != 0;
  std::vector<int64_t> bp_off(n + 1, 0);
  std::vector<size_t> need(n, 0);
  std::vector<int64_t> work(n, 0);
  int64_t bp_total = 0;
  for (int ul : subset) {
    const int ug = a.utt0 + ul;
    const int64_t S = g->st_off[ug + 1] - g->st_off[ug], A = g->arc_off[ug + 1] - g->arc_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug];
    const int64_t T = a.h_frame_off[ul + 1] - a.h_frame_off[ul];
    bp_off[ul] = bp_total;
    bp_total += (T * BandK<GM_MAIN>::ROWB + T * 2 + 15) / 16 * 16;
    need[ul] = viterbi_band_smem(S, A, P, graph_smem, false);
    work[ul] = T;
    if (need[ul] > limit) return set_error(MFA_ERR_UNSUPPORTED, "internal: band utterance exceeds shared memory");
  }
  uint8_t *d_bp; int64_t *d_bp_off; int32_t *d_order;
  MFA_TRY(e->getT<uint8_t>(DB_BBP, (size_t)bp_total + BT_ROWS * BandK<GM_MAIN>::ROWB + 64, &d_bp));   // the back-trace stages whole batches of rows
  MFA_TRY(e->upload(DB_BBP_OFF, bp_off.data(), bp_off.size(), &d_bp_off));
  constexpr int NC = mfa_engine::kSide;
  size_t bounds[NC];
  for (int c = 0; c < NC; c++) bounds[c] = std::min<size_t>(limit, (size_t)(14336.0 * std::pow(2.0, 0.34 * c)));   // 14 KB ... 190 KB
  bounds[NC - 1] = limit;
  std::vector<int> cls(n, 0);
  for (int ul : subset) { int c = 0; while (c < NC - 1 && need[ul] > bounds[c]) c++; cls[ul] = c; }
  std::vector<int32_t> order(subset);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cls[x] != cls[y] ? cls[x] < cls[y] : work[x] > work[y]; });
  MFA_TRY(e->upload(DB_BORDER, order.data(), order.size(), &d_order));
  BandParams p;
  p.st_off = g->d_st_off; p.arc_off = g->d_arc_off; p.lp_off = g->d_lp_off;
  p.b_start = g->d_b_start; p.b_maxback = g->d_b_maxback; p.a_tid = g->d_a_tid; p.a_olabel = g->d_a_olabel; p.lp2pdf = g->d_lp2pdf;
  p.b_stw = g->d_b_stw; p.b_arc = (const uint2 *)g->d_b_arc; p.b_fin = g->d_b_fin; p.b_arcid = g->d_b_arcid; p.b_orig = g->d_b_orig;
  p.utt0 = a.utt0; p.llT = a.d_llT; p.ld = a.ld; p.col_off = a.d_col_off; p.frame_off = a.d_frame_off; p.word_off = a.d_word_off;
  p.ll_off = a.d_ll_off; p.ld_u = a.d_ld_u; p.bp_off = d_bp_off; p.bp = d_bp;
  p.ali = a.d_ali; p.num_words = a.d_num_words; p.words = a.d_words; p.status = a.d_status; p.fallback = d_fallback; p.done = d_ctl;
  p.per_frame = a.d_per_frame; p.total_like = a.d_total_like;
  p.acwt = a.opts.acoustic_scale; p.beam = a.opts.beam; p.retry_beam = a.opts.retry_beam; p.beam_delta = a.opts.beam_delta;
  p.min_active = a.opts.min_active; p.max_groups = std::max(1, std::min(max_groups, GM_MAIN));
  // shared-memory share of the unified L1 (percent).  Graph through L1, 10 h workload: 25 -> 16.8 ms, 50 -> 9.4, 65 -> 7.6, 75 -> 7.6,
  // 88 -> 8.8, 100 -> 10.5: enough shared memory for ~9 resident utterances per SM, the rest as L1 for their graph windows
  const int carve = e->cfg.vit_carveout_band >= 0 ? e->cfg.vit_carveout_band : (graph_smem ? 100 : 70);
  for (auto fn : {(const void *)viterbi_band_kernel<4, true>, (const void *)viterbi_band_kernel<2, true>, (const void *)viterbi_band_kernel<4, false>,
                  (const void *)viterbi_band_kernel<2, false>}) {
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  }
  // two-warp CTAs where many utterances share an SM, four-warp CTAs for the size classes above a shared-memory threshold
  // (engine option vit_nw2_kb overrides it): 44 KB when the graph is copied to shared memory; 20 KB when it is read through L1 -- a CTA then needs
  // ~15 KB + 32 B per pdf of its graph, so the classes above 20 KB are the longest utterances, the ones on the launch's critical path,
  // and four warps shorten their per-frame chain while the bulk keeps the residency of two-warp CTAs.  Measured on the 10 h config-2
  // workload, graph through L1 (three boxes): all classes 2 warps 9.4-9.6 ms, all 4 warps 10.0 ms, threshold 20 KB 8.2-8.45 ms
  // (19 KB 8.15, 21 KB 9.5-9.6, 22 KB 8.2-9.3: the launch is bounded by a handful of utterances, so neighbouring thresholds scatter);
  // graph in shared memory: 11.7 ms.
  const size_t nw2_below = (size_t)(e->cfg.vit_nw2_kb >= 0 ? e->cfg.vit_nw2_kb : (graph_smem ? 44 : 20)) * 1024;
  CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
  for (int c = NC - 1; c >= 0; c--) {
    int pos = 0, cnt = 0;
    for (int k = 0; k < ns; k++) if (cls[order[k]] < c) pos++;
    size_t mx = 0;
    while (pos + cnt < ns && cls[order[pos + cnt]] == c) { mx = std::max(mx, need[order[pos + cnt]]); cnt++; }
    if (cnt == 0) continue;
    p.order = d_order + pos;
    cudaStream_t st = e->side[c];
    CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fork, 0));
    const size_t sm = (mx + 15) / 16 * 16;
    if (graph_smem) { if (mx <= nw2_below) viterbi_band_kernel<2, true><<<cnt, 64, sm, st>>>(p); else viterbi_band_kernel<4, true><<<cnt, 128, sm, st>>>(p); }
    else { if (mx <= nw2_below) viterbi_band_kernel<2, false><<<cnt, 64, sm, st>>>(p); else viterbi_band_kernel<4, false><<<cnt, 128, sm, st>>>(p); }
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e->ev_join[c], st));
    CUDA_TRY(cudaStreamWaitEvent(e->sj, e->ev_join[c], 0));
  }
  return MFA_OK;
}

// Wide-band fallback level over the device-side list d_fb (see viterbi_band_wide_kernel); d_fb2 receives what overflows again.
int launch_viterbi_band_wide(mfa_engine *e, const ViterbiArgs &a, const std::vector<int32_t> &subset, const int32_t *d_fb, int32_t *d_fb2,
                             int32_t *h_count, int32_t *d_ctl) {
  const mfa_graphs *g = a.g;
  CUDA_TRY(cudaMemsetAsync(d_fb2, 0, sizeof(int32_t), e->stream));
  if (subset.empty()) return MFA_OK;
  const int kCtas = std::max(1, std::min(e->cfg.vit_wide_ctas, 64));
  size_t smem = 0; int64_t slab = 0;
  for (int ul : subset) {
    const int ug = a.utt0 + ul;
    const int64_t S = g->st_off[ug + 1] - g->st_off[ug], A = g->arc_off[ug + 1] - g->arc_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug];
    const int64_t T = a.h_frame_off[ul + 1] - a.h_frame_off[ul];
    smem = std::max(smem, viterbi_band_smem(S, A, P, false, true));
    slab = std::max(slab, (T * BandK<GM_WIDE>::ROWB + T * 2 + 15) / 16 * 16 + BT_ROWS * BandK<GM_WIDE>::ROWB + 64);
  }
  const int ctas = (int)std::min<size_t>(kCtas, subset.size());
  uint8_t *d_bp;
  MFA_TRY(e->getT<uint8_t>(DB_WIDE_BP, (size_t)slab * ctas + 64, &d_bp));
  BandParams p;
  p.st_off = g->d_st_off; p.arc_off = g->d_arc_off; p.lp_off = g->d_lp_off;
  p.b_start = g->d_b_start; p.b_maxback = g->d_b_maxback; p.a_tid = g->d_a_tid; p.a_olabel = g->d_a_olabel; p.lp2pdf = g->d_lp2pdf;
  p.b_stw = g->d_b_stw; p.b_arc = (const uint2 *)g->d_b_arc; p.b_fin = g->d_b_fin; p.b_arcid = g->d_b_arcid; p.b_orig = g->d_b_orig;
  p.utt0 = a.utt0; p.order = nullptr; p.llT = a.d_llT; p.ld = a.ld; p.col_off = a.d_col_off; p.frame_off = a.d_frame_off; p.word_off = a.d_word_off;
  p.ll_off = a.d_ll_off; p.ld_u = a.d_ld_u; p.bp_off = nullptr; p.bp = d_bp;
  p.ali = a.d_ali; p.num_words = a.d_num_words; p.words = a.d_words; p.status = a.d_status; p.fallback = d_fb2; p.done = nullptr;
  p.per_frame = a.d_per_frame; p.total_like = a.d_total_like;
  p.acwt = a.opts.acoustic_scale; p.beam = a.opts.beam; p.retry_beam = a.opts.retry_beam; p.beam_delta = a.opts.beam_delta;
  p.min_active = a.opts.min_active; p.max_groups = GM_WIDE;
  smem = (smem + 15) / 16 * 16;
  if (smem + 8192 > e->smem_optin) return set_error(MFA_ERR_UNSUPPORTED, "internal: wide-band utterance exceeds shared memory");
  CUDA_TRY(cudaFuncSetAttribute(viterbi_band_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Polling mode: its own highest-priority stream, ordered behind the main stream only (list resets, slab allocation): the CTAs become
  // resident as soon as primary CTAs retire and poll from then on; the join stream waits for it before the sparse level.  Otherwise the
  // kernel goes on the join stream behind all classes (it then finds the list complete).  Default: poll in device-buffer calls, where
  // nothing else of this engine competes for the SMs the pollers hold; host-buffer calls are issued by several jobs (engines) per GPU at
  // once, and resident pollers of one job measurably slow the other jobs' launches (end to end, 2 jobs: 1.31 -> 1.15 M x RT).
  const bool poll = e->cfg.vit_wide_poll >= 0 ? e->cfg.vit_wide_poll != 0 : !a.host_call;
  cudaStream_t st = poll ? e->sw : e->sj;
  CUDA_TRY(cudaEventRecord(e->ev_fb, e->stream));
  CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fb, 0));
  viterbi_band_wide_kernel<<<ctas, 128, smem, st>>>(p, d_fb, d_fb2, slab, h_count, d_ctl, (int)subset.size());
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  if (poll) {
    CUDA_TRY(cudaEventRecord(e->ev_wide, e->sw));
    CUDA_TRY(cudaStreamWaitEvent(e->sj, e->ev_wide, 0));
  }
  return MFA_OK;
}

}  // namespace mfa
