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
// Formulation: sparse token passing.  One warp (= one CTA of 32 threads) per utterance keeps in shared memory a compact list of live states with their
// costs, and per frame (a) pushes every live token under the cutoff over its out-arcs with one 64-bit atomicMin per arc on a
// packed (ordered cost, arc index) word per destination state -- the first toucher appends the state to the next list --,
// (b) block-reduces the best new cost, (c) prunes against best + adaptive_beam, renormalises, writes the surviving
// back-pointers and compacts the list, (d) runs GetCutoff on the compact list (histogram + exact rank for the min_active-th
// cost).  Work per frame is proportional to the live tokens, not to the graph.  Arcs are read from global memory (L1/L2
// resident, only the live region of the graph is touched); an 8-frame block of the utterance's acoustic costs is staged in
// shared memory from the pdf-major log-likelihood matrix (one 32-byte sector per pdf per block); back-pointers (uint16 arc
// index per frame x state, only live entries written) go to HBM and are walked back by one thread at the end.
// Token costs are kept relative to the frame's best token (fp32) with the running offset in fp64, which is
// as accurate as FasterDecoder's double-precision token costs at the magnitudes that matter for comparisons.
//
// Differences from FasterDecoder that cannot change a surviving best path: (1) new tokens are pruned against the
// final per-frame cutoff rather than the cutoff as it tightens in hash-list order (a superset of tokens can exist in
// Kaldi for one frame; they are pruned by the next GetCutoff unless fewer than min_active tokens are in the beam);
// (2) ties between equal-cost arcs into a state resolve to the lowest arc index.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {
constexpr int VT = 32;   // one warp per utterance: every barrier is a __syncwarp, every reduction a shuffle
constexpr unsigned kNoArc = 0xFFFFu;
constexpr unsigned kEps = 0xFFFFu;

struct VitParams {
  const int64_t *st_off, *arc_off, *lp_off, *inb_off;
  const int32_t *start, *n_eps, *in_begin, *a_tid, *a_olabel, *lp2pdf, *a_src;
  const uint32_t *a_pack;
  const float *a_w, *final_w;
  int utt0;
  const int32_t *order;
  const float *llT;
  int64_t ld;
  const int64_t *col_off, *frame_off, *bp_off, *word_off, *ll_off, *ld_u;   // ll_off/ld_u non-null: per-utterance [local pdf][ld_u] blocks
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

// monotone map float -> uint32 (valid for any finite value and +inf) and back
__device__ __forceinline__ uint32_t f2key(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
constexpr unsigned long long kEmpty = ~0ull;

// one utterance (chunk-local id `ul`) on the calling warp; `bp` = its back-pointer rows
__device__ __forceinline__ void viterbi_utt(const VitParams &p, const int ul, uint16_t *const bp, unsigned char *smraw) {
  __shared__ int sh_cnt, sh_nnext, sh_best_state, sh_bin, sh_rank, sh_head;
  __shared__ float sh_sel;
  __shared__ int sh_hist[128];
  __shared__ float sh_cand[64];

  const int tid = threadIdx.x, warp = 0, lane = tid;
  const int ug = p.utt0 + ul;
  const int S = (int)(p.st_off[ug + 1] - p.st_off[ug]);
  const int P = (int)(p.lp_off[ug + 1] - p.lp_off[ug]);
  const int64_t T = p.frame_off[ul + 1] - p.frame_off[ul];
  const int start = p.start[ug];
  const bool has_eps = p.n_eps[ug] > 0;
  if (start < 0 || S == 0) { if (tid == 0) { p.status[ul] = MFA_ALIGN_EMPTY_GRAPH; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }
  if (T == 0) { if (tid == 0) { p.status[ul] = MFA_ALIGN_ZERO_FRAMES; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; } return; }

  unsigned long long *nxt = (unsigned long long *)smraw;   // [S] packed (ordered cost << 32 | arc), kEmpty = untouched
  float *cost = (float *)(nxt + S);                         // [S] normalised cost of live states (stale elsewhere)
  float *ac = cost + S;                                     // [4][P]
  uint16_t *list_a = (uint16_t *)(ac + 4 * P);              // live states (current frame)
  uint16_t *list_b = list_a + ((S + 1) & ~1);               // states touched while expanding
  const int32_t *outb = p.in_begin + p.inb_off[ug];         // out-arc offsets [S+1]
  const uint32_t *pack = p.a_pack + p.arc_off[ug];          // dst | lp << 16
  const float *aw = p.a_w + p.arc_off[ug];
  const float *fin = p.final_w + p.st_off[ug];
  const int32_t *lp2pdf = p.lp2pdf + p.lp_off[ug];
  const bool rag = p.ll_off != nullptr;
  const float *ll = p.llT + (rag ? p.ll_off[ul] : p.col_off[ul]);
  const int64_t ldu = rag ? p.ld_u[ul] : p.ld;
  const float inf = INFINITY;

  int result = MFA_ALIGN_NO_FINAL;
  double offset = 0.0;
  int n_cur = 0;

  for (int attempt = 0; attempt < 2; attempt++) {
    const float beam = attempt == 0 ? p.beam : p.retry_beam;
    if (attempt == 1 && !(p.retry_beam > 0.0f)) break;
    __syncwarp();
    offset = 0.0;
    for (int s = tid; s < S; s += VT) nxt[s] = kEmpty;
    if (has_eps) for (int s = tid; s < S; s += VT) bp[(size_t)T * S + s] = (uint16_t)kNoArc;
    if (tid == 0) { list_a[0] = (uint16_t)start; cost[start] = 0.0f; }
    n_cur = 1;
    __syncwarp();
    float cutoff = inf, adaptive = inf;
    int n_tot = 1, n_beam = 1;
    if (has_eps) {
      // ProcessNonemitting(+inf) from the start state: label-correcting relaxation over epsilon arcs.
      // Live states are marked in nxt with their current (cost, arc) so improvements are detected by atomicMin.
      if (tid == 0) { nxt[start] = ((unsigned long long)f2key(0.0f) << 32) | kNoArc; sh_cnt = 1; sh_head = 0; }
      __syncwarp();
      for (;;) {
        const int head = sh_head, n = sh_cnt;
        __syncwarp();
        if (head >= n) break;
        for (int i = head + tid; i < n; i += VT) {
          const int s = list_a[i];
          const float c = key2f((uint32_t)(nxt[s] >> 32));
          for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
            const uint32_t pk = pack[a];
            if ((pk >> 16) != kEps) continue;
            const int d = pk & 0xFFFF;
            const unsigned long long old = atomicMin(&nxt[d], ((unsigned long long)f2key(c + aw[a]) << 32) | (unsigned)a);
            if (old == kEmpty) list_a[atomicAdd(&sh_cnt, 1)] = (uint16_t)d;
          }
        }
        __syncwarp();
        if (tid == 0) sh_head = n;
        __syncwarp();
      }
      // a state's cost may have improved after it was expanded: iterate the whole list to a fixed point
      for (;;) {
        __syncwarp();
        if (tid == 0) sh_head = 0;
        __syncwarp();
        const int n = sh_cnt;
        for (int i = tid; i < n; i += VT) {
          const int s = list_a[i];
          const float c = key2f((uint32_t)(nxt[s] >> 32));
          for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
            const uint32_t pk = pack[a];
            if ((pk >> 16) != kEps) continue;
            const int d = pk & 0xFFFF;
            const unsigned long long cand = ((unsigned long long)f2key(c + aw[a]) << 32) | (unsigned)a;
            const unsigned long long old = atomicMin(&nxt[d], cand);
            if (cand < old) { sh_head = 1; if (old == kEmpty) list_a[atomicAdd(&sh_cnt, 1)] = (uint16_t)d; }
          }
        }
        __syncwarp();
        if (!sh_head) break;
      }
      n_cur = sh_cnt;
      int ct = 0, cb = 0;
      for (int i = tid; i < n_cur; i += VT) {
        const int s = list_a[i];
        const unsigned long long w = nxt[s];
        const float v = key2f((uint32_t)(w >> 32));
        cost[s] = v; bp[(size_t)T * S + s] = (uint16_t)(w & 0xFFFF);
        nxt[s] = kEmpty;
        ct++; if (v <= beam) cb++;
      }
      n_tot = warp_sum(ct); n_beam = warp_sum(cb);
      __syncwarp();
    }
    bool dead = false;
    for (int64_t t = 0; t < T; t++) {
      if (!has_eps && n_cur <= 32) {
        // ===== fast frame (the common case): at most one live token per lane, kept in registers; reductions by REDUX / ballot,
        // compaction by ballot prefix; same semantics as the general frame below =====
        const unsigned full = 0xffffffffu;
        const int s = lane < n_cur ? (int)list_a[lane] : 0;
        const float c = lane < n_cur ? cost[s] : inf;
        if (n_tot <= p.min_active) { cutoff = inf; adaptive = inf; }
        else if (n_beam > p.min_active) { cutoff = beam; adaptive = beam; }
        else {
          int r = 0;
#pragma unroll
          for (int j = 0; j < 32; j++) { const float y = __shfl_sync(full, c, j); r += (y < c) || (y == c && j < lane); }
          const unsigned m = __ballot_sync(full, r == p.min_active && lane < n_cur);
          cutoff = __shfl_sync(full, c, __ffs(m) - 1); adaptive = cutoff + p.beam_delta;
        }
        if ((t & 3) == 0) {
          __syncwarp();
          for (int lp = tid; lp < P; lp += VT) {
            const float4 v0 = __ldcs((const float4 *)(ll + (size_t)(rag ? lp : lp2pdf[lp]) * ldu + t));
            ac[0 * P + lp] = -p.acwt * v0.x; ac[1 * P + lp] = -p.acwt * v0.y; ac[2 * P + lp] = -p.acwt * v0.z; ac[3 * P + lp] = -p.acwt * v0.w;
          }
        }
        if (lane == 0) sh_cnt = 0;
        __syncwarp();
        const float *acf = ac + (int)(t & 3) * P;
        uint16_t *bprow = bp + (size_t)t * S;
        uint32_t *key32 = (uint32_t *)nxt, *arc32 = key32 + S;
        int a0 = 0, a1 = 0;
        if (c < cutoff) { a0 = outb[s]; a1 = outb[s + 1]; }
        for (int a = a0; a < a1; a++) {
          const uint32_t pk = pack[a];
          const float v = (c + aw[a]) + acf[pk >> 16];
          const int d = pk & 0xFFFF;
          if (atomicMin(&key32[d], f2key(v)) == 0xFFFFFFFFu) list_b[atomicAdd(&sh_cnt, 1)] = (uint16_t)d;
        }
        __syncwarp();
        const int n_new = sh_cnt;
        uint32_t kmin = 0xFFFFFFFFu;
        for (int i = lane; i < n_new; i += 32) kmin = min(kmin, key32[list_b[i]]);
        kmin = __reduce_min_sync(full, kmin);
        const float best_new = key2f(kmin);
        if (!(best_new < inf)) { dead = true; break; }
        const float next_cutoff = best_new + adaptive;
        for (int a = a0; a < a1; a++) {
          const uint32_t pk = pack[a];
          const float v = (c + aw[a]) + acf[pk >> 16];
          const int d = pk & 0xFFFF;
          if (v < next_cutoff && f2key(v) == key32[d]) atomicMin(&arc32[d], (uint32_t)a);
        }
        __syncwarp();
        int base = 0, nb = 0;
        for (int i0 = 0; i0 < n_new; i0 += 32) {
          const int i = i0 + lane;
          bool keep = false; int d = 0; float v = inf; uint32_t a = 0;
          if (i < n_new) {
            d = list_b[i]; v = key2f(key32[d]); a = arc32[d];
            key32[d] = 0xFFFFFFFFu; arc32[d] = 0xFFFFFFFFu;
            keep = v < next_cutoff;
          }
          const unsigned km = __ballot_sync(full, keep);
          if (keep) { v -= best_new; cost[d] = v; bprow[d] = (uint16_t)a; list_a[base + __popc(km & ((1u << lane) - 1u))] = (uint16_t)d; }
          nb += __popc(__ballot_sync(full, keep && v <= beam));
          base += __popc(km);
        }
        __syncwarp();
        n_tot = base; n_beam = nb; n_cur = base;
        offset += (double)best_new;
        continue;
      }
      // ---- GetCutoff for the live tokens (normalised: best == 0)
      if (n_tot <= p.min_active) { cutoff = inf; adaptive = inf; }
      else if (n_beam > p.min_active) { cutoff = beam; adaptive = beam; }
      else {
        // min_active-th order statistic (0-based) of the live costs (Kaldi: nth_element on the token costs).
        if (n_cur <= 64) {
          // the common case: each lane holds two costs, ranks them against all others by shuffle broadcast
          const float x0 = lane < n_cur ? cost[list_a[lane]] : inf, x1 = lane + 32 < n_cur ? cost[list_a[lane + 32]] : inf;
          int r0 = 0, r1 = 0;
          for (int j = 0; j < n_cur; j++) {
            const float y = __shfl_sync(0xffffffffu, j < 32 ? x0 : x1, j & 31);
            r0 += (y < x0) || (y == x0 && j < lane);
            r1 += (y < x1) || (y == x1 && j < lane + 32);
          }
          const unsigned m0 = __ballot_sync(0xffffffffu, r0 == p.min_active && lane < n_cur);
          const unsigned m1 = __ballot_sync(0xffffffffu, r1 == p.min_active && lane + 32 < n_cur);
          cutoff = m0 ? __shfl_sync(0xffffffffu, x0, __ffs(m0) - 1) : __shfl_sync(0xffffffffu, x1, __ffs(m1) - 1);
          adaptive = cutoff + p.beam_delta;
        } else {
        // many live tokens: 128-bin histogram over [0, max] -> the bin holding that rank -> exact rank inside the bin.
        float lmx = 0.0f;
        for (int i = tid; i < n_cur; i += VT) lmx = fmaxf(lmx, cost[list_a[i]]);
#pragma unroll
        for (int o = 16; o; o >>= 1) lmx = fmaxf(lmx, __shfl_xor_sync(0xffffffffu, lmx, o));
        for (int i = tid; i < 128; i += VT) sh_hist[i] = 0;
        if (tid == 0) sh_cnt = 0;
        __syncwarp();
        const float vmax = lmx;
        const float scale = 128.0f / (vmax * 1.000001f + 1e-30f);
        for (int i = tid; i < n_cur; i += VT) atomicAdd(&sh_hist[min(127, (int)(cost[list_a[i]] * scale))], 1);
        __syncwarp();
        if (warp == 0) {
          int c0 = sh_hist[4 * lane], c1 = sh_hist[4 * lane + 1], c2 = sh_hist[4 * lane + 2], c3 = sh_hist[4 * lane + 3];
          int tot = c0 + c1 + c2 + c3, incl = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
          int excl = incl - tot;
          const int k = p.min_active;
          if (excl <= k && k < incl) {  // exactly one lane
            int b = 4 * lane, e = excl;
            if (k >= e + c0) { e += c0; b++; if (k >= e + c1) { e += c1; b++; if (k >= e + c2) { e += c2; b++; } } }
            sh_bin = b; sh_rank = k - e;
          }
        }
        __syncwarp();
        const int bin = sh_bin;
        for (int i = tid; i < n_cur; i += VT) {
          const float v = cost[list_a[i]];
          if (min(127, (int)(v * scale)) == bin) { int idx = atomicAdd(&sh_cnt, 1); if (idx < 64) sh_cand[idx] = v; }
        }
        __syncwarp();
        const int nc = sh_cnt;
        if (nc <= 64) {
          if (warp == 0) {
            const int want = sh_rank;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int i = lane + 32 * h;
              if (i < nc) {
                const float x = sh_cand[i];
                int rk = 0;
                for (int jx = 0; jx < nc; jx++) { float y = sh_cand[jx]; rk += (y < x) || (y == x && jx < i); }
                if (rk == want) sh_sel = x;
              }
            }
          }
          __syncwarp();
        } else {
          // many tokens share one bin (clustered costs): exact bitwise radix select over the live costs (warp 0)
          if (warp == 0) {
            unsigned prefix = 0, mask = 0;
            int want = p.min_active;
            for (int bit = 31; bit >= 0; bit--) {
              unsigned b = 1u << bit;
              int c0 = 0;
              for (int i = lane; i < n_cur; i += 32) { unsigned x = __float_as_uint(cost[list_a[i]]); if ((x & mask) == prefix && !(x & b)) c0++; }
              c0 = warp_sum(c0);
              if (want >= c0) { prefix |= b; want -= c0; }
              mask |= b;
            }
            if (lane == 0) sh_sel = __uint_as_float(prefix);
          }
          __syncwarp();
        }
        cutoff = sh_sel; adaptive = cutoff + p.beam_delta;
        __syncwarp();
        }
      }
      // ---- acoustic costs for frames t..t+3 (one 16-byte streaming load per pdf: the matrix is read exactly once)
      if ((t & 3) == 0) {
        __syncwarp();
        for (int lp = tid; lp < P; lp += VT) {
          const float4 v0 = __ldcs((const float4 *)(ll + (size_t)(rag ? lp : lp2pdf[lp]) * ldu + t));
          ac[0 * P + lp] = -p.acwt * v0.x; ac[1 * P + lp] = -p.acwt * v0.y; ac[2 * P + lp] = -p.acwt * v0.z; ac[3 * P + lp] = -p.acwt * v0.w;
        }
      }
      if (tid == 0) sh_cnt = 0;
      __syncwarp();
      const float *acf = ac + (int)(t & 3) * P;
      uint16_t *bprow = bp + (size_t)t * S;
      // ---- ProcessEmitting: push every live token under the cutoff over its emitting out-arcs
      float best_new;
      int ct = 0, cb = 0;
      if (!has_eps) {
        // fast path (no input-epsilon arcs): native 32-bit shared-memory atomics in two sweeps over the live tokens --
        // (1) atomicMin of the ordered cost per destination, first toucher appends the state; (2) the arc whose cost equals the
        // winning cost records itself (lowest arc index among ties).  key32 / arc32 alias the two halves of `nxt`.
        uint32_t *key32 = (uint32_t *)nxt, *arc32 = key32 + S;
        for (int i = tid; i < n_cur; i += VT) {
          const int s = list_a[i];
          const float c = cost[s];
          if (!(c < cutoff)) continue;
          for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
            const uint32_t pk = pack[a];
            const float v = (c + aw[a]) + acf[pk >> 16];
            const int d = pk & 0xFFFF;
            if (atomicMin(&key32[d], f2key(v)) == 0xFFFFFFFFu) list_b[atomicAdd(&sh_cnt, 1)] = (uint16_t)d;
          }
        }
        __syncwarp();
        const int n_new = sh_cnt;
        float lmin = inf;
        for (int i = tid; i < n_new; i += VT) lmin = fminf(lmin, key2f(key32[list_b[i]]));
        best_new = warp_min(lmin);
        if (tid == 0) sh_nnext = 0;
        if (!(best_new < inf)) { dead = true; break; }
        const float next_cutoff = best_new + adaptive;  // inf stays inf
        for (int i = tid; i < n_cur; i += VT) {
          const int s = list_a[i];
          const float c = cost[s];
          if (!(c < cutoff)) continue;
          for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
            const uint32_t pk = pack[a];
            const float v = (c + aw[a]) + acf[pk >> 16];
            const int d = pk & 0xFFFF;
            if (v < next_cutoff && f2key(v) == key32[d]) atomicMin(&arc32[d], (uint32_t)a);
          }
        }
        __syncwarp();
        // prune, renormalise, record back-pointers, compact into list_a
        for (int i = tid; i < n_new; i += VT) {
          const int d = list_b[i];
          float v = key2f(key32[d]);
          const uint32_t a = arc32[d];
          key32[d] = 0xFFFFFFFFu; arc32[d] = 0xFFFFFFFFu;
          if (v < next_cutoff) {
            v -= best_new;
            cost[d] = v; bprow[d] = (uint16_t)a;
            list_a[atomicAdd(&sh_nnext, 1)] = (uint16_t)d;
            ct++; if (v <= beam) cb++;
          }
        }
      } else {
        // general path (graphs with input-epsilon arcs, e.g. read from a Kaldi fsts.ark): 64-bit packed (cost, arc) atomics
        for (int i = tid; i < n_cur; i += VT) {
          const int s = list_a[i];
          const float c = cost[s];
          if (!(c < cutoff)) continue;
          for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
            const uint32_t pk = pack[a];
            const unsigned lp = pk >> 16;
            if (lp == kEps) continue;
            const float v = (c + aw[a]) + acf[lp];
            const int d = pk & 0xFFFF;
            const unsigned long long old = atomicMin(&nxt[d], ((unsigned long long)f2key(v) << 32) | (unsigned)a);
            if (old == kEmpty) list_b[atomicAdd(&sh_cnt, 1)] = (uint16_t)d;
          }
        }
        __syncwarp();
        const int n_new = sh_cnt;
        float lmin = inf;
        for (int i = tid; i < n_new; i += VT) lmin = fminf(lmin, key2f((uint32_t)(nxt[list_b[i]] >> 32)));
        best_new = warp_min(lmin);
        if (tid == 0) sh_nnext = 0;
        __syncwarp();
        if (!(best_new < inf)) { dead = true; break; }
        const float next_cutoff = best_new + adaptive;  // inf stays inf
        // keep survivors marked in nxt (normalised), run ProcessNonemitting(next_cutoff) to a fixed point
        for (int i = tid; i < n_new; i += VT) {
          const int d = list_b[i];
          const unsigned long long w = nxt[d];
          const float v = key2f((uint32_t)(w >> 32));
          if (v < next_cutoff) {
            nxt[d] = ((unsigned long long)f2key(v - best_new) << 32) | (w & 0xFFFFFFFFull);
            list_a[atomicAdd(&sh_nnext, 1)] = (uint16_t)d;
          } else nxt[d] = kEmpty;
        }
        for (;;) {
          __syncwarp();
          if (tid == 0) sh_head = 0;
          __syncwarp();
          const int n = sh_nnext;
          for (int i = tid; i < n; i += VT) {
            const int s = list_a[i];
            const float c = key2f((uint32_t)(nxt[s] >> 32));
            if (!(c <= adaptive)) continue;
            for (int a = outb[s], a1 = outb[s + 1]; a < a1; a++) {
              const uint32_t pk = pack[a];
              if ((pk >> 16) != kEps) continue;
              const float v = c + aw[a];
              if (!(v <= adaptive)) continue;
              const int d = pk & 0xFFFF;
              const unsigned long long cand = ((unsigned long long)f2key(v) << 32) | (unsigned)a;
              const unsigned long long old = atomicMin(&nxt[d], cand);
              if (cand < old) { sh_head = 1; if (old == kEmpty) list_a[atomicAdd(&sh_nnext, 1)] = (uint16_t)d; }
            }
          }
          __syncwarp();
          if (!sh_head) break;
        }
        const int n = sh_nnext;
        for (int i = tid; i < n; i += VT) {
          const int d = list_a[i];
          const unsigned long long w = nxt[d];
          nxt[d] = kEmpty;
          const float v = key2f((uint32_t)(w >> 32));
          cost[d] = v; bprow[d] = (uint16_t)(w & 0xFFFF);
          ct++; if (v <= beam) cb++;
        }
      }
      n_tot = warp_sum(ct); n_beam = warp_sum(cb);
      __syncwarp();
      n_cur = sh_nnext;
      offset += (double)best_new;
    }
    if (dead) {
      // leave nxt clean for the retry
      __syncwarp();
      continue;
    }
    // ---- ReachedFinal / best final token
    float lbest = inf;
    for (int i = tid; i < n_cur; i += VT) { const int s = list_a[i]; lbest = fminf(lbest, cost[s] + fin[s]); }
    const float fbest = warp_min(lbest);
    if (tid == 0) sh_best_state = 0x7fffffff;
    __syncwarp();
    if (fbest < inf) {
      for (int i = tid; i < n_cur; i += VT) { const int s = list_a[i]; if (cost[s] + fin[s] == fbest) atomicMin(&sh_best_state, s); }
      __syncwarp();
      result = attempt == 0 ? MFA_ALIGN_OK : MFA_ALIGN_RETRIED;
      if (tid == 0) p.total_like[ul] = (float)(-(offset + (double)fbest) / (double)p.acwt);
      break;
    }
    __syncwarp();
  }
  __syncwarp();  // all back-pointer writes of this CTA are visible to thread 0 below (same CTA, global memory)
  if (tid != 0) return;
  p.status[ul] = result;
  if (result == MFA_ALIGN_NO_FINAL) { p.num_words[ul] = 0; p.total_like[ul] = 0.0f; return; }
  // ---- back-trace (one thread): frames T-1..0, then the initial epsilon closure
  const int32_t *a_tid = p.a_tid + p.arc_off[ug], *a_ol = p.a_olabel + p.arc_off[ug], *a_src = p.a_src + p.arc_off[ug];
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
    const uint32_t pk = pack[a];
    int ol = a_ol[a];
    if (ol != 0) { if (nw < wcap) words[nw] = ol; nw++; }
    s = a_src[a];
    if ((pk >> 16) == kEps) { if (++guard > S) { p.status[ul] = MFA_ALIGN_NO_FINAL; p.num_words[ul] = 0; return; } continue; }
    guard = 0;
    ali[t] = a_tid[a];
    pf[t] = ll[(size_t)(rag ? (int)(pk >> 16) : lp2pdf[pk >> 16]) * ldu + t];
    t--;
  }
  if (has_eps) {
    guard = 0;
    while (s != start && guard++ <= S) {
      unsigned a = bp[(size_t)T * S + s];
      if (a == kNoArc) break;
      int ol = a_ol[a];
      if (ol != 0) { if (nw < wcap) words[nw] = ol; nw++; }
      s = a_src[a];
    }
  }
  int n = nw < wcap ? nw : wcap;
  for (int i = 0; i < n / 2; i++) { int32_t x = words[i]; words[i] = words[n - 1 - i]; words[n - 1 - i] = x; }
  p.num_words[ul] = nw;
}

__global__ void __launch_bounds__(VT, 16)
viterbi_kernel(VitParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int ul = p.order[blockIdx.x];
  viterbi_utt(p, ul, p.bp + p.bp_off[ul], smraw);   // rows 0..T-1 (+ row T: initial epsilon closure)
}

// Fallback pass behind the band kernel: `fb` = {count, utterance ids ...} was filled on the device by the band kernel for utterances
// whose live window outgrew the band.  Always launched with a small fixed grid -- the count is read HERE, on the device, so the host
// never synchronises inside a step; each CTA owns one back-pointer slab and walks the list with a grid stride.  The count is also
// stored to a host-mapped slot so that mfa_engine_band_fallbacks can report it later.
__global__ void __launch_bounds__(VT, 16)
viterbi_fallback_kernel(VitParams p, const int32_t *__restrict__ fb, int64_t slab, const unsigned char *__restrict__ too_big, int32_t *h_count) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int n = fb[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) { *(volatile int32_t *)h_count = n; __threadfence_system(); }
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int ul = fb[1 + i] & 0x3FFFFFFF;   // (the top bits carry the attempt at which the band overflowed: the sparse kernel starts over)
    if (too_big[ul]) {   // graph beyond the sparse kernel's shared memory: cannot be retried (reported as a failed alignment)
      if (threadIdx.x == 0) { p.status[ul] = MFA_ALIGN_NO_FINAL; p.num_words[ul] = 0; p.total_like[ul] = 0.0f; }
    } else {
      viterbi_utt(p, ul, p.bp + (size_t)blockIdx.x * slab, smraw);
    }
    __syncwarp();
  }
}

}  // namespace

namespace mfa {

static VitParams sparse_params(const ViterbiArgs &a) {
  const mfa_graphs *g = a.g;
  VitParams p{};
  p.st_off = g->d_st_off; p.arc_off = g->d_arc_off; p.lp_off = g->d_lp_off; p.inb_off = g->d_inb_off;
  p.start = g->d_start; p.n_eps = g->d_n_eps; p.in_begin = g->d_in_begin; p.a_tid = g->d_a_tid; p.a_olabel = g->d_a_olabel; p.lp2pdf = g->d_lp2pdf;
  p.a_src = g->d_a_src; p.a_pack = g->d_a_pack; p.a_w = g->d_a_w; p.final_w = g->d_final_w;
  p.utt0 = a.utt0; p.llT = a.d_llT; p.ld = a.ld; p.col_off = a.d_col_off; p.frame_off = a.d_frame_off; p.word_off = a.d_word_off;
  p.ali = a.d_ali; p.num_words = a.d_num_words; p.words = a.d_words; p.status = a.d_status; p.per_frame = a.d_per_frame;
  p.total_like = a.d_total_like;
  p.ll_off = a.d_ll_off; p.ld_u = a.d_ld_u;
  p.acwt = a.opts.acoustic_scale; p.beam = a.opts.beam; p.retry_beam = a.opts.retry_beam; p.beam_delta = a.opts.beam_delta; p.min_active = a.opts.min_active;
  return p;
}

// Sparse kernel over the utterances in `subset` (chunk-local ids).
static int launch_viterbi_sparse(mfa_engine *e, const ViterbiArgs &a, const std::vector<int32_t> &subset) {
  const mfa_graphs *g = a.g;
  const int n = a.n_utts, ns = (int)subset.size();
  if (ns == 0) return MFA_OK;
  if (a.ld % 8 != 0) return set_error(MFA_ERR_INVALID, "log-likelihood leading dimension must be a multiple of 8");
  // shared-memory need and back-pointer offsets per utterance
  std::vector<int64_t> bp_off(n + 1, 0);
  std::vector<size_t> need(n, 0);
  std::vector<double> work(n, 0.0);
  const size_t limit = e->smem_optin - 2048;
  int64_t bp_total = 0;
  for (int u : subset) {
    int ug = a.utt0 + u;
    int64_t S = g->st_off[ug + 1] - g->st_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug];
    int64_t T = a.h_frame_off[u + 1] - a.h_frame_off[u];
    bp_off[u] = bp_total;
    bp_total += (T + (g->n_eps[ug] > 0 ? 1 : 0)) * S;
    need[u] = (size_t)S * 12 + (size_t)P * 16 + (size_t)((S + 1) & ~1) * 4 + 16;
    work[u] = (double)T;
    if (need[u] > limit) return set_error(MFA_ERR_UNSUPPORTED, "utterance graph too large for the Viterbi kernel's shared memory");
  }
  uint16_t *d_bp; int64_t *d_bp_off; int32_t *d_order;
  MFA_TRY(e->getT<uint16_t>(DB_BP, (size_t)bp_total + 8, &d_bp));
  MFA_TRY(e->upload(DB_BP_OFF, bp_off.data(), bp_off.size(), &d_bp_off));
  // classes by shared-memory need (occupancy); inside a class, longest utterance first
  constexpr int NC = mfa_engine::kSide;
  size_t bounds[NC];
  for (int c = 0; c < NC; c++) bounds[c] = std::min<size_t>(limit, (size_t)(7168.0 * std::pow(2.0, 0.5 * c)));   // 7, 9.9, 14, ... KB
  bounds[NC - 1] = limit;
  std::vector<int> cls(n, 0);
  for (int u : subset) { int c = 0; while (c < NC - 1 && need[u] > bounds[c]) c++; cls[u] = c; }
  std::vector<int32_t> order(subset);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cls[x] != cls[y] ? cls[x] < cls[y] : work[x] > work[y]; });
  MFA_TRY(e->upload(DB_UTT_ORDER, order.data(), order.size(), &d_order));
  VitParams p = sparse_params(a);
  p.bp_off = d_bp_off; p.bp = d_bp;
  CUDA_TRY(cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
  {
    // split of the unified L1/shared array (percent shared): the arcs of the live tokens are re-read every frame through L1/L2
    // measured (10 h config-2 workload): 100 -> 60.1 ms/step, 75 -> 60.8, 50 -> 69.9, 25 -> 112.4: resident warps matter more than L1
    CUDA_TRY(cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, e->cfg.vit_carveout));
  }
  CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
  // largest-need classes first: they hold the biggest graphs; the small ones fill in around them on the side streams
  for (int c = NC - 1; c >= 0; c--) {
    int pos = 0, cnt = 0;
    for (int k = 0; k < ns; k++) { if (cls[order[k]] < c) pos++; }
    size_t mx = 0;
    while (pos + cnt < ns && cls[order[pos + cnt]] == c) { mx = std::max(mx, need[order[pos + cnt]]); cnt++; }
    if (cnt == 0) continue;
    p.order = d_order + pos;
    size_t smem = (mx + 15) / 16 * 16;
    cudaStream_t st = e->side[c];
    CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fork, 0));
    viterbi_kernel<<<cnt, VT, smem, st>>>(p);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e->ev_join[c], st));
    CUDA_TRY(cudaStreamWaitEvent(e->sj, e->ev_join[c], 0));
  }
  return MFA_OK;
}

// K3 dispatch: graphs with a band view run on the band kernel (viterbi_band.cu); graphs without one (input-epsilon arcs, very
// high in-degree) run on the sparse kernel, and so does any utterance whose live window outgrew the band at run time -- the
// band kernel lists those on the device and a small fallback launch, always enqueued, walks that list (usually empty) without the
// host ever reading the count inside the step.
// Engine option vit_band = 0 forces the sparse kernel; vit_maxgroups = k (1..8) narrows the band (tests use it to exercise the fallback).
int launch_viterbi(mfa_engine *e, const ViterbiArgs &a_in) {
  const mfa_graphs *g = a_in.g;
  const int n = a_in.n_utts;
  if (n == 0) return MFA_OK;
  MFA_TRY(e->join_k3());   // one Viterbi launch in flight per engine: its scratch buffers are reused here
  // K3 is not joined on the main stream (cuda_internal.cuh: join_k3), so the next call may overwrite the small offset arrays of THIS call
  // (they live in per-engine upload slots) while the kernels still read them: the launch works on private copies
  ViterbiArgs a = a_in;
  {
    const bool rag = a_in.d_ll_off != nullptr;
    int64_t *k3o;
    MFA_TRY(e->getT<int64_t>(DB_K3_OFFS, (size_t)4 * ((size_t)n + 1), &k3o));
    CUDA_TRY(cudaMemcpyAsync(k3o, a_in.d_frame_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(k3o + (n + 1), a_in.d_word_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToDevice, e->stream));
    a.d_frame_off = k3o; a.d_word_off = k3o + (n + 1);
    if (rag) {
      CUDA_TRY(cudaMemcpyAsync(k3o + 2 * (n + 1), a_in.d_ll_off, (size_t)n * 8, cudaMemcpyDeviceToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(k3o + 3 * (n + 1), a_in.d_ld_u, (size_t)n * 8, cudaMemcpyDeviceToDevice, e->stream));
      a.d_ll_off = k3o + 2 * (n + 1); a.d_ld_u = k3o + 3 * (n + 1);
    } else {
      CUDA_TRY(cudaMemcpyAsync(k3o + 2 * (n + 1), a_in.d_col_off, (size_t)n * 8, cudaMemcpyDeviceToDevice, e->stream));
      a.d_col_off = k3o + 2 * (n + 1);
    }
  }
  if (a.ld % 8 != 0) return set_error(MFA_ERR_INVALID, "log-likelihood leading dimension must be a multiple of 8");
  if (!a.d_ll_off) {
    for (int u = 0; u < n; u++) {
      const int64_t T = a.h_frame_off[u + 1] - a.h_frame_off[u];
      if (a.h_col_off[u] % 4 != 0) return set_error(MFA_ERR_INVALID, "col_off must be a multiple of 4");
      if (a.h_col_off[u] + ((T + 3) / 4) * 4 > a.ld) return set_error(MFA_ERR_INVALID, "log-likelihood leading dimension too small for 4-frame blocks");
    }
  }
  const bool use_band = e->cfg.vit_band != 0;
  const int max_groups = e->cfg.vit_maxgroups;
  std::vector<int32_t> band, sparse;
  for (int u = 0; u < n; u++) {
    const int ug = a.utt0 + u;
    bool ok = use_band && g->band_ok[ug];
    if (ok) {
      const int64_t S = g->st_off[ug + 1] - g->st_off[ug], A = g->arc_off[ug + 1] - g->arc_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug];
      ok = viterbi_band_smem(S, A, P, e->cfg.vit_graph_smem != 0) <= e->smem_optin - 4096;
    }
    (ok ? band : sparse).push_back(u);
  }
  int32_t *d_fb = nullptr, *d_ctl = nullptr;
  if (!band.empty()) {
    MFA_TRY(e->getT<int32_t>(DB_FALLBACK, (size_t)n + 1, &d_fb));
    MFA_TRY(e->getT<int32_t>(DB_FB_CTL, 2, &d_ctl));
    MFA_TRY(launch_viterbi_band(e, a, band, max_groups, d_fb, d_ctl));
  }
  MFA_TRY(launch_viterbi_sparse(e, a, sparse));
  if (!band.empty()) {
    // fallback levels over device-side lists (no host read of a count: see viterbi_band_wide_kernel / viterbi_fallback_kernel):
    // band (8 groups) -> d_fb -> wide band (32 groups) -> d_fb2 -> sparse kernel
    if (e->fb_pending + 2 > mfa_engine::kFbRing) { CUDA_TRY(cudaStreamSynchronize(e->stream)); e->harvest_fallbacks(); }
    const int32_t *d_list = d_fb;
    if (e->cfg.vit_wide) {
      int32_t *d_fb2;
      MFA_TRY(e->getT<int32_t>(DB_FALLBACK2, (size_t)n + 1, &d_fb2));
      MFA_TRY(launch_viterbi_band_wide(e, a, band, d_fb, d_fb2, e->h_fb_ring + e->fb_pending, d_ctl));
      e->fb_pending++;
      d_list = d_fb2;
    }
    constexpr int kFbCtas = 16;
    const size_t limit = e->smem_optin - 2048;
    std::vector<unsigned char> too_big(n, 0);
    size_t smem = 0; int64_t slab = 0;
    for (int u : band) {
      const int ug = a.utt0 + u;
      const int64_t S = g->st_off[ug + 1] - g->st_off[ug], P = g->lp_off[ug + 1] - g->lp_off[ug], T = a.h_frame_off[u + 1] - a.h_frame_off[u];
      const size_t need = (size_t)S * 12 + (size_t)P * 16 + (size_t)((S + 1) & ~1) * 4 + 16;
      if (need > limit) { too_big[u] = 1; continue; }
      smem = std::max(smem, need); slab = std::max(slab, (T + 1) * S);
    }
    slab = (slab + 7) & ~(int64_t)7;
    const int ctas = (int)std::min<size_t>(kFbCtas, band.size());
    uint16_t *d_bp; unsigned char *d_big;
    MFA_TRY(e->getT<uint16_t>(DB_FB_BP, (size_t)slab * ctas + 8, &d_bp));
    MFA_TRY(e->upload(DB_FB_BIG, too_big.data(), too_big.size(), &d_big));
    VitParams p = sparse_params(a);
    p.bp = d_bp;
    smem = (smem + 15) / 16 * 16;
    CUDA_TRY(cudaFuncSetAttribute(viterbi_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    // with the wide level in front, band_fallbacks counts the utterances that left the 8-group band; the sparse level's own count goes to
    // a scratch slot of the ring that is never harvested
    CUDA_TRY(cudaEventRecord(e->ev_fb, e->stream));    // the uploads above precede the kernel on the join stream
    CUDA_TRY(cudaStreamWaitEvent(e->sj, e->ev_fb, 0));
    viterbi_fallback_kernel<<<ctas, VT, smem, e->sj>>>(p, d_list, slab, d_big, e->cfg.vit_wide ? e->h_fb_ring + mfa_engine::kFbRing : e->h_fb_ring + e->fb_pending);
    if (!e->cfg.vit_wide) e->fb_pending++;
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaEventRecord(e->ev_fb, e->stream));
  CUDA_TRY(cudaStreamWaitEvent(e->sj, e->ev_fb, 0));   // (a launch without any class still orders the join stream behind the main stream)
  if (g->n_too_large) {
    // graphs that were packed as empty because they exceed the 16-bit views: their own status instead of EMPTY_GRAPH, after the kernels
    static const int32_t kTooLarge = MFA_ALIGN_GRAPH_TOO_LARGE;
    for (int u = 0; u < n; u++)
      if (g->too_large[a.utt0 + u]) CUDA_TRY(cudaMemcpyAsync(a.d_status + u, &kTooLarge, sizeof(int32_t), cudaMemcpyHostToDevice, e->sj));
  }
  CUDA_TRY(cudaEventRecord(e->ev_k3_done, e->sj));
  e->k3_pending = true;
  for (int i = 0; i < 6; i++) e->pend_out[i] = nullptr;   // the caller that defers the join says which buffers are being written
  return MFA_OK;
}

}  // namespace mfa
