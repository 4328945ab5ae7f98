// internal.h -- structures shared by the translation units behind include/mfa_b200.h.
#pragma once
#include <cstdint>
#include <string>
#include <memory>
#include <utility>
#include <vector>

#include "../../include/mfa_b200.h"

namespace mfa {
int set_error(int code, const std::string &msg);
// Recycled device blocks for the per-batch graph mirrors (hundreds of MB each, one per new batch): cudaMalloc / cudaFree have a long
// latency tail on this platform and cudaFree synchronises the device, so a destroyed graphs object parks its blocks here and the next
// upload takes one that is large enough (after a device synchronisation: the previous user's kernels may still be running).
void *dev_cache_take(int device, size_t bytes, size_t *cap);   // nullptr when nothing fits
void dev_cache_give(int device, void *p, size_t cap);          // may cudaFree (the cache holds a few blocks per process)

// std::vector whose resize() leaves trivially constructible elements uninitialised.  The per-batch graph arrays are hundreds of MB that
// worker threads fill completely: a zero-filling resize touched every page on ONE thread first (180 ms of page faults per 10 h batch);
// now the first touch happens in the parallel fill.
template <class T> struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
  template <class U, class... A> void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
    else ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T> using BigVec = std::vector<T, NoInitAlloc<T>>;
}

// Decoder-ready batch of graphs (host copy; device mirror uploaded on first use by an engine).
// Arcs of utterance u are arc_off[u]..arc_off[u+1]-1, grouped by SOURCE state (OpenFst order inside a state):
// the out-arcs of local state s are in_begin[inb_off[u]+s] .. in_begin[inb_off[u]+s+1]-1 (utterance-local indices).
struct mfa_graphs {
  int32_t n_utts = 0;
  std::vector<int64_t> st_off, arc_off, lp_off, inb_off;
  std::vector<int32_t> start, n_eps, max_words;
  mfa::BigVec<uint32_t> h_barc, h_apack;   // device images built at pack time (per utterance, in parallel): {band arc word, weight} pairs; dst | local pdf << 16
  std::vector<uint8_t> too_large;   // graph beyond the 16-bit packed views: packed as an empty graph, status MFA_ALIGN_GRAPH_TOO_LARGE
  int32_t n_too_large = 0;
  mfa::BigVec<int32_t> in_begin;             // [sum(S_u+1)]
  mfa::BigVec<int32_t> a_src, a_dst, a_lp, a_tid, a_olabel;  // per arc; a_lp = local pdf index or -1 (epsilon input)
  mfa::BigVec<float> a_w;                    // graph weight + AddTransitionProbs cost
  mfa::BigVec<float> final_w;                // per state, +inf = non-final
  mfa::BigVec<int32_t> lp2pdf;               // per utterance local pdf list (sorted pdf ids)
  // Band view (viterbi_band.cu): states renumbered in a topological order of the strongly-connected-component DAG (self-loops and
  // the small cycles of Kaldi's silence topology stay inside a component), arcs grouped by DESTINATION state in that order.
  // b_stw[st_off[u]+s]  = first in-arc (low 16) | in-degree << 16 | largest forward jump of s's out-arcs << 24
  // b_apk[arc_off[u]+j] = source band state (low 16) | local pdf << 16;  b_arcid = index of the same arc in the by-source arrays
  std::vector<int32_t> band_ok, b_start, b_maxback;   // per utterance; band_ok = 0: graph only runs on the sparse kernel
  mfa::BigVec<uint32_t> b_stw, b_apk;
  mfa::BigVec<float> b_aw, b_fin;
  mfa::BigVec<uint16_t> b_arcid, b_orig;              // b_orig: band state -> original state id
  // device mirror
  int device = -1;
  void *d_blob = nullptr;
  size_t d_bytes = 0;
  size_t d_blob_cap = 0;          // bytes d_blob was allocated with (blocks are recycled through mfa::dev_cache_*)
  int64_t *d_st_off = nullptr, *d_arc_off = nullptr, *d_lp_off = nullptr, *d_inb_off = nullptr;
  int32_t *d_start = nullptr, *d_n_eps = nullptr, *d_in_begin = nullptr, *d_a_tid = nullptr, *d_a_olabel = nullptr, *d_lp2pdf = nullptr, *d_a_src = nullptr;
  uint32_t *d_a_pack = nullptr;  // dst (low 16) | lp (high 16, 0xFFFF = epsilon)
  int32_t *d_b_start = nullptr, *d_b_maxback = nullptr;
  uint32_t *d_b_stw = nullptr;
  void *d_b_arc = nullptr;       // uint2 per arc: {b_apk, bits of b_aw}
  float *d_b_fin = nullptr;
  uint16_t *d_b_arcid = nullptr, *d_b_orig = nullptr;
  float *d_a_w = nullptr, *d_final_w = nullptr;
  // graph weights WITHOUT the AddTransitionProbs cost (a_w = a_w0 + tid_cost[a_tid]): what mfa_graphs_set_transitions re-folds on the
  // device when the transition model has been re-estimated.  After a re-fold the host copies of a_w / b_aw are stale.
  int32_t num_tids = 0;
  mfa::BigVec<float> a_w0;
  float *d_a_w0 = nullptr;
  bool host_w_stale = false;
  // per-utterance Gaussian tiling for the ragged K2 path (cached against the model's tiling version)
  uint64_t rag_version = 0;
  std::vector<int64_t> rag_tile_off;
  void *d_rag = nullptr;
  size_t rag_meta_bytes = 0;
  ~mfa_graphs();
};
