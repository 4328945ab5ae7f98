// graph.cc -- host-side training-graph compiler and decoder-ready graph packing.
//
// Replaces kalpy TrainingGraphCompiler.compile_fst / export_graphs (reference call sites:
// montreal_forced_aligner/alignment/multiprocessing.py:537-571, online/alignment.py:77-96), i.e. Kaldi's
// decoder/training-graph-compiler.cc pipeline  L o G -> context -> H -> self-loops(reorder=true),
// WITHOUT OpenFst: for a linear transcript the composition has a closed form, so the graph is
// constructed directly.  The weighted set of (transition-id sequence, word sequence) paths equals
// Kaldi's; state numbering and weight placement along a path are not (no determinise/minimise pass),
// see DESIGN.md.
//
// Structure:
//   phone graph   : per word boundary i the lexicon states NS_i (no silence), P_i (before optional
//                   silence), S_i (after silence); pronunciation chains between boundaries
//                   (tests/data/dictionaries/expected/lexicon.text.fst documents the layout).
//   instances     : (phone arc, left phone, right phone) -- triphone context (N=3,P=1) or the arc alone (N=1).
//   HMM expansion : node (instance, hmm state j', source state j) carries the self-loop of j
//                   (reorder=true: [forward tid, self-loop tid x (n-1)]); nodes with j' = final are the
//                   junctions from which the successors' state-0 transitions leave.
#include <atomic>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/mfa_b200.h"
#include "internal.h"

namespace mfa {

static const float kInf = std::numeric_limits<float>::infinity();

struct GraphCompiler {
  // topology
  std::vector<int32_t> phone2entry, entry_state_off, fwd_class, self_class, trans_off, trans_dst;
  // tuples
  std::vector<int32_t> tuples, first_tid;
  std::unordered_map<uint64_t, std::vector<std::pair<uint64_t, int32_t>>> tstate_map;
  // tree
  int ctx_width = 1, central = 0, tree_root = 0;
  std::vector<int32_t> tree_nodes, tree_aux_off, tree_aux;
  // lexicon
  std::vector<int32_t> word_pron_off, pron_phone_off, pron_phones;
  std::vector<float> pron_cost, sil_after, nonsil_after, sil_before, nonsil_before;
  int sil_phone = 1;
  float init_sil = 0, init_nonsil = 0, final_sil = 0, final_nonsil = 0;

  int tree_lookup(const int *ctx, int pdf_class) const {
    int n = tree_root;
    for (;;) {
      const int32_t *nd = &tree_nodes[4 * n];
      if (nd[0] == 0) return nd[2];
      int key = nd[1];
      int v;
      if (key == -1) v = pdf_class;
      else if (key >= 0 && key < ctx_width) v = ctx[key];
      else return -1;
      const int32_t *aux = &tree_aux[tree_aux_off[n]];
      int na = tree_aux_off[n + 1] - tree_aux_off[n];
      if (nd[0] == 1) {
        bool yes = std::binary_search(aux, aux + na, v);
        n = yes ? nd[2] : nd[3];
      } else {
        if (v < 0 || v >= na || aux[v] < 0) return -1;
        n = aux[v];
      }
      if (n < 0) return -1;
    }
  }

  int find_tstate(int phone, int hs, int fpdf, int spdf) const {
    uint64_t k1 = ((uint64_t)(uint32_t)phone << 32) | (uint32_t)hs;
    auto it = tstate_map.find(k1);
    if (it == tstate_map.end()) return -1;
    uint64_t k2 = ((uint64_t)(uint32_t)fpdf << 32) | (uint32_t)spdf;
    for (auto &p : it->second) if (p.first == k2) return p.second;
    return -1;
  }
};

struct PhoneArc { int src, dst, phone, olabel; float cost; };

struct FstBuilder {
  std::vector<int32_t> src, dst, il, ol;
  std::vector<float> w, finals;
  int add_state() { finals.push_back(kInf); return (int)finals.size() - 1; }
  void add_arc(int s, int d, int i, int o, float c) { src.push_back(s); dst.push_back(d); il.push_back(i); ol.push_back(o); w.push_back(c); }
};

struct FstBatch {
  std::vector<int64_t> state_off{0}, arc_off{0};
  std::vector<int32_t> start;
  BigVec<int32_t> src, dst, il, ol;   // (filled completely by worker threads: see NoInitAlloc)
  BigVec<float> finals, w;
  int32_t n() const { return (int32_t)start.size(); }
};

// One utterance.  Returns "" or an error string.
// transition-states of a phone in context, (left, phone, right) -> one per HMM state: six decision-tree walks and three map look-ups per
// phone instance were most of the compile time; a worker thread keeps what it has resolved (contexts repeat across utterances)
using CtxCache = std::unordered_map<uint64_t, std::vector<int>>;

static std::string compile_one(const GraphCompiler &C, const int32_t *words, int64_t nw, FstBuilder &out, int &start_state, CtxCache &ctx_cache) {
  // ---- phone graph --------------------------------------------------------------------------
  // node ids: 0 = Start; per boundary i in 0..nw: NS_i = 1+3i, P_i = 2+3i, S_i = 3+3i; then chain nodes.
  std::vector<PhoneArc> arcs;
  int n_nodes = 1 + 3 * (int)(nw + 1);
  auto NS = [](int i) { return 1 + 3 * i; };
  auto PP = [](int i) { return 2 + 3 * i; };
  auto SS = [](int i) { return 3 + 3 * i; };
  for (int i = 0; i <= nw; i++) arcs.push_back({PP(i), SS(i), C.sil_phone, 0, 0.0f});
  int nwords_tab = (int)C.word_pron_off.size() - 1;
  for (int i = 0; i < nw; i++) {
    int wid = words[i];
    if (wid < 0 || wid >= nwords_tab || C.word_pron_off[wid] == C.word_pron_off[wid + 1])
      return "word id " + std::to_string(wid) + " has no pronunciation";
    for (int pr = C.word_pron_off[wid]; pr < C.word_pron_off[wid + 1]; pr++) {
      int a = C.pron_phone_off[pr], b = C.pron_phone_off[pr + 1], len = b - a;
      if (len <= 0) return "empty pronunciation";
      // entry variants: from NS_i (non-silence before) and S_i (silence before); exit variants: to NS_{i+1} / P_{i+1}
      // interior chain shared: nodes c_1..c_{len-1}
      std::vector<int> chain(len + 1, -1);
      for (int k = 1; k < len; k++) chain[k] = n_nodes++;
      for (int k = 0; k < len; k++) {
        int ph = C.pron_phones[a + k];
        std::vector<std::pair<int, float>> srcs, dsts;
        if (k == 0) { srcs.push_back({NS(i), C.pron_cost[pr] + C.nonsil_before[pr]}); srcs.push_back({SS(i), C.pron_cost[pr] + C.sil_before[pr]}); }
        else srcs.push_back({chain[k], 0.0f});
        if (k == len - 1) { dsts.push_back({NS(i + 1), C.nonsil_after[pr]}); dsts.push_back({PP(i + 1), C.sil_after[pr]}); }
        else dsts.push_back({chain[k + 1], 0.0f});
        for (auto &s : srcs) for (auto &d : dsts) arcs.push_back({s.first, d.first, ph, k == 0 ? wid : 0, s.second + d.second});
      }
    }
  }
  std::vector<float> node_final(n_nodes, kInf);
  node_final[NS((int)nw)] = C.final_nonsil;
  node_final[SS((int)nw)] = C.final_sil;
  // Start: silence into S_0 with init_sil; copies of NS_0's out-arcs with init_nonsil added.
  {
    size_t na = arcs.size();
    arcs.push_back({0, SS(0), C.sil_phone, 0, C.init_sil});
    for (size_t k = 0; k < na; k++) if (arcs[k].src == NS(0)) { PhoneArc c = arcs[k]; c.src = 0; c.cost += C.init_nonsil; arcs.push_back(c); }
    if (nw == 0) node_final[0] = C.init_nonsil + C.final_nonsil;
  }
  // out-adjacency
  std::vector<std::vector<int>> out_arcs(n_nodes);
  for (int k = 0; k < (int)arcs.size(); k++) out_arcs[arcs[k].src].push_back(k);

  // ---- instances + HMM expansion ------------------------------------------------------------
  const bool tri = C.ctx_width == 3;
  struct Inst { int arc, l, r; std::vector<int> junctions; };  // junction graph nodes (hmm final reached)
  std::vector<Inst> insts;
  std::unordered_map<uint64_t, int> inst_id;   // (phone arc, left phone, right phone) -> instance
  inst_id.reserve(arcs.size() * 4);
  out = FstBuilder();
  start_state = out.add_state();
  if (node_final[0] != kInf) out.finals[start_state] = node_final[0];

  // create instance (lazily), returns id; hmm nodes created on creation
  std::vector<int> work;
  std::string err;
  auto get_inst = [&](int arc, int l, int r) -> int {
    if (!tri) { l = 0; r = 0; }
    const uint64_t key = ((uint64_t)(uint32_t)arc << 40) | ((uint64_t)(uint32_t)l << 20) | (uint64_t)(uint32_t)r;   // phone ids < 2^20
    auto it = inst_id.find(key);
    if (it != inst_id.end()) return it->second;
    int id = (int)insts.size();
    insts.push_back({arc, l, r, {}});
    inst_id[key] = id;
    work.push_back(id);
    return id;
  };
  // per instance: entry arcs description = list of (tid, target graph node) for state-0 forward transitions
  struct Entry { int tid, node; };
  std::vector<std::vector<Entry>> entries;
  std::vector<int> node_tab, tstate;          // scratch of expand(), reused across instances
  std::vector<std::pair<int, int>> order;
  auto expand = [&](int id) {
    // NOTE: insts may reallocate inside; copy fields first
    int arc = insts[id].arc, l = insts[id].l, r = insts[id].r;
    int ph = arcs[arc].phone;
    if (ph < 0 || ph >= (int)C.phone2entry.size() || C.phone2entry[ph] < 0) { err = "phone " + std::to_string(ph) + " has no topology"; return; }
    int ent = C.phone2entry[ph];
    int s0 = C.entry_state_off[ent], ns = C.entry_state_off[ent + 1] - s0;
    int nfinal = ns - 1;
    int ctx[3] = {l, ph, r};
    int ctx1[1] = {ph};
    const uint64_t ckey = tri ? (((uint64_t)(uint32_t)l << 40) | ((uint64_t)(uint32_t)ph << 20) | (uint64_t)(uint32_t)r) : (uint64_t)(uint32_t)ph;
    auto hit = ctx_cache.find(ckey);
    if (hit != ctx_cache.end()) tstate = hit->second;
    else {
      tstate.assign(ns, -1);
      for (int j = 0; j < ns; j++) {
        if (C.fwd_class[s0 + j] < 0) continue;
        int fpdf = C.tree_lookup(tri ? ctx : ctx1, C.fwd_class[s0 + j]);
        int spdf = C.tree_lookup(tri ? ctx : ctx1, C.self_class[s0 + j]);
        if (fpdf < 0 || spdf < 0) { err = "tree has no pdf for context (" + std::to_string(l) + "," + std::to_string(ph) + "," + std::to_string(r) + ")"; return; }
        tstate[j] = C.find_tstate(ph, j, fpdf, spdf);
        if (tstate[j] < 0) { err = "no transition-state for phone " + std::to_string(ph) + " state " + std::to_string(j); return; }
      }
      ctx_cache.emplace(ckey, tstate);
    }
    // nodes keyed by (j' dst, j src): a flat ns x ns table (HMMs have a handful of states)
    node_tab.assign((size_t)ns * ns, -1);
    order.clear();                           // creation order (j', j)
    auto get_node = [&](int jd, int js) {
      int &slot = node_tab[(size_t)jd * ns + js];
      if (slot >= 0) return slot;
      slot = out.add_state();
      order.push_back({jd, js});
      return slot;
    };
    if ((int)entries.size() <= id) entries.resize(id + 1);
    // state-0 forward transitions become entry arcs (attached to predecessors' junctions later)
    {
      int j = 0;
      for (int t = C.trans_off[s0 + j], k = 0; t < C.trans_off[s0 + j + 1]; t++, k++) {
        int jd = C.trans_dst[t];
        if (jd == j) continue;
        entries[id].push_back({C.first_tid[tstate[j]] + k, get_node(jd, j)});
      }
    }
    // closure over internal nodes
    for (size_t oi = 0; oi < order.size(); oi++) {
      int jd = order[oi].first, js = order[oi].second;
      int n = node_tab[(size_t)jd * ns + js];
      if (jd != nfinal) {
        for (int t = C.trans_off[s0 + jd], k = 0; t < C.trans_off[s0 + jd + 1]; t++, k++) {
          int j2 = C.trans_dst[t];
          if (j2 == jd) continue;
          out.add_arc(n, get_node(j2, jd), C.first_tid[tstate[jd]] + k, 0, 0.0f);
        }
      } else {
        insts[id].junctions.push_back(n);
      }
      // self-loop of the SOURCE state js sits here (reorder=true); added last like AddSelfLoops does
      for (int t = C.trans_off[s0 + js], k = 0; t < C.trans_off[s0 + js + 1]; t++, k++)
        if (C.trans_dst[t] == js) out.add_arc(n, n, C.first_tid[tstate[js]] + k, 0, 0.0f);
    }
  };

  // seeds: arcs out of Start
  struct Pending { int from_node; int inst; int olabel; float cost; };  // attach inst's entry arcs at graph node
  std::vector<Pending> pend;
  std::vector<int> rs;
  auto successors = [&](int pnode, int lphone, int from_graph_node, int rfilter) {
    // all instances following phone-graph node `pnode` with left phone `lphone`; rfilter = required phone (tri) or -1
    for (int ea : out_arcs[pnode]) {
      if (tri && rfilter >= 0 && arcs[ea].phone != rfilter) continue;
      int d = arcs[ea].dst;
      if (tri) {
        rs.clear();
        for (int e2 : out_arcs[d]) if (std::find(rs.begin(), rs.end(), arcs[e2].phone) == rs.end()) rs.push_back(arcs[e2].phone);
        if (node_final[d] != kInf) rs.push_back(0);
        for (int rr : rs) pend.push_back({from_graph_node, get_inst(ea, lphone, rr), arcs[ea].olabel, arcs[ea].cost});
      } else {
        pend.push_back({from_graph_node, get_inst(ea, 0, 0), arcs[ea].olabel, arcs[ea].cost});
      }
    }
  };
  successors(0, 0, start_state, -1);
  size_t wi = 0;
  while (wi < work.size()) {
    int id = work[wi++];
    expand(id);
    if (!err.empty()) return err;
    int arc = insts[id].arc, r = insts[id].r;
    int d = arcs[arc].dst, ph = arcs[arc].phone;
    std::vector<int> junc = insts[id].junctions;
    for (int jn : junc) {
      if (tri) {
        if (r == 0) { if (node_final[d] != kInf) out.finals[jn] = node_final[d]; }
        else successors(d, ph, jn, r);
      } else {
        if (node_final[d] != kInf) out.finals[jn] = node_final[d];
        successors(d, 0, jn, -1);
      }
    }
  }
  // attach entry arcs
  for (auto &p : pend)
    for (auto &en : entries[p.inst]) out.add_arc(p.from_node, en.node, en.tid, p.olabel, p.cost);
  return "";
}

// trim states that cannot reach a final state / are unreachable (keeps the decoder's token counts honest)
static void trim(FstBuilder &g, int &start) {
  int S = (int)g.finals.size();
  size_t A = g.src.size();
  std::vector<char> fwd(S, 0), bwd(S, 0);
  // CSR adjacency in both directions (one allocation each: a vector per state made this the most expensive part of graph compilation)
  std::vector<int> out_off(S + 1, 0), in_off(S + 1, 0), out_arc(A), in_arc(A);
  for (size_t a = 0; a < A; a++) { out_off[g.src[a] + 1]++; in_off[g.dst[a] + 1]++; }
  for (int s = 0; s < S; s++) { out_off[s + 1] += out_off[s]; in_off[s + 1] += in_off[s]; }
  {
    std::vector<int> oc(out_off.begin(), out_off.end() - 1), ic(in_off.begin(), in_off.end() - 1);
    for (size_t a = 0; a < A; a++) { out_arc[oc[g.src[a]]++] = (int)a; in_arc[ic[g.dst[a]]++] = (int)a; }
  }
  std::vector<int> st{start}; fwd[start] = 1;
  while (!st.empty()) {
    int s = st.back(); st.pop_back();
    for (int k = out_off[s]; k < out_off[s + 1]; k++) { const int d = g.dst[out_arc[k]]; if (!fwd[d]) { fwd[d] = 1; st.push_back(d); } }
  }
  for (int s = 0; s < S; s++) if (g.finals[s] != kInf) { bwd[s] = 1; st.push_back(s); }
  while (!st.empty()) {
    int s = st.back(); st.pop_back();
    for (int k = in_off[s]; k < in_off[s + 1]; k++) { const int d = g.src[in_arc[k]]; if (!bwd[d]) { bwd[d] = 1; st.push_back(d); } }
  }
  std::vector<int> remap(S, -1); int n = 0;
  for (int s = 0; s < S; s++) if (fwd[s] && bwd[s]) remap[s] = n++;
  if (remap[start] < 0) { g = FstBuilder(); start = -1; return; }
  FstBuilder o; o.finals.resize(n);
  o.src.reserve(A); o.dst.reserve(A); o.il.reserve(A); o.ol.reserve(A); o.w.reserve(A);
  for (int s = 0; s < S; s++) if (remap[s] >= 0) o.finals[remap[s]] = g.finals[s];
  for (size_t a = 0; a < A; a++) if (remap[g.src[a]] >= 0 && remap[g.dst[a]] >= 0) o.add_arc(remap[g.src[a]], remap[g.dst[a]], g.il[a], g.ol[a], g.w[a]);
  start = remap[start];
  g = std::move(o);
}

// ---- band view for viterbi_band.cu ---------------------------------------------------------------
// Renumbers one utterance's states so that every arc either stays inside a strongly connected component (self-loops; the
// ergodic middle states of Kaldi's silence topology) or goes forward: Tarjan SCCs, then a FIFO Kahn order over the component
// DAG (states of a component adjacent, in original order).  A beam-pruned token set then occupies a narrow, forward-moving
// index window (measured on LibriSpeech-shaped synthetic graphs: median 29 states, maximum 170), which is what the band
// kernel processes densely.  Returns false when the graph cannot use the band kernel (epsilon input arcs, in-degree or jump
// beyond the packed 8-bit fields); such utterances run on the sparse kernel.
struct BandOut {
  int start = 0, maxback = 0;
  std::vector<uint32_t> stw, apk;
  std::vector<float> aw, fin;
  std::vector<uint16_t> arcid, orig;
};

static bool build_band(int S, int A, int start, const int32_t *inb, const int32_t *a_src, const int32_t *a_dst, const int32_t *a_lp,
                       const float *a_w, const float *finals, BandOut &o) {
  if (start < 0 || S <= 0 || S > 65534 || A > 65534) return false;
  for (int a = 0; a < A; a++) if (a_lp[a] < 0) return false;
  // Tarjan (iterative) over non-self-loop arcs
  std::vector<int> index(S, -1), low(S, 0), comp(S, -1), stk, call, it(S, 0);
  std::vector<char> on(S, 0);
  int counter = 0, ncomp = 0;
  for (int root = 0; root < S; root++) {
    if (index[root] >= 0) continue;
    call.push_back(root);
    index[root] = low[root] = counter++; stk.push_back(root); on[root] = 1; it[root] = inb[root];
    while (!call.empty()) {
      int v = call.back();
      if (it[v] < inb[v + 1]) {
        int w = a_dst[it[v]++];
        if (w == v) continue;
        if (index[w] < 0) { index[w] = low[w] = counter++; stk.push_back(w); on[w] = 1; it[w] = inb[w]; call.push_back(w); }
        else if (on[w]) low[v] = std::min(low[v], index[w]);
      } else {
        call.pop_back();
        if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
        if (low[v] == index[v]) {
          for (;;) { int w = stk.back(); stk.pop_back(); on[w] = 0; comp[w] = ncomp; if (w == v) break; }
          ncomp++;
        }
      }
    }
  }
  // Kahn over the component DAG (arc multiplicities counted on both sides)
  std::vector<int> indeg(ncomp, 0), cbeg(ncomp + 1, 0), cstates(S);
  for (int s = 0; s < S; s++) cbeg[comp[s] + 1]++;
  for (int c = 0; c < ncomp; c++) cbeg[c + 1] += cbeg[c];
  { std::vector<int> fill(cbeg.begin(), cbeg.end() - 1); for (int s = 0; s < S; s++) cstates[fill[comp[s]]++] = s; }
  for (int a = 0; a < A; a++) if (comp[a_src[a]] != comp[a_dst[a]]) indeg[comp[a_dst[a]]]++;
  std::vector<int> queue; queue.reserve(ncomp);
  if (indeg[comp[start]] == 0) queue.push_back(comp[start]);
  for (int c = 0; c < ncomp; c++) if (indeg[c] == 0 && c != comp[start]) queue.push_back(c);
  std::vector<int> pos(S, -1);
  int np = 0;
  for (size_t qi = 0; qi < queue.size(); qi++) {
    int c = queue[qi];
    for (int k = cbeg[c]; k < cbeg[c + 1]; k++) pos[cstates[k]] = np++;
    for (int k = cbeg[c]; k < cbeg[c + 1]; k++) {
      int s = cstates[k];
      for (int a = inb[s]; a < inb[s + 1]; a++) { int d = comp[a_dst[a]]; if (d != c && --indeg[d] == 0) queue.push_back(d); }
    }
  }
  if (np != S) return false;
  // in-arcs by band destination, by-source arc index ascending inside a destination (ties resolve like the sparse kernel)
  std::vector<int> cnt(S + 1, 0), fwd(S, 0);
  int maxback = 0;
  for (int a = 0; a < A; a++) {
    int ps = pos[a_src[a]], pd = pos[a_dst[a]];
    cnt[pd + 1]++;
    fwd[ps] = std::max(fwd[ps], pd - ps);
    maxback = std::max(maxback, ps - pd);
  }
  // in-degree, forward reach and the back-pointer's source-delta code (dst - src + 16) are packed into bytes
  for (int s = 0; s < S; s++) { if (cnt[s + 1] > 254 || fwd[s] > 239) return false; cnt[s + 1] += cnt[s]; }
  if (maxback > 16) return false;
  o.start = pos[start]; o.maxback = maxback;
  o.stw.assign(S, 0); o.fin.assign(S, 0.0f); o.orig.assign(S, 0);
  o.apk.assign(A, 0); o.aw.assign(A, 0.0f); o.arcid.assign(A, 0);
  for (int s = 0; s < S; s++) {
    int b = pos[s];
    o.stw[b] = (uint32_t)cnt[b] | ((uint32_t)(cnt[b + 1] - cnt[b]) << 16) | ((uint32_t)fwd[b] << 24);
    o.fin[b] = finals[s]; o.orig[b] = (uint16_t)s;
  }
  // fwd was indexed by band position above (fwd[ps]); stw reads fwd[b] accordingly
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  for (int a = 0; a < A; a++) {
    int j = fill[pos[a_dst[a]]]++;
    o.apk[j] = (uint32_t)pos[a_src[a]] | ((uint32_t)a_lp[a] << 16);
    o.aw[j] = a_w[a]; o.arcid[j] = (uint16_t)a;
  }
  return true;
}

}  // namespace mfa

using namespace mfa;

struct mfa_graph_compiler { GraphCompiler c; };
struct mfa_fst_batch { FstBatch b; };

// ---- a11: equal alignment ----------------------------------------------------------------------------------------------
// MonoAlignEqualFunction (acoustic_modeling/monophone.py:63-139) -> kalpy gmm_align_equal -> Kaldi EqualAlign
// (fstext/fstext-utils-inl.h) + GetLinearSymbolSequence: a random start-to-final path without self-loops (uniform choice among
// the out-arcs and, on a final state, "stop"; self-loop draws are re-drawn), retried while it has more input labels than
// frames; the missing frames are spread over the path's self-loops, the first `extra % loops` of them getting one more.
// Kaldi draws with srand(seed) / rand(); glibc's TYPE_3 additive-feedback generator is restated here so the batch can run
// re-entrantly (tests/test_host_logic.py checks it against libc).
namespace {
struct GlibcRand {
  int32_t r[34];
  uint32_t ring[31];
  int pos = 0;
  explicit GlibcRand(uint32_t seed) {
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
      int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
      if (v < 0) v += 2147483647LL;
      r[i] = (int32_t)v;
    }
    uint32_t st[344];
    for (int i = 0; i < 31; i++) st[i] = (uint32_t)r[i];
    for (int i = 31; i < 34; i++) st[i] = st[i - 31];
    for (int i = 34; i < 344; i++) st[i] = st[i - 31] + st[i - 3];
    for (int i = 0; i < 31; i++) ring[i] = st[313 + i];   // the last 31 values; ring[k] = st[313 + k]
    pos = 0;
  }
  int next() {   // value i = value[i-31] + value[i-3]
    const uint32_t v = ring[pos] + ring[(pos + 28) % 31];
    ring[pos] = v;
    pos = (pos + 1) % 31;
    return (int)(v >> 1);
  }
  int rand_int(int lo, int hi) { return lo == hi ? lo : lo + next() % (hi - lo + 1); }   // kaldi::RandInt
};
}  // namespace

// ---- binary OpenFst bodies (FstArchive import): see include/mfa_b200.h
extern "C" int mfa_fst_body_scan(const uint8_t *body, int64_t len, int64_t n_states, int64_t *n_arcs, int64_t *n_bytes) {
  if (!body || !n_arcs || !n_bytes || n_states < 0) return set_error(MFA_ERR_INVALID, "null argument");
  int64_t pos = 0, arcs = 0;
  for (int64_t s = 0; s < n_states; s++) {
    if (pos + 12 > len) return set_error(MFA_ERR_INVALID, "truncated FST: state header beyond the buffer");
    int64_t na;
    memcpy(&na, body + pos + 4, 8);
    if (na < 0 || na > (len - pos - 12) / 16) return set_error(MFA_ERR_INVALID, "truncated FST: arcs beyond the buffer");
    pos += 12 + 16 * na;
    arcs += na;
  }
  *n_arcs = arcs; *n_bytes = pos;
  return MFA_OK;
}
extern "C" int mfa_fst_body_fill(const uint8_t *body, int64_t n_states, float *finals, int32_t *src, int32_t *dst, int32_t *ilabel,
                                 int32_t *olabel, float *weight) {
  if (!body || (n_states > 0 && !finals)) return set_error(MFA_ERR_INVALID, "null argument");
  int64_t pos = 0, a = 0;
  for (int64_t s = 0; s < n_states; s++) {
    int64_t na;
    memcpy(finals + s, body + pos, 4);
    memcpy(&na, body + pos + 4, 8);
    pos += 12;
    for (int64_t k = 0; k < na; k++, a++, pos += 16) {
      src[a] = (int32_t)s;
      memcpy(ilabel + a, body + pos, 4); memcpy(olabel + a, body + pos + 4, 4);
      memcpy(weight + a, body + pos + 8, 4); memcpy(dst + a, body + pos + 12, 4);
    }
  }
  return MFA_OK;
}

extern "C" int mfa_rand_sequence(uint32_t seed, int32_t n, int32_t *out) {
  if (n < 0 || (n && !out)) return set_error(MFA_ERR_INVALID, "bad argument");
  GlibcRand g(seed);
  for (int i = 0; i < n; i++) out[i] = g.next();
  return MFA_OK;
}

extern "C" int mfa_equal_align(const mfa_fst_batch *fb, const int64_t *frame_off, const uint32_t *seeds, int32_t num_retries, int32_t *ali,
                               int32_t *words, const int64_t *word_off, int32_t *num_words, int32_t *status) {
  if (!fb || !frame_off || !seeds || !ali || !status) return set_error(MFA_ERR_INVALID, "null argument");
  const FstBatch &b = fb->b;
  if (num_retries <= 0) num_retries = 10;
  for (int u = 0; u < b.n(); u++) {
    const int64_t T = frame_off[u + 1] - frame_off[u];
    const int ns = (int)(b.state_off[u + 1] - b.state_off[u]);
    const int64_t a0 = b.arc_off[u], a1 = b.arc_off[u + 1];
    if (num_words) num_words[u] = 0;
    if (T <= 0) { status[u] = MFA_ALIGN_ZERO_FRAMES; continue; }
    if (b.start[u] < 0 || ns == 0) { status[u] = MFA_ALIGN_EMPTY_GRAPH; continue; }
    // out-arcs per state, in the batch's (= OpenFst's) order
    std::vector<int> first(ns + 1, 0);
    for (int64_t a = a0; a < a1; a++) first[b.src[a] + 1]++;
    for (int s = 0; s < ns; s++) first[s + 1] += first[s];
    std::vector<int64_t> arcs(a1 - a0);
    { std::vector<int> cur(first.begin(), first.end() - 1); for (int64_t a = a0; a < a1; a++) arcs[cur[b.src[a]]++] = a; }
    const float *fin = &b.finals[b.state_off[u]];
    GlibcRand rng(seeds[u]);
    std::vector<int> path;
    std::vector<int64_t> taken;
    int64_t n_il = 0;
    int retry = 0;
    bool stuck = false;
    do {
      n_il = 0; taken.clear(); path.clear(); path.push_back(b.start[u]);
      int64_t guard = 0;
      for (;;) {
        const int s = path.back();
        const int na = first[s + 1] - first[s];
        const int tot = na + (fin[s] < kInf ? 1 : 0);
        if (tot == 0 || ++guard > (int64_t)50000000) { stuck = true; break; }   // dead end: Kaldi asserts co-accessibility
        const int off = rng.rand_int(0, tot - 1);
        if (off >= na) break;   // chose the final weight
        const int64_t a = arcs[first[s] + off];
        if (b.dst[a] == s) continue;   // self-loops are not taken here
        taken.push_back(a); path.push_back(b.dst[a]);
        if (b.il[a] != 0) n_il++;
      }
    } while (!stuck && ++retry < num_retries && n_il > T);
    if (stuck || n_il > T) { status[u] = MFA_ALIGN_NO_FINAL; continue; }
    std::vector<int64_t> loop(path.size(), -1);
    int64_t n_loops = 0;
    for (size_t i = 0; i < path.size(); i++) {
      const int s = path[i];
      for (int k = first[s]; k < first[s + 1]; k++)
        if (b.dst[arcs[k]] == s && b.il[arcs[k]] != 0) { loop[i] = arcs[k]; n_loops++; break; }
    }
    if (n_loops == 0 && n_il < T) { status[u] = MFA_ALIGN_NO_FINAL; continue; }
    const int64_t extra = T - n_il;
    const int64_t min_loops = extra != 0 ? extra / n_loops : 0;
    const int64_t one_more = extra - min_loops * n_loops;
    int64_t counter = 0, t = 0, nw = 0;
    const int64_t wcap = word_off ? word_off[u + 1] - word_off[u] : 0;
    int32_t *out = ali + frame_off[u];
    bool overflow = false;
    auto emit = [&](int64_t a) {
      if (b.il[a] != 0) { if (t < T) out[t] = b.il[a]; t++; }
      if (b.ol[a] != 0) { if (words && nw < wcap) words[word_off[u] + nw] = b.ol[a]; else if (words) overflow = true; nw++; }
    };
    for (size_t i = 0; i < path.size(); i++) {
      if (loop[i] >= 0) {
        const int64_t k = min_loops + (counter < one_more ? 1 : 0);
        counter++;
        for (int64_t j = 0; j < k; j++) emit(loop[i]);
      }
      if (i + 1 < path.size()) emit(taken[i]);
    }
    if (t != T) return set_error(MFA_ERR_GRAPH, "equal alignment length mismatch");
    if (overflow) return set_error(MFA_ERR_INVALID, "word buffer too small for the equal-alignment path");
    if (num_words) num_words[u] = (int32_t)nw;
    status[u] = MFA_ALIGN_OK;
  }
  return MFA_OK;
}

extern "C" {

int mfa_graph_compiler_create(const mfa_hmm_desc *h, const mfa_lexicon_desc *l, mfa_graph_compiler **out) {
  if (!h || !l || !out) return set_error(MFA_ERR_INVALID, "null argument");
  if (!((h->ctx_width == 1 && h->central_pos == 0) || (h->ctx_width == 3 && h->central_pos == 1)))
    return set_error(MFA_ERR_UNSUPPORTED, "only (N,P)=(1,0) and (3,1) context trees are supported");
  auto *gc = new mfa_graph_compiler();
  GraphCompiler &c = gc->c;
  c.phone2entry.assign(h->phone2entry, h->phone2entry + h->num_phones);
  c.entry_state_off.assign(h->entry_state_off, h->entry_state_off + h->num_entries + 1);
  int nhs = c.entry_state_off.back();
  c.fwd_class.assign(h->state_fwd_class, h->state_fwd_class + nhs);
  c.self_class.assign(h->state_self_class, h->state_self_class + nhs);
  c.trans_off.assign(h->state_trans_off, h->state_trans_off + nhs + 1);
  c.trans_dst.assign(h->trans_dst, h->trans_dst + c.trans_off.back());
  c.tuples.assign(h->tuples, h->tuples + 4 * (size_t)h->num_tstates);
  c.first_tid.assign(h->tstate_first_tid, h->tstate_first_tid + h->num_tstates + 2);
  for (int ts = 1; ts <= h->num_tstates; ts++) {
    const int32_t *t = &c.tuples[4 * (ts - 1)];
    uint64_t k1 = ((uint64_t)(uint32_t)t[0] << 32) | (uint32_t)t[1];
    uint64_t k2 = ((uint64_t)(uint32_t)t[2] << 32) | (uint32_t)t[3];
    c.tstate_map[k1].push_back({k2, ts});
  }
  c.ctx_width = h->ctx_width; c.central = h->central_pos; c.tree_root = h->tree_root;
  c.tree_nodes.assign(h->tree_nodes, h->tree_nodes + 4 * (size_t)h->num_tree_nodes);
  c.tree_aux_off.assign(h->tree_aux_off, h->tree_aux_off + h->num_tree_nodes + 1);
  c.tree_aux.assign(h->tree_aux, h->tree_aux + c.tree_aux_off.back());
  c.word_pron_off.assign(l->word_pron_off, l->word_pron_off + l->num_words + 1);
  int np = c.word_pron_off.back();
  c.pron_phone_off.assign(l->pron_phone_off, l->pron_phone_off + np + 1);
  c.pron_phones.assign(l->pron_phones, l->pron_phones + c.pron_phone_off.back());
  c.pron_cost.assign(l->pron_cost, l->pron_cost + np);
  auto fill = [&](std::vector<float> &v, const float *p, float dflt) { if (p) v.assign(p, p + np); else v.assign(np, dflt); };
  fill(c.sil_after, l->pron_sil_after_cost, l->sil_cost);
  fill(c.nonsil_after, l->pron_nonsil_after_cost, l->nonsil_cost);
  fill(c.sil_before, l->pron_sil_before_cost, 0.0f);
  fill(c.nonsil_before, l->pron_nonsil_before_cost, 0.0f);
  c.sil_phone = l->sil_phone;
  c.init_sil = l->init_sil_cost; c.init_nonsil = l->init_nonsil_cost;
  c.final_sil = l->final_sil_cost; c.final_nonsil = l->final_nonsil_cost;
  *out = gc;
  return MFA_OK;
}

int mfa_graph_compiler_destroy(mfa_graph_compiler *c) { delete c; return MFA_OK; }

int mfa_graph_compile(mfa_graph_compiler *c, const int32_t *words, const int64_t *word_off, int32_t n_utts, int32_t n_threads,
                      mfa_fst_batch **out) {
  if (!c || !word_off || !out || n_utts < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  std::vector<FstBuilder> gs(n_utts);
  std::vector<int> starts(n_utts, -1);
  std::vector<std::string> errs(n_utts);
  if (n_threads < 1) n_threads = 1;
  n_threads = std::min<int>(n_threads, std::max(1, n_utts));
  auto worker = [&](int tid) {
    CtxCache ctx_cache;
    for (int u = tid; u < n_utts; u += n_threads) {
      errs[u] = compile_one(c->c, words + word_off[u], word_off[u + 1] - word_off[u], gs[u], starts[u], ctx_cache);
      if (errs[u].empty()) trim(gs[u], starts[u]);
    }
  };
  if (n_threads == 1) worker(0);
  else { std::vector<std::thread> th; for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t); for (auto &t : th) t.join(); }
  for (int u = 0; u < n_utts; u++) if (!errs[u].empty()) return set_error(MFA_ERR_GRAPH, "utterance " + std::to_string(u) + ": " + errs[u]);
  auto *fb = new mfa_fst_batch();
  FstBatch &b = fb->b;
  // offsets first, then every thread copies its utterances into place (the serial append of ~180 MB was a third of the call)
  b.state_off.assign((size_t)n_utts + 1, 0); b.arc_off.assign((size_t)n_utts + 1, 0);
  b.start.assign(starts.begin(), starts.end());
  for (int u = 0; u < n_utts; u++) {
    b.state_off[u + 1] = b.state_off[u] + (int64_t)gs[u].finals.size();
    b.arc_off[u + 1] = b.arc_off[u] + (int64_t)gs[u].src.size();
  }
  const size_t TS = (size_t)b.state_off[n_utts], TA = (size_t)b.arc_off[n_utts];
  b.finals.resize(TS); b.src.resize(TA); b.dst.resize(TA); b.il.resize(TA); b.ol.resize(TA); b.w.resize(TA);
  auto gather = [&](int tid) {
    for (int u = tid; u < n_utts; u += n_threads) {
      const size_t so = (size_t)b.state_off[u], ao = (size_t)b.arc_off[u];
      std::copy(gs[u].finals.begin(), gs[u].finals.end(), b.finals.begin() + so);
      std::copy(gs[u].src.begin(), gs[u].src.end(), b.src.begin() + ao); std::copy(gs[u].dst.begin(), gs[u].dst.end(), b.dst.begin() + ao);
      std::copy(gs[u].il.begin(), gs[u].il.end(), b.il.begin() + ao); std::copy(gs[u].ol.begin(), gs[u].ol.end(), b.ol.begin() + ao);
      std::copy(gs[u].w.begin(), gs[u].w.end(), b.w.begin() + ao);
      gs[u] = FstBuilder();
    }
  };
  if (n_threads == 1) gather(0);
  else { std::vector<std::thread> th; for (int t = 0; t < n_threads; t++) th.emplace_back(gather, t); for (auto &t : th) t.join(); }
  *out = fb;
  return MFA_OK;
}

int mfa_fst_batch_create(int32_t n_utts, const int64_t *state_off, const int64_t *arc_off, const int32_t *start, const float *finals,
                         const int32_t *src, const int32_t *dst, const int32_t *ilabel, const int32_t *olabel, const float *weight,
                         mfa_fst_batch **out) {
  if (n_utts < 0 || !state_off || !arc_off || !out) return set_error(MFA_ERR_INVALID, "bad argument");
  auto *fb = new mfa_fst_batch();
  FstBatch &b = fb->b;
  int64_t S = state_off[n_utts], A = arc_off[n_utts];
  b.state_off.assign(state_off, state_off + n_utts + 1);
  b.arc_off.assign(arc_off, arc_off + n_utts + 1);
  b.start.assign(start, start + n_utts);
  b.finals.assign(finals, finals + S);
  b.src.assign(src, src + A); b.dst.assign(dst, dst + A); b.il.assign(ilabel, ilabel + A); b.ol.assign(olabel, olabel + A);
  b.w.assign(weight, weight + A);
  for (int u = 0; u < n_utts; u++) {
    int64_t ns = state_off[u + 1] - state_off[u];
    for (int64_t a = arc_off[u]; a < arc_off[u + 1]; a++)
      if (src[a] < 0 || src[a] >= ns || dst[a] < 0 || dst[a] >= ns) { delete fb; return set_error(MFA_ERR_INVALID, "arc endpoint out of range"); }
    if (start[u] >= ns) { delete fb; return set_error(MFA_ERR_INVALID, "start state out of range"); }
  }
  *out = fb;
  return MFA_OK;
}

int mfa_fst_batch_destroy(mfa_fst_batch *b) { delete b; return MFA_OK; }

int mfa_fst_batch_sizes(const mfa_fst_batch *b, int32_t *n_utts, int64_t *n_states, int64_t *n_arcs) {
  if (!b) return set_error(MFA_ERR_INVALID, "null batch");
  if (n_utts) *n_utts = b->b.n();
  if (n_states) *n_states = b->b.state_off.back();
  if (n_arcs) *n_arcs = b->b.arc_off.back();
  return MFA_OK;
}

int mfa_fst_batch_export(const mfa_fst_batch *fb, int64_t *state_off, int64_t *arc_off, int32_t *start, float *finals, int32_t *src,
                         int32_t *dst, int32_t *ilabel, int32_t *olabel, float *weight) {
  if (!fb) return set_error(MFA_ERR_INVALID, "null batch");
  const FstBatch &b = fb->b;
  auto cp = [](auto *d, const auto &v) { if (d && !v.empty()) std::memcpy(d, v.data(), v.size() * sizeof(v[0])); };
  cp(state_off, b.state_off); cp(arc_off, b.arc_off); cp(start, b.start); cp(finals, b.finals);
  cp(src, b.src); cp(dst, b.dst); cp(ilabel, b.il); cp(olabel, b.ol); cp(weight, b.w);
  return MFA_OK;
}

// ---- packing for the Viterbi kernel -----------------------------------------------------------
int mfa_graphs_pack(const mfa_fst_batch *fb, const float *tid_cost, const int32_t *tid2pdf, int32_t num_tids, mfa_graphs **out) {
  if (!fb || !tid_cost || !tid2pdf || !out) return set_error(MFA_ERR_INVALID, "null argument");
  const FstBatch &b = fb->b;
  const int n = b.n();
  // Phase 1, utterances in parallel on host threads (the band renumbering -- Tarjan + Kahn per graph -- is most of the time: 0.26 ms per
  // utterance, 0.78 s for the 10 h workload when serial): everything of one utterance goes into its own PackedUtt.
  struct PackedUtt {
    std::vector<int32_t> inb, a_src, a_dst, a_lp, a_tid, a_olabel, pdfs;
    std::vector<float> a_w, a_w0;
    int32_t n_eps = 0, words = 0;
    bool ok = false;
    BandOut bo;
    int err = 0;
  };
  static const bool trace = getenv("MFA_PACK_TRACE") != nullptr;   // read once: phase times on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  std::vector<PackedUtt> pu(n);
  // graphs beyond the 16-bit indices of the packed views are packed as EMPTY graphs and flagged: that utterance fails with its own
  // status (MFA_ALIGN_GRAPH_TOO_LARGE), the rest of the batch is aligned
  std::vector<uint8_t> big(n, 0);
  for (int u = 0; u < n; u++) {
    const int64_t S = b.state_off[u + 1] - b.state_off[u], A = b.arc_off[u + 1] - b.arc_off[u];
    big[u] = S > 65534 || A > 65534;
  }
  auto work = [&](int u, std::vector<int32_t> &lpmap) {
    PackedUtt &P = pu[u];
    const int64_t s0 = b.state_off[u], a0 = b.arc_off[u];
    const int64_t S = big[u] ? 0 : b.state_off[u + 1] - s0, A = big[u] ? 0 : b.arc_off[u + 1] - a0;
    std::vector<int32_t> order(A);   // arcs by (src, original index)
    for (int64_t a = 0; a < A; a++) order[a] = (int32_t)a;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return b.src[a0 + x] < b.src[a0 + y]; });
    P.inb.assign(S + 1, 0);
    for (int64_t a = 0; a < A; a++) P.inb[b.src[a0 + a] + 1]++;
    for (int64_t s = 0; s < S; s++) P.inb[s + 1] += P.inb[s];
    for (int64_t a = 0; a < A; a++) { const int il = b.il[a0 + a]; if (il > 0) { if (il > num_tids) { P.err = 1; return; } P.pdfs.push_back(tid2pdf[il]); } }
    std::sort(P.pdfs.begin(), P.pdfs.end()); P.pdfs.erase(std::unique(P.pdfs.begin(), P.pdfs.end()), P.pdfs.end());
    const int maxpdf = P.pdfs.empty() ? 0 : P.pdfs.back();
    if ((int)lpmap.size() <= maxpdf) lpmap.resize(maxpdf + 1);
    for (size_t k = 0; k < P.pdfs.size(); k++) lpmap[P.pdfs[k]] = (int32_t)k;
    P.a_src.resize(A); P.a_dst.resize(A); P.a_lp.resize(A); P.a_tid.resize(A); P.a_olabel.resize(A); P.a_w.resize(A); P.a_w0.resize(A);
    for (int64_t k = 0; k < A; k++) {
      const int64_t a = a0 + order[k];
      const int il = b.il[a];
      P.a_src[k] = b.src[a]; P.a_dst[k] = b.dst[a]; P.a_tid[k] = il; P.a_olabel[k] = b.ol[a];
      if (b.ol[a] != 0) P.words++;   // loose bound (a path crosses each labelled arc at most once in an acyclic word graph)
      P.a_w0[k] = b.w[a];
      if (il > 0) { P.a_w[k] = b.w[a] + tid_cost[il]; P.a_lp[k] = lpmap[tid2pdf[il]]; }
      else { P.a_w[k] = b.w[a]; P.a_lp[k] = -1; P.n_eps++; }
    }
    P.ok = build_band((int)S, (int)A, big[u] ? -1 : b.start[u], P.inb.data(), P.a_src.data(), P.a_dst.data(), P.a_lp.data(), P.a_w.data(), b.finals.data() + s0, P.bo);
    if (!P.ok) { P.bo.stw.assign(S, 0); P.bo.fin.assign(S, 0.0f); P.bo.orig.assign(S, 0); P.bo.apk.assign(A, 0); P.bo.aw.assign(A, 0.0f); P.bo.arcid.assign(A, 0); }
  };
  {
    static const int env_nt = [] { const char *ev = getenv("MFA_PACK_THREADS"); return ev ? atoi(ev) : 0; }();   // read once per process
    int nt = env_nt > 0 ? env_nt : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(32, n)));
    std::atomic<int> next{0};
    auto loop = [&]() { std::vector<int32_t> lpmap; for (int u = next.fetch_add(1); u < n; u = next.fetch_add(1)) work(u, lpmap); };
    if (nt <= 1) loop();
    else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(loop); for (auto &t : th) t.join(); }
  }
  for (int u = 0; u < n; u++) if (pu[u].err) return set_error(MFA_ERR_INVALID, "ilabel exceeds num_tids");
  const double ms_phase1 = since(t_begin);
  const auto t_alloc = std::chrono::steady_clock::now();
  // Phase 2: concatenate
  auto *g = new mfa_graphs();
  g->n_utts = n;
  g->num_tids = num_tids;
  g->st_off.assign(n + 1, 0); g->arc_off.assign(n + 1, 0); g->lp_off.assign(n + 1, 0); g->inb_off.assign(n + 1, 0);
  g->start.resize(n); g->n_eps.assign(n, 0); g->max_words.assign(n, 0);
  g->too_large = big;
  for (int u = 0; u < n; u++) g->n_too_large += big[u];
  for (int u = 0; u < n; u++) {
    const int64_t S = big[u] ? 0 : b.state_off[u + 1] - b.state_off[u], A = big[u] ? 0 : b.arc_off[u + 1] - b.arc_off[u];
    g->st_off[u + 1] = g->st_off[u] + S; g->arc_off[u + 1] = g->arc_off[u] + A;
    g->lp_off[u + 1] = g->lp_off[u] + (int64_t)pu[u].pdfs.size(); g->inb_off[u + 1] = g->inb_off[u] + S + 1;
  }
  const size_t TS = (size_t)g->st_off[n], TA = (size_t)g->arc_off[n];
  g->in_begin.resize((size_t)g->inb_off[n]); g->lp2pdf.resize((size_t)g->lp_off[n]);
  g->a_src.resize(TA); g->a_dst.resize(TA); g->a_lp.resize(TA); g->a_tid.resize(TA); g->a_olabel.resize(TA); g->a_w.resize(TA); g->a_w0.resize(TA);
  g->final_w.resize(TS);
  for (int u = 0; u < n; u++)
    if (!big[u]) std::copy(b.finals.begin() + b.state_off[u], b.finals.begin() + b.state_off[u + 1], g->final_w.begin() + g->st_off[u]);
  g->band_ok.resize(n); g->b_start.resize(n); g->b_maxback.resize(n);
  g->b_stw.resize(TS); g->b_fin.resize(TS); g->b_orig.resize(TS); g->b_apk.resize(TA); g->b_aw.resize(TA); g->b_arcid.resize(TA);
  g->h_barc.resize(2 * TA); g->h_apack.resize(TA);
  const double ms_alloc = since(t_alloc);
  const auto t_place = std::chrono::steady_clock::now();
  auto place = [&](int u) {
    PackedUtt &P = pu[u];
    const size_t so = (size_t)g->st_off[u], ao = (size_t)g->arc_off[u];
    g->start[u] = big[u] ? -1 : b.start[u]; g->n_eps[u] = P.n_eps; g->max_words[u] = P.words;
    std::copy(P.inb.begin(), P.inb.end(), g->in_begin.begin() + g->inb_off[u]);
    std::copy(P.pdfs.begin(), P.pdfs.end(), g->lp2pdf.begin() + g->lp_off[u]);
    std::copy(P.a_src.begin(), P.a_src.end(), g->a_src.begin() + ao); std::copy(P.a_dst.begin(), P.a_dst.end(), g->a_dst.begin() + ao);
    std::copy(P.a_lp.begin(), P.a_lp.end(), g->a_lp.begin() + ao); std::copy(P.a_tid.begin(), P.a_tid.end(), g->a_tid.begin() + ao);
    std::copy(P.a_olabel.begin(), P.a_olabel.end(), g->a_olabel.begin() + ao); std::copy(P.a_w.begin(), P.a_w.end(), g->a_w.begin() + ao);
    std::copy(P.a_w0.begin(), P.a_w0.end(), g->a_w0.begin() + ao);
    g->band_ok[u] = P.ok ? 1 : 0; g->b_start[u] = P.ok ? P.bo.start : -1; g->b_maxback[u] = P.ok ? P.bo.maxback : 0;
    std::copy(P.bo.stw.begin(), P.bo.stw.end(), g->b_stw.begin() + so); std::copy(P.bo.fin.begin(), P.bo.fin.end(), g->b_fin.begin() + so);
    std::copy(P.bo.orig.begin(), P.bo.orig.end(), g->b_orig.begin() + so); std::copy(P.bo.apk.begin(), P.bo.apk.end(), g->b_apk.begin() + ao);
    std::copy(P.bo.aw.begin(), P.bo.aw.end(), g->b_aw.begin() + ao); std::copy(P.bo.arcid.begin(), P.bo.arcid.end(), g->b_arcid.begin() + ao);
    for (size_t k = 0; k < P.a_dst.size(); k++) {
      g->h_barc[2 * (ao + k)] = P.bo.apk[k]; std::memcpy(&g->h_barc[2 * (ao + k) + 1], &P.bo.aw[k], 4);
      g->h_apack[ao + k] = (uint32_t)(P.a_dst[k] & 0xFFFF) | ((uint32_t)(P.a_lp[k] < 0 ? 0xFFFF : P.a_lp[k]) << 16);
    }
    P = PackedUtt();   // release as we go
  };
  {
    // (disjoint destination ranges per utterance: the placement runs on the same threads as phase 1)
    static const int env_nt2 = [] { const char *ev = getenv("MFA_PACK_THREADS"); return ev ? atoi(ev) : 0; }();
    int nt = env_nt2 > 0 ? env_nt2 : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(32, n)));
    std::atomic<int> next{0};
    auto loop = [&]() { for (int u = next.fetch_add(1); u < n; u = next.fetch_add(1)) place(u); };
    if (nt <= 1) loop();
    else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(loop); for (auto &t : th) t.join(); }
  }
  if (trace) fprintf(stderr, "[mfa_graphs_pack] %d utterances: per-utterance packing %.1f ms, allocation %.1f ms, placement %.1f ms\n", n, ms_phase1, ms_alloc, since(t_place));
  *out = g;
  return MFA_OK;
}

int mfa_graphs_offsets(const mfa_graphs *g, int64_t *state_off, int64_t *arc_off, int64_t *pdf_off) {
  if (!g) return set_error(MFA_ERR_INVALID, "null argument");
  const size_t n = (size_t)g->n_utts + 1;
  if (state_off) std::memcpy(state_off, g->st_off.data(), n * sizeof(int64_t));
  if (arc_off) std::memcpy(arc_off, g->arc_off.data(), n * sizeof(int64_t));
  if (pdf_off) std::memcpy(pdf_off, g->lp_off.data(), n * sizeof(int64_t));
  return MFA_OK;
}

int mfa_graphs_band_view(const mfa_graphs *g, int32_t *band_ok, int32_t *start, int32_t *maxback, uint32_t *state_word, uint16_t *orig_state,
                         uint32_t *arc_word, uint16_t *arc_index) {
  if (!g) return set_error(MFA_ERR_INVALID, "null argument");
  auto cp = [](auto *d, const auto &v) { if (d && !v.empty()) std::memcpy(d, v.data(), v.size() * sizeof(v[0])); };
  cp(band_ok, g->band_ok); cp(start, g->b_start); cp(maxback, g->b_maxback); cp(state_word, g->b_stw); cp(orig_state, g->b_orig);
  cp(arc_word, g->b_apk); cp(arc_index, g->b_arcid);
  return MFA_OK;
}

int mfa_graphs_max_words(const mfa_graphs *g, int32_t *max_words) {
  if (!g || !max_words) return set_error(MFA_ERR_INVALID, "null argument");
  std::memcpy(max_words, g->max_words.data(), sizeof(int32_t) * g->max_words.size());
  return MFA_OK;
}

}  // extern "C"
