"""CPU restatement (pure Python) of training-graph compilation for a linear transcript.  TEST INFRASTRUCTURE ONLY: imported by tests/
(as an independent check of csrc/graph.cc) and by bench.py's reference arm (so that the CPU arm never loads libmfa_b200.so).

Restates what kalpy ``TrainingGraphCompiler.compile_fst`` produces (reference call sites: montreal_forced_aligner/alignment/
multiprocessing.py:537-571, online/alignment.py:77-96) -- Kaldi decoder/training-graph-compiler.cc: L o G -> context expansion (tree
look-ups) -> H -> self-loops with reorder=true, transition probabilities NOT included (they are added at alignment time,
hmm/hmm-utils.cc AddTransitionProbs) -- up to the weighted set of (transition-id sequence, word sequence) paths: state numbering
and the position of lexicon weights along a path are not Kaldi's (no determinisation / minimisation).
Lexicon layout: tests/data/dictionaries/expected/lexicon.text.fst of the reference tree (start -> optional silence; per word the
pronunciation chain, ending either in the loop state or in the pre-silence state).
Parity unpinned against real Kaldi; pinned against the reference's own lexicon FST fixture through
tests/test_reference_dictionary_fixtures.py and against csrc/graph.cc through tests/test_graph_oracle.py.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def compile_fst(tm, tree, lexicon, words: List[int]):
    """-> mfa_b200.kaldi_io.Fst.  tm: TransitionModel, tree: ContextDependency, lexicon: mfa_b200.lexicon.Lexicon, words: word ids."""
    from mfa_b200.kaldi_io import Fst
    A = lexicon._arrs   # word_pron_off, pron_phone_off, pron_phones, pron_cost, sil_after, nonsil_after, sil_before_corr, nonsil_before_corr
    d, _keep = lexicon.desc()
    sil = int(d.sil_phone)
    n = len(words)
    tri = tree.N == 3
    # ---- phone lattice.  Nodes: ("NS", i) boundary i reached without silence, ("P", i) before the optional silence, ("AS", i) after it.
    # A phone arc = (src node, dst node, phone, olabel, cost); start alternatives carry their own cost.
    arcs: List[Tuple[tuple, tuple, int, int, float]] = []
    starts = [(("NS", 0), float(d.init_nonsil_cost))]          # no initial silence
    arcs.append((("START",), ("AS", 0), sil, 0, float(d.init_sil_cost)))
    for i, w in enumerate(words):
        for pr in range(int(A[0][w]), int(A[0][w + 1])):
            ph = [int(x) for x in A[2][int(A[1][pr]):int(A[1][pr + 1])]]
            for src, before in ((("NS", i), float(A[7][pr])), (("AS", i), float(A[6][pr]))):
                base = float(A[3][pr]) + before
                for k, p in enumerate(ph):
                    a = src if k == 0 else ("C", i, pr, src[0], k)
                    ol = w if k == 0 else 0
                    c = base if k == 0 else 0.0
                    if k == len(ph) - 1:
                        arcs.append((a, ("NS", i + 1), p, ol, c + float(A[5][pr])))
                        arcs.append((a, ("P", i + 1), p, ol, c + float(A[4][pr])))
                    else:
                        arcs.append((a, ("C", i, pr, src[0], k + 1), p, ol, c))
        arcs.append((("P", i + 1), ("AS", i + 1), sil, 0, 0.0))
    finals = {("NS", n): float(d.final_nonsil_cost), ("AS", n): float(d.final_sil_cost)}
    out_of: Dict[tuple, List[int]] = {}
    for ai, a in enumerate(arcs):
        out_of.setdefault(a[0], []).append(ai)
    # ---- instances (arc, left phone, right phone): right = phone of the successor arc actually taken (0 at the end of the utterance)
    def rights(ai):
        dst = arcs[ai][1]
        r = {arcs[b][2] for b in out_of.get(dst, [])}
        if dst in finals:
            r.add(0)
        return sorted(r) if tri else [0]
    topo = tm.topo
    st_id: Dict[tuple, int] = {}
    src_l, dst_l, il_l, ol_l, w_l = [], [], [], [], []
    loops: Dict[int, int] = {}

    def state(key):
        if key not in st_id:
            st_id[key] = len(st_id)
        return st_id[key]

    def add_arc(s, t, il, ol, w):
        src_l.append(s); dst_l.append(t); il_l.append(il); ol_l.append(ol); w_l.append(w)

    def tids_of(phone, l, r, j):
        """(tstate, transitions) of HMM state j of `phone` in context (l, r)."""
        st = topo.states_for(phone)[j]
        ctx = [l, phone, r] if tri else [phone]
        fp = tree.lookup(ctx, st.forward_pdf_class)
        sp = tree.lookup(ctx, st.self_loop_pdf_class)
        ts = tm.tuple_to_tstate(phone, j, fp, sp)
        return ts, st.transitions

    start_state = state(("S",))
    final_w: Dict[int, float] = {}
    done = set()

    def expand(ai, l, r, entry_state, entry_cost):
        """HMM of instance (ai, l, r) entered from graph state `entry_state` (a junction) with lexicon cost `entry_cost` on its entry arcs."""
        _, dst, phone, ol, _c = arcs[ai]
        states = topo.states_for(phone)
        J = len(states) - 1
        # entry: forward transitions of HMM state 0
        ts0, tr0 = tids_of(phone, l, r, 0)
        for k, (j2, _p) in enumerate(tr0):
            if j2 == 0:
                continue
            add_arc(entry_state, state((ai, l, r, j2, 0)), int(tm.state2id[ts0]) + k, ol, entry_cost)
        key = (ai, l, r)
        if key in done:
            return
        done.add(key)
        # inside: node (j', j) = "entered j' from j": carries the self-loop of j (reorder = true), leaves by the forward transitions of j'
        seen, todo = set(), [(j2, 0) for j2, _p in tr0 if j2 != 0]
        while todo:
            j2, j = todo.pop()
            if (j2, j) in seen:
                continue
            seen.add((j2, j))
            node = state((ai, l, r, j2, j))
            tsj, trj = tids_of(phone, l, r, j)
            for k, (jj, _p) in enumerate(trj):
                if jj == j:
                    loops[node] = int(tm.state2id[tsj]) + k
            if j2 == J:
                # junction: the instance is over; continue with every successor arc whose phone is r, or end the utterance
                if dst in finals and r == 0:
                    final_w[node] = finals[dst]
                for b in out_of.get(dst, []):
                    if tri and arcs[b][2] != r:
                        continue
                    for r2 in rights(b):
                        expand(b, phone if tri else 0, r2, node, arcs[b][4])
                continue
            ts2, tr2 = tids_of(phone, l, r, j2)
            for k, (j3, _p) in enumerate(tr2):
                if j3 == j2:
                    continue
                add_arc(node, state((ai, l, r, j3, j2)), int(tm.state2id[ts2]) + k, 0, 0.0)
                todo.append((j3, j2))

    import sys
    sys.setrecursionlimit(max(10000, 200 * (n + 2)))
    for node, cost in starts:
        for b in out_of.get(node, []):
            for r2 in rights(b):
                expand(b, 0, r2, start_state, cost + arcs[b][4])
        if node in finals:      # empty transcript without silence
            final_w[start_state] = cost + finals[node]
    for b in out_of.get(("START",), []):
        for r2 in rights(b):
            expand(b, 0, r2, start_state, arcs[b][4])
    S = len(st_id)
    for node, tid in loops.items():
        add_arc(node, node, tid, 0, 0.0)
    fin = np.full(S, np.inf, np.float32)
    for s, w in final_w.items():
        fin[s] = w
    return Fst(start_state, S, np.asarray(src_l, np.int32), np.asarray(il_l, np.int32), np.asarray(ol_l, np.int32), np.asarray(dst_l, np.int32),
               np.asarray(w_l, np.float32), fin)
