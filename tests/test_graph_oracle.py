"""oracle/graph_oracle.py (pure-Python restatement of training-graph compilation) against the host C++ compiler csrc/graph.cc: the same
weighted path set on small lexicons (monophone / triphone / position-dependent), and -- on a synthetic corpus -- identical oracle
alignments and log-likelihoods on both compilers' graphs.  The Python restatement is what bench.py's reference arm aligns on."""
import numpy as np
import pytest

from helpers import build_synth_scenario
from mfa_b200 import engine as E
from oracle import graph_oracle as GO, oracle as O
from test_graph_compiler import _paths, _setup


def _path_set(fst, tm):
    out = {}
    for arcs, cost in _paths(fst):
        key = (tuple(int(fst.arc_ilabel[a]) for a in arcs), tuple(int(fst.arc_olabel[a]) for a in arcs if fst.arc_olabel[a] != 0))
        out[key] = min(out.get(key, np.inf), cost)
    return out


@pytest.mark.parametrize("triphone,posdep", [(False, False), (True, False), (True, True)])
def test_python_restatement_equals_cpp_compiler_path_sets(triphone, posdep):
    lex, pt, topo, tree, tm = _setup(triphone, position_dependent=posdep)
    for words in ([lex.word_table["x"]], [lex.word_table[w] for w in ("x", "z", "y")], []):
        a = E.GraphCompiler(tm, tree, lex).compile([words]).export()[0]
        b = GO.compile_fst(tm, tree, lex, words)
        pa, pb = _path_set(a, tm), _path_set(b, tm)
        assert set(pa) == set(pb) and len(pa) > 0
        for k in pa:
            assert abs(pa[k] - pb[k]) < 1e-4
        # self-loops: same transition-id on the states a given forward arc enters (reorder = true)
        la = {int(a.arc_ilabel[i]) for i in range(len(a.arc_src)) if a.arc_src[i] == a.arc_dst[i]}
        lb = {int(b.arc_ilabel[i]) for i in range(len(b.arc_src)) if b.arc_src[i] == b.arc_dst[i]}
        assert la == lb


def test_oracle_alignments_agree_on_both_compilers_graphs():
    sc = build_synth_scenario(seconds=40.0, seed=21, triphone=True, n_phones=10, n_words=40, target_pdfs=90, gauss_per_pdf=2)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    fsts = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts).export()
    g = O.GmmModel.from_am(am)
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    n = 0
    for u in range(min(6, c.n_utts)):
        f = sc["feats"][u]
        ra = O.align(fsts[u], tc, g, tm.tid2pdf, f, f.shape[0], 0.1, 10.0, 40.0)
        rb = O.align(GO.compile_fst(tm, sc["tree"], c.lexicon, c.transcripts[u]), tc, g, tm.tid2pdf, f, f.shape[0], 0.1, 10.0, 40.0)
        assert ra["status"] == rb["status"]
        if ra["status"] < 2:
            assert np.array_equal(ra["ali"], rb["ali"]) and list(ra["words"]) == list(rb["words"])
            assert abs(ra["like"] - rb["like"]) <= 1e-5 * abs(ra["like"])
            n += 1
    assert n >= 4
