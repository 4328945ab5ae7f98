"""Pins the lexicon / topology handling and the host graph compiler against the fixtures the REFERENCE holds for this path
(/root/reference/tests/data/dictionaries/expected/{lexicon.text.fst, topo, phones.txt, words.txt}; committed as
tests/golden/dictionary_fixtures.npz by tests/golden/make_dictionary_golden.py):

  * phones.txt / words.txt: the symbol numbering of lexicon.make_phone_table / Lexicon.word_table;
  * topo: the Kaldi text topology parser and the MFA-shaped topology synth.make_topology builds;
  * lexicon.text.fst: the language (phone sequences with position tags, word sequences, weights) of the reference's lexicon FST,
    enumerated from the fixture itself, equals the phone-level path set of the training graphs csrc/graph.cc compiles for the same
    dictionary -- the graph compiler never materialises L, so this is the check that its closed-form composition means the same FST.
"""
import math
import os

import numpy as np
import pytest

from mfa_b200 import engine as E, kaldi_io as K, lexicon as LX, synth as SY
from test_graph_compiler import _paths

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dictionary_fixtures.npz")


@pytest.fixture(scope="module")
def fx():
    g = np.load(GOLD)
    txt = {k: bytes(g[k]).decode() for k in g.files}
    sym = lambda t: {a: int(b) for a, b in (ln.split() for ln in t.strip().splitlines())}
    return dict(phones=sym(txt["phones_txt"]), words=sym(txt["words_txt"]), topo=txt["topo"], lexfst=txt["lexicon_text_fst"], dict=txt["abstract_dict"])


def _lexicon(fx):
    """The dictionary behind the fixture: test_abstract.txt's worda / wordb, plus MFA's silence word and OOV word."""
    prons = {}
    for ln in fx["dict"].strip().splitlines():
        w, *ph = ln.split()
        if w in fx["words"]:
            prons[w] = [LX.Pron(ph)]
    prons["!SIL"] = [LX.Pron(["sil"])]
    prons["<unk>"] = [LX.Pron(["spn"])]
    pt = LX.make_phone_table(["phonea", "phoneb", "phonec"], ("sil", "spn"), True)
    return LX.Lexicon(prons, pt, silence_probability=0.5, initial_silence_probability=0.5, position_dependent_phones=True), pt


def test_symbol_tables_match_reference_fixture(fx):
    lex, pt = _lexicon(fx)
    assert pt == fx["phones"]
    ours = {w: i for w, i in lex.word_table.items()}
    ref = {w: i for w, i in fx["words"].items() if not w.startswith("#") and w not in ("<s>", "</s>")}   # disambiguation / LM symbols follow the words
    assert ours == ref


def test_text_topology_matches_reference_fixture(fx, tmp_path):
    topo = K.read_topology_text(fx["topo"])
    ours = SY.make_topology(fx["phones"])
    assert np.array_equal(topo.phones, ours.phones) and np.array_equal(topo.phone2idx[1:], ours.phone2idx[1:])
    # the fixture lists the non-silence entry first; compare entry by entry through the phone -> entry map
    for ph in (int(topo.phones[0]), int(topo.phones[-1])):
        a, b = topo.states_for(ph), ours.states_for(ph)
        assert len(a) == len(b)
        for sa, sb in zip(a, b):
            assert (sa.forward_pdf_class, sa.self_loop_pdf_class) == (sb.forward_pdf_class, sb.self_loop_pdf_class)
            assert [(d, round(p, 6)) for d, p in sa.transitions] == [(d, round(p, 6)) for d, p in sb.transitions]
    # binary round trip of the parsed text topology through the model writer / reader
    rng = np.random.default_rng(0)
    tree, n_pdfs = SY.make_tree(rng, topo, False, 0)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    am = K.AmDiagGmm(2, np.arange(n_pdfs + 1, dtype=np.int32), np.ones(n_pdfs, np.float32), np.zeros((n_pdfs, 2), np.float32), np.ones((n_pdfs, 2), np.float32))
    K.write_gmm_model(tmp_path / "t.mdl", tm, am)
    tm2, _ = K.read_gmm_model(tmp_path / "t.mdl")
    assert np.array_equal(tm2.tuples, tm.tuples) and tm2.num_tids == tm.num_tids


def _fixture_language(fx, words):
    """All (phone-id sequence -> cost) pairs the reference lexicon FST accepts for the word-id sequence `words`."""
    arcs, finals = {}, {}
    for ln in fx["lexfst"].strip().splitlines():
        p = ln.split()
        if len(p) <= 2:
            finals[int(p[0])] = float(p[1]) if len(p) == 2 else 0.0
        else:
            arcs.setdefault(int(p[0]), []).append((int(p[1]), fx["phones"][p[2]], fx["words"][p[3]], float(p[4]) if len(p) == 5 else 0.0))
    out = {}

    def rec(s, pos, seq, cost, depth):
        assert depth < 64
        if s in finals and pos == len(words):
            key = tuple(seq)
            out[key] = min(out.get(key, math.inf), cost + finals[s])
        for d, il, ol, w in arcs.get(s, []):
            if ol != 0:
                if pos >= len(words) or words[pos] != ol:
                    continue
                rec(d, pos + 1, seq + ([il] if il else []), cost + w, depth + 1)
            else:
                rec(d, pos, seq + ([il] if il else []), cost + w, depth + 1)
    rec(0, 0, [], 0.0, 0)
    return out


@pytest.mark.parametrize("text", ["worda", "wordb worda", "worda <unk> wordb", "!SIL worda"])
def test_compiled_graphs_accept_the_reference_lexicon_language(fx, text):
    lex, pt = _lexicon(fx)
    topo = K.read_topology_text(fx["topo"])
    # (the lexicon language does not depend on the HMMs: a small non-ergodic silence model instead of the fixture's 5-state one keeps the
    # exhaustive enumeration of HMM paths tractable, as in test_graph_compiler.py)
    topo.entries[int(topo.phone2idx[fx["phones"]["sil"]])] = [K.HmmState(0, 0, [(0, 0.5), (1, 0.25), (2, 0.25)]), K.HmmState(1, 1, [(2, 1.0)]),
                                                             K.HmmState(2, 2, [(2, 0.6), (3, 0.4)]), K.HmmState(-1, -1, [])]
    rng = np.random.default_rng(1)
    tree, n_pdfs = SY.make_tree(rng, topo, False, 0)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    words = [fx["words"][w] for w in text.split()]
    assert words == lex.to_int(text)
    ref = _fixture_language(fx, words)
    assert len(ref) == 2 ** (len(words) + 1)        # optional silence before, between and after the words
    fst = E.GraphCompiler(tm, tree, lex).compile([words]).export()[0]
    got = {}
    for arcs, cost in _paths(fst):
        tids = [int(fst.arc_ilabel[a]) for a in arcs]
        assert [int(fst.arc_olabel[a]) for a in arcs if fst.arc_olabel[a] != 0] == words
        seq = tuple(int(tm.tid2phone[t]) for t in tids if tm.is_final_tid[t])
        # several HMM paths (the silence model's skips) realise one phone sequence: all carry the same lexicon cost
        if seq in got:
            assert abs(got[seq] - cost) < 1e-5
        got[seq] = cost
    assert set(got) == set(ref)
    for k, c in ref.items():
        assert abs(got[k] - c) < 1e-5, (k, got[k], c)
