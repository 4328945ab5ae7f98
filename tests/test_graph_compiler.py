"""Host training-graph compiler (csrc/graph.cc) through the C ABI: path-set equality with a brute-force enumeration of the
lexicon semantics, context-dependent pdf selection via the tree, reorder=true self-loop placement, error behaviour."""
import math

import numpy as np
import pytest

from mfa_b200 import engine as E, kaldi_io as K, lexicon as LX, synth as SY, _lib as L


def _setup(triphone, seed=0, position_dependent=False):
    rng = np.random.default_rng(seed)
    phones = ["a", "b", "c", "d"]
    pt = LX.make_phone_table(phones, ("sil", "spn"), position_dependent)
    prons = {"x": [LX.Pron(["a", "b"]), LX.Pron(["a", "c", "d"], 0.5)], "y": [LX.Pron(["d"])], "z": [LX.Pron(["b", "a"]), LX.Pron(["c", "a"], 0.25)]}
    lex = LX.Lexicon(prons, pt, silence_probability=0.3, initial_silence_probability=0.6, position_dependent_phones=position_dependent)
    topo = SY.make_topology(pt)
    # a smaller non-Bakis silence model (skip + no self-loop on state 1) keeps the exhaustive path enumeration tractable
    topo.entries[1] = [K.HmmState(0, 0, [(0, 0.5), (1, 0.25), (2, 0.25)]), K.HmmState(1, 1, [(2, 1.0)]), K.HmmState(2, 2, [(2, 0.6), (3, 0.4)]),
                       K.HmmState(-1, -1, [])]
    tree, n_pdfs = SY.make_tree(rng, topo, triphone, 40)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    return lex, pt, topo, tree, tm


def _paths(fst, max_paths=200000):
    """All simple start->final paths over non-self-loop arcs: yields (tids, olabels, cost)."""
    out_arcs = [[] for _ in range(fst.num_states)]
    for a in range(fst.arc_src.shape[0]):
        if fst.arc_src[a] != fst.arc_dst[a]:
            out_arcs[fst.arc_src[a]].append(a)
    res = []
    stack = [(fst.start, [], 0.0, {fst.start})]
    while stack:
        s, arcs, c, seen = stack.pop()
        if np.isfinite(fst.finals[s]):
            res.append((arcs, c + float(fst.finals[s])))
            assert len(res) < max_paths
        for a in out_arcs[s]:
            d = int(fst.arc_dst[a])
            if d in seen:
                continue
            stack.append((d, arcs + [a], c + float(fst.arc_weight[a]), seen | {d}))
    return res


def _brute(lex, words):
    """(phone-id sequence) -> cost, straight from the lexicon definition (lexicon.text.fst layout)."""
    sil = lex.phone_table["sil"]
    c = lambda p: -math.log(p)
    out = {}

    def rec(i, seq, cost):
        if i == len(words):
            out[tuple(seq)] = cost
            return
        for pr in range(lex._arrs[0][words[i]], lex._arrs[0][words[i] + 1]):
            ph = [int(x) for x in lex._arrs[2][lex._arrs[1][pr]:lex._arrs[1][pr + 1]]]
            pc = float(lex._arrs[3][pr])
            rec(i + 1, seq + ph, cost + pc + c(1 - lex.silence_probability))
            rec(i + 1, seq + ph + [sil], cost + pc + c(lex.silence_probability))

    rec(0, [], c(1 - lex.initial_silence_probability))
    rec(0, [sil], c(lex.initial_silence_probability))
    return out


@pytest.mark.parametrize("triphone,posdep", [(False, False), (True, False), (True, True)])
def test_path_set_equals_lexicon_language(triphone, posdep):
    lex, pt, topo, tree, tm = _setup(triphone, position_dependent=posdep)
    words = [lex.word_table[w] for w in ("x", "z", "y")]
    fst = E.GraphCompiler(tm, tree, lex).compile([words]).export()[0]
    brute = _brute(lex, words)
    got = {}
    for arcs, cost in _paths(fst):
        tids = [int(fst.arc_ilabel[a]) for a in arcs]
        assert all(t > 0 for t in tids)
        ols = [int(fst.arc_olabel[a]) for a in arcs if fst.arc_olabel[a] != 0]
        assert ols == words
        # split into phone instances at transitions into the HMM's final state
        seq, inst, cur = [], [], []
        for t in tids:
            cur.append(t)
            if tm.is_final_tid[t]:
                seq.append(int(tm.tid2phone[t])); inst.append(cur); cur = []
        assert not cur
        # every instance: consistent phone, consecutive HMM path starting in state 0, pdfs picked by the tree for (l, p, r)
        for k, ts in enumerate(inst):
            ph = seq[k]
            l = seq[k - 1] if k > 0 else 0
            r = seq[k + 1] if k + 1 < len(seq) else 0
            hs = 0
            for t in ts:
                tstate = tm.id2state[t]
                p_, h_, fpdf, _ = tm.tuples[tstate - 1]
                assert p_ == ph and h_ == hs and not tm.is_self_loop[t]
                st = topo.states_for(ph)[hs]
                assert fpdf == tree.lookup([l, ph, r] if triphone else [ph], st.forward_pdf_class)
                hs = st.transitions[t - tm.state2id[tstate]][0]
        key = tuple(seq)
        if key in got:
            assert abs(got[key] - cost) < 1e-4
        got[key] = cost
    assert set(got) == set(brute)
    for k in brute:
        assert abs(got[k] - brute[k]) < 1e-4, (k, got[k], brute[k])


def test_self_loops_follow_reorder_convention():
    lex, pt, topo, tree, tm = _setup(True)
    fst = E.GraphCompiler(tm, tree, lex).compile([[lex.word_table["x"], lex.word_table["y"]]]).export()[0]
    loops = {}
    for a in range(fst.arc_src.shape[0]):
        if fst.arc_src[a] == fst.arc_dst[a]:
            assert fst.arc_dst[a] not in loops, "one self-loop per state"
            loops[int(fst.arc_dst[a])] = int(fst.arc_ilabel[a])
            assert tm.is_self_loop[fst.arc_ilabel[a]] and fst.arc_weight[a] == 0 and fst.arc_olabel[a] == 0
    # reorder=true: the self-loop on a state belongs to the transition-state of every arc ENTERING it
    for a in range(fst.arc_src.shape[0]):
        s, d, t = int(fst.arc_src[a]), int(fst.arc_dst[a]), int(fst.arc_ilabel[a])
        if s == d:
            continue
        sl = tm.self_loop_tid[tm.id2state[t]]
        assert loops.get(d, 0) == sl
    assert fst.start not in loops


def test_batch_compile_threads_and_errors():
    lex, pt, topo, tree, tm = _setup(False)
    gc = E.GraphCompiler(tm, tree, lex)
    seqs = [[lex.word_table["x"]], [], [lex.word_table["y"], lex.word_table["z"], lex.word_table["x"]]] * 7
    a = gc.compile(seqs, n_threads=1).export()
    b = gc.compile(seqs, n_threads=4).export()
    for f1, f2 in zip(a, b):
        assert np.array_equal(f1.arc_ilabel, f2.arc_ilabel) and np.array_equal(f1.arc_weight, f2.arc_weight) and f1.start == f2.start
    # empty transcript: optional silence only
    e = a[1]
    assert e.start >= 0 and np.isfinite(e.finals).any()
    with pytest.raises(L.MfaError):
        gc.compile([[len(lex.word_table) + 5]])
    # round trip through the OpenFst binary container used by fsts.*.ark
    import io
    buf = io.BytesIO()
    K.write_fst(buf, a[2])
    back = K.read_fst(io.BytesIO(buf.getvalue()))
    rt = E.FstBatch.from_fsts([back]).export()[0]
    assert rt.num_states == a[2].num_states and np.array_equal(np.sort(rt.arc_ilabel), np.sort(a[2].arc_ilabel))


def _check_band(fst, tm, bv, u):
    """The band view is a renumbering of the same graph: a permutation, arcs grouped by destination, reach fields valid."""
    s0, s1, a0, a1 = int(bv["state_off"][u]), int(bv["state_off"][u + 1]), int(bv["arc_off"][u]), int(bv["arc_off"][u + 1])
    S, A = s1 - s0, a1 - a0
    assert S == fst.num_states and A == fst.arc_src.shape[0]
    orig = bv["orig_state"][s0:s1].astype(np.int64)
    assert sorted(orig.tolist()) == list(range(S))              # permutation
    pos = np.empty(S, np.int64)
    pos[orig] = np.arange(S)
    assert int(bv["start"][u]) == pos[fst.start]
    stw, apk, aidx = bv["state_word"][s0:s1], bv["arc_word"][a0:a1], bv["arc_index"][a0:a1].astype(np.int64)
    begin, cnt, reach = (stw & 0xFFFF).astype(np.int64), ((stw >> 16) & 0xFF).astype(np.int64), (stw >> 24).astype(np.int64)
    assert begin[0] == 0 and np.array_equal(begin[1:], np.cumsum(cnt)[:-1]) and cnt.sum() == A
    # by-source order of the packed graph = stable sort of the fst's arcs by source state
    by_src = np.argsort(fst.arc_src, kind="stable")
    assert sorted(aidx.tolist()) == list(range(A))              # every arc exactly once
    tid2pdf = np.maximum(tm.tid2pdf, 0)
    maxback, fwd = 0, np.zeros(S, np.int64)
    for d in range(S):
        lst = aidx[begin[d]:begin[d] + cnt[d]]
        assert np.all(np.diff(lst) > 0)                         # ties resolve to the lowest by-source arc index
        for j, k in zip(range(begin[d], begin[d] + cnt[d]), lst):
            a = by_src[k]
            assert pos[fst.arc_dst[a]] == d and (apk[j] & 0xFFFF) == pos[fst.arc_src[a]]
            ps = int(pos[fst.arc_src[a]])
            fwd[ps] = max(fwd[ps], d - ps)
            maxback = max(maxback, ps - d)
    assert np.array_equal(fwd, reach) and maxback == int(bv["maxback"][u])
    # local pdf ids in the arc words index the utterance's sorted pdf list
    pdfs = np.unique(tid2pdf[fst.arc_ilabel[fst.arc_ilabel > 0]])
    for j in range(A):
        a = by_src[aidx[j]]
        assert pdfs[apk[j] >> 16] == tid2pdf[fst.arc_ilabel[a]]
    return maxback, int(reach.max())


@pytest.mark.parametrize("triphone", [False, True])
def test_band_view_is_a_forward_renumbering(triphone):
    lex, pt, topo, tree, tm = _setup(triphone)
    # the default 5-state silence model: its ergodic middle states are the only cycles (besides self-loops) in a training graph
    rng = np.random.default_rng(3)
    topo = SY.make_topology(pt)
    tree, n_pdfs = SY.make_tree(rng, topo, triphone, 40)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    w = lex.word_table
    seqs = [[w["x"], w["y"], w["z"], w["x"]], [], [w["z"]] * 6, [w["y"]]]
    batch = E.GraphCompiler(tm, tree, lex).compile(seqs)
    fsts = batch.export()
    bv = E.Graphs(batch, tm).band_view()
    assert bv["band_ok"].tolist() == [1] * len(seqs)
    for u, f in enumerate(fsts):
        maxback, reach = _check_band(f, tm, bv, u)
        assert maxback <= 16 and reach <= 255


def test_band_view_refuses_epsilon_graphs():
    lex, pt, topo, tree, tm = _setup(False)
    f = E.GraphCompiler(tm, tree, lex).compile([[lex.word_table["x"]]]).export()[0]
    # splice an input-epsilon arc in front of the start state (as a Kaldi fsts.ark graph may have)
    g = K.Fst(f.num_states, f.num_states + 1, np.append(f.arc_src, f.num_states).astype(np.int32), np.append(f.arc_ilabel, 0).astype(np.int32),
              np.append(f.arc_olabel, 0).astype(np.int32), np.append(f.arc_dst, f.start).astype(np.int32),
              np.append(f.arc_weight, 0.25).astype(np.float32), np.append(f.finals, np.inf).astype(np.float32))
    bv = E.Graphs(E.FstBatch.from_fsts([g, f]), tm).band_view()
    assert bv["band_ok"].tolist() == [0, 1]


def test_fst_archive_c_parser_equals_python_reader(tmp_path):
    """kaldi_io.read_fst_ark (header in Python, state / arc body by mfa_fst_body_scan / _fill) against the pure-Python read_ark on a
    written archive, including an empty FST and a state without arcs; truncated input is refused."""
    import io
    from mfa_b200 import _lib as L
    rng = np.random.default_rng(5)
    fsts = []
    for n_states in (1, 7, 40, 0):
        na = 0 if n_states == 0 else int(rng.integers(0, 4 * n_states + 1))
        src = np.sort(rng.integers(0, max(n_states, 1), na)).astype(np.int32)
        f = K.Fst(0 if n_states else -1, n_states, src, rng.integers(0, 50, na).astype(np.int32), rng.integers(0, 9, na).astype(np.int32),
                  rng.integers(0, max(n_states, 1), na).astype(np.int32), rng.random(na).astype(np.float32),
                  np.where(rng.random(n_states) < 0.3, rng.random(n_states), np.inf).astype(np.float32))
        fsts.append(f)
    w = K.ArkWriter(str(tmp_path / "fsts.ark"))
    for i, f in enumerate(fsts):
        w.write_fst(f"utt-{i}", f)
    w.close()
    a = list(K.read_ark(str(tmp_path / "fsts.ark"), "fst"))
    b = K.read_fst_ark(str(tmp_path / "fsts.ark"))
    assert [k for k, _ in a] == [k for k, _ in b] == [f"utt-{i}" for i in range(len(fsts))]
    for (_, x), (_, y), z in zip(a, b, fsts):
        assert (x.start, x.num_states) == (y.start, y.num_states) == (z.start, z.num_states)
        for name in ("arc_src", "arc_ilabel", "arc_olabel", "arc_dst", "arc_weight", "finals"):
            assert np.array_equal(getattr(x, name), getattr(y, name)) and np.array_equal(getattr(y, name), getattr(z, name)), name
    raw = (tmp_path / "fsts.ark").read_bytes()
    (tmp_path / "cut.ark").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(L.MfaError, match="truncated FST"):
        K.read_fst_ark(str(tmp_path / "cut.ark"))
