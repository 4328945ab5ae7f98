"""Host-side pieces added for rows a11 (equal alignment) and N2 (fMLLR update): no GPU needed.

* the glibc rand() restatement behind mfa_equal_align == libc srand()/rand();
* mfa_equal_align (product, C ABI) == the oracle's EqualAlign (libc rand) bit for bit, incl. the failure cases;
* equal alignments are valid paths of the graph with the frames spread evenly over the self-loops;
* the numpy fMLLR row update == the oracle's restatement of ComputeFmllrMatrixDiagGmmFull, recovers a known transform,
  never lowers the auxiliary function, honours min_count; compose_transforms == matrix algebra.
"""
import ctypes as C

import numpy as np
import pytest

from mfa_b200 import engine as E, fmllr as F, kaldi_io as K, kalpy_compat as KC, lexicon as LX, synth as SY, _lib as L
from oracle import oracle as O


def test_rand_restatement_matches_libc():
    libc = C.CDLL("libc.so.6")
    for seed in (0, 1, 42, 123456789, 0xFFFFFFFF, KC.string_hash("12-345")):
        out = np.zeros(500, np.int32)
        L.check(L.lib().mfa_rand_sequence(C.c_uint32(seed), C.c_int32(500), out.ctypes.data_as(C.c_void_p)))
        libc.srand(C.c_uint(seed))
        ref = [libc.rand() for _ in range(500)]
        assert out.tolist() == ref, seed


def test_string_hash_is_kaldi_stringhasher():
    assert KC.string_hash("") == 0 and KC.string_hash("a") == 97 and KC.string_hash("ab") == (97 * 7853 + 98)
    assert KC.string_hash("speaker1-utt000123") < 2 ** 32


def _graphs(triphone=False, n=12, seed=3):
    rng = np.random.default_rng(seed)
    lex, _ = SY.make_lexicon(rng, n_phones=10, n_words=50)
    topo = SY.make_topology(lex.phone_table)
    tree, n_pdfs = SY.make_tree(rng, topo, triphone, 60)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    words = list(lex.word_table.values())
    seqs = [[int(w) for w in rng.choice([w for w in words if w > 0], size=int(rng.integers(1, 7)))] for _ in range(n)]
    gc = E.GraphCompiler(tm, tree, lex)
    return tm, gc.compile(seqs), seqs


def _check_path(fst, ali, words):
    """ali/words must label a start->final path of fst."""
    out = {}
    for a in range(fst.arc_src.shape[0]):
        out.setdefault(int(fst.arc_src[a]), []).append(a)
    # the graphs have no input-epsilon arcs, so every arc consumes one frame: simple frontier search over (state, word position)
    frontier = {(fst.start, 0)}
    for tid in ali:
        nxt = set()
        for s, wp in frontier:
            for a in out.get(s, []):
                if int(fst.arc_ilabel[a]) != tid:
                    continue
                ol = int(fst.arc_olabel[a])
                if ol == 0:
                    nxt.add((int(fst.arc_dst[a]), wp))
                elif wp < len(words) and words[wp] == ol:
                    nxt.add((int(fst.arc_dst[a]), wp + 1))
        frontier = nxt
        assert frontier
    assert any(np.isfinite(fst.finals[s]) and wp == len(words) for s, wp in frontier)


@pytest.mark.parametrize("triphone", [False, True])
def test_equal_align_matches_oracle_and_is_a_valid_path(triphone):
    tm, batch, seqs = _graphs(triphone)
    fsts = batch.export()
    n = len(fsts)
    assert all((f.arc_ilabel != 0).all() for f in fsts)
    rng = np.random.default_rng(5)
    T = rng.integers(40, 400, size=n)
    T[0] = 0          # zero frames
    T[1] = 2          # shorter than any path: EqualAlign fails after its retries
    fo = np.zeros(n + 1, np.int64); fo[1:] = np.cumsum(T)
    seeds = [KC.string_hash(f"{u % 3}-{u}") for u in range(n)]
    ali, words, wo, nw, st = batch.equal_align(fo, seeds)
    assert st[0] == 4 and st[1] == 2
    for u in range(n):
        r = O.equal_align(fsts[u], int(T[u]), seeds[u])
        assert r["status"] == st[u], u
        if st[u] != 0:
            continue
        a = ali[fo[u]:fo[u + 1]]
        w = words[wo[u]:wo[u] + nw[u]]
        assert (a == r["ali"]).all() and w.tolist() == r["words"].tolist()
        assert w.tolist() == seqs[u]                       # the olabels are the transcript
        _check_path(fsts[u], a.tolist(), w.tolist())
        # even spread: run lengths of self-loop tids on the path differ by at most one
        runs, k = [], 0
        is_loop = np.zeros(tm.num_tids + 1, bool)
        is_loop[tm.self_loop_tid[tm.self_loop_tid > 0]] = True
        for t in a:
            if is_loop[t]:
                k += 1
            else:
                if k:
                    runs.append(k)
                k = 0
        if k:
            runs.append(k)
        assert max(runs) - min(runs) <= 1
    # kalpy-shaped wrapper: same result, (None, None) on failure
    a0, w0 = KC.gmm_align_equal(fsts[2], np.zeros((int(T[2]), 3), np.float32), utterance_id="0-2")
    r = O.equal_align(fsts[2], int(T[2]), KC.string_hash("0-2"))
    assert a0 == r["ali"].tolist() and w0 == r["words"].tolist()
    assert KC.gmm_align_equal(fsts[1], np.zeros((2, 3), np.float32), "x") == (None, None)


def _toy_gmm(rng, D, P, M):
    off = np.arange(0, P * M + 1, M).astype(np.int32)
    G = int(off[-1])
    means = rng.normal(size=(G, D))
    var = np.exp(rng.normal(scale=0.3, size=(G, D)))
    w = rng.dirichlet(np.ones(M), size=P).ravel()
    iv = 1.0 / var
    gc = np.log(w) - 0.5 * (D * np.log(2 * np.pi) + np.log(var).sum(1) + (means * means * iv).sum(1))
    return O.GmmModel(D, off, gc, means * iv, iv), means, var


def test_fmllr_update_matches_oracle_and_recovers_transform():
    rng = np.random.default_rng(0)
    D, P, M, T = 13, 6, 3, 6000
    g, means, var = _toy_gmm(rng, D, P, M)
    tid2pdf = np.concatenate([[0], np.arange(P)]).astype(np.int32)
    stats = []
    truth = []
    for spk in range(3):
        ali = rng.integers(1, P + 1, size=T).astype(np.int32)
        comp = rng.integers(0, M, size=T)
        A = np.eye(D) + 0.1 * rng.normal(size=(D, D))
        b = rng.normal(size=D) * 0.3
        x = means[(ali - 1) * M + comp] + rng.normal(size=(T, D)) * np.sqrt(var[(ali - 1) * M + comp])
        y = ((x - b) @ np.linalg.inv(A).T).astype(np.float32)      # A y + b = x
        n = T if spk < 2 else 300                                   # third speaker: below min_count
        stats.append(O.fmllr_acc(g, g, tid2pdf, None, y[:n], ali[:n]))
        truth.append(np.hstack([A, b[:, None]]))
    stats = np.stack(stats)
    W, impr, count = F.compute_transforms(stats, D)
    assert np.allclose(count, [T, T, 300], rtol=1e-5)
    for s in range(3):
        Wo, io = O.fmllr_update(stats[s], D)
        assert np.abs(W[s] - Wo).max() <= 1e-5 and abs(impr[s] - io) <= 1e-4 * max(1.0, abs(io))
    assert np.abs(W[0] - truth[0]).max() < 0.08 and np.abs(W[1] - truth[1]).max() < 0.08
    assert (W[2] == np.eye(D, D + 1, dtype=np.float32)).all() and impr[2] == 0.0    # unit transform kept
    assert impr[0] > 0 and impr[1] > 0
    # the auxiliary function at the estimate is not below its value at the unit transform (each row update maximises it)
    beta, Kk, G = F.unpack_stats(stats[:2], D)
    unit = np.tile(np.eye(D, D + 1), (2, 1, 1))
    assert (F.aux_function(W[:2].astype(np.float64), beta, Kk, G) >= F.aux_function(unit, beta, Kk, G)).all()
    # silence weighting: weight-0 transition-ids drop out of the statistics, fractional weights scale them
    ali = rng.integers(1, P + 1, size=500).astype(np.int32)
    y = rng.normal(size=(500, D)).astype(np.float32)
    tw = np.ones(P + 1, np.float32); tw[1] = 0.0; tw[2] = 0.5
    s_w = O.fmllr_acc(g, g, tid2pdf, tw, y, ali)
    keep = ali != 1
    s_a = O.fmllr_acc(g, g, tid2pdf, None, y[keep & (ali != 2)], ali[keep & (ali != 2)])
    s_b = O.fmllr_acc(g, g, tid2pdf, None, y[ali == 2], ali[ali == 2])
    assert np.allclose(s_w, s_a + 0.5 * s_b, rtol=1e-5, atol=1e-5)


def test_compose_transforms():
    rng = np.random.default_rng(2)
    D = 5
    a, b = rng.normal(size=(D, D + 1)).astype(np.float32), rng.normal(size=(D, D + 1)).astype(np.float32)
    x = rng.normal(size=D)
    c = F.compose_transforms(a, b)
    y = b[:, :D] @ x + b[:, D]
    assert np.allclose(c[:, :D] @ x + c[:, D], a[:, :D] @ y + a[:, D], atol=1e-5)
