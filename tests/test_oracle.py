"""Pins the CPU oracle (oracle/oracle.c) before it is trusted as the checker: independent MFCC port (torchaudio golden),
closed-form / float64 numpy restatements, and exhaustive (no-beam) Viterbi."""
import numpy as np
import pytest

from helpers import build_synth_scenario, gold, load_model, oracle_align_all
from mfa_b200 import engine as E, kaldi_io as K
from oracle import oracle as O


def _relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_mfcc_matches_independent_kaldi_port():
    g = gold()
    pcm = g["acoustic_corpus_pcm"]
    # tolerance: north_star's "MFCC values agree within 1e-4 relative" (max-norm relative)
    assert _relmax(O.mfcc(pcm), g["ta_mfcc_snip"]) < 1e-4
    assert _relmax(O.mfcc(pcm, O.mfcc_opts(snip_edges=0)), g["ta_mfcc_nosnip"]) < 1e-4
    assert _relmax(O.mfcc(pcm, O.mfcc_opts(use_energy=1, energy_floor=1.0)), g["ta_mfcc_energy"]) < 1e-4
    assert O.mfcc(pcm).shape == (2670, 13) and O.mfcc(pcm, O.mfcc_opts(snip_edges=0)).shape == (2672, 13)
    assert O.mfcc(pcm[:399]).shape == (0, 13) and O.mfcc(pcm[:400]).shape == (1, 13)


def test_cmvn_deltas_splice_transform():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((57, 13)).astype(np.float32) * 5 + 3
    st = O.cmvn_stats([x[:20], x[20:]])
    assert np.allclose(st[0, :13], x.astype(np.float64).sum(0)) and st[0, 13] == 57
    assert np.allclose(st[1, :13], (x.astype(np.float64) ** 2).sum(0))
    y = O.cmvn_apply(x, st)
    assert np.abs(y.mean(0)).max() < 1e-5
    d = O.add_deltas(y)
    assert d.shape == (57, 39) and np.array_equal(d[:, :13], y)
    idx = lambda t: np.clip(t, 0, 56)
    t = np.arange(57)
    d1 = sum(k * y[idx(t + k)] for k in (-2, -1, 1, 2)) / 10.0
    assert np.allclose(d[:, 13:26], d1, atol=1e-5)
    k2 = np.convolve([-0.2, -0.1, 0, 0.1, 0.2], [-0.2, -0.1, 0, 0.1, 0.2])
    d2 = sum(k2[k + 4] * y[idx(t + k)] for k in range(-4, 5))
    assert np.allclose(d[:, 26:], d2, atol=1e-5)
    s = O.splice(y, 3, 3)
    assert s.shape == (57, 91) and np.array_equal(s[10, 39:52], y[10]) and np.array_equal(s[0, :13], y[0]) and np.array_equal(s[56, 78:], y[56])
    M = rng.standard_normal((40, 92)).astype(np.float32)
    assert np.allclose(O.transform(s, M), s @ M[:, :91].T + M[:, 91], atol=1e-4)
    assert np.allclose(O.transform(s, M[:, :91].copy()), s @ M[:, :91].T, atol=1e-4)


def test_gmm_loglikes_against_float64():
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((50, am.dim)) * 2).astype(np.float32)
    ll = O.gmm_loglikes(O.GmmModel.from_am(am), x)
    x64 = x.astype(np.float64)
    comp = am.gconsts.astype(np.float64)[None] + x64 @ am.means_invvars.astype(np.float64).T - 0.5 * (x64 ** 2) @ am.inv_vars.astype(np.float64).T
    ref = np.zeros((50, am.NumPdfs()))
    for j in range(am.NumPdfs()):
        c = comp[:, am.offsets[j]:am.offsets[j + 1]]
        mx = c.max(1, keepdims=True)
        ref[:, j] = (mx + np.log(np.exp(c - mx).sum(1, keepdims=True)))[:, 0]
    assert np.abs(ll - ref).max() / np.abs(ref).max() < 1e-5


def _exhaustive_viterbi(fst, tid_cost, loglikes, tid2pdf, acwt):
    """Float64 Viterbi over all states, no pruning (no epsilon arcs)."""
    S, T = fst.num_states, loglikes.shape[0]
    cost = np.full(S, np.inf)
    cost[fst.start] = 0.0
    w = fst.arc_weight.astype(np.float64) + tid_cost[fst.arc_ilabel].astype(np.float64)
    pdf = tid2pdf[fst.arc_ilabel]
    bps = []
    for t in range(T):
        cand = cost[fst.arc_src] + w - acwt * loglikes[t, pdf].astype(np.float64)
        new = np.full(S, np.inf)
        np.minimum.at(new, fst.arc_dst, cand)
        bp = np.full(S, -1, dtype=np.int64)
        for a in np.argsort(cand, kind="stable"):
            d = fst.arc_dst[a]
            if bp[d] < 0 and cand[a] == new[d]:
                bp[d] = a
        bps.append(bp)
        cost = new
    fin = cost + fst.finals.astype(np.float64)
    s = int(np.argmin(fin))
    total = fin[s]
    ali = np.zeros(T, dtype=np.int32)
    for t in range(T - 1, -1, -1):
        a = bps[t][s]
        ali[t] = fst.arc_ilabel[a]
        s = fst.arc_src[a]
    return total, ali


@pytest.mark.parametrize("triphone", [False, True])
def test_decoder_wide_beam_equals_exhaustive_search(triphone):
    sc = build_synth_scenario(seconds=16.0, seed=11, triphone=triphone, n_phones=8, n_words=30, target_pdfs=60, gauss_per_pdf=2)
    gc = E.GraphCompiler(sc["tm"], sc["tree"], sc["corpus"].lexicon)
    fsts = gc.compile(sc["corpus"].transcripts).export()
    res = oracle_align_all(sc, fsts, beam=1e4, retry_beam=0.0)
    g = O.GmmModel.from_am(sc["am"])
    tid_cost = -sc["tm"].scaled_transition_log_probs(1.0, 0.1)
    for u, r in enumerate(res):
        assert r["status"] == 0
        ll = O.gmm_loglikes(g, sc["feats"][u])
        total, ali = _exhaustive_viterbi(fsts[u], tid_cost, ll, sc["tm"].tid2pdf, 0.1)
        assert abs(-total / 0.1 - r["like"]) <= 1e-4 * abs(r["like"])
        assert (ali == r["ali"]).mean() >= 0.999
        assert list(r["words"]) == sc["corpus"].transcripts[u]
        # per-frame log-likelihoods are those of the aligned pdf
        assert np.allclose(r["per_frame"], ll[np.arange(len(ali)), sc["tm"].tid2pdf[r["ali"]]], rtol=1e-4, atol=1e-3)


def test_decoder_beam_and_retry_semantics(tmp_path):
    """Config 1 (the reference's sample utterance + fixture monophone model): beam 10 loses every final-state token,
    AlignUtteranceWrapper's second pass with retry_beam 40 succeeds."""
    from helpers import mono_sample_setup
    ms = mono_sample_setup(tmp_path)
    tm, am, lex = ms["tm"], ms["am"], ms["lex"]
    fst = E.GraphCompiler(tm, ms["tree"], lex).compile([lex.to_int(ms["text"])]).export()[0]
    m = O.mfcc(ms["pcm"])
    f = O.add_deltas(O.cmvn_apply(m, O.cmvn_stats([m])))
    g = O.GmmModel.from_am(am)
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    r_no_retry = O.align(fst, tc, g, tm.tid2pdf, f, f.shape[0], 0.1, 10.0, 0.0)
    r = O.align(fst, tc, g, tm.tid2pdf, f, f.shape[0], 0.1, 10.0, 40.0)
    r_wide = O.align(fst, tc, g, tm.tid2pdf, f, f.shape[0], 0.1, 400.0, 0.0)
    assert r_no_retry["status"] == 2 and r["status"] == 1 and r_wide["status"] == 0
    assert (r["ali"] == r_wide["ali"]).mean() > 0.99 and r["like"] <= r_wide["like"] + 1e-3 * abs(r_wide["like"])
    assert [lex.id2word[w] for w in r["words"]] == [w if w in lex.prons else "<unk>" for w in ms["text"].split()]
    # a feature sequence shorter than the shortest path can never reach a final state
    assert O.align(fst, tc, g, tm.tid2pdf, f[:3], 3, 0.1, 10.0, 40.0)["status"] == 2
    # lazily cached per-(frame,pdf) scoring == dense precomputed matrix
    rd = O.align(fst, tc, g, tm.tid2pdf, None, f.shape[0], 0.1, 10.0, 40.0, dense=O.gmm_loglikes(g, f))
    assert np.array_equal(rd["ali"], r["ali"]) and rd["like"] == r["like"]


def test_acc_stats_against_float64():
    sc = build_synth_scenario(seconds=10.0, seed=3, n_phones=6, n_words=20, gauss_per_pdf=3)
    tm, am = sc["tm"], sc["am"]
    gc = E.GraphCompiler(tm, sc["tree"], sc["corpus"].lexicon)
    fsts = gc.compile(sc["corpus"].transcripts).export()
    res = oracle_align_all(sc, fsts, beam=100, retry_beam=0)
    g = O.GmmModel.from_am(am)
    accs = None
    for u, r in enumerate(res):
        accs = O.acc_stats(g, tm.tid2pdf, sc["feats"][u], r["ali"], tm.num_tids, accs)
    x = np.concatenate(sc["feats"]).astype(np.float64)
    ali = np.concatenate([r["ali"] for r in res])
    occ = np.zeros(am.NumGauss()); mean = np.zeros((am.NumGauss(), am.dim)); var = np.zeros_like(mean); like = 0.0
    comp_all = am.gconsts.astype(np.float64)[None] + x @ am.means_invvars.astype(np.float64).T - 0.5 * (x ** 2) @ am.inv_vars.astype(np.float64).T
    for t in range(x.shape[0]):
        j = tm.tid2pdf[ali[t]]
        a, b = am.offsets[j], am.offsets[j + 1]
        c = comp_all[t, a:b]
        mx = c.max(); p = np.exp(c - mx); s = p.sum(); p /= s
        like += mx + np.log(s)
        occ[a:b] += p; mean[a:b] += p[:, None] * x[t]; var[a:b] += p[:, None] * x[t] ** 2
    assert np.allclose(accs["occ"], occ, rtol=1e-4, atol=1e-6)
    assert np.allclose(accs["mean"], mean, rtol=1e-4, atol=1e-4) and np.allclose(accs["var"], var, rtol=1e-4, atol=1e-3)
    assert abs(accs["like"][0] - like) < 1e-5 * abs(like)
    assert accs["trans"].sum() == x.shape[0] and abs(accs["occ"].sum() - x.shape[0]) < 1e-3


def test_gmm_loglikes_and_posteriors_against_sklearn():
    """An independent implementation of the same model: scikit-learn's diagonal GaussianMixture, loaded with the reference's own
    fixture model (tests/data/am/acoustic_g2p_output_model.zip -> golden arrays).  score_samples == per-pdf log-likelihood,
    predict_proba == the component posteriors behind the K4 / fMLLR statistics."""
    from sklearn.mixture import GaussianMixture
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(5)
    D = am.dim
    X = (am.means()[rng.integers(0, am.NumGauss(), size=64)] + 0.5 * rng.standard_normal((64, D))).astype(np.float32)
    ll = O.gmm_loglikes(O.GmmModel.from_am(am), X)
    g = O.GmmModel.from_am(am)
    checked = 0
    for j in list(range(0, am.NumPdfs(), max(1, am.NumPdfs() // 25)))[:25]:
        a, b = int(am.offsets[j]), int(am.offsets[j + 1])
        gm = GaussianMixture(n_components=b - a, covariance_type="diag")
        gm.weights_ = am.weights[a:b].astype(np.float64) / am.weights[a:b].sum()
        gm.means_ = am.means()[a:b].astype(np.float64)
        gm.covariances_ = am.variances()[a:b].astype(np.float64)
        gm.precisions_cholesky_ = 1.0 / np.sqrt(gm.covariances_)
        ref = gm.score_samples(X.astype(np.float64))
        assert np.abs(ll[:, j] - ref).max() <= 1e-4 * np.abs(ref).max(), j
        # posteriors through the accumulator: one frame at a time, occ == predict_proba
        tid = int(np.nonzero(tm.tid2pdf == j)[0][0])
        acc = O.acc_stats(g, tm.tid2pdf, X[:8], np.full(8, tid, np.int32), tm.num_tids)
        assert np.allclose(acc["occ"][a:b], gm.predict_proba(X[:8].astype(np.float64)).sum(0), rtol=1e-3, atol=1e-4)
        checked += 1
    assert checked >= 10


def test_avx2_and_generic_oracle_builds_agree_bit_for_bit():
    """oracle/Makefile builds oracle.c twice (-O2 generic; -O3 -mavx2, the one bench.py's CPU arm uses on AVX2 hosts).  No fused or
    re-associated arithmetic in either (-ffp-contract=off, no -ffast-math): MFCC, log-likelihoods and a beam alignment must be identical
    down to the last bit, otherwise the faster build could not stand in as the checker."""
    import os
    import subprocess
    import sys
    if O.VARIANT != "avx2":
        pytest.skip("host without AVX2: only the generic build is in use")
    code = (
        "import sys, zlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from helpers import build_synth_scenario, oracle_align_all\n"
        "from oracle import oracle as O\n"
        "from mfa_b200 import engine as E\n"
        "sc = build_synth_scenario(seconds=20.0, seed=11, triphone=True, n_phones=8, n_words=30, target_pdfs=60, gauss_per_pdf=3)\n"
        "fsts = E.GraphCompiler(sc['tm'], sc['tree'], sc['corpus'].lexicon).compile(sc['corpus'].transcripts).export()\n"
        "ref = oracle_align_all(sc, fsts, 10.0, 40.0)\n"
        "g = O.GmmModel.from_am(sc['am'])\n"
        "h = 0\n"
        "for f in sc['feats']: h = zlib.crc32(np.ascontiguousarray(f).tobytes(), h); h = zlib.crc32(np.ascontiguousarray(O.gmm_loglikes(g, f)).tobytes(), h)\n"
        "for r in ref: h = zlib.crc32(np.asarray(r['ali'], np.int32).tobytes(), h); h = zlib.crc32(np.asarray(r['per_frame'], np.float32).tobytes(), h); h = zlib.crc32(np.float64(r['like']).tobytes(), h)\n"
        "print(O.VARIANT, h)\n"
    ) % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = {}
    for generic in ("", "1"):
        env = dict(os.environ)
        env.pop("MFA_ORACLE_GENERIC", None)
        if generic:
            env["MFA_ORACLE_GENERIC"] = "1"
        variant, crc = subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.split()[-2:]
        outs[variant] = crc
    assert set(outs) == {"avx2", "generic"} and outs["avx2"] == outs["generic"], outs
