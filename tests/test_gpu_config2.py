"""Parity at the BENCHMARKED shape (BASELINE.json configs[1]): 4 000 pdfs / ~40 000 Gaussians / D = 40 (splice +-3 + 40x91 projection),
LibriSpeech-shaped utterances up to 30 s, beam 10 / retry 40 -- the very scenario bench.py times (mfa_b200.scenario.build, seed 1234).

The oracle (oracle/oracle.c) finishes a sample of it in seconds on the host cores; north_star's bars are written into the asserts:
transition-ids identical on >= 99.9 % of frames, per-utterance and per-value log-likelihoods within 1e-4 relative, phone boundaries
within one 10 ms frame, words / statuses exact.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from mfa_b200 import engine as E, kalpy_compat as KC, scenario as SC
from oracle import oracle as O

pytestmark = pytest.mark.gpu

HOURS = float(os.environ.get("MFA_TEST_CONFIG2_HOURS", "10"))


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def sc(eng):
    import torch
    s = SC.build(eng, HOURS * 3600.0, seed=1234, target_pdfs=4000, gauss_per_pdf=10, n_threads=os.cpu_count() or 8,
                 synth_device=torch.device("cuda", 0))
    yield s
    s.model.close()


def _oracle_feats(sc, utts, pool):
    """Oracle feature chain for `utts` (+ every utterance of their speakers, whose frames enter the CMVN statistics)."""
    c = sc.corpus
    spk = sorted({int(c.utt2spk[u]) for u in utts})
    need = [u for u in range(c.n_utts) if int(c.utt2spk[u]) in spk]
    opts = O.mfcc_opts()
    raw = dict(zip(need, pool.map(lambda u: O.mfcc(c.pcm[c.sample_off[u]:c.sample_off[u + 1]], opts), need)))
    stats = {s: O.cmvn_stats([raw[u] for u in need if c.utt2spk[u] == s]) for s in spk}
    return {u: O.transform(O.splice(O.cmvn_apply(raw[u], stats[int(c.utt2spk[u])]), 3, 3), sc.lda) for u in utts}


def test_k2_dense_and_ffma_at_config2_model_size(eng, sc):
    """K2 on the 4 000-pdf / 40 k-Gaussian model, >= 20 k frames of the corpus' own features: the tcgen05 kernel (dense tiling, 316 tiles,
    ragged pdf boundaries) and the fp32 CUDA-core kernel against the oracle's Kaldi-order evaluation (f64 LogSumExp with the
    -15.94 cutoff), per value."""
    c = sc.corpus
    n_utts = 0
    while sc.frame_off[n_utts] < 20480:
        n_utts += 1
    T = int(sc.frame_off[n_utts])
    with ThreadPoolExecutor(os.cpu_count() or 8) as pool:
        feats = _oracle_feats(sc, list(range(n_utts)), pool)
        x = np.concatenate([feats[u] for u in range(n_utts)]).astype(np.float32)
        assert x.shape == (T, 40) and T >= 20480
        g = O.GmmModel.from_am(sc.am)
        step = 256
        ref = np.concatenate(list(pool.map(lambda i: O.gmm_loglikes(g, x[i:i + step]), range(0, T, step))))
    assert sc.am.NumPdfs() == 4000 and sc.am.NumGauss() > 35000
    for impl in (0, 1):
        ll = sc.model.loglikes(x, impl=impl)
        err = np.abs(ll - ref)
        rel = err / np.maximum(np.abs(ref), 1.0)
        assert rel.max() <= 1e-4, (impl, float(rel.max()), float(err.max()))
        # the deviations of the tensor-core path (ex2/lg2.approx, fp32 sums, no -15.94 cutoff, truncating accumulation) stay ~1e-6 relative
        assert rel.mean() <= 5e-6, (impl, float(rel.mean()))


def test_fused_align_pcm_at_config2_scale_against_oracle(eng, sc):
    """The whole 10 h batch through mfa_align_pcm (ragged per-utterance K2 tiles, band Viterbi, retry beam); >= 50 utterances of it --
    the longest ones, every utterance that needed the retry beam or failed, and a spread of the rest -- against the oracle."""
    import torch
    c = sc.corpus
    dev = torch.device("cuda", 0)
    d_pcm = torch.from_numpy(c.pcm).to(dev)
    res = E.align_pcm(eng, sc.model, sc.graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), sc.feat_mode, lda=sc.lda,
                      workspace_bytes=100 << 30)
    eng.sync()
    ali, pf, words, nw, tl, st = (x.cpu().numpy() for x in (res.ali, res.per_frame, res.words, res.num_words, res.total_like, res.status))
    fo, wo = sc.frame_off, res.word_off
    T_u = fo[1:] - fo[:-1]
    special = [int(u) for u in np.nonzero(st != 0)[0]]
    longest = [int(u) for u in np.argsort(-T_u)[:8]]
    spread = [int(u) for u in np.linspace(0, c.n_utts - 1, 48).astype(int)]
    utts = sorted(set(special + longest + spread))
    assert len(utts) >= 50
    if HOURS >= 10:
        assert any(st[u] == 1 for u in utts), "the 10 h corpus holds an utterance that needs the retry beam"
    fsts = sc.batch.export()
    g = O.GmmModel.from_am(sc.am)
    tid_cost = -sc.tm.scaled_transition_log_probs(1.0, 0.1)
    with ThreadPoolExecutor(os.cpu_count() or 8) as pool:
        feats = _oracle_feats(sc, utts, pool)
        ref = list(pool.map(lambda u: O.align(O.FstCsr(fsts[u]), tid_cost, g, sc.tm.tid2pdf, feats[u], feats[u].shape[0], 0.1, 10.0, 40.0), utts))
    same = total = b_ok = b_tot = 0
    for u, r in zip(utts, ref):
        assert int(st[u]) == r["status"], (u, int(st[u]), r["status"])
        if r["status"] >= 2:
            continue
        a = ali[fo[u]:fo[u + 1]]
        eq = a == r["ali"]
        same += int(eq.sum()); total += len(a)
        assert list(words[wo[u]:wo[u] + nw[u]]) == list(r["words"]), u
        assert abs(float(tl[u]) - r["like"]) <= 1e-4 * abs(r["like"]), (u, float(tl[u]), r["like"])
        d = np.abs(pf[fo[u]:fo[u + 1]][eq] - r["per_frame"][eq])
        assert np.all(d <= 1e-4 * np.maximum(1.0, np.abs(r["per_frame"][eq]))), (u, float(d.max()))
        cg = KC.Alignment(str(u), a, [], float(tl[u])).generate_ctm(sc.tm, None)
        cr = KC.Alignment(str(u), r["ali"], [], r["like"]).generate_ctm(sc.tm, None)
        assert [x.label for x in cg] == [x.label for x in cr], u
        b_tot += 2 * len(cr)
        b_ok += sum(int(abs(x.begin - y.begin) <= 0.0101) + int(abs(x.end - y.end) <= 0.0101) for x, y in zip(cg, cr))
    assert total > 50000 and same / total >= 0.999, (same, total)
    assert b_ok / b_tot >= 0.999, (b_ok, b_tot)
    assert eng.band_fallbacks >= 0
