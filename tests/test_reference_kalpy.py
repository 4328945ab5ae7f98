"""Upgrade path of SURVEY.md section 8c: parity against the REAL reference stack (kalpy over Kaldi), stage by stage.

The vectors come from tools/make_kalpy_fixtures.py, which needs a machine where ``import kalpy`` is the real package (not installable in
the offline build image: there the file tests/golden/kalpy_fixtures.npz does not exist, every test here skips, and DESIGN.md says "parity
unpinned").  With the file present -- one command on any kalpy box -- the ORACLE (CPU tests) and the ENGINE (``-m gpu`` tests) are diffed
against Kaldi's own outputs on the reference's fixture model + sample utterance: MFCC, CMVN statistics, delta / splice+LDA features, all-pdf
log-likelihoods (gmm_compute_likes), the compiled training graph's language, GmmAligner.align_utterance (transition-ids, words, likelihood,
phone CTM), gmm_align_equal, and the accumulator statistics (transition counts, total log-likelihood).
Tolerances are north_star's: 1e-4 relative for MFCC / log-likelihoods, >= 99.9 % identical transition-ids, boundaries within one frame."""
import json
import os

import numpy as np
import pytest

PATH = os.environ.get("MFA_KALPY_FIXTURES") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kalpy_fixtures.npz")
if not os.path.exists(PATH):
    pytest.skip("no vectors from a real kalpy / Kaldi install (tools/make_kalpy_fixtures.py): oracle parity only, 'parity unpinned' "
                "(DESIGN.md section 2)", allow_module_level=True)

from helpers import load_model, mono_sample_setup  # noqa: E402
from mfa_b200 import kaldi_io as K  # noqa: E402
from oracle import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def fx():
    g = np.load(PATH)
    d = {k: g[k] for k in g.files}
    d["errors"] = json.loads(bytes(d["errors"]).decode())
    return d


def need(fx, *keys):
    for k in keys:
        if k not in fx:
            pytest.skip(f"the kalpy run did not produce {k!r}: {fx['errors']}")


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(1e-30, np.abs(np.asarray(b, np.float64)).max()))


def _oracle_feats(fx):
    m = O.mfcc(fx["pcm"])
    st = O.cmvn_stats([m])
    return m, st, O.add_deltas(O.cmvn_apply(m, st))


def test_oracle_mfcc_cmvn_features_against_kalpy(fx):
    need(fx, "pcm", "mfcc")
    m, st, f = _oracle_feats(fx)
    assert m.shape == fx["mfcc"].shape and rel(m, fx["mfcc"]) <= 1e-4
    if "cmvn_stats" in fx:
        assert np.allclose(st, fx["cmvn_stats"], rtol=1e-6, atol=1e-3)
    if "feats_deltas" in fx:
        assert rel(f, fx["feats_deltas"]) <= 1e-4
    if "feats_lda" in fx:
        got = O.transform(O.splice(O.cmvn_apply(m, st), 3, 3), fx["lda"])
        assert rel(got, fx["feats_lda"]) <= 1e-4


def test_oracle_loglikes_against_kalpy(fx):
    need(fx, "loglikes", "feats_deltas")
    _, am, _ = load_model("mono")
    got = O.gmm_loglikes(O.GmmModel.from_am(am), fx["feats_deltas"].astype(np.float32))
    err = np.abs(got - fx["loglikes"])
    assert np.all(err <= 1e-4 * np.abs(fx["loglikes"]) + 1e-4), float(err.max())


def _kalpy_fst(fx):
    return K.Fst(int(fx["fst_start"][0]), fx["fst_finals"].shape[0], fx["fst_src"], fx["fst_ilabel"], fx["fst_olabel"], fx["fst_dst"],
                 fx["fst_weight"], fx["fst_finals"])


def test_oracle_alignment_on_kalpy_graph_against_kalpy(fx):
    """FasterDecoder restatement on the very graph Kaldi compiled and the very features it computed: isolates the decoder."""
    need(fx, "ali", "fst_src", "feats_deltas")
    tm, am, _ = load_model("mono")
    r = O.align(_kalpy_fst(fx), -tm.scaled_transition_log_probs(1.0, 0.1), O.GmmModel.from_am(am), tm.tid2pdf, fx["feats_deltas"].astype(np.float32),
                fx["feats_deltas"].shape[0], 0.1, 10.0, 40.0)
    assert r["status"] < 2
    assert (r["ali"] == fx["ali"]).mean() >= 0.999
    assert list(r["words"]) == list(fx["words"])
    assert abs(r["like"] - float(fx["likelihood"][0])) <= 1e-4 * abs(float(fx["likelihood"][0]))
    if "per_frame" in fx:
        eq = r["ali"] == fx["ali"]
        assert np.allclose(r["per_frame"][eq], fx["per_frame"][eq], rtol=1e-4, atol=1e-3)


def test_graph_compiler_language_against_kalpy_graph(fx, tmp_path):
    """Our training graph and Kaldi's (determinised / minimised by OpenFst, ours not) must give the same best path and cost for the
    same features: equal alignments, log-likelihood within 1e-4."""
    need(fx, "ali", "feats_deltas")
    from mfa_b200 import engine as E
    ms = mono_sample_setup(tmp_path)
    fst = E.GraphCompiler(ms["tm"], ms["tree"], ms["lex"]).compile([ms["lex"].to_int(bytes(fx["text"]).decode())]).export()[0]
    r = O.align(fst, -ms["tm"].scaled_transition_log_probs(1.0, 0.1), O.GmmModel.from_am(ms["am"]), ms["tm"].tid2pdf,
                fx["feats_deltas"].astype(np.float32), fx["feats_deltas"].shape[0], 0.1, 10.0, 40.0)
    assert (r["ali"] == fx["ali"]).mean() >= 0.999
    assert abs(r["like"] - float(fx["likelihood"][0])) <= 1e-4 * abs(float(fx["likelihood"][0]))


def test_oracle_equal_align_and_acc_stats_against_kalpy(fx):
    tm, am, _ = load_model("mono")
    if "equal_ali" in fx and "fst_src" in fx:
        # kalpy's seed is not visible from the reference tree: only the invariants are compared (length, graph membership, word labels)
        assert fx["equal_ali"].shape[0] == fx["feats_deltas"].shape[0]
        assert set(np.unique(fx["equal_ali"])) <= set(np.unique(fx["fst_ilabel"]))
    need(fx, "acc_trans", "ali", "feats_deltas")
    acc = O.acc_stats(O.GmmModel.from_am(am), tm.tid2pdf, fx["feats_deltas"].astype(np.float32), fx["ali"].astype(np.int32), tm.num_tids)
    assert np.array_equal(np.asarray(acc["trans"])[1:], np.asarray(fx["acc_trans"], np.float64).reshape(-1)[1:])
    if "acc_like" in fx:
        assert abs(acc["like"][0] - float(fx["acc_like"][0])) <= 1e-4 * abs(float(fx["acc_like"][0]))


@pytest.mark.gpu
def test_engine_against_kalpy(fx, tmp_path):
    """The CUDA path through the C ABI against Kaldi's outputs: MFCC -> features -> log-likelihoods -> alignment on Kaldi's own graph."""
    need(fx, "pcm", "mfcc", "feats_deltas", "ali", "fst_src")
    from mfa_b200 import engine as E
    tm, am, _ = load_model("mono")
    eng = E.Engine(0)
    pcm = fx["pcm"]
    so = np.asarray([0, pcm.shape[0]], np.int64)
    raw, fo = eng.mfcc(pcm, so, E.mfcc_opts())
    assert rel(raw, fx["mfcc"]) <= 1e-4
    st = eng.cmvn_stats(raw, fo, np.zeros(1, np.int32), 1)
    feats = eng.features(raw, fo, "deltas", cmvn_stats=st, utt2spk=np.zeros(1, np.int32), n_spk=1)
    assert rel(feats, fx["feats_deltas"]) <= 1e-4
    dm = E.DeviceModel(eng, tm, am)
    if "loglikes" in fx:
        ll = dm.loglikes(fx["feats_deltas"].astype(np.float32))
        assert np.all(np.abs(ll - fx["loglikes"]) <= 1e-4 * np.abs(fx["loglikes"]) + 1e-4)
    graphs = E.Graphs(E.FstBatch.from_fsts([_kalpy_fst(fx)]), tm, 1.0, 0.1)
    res = E.align_feats(eng, dm, graphs, fx["feats_deltas"].astype(np.float32), fo, E.align_opts())
    assert int(res.status[0]) < 2 and (res.ali == fx["ali"]).mean() >= 0.999
    assert abs(float(res.total_like[0]) - float(fx["likelihood"][0])) <= 1e-4 * abs(float(fx["likelihood"][0]))
    assert list(res.words[:int(res.num_words[0])]) == list(fx["words"])
    dm.close(); eng.close()
