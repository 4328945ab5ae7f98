"""-m gpu: the CUDA hot path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): transition-id alignments identical on >= 99.9 % of frames, per-utterance
log-likelihood and MFCC values within 1e-4 relative; integer/index outputs (words, statuses) exact."""
import numpy as np
import pytest

from helpers import build_synth_scenario, gold, load_model, mono_sample_setup, oracle_align_all, oracle_features
from mfa_b200 import engine as E, _lib as L
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


def relmax(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(1e-30, np.abs(b).max()))


def _cat(pcm_list):
    off = np.zeros(len(pcm_list) + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in pcm_list])
    return (np.concatenate(pcm_list) if len(pcm_list) else np.zeros(0, np.int16)).astype(np.int16), off


@pytest.mark.parametrize("kw", [dict(), dict(snip_edges=False), dict(use_energy=True, energy_floor=1.0), dict(use_energy=True, raw_energy=False),
                                dict(num_mel_bins=40, num_coefficients=20, low_frequency=0.0, high_frequency=-200.0, cepstral_lifter=0.0)])
def test_mfcc_parity_ragged_batch(eng, kw):
    g = gold()
    a, b = g["acoustic_corpus_pcm"], g["cold_corpus_pcm"]
    # ragged: full utterance, shorter than one window (0 frames with snip_edges), exactly one window, odd lengths, empty
    pcm_list = [a, a[:399], a[1000:1400], b[:16001], np.zeros(0, np.int16), b[3333:77777], a[5:1234]]
    pcm, off = _cat(pcm_list)
    opts = E.mfcc_opts(**kw)
    okw = dict(snip_edges=int(opts.snip_edges), use_energy=int(opts.use_energy), raw_energy=int(opts.raw_energy), energy_floor=opts.energy_floor,
               num_mel_bins=opts.num_mel_bins, num_ceps=opts.num_ceps, low_freq=opts.low_freq, high_freq=opts.high_freq,
               cepstral_lifter=opts.cepstral_lifter)
    out, fo = eng.mfcc(pcm, off, opts)
    ref = [O.mfcc(p, O.mfcc_opts(**okw)) for p in pcm_list]
    assert [int(fo[i + 1] - fo[i]) for i in range(len(pcm_list))] == [r.shape[0] for r in ref]
    refc = np.concatenate(ref)
    assert relmax(out, refc) < 1e-4           # tolerance stated by north_star: MFCC within 1e-4 relative
    # also against the independent torchaudio port for the default options
    if not kw:
        assert relmax(out[: 2670], g["ta_mfcc_snip"]) < 1e-4


def test_mfcc_device_buffers_and_linearity(eng):
    import torch
    g = gold()
    pcm = g["acoustic_corpus_pcm"][:160000]
    off = np.asarray([0, 50000, 160000], np.int64)
    host, fo = eng.mfcc(pcm, off, E.mfcc_opts())
    dev, fo2 = eng.mfcc(torch.from_numpy(pcm.copy()).cuda(), off, E.mfcc_opts())
    eng.sync()
    assert np.array_equal(fo, fo2) and np.array_equal(host, dev.cpu().numpy())
    # size-independent property: scaling the waveform by 2 shifts c0 by sqrt(23)*2*log(2)*lifter0 and leaves c1.. unchanged
    half = (pcm // 2 * 2).astype(np.int16)
    o1, _ = eng.mfcc((half // 2).astype(np.int16), off, E.mfcc_opts())
    o2, _ = eng.mfcc(half, off, E.mfcc_opts())
    assert np.allclose(o2[:, 1:], o1[:, 1:], atol=2e-3)
    assert np.allclose(o2[:, 0] - o1[:, 0], np.sqrt(23.0) * 2 * np.log(2.0), atol=2e-3)


def test_cmvn_and_feature_pipeline_parity(eng):
    g = gold()
    a, b = g["acoustic_corpus_pcm"], g["cold_corpus_pcm"]
    pcm_list = [a[:80000], b[:64000], a[80000:200000], b[64000:], a[200000:200650]]
    utt2spk = np.asarray([0, 1, 0, 1, 2], np.int32)
    rng = np.random.default_rng(0)
    lda = g["g2p_lda"]
    fm = (np.eye(40, 41)[None] + 0.05 * rng.standard_normal((3, 40, 41))).astype(np.float32)
    pcm, off = _cat(pcm_list)
    raw, fo = eng.mfcc(pcm, off, E.mfcc_opts())
    stats = eng.cmvn_stats(raw, fo, utt2spk, 3)
    raw_o, stats_o, feats_d = oracle_features(pcm_list, utt2spk, 3, "deltas")
    assert relmax(stats, stats_o) < 1e-5      # f64 sums of f32 values that themselves agree to ~1e-6
    # feature kernels are checked on the ORACLE's MFCCs and stats so that only the op under test differs
    rawc = np.concatenate(raw_o)
    out_d = eng.features(rawc, fo, "deltas", cmvn_stats=stats_o, utt2spk=utt2spk, n_spk=3)
    assert relmax(out_d, np.concatenate(feats_d)) < 1e-6
    _, _, feats_l = oracle_features(pcm_list, utt2spk, 3, "lda", lda=lda, fmllr=fm)
    out_l = eng.features(rawc, fo, "lda", lda=lda, fmllr=fm, cmvn_stats=stats_o, utt2spk=utt2spk, n_spk=3)
    assert relmax(out_l, np.concatenate(feats_l)) < 1e-5
    lda92 = np.concatenate([lda, rng.standard_normal((40, 1)).astype(np.float32)], 1)
    _, _, feats_a = oracle_features(pcm_list, utt2spk, 3, "lda", lda=lda92)
    out_a = eng.features(rawc, fo, "lda", lda=lda92, cmvn_stats=stats_o, utt2spk=utt2spk, n_spk=3)
    assert relmax(out_a, np.concatenate(feats_a)) < 1e-5


@pytest.mark.parametrize("impl", [1, 0, 96])
@pytest.mark.parametrize("which", ["mono", "g2p"])
def test_gmm_loglikes_parity(eng, impl, which, request):
    if impl == 96:   # the K = 96 geometry of the tensor-core kernel (gconst as fp16 columns, two-stage ring) instead of K = 80
        eng.set_option("tc_k96", 1)   # read when the model's operand images are built
        request.addfinalizer(lambda: eng.set_option("tc_k96", 0))
        impl = 0
    tm, am, _ = load_model(which)
    rng = np.random.default_rng(2)
    means = am.means()
    # frames drawn around the model's own Gaussians (+ a few outliers), ragged count to exercise tile tails
    pick = rng.integers(0, am.NumGauss(), 777)
    x = (means[pick] + rng.standard_normal((777, am.dim)) * np.sqrt(am.variances()[pick])).astype(np.float32)
    x[::97] *= 3.0
    dm = E.DeviceModel(eng, tm, am)
    ll = dm.loglikes(x, impl=impl)
    ref = O.gmm_loglikes(O.GmmModel.from_am(am), x)
    # per value: within 1e-4 relative (north_star's tolerance for log-likelihoods); typical error is ~1e-6 relative
    err = np.abs(ll - ref)
    assert np.all(err <= 1e-4 * np.abs(ref) + 1e-4), (err.max(), np.abs(ref).mean())
    assert err.mean() <= 2e-6 * np.abs(ref).mean() + (0 if impl == 1 else 2e-4)
    dm.close()


def test_boost_silence_shifts_only_those_pdfs(eng):
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(3)
    x = rng.standard_normal((64, am.dim)).astype(np.float32)
    dm = E.DeviceModel(eng, tm, am)
    base = dm.loglikes(x, impl=1)
    dm.boost_pdfs(1.5, [0, 7])
    b = dm.loglikes(x, impl=1)
    d = b - base
    assert np.allclose(d[:, [0, 7]], np.log(1.5), atol=1e-4) and np.abs(np.delete(d, [0, 7], axis=1)).max() == 0.0
    dm.close()


def _gpu_align_from_oracle_loglikes(eng, sc, fsts_batch, beam, retry_beam):
    tm, am = sc["tm"], sc["am"]
    dm = E.DeviceModel(eng, tm, am)
    graphs = E.Graphs(fsts_batch, tm, 1.0, 0.1)
    g = O.GmmModel.from_am(am)
    ll = np.concatenate([O.gmm_loglikes(g, f) for f in sc["feats"]])
    res = E.align_loglikes(eng, dm, graphs, ll, sc["frame_off"], E.align_opts(0.1, beam, retry_beam))
    dm.close()
    return res


@pytest.mark.parametrize("triphone,beam,retry", [(False, 10.0, 40.0), (True, 10.0, 40.0), (True, 200.0, 0.0), (False, 2.0, 8.0)])
def test_viterbi_parity_with_oracle(eng, triphone, beam, retry):
    sc = build_synth_scenario(seconds=60.0, seed=21, triphone=triphone, n_phones=10, n_words=50, target_pdfs=90, gauss_per_pdf=2)
    batch = E.GraphCompiler(sc["tm"], sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts)
    fsts = batch.export()
    ref = oracle_align_all(sc, fsts, beam, retry)
    res = _gpu_align_from_oracle_loglikes(eng, sc, batch, beam, retry)
    same = total = 0
    for u, r in enumerate(ref):
        got = res.utterance(u)
        assert got["status"] == r["status"], (u, got["status"], r["status"])
        if r["status"] >= 2:
            continue
        same += int((got["ali"] == r["ali"]).sum()); total += len(r["ali"])
        assert list(got["words"]) == list(r["words"])
        assert abs(got["like"] - r["like"]) <= 1e-4 * abs(r["like"])
        assert np.allclose(got["per_frame"], r["per_frame"], rtol=1e-4, atol=1e-3)
    assert total > 0 and same / total >= 0.999, same / total


def _align_env(eng, sc, batch, beam, retry, **opts):
    """One alignment run under the given engine options (mfa_engine_set_option), restored afterwards."""
    with eng.options(**opts):
        return _gpu_align_from_oracle_loglikes(eng, sc, batch, beam, retry)


@pytest.mark.parametrize("triphone,beam,retry", [(True, 10.0, 40.0), (False, 3.0, 60.0), (True, 200.0, 0.0)])
def test_band_kernel_equals_sparse_kernel(eng, triphone, beam, retry):
    """K3's two kernels implement one recursion: the band kernel (primary), the sparse kernel (epsilon graphs, fallback) and the
    band kernel with a 1- or 2-group band (most utterances overflow and are re-run by the sparse kernel) must agree exactly."""
    sc = build_synth_scenario(seconds=80.0, seed=5, triphone=triphone, n_phones=12, n_words=60, target_pdfs=100, gauss_per_pdf=2)
    batch = E.GraphCompiler(sc["tm"], sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts)
    f0 = eng.band_fallbacks
    band = _align_env(eng, sc, batch, beam, retry)
    f1 = eng.band_fallbacks
    sparse = _align_env(eng, sc, batch, beam, retry, vit_band=0)
    assert eng.band_fallbacks == f1
    narrow = _align_env(eng, sc, batch, beam, retry, vit_maxgroups=1 if beam < 100 else 2)    # overflow -> 32-group band kernel (first fallback level)
    f2 = eng.band_fallbacks
    assert f2 > f1                                      # the fallback path really ran
    narrow_sparse = _align_env(eng, sc, batch, beam, retry, vit_maxgroups=1 if beam < 100 else 2, vit_wide=0)   # overflow -> sparse kernel directly
    assert eng.band_fallbacks - f2 == f2 - f1
    f3 = eng.band_fallbacks
    narrow_seq = _align_env(eng, sc, batch, beam, retry, vit_maxgroups=1 if beam < 100 else 2, vit_wide_poll=0)   # wide level after the join instead of polling
    assert eng.band_fallbacks - f3 == f2 - f1
    narrow_few = _align_env(eng, sc, batch, beam, retry, vit_maxgroups=1 if beam < 100 else 2, vit_wide_poll=1, vit_wide_ctas=1)   # one poller takes the whole list
    if beam <= 10.0:
        assert f1 == f0                                   # ... and the default band is wide enough for ordinary beams
    assert np.isin(sparse.status, (0, 1)).sum() > 0
    smem4 = _align_env(eng, sc, batch, beam, retry, vit_graph_smem=1, vit_nw2_kb=0)   # graph in shared memory, 4 warps
    l1w4 = _align_env(eng, sc, batch, beam, retry, vit_nw2_kb=0)                           # graph through L1, 4 warps
    for other in (band, narrow, narrow_sparse, narrow_seq, narrow_few, smem4, l1w4):
        assert np.array_equal(other.status, sparse.status) and np.array_equal(other.num_words, sparse.num_words)
        assert np.array_equal(other.ali, sparse.ali) and np.array_equal(other.words, sparse.words)
        assert np.array_equal(other.total_like, sparse.total_like) and np.array_equal(other.per_frame, sparse.per_frame)


def test_viterbi_edge_cases(eng):
    sc = build_synth_scenario(seconds=12.0, seed=4, n_phones=6, n_words=20, gauss_per_pdf=2)
    tm, am = sc["tm"], sc["am"]
    from mfa_b200.kaldi_io import Fst
    fsts = E.GraphCompiler(tm, sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts).export()
    n = len(fsts)
    empty = Fst(-1, 0, *(np.zeros(0, np.int32) for _ in range(4)), np.zeros(0, np.float32), np.zeros(0, np.float32))
    fsts2 = [fsts[0], empty, fsts[1], fsts[2 % n]]
    g = O.GmmModel.from_am(am)
    lls = [O.gmm_loglikes(g, sc["feats"][0]), O.gmm_loglikes(g, sc["feats"][1]), O.gmm_loglikes(g, sc["feats"][1])[:0],
           O.gmm_loglikes(g, sc["feats"][2 % n])[:4]]
    fo = np.zeros(5, np.int64)
    fo[1:] = np.cumsum([x.shape[0] for x in lls])
    dm = E.DeviceModel(eng, tm, am)
    graphs = E.Graphs(E.FstBatch.from_fsts(fsts2), tm)
    res = E.align_loglikes(eng, dm, graphs, np.concatenate(lls), fo, E.align_opts())
    assert int(res.status[0]) in (0, 1) and [int(s) for s in res.status[1:]] == [3, 4, 2]
    dm.close()


def test_epsilon_arcs_are_supported(eng):
    """Graphs read from a Kaldi fsts.ark may keep input-epsilon arcs; splice one into every graph and compare with the oracle."""
    sc = build_synth_scenario(seconds=20.0, seed=9, n_phones=6, n_words=20, gauss_per_pdf=2)
    tm = sc["tm"]
    from mfa_b200.kaldi_io import Fst
    fsts = E.GraphCompiler(tm, sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts).export()
    mod = []
    for f in fsts:
        # new start state with an epsilon arc (carrying a weight) into the old start, and an epsilon arc into a new final state
        S = f.num_states
        fin_states = np.where(np.isfinite(f.finals))[0]
        src = np.concatenate([f.arc_src, [S], fin_states]).astype(np.int32)
        dst = np.concatenate([f.arc_dst, [f.start], np.full(len(fin_states), S + 1)]).astype(np.int32)
        il = np.concatenate([f.arc_ilabel, [0], np.zeros(len(fin_states))]).astype(np.int32)
        ol = np.concatenate([f.arc_olabel, [0], np.zeros(len(fin_states))]).astype(np.int32)
        w = np.concatenate([f.arc_weight, [0.25], f.finals[fin_states]]).astype(np.float32)
        finals = np.full(S + 2, np.inf, np.float32)
        finals[S + 1] = 0.0
        mod.append(Fst(S, S + 2, src, il, ol, dst, w, finals))
    ref = oracle_align_all(sc, mod, 10.0, 40.0)
    res = _gpu_align_from_oracle_loglikes(eng, sc, E.FstBatch.from_fsts(mod), 10.0, 40.0)
    base = oracle_align_all(sc, fsts, 10.0, 40.0)
    for u, r in enumerate(ref):
        got = res.utterance(u)
        assert got["status"] == r["status"] and (got["ali"] == r["ali"]).mean() >= 0.999
        assert abs(got["like"] - r["like"]) <= 1e-4 * abs(r["like"])
        assert abs(r["like"] - (base[u]["like"] - 0.25 / 0.1)) <= 1e-4 * abs(r["like"])


def test_fused_pipeline_config1_sample(eng, tmp_path):
    """Config 1: the reference's sample utterance, its fixture monophone model and dictionary, PCM in -> alignment out."""
    ms = mono_sample_setup(tmp_path)
    tm, am, lex = ms["tm"], ms["am"], ms["lex"]
    ids = lex.to_int(ms["text"])
    batch = E.GraphCompiler(tm, ms["tree"], lex).compile([ids, ids[:20]])
    fsts = batch.export()
    pcm_list = [ms["pcm"], ms["pcm"][:140000]]
    pcm, off = _cat(pcm_list)
    u2s = np.asarray([0, 0], np.int32)
    dm = E.DeviceModel(eng, tm, am)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    for impl in (1, 0, 2):   # fp32 CUDA-core (all pdfs), tcgen05 on per-utterance pdf subsets, tcgen05 on all pdfs
        res = E.align_pcm(eng, dm, graphs, pcm, off, u2s, 1, E.mfcc_opts(), "deltas", gmm_impl=impl)
        _, _, feats = oracle_features(pcm_list, u2s, 1, "deltas")
        g = O.GmmModel.from_am(am)
        tc = -tm.scaled_transition_log_probs(1.0, 0.1)
        for u in range(2):
            r = O.align(fsts[u], tc, g, tm.tid2pdf, feats[u], feats[u].shape[0])
            got = res.utterance(u)
            assert got["status"] == r["status"]
            if r["status"] < 2:
                assert (got["ali"] == r["ali"]).mean() >= 0.999
                assert abs(got["like"] - r["like"]) <= 1e-4 * abs(r["like"])
                assert list(got["words"]) == list(r["words"])
    dm.close()


def test_fused_pipeline_triphone_lda_fmllr_chunked(eng):
    """Config 2/3 shape in miniature: splice+LDA(+fMLLR) features, triphone tree, several speakers, forced small workspace
    so the chunk loop runs more than once; device-resident PCM."""
    import torch
    sc = build_synth_scenario(seconds=90.0, seed=33, triphone=True, n_phones=10, n_words=60, target_pdfs=150, gauss_per_pdf=4, use_lda=True, n_spk=4)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    fsts = batch.export()
    ref = oracle_align_all(sc, fsts, 10.0, 40.0)
    dm = E.DeviceModel(eng, tm, am)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    res = E.align_pcm(eng, dm, graphs, torch.from_numpy(c.pcm).cuda(), c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"],
                      workspace_bytes=6 << 20)
    eng.sync()
    ali = res.ali.cpu().numpy(); st = res.status.cpu().numpy(); tl = res.total_like.cpu().numpy()
    same = total = 0
    for u, r in enumerate(ref):
        assert int(st[u]) == r["status"]
        if r["status"] >= 2:
            continue
        a = ali[res.frame_off[u]:res.frame_off[u + 1]]
        same += int((a == r["ali"]).sum()); total += len(a)
        assert abs(float(tl[u]) - r["like"]) <= 1e-4 * abs(r["like"])
    assert same / total >= 0.999, same / total
    # host-buffer path cut into segments at speaker boundaries (uploads overlap work on earlier segments): same outputs as the
    # single-segment run; the corpus generator keeps each speaker's utterances contiguous, like MFA's job ordering
    with eng.options(pipeline_split=1):
        res_split = E.align_pcm(eng, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"])
    with eng.options(pipeline_split=0):
        res_whole = E.align_pcm(eng, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"])
    # streamed: K1 / CMVN / features / K2 per segment as its PCM arrives, one Viterbi launch at the end (the default for big batches)
    with eng.options(pipeline_split=2):
        res_stream = E.align_pcm(eng, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"])
    assert len(set(c.utt2spk.tolist())) > 1
    for r_ in (res_split, res_whole, res_stream):
        assert np.array_equal(r_.ali, ali) and np.array_equal(r_.status, st) and np.array_equal(r_.total_like, tl)
    # per-speaker fMLLR path: identity transforms must reproduce the result exactly
    fm = np.tile(np.eye(40, 41, dtype=np.float32)[None], (c.n_spk, 1, 1))
    res2 = E.align_pcm(eng, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"], fmllr=fm)
    assert np.array_equal(res2.ali, ali) and np.array_equal(res2.status, st)
    dm.close()


def test_acc_stats_parity(eng):
    sc = build_synth_scenario(seconds=30.0, seed=8, n_phones=8, n_words=30, gauss_per_pdf=3)
    tm, am = sc["tm"], sc["am"]
    fsts = E.GraphCompiler(tm, sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts).export()
    ref = oracle_align_all(sc, fsts, 10.0, 40.0)
    g = O.GmmModel.from_am(am)
    accs = None
    dm = E.DeviceModel(eng, tm, am)
    dm.acc_zero()
    for u, r in enumerate(ref):
        if r["status"] >= 2:
            continue
        accs = O.acc_stats(g, tm.tid2pdf, sc["feats"][u], r["ali"], tm.num_tids, accs)
        dm.acc_stats(sc["feats"][u], r["ali"])
    got = dm.acc_read()
    assert got["frames"] == accs["frames"] and np.array_equal(got["trans"], accs["trans"])     # integer counts: exact
    assert abs(got["like"] - accs["like"][0]) <= 1e-5 * abs(accs["like"][0])
    assert np.allclose(got["occ"], accs["occ"], rtol=1e-4, atol=1e-5)
    assert np.allclose(got["mean"], accs["mean"], rtol=1e-4, atol=1e-3) and np.allclose(got["var"], accs["var"], rtol=1e-4, atol=1e-2)
    # linearity: accumulating the same frames twice doubles everything
    for u, r in enumerate(ref):
        if r["status"] < 2:
            dm.acc_stats(sc["feats"][u], r["ali"])
    got2 = dm.acc_read()
    assert np.allclose(got2["occ"], 2 * got["occ"], rtol=1e-9) and got2["frames"] == 2 * got["frames"]
    dm.close()


def test_acc_stats_segmented_equals_atomic_kernel_and_handles_ragged_input(eng, request):
    """K4's two kernels (counting sort by pdf + register accumulation per item | one red per frame) on one concatenated batch:
    integer fields identical, f64 sums equal to rounding; frames with tid 0 / out-of-range are skipped; a pdf with many
    components and a pdf with thousands of frames (several items, one CTA reusing its staged Gaussians) are both present."""
    rng = np.random.default_rng(3)
    D, P = 39, 50
    comps = rng.integers(1, 6, size=P); comps[7] = 70; comps[0] = 1
    off = np.zeros(P + 1, np.int32); off[1:] = np.cumsum(comps)
    G = int(off[-1])
    means, var = rng.normal(size=(G, D)), np.exp(rng.normal(scale=0.3, size=(G, D)))
    w = np.concatenate([rng.dirichlet(np.ones(c)) for c in comps])
    from mfa_b200 import kaldi_io as K
    am = K.AmDiagGmm(D, off, w.astype(np.float32), (means / var).astype(np.float32), (1.0 / var).astype(np.float32))
    tm, _, _ = load_model("mono")
    # a fake tid -> pdf map over the fixture's transition model: only tid2pdf matters to K4
    tid2pdf = np.concatenate([[0], rng.integers(0, P, size=tm.num_tids)]).astype(np.int32)
    tid2pdf[1:40] = 3                                   # many tids share a heavy pdf
    tm.tid2pdf = tid2pdf
    T = 20000
    ali = rng.integers(1, tm.num_tids + 1, size=T).astype(np.int32)
    heavy = rng.random(T) < 0.3
    ali[heavy] = rng.integers(1, 40, size=int(heavy.sum()))
    ali[::97] = 0
    ali[5::1013] = tm.num_tids + 5
    feats = rng.normal(size=(T, D)).astype(np.float32)
    dm = E.DeviceModel(eng, tm, am)
    res = {}
    for impl in ("atomic", "segmented"):
        eng.set_option("acc_impl", 1 if impl == "atomic" else 0)
        request.addfinalizer(lambda: eng.set_option("acc_impl", 0))
        dm.acc_zero()
        dm.acc_stats(feats, ali)
        dm.acc_stats(feats[:0], ali[:0])               # empty call is a no-op
        res[impl] = dm.acc_read()
    a, s = res["atomic"], res["segmented"]
    valid = (ali > 0) & (ali <= tm.num_tids)
    assert s["frames"] == a["frames"] == int(valid.sum()) and np.array_equal(s["trans"], a["trans"])
    assert np.array_equal(s["trans"][1:], np.bincount(ali[valid], minlength=tm.num_tids + 1)[1:])
    assert abs(s["like"] - a["like"]) <= 1e-9 * abs(a["like"])
    # fp32 dot products are summed in a different lane order: posteriors differ by ~1e-7 per frame
    assert np.allclose(s["occ"], a["occ"], rtol=1e-5, atol=1e-4)
    assert np.allclose(s["mean"], a["mean"], rtol=1e-5, atol=1e-3) and np.allclose(s["var"], a["var"], rtol=1e-5, atol=1e-3)
    g = O.GmmModel.from_am(am)
    ref = O.acc_stats(g, tid2pdf, feats[valid], ali[valid], tm.num_tids)
    assert np.allclose(s["occ"], ref["occ"], rtol=1e-4, atol=1e-5) and abs(s["like"] - ref["like"][0]) <= 1e-5 * abs(ref["like"][0])
    assert np.allclose(s["mean"], ref["mean"], rtol=1e-4, atol=1e-3) and np.allclose(s["var"], ref["var"], rtol=1e-4, atol=1e-2)
    assert abs(s["occ"].sum() - valid.sum()) < 1e-3 * valid.sum()       # posteriors of a frame sum to one
    dm.close()


@pytest.mark.parametrize("two_models,use_lda", [(False, False), (True, True)])
def test_fmllr_stats_and_transforms_parity(eng, two_models, use_lda):
    """K5 against the oracle's FmllrDiagGmmAccs restatement, per speaker, with silence weighting; then the transforms."""
    from mfa_b200 import fmllr as F
    sc = build_synth_scenario(seconds=60.0, seed=21, n_phones=8, n_words=30, gauss_per_pdf=3, n_spk=3, use_lda=use_lda)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    fsts = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts).export()
    ref = oracle_align_all(sc, fsts, 10.0, 40.0)
    D = am.dim
    am_post = am
    if two_models:   # an "alignment model": same layout, perturbed parameters
        rng = np.random.default_rng(4)
        from mfa_b200 import kaldi_io as K
        am_post = K.AmDiagGmm(D, am.offsets, am.weights, am.means_invvars * (1 + 0.05 * rng.normal(size=am.means_invvars.shape)).astype(np.float32),
                              am.inv_vars * np.exp(0.05 * rng.normal(size=am.inv_vars.shape)).astype(np.float32))
    sil_phone = c.lexicon.phone_table["sil"]
    tw = np.where(tm.tid2phone == sil_phone, np.float32(0.0), np.float32(1.0)).astype(np.float32)
    tw[0] = 0.0
    g, gp = O.GmmModel.from_am(am), O.GmmModel.from_am(am_post)
    fo = sc["frame_off"]
    ali = np.zeros(int(fo[-1]), np.int32)
    stats_ref = np.zeros((c.n_spk, O.fmllr_stats_size(D)))
    for u, r in enumerate(ref):
        if r["status"] >= 2:
            continue
        ali[fo[u]:fo[u + 1]] = r["ali"]
        O.fmllr_acc(gp, g, tm.tid2pdf, tw, sc["feats"][u], r["ali"], stats_ref[c.utt2spk[u]])
    dm = E.DeviceModel(eng, tm, am)
    dmp = E.DeviceModel(eng, tm, am_post) if two_models else None
    feats = np.concatenate(sc["feats"])
    got = dm.fmllr_acc(feats, ali, fo, c.utt2spk, c.n_spk, tid_weight=tw, post_model=dmp)
    assert got.shape == stats_ref.shape
    n_sil = int((tw[ali] == 0).sum())
    assert 0 < n_sil < ali.shape[0]
    assert np.allclose(got[:, 0], stats_ref[:, 0], rtol=1e-6)
    assert abs(got[:, 0].sum() - (ali.shape[0] - n_sil)) < 1e-3 * ali.shape[0]          # beta = number of weighted frames
    scale = np.abs(stats_ref).max(axis=1, keepdims=True)
    assert (np.abs(got - stats_ref) / scale).max() < 1e-6
    assert relmax(got[:, 1:1 + D * (D + 1)], stats_ref[:, 1:1 + D * (D + 1)]) < 1e-5
    # device buffers in -> device statistics out, same numbers
    import torch
    got_d = dm.fmllr_acc(torch.from_numpy(feats).cuda(), torch.from_numpy(ali).cuda(), fo, c.utt2spk, c.n_spk, tid_weight=tw, post_model=dmp)
    eng.sync()
    assert np.allclose(got_d.cpu().numpy(), got, rtol=1e-9, atol=1e-9 * scale.max())
    # transforms: product (numpy on the GPU statistics) vs oracle update on the oracle statistics
    W, impr, count = F.compute_transforms(got, D, min_count=100.0)
    Wd, impr_d, count_d = eng.fmllr_update(got, D, 40, 100.0)                      # the CUDA update kernel (host buffers)
    Wt, impr_t, _ = eng.fmllr_update(got_d, D, 40, 100.0)                           # ... and on the device-resident statistics
    eng.sync()
    # (the two statistics blocks come from two accumulation runs: f64 atomics reorder sums, so the transforms agree to rounding only)
    assert np.allclose(count_d, got[:, 0]) and np.abs(Wt.cpu().numpy() - Wd).max() < 1e-4 and np.allclose(impr_t, impr_d, rtol=1e-4, atol=1e-2)
    for s in range(c.n_spk):
        Wo, io = O.fmllr_update(stats_ref[s], D, min_count=100.0)
        assert np.abs(W[s] - Wo).max() < 1e-3 and abs(impr[s] - io) <= 1e-3 * max(1.0, abs(io))
        assert np.abs(Wd[s] - Wo).max() < 1e-3 and abs(impr_d[s] - io) <= 1e-3 * max(1.0, abs(io))
        assert impr[s] >= 0.0 and impr_d[s] >= 0.0
    # below min_count the unit transform is kept
    Wu, impr_u, _ = eng.fmllr_update(got, D, 40, 1e9)
    assert (Wu == np.eye(D, D + 1, dtype=np.float32)).all() and not impr_u.any()
    # edge: a speaker without frames and an utterance without alignment contribute nothing
    got2 = dm.fmllr_acc(feats, np.zeros_like(ali), fo, c.utt2spk, c.n_spk + 1, tid_weight=tw, post_model=dmp)
    assert not got2.any()
    dm.close()
    if dmp:
        dmp.close()


def test_acc_stats_large_many_pdfs_against_numpy(eng, request):
    """K4 at a size where every CTA walks several items and pdfs (4 000 pdfs x 10 components, 600 k frames): both kernels against a
    float64 numpy evaluation of the same posteriors (total log-likelihood, occupancies, first-order sums)."""
    from mfa_b200 import kaldi_io as K
    rng = np.random.default_rng(12)
    D, P, M, T = 40, 4000, 10, 600_000
    off = (np.arange(P + 1) * M).astype(np.int32)
    G = P * M
    means, var = rng.normal(size=(G, D)), np.exp(rng.normal(scale=0.2, size=(G, D)))
    w = rng.dirichlet(np.ones(M), size=P).ravel()
    am = K.AmDiagGmm(D, off, w.astype(np.float32), (means / var).astype(np.float32), (1.0 / var).astype(np.float32))
    tm, _, _ = load_model("mono")
    nt = tm.num_tids
    tid2pdf = np.concatenate([[0], rng.integers(0, P, size=nt)]).astype(np.int32)
    tm.tid2pdf = tid2pdf
    # skewed usage: a handful of pdfs take a quarter of the frames (silence-like), the rest spread thin
    ali = rng.integers(1, nt + 1, size=T).astype(np.int32)
    hot = rng.random(T) < 0.25
    ali[hot] = rng.integers(1, 6, size=int(hot.sum()))
    pdf = tid2pdf[ali]
    comp = rng.integers(0, M, size=T)
    feats = (means[pdf * M + comp] + rng.normal(size=(T, D)) * np.sqrt(var[pdf * M + comp])).astype(np.float32)
    # float64 reference, frames grouped by pdf
    x = feats.astype(np.float64)
    miv, iv, gc = am.means_invvars.astype(np.float64), am.inv_vars.astype(np.float64), am.gconsts.astype(np.float64)
    ll = gc[(pdf * M)[:, None] + np.arange(M)[None]]
    idx = (pdf * M)[:, None] + np.arange(M)[None]
    ll = ll + np.einsum("tmd,td->tm", miv[idx], x) - 0.5 * np.einsum("tmd,td->tm", iv[idx], x * x)
    mx = ll.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(ll - mx).sum(1))
    post = np.exp(ll - lse[:, None])
    occ_ref = np.bincount(idx.ravel(), weights=post.ravel(), minlength=G)
    mean0_ref = np.bincount(idx.ravel(), weights=(post * x[:, :1]).ravel(), minlength=G)
    dm = E.DeviceModel(eng, tm, am)
    for impl in ("segmented", "atomic"):
        eng.set_option("acc_impl", 1 if impl == "atomic" else 0)
        request.addfinalizer(lambda: eng.set_option("acc_impl", 0))
        dm.acc_zero()
        dm.acc_stats(feats, ali)
        got = dm.acc_read()
        assert got["frames"] == T
        assert abs(got["like"] - lse.sum()) <= 2e-6 * abs(lse.sum()), impl
        assert np.allclose(got["occ"], occ_ref, rtol=1e-3, atol=2e-3), impl
        assert np.allclose(got["mean"][:, 0], mean0_ref, rtol=1e-3, atol=5e-3), impl
        assert np.array_equal(got["trans"][1:], np.bincount(ali, minlength=nt + 1)[1:])
    dm.close()


@pytest.mark.parametrize("k96", [False, True])
def test_gmm_loglikes_odd_gaussian_count_and_ragged_pdfs(eng, request, k96):
    """Dense K2 on a model whose Gaussian count is odd and whose pdfs have 1..17 components (tiles with ragged pdf boundaries, a
    trailing partial tile): both operand geometries against the oracle.  (A Gaussian count that is not a multiple of 4 once put the
    per-tile gconst array of the K = 80 geometry on a misaligned address.)"""
    from mfa_b200 import kaldi_io as K
    if k96:
        eng.set_option("tc_k96", 1)
        request.addfinalizer(lambda: eng.set_option("tc_k96", 0))
    rng = np.random.default_rng(77)
    D, P = 39, 53
    comps = rng.integers(1, 18, size=P)
    if comps.sum() % 2 == 0:
        comps[0] += 1
    off = np.zeros(P + 1, np.int32); off[1:] = np.cumsum(comps)
    G = int(off[-1])
    assert G % 4 != 0 or G % 2 == 1
    means, var = rng.normal(size=(G, D)) * 2.0, np.exp(rng.normal(scale=0.4, size=(G, D)))
    w = np.concatenate([rng.dirichlet(np.ones(c)) for c in comps])
    am = K.AmDiagGmm(D, off, w.astype(np.float32), (means / var).astype(np.float32), (1.0 / var).astype(np.float32))
    tm, _, _ = load_model("mono")
    tm.tid2pdf = (np.maximum(tm.tid2pdf, 0) % P).astype(np.int32)   # only the transition-id -> pdf range matters to the device model
    x = (means[rng.integers(0, G, 300)] + rng.standard_normal((300, D)) * 1.5).astype(np.float32)
    dm = E.DeviceModel(eng, tm, am)
    for T in (300, 1, 129):
        ll = dm.loglikes(x[:T], impl=0)
        ref = O.gmm_loglikes(O.GmmModel.from_am(am), x[:T])
        assert np.all(np.abs(ll - ref) <= 1e-4 * np.abs(ref) + 1e-4), (T, np.abs(ll - ref).max())
    dm.close()


@pytest.mark.parametrize("triphone,use_lda", [(False, False), (True, True)])
def test_align_feats_equals_dense_loglikes_path(eng, triphone, use_lda):
    """mfa_align_feats (final features in, per-utterance pdf subsets scored on the device) against mfa_gmm_loglikes + mfa_align (the dense
    frames x all-pdfs matrix through host memory) and the oracle; host and device inputs; a tiny workspace forces several chunks."""
    import torch
    sc = build_synth_scenario(seconds=40.0, seed=51, triphone=triphone, n_phones=8, n_words=30, gauss_per_pdf=3, use_lda=use_lda)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    dm = E.DeviceModel(eng, tm, am)
    feats = np.concatenate(sc["feats"]).astype(np.float32)
    fo = sc["frame_off"]
    opts = E.align_opts()
    a = E.align_feats(eng, dm, graphs, feats, fo, opts)
    b = E.align_loglikes(eng, dm, graphs, dm.loglikes(feats, impl=0), fo, opts)
    assert np.array_equal(a.status, b.status) and (a.ali == b.ali).mean() >= 0.999
    assert np.allclose(a.total_like, b.total_like, rtol=1e-4)
    ref = oracle_align_all(sc, batch.export())
    same = total = 0
    for u, r in enumerate(ref):
        got = a.utterance(u)
        assert got["status"] == r["status"]
        if r["status"] < 2:
            same += int((got["ali"] == r["ali"]).sum()); total += len(r["ali"])
            assert list(got["words"]) == list(r["words"]) and abs(got["like"] - r["like"]) <= 1e-4 * abs(r["like"])
    assert same / total >= 0.999
    small = E.align_feats(eng, dm, graphs, feats, fo, opts, workspace_bytes=2 << 20)       # several chunks
    assert np.array_equal(small.ali, a.ali) and np.array_equal(small.status, a.status)
    d = E.align_feats(eng, dm, graphs, torch.from_numpy(feats).cuda(), fo, opts)            # device buffers
    eng.sync()
    assert np.array_equal(d.ali.cpu().numpy()[: a.ali.shape[0]], a.ali)
    graphs.close(); batch.close(); dm.close()


def test_new_entry_points_edge_cases(eng):
    """Empty / degenerate inputs through the entry points added for rows a11, N2 and the features-in aligner: nothing crashes, statuses and
    shapes are what the callers rely on."""
    sc = build_synth_scenario(seconds=12.0, seed=61, n_phones=6, n_words=20, gauss_per_pdf=2)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    dm = E.DeviceModel(eng, tm, am)
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    n = c.n_utts
    D = am.dim
    # features-in aligner with utterances that have no frames at all
    fo0 = np.zeros(n + 1, np.int64)
    r = E.align_feats(eng, dm, graphs, np.zeros((0, D), np.float32), fo0, E.align_opts())
    assert (r.status == 4).all() and r.ali.shape[0] == 0                       # MFA_ALIGN_ZERO_FRAMES
    # ... and with one real utterance among empty ones
    fo1 = np.zeros(n + 1, np.int64); fo1[1:] = sc["feats"][0].shape[0]
    r = E.align_feats(eng, dm, graphs, sc["feats"][0], fo1, E.align_opts())
    ref = oracle_align_all(sc, batch.export()[:1])[0]
    assert r.status[0] == ref["status"] and (r.status[1:] == 4).all() and (r.utterance(0)["ali"] == ref["ali"]).mean() >= 0.999
    # fMLLR statistics: zero frames, zero speakers
    s = dm.fmllr_acc(np.zeros((0, D), np.float32), np.zeros(0, np.int32), fo0, c.utt2spk, c.n_spk)
    assert s.shape == (c.n_spk, dm.fmllr_stats_size()) and not s.any()
    W, impr, cnt = eng.fmllr_update(s, D)
    assert (W == np.eye(D, D + 1, dtype=np.float32)).all() and not impr.any() and not cnt.any()
    W0, i0, c0 = eng.fmllr_update(np.zeros((0, dm.fmllr_stats_size())), D)
    assert W0.shape == (0, D, D + 1) and i0.shape == (0,)
    # statistics with a mismatching posterior model are refused
    from mfa_b200 import kaldi_io as K, _lib as L
    am2 = K.AmDiagGmm(D, np.arange(am.NumPdfs() + 1, dtype=np.int32), np.ones(am.NumPdfs(), np.float32),
                      np.zeros((am.NumPdfs(), D), np.float32), np.ones((am.NumPdfs(), D), np.float32))
    dm2 = E.DeviceModel(eng, tm, am2)
    if am2.NumGauss() != am.NumGauss():
        with pytest.raises(L.MfaError):
            dm.fmllr_acc(sc["feats"][0], np.ones(sc["feats"][0].shape[0], np.int32), fo1[:2], np.zeros(1, np.int32), 1, post_model=dm2)
    # equal alignment of an empty batch / accumulators on no frames
    empty = E.FstBatch.from_fsts([])
    a, w, wo, nw, st = empty.equal_align(np.zeros(1, np.int64), [])
    assert a.shape[0] == 0 and st.shape[0] == 0
    dm.acc_zero(); dm.acc_stats(np.zeros((0, D), np.float32), np.zeros(0, np.int32))
    assert dm.acc_read()["frames"] == 0
    for x in (graphs, batch, dm, dm2, empty):
        x.close()


def test_two_engines_on_one_gpu_from_two_threads(eng):
    """MFA runs several jobs per device.  Two engines driven from two host threads (ctypes releases the GIL), each aligning the same batch
    repeatedly with a 1-group band -- so that the polling wide-band level and the sparse level both run while the other engine's launches
    are in flight -- must each reproduce the single-engine result bit for bit."""
    import threading
    sc = build_synth_scenario(seconds=80.0, seed=5, triphone=True, n_phones=12, n_words=60, target_pdfs=100, gauss_per_pdf=2)
    batch = E.GraphCompiler(sc["tm"], sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts)
    want = _align_env(eng, sc, batch, 10.0, 40.0)
    eng2 = E.Engine(0)
    results, errors = {}, []

    def work(j, en):
        try:
            with en.options(vit_maxgroups=1, vit_wide_poll=1):
                results[j] = [_gpu_align_from_oracle_loglikes(en, sc, batch, 10.0, 40.0) for _ in range(4)]
        except Exception as ex:   # surfaced in the main thread
            errors.append(repr(ex))

    th = [threading.Thread(target=work, args=(j, en)) for j, en in enumerate((eng, eng2))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for j in (0, 1):
        for got in results[j]:
            assert np.array_equal(got.status, want.status) and np.array_equal(got.ali, want.ali) and np.array_equal(got.words, want.words)
            assert np.array_equal(got.total_like, want.total_like) and np.array_equal(got.per_frame, want.per_frame)
    assert eng2.band_fallbacks > 0
    eng2.close()


def test_very_long_utterances_fused_path(eng):
    """Maximum sizes: utterances of 2-4 minutes (12 000-24 000 frames, graphs of ~10^4 states -- near the 16-bit state / arc indices of the
    packed graphs) next to short ones, through the fused PCM -> alignment call; every utterance against the oracle chain."""
    sc = build_synth_scenario(seconds=600.0, seed=31, triphone=True, n_phones=10, n_words=50, target_pdfs=80, gauss_per_pdf=2, n_spk=2,
                              mean_utt_s=150.0, min_utt_s=2.0, max_utt_s=240.0)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    T = np.diff(sc["frame_off"])
    assert T.max() >= 11000
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    fsts = batch.export()
    assert max(f.num_states for f in fsts) > 8000
    ref = oracle_align_all(sc, fsts, 10.0, 40.0)
    dm = E.DeviceModel(eng, tm, am)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    res = E.align_pcm(eng, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "deltas")
    eng.sync()
    same = total = 0
    for u, r in enumerate(ref):
        got = res.utterance(u)
        assert got["status"] == r["status"], (u, got["status"], r["status"])
        if r["status"] >= 2:
            continue
        same += int((got["ali"] == r["ali"]).sum()); total += len(r["ali"])
        assert list(got["words"]) == list(r["words"])
        assert abs(got["like"] - r["like"]) <= 1e-4 * abs(r["like"])
    assert total >= 50000 and same / total >= 0.999, (same, total)
    dm.close()


def test_graph_beyond_16_bit_views_fails_alone(eng):
    """A training graph with more than 65 534 states (minutes of continuous speech) cannot use the packed 16-bit graph views: that
    utterance gets MFA_ALIGN_GRAPH_TOO_LARGE (5), its neighbours in the batch are aligned exactly as without it."""
    sc = build_synth_scenario(seconds=12.0, seed=4, n_phones=6, n_words=20, gauss_per_pdf=2)
    tm, am = sc["tm"], sc["am"]
    from mfa_b200.kaldi_io import Fst
    fsts = E.GraphCompiler(tm, sc["tree"], sc["corpus"].lexicon).compile(sc["corpus"].transcripts).export()
    S = 70000
    tid = int(fsts[0].arc_ilabel[fsts[0].arc_ilabel > 0][0])
    fin = np.full(S, np.inf, np.float32); fin[-1] = 0.0
    chain = Fst(0, S, np.arange(S - 1, dtype=np.int32), np.full(S - 1, tid, np.int32), np.zeros(S - 1, np.int32), np.arange(1, S, dtype=np.int32),
                np.zeros(S - 1, np.float32), fin)
    g = O.GmmModel.from_am(am)
    lls = [O.gmm_loglikes(g, sc["feats"][0]), O.gmm_loglikes(g, sc["feats"][1]), O.gmm_loglikes(g, sc["feats"][2])]
    fo = np.zeros(4, np.int64)
    fo[1:] = np.cumsum([x.shape[0] for x in lls])
    dm = E.DeviceModel(eng, tm, am)
    with_big = E.align_loglikes(eng, dm, E.Graphs(E.FstBatch.from_fsts([fsts[0], chain, fsts[2]]), tm), np.concatenate(lls), fo, E.align_opts())
    fo2 = np.asarray([0, lls[0].shape[0], lls[0].shape[0] + lls[2].shape[0]], np.int64)
    without = E.align_loglikes(eng, dm, E.Graphs(E.FstBatch.from_fsts([fsts[0], fsts[2]]), tm), np.concatenate([lls[0], lls[2]]), fo2, E.align_opts())
    assert [int(x) for x in with_big.status] == [int(without.status[0]), 5, int(without.status[1])]
    assert int(with_big.status[0]) in (0, 1) and int(with_big.status[2]) in (0, 1)
    for a, b in ((0, 0), (2, 1)):
        x, y = with_big.utterance(a), without.utterance(b)
        assert np.array_equal(x["ali"], y["ali"]) and list(x["words"]) == list(y["words"]) and x["like"] == y["like"]
    assert with_big.utterance(1)["status"] == 5 and int(with_big.num_words[1]) == 0
    dm.close()


def test_single_shot_from_transcripts_pipelined_equals_one_call(eng):
    """align_pcm_from_transcripts compiles / packs piece k + 1 while piece k is aligned; cut at speaker boundaries, every utterance must
    come out exactly as from ONE compile + ONE fused call over the whole batch (per-speaker CMVN, per-utterance scoring and Viterbi)."""
    sc = build_synth_scenario(seconds=90.0, seed=13, triphone=True, n_phones=10, n_words=50, target_pdfs=80, gauss_per_pdf=2, n_spk=5)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    order = np.argsort(c.utt2spk, kind="stable")             # a job ordered by speaker, as MFA's are
    pcm = np.concatenate([c.pcm[c.sample_off[u]:c.sample_off[u + 1]] for u in order])
    so = np.zeros(c.n_utts + 1, np.int64); so[1:] = np.cumsum([c.sample_off[u + 1] - c.sample_off[u] for u in order])
    u2s = c.utt2spk[order].astype(np.int32)
    tr = [c.transcripts[u] for u in order]
    dm = E.DeviceModel(eng, tm, am)
    gc = E.GraphCompiler(tm, sc["tree"], c.lexicon)
    batch = gc.compile(tr)
    one = E.align_pcm(eng, dm, E.Graphs(batch, tm, 1.0, 0.1), pcm, so, u2s, c.n_spk, E.mfcc_opts(), "deltas")
    piped, times = E.align_pcm_from_transcripts(eng, gc, dm, tr, pcm, so, u2s, c.n_spk, E.mfcc_opts(), "deltas", n_segments=3, n_threads=4)
    assert times["segments"] == 3 and len(times["align_ms"]) == 3
    assert np.array_equal(piped.status, one.status) and np.array_equal(piped.frame_off, one.frame_off) and np.array_equal(piped.num_words, one.num_words)
    assert np.array_equal(piped.total_like, one.total_like)
    for u in range(c.n_utts):
        a, b = piped.utterance(u), one.utterance(u)
        assert np.array_equal(a["ali"], b["ali"]) and np.array_equal(a["per_frame"], b["per_frame"]) and list(a["words"]) == list(b["words"])
    assert int((one.status < 2).sum()) >= c.n_utts - 1
    dm.close(); gc.close()
