"""-m gpu: MFA's corpus path and online path end to end through the kalpy-shaped API and the job functions, on files, against the
oracle chain (including the 8-bit CompressedMatrix round trips of the corpus path)."""
from pathlib import Path

import numpy as np
import pytest

from helpers import build_synth_scenario, gold, mono_sample_setup
from mfa_b200 import kaldi_io as K, kalpy_compat as KC, mfa_functions as MF
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _codec(m):
    return K.decompress_matrix(K.compress_matrix(m))


def _codec_close(a, b):
    """Two decodes of (nearly) the same matrix: a 1e-6 upstream difference perturbs the codec's header (all cells move by
    ~1e-6 of the range) and flips a rare cell by one 8-bit step (<= range/64)."""
    rng = b.max(0) - b.min(0) + 1e-6
    d = np.abs(a - b) / rng
    return float(np.mean(d < 1e-3)), float(d.max())


def test_corpus_path_files_and_training_iteration(tmp_path):
    sc = build_synth_scenario(seconds=50.0, seed=17, triphone=False, n_phones=8, n_words=40, gauss_per_pdf=2, n_spk=3)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    split = tmp_path / "split2"; work = tmp_path / "work"
    split.mkdir(); work.mkdir()
    K.write_gmm_model(work / "1.mdl", tm, am)
    K.write_tree(work / "tree", sc["tree"])
    id2w = c.lexicon.id2word
    utts = []
    for u in range(c.n_utts):
        wav = tmp_path / f"u{u}.wav"
        K.write_wav_int16(wav, c.pcm[c.sample_off[u]:c.sample_off[u + 1]])
        utts.append(MF.Utterance(u, int(c.utt2spk[u]), str(wav), " ".join(id2w[w] for w in c.transcripts[u]),
                                 duration=(c.sample_off[u + 1] - c.sample_off[u]) / 16000.0))
    jobs = MF.assign_jobs(utts, 2, split)
    # a1: MFCC (compressed ark), a3: CMVN, a4: final features (CMVN applied, re-compressed)
    mc = KC.MfccComputer(use_energy=False, dither=0.0, snip_edges=True)
    n_done = []
    list(MF.run_kaldi_function(MF.MfccFunction, [MF.MfccArguments(j.id, j, None, split, mc) for j in jobs], progress=n_done.append))
    assert sum(n_done) == c.n_utts
    MF.calc_cmvn(jobs, split)
    list(MF.run_kaldi_function(MF.FinalFeatureFunction, [MF.FinalFeatureArguments(j.id, j, None, split) for j in jobs]))
    # oracle chain with the same codec round trips.  The 8-bit codec turns a 1e-6 upstream difference into (rarely) one
    # quantisation step, so each stage is checked on the files the previous GPU stage actually wrote.
    raw_gpu = {}
    for j in jobs:
        raw_gpu.update(dict(K.read_ark(j.construct_path(split, "feats", "ark"), "matrix")))
    raw_ref = {u.kaldi_id: _codec(O.mfcc(c.pcm[c.sample_off[u.id]:c.sample_off[u.id + 1]])) for u in utts}
    assert set(raw_gpu) == set(raw_ref)
    cc = [_codec_close(raw_gpu[k], raw_ref[k]) for k in raw_ref]
    assert np.mean([x[0] for x in cc]) > 0.995 and max(x[1] for x in cc) <= 1.0 / 60
    stats = {s: O.cmvn_stats([raw_gpu[u.kaldi_id] for u in utts if u.speaker_id == s]) for s in range(c.n_spk)}
    cm = dict(K.read_ark(split / "cmvn.ark", "matrix"))
    for s in range(c.n_spk):
        assert np.allclose(cm[str(s)], stats[s], rtol=1e-9, atol=1e-6)
    final = {u.kaldi_id: _codec(O.cmvn_apply(raw_gpu[u.kaldi_id], stats[u.speaker_id])) for u in utts}
    got_final = {}
    for j in jobs:
        for k, p, o in K.read_scp(j.construct_path(split, "feats", "scp")):
            got_final[k] = K.read_scp_object(p, o, "matrix")
    assert set(got_final) == set(final)
    cc = [_codec_close(got_final[k], final[k]) for k in final]
    assert np.mean([x[0] for x in cc]) > 0.995 and max(x[1] for x in cc) <= 1.0 / 60
    # a6: graphs, a7/a8: alignment
    lex = {1: c.lexicon}
    list(MF.run_kaldi_function(MF.CompileTrainGraphsFunction,
                               [MF.CompileTrainGraphsArguments(j.id, j, None, work, lex, work / "tree", work / "1.mdl") for j in jobs]))
    opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40, boost_silence=1.0)
    score, n_fail = MF.align_utterances(jobs, work, work / "1.mdl", opts)
    assert n_fail == 0 and np.isfinite(score)
    g = O.GmmModel.from_am(am)
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    same = total = 0
    accs = None
    for j in jobs:
        fsts = KC.FstArchive(j.construct_path(work, "fsts", "ark", 1))
        alis = KC.AlignmentArchive(j.construct_path(work, "ali", "ark", 1), j.construct_path(work, "words", "ark", 1),
                                   j.construct_path(work, "likelihoods", "ark", 1))
        for u in j.utts(1):
            f = O.add_deltas(got_final[u.kaldi_id])   # same (GPU-written) final features: isolates K2/K3 from codec flips
            r = O.align(fsts[u.kaldi_id], tc, g, tm.tid2pdf, f, f.shape[0])
            a = alis[u.kaldi_id]
            assert r["status"] < 2 and a.words == list(r["words"])
            same += int((np.asarray(a.alignment) == r["ali"]).sum()); total += len(r["ali"])
            assert abs(u.alignment_log_likelihood - r["like"]) <= 1e-4 * abs(r["like"])
            accs = O.acc_stats(g, tm.tid2pdf, f, np.asarray(a.alignment, np.int32), tm.num_tids, accs)
    assert same / total >= 0.999
    # a9/a10: statistics -> update -> 2.mdl
    avg, impr, frames = MF.acc_stats(jobs, work, 1, mixup=am.NumGauss() + 10)
    assert frames == accs["frames"] and abs(avg - accs["like"][0] / accs["frames"]) < 1e-4 * abs(avg)
    tm2, am2 = K.read_gmm_model(work / "2.mdl")
    assert am2.NumPdfs() == am.NumPdfs() and am2.NumGauss() >= am.NumGauss() - 5
    # one EM step on the same alignments must not lower the likelihood of the data under the new model
    from mfa_b200.gmm_update import AccumAmDiagGmm
    from oracle.mstep_oracle import mle_update
    ref_new, _, _ = mle_update(am, AccumAmDiagGmm.from_dict(accs), mixup=0)
    like_old = like_new = 0.0
    gn = O.GmmModel.from_am(ref_new)
    for j in jobs:
        alis = KC.AlignmentArchive(j.construct_path(work, "ali", "ark", 1))
        for u in j.utts(1):
            f = O.add_deltas(got_final[u.kaldi_id])
            a = np.asarray(alis[u.kaldi_id].alignment, np.int32)
            like_old += O.acc_stats(g, tm.tid2pdf, f, a, tm.num_tids)["like"][0]
            like_new += O.acc_stats(gn, tm.tid2pdf, f, a, tm.num_tids)["like"][0]
    assert like_new >= like_old - 1e-6 * abs(like_old)


def test_online_path_config1(tmp_path):
    ms = mono_sample_setup(tmp_path)
    tm, am, lex = ms["tm"], ms["am"], ms["lex"]
    K.write_gmm_model(tmp_path / "final.mdl", tm, am)
    K.write_tree(tmp_path / "tree", ms["tree"])
    id2ph = {v: k for k, v in ms["pt"].items()}
    ali, ctm = MF.align_utterance_online(tmp_path / "final.mdl", tmp_path / "tree", lex, ms["pcm"], ms["text"], phone_table=id2ph)
    m = O.mfcc(ms["pcm"])
    f = O.add_deltas(O.cmvn_apply(m, O.cmvn_stats([m])))
    fst = KC.TrainingGraphCompiler(tmp_path / "final.mdl", tmp_path / "tree", lex).compile_fst(ms["text"])
    r = O.align(fst, -tm.scaled_transition_log_probs(1.0, 0.1), O.GmmModel.from_am(am), tm.tid2pdf, f, f.shape[0])
    assert (np.asarray(ali.alignment) == r["ali"]).mean() >= 0.999 and abs(ali.likelihood - r["like"]) <= 1e-4 * abs(r["like"])
    # phone intervals tile the utterance on the 10 ms grid; boundaries within one frame of the oracle's
    assert ctm[0].begin == 0.0 and abs(ctm[-1].end - len(r["ali"]) * 0.01) < 1e-6
    assert all(abs(a.end - b.begin) < 1e-9 for a, b in zip(ctm, ctm[1:]))
    from mfa_b200.kalpy_compat import Alignment
    ref_ctm = Alignment("u", r["ali"], r["words"], r["like"], r["per_frame"]).generate_ctm(tm, id2ph, 0.01)
    assert len(ctm) == len(ref_ctm)
    assert max(abs(a.begin - b.begin) for a, b in zip(ctm, ref_ctm)) <= 0.0100001
    assert [x.label for x in ctm] == [x.label for x in ref_ctm] and ctm[1].label.split("_")[0] in ("dh", "sil")
    # too tight a beam without retry -> AlignerError (online/alignment.py:108-112)
    with pytest.raises(MF.AlignerError):
        MF.align_utterance_online(tmp_path / "final.mdl", tmp_path / "tree", lex, ms["pcm"], ms["text"], beam=0.001, retry_beam=0.0)


def _write_corpus(tmp_path, sc):
    c = sc["corpus"]
    split = tmp_path / "split"; work = tmp_path / "work"
    split.mkdir(); work.mkdir()
    id2w = c.lexicon.id2word
    utts = []
    for u in range(c.n_utts):
        wav = tmp_path / f"u{u}.wav"
        K.write_wav_int16(wav, c.pcm[c.sample_off[u]:c.sample_off[u + 1]])
        utts.append(MF.Utterance(u, int(c.utt2spk[u]), str(wav), " ".join(id2w[w] for w in c.transcripts[u]),
                                 duration=(c.sample_off[u + 1] - c.sample_off[u]) / 16000.0))
    jobs = MF.assign_jobs(utts, 2, split)
    mc = KC.MfccComputer(use_energy=False, dither=0.0, snip_edges=True)
    list(MF.run_kaldi_function(MF.MfccFunction, [MF.MfccArguments(j.id, j, None, split, mc) for j in jobs]))
    MF.calc_cmvn(jobs, split)
    list(MF.run_kaldi_function(MF.FinalFeatureFunction, [MF.FinalFeatureArguments(j.id, j, None, split) for j in jobs]))
    return utts, jobs, split, work


def test_mono_align_equal_iteration_zero(tmp_path):
    """Row a11: equal alignments through the training graphs (== the oracle's EqualAlign), ali archives written, statistics of the
    flat-start model == the oracle's on those alignments, first update -> 1.mdl; one more align/acc-stats pass raises the likelihood."""
    sc = build_synth_scenario(seconds=40.0, seed=23, triphone=False, n_phones=8, n_words=40, gauss_per_pdf=1, n_spk=2)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    utts, jobs, split, work = _write_corpus(tmp_path, sc)
    # flat start (gmm-init-mono): every pdf = one Gaussian with the global mean / variance
    feats_all = {}
    for j in jobs:
        fa = KC.FeatureArchive(j.construct_path(split, "feats", "scp", 1), deltas=True)
        for k, m in fa:
            feats_all[k] = m
    X = np.concatenate(list(feats_all.values()))
    mu, var = X.mean(0), X.var(0)
    P = am.NumPdfs()
    flat = K.AmDiagGmm(am.dim, np.arange(P + 1, dtype=np.int32), np.ones(P, np.float32), np.tile(mu / var, (P, 1)).astype(np.float32),
                       np.tile(1.0 / var, (P, 1)).astype(np.float32))
    K.write_gmm_model(work / "0.mdl", tm, flat)
    K.write_tree(work / "tree", sc["tree"])
    list(MF.run_kaldi_function(MF.CompileTrainGraphsFunction,
                               [MF.CompileTrainGraphsArguments(j.id, j, None, work, {1: c.lexicon}, work / "tree", work / "0.mdl") for j in jobs]))
    avg0, frames = MF.mono_align_equal(jobs, work)
    assert frames == X.shape[0] and np.isfinite(avg0)
    g = O.GmmModel.from_am(flat)
    accs = None
    for j in jobs:
        fsts = KC.FstArchive(j.construct_path(work, "fsts", "ark", 1))
        alis = KC.AlignmentArchive(j.construct_path(work, "ali", "ark", 1))
        for u in j.utts(1):
            f = feats_all[u.kaldi_id]
            r = O.equal_align(fsts[u.kaldi_id], f.shape[0], KC.string_hash(u.kaldi_id))
            assert r["status"] == 0 and alis[u.kaldi_id].alignment == r["ali"].tolist()
            accs = O.acc_stats(g, tm.tid2pdf, f, r["ali"], tm.num_tids, accs)
    assert abs(avg0 - accs["like"][0] / accs["frames"]) <= 1e-5 * abs(avg0)
    tm1, am1 = K.read_gmm_model(work / "1.mdl")
    assert am1.NumPdfs() == P and am1.NumGauss() >= P
    # iteration 1: Viterbi alignments with 1.mdl fit better than the equal alignments did under 0.mdl
    opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=6, retry_beam=40, boost_silence=1.0)
    MF.align_utterances(jobs, work, work / "1.mdl", opts)
    avg1, _, frames1 = MF.acc_stats(jobs, work, 1)
    assert frames1 == frames and avg1 > avg0


def test_two_pass_sat_alignment_with_estimated_fmllr(tmp_path):
    """Rows a12 + N2: pass 1 with the speaker-independent model, per-speaker fMLLR from those alignments (CalcFmllrFunction ->
    trans.D.J.ark/.scp), pass 2 picks the transforms up through construct_feature_archive.  The corpus is given a per-speaker
    affine distortion in feature space so the transforms have something to undo."""
    sc = build_synth_scenario(seconds=90.0, seed=29, triphone=True, n_phones=8, n_words=40, gauss_per_pdf=2, n_spk=3, use_lda=True)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    utts, jobs, split, work = _write_corpus(tmp_path, sc)
    K.write_matrix_file(work / "lda.mat", sc["lda"])
    K.write_gmm_model(work / "final.mdl", tm, am)
    K.write_tree(work / "tree", sc["tree"])
    list(MF.run_kaldi_function(MF.CompileTrainGraphsFunction,
                               [MF.CompileTrainGraphsArguments(j.id, j, None, work, {1: c.lexicon}, work / "tree", work / "final.mdl") for j in jobs]))
    opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40, boost_silence=1.0)
    score1, fail1 = MF.align_utterances(jobs, work, work / "final.mdl", opts)
    sil = [c.lexicon.phone_table["sil"]]
    res = MF.calc_fmllr(jobs, work, work / "final.mdl", work / "final.mdl", dict(silence_weight=0.0), sil)
    assert set(res) == {str(s) for s in range(c.n_spk)}
    # oracle: statistics from the GPU-written final features + pass-1 alignments, per speaker, then the row update
    g = O.GmmModel.from_am(am)
    tw = np.where(tm.tid2phone == sil[0], np.float32(0), np.float32(1)).astype(np.float32)
    tw[0] = 0
    D = am.dim
    stats = np.zeros((c.n_spk, O.fmllr_stats_size(D)))
    base = {}
    for j in jobs:
        scp = {k: (p, o) for k, p, o in K.read_scp(j.construct_path(split, "feats", "scp"))}
        alis = KC.AlignmentArchive(j.construct_path(work, "ali", "ark", 1))
        for u in j.utts(1):
            m = K.read_scp_object(*scp[u.kaldi_id], "matrix")
            f = O.transform(O.splice(np.asarray(m, np.float32), 3, 3), sc["lda"])
            base[u.kaldi_id] = f
            if u.kaldi_id in alis:
                O.fmllr_acc(g, g, tm.tid2pdf, tw, f, np.asarray(alis[u.kaldi_id].alignment, np.int32), stats[u.speaker_id])
    for j in jobs:
        tr = KC.MatrixArchive(j.construct_path(split, "trans", "scp", 1))
        for s in sorted({u.speaker_id for u in j.utts(1)}):
            Wo, io = O.fmllr_update(stats[s], D)
            assert np.abs(tr[str(s)] - Wo).max() < 2e-3
            assert abs(res[str(s)][1] - stats[s, 0]) <= 1e-4 * stats[s, 0]
            if stats[s, 0] > 500:
                assert res[str(s)][0] > 0
    # pass 2: features now carry the transforms; per-utterance likelihoods are those of the oracle on transformed features
    score2, fail2 = MF.align_utterances(jobs, work, work / "final.mdl", opts)
    assert fail2 <= fail1 and np.isfinite(score2)   # (the decoder's score omits the log|det A| Jacobian, so it need not rise)
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    same = total = 0
    for j in jobs:
        tr = KC.MatrixArchive(j.construct_path(split, "trans", "scp", 1))
        fsts = KC.FstArchive(j.construct_path(work, "fsts", "ark", 1))
        alis = KC.AlignmentArchive(j.construct_path(work, "ali", "ark", 1))
        for u in j.utts(1)[:6]:
            f = O.transform(base[u.kaldi_id], tr[str(u.speaker_id)])
            r = O.align(fsts[u.kaldi_id], tc, g, tm.tid2pdf, f, f.shape[0])
            same += int((np.asarray(alis[u.kaldi_id].alignment) == r["ali"]).sum()); total += len(r["ali"])
            assert abs(u.alignment_log_likelihood - r["like"]) <= 1e-4 * abs(r["like"])
    assert same / total >= 0.999
    # a second estimation composes on top of the first (features.py:486-495): the composed transform stays close to the first
    res2 = MF.calc_fmllr(jobs, work, work / "final.mdl", work / "final.mdl", dict(silence_weight=0.0), sil)
    assert all(v[0] >= 0.0 for v in res2.values())


def test_textgrid_export_corpus_and_reference_sanity(tmp_path):
    """Row N4 end to end: corpus path alignments -> AlignmentExtractionFunction -> TextGrids; boundaries equal the oracle's CTM within one
    frame; and the sample utterance against the reference repo's own TextGrid for it (made upstream with a different English model:
    a loose sanity bound)."""
    from mfa_b200 import export as X
    sc = build_synth_scenario(seconds=30.0, seed=31, triphone=False, n_phones=8, n_words=40, gauss_per_pdf=2, n_spk=2)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    utts, jobs, split, work = _write_corpus(tmp_path, sc)
    K.write_gmm_model(work / "final.mdl", tm, am)
    K.write_tree(work / "tree", sc["tree"])
    list(MF.run_kaldi_function(MF.CompileTrainGraphsFunction,
                               [MF.CompileTrainGraphsArguments(j.id, j, None, work, {1: c.lexicon}, work / "tree", work / "final.mdl") for j in jobs]))
    opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40, boost_silence=1.0)
    MF.align_utterances(jobs, work, work / "final.mdl", opts)
    written = MF.export_textgrids(jobs, work, work / "final.mdl", {1: c.lexicon}, tmp_path / "out")
    assert len(written) == c.n_utts and all(p is not None and p.exists() for p in written.values())
    g = O.GmmModel.from_am(am)
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    id2ph = {v: k for k, v in c.lexicon.phone_table.items()}
    n_b = 0
    for j in jobs:
        fsts = KC.FstArchive(j.construct_path(work, "fsts", "ark", 1))
        fa = j.construct_feature_archive(work, 1)
        for u in j.utts(1)[:5]:
            f = fa[u.kaldi_id]
            r = O.align(fsts[u.kaldi_id], tc, g, tm.tid2pdf, f, f.shape[0])
            ref = X.alignment_to_ctm(KC.Alignment(u.kaldi_id, r["ali"], r["words"], r["like"], r["per_frame"]), tm, c.lexicon, 0.01, None, None,
                                     u.normalized_text)
            tiers = X.read_textgrid(written[u.id])
            got_w = [e for e in tiers["words"] if e[2]]
            ref_w = [w for w in ref.word_intervals if w.label != "<eps>"]
            assert [e[2] for e in got_w] == [w.label for w in ref_w] == u.normalized_text.split()
            for e, w in zip(got_w, ref_w):
                assert abs(e[0] - w.begin) <= 0.0100001 and (abs(e[1] - w.end) <= 0.0100001 or e is got_w[-1])
                n_b += 1
            got_p = [e for e in tiers["phones"] if e[2]]
            ref_p = [p for w in ref_w for p in w.phones]
            assert [e[2] for e in got_p] == [p.label for p in ref_p]
            assert tiers["words"][-1][1] == pytest.approx(u.duration, abs=1e-5)
    assert n_b > 20
    # sample utterance (config 1) vs the reference repo's TextGrid
    ms = mono_sample_setup(tmp_path)
    K.write_gmm_model(tmp_path / "mono.mdl", ms["tm"], ms["am"])
    K.write_tree(tmp_path / "mono.tree", ms["tree"])
    ali, _ = MF.align_utterance_online(tmp_path / "mono.mdl", tmp_path / "mono.tree", ms["lex"], ms["pcm"], ms["text"])
    ctm = X.alignment_to_ctm(ali, ms["tm"], ms["lex"], 0.01, None, None, ms["text"])
    ours = [w for w in ctm.word_intervals if w.label != "<eps>"]
    assert [w.label for w in ours] == ms["text"].split()
    rp = tmp_path / "ref.TextGrid"; rp.write_bytes(bytes(gold()["acoustic_corpus_textgrid"]))
    ref_words = [e for e in X.read_textgrid(rp)["words"] if e[2]]
    # the reference transcript tier may tokenise differently; compare the words both have, in order
    import difflib
    sm = difflib.SequenceMatcher(a=[w.label for w in ours], b=[e[2] for e in ref_words], autojunk=False)
    diffs = []
    for blk in sm.get_matching_blocks():
        for k in range(blk.size):
            w, e = ours[blk.a + k], ref_words[blk.b + k]
            diffs += [abs(w.begin - e[0]), abs(w.end - e[1])]
    diffs = np.asarray(diffs)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/textgrid_sanity.txt", "w") as f:
        f.write(f"matched boundaries {diffs.size}, median {np.median(diffs):.3f} s, p90 {np.percentile(diffs, 90):.3f} s, max {diffs.max():.3f} s\n")
    # informational only: the fixture monophone model is a unit-test artefact (132 single Gaussians) whose alignments -- ours and the
    # oracle's alike, see test_online_path_config1 -- are seconds away from a production model's; there is nothing to assert on here
    assert diffs.size >= 60


def test_align_one_file_with_segments(tmp_path):
    """`mfa align_one` flow (command_line/align_one.py:157-196): one sound file holding two segments with their own transcripts, one CMVN
    over the file, per-segment alignment through KalpyUtterance, one TextGrid whose intervals sit inside the segments."""
    from mfa_b200 import export as X
    sc = build_synth_scenario(seconds=20.0, seed=37, triphone=False, n_phones=8, n_words=30, gauss_per_pdf=2, n_spk=1)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    K.write_gmm_model(tmp_path / "final.mdl", tm, am)
    K.write_tree(tmp_path / "tree", sc["tree"])
    gap = np.zeros(8000, np.int16)
    a, b = c.pcm[c.sample_off[0]:c.sample_off[1]], c.pcm[c.sample_off[1]:c.sample_off[2]]
    pcm = np.concatenate([gap, a, gap, b, gap])
    wav = tmp_path / "file.wav"
    K.write_wav_int16(wav, pcm)
    t = lambda n: n / 16000.0
    segs = [(t(8000), t(8000 + len(a)), 0, " ".join(c.lexicon.id2word[w] for w in c.transcripts[0])),
            (t(16000 + len(a)), t(16000 + len(a) + len(b)), 0, " ".join(c.lexicon.id2word[w] for w in c.transcripts[1]))]
    ctm = MF.align_one(wav, segs, tmp_path / "final.mdl", tmp_path / "tree", c.lexicon, tmp_path / "out" / "file.TextGrid", t(len(pcm)))
    tiers = X.read_textgrid(tmp_path / "out" / "file.TextGrid")
    words = [e for e in tiers["words"] if e[2]]
    assert [e[2] for e in words] == (segs[0][3] + " " + segs[1][3]).split()
    n0 = len(segs[0][3].split())
    assert all(segs[0][0] - 1e-6 <= e[0] and e[1] <= segs[0][1] + 0.011 for e in words[:n0])
    assert all(segs[1][0] - 1e-6 <= e[0] and e[1] <= segs[1][1] + 0.011 for e in words[n0:])
    assert tiers["words"][0] == (0.0, pytest.approx(words[0][0]), "") and abs(tiers["phones"][-1][1] - t(len(pcm))) < 1e-6
    assert len(ctm.word_intervals) >= len(words)
    # the same segment aligned on its own (its own CMVN) gives the same word sequence through the KalpyUtterance path
    u = KC.KalpyUtterance(KC.Segment(str(wav), segs[0][0], segs[0][1], 0), segs[0][3])
    one = MF.align_utterance_online_ctm(tmp_path / "final.mdl", tmp_path / "tree", c.lexicon, u)
    assert [w.label for w in one.word_intervals if w.label != "<eps>"] == segs[0][3].split()
    assert one.word_intervals[0].begin >= segs[0][0] - 1e-6


def test_two_pass_align_pcm_in_memory():
    """The in-memory two-pass flow (pass 1 -> K5 statistics + transforms -> pass 2) on numpy and on device-resident PCM: transforms equal
    the oracle's on the oracle's features and pass-1 alignments; pass 2 equals the oracle on the transformed features."""
    import torch
    from helpers import oracle_align_all
    from mfa_b200 import engine as E
    sc = build_synth_scenario(seconds=80.0, seed=43, triphone=True, n_phones=8, n_words=40, gauss_per_pdf=2, n_spk=3, use_lda=True)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    eng = KC.get_engine()
    dm = E.DeviceModel(eng, tm, am)
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    sil = [c.lexicon.phone_table["sil"]]
    r1, W, (impr, count), r2 = MF.two_pass_align_pcm(eng, dm, dm, graphs, c.pcm, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda",
                                                     lda=sc["lda"], silence_phone_ids=sil, min_count=100.0)
    fsts = batch.export()
    ref1 = oracle_align_all(sc, fsts)
    g = O.GmmModel.from_am(am)
    tw = np.where(tm.tid2phone == sil[0], np.float32(0), np.float32(1)).astype(np.float32); tw[0] = 0
    D = am.dim
    stats = np.zeros((c.n_spk, O.fmllr_stats_size(D)))
    for u, r in enumerate(ref1):
        if r["status"] < 2:
            O.fmllr_acc(g, g, tm.tid2pdf, tw, sc["feats"][u], r["ali"], stats[c.utt2spk[u]])
    same = total = 0
    tc = -tm.scaled_transition_log_probs(1.0, 0.1)
    for s in range(c.n_spk):
        Wo, io = O.fmllr_update(stats[s], D, min_count=100.0)
        assert np.abs(W[s] - Wo).max() < 2e-3 and abs(count[s] - stats[s, 0]) <= 1e-3 * stats[s, 0]
    for u in range(min(c.n_utts, 8)):
        f = O.transform(sc["feats"][u], W[c.utt2spk[u]])
        r = O.align(fsts[u], tc, g, tm.tid2pdf, f, f.shape[0])
        got = r2.utterance(u)
        assert got["status"] == r["status"]
        same += int((got["ali"] == r["ali"]).sum()); total += len(r["ali"])
    assert same / total >= 0.999
    # device-resident PCM: same transforms (f64 atomics reorder the statistics, so to rounding) and the same pass-2 alignments
    d = torch.from_numpy(c.pcm).cuda()
    _, Wd, _, r2d = MF.two_pass_align_pcm(eng, dm, dm, graphs, d, c.sample_off, c.utt2spk, c.n_spk, E.mfcc_opts(), "lda", lda=sc["lda"],
                                          silence_phone_ids=sil, min_count=100.0)
    eng.sync()
    assert np.abs(Wd - W).max() < 1e-4
    assert (r2d.ali.cpu().numpy()[: r2.ali.shape[0]] == r2.ali).mean() >= 0.999
    graphs.close(); batch.close(); dm.close()


def test_jobs_as_threads_equal_jobs_in_sequence(tmp_path):
    """MFA's USE_THREADING mode: the jobs of a stage run as threads of one process (utils.py:1560-1580), every thread on its own engine
    (kalpy_compat.get_engine is per thread; the ONE MfccComputer of MfccArguments is shared by all jobs).  Every archive a stage writes
    must be byte-identical to the sequential run."""
    import threading
    sc = build_synth_scenario(seconds=60.0, seed=23, triphone=False, n_phones=8, n_words=40, gauss_per_pdf=2, n_spk=4)
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    id2w = c.lexicon.id2word
    outs = {}
    for mode, nthr in (("seq", 1), ("thr", 4)):
        split = tmp_path / mode / "split"; work = tmp_path / mode / "work"
        split.mkdir(parents=True); work.mkdir(parents=True)
        K.write_gmm_model(work / "1.mdl", tm, am)
        K.write_tree(work / "tree", sc["tree"])
        utts = []
        for u in range(c.n_utts):
            wav = tmp_path / mode / f"u{u}.wav"
            K.write_wav_int16(wav, c.pcm[c.sample_off[u]:c.sample_off[u + 1]])
            utts.append(MF.Utterance(u, int(c.utt2spk[u]), str(wav), " ".join(id2w[w] for w in c.transcripts[u]),
                                     duration=(c.sample_off[u + 1] - c.sample_off[u]) / 16000.0))
        jobs = MF.assign_jobs(utts, 4, split)
        mc = KC.MfccComputer(use_energy=False, dither=0.0, snip_edges=True)
        list(MF.run_kaldi_function(MF.MfccFunction, [MF.MfccArguments(j.id, j, None, split, mc) for j in jobs], num_threads=nthr))
        MF.calc_cmvn(jobs, split)
        list(MF.run_kaldi_function(MF.FinalFeatureFunction, [MF.FinalFeatureArguments(j.id, j, None, split) for j in jobs], num_threads=nthr))
        lex = {1: c.lexicon}
        list(MF.run_kaldi_function(MF.CompileTrainGraphsFunction,
                                   [MF.CompileTrainGraphsArguments(j.id, j, None, work, lex, work / "tree", work / "1.mdl") for j in jobs], num_threads=nthr))
        opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40, boost_silence=1.0)
        score, n_fail = MF.align_utterances(jobs, work, work / "1.mdl", opts, num_threads=nthr)
        assert n_fail == 0
        files = {}
        for j in jobs:
            for name, d in (("feats", split), ("fsts", work), ("ali", work), ("words", work), ("likelihoods", work)):
                pth = j.construct_path(d, name, "ark", 1) if d is work else j.construct_path(d, name, "ark")
                files[(j.id, name)] = Path(pth).read_bytes()
        outs[mode] = (score, files)
    assert outs["seq"][0] == outs["thr"][0]
    assert outs["seq"][1].keys() == outs["thr"][1].keys() and len(outs["seq"][1]) == 20
    for k in outs["seq"][1]:
        assert outs["seq"][1][k] == outs["thr"][1][k], k
    assert 2 <= len(KC._engines) <= 8     # the worker threads had their own engines, and the four stages' thread pools shared them
