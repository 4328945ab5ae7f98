"""Generates tests/golden/kalpy_fixtures.npz from the REAL reference stack (kalpy over Kaldi + OpenFst), on any machine where
``import kalpy`` gives the real package (conda-forge kalpy / PyPI kalpy-kaldi 0.6.x; not installable in the offline build image, so the
committed repo carries no such file and DESIGN.md says "parity unpinned").  One command turns every comparison in
tests/test_reference_kalpy.py from "skipped" into a diff against Kaldi itself:

    python tools/make_kalpy_fixtures.py [--model-dir DIR --dict DICT.txt --wav X.wav --text "..."]

By default it uses the reference's own fixtures as committed in tests/golden/reference_fixtures.npz (mono_model, test_acoustic dictionary,
acoustic_corpus.wav + .lab): the same inputs the oracle / engine tests use.  The flow is MFA's DB-free path (online/alignment.py:29-123,
command_line/align_one.py:118-196), dither = 0.  Every stage is guarded: what a given kalpy version cannot produce is listed under
``errors`` in the file instead of aborting the run.
"""
import argparse
import json
import os
import sys
import tempfile
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "kalpy_fixtures.npz")


def np_of(x):
    for attr in ("numpy",):
        if hasattr(x, attr):
            return np.asarray(getattr(x, attr)())
    return np.asarray(x)


def fst_arrays(fst):
    """pywrapfst.VectorFst -> (start, finals, src, dst, ilabel, olabel, weight)"""
    if hasattr(fst, "arc_src"):   # the shim's Fst (plumbing check only)
        return (int(fst.start), np.asarray(fst.finals, np.float32), fst.arc_src, fst.arc_dst, fst.arc_ilabel, fst.arc_olabel, fst.arc_weight)
    src, dst, il, ol, w, fin = [], [], [], [], [], []
    for s in fst.states():
        fw = fst.final(s)
        fin.append(float(fw) if str(fw) != "Infinity" else np.inf)
        for a in fst.arcs(s):
            src.append(s); dst.append(a.nextstate); il.append(a.ilabel); ol.append(a.olabel); w.append(float(a.weight))
    return (int(fst.start()), np.asarray(fin, np.float32), np.asarray(src, np.int32), np.asarray(dst, np.int32), np.asarray(il, np.int32),
            np.asarray(ol, np.int32), np.asarray(w, np.float32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-dir"); ap.add_argument("--dict"); ap.add_argument("--wav"); ap.add_argument("--text")
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--allow-shim", action="store_true", help="PLUMBING CHECK ONLY: run the same flow through the mfa_b200 kalpy shim (needs a "
                    "GPU) and write to --out; such a file pins nothing and must not be committed as kalpy_fixtures.npz")
    args = ap.parse_args()
    if args.allow_shim:
        import mfa_b200
        mfa_b200.install_kalpy_shim()
        if os.path.abspath(args.out) == os.path.abspath(OUT):
            raise SystemExit("--allow-shim needs an --out path other than the golden file")
    import kalpy
    if getattr(kalpy, "__mfa_b200_shim__", False) and not args.allow_shim:
        raise SystemExit("`import kalpy` resolved to the mfa_b200 shim: run this with the real kalpy on the path")
    from helpers import gold
    from mfa_b200 import kaldi_io as K
    g = gold()
    tmp = tempfile.mkdtemp()
    if args.model_dir:
        mdl, tree = os.path.join(args.model_dir, "final.mdl"), os.path.join(args.model_dir, "tree")
        phones = None
    else:
        from helpers import load_model
        tm, am, cd = load_model("mono")
        mdl, tree = os.path.join(tmp, "final.mdl"), os.path.join(tmp, "tree")
        K.write_gmm_model(mdl, tm, am); K.write_tree(tree, cd)
        phones = json.loads(bytes(g["mono_phones"]).decode())
    dict_path = args.dict or os.path.join(tmp, "dict.txt")
    if not args.dict:
        open(dict_path, "wb").write(bytes(g["test_acoustic_dict"]))
    wav = args.wav or os.path.join(tmp, "utt.wav")
    if not args.wav:
        K.write_wav_int16(wav, g["acoustic_corpus_pcm"])
    text = args.text or bytes(g["acoustic_corpus_lab"]).decode().strip().lower()
    out, errors = {"kalpy_version": np.frombuffer(str(getattr(kalpy, "__version__", "?")).encode(), np.uint8),
                   "text": np.frombuffer(text.encode(), np.uint8)}, {}

    def stage(name, fn):
        try:
            r = fn()
            if r is not None:
                out.update(r)
        except Exception:
            errors[name] = traceback.format_exc(limit=4)
            print(f"[kalpy-fixtures] stage {name} failed:\n{errors[name]}", file=sys.stderr)

    from kalpy.feat.cmvn import CmvnComputer
    from kalpy.feat.mfcc import MfccComputer
    from kalpy.utterance import Segment, Utterance
    opts = dict(use_energy=False, dither=0.0, energy_floor=0.0, snip_edges=True, sample_frequency=16000, frame_length=25, frame_shift=10,
                num_mel_bins=23, num_coefficients=13, low_frequency=20, high_frequency=7800, preemphasis_coefficient=0.97, cepstral_lifter=22)
    mc = MfccComputer(**opts)
    pcm, sr = K.read_wav_int16(wav)
    utt = Utterance(Segment(wav, 0.0, pcm.shape[0] / sr, 0), text)
    st = {}

    def s_mfcc():
        utt.generate_mfccs(mc)
        st["mfcc"] = utt.mfccs
        return {"pcm": pcm, "mfcc": np_of(utt.mfccs)}
    stage("mfcc", s_mfcc)

    def s_cmvn():
        st["cmvn"] = CmvnComputer().compute_cmvn_from_features([utt.mfccs])
        utt.apply_cmvn(st["cmvn"])
        return {"cmvn_stats": np_of(st["cmvn"])}
    stage("cmvn", s_cmvn)

    def s_feats():
        st["feats"] = utt.generate_features(mc, None, lda_mat=None, fmllr_trans=None)
        return {"feats_deltas": np_of(st["feats"])}
    stage("features", s_feats)

    def s_lda():
        from helpers import gold as _g
        lda = _g()["g2p_lda"]
        from _kalpy.matrix import FloatMatrix
        m = FloatMatrix()
        m.from_numpy(np.ascontiguousarray(lda, np.float32))
        u2 = Utterance(Segment(wav, 0.0, pcm.shape[0] / sr, 0), text)
        u2.generate_mfccs(mc); u2.apply_cmvn(st["cmvn"])
        return {"feats_lda": np_of(u2.generate_features(mc, None, lda_mat=m, fmllr_trans=None)), "lda": lda}
    stage("splice_lda", s_lda)

    from kalpy.gmm.utils import read_gmm_model
    tmk, amk = read_gmm_model(mdl)

    def s_likes():
        try:
            from _kalpy.gmm import gmm_compute_likes
        except ImportError:
            from kalpy.gmm.utils import gmm_compute_likes
        return {"loglikes": np_of(gmm_compute_likes(amk, st["feats"]))}
    stage("gmm_compute_likes", s_likes)

    def s_graph():
        from kalpy.decoder.training_graphs import TrainingGraphCompiler
        from kalpy.fstext.lexicon import LexiconCompiler
        kw = dict(silence_probability=0.5, initial_silence_probability=0.5, position_dependent_phones=True)
        if phones is not None:
            kw["phones"] = set(phones)
        lc = LexiconCompiler(**kw)
        lc.load_pronunciations(dict_path)
        st["lc"] = lc
        gc = TrainingGraphCompiler(mdl, tree, lc)
        st["fst"] = gc.compile_fst(text)
        start, fin, src, dst, il, ol, w = fst_arrays(st["fst"])
        words = {lc.word_table.find(i): i for i in range(lc.word_table.num_symbols())}
        return {"fst_start": np.asarray([start]), "fst_finals": fin, "fst_src": src, "fst_dst": dst, "fst_ilabel": il, "fst_olabel": ol,
                "fst_weight": w, "word_table": np.frombuffer(json.dumps(words).encode(), np.uint8)}
    stage("training_graph", s_graph)

    def s_align():
        from kalpy.gmm.align import GmmAligner
        al = GmmAligner(mdl, beam=10, retry_beam=40, transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1)
        a = al.align_utterance(st["fst"], st["feats"])
        st["ali"] = a
        r = {"ali": np.asarray(a.alignment, np.int32), "words": np.asarray(a.words, np.int32), "likelihood": np.asarray([a.likelihood], np.float64)}
        if getattr(a, "per_frame_likelihoods", None) is not None:
            r["per_frame"] = np_of(a.per_frame_likelihoods).astype(np.float32)
        ctm = a.generate_ctm(al.transition_model, st["lc"].phone_table, 0.01)
        r["ctm"] = np.frombuffer(json.dumps([[float(i.begin), float(i.end), str(i.label)] for i in ctm]).encode(), np.uint8)
        return r
    stage("align_utterance", s_align)

    def s_equal():
        from kalpy.gmm.align import gmm_align_equal
        a, w = gmm_align_equal(st["fst"], st["feats"])
        return {"equal_ali": np.asarray(a, np.int32), "equal_words": np.asarray(w, np.int32)}
    stage("gmm_align_equal", s_equal)

    def s_acc():
        from kalpy.gmm.train import GmmStatsAccumulator
        from kalpy.utils import generate_write_specifier  # noqa: F401
        acc = GmmStatsAccumulator(mdl)
        like = acc.gmm_accs.acc_stats(acc.acoustic_model, acc.transition_model, st["ali"].alignment, st["feats"])
        acc.transition_model.acc_stats(st["ali"].alignment, acc.transition_accs)
        r = {"acc_like": np.asarray([like], np.float64), "acc_trans": np_of(acc.transition_accs)}
        path = os.path.join(tmp, "1.acc")
        try:   # the GMM accumulators only leave kalpy through Kaldi's writer: gmm-acc-stats format, parsed by kaldi_io.read_gmm_accs
            from _kalpy.util import Output
            ko = Output(path, True)
            acc.transition_accs.Write(ko.Stream(), True)
            acc.gmm_accs.Write(ko.Stream(), True)
            ko.Close()
            r["acc_file"] = np.frombuffer(open(path, "rb").read(), np.uint8)
        except Exception:
            errors["acc_file"] = traceback.format_exc(limit=3)
        return r
    stage("acc_stats", s_acc)

    out["errors"] = np.frombuffer(json.dumps(errors).encode(), np.uint8)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "stages failed:", sorted(errors) or "none")


if __name__ == "__main__":
    main()
