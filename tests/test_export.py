"""Row N4 on the host: alignment -> phone CTM -> word intervals -> TextGrid / json / csv, and the TextGrid reader on the reference's
own fixtures (short and long format).  Alignments come from mfa_equal_align so no GPU is needed."""
import json

import numpy as np
import pytest

from helpers import gold
from mfa_b200 import engine as E, export as X, kaldi_io as K, kalpy_compat as KC, lexicon as LX, synth as SY


def _setup(position_dependent):
    rng = np.random.default_rng(11)
    phones = ["a", "b", "c", "d", "e"]
    pt = LX.make_phone_table(phones, ("sil", "spn"), position_dependent)
    prons = {"ab": [LX.Pron(["a", "b"]), LX.Pron(["a", "b", "c"], 0.5)], "cab": [LX.Pron(["c", "a", "b"])], "d": [LX.Pron(["d"])],
             "eda": [LX.Pron(["e", "d", "a"]), LX.Pron(["e", "a"], 0.3)]}
    lex = LX.Lexicon(prons, pt, silence_probability=0.5, initial_silence_probability=0.5, position_dependent_phones=position_dependent)
    topo = SY.make_topology(pt)
    tree, n_pdfs = SY.make_tree(rng, topo, False, 30)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    return lex, tm, tree


@pytest.mark.parametrize("position_dependent", [False, True])
def test_alignment_to_word_and_phone_intervals(position_dependent, tmp_path):
    lex, tm, tree = _setup(position_dependent)
    text = "ab zzz cab d eda ab"                     # zzz is out of vocabulary -> <unk> -> spn
    wids = lex.to_int(text)
    assert wids[1] == lex.word_table["<unk>"]
    batch = E.GraphCompiler(tm, tree, lex).compile([wids])
    fst = batch.export()[0]
    T = 437
    for seed in range(6):                            # different random paths: with / without optional silences, both pronunciations
        ali, words = KC.gmm_align_equal_batch(["k"], [fst], [T], seeds=[seed])[0]
        assert words == wids
        a = KC.Alignment("0-7", ali, words, -1234.5, np.full(T, -3.0, np.float32))
        ctm = X.alignment_to_ctm(a, tm, lex, 0.01, begin=2.0, end=2.0 + T * 0.01, text=text)
        real = [w for w in ctm.word_intervals if w.label != "<eps>"]
        assert [w.label for w in real] == text.split()          # <unk> got its original word back
        ph = ctm.phone_intervals
        assert ph[0].begin == 2.0 and abs(ph[-1].end - (2.0 + T * 0.01)) < 1e-6
        assert all(abs(x.end - y.begin) < 1e-9 for x, y in zip(ph, ph[1:]))          # phones tile the utterance
        assert all(w.begin == w.phones[0].begin and w.end == w.phones[-1].end for w in ctm.word_intervals)
        assert all(not p.label.endswith(("_B", "_E", "_I", "_S")) for p in ph)        # position markers stripped
        assert real[1].pronunciation == "spn" and real[3].pronunciation == "d"
        assert real[0].pronunciation in ("a b", "a b c") and real[4].pronunciation in ("e d a", "e a")
        assert all(w.pronunciation == "sil" for w in ctm.word_intervals if w.label == "<eps>")
    # export: all four formats; silences are blank; last interval snapped to the file duration
    dur = 2.0 + T * 0.01 + 0.013
    data = X.ctm_to_speaker_data(ctm, "spk1", lex)
    for fmt, name in (("long_textgrid", "l.TextGrid"), ("short_textgrid", "s.TextGrid")):
        assert X.export_textgrid(data, tmp_path / name, dur, 0.01, fmt)
        tiers = X.read_textgrid(tmp_path / name)
        assert list(tiers) == ["words", "phones"]
        for ent in tiers.values():
            assert ent[0][0] == 0.0 and abs(ent[-1][1] - round(dur, 6)) < 1e-9
            assert all(abs(x[1] - y[0]) < 1e-9 for x, y in zip(ent, ent[1:]))         # blanks filled: the tier tiles [0, duration]
        assert [e[2] for e in tiers["words"] if e[2]] == text.split()
        assert "sil" not in [e[2] for e in tiers["phones"]] and tiers["words"][0][2] == ""   # [0, 2.0) precedes the utterance
    assert X.export_textgrid(data, tmp_path / "o.json", dur, 0.01, "json")
    js = json.load(open(tmp_path / "o.json"))
    assert js["end"] == round(dur, 6) and [e[2] for e in js["tiers"]["words"]["entries"]] == text.split()
    assert X.export_textgrid(data, tmp_path / "o.csv", dur, 0.01, "csv")
    rows = open(tmp_path / "o.csv").read().strip().splitlines()
    assert rows[0] == "Begin,End,Label,Type,Speaker" and len(rows) == 1 + len(data["spk1"]["words"]) + len(data["spk1"]["phones"])
    # two speakers in one file -> "speaker - tier" names; no intervals -> nothing written
    two = dict(data); two.update(X.ctm_to_speaker_data(ctm, "spk2", lex))
    X.export_textgrid(two, tmp_path / "two.TextGrid", dur, 0.01)
    assert list(X.read_textgrid(tmp_path / "two.TextGrid")) == ["spk1 - words", "spk1 - phones", "spk2 - words", "spk2 - phones"]
    assert not X.export_textgrid({"s": {"words": [], "phones": []}}, tmp_path / "none.TextGrid", dur, 0.01)
    assert not (tmp_path / "none.TextGrid").exists()


def test_phones_to_pronunciations_rejects_foreign_phone_sequences():
    lex, tm, tree = _setup(False)
    iv = [KC.CtmInterval(0.0, 0.1, "a"), KC.CtmInterval(0.1, 0.2, "e")]
    with pytest.raises(ValueError):
        X.phones_to_pronunciations(lex, [lex.word_table["ab"]], iv)


def test_textgrid_reader_on_reference_fixtures(tmp_path):
    g = gold()
    p = tmp_path / "ref_short.TextGrid"; p.write_bytes(bytes(g["acoustic_corpus_textgrid"]))
    t = X.read_textgrid(p)
    assert list(t) == ["words", "phones"] and len(t["words"]) == 71
    assert t["words"][1] == (1.05, 1.2, "this") and abs(t["words"][-1][1] - 26.72325) < 1e-9
    lab = bytes(g["acoustic_corpus_lab"]).decode().strip().lower().split()
    assert [w for _, _, w in t["words"] if w][:8] == lab[:8]
    q = tmp_path / "ref_long.TextGrid"; q.write_bytes(bytes(g["long_textgrid_fixture"]))
    t2 = X.read_textgrid(q)
    assert list(t2)[0] == "michael" and len(t2["michael"]) == 7 and abs(t2["michael"][0][1] - 1.059222833923831) < 1e-12
    # writer -> reader round trip of the reference tiers, both flavours
    for short in (False, True):
        X.write_textgrid(tmp_path / "rt.TextGrid", t, 26.72325, short=short)
        assert X.read_textgrid(tmp_path / "rt.TextGrid") == t
