"""The drop-in boundary, literally: every ``from kalpy... import ...`` of the reference's hot-path modules resolves
against the shim package (montreal-forced-aligner_b200/shim/kalpy), and every call shape of SURVEY.md section 8(b) -- the constructor
and method signatures MFA's own code uses, copied from the cited call sites -- binds to the shim's classes (inspect.signature).
No GPU: nothing is constructed that needs a device."""
import importlib
import inspect
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "kalpy_imports.json")


@pytest.fixture(scope="module")
def shim():
    import mfa_b200
    real = sys.modules.get("kalpy")
    if real is not None and not getattr(real, "__mfa_b200_shim__", False):
        pytest.skip("a real kalpy is imported in this process")
    path = mfa_b200.install_kalpy_shim()
    yield path
    sys.path.remove(path)
    for k in [k for k in sys.modules if k == "kalpy" or k.startswith("kalpy.")]:
        del sys.modules[k]


def test_golden_import_list_is_current():
    """Where the reference tree is present (the build container) the committed list is re-derived from it."""
    if not os.path.isdir("/root/reference/montreal_forced_aligner"):
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import extract_kalpy_imports as X
        assert X.extract() == json.load(open(GOLD))
    finally:
        sys.path.pop(0)


def test_every_kalpy_import_of_the_hot_path_modules_resolves(shim):
    wanted = json.load(open(GOLD))
    assert len(wanted) >= 10
    n = 0
    for rel, rows in wanted.items():
        for module, names, lineno in rows:
            mod = importlib.import_module(module)
            assert getattr(importlib.import_module("kalpy"), "__mfa_b200_shim__", False)
            for name in names:
                assert hasattr(mod, name), f"{rel}:{lineno}: from {module} import {name}"
                n += 1
    assert n >= 70


def _binds(fn, *args, **kwargs):
    sig = inspect.signature(fn)
    try:
        sig.bind(*args, **kwargs)
    except TypeError as ex:
        raise AssertionError(f"{getattr(fn, '__qualname__', fn)}{sig} does not accept args={args} kwargs={sorted(kwargs)}: {ex}")


def test_call_shapes_of_the_reference_bind(shim):
    """Argument lists as written at the reference's call sites (file:line in each comment); values are placeholders."""
    from kalpy.data import KaldiMapping, MatrixArchive, Segment
    from kalpy.decoder.data import FstArchive
    from kalpy.decoder.training_graphs import TrainingGraphCompiler
    from kalpy.feat.cmvn import CmvnComputer
    from kalpy.feat.data import FeatureArchive
    from kalpy.feat.fmllr import FmllrComputer
    from kalpy.feat.mfcc import MfccComputer
    from kalpy.fstext.lexicon import LexiconCompiler, Pronunciation
    from kalpy.gmm.align import GmmAligner
    from kalpy.gmm.data import Alignment, AlignmentArchive
    from kalpy.gmm.train import GmmStatsAccumulator
    from kalpy.gmm.utils import read_gmm_model, read_topology, read_transition_model, read_tree, write_gmm_model
    from kalpy.utils import generate_read_specifier, generate_write_specifier, read_kaldi_object
    from kalpy.utterance import Utterance as KalpyUtterance
    P = "x"
    # corpus/features.py:780-820 FeatureConfigMixin.mfcc_options -> MfccComputer(**opts) (features.py:685, models.py:515)
    mfcc_options = dict(use_energy=False, dither=0.0, energy_floor=0.0, num_coefficients=13, num_mel_bins=23, cepstral_lifter=22,
                        preemphasis_coefficient=0.97, frame_shift=10, frame_length=25, low_frequency=20, high_frequency=7800,
                        sample_frequency=16000, allow_downsample=True, allow_upsample=True, snip_edges=True)
    _binds(MfccComputer.__init__, None, **mfcc_options)
    _binds(MfccComputer.compute_mfccs_for_export, None, object(), compress=True)                  # corpus/features.py:235
    _binds(Segment, P, 0.0, 1.0, 0)                                                               # corpus/features.py:232
    _binds(KalpyUtterance, object(), "text")                                                      # command_line/align_one.py:163-166
    _binds(KalpyUtterance.generate_mfccs, None, object())                                         # align_one.py:167
    _binds(KalpyUtterance.apply_cmvn, None, object())                                             # align_one.py:183
    _binds(KalpyUtterance.generate_features, None, object(), None, lda_mat=None, fmllr_trans=None)   # online/alignment.py:82-94
    _binds(CmvnComputer.export_cmvn, None, P, object(), object(), write_scp=True)                 # corpus/acoustic_corpus.py:1337
    _binds(CmvnComputer.compute_cmvn_from_features, None, [object()])                             # align_one.py:168
    _binds(FeatureArchive.__init__, None, P, utt2spk=None, lda_mat_file_name=None, transform_file_name=None, vad_file_name=None,
           deltas=True, splices=True, splice_frames=3)                                            # db.py:2127-2135 (+ **kwargs of its callers)
    _binds(FeatureArchive.__init__, None, P, utt2spk=None, cmvn_file_name=P, vad_file_name=None, subsample_n=0)   # corpus/features.py:323-329
    _binds(FeatureArchive.__init__, None, P, utt2spk=None, vad_file_name=None, subsample_n=0, use_sliding_cmvn=True)   # features.py:331-337
    _binds(FeatureArchive.__init__, None, P, deltas=True)                                         # acoustic_modeling/monophone.py:89-92
    for attr in ("__iter__", "__getitem__", "close"):
        assert hasattr(FeatureArchive, attr)
    _binds(TrainingGraphCompiler.__init__, None, P, P, object(), use_g2p=False, batch_size=500)   # alignment/multiprocessing.py:537-545
    _binds(TrainingGraphCompiler.export_graphs, None, P, [("k", "t")], interjection_words=None, callback=print)   # multiprocessing.py:565-571
    _binds(TrainingGraphCompiler.compile_fst, None, "text")                                       # online/alignment.py:96
    _binds(FstArchive.__init__, None, P)                                                          # multiprocessing.py:831
    _binds(GmmAligner.__init__, None, P, transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40,
           disambiguation_symbols=None)                                                           # multiprocessing.py:814 + mixins.py:192-203
    _binds(GmmAligner.boost_silence, None, 1.0, [1, 2])                                           # multiprocessing.py:815
    _binds(GmmAligner.align_utterance, None, object(), object())                                  # online/alignment.py:107
    _binds(GmmAligner.export_alignments, None, P, object(), object(), word_file_name=P, likelihood_file_name=P, callback=print)   # :846-853
    for attr in ("acoustic_model_path", "transition_model", "acoustic_scale", "beam"):
        assert attr in inspect.getsource(GmmAligner.__init__)
    _binds(Alignment.generate_ctm, None, object(), object(), 0.01)                                # multiprocessing.py:1316-1320
    _binds(AlignmentArchive.__init__, None, P, words_file_name=P, likelihood_file_name=P)         # multiprocessing.py:1729-1731
    _binds(GmmStatsAccumulator.__init__, None, P)                                                 # multiprocessing.py:652
    _binds(GmmStatsAccumulator.accumulate_stats, None, object(), object(), callback=print)        # multiprocessing.py:658-662
    _binds(FmllrComputer.__init__, None, P, P, [1], spk2utt={}, fmllr_update_type="full", silence_weight=0.0, acoustic_scale=0.1)   # features.py:506-512
    _binds(FmllrComputer.export_transforms, None, P, object(), object(), previous_transform_archive=None, callback=print)   # features.py:521-527
    _binds(MatrixArchive.__init__, None, P)                                                       # corpus/features.py:494
    _binds(read_gmm_model, P); _binds(write_gmm_model, P, object(), object())                     # monophone.py:257,296
    _binds(read_topology, P); _binds(read_tree, P); _binds(read_transition_model, P)
    _binds(read_kaldi_object, object, P)                                                          # multiprocessing.py:1218
    _binds(generate_read_specifier, P); _binds(generate_write_specifier, P, write_scp=True)       # features.py:207,317
    m = KaldiMapping(list_mapping=True)                                                           # corpus/features.py:298
    m["spk"] = ["u1", "u2"]
    assert hasattr(m, "load") and "spk" in m
    # models.py:495-504
    lc = LexiconCompiler(silence_probability=0.5, initial_silence_probability=0.5, final_silence_correction=None,
                         final_non_silence_correction=None, silence_phone="sil", oov_phone="sil", position_dependent_phones=False,
                         phones={"a", "b"})
    lc.add_pronunciation(Pronunciation("ab", "a b", None, None, None, None, None))                # online/alignment.py:56-66
    assert lc.word_table.member("ab") and not lc.word_table.member("zz")                          # online/alignment.py:53
    assert lc.to_int("ab zz") == [lc.word_table.find("ab"), lc.word_table.find("<unk>")]
    assert lc.phone_table.find("a") > 0 and lc.silence_symbols == [lc.phone_table.find("sil")]    # online/alignment.py:106
    for attr in ("phones_to_pronunciations", "load_pronunciations", "clear"):
        assert hasattr(lc, attr)


def test_out_of_scope_names_import_but_refuse_construction(shim):
    from kalpy.feat.pitch import PitchComputer
    from kalpy.feat.vad import VadComputer
    from kalpy.gmm.data import TranscriptionArchive
    from kalpy.gmm.train import TwoFeatsStatsAccumulator
    from kalpy.ivector.extractor import IvectorExtractor
    from mfa_b200._lib import MfaError
    for cls in (PitchComputer, VadComputer, TranscriptionArchive, TwoFeatsStatsAccumulator, IvectorExtractor):
        with pytest.raises(MfaError, match="outside the alignment hot path"):
            cls()
