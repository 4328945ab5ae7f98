"""kalpy-shaped Python surface over the B200 engine (the drop-in boundary of SURVEY.md section 8b).

The reference reaches its hot path only through kalpy objects; these classes keep the names, argument meaning and error
behaviour MFA relies on (call sites cited per class, paths relative to /root/reference/montreal_forced_aligner), and run
the arithmetic on the GPU through the C ABI.  Matrices cross the boundary as numpy arrays (kalpy ``FloatMatrix`` role);
files are Kaldi ark/scp.  There is no CPU fallback: constructing an engine without a B200 raises ``MfaError``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import threading

import numpy as np

from . import engine as E, kaldi_io as K
from ._lib import MfaError
from .gmm_update import AccumAmDiagGmm, mle_update
from .lexicon import Lexicon

_engines: Dict[int, E.Engine] = {}            # every engine this module created, by id (introspection / tests)
_idle: Dict[int, List[E.Engine]] = {}         # per device: engines whose thread has ended, waiting for the next thread
_engines_lock = threading.Lock()
_tls = threading.local()


class _Lease:
    """Ties an engine to the thread that holds this object in its thread-local storage; when the thread ends the storage is dropped, the
    lease is collected and the engine goes back to the idle list instead of being destroyed (an engine owns pinned staging buffers and
    device work space: creating one per short-lived job thread would cost more than the jobs)."""

    def __init__(self, device: int, eng: E.Engine):
        self.device, self.eng = device, eng

    def __del__(self):
        try:
            with _engines_lock:
                _idle.setdefault(self.device, []).append(self.eng)
        except Exception:   # interpreter shutdown
            pass


def get_engine(device: Optional[int] = None) -> E.Engine:
    """One engine per device and per host THREAD.  MFA runs one aligner / accumulator per job and, with USE_THREADING, several jobs as
    threads of one process (native code must be re-entrant): an engine owns one main stream and its scratch buffers and is not
    re-entrant, so every job thread gets its own; the objects of this module remember the engine they were built with.  Engines of ended
    threads are handed to the next new thread (a stage's thread pool comes and goes with the stage).  Jobs map to GPUs via LOCAL_RANK /
    MFA_B200_DEVICE."""
    if device is None:
        device = int(os.environ.get("MFA_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    leases = getattr(_tls, "leases", None)
    if leases is None:
        leases = _tls.leases = {}
    lease = leases.get(device)
    if lease is not None:
        return lease.eng
    with _engines_lock:
        pool = _idle.get(device)
        eng = pool.pop() if pool else None
    if eng is None:
        eng = E.Engine(device)
        with _engines_lock:
            _engines[id(eng)] = eng
    else:
        eng.sync()   # whatever its previous thread left in flight
    leases[device] = _Lease(device, eng)
    return eng


read_gmm_model = K.read_gmm_model
write_gmm_model = K.write_gmm_model


import logging

kalpy_logger = logging.getLogger("kalpy")   # kalpy.utils.kalpy_logger (acoustic_modeling/monophone.py:15, corpus/acoustic_corpus.py:20)


def read_topology(path):
    """kalpy.gmm.utils.read_topology: Kaldi text topology (`topo`)."""
    return K.read_topology_text(path)


def read_tree(path):
    return K.read_tree(path)


def read_transition_model(path):
    """kalpy.gmm.utils.read_transition_model (alignment/base.py:27): the TransitionModel at the head of a .mdl."""
    return K.read_gmm_model(path)[0]


class FloatMatrix(np.ndarray):
    """numpy array carrying the three Kaldi matrix methods MFA's job functions call on what the archives hand back
    (``feats.NumRows()`` acoustic_modeling/monophone.py:100, ``mfccs.NumRows()`` corpus/features.py:353)."""

    def NumRows(self) -> int:
        return int(self.shape[0])

    def NumCols(self) -> int:
        return int(self.shape[1]) if self.ndim > 1 else 0

    def numpy(self) -> np.ndarray:
        return np.asarray(self)


def as_matrix(a) -> FloatMatrix:
    return np.asarray(a).view(FloatMatrix)


class KaldiMapping(dict):
    """kalpy.data.KaldiMapping (corpus/features.py:298-305,492; db.py:2114): utt2spk / spk2utt text maps (``key value...`` per line)."""

    def __init__(self, list_mapping: bool = False):
        super().__init__()
        self.list_mapping = list_mapping

    def load(self, file_name):
        with open(file_name, "r", encoding="utf8") as f:
            for line in f:
                parts = line.split()
                if not parts:
                    continue
                self[parts[0]] = parts[1:] if self.list_mapping else (parts[1] if len(parts) > 1 else "")

    def export(self, file_name, skip_safe: bool = False):
        with open(file_name, "w", encoding="utf8") as f:
            for k in sorted(self):
                v = self[k]
                f.write(f"{k} {' '.join(str(x) for x in v) if self.list_mapping else v}\n")


def _outside(name: str, what: str):
    class _Outside:
        __doc__ = f"kalpy {name}: {what} -- outside the alignment hot path this engine replaces (SURVEY.md section 8); importable, not constructible."

        def __init__(self, *a, **k):
            raise MfaError(f"{name}: {what} is outside the alignment hot path this engine replaces (SURVEY.md section 8)")
    _Outside.__name__ = _Outside.__qualname__ = name
    return _Outside


PitchComputer = _outside("PitchComputer", "pitch features")
VadComputer = _outside("VadComputer", "voice activity detection")
IvectorExtractor = _outside("IvectorExtractor", "i-vector extraction")
TranscriptionArchive = _outside("TranscriptionArchive", "lattice / transcription archives")
TwoFeatsStatsAccumulator = _outside("TwoFeatsStatsAccumulator", "SAT two-feature accumulation (train_sat model surgery)")


# ------------------------------------------------------------------------------------------------ audio / MFCC
@dataclass
class Segment:
    """kalpy.data.Segment (corpus/features.py:232; command_line/align_one.py:163)."""
    file_name: str
    begin: Optional[float] = None
    end: Optional[float] = None
    channel: int = 0

    def load_audio(self) -> np.ndarray:
        pcm, sr = K.read_wav_int16(self.file_name, self.channel or 0)
        if sr != 16000:
            raise MfaError(f"{self.file_name}: {sr} Hz audio; resampling is out of scope (SURVEY.md A.1), provide 16 kHz WAV")
        a = 0 if self.begin is None else int(round(self.begin * sr))
        b = pcm.shape[0] if self.end is None else int(round(self.end * sr))
        return pcm[max(a, 0):min(b, pcm.shape[0])]


@dataclass
class KalpyUtterance:
    """kalpy.utterance.Utterance as MFA drives it in align_one (command_line/align_one.py:163-183) and align_utterance_online
    (online/alignment.py:44,82-94,122): a segment of a sound file with its transcript, raw MFCCs, and the feature chain
    CMVN -> deltas | splice+LDA -> fMLLR (order of alignment/multiprocessing.py:1287-1304)."""
    segment: Segment
    transcript: str = ""
    cmvn_string: Optional[str] = None
    fmllr_string: Optional[str] = None
    mfccs: Optional[np.ndarray] = None
    _cmvn: Optional[np.ndarray] = None

    def generate_mfccs(self, mfcc_computer: "MfccComputer"):
        self.mfccs = mfcc_computer.compute_mfccs(self.segment)
        self._cmvn = None
        return self.mfccs

    def apply_cmvn(self, cmvn: np.ndarray):
        """Records the (file / speaker level) statistics; the subtraction is fused into the feature kernel (norm_vars = false)."""
        self._cmvn = np.asarray(cmvn, np.float64)

    def generate_features(self, mfcc_computer: "MfccComputer", pitch_computer=None, lda_mat: Optional[np.ndarray] = None,
                          fmllr_trans: Optional[np.ndarray] = None, uses_deltas: bool = True) -> np.ndarray:
        if pitch_computer is not None:
            raise MfaError("pitch features are outside the hot path (SURVEY.md section 2a)")
        if self.mfccs is None:
            self.generate_mfccs(mfcc_computer)
        fo = np.asarray([0, self.mfccs.shape[0]], np.int64)
        mode = "lda" if lda_mat is not None else ("deltas" if uses_deltas else "none")
        fm = None if fmllr_trans is None else np.asarray(fmllr_trans, np.float32)[None]
        cm = None if self._cmvn is None else self._cmvn[None]
        return get_engine().features(self.mfccs, fo, mode, lda=lda_mat, fmllr=fm, cmvn_stats=cm, utt2spk=np.zeros(1, np.int32), n_spk=1)


def generate_read_specifier(path) -> str:
    """kalpy.utils.generate_read_specifier: 'scp:...' / 'ark:...' by extension (the archives of this module take plain paths too)."""
    path = str(path)
    return ("scp:" if path.endswith(".scp") else "ark:") + path


def generate_write_specifier(path, write_scp: bool = False) -> str:
    path = str(path)
    return f"ark,scp:{path},{path[:-4]}.scp" if write_scp else f"ark:{path}"


def read_kaldi_object(obj_type, path):
    """kalpy.utils.read_kaldi_object for the types the hot path reads: TransitionModel / AmDiagGmm (from a .mdl), ContextDependency
    (tree), FloatMatrix (lda.mat)."""
    name = getattr(obj_type, "__name__", str(obj_type))
    if name == "TransitionModel":
        return K.read_gmm_model(path)[0]
    if name == "AmDiagGmm":
        return K.read_gmm_model(path)[1]
    if name == "ContextDependency":
        return K.read_tree(path)
    return K.read_matrix_file(path)


class CompressedMatrix:
    """Kaldi CompressedMatrix value: holds the codec bytes; ``numpy()`` decodes (corpus/features.py:209,318,356)."""

    def __init__(self, mat: np.ndarray):
        self.blob = K.compress_matrix(mat)
        self.shape = mat.shape

    def numpy(self) -> np.ndarray:
        return K.decompress_matrix(self.blob)


class MfccComputer:
    """kalpy.feat.mfcc.MfccComputer (corpus/features.py:685,198,235; alignment/multiprocessing.py:1284)."""

    def __init__(self, **mfcc_options):
        self.parameters = dict(mfcc_options)
        self.opts = E.mfcc_opts(**{k: v for k, v in mfcc_options.items() if k not in ("uses_cmvn", "use_pitch")})
        self.frame_shift = self.opts.frame_shift_ms

    def _eng(self):
        return get_engine()   # the calling thread's engine: MFA hands ONE MfccComputer to all jobs (MfccArguments), which may be threads

    def compute_mfccs_batch(self, pcm_list: Sequence[np.ndarray]) -> List[np.ndarray]:
        off = np.zeros(len(pcm_list) + 1, np.int64)
        off[1:] = np.cumsum([len(p) for p in pcm_list])
        pcm = np.concatenate(pcm_list).astype(np.int16) if len(pcm_list) else np.zeros(0, np.int16)
        out, fo = self._eng().mfcc(pcm, off, self.opts)
        return [out[fo[i]:fo[i + 1]] for i in range(len(pcm_list))]

    def compute_mfccs(self, segment) -> np.ndarray:
        pcm = segment.load_audio() if isinstance(segment, Segment) else np.asarray(segment, dtype=np.int16)
        return self.compute_mfccs_batch([pcm])[0]

    def compute_mfccs_for_export(self, segment, compress: bool = True):
        m = self.compute_mfccs(segment)
        return CompressedMatrix(m) if compress else m


class CmvnComputer:
    """kalpy.feat.cmvn.CmvnComputer (corpus/acoustic_corpus.py:1336-1337; command_line/align_one.py:161,168)."""

    def compute_cmvn_from_features(self, feats: Sequence[np.ndarray]) -> np.ndarray:
        feats = [np.asarray(f.numpy() if isinstance(f, CompressedMatrix) else f, dtype=np.float32) for f in feats]
        fo = np.zeros(len(feats) + 1, np.int64)
        fo[1:] = np.cumsum([f.shape[0] for f in feats])
        stats = get_engine().cmvn_stats(np.concatenate(feats), fo, np.zeros(len(feats), np.int32), 1)
        return stats[0]

    def export_cmvn(self, file_name, feature_archive: "FeatureArchive", spk2utt: Dict[str, List[str]], write_scp: bool = True):
        """Per-speaker stats over the raw features -> cmvn.ark (+ cmvn.scp), DoubleMatrix 2 x (D+1) per speaker."""
        spks = list(spk2utt)
        mats, u2s = [], []
        for si, s in enumerate(spks):
            for u in spk2utt[s]:
                mats.append(feature_archive.raw(u))
                u2s.append(si)
        fo = np.zeros(len(mats) + 1, np.int64)
        fo[1:] = np.cumsum([m.shape[0] for m in mats])
        stats = get_engine().cmvn_stats(np.concatenate(mats), fo, np.asarray(u2s, np.int32), len(spks))
        scp = str(file_name)[:-4] + ".scp" if write_scp else None
        with K.ArkWriter(file_name, scp) as w:
            for si, s in enumerate(spks):
                w.write_matrix(str(s), stats[si].astype(np.float64))
        return {s: stats[i] for i, s in enumerate(spks)}


def _read_table(path, kind) -> Dict[str, object]:
    path = str(path)
    if path.endswith(".scp"):
        return {k: K.read_scp_object(p, off, kind) for k, p, off in K.read_scp(path)}
    return dict(K.read_ark(path, kind))


def _read_map(path) -> Dict[str, str]:
    out = {}
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2:
                out[parts[0]] = parts[1]
    return out


class FeatureArchive:
    """kalpy.feat.data.FeatureArchive (db.py:2127-2135; corpus/features.py:323-339; acoustic_modeling/monophone.py:89-99).

    Lazily applies CMVN -> deltas | splice+LDA -> fMLLR (order of alignment/multiprocessing.py:1287-1304) on the GPU."""

    def __init__(self, file_name, utt2spk=None, cmvn_file_name=None, lda_mat_file_name=None, transform_file_name=None,
                 vad_file_name=None, deltas: bool = False, splices: bool = False, splice_frames: int = 3, subsample_n: int = 0,
                 use_sliding_cmvn: bool = False, utt2spk_file_name=None):
        """``utt2spk``: a KaldiMapping / dict (what db.py:2114-2129 passes) or the path of an utt2spk file."""
        if utt2spk is None:
            utt2spk = utt2spk_file_name
        if vad_file_name or subsample_n or use_sliding_cmvn:
            raise MfaError("vad / subsampling / sliding CMVN are outside the alignment hot path (SURVEY.md section 2a)")
        self.file_name = str(file_name)
        self._entries = K.read_scp(self.file_name) if self.file_name.endswith(".scp") else None
        self._ark = None if self._entries is not None else dict(K.read_ark(self.file_name, "matrix"))
        self.keys = [e[0] for e in self._entries] if self._entries is not None else list(self._ark)
        self._index = {e[0]: e for e in self._entries} if self._entries is not None else None
        self.utt2spk = {} if utt2spk is None else (dict(utt2spk) if isinstance(utt2spk, dict) else _read_map(utt2spk))
        self.cmvn_read_specifier = str(cmvn_file_name) if cmvn_file_name else None
        self._cmvn = {k: np.asarray(v, np.float64) for k, v in _read_table(cmvn_file_name, "matrix").items()} if cmvn_file_name else None
        self.lda_mat_file_name = str(lda_mat_file_name) if lda_mat_file_name else None
        self._lda = K.read_matrix_file(lda_mat_file_name).astype(np.float32) if lda_mat_file_name else None
        self.transform_read_specifier = str(transform_file_name) if transform_file_name else None
        self._trans = {k: np.asarray(v, np.float32) for k, v in _read_table(transform_file_name, "matrix").items()} if transform_file_name else None
        self.use_deltas, self.use_splices, self.splice_frames = bool(deltas), bool(splices) or self._lda is not None, splice_frames

    def raw(self, key: str) -> np.ndarray:
        if self._ark is not None:
            return np.asarray(self._ark[key], np.float32)
        _, p, off = self._index[key]
        return np.asarray(K.read_scp_object(p, off, "matrix"), np.float32)

    def batch(self, keys: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """Final features of `keys`, concatenated, + frame offsets: one fused kernel launch for the whole batch."""
        mats = [self.raw(k) for k in keys]
        fo = np.zeros(len(mats) + 1, np.int64)
        fo[1:] = np.cumsum([m.shape[0] for m in mats])
        if not mats:
            return np.zeros((0, 0), np.float32), fo
        spk_names = sorted({self.utt2spk.get(k, k) for k in keys})
        sidx = {s: i for i, s in enumerate(spk_names)}
        u2s = np.asarray([sidx[self.utt2spk.get(k, k)] for k in keys], np.int32)
        cmvn = fm = None
        dim = mats[0].shape[1]
        if self._cmvn is not None:
            cmvn = np.stack([self._cmvn[s] if s in self._cmvn else np.concatenate([np.zeros((2, dim)), np.ones((2, 1))], 1) for s in spk_names])
        mode = "lda" if self._lda is not None else ("deltas" if self.use_deltas else "none")
        if self._trans is not None:
            od = self._lda.shape[0] if self._lda is not None else (3 * dim if self.use_deltas else dim)
            fm = np.stack([self._trans.get(s, np.eye(od, od + 1, dtype=np.float32)) for s in spk_names])
        out = get_engine().features(np.concatenate(mats), fo, mode, lda=self._lda, splice_ctx=self.splice_frames, fmllr=fm, cmvn_stats=cmvn,
                                    utt2spk=u2s, n_spk=len(spk_names))
        return out, fo

    def __getitem__(self, key: str) -> np.ndarray:
        return as_matrix(self.batch([key])[0])

    def __iter__(self) -> Iterator[Tuple[str, np.ndarray]]:
        B = 256
        for i in range(0, len(self.keys), B):
            ks = self.keys[i:i + B]
            out, fo = self.batch(ks)
            for j, k in enumerate(ks):
                yield k, as_matrix(out[fo[j]:fo[j + 1]])

    def close(self):
        self._ark = None


# ------------------------------------------------------------------------------------------------ graphs
class TrainingGraphCompiler:
    """kalpy.decoder.training_graphs.TrainingGraphCompiler (alignment/multiprocessing.py:537-571,1189,1282;
    online/alignment.py:77-96).  ``lexicon_compiler`` is a mfa_b200.lexicon.Lexicon."""

    def __init__(self, model_path, tree_path, lexicon_compiler: Lexicon, use_g2p: bool = False, batch_size: int = 500):
        if use_g2p:
            raise MfaError("G2P-backed lexicons are outside the hot path (SURVEY.md section 2a)")
        self.transition_model, self.acoustic_model = K.read_gmm_model(model_path)
        self.tree = K.read_tree(tree_path)
        self.lexicon_compiler = lexicon_compiler
        self.batch_size = batch_size
        self._gc = E.GraphCompiler(self.transition_model, self.tree, getattr(lexicon_compiler, "lexicon", lexicon_compiler))

    def compile_fst(self, text: str) -> K.Fst:
        return self._gc.compile([self.lexicon_compiler.to_int(text)]).export()[0]

    def compile_batch(self, texts: Sequence[str], n_threads: int = 8) -> E.FstBatch:
        return self._gc.compile([self.lexicon_compiler.to_int(t) for t in texts], n_threads=n_threads)

    def export_graphs(self, file_name, records: Iterable[Tuple[str, str]], interjection_words=None, callback: Optional[Callable] = None):
        if interjection_words:
            raise MfaError("interjection-word graphs (transcript verification) are outside the hot path")
        records = list(records)
        with K.ArkWriter(file_name) as w:
            for i in range(0, len(records), self.batch_size):
                chunk = records[i:i + self.batch_size]
                fsts = self.compile_batch([t for _, t in chunk]).export()
                for (key, _), fst in zip(chunk, fsts):
                    w.write_fst(key, fst)
                    if callback:
                        callback(1)


class FstArchive:
    """kalpy.decoder.data.FstArchive (alignment/multiprocessing.py:831; acoustic_modeling/monophone.py:93-104)."""

    def __init__(self, file_name):
        self.file_name = str(file_name)
        self._fsts = dict(K.read_fst_ark(self.file_name))

    def __getitem__(self, key) -> K.Fst:
        return self._fsts[key]

    def __contains__(self, key):
        return key in self._fsts

    def __iter__(self):
        return iter(self._fsts.items())

    def keys(self):
        return list(self._fsts)

    def close(self):
        pass


# ------------------------------------------------------------------------------------------------ alignment
@dataclass
class CtmInterval:
    begin: float
    end: float
    label: object
    confidence: float = 0.0


class Alignment:
    """kalpy.gmm.data.Alignment (alignment/multiprocessing.py:1316-1320,1542-1546; online/alignment.py:113-121)."""

    def __init__(self, utterance_id, alignment, words, likelihood=None, per_frame_likelihoods=None):
        self.utterance_id = utterance_id
        self.alignment = np.asarray(alignment, np.int64).reshape(-1).tolist()   # plain ints, as kalpy returns them
        self.words = np.asarray(words, np.int64).reshape(-1).tolist()
        self.likelihood = likelihood
        self.per_frame_likelihoods = None if per_frame_likelihoods is None else np.asarray(per_frame_likelihoods, np.float32)

    def generate_ctm(self, transition_model: K.TransitionModel, phone_table: Dict[int, str], frame_shift: float = 0.01) -> List[CtmInterval]:
        """Phone intervals from transition-ids: a phone ends at a transition into its HMM's final state followed (reorder=true)
        by that state's trailing self-loops (Kaldi SplitToPhones)."""
        tids = np.asarray(self.alignment, np.int64)
        n = tids.shape[0]
        tm = transition_model
        idx = np.nonzero(np.asarray(tm.is_final_tid)[tids])[0] if n else np.zeros(0, np.int64)
        if idx.size == 0:
            return []
        # phone k ends behind its final transition idx[k] and the self-loops of that state which follow it: the first position after
        # idx[k] whose transition-id differs (the next final transition at the latest) -- one segmented minimum instead of a Python loop
        sl = np.asarray(tm.self_loop_tid)[np.asarray(tm.id2state)[tids[idx]]]
        starts = idx + 1
        seg = np.diff(np.append(starts, n + 1))
        tp = np.append(tids, -1)                                   # position n: always a mismatch
        cand = np.where(tp[starts[0]:] != np.repeat(sl, seg), np.arange(starts[0], n + 1), n)
        ends = np.minimum.reduceat(cand, starts - starts[0])
        begins = np.append(0, ends[:-1])
        phones = np.asarray(tm.tid2phone)[tids[idx]]
        if self.per_frame_likelihoods is not None:
            cs = np.append(0.0, np.cumsum(self.per_frame_likelihoods, dtype=np.float64))
            conf = (cs[ends] - cs[begins]) / (ends - begins)
        else:
            conf = np.zeros(idx.size)
        out: List[CtmInterval] = []
        for b, e, ph, cf in zip(begins.tolist(), ends.tolist(), phones.tolist(), conf.tolist()):
            out.append(CtmInterval(round(b * frame_shift, 4), round(e * frame_shift, 4), phone_table.get(ph, ph) if phone_table else ph, cf))
        return out


class GmmAligner:
    """kalpy.gmm.align.GmmAligner (alignment/multiprocessing.py:814-853,1204-1315; online/alignment.py:97-117)."""

    def __init__(self, acoustic_model_path, transition_scale: float = 1.0, acoustic_scale: float = 0.1, self_loop_scale: float = 0.1,
                 beam: float = 10, retry_beam: float = 40, disambiguation_symbols=None, careful: bool = False, gmm_impl: int = 0):
        self.acoustic_model_path = str(acoustic_model_path)
        self.transition_model, self.acoustic_model = K.read_gmm_model(acoustic_model_path)
        self.transition_scale, self.acoustic_scale, self.self_loop_scale = transition_scale, acoustic_scale, self_loop_scale
        self.beam, self.retry_beam = beam, retry_beam
        self.gmm_impl = gmm_impl
        self.engine = get_engine()
        self._dm = E.DeviceModel(self.engine, self.transition_model, self.acoustic_model)
        self.num_done = self.num_error = self.num_retry = 0
        self.total_likelihood = 0.0
        self.total_frames = 0

    def boost_silence(self, factor: float, silence_phones: Sequence[int]):
        """gmm-boost-silence: weights of every pdf reachable from a silence phone are scaled by `factor` (no renormalisation)."""
        tm = self.transition_model
        sil = set(int(p) for p in silence_phones)
        pdfs = set()
        for ph, _hs, fpdf, spdf in tm.tuples:
            if int(ph) in sil:
                pdfs.add(int(fpdf)); pdfs.add(int(spdf))
        self._dm.boost_pdfs(factor, sorted(pdfs))
        for j in sorted(pdfs):
            a, b = self.acoustic_model.offsets[j], self.acoustic_model.offsets[j + 1]
            self.acoustic_model.weights[a:b] *= factor
        self.acoustic_model.gconsts = self.acoustic_model.compute_gconsts()

    def _opts(self):
        return E.align_opts(self.acoustic_scale, self.beam, self.retry_beam)

    def align_batch(self, keys: Sequence[str], fsts: Sequence[K.Fst], feats: np.ndarray, frame_off: np.ndarray) -> List[Optional[Alignment]]:
        """K2 + K3 for a batch of utterances whose final features are concatenated in `feats`."""
        batch = E.FstBatch.from_fsts(list(fsts))
        graphs = E.Graphs(batch, self.transition_model, self.transition_scale, self.self_loop_scale)
        # features in, alignments out: the log-likelihoods (only the pdfs each utterance's graph references) never leave the GPU
        res = E.align_feats(self.engine, self._dm, graphs, np.ascontiguousarray(feats, np.float32), frame_off, self._opts(), gmm_impl=self.gmm_impl)
        out: List[Optional[Alignment]] = []
        for u, k in enumerate(keys):
            r = res.utterance(u)
            if r["status"] >= 2:
                self.num_error += 1
                out.append(None)
                continue
            if r["status"] == 1:
                self.num_retry += 1
            self.num_done += 1
            self.total_likelihood += r["like"]
            self.total_frames += len(r["ali"])
            out.append(Alignment(k, r["ali"], r["words"], r["like"], r["per_frame"]))
        graphs.close(); batch.close()
        return out

    def align_utterance(self, training_graph: K.Fst, features: np.ndarray, utterance_id: Optional[str] = None) -> Optional[Alignment]:
        features = np.ascontiguousarray(features, np.float32)
        fo = np.asarray([0, features.shape[0]], np.int64)
        return self.align_batch([utterance_id], [training_graph], features, fo)[0]

    def export_alignments(self, file_name, training_graph_archive: FstArchive, feature_archive: FeatureArchive, word_file_name=None,
                          likelihood_file_name=None, callback: Optional[Callable] = None, batch_size: int = 512):
        """ali.ark (+ words.ark, likelihoods.ark); failed utterances are skipped; callback gets (utt_id, log-likelihood)."""
        wa = K.ArkWriter(file_name)
        ww = K.ArkWriter(word_file_name) if word_file_name else None
        wl = K.ArkWriter(likelihood_file_name) if likelihood_file_name else None
        keys = [k for k in feature_archive.keys if k in training_graph_archive]
        try:
            for i in range(0, len(keys), batch_size):
                ks = keys[i:i + batch_size]
                feats, fo = feature_archive.batch(ks)
                alis = self.align_batch(ks, [training_graph_archive[k] for k in ks], feats, fo)
                for k, a in zip(ks, alis):
                    if a is None:
                        continue
                    wa.write_int_vector(k, np.asarray(a.alignment, np.int32))
                    if ww:
                        ww.write_int_vector(k, np.asarray(a.words, np.int32))
                    if wl:
                        wl.write_vector(k, a.per_frame_likelihoods)
                    if callback:
                        callback((k, a.likelihood))
        finally:
            wa.close()
            if ww:
                ww.close()
            if wl:
                wl.close()


class AlignmentArchive:
    """kalpy.gmm.data.AlignmentArchive (alignment/multiprocessing.py:657,1540,1729-1731,1809)."""

    def __init__(self, file_name, words_file_name=None, likelihood_file_name=None):
        self._ali = dict(K.read_ark(str(file_name), "int_vector"))
        self._words = dict(K.read_ark(str(words_file_name), "int_vector")) if words_file_name else {}
        self._likes = dict(K.read_ark(str(likelihood_file_name), "vector")) if likelihood_file_name else {}

    def __getitem__(self, key) -> Alignment:
        pf = self._likes.get(key)
        return Alignment(key, self._ali[key], self._words.get(key, []), None if pf is None else float(np.sum(pf)), pf)

    def __contains__(self, key):
        return key in self._ali

    def __iter__(self):
        for k in self._ali:
            yield self[k]

    def keys(self):
        return list(self._ali)

    def close(self):
        pass


def string_hash(key: str) -> int:
    """Kaldi StringHasher (util/stl-utils.h): the srand() seed align-equal-compiled derives from the utterance id."""
    h = 0
    for ch in key.encode("utf8"):
        h = (h * 7853 + ch) & 0xFFFFFFFFFFFFFFFF
    return h & 0xFFFFFFFF


def gmm_align_equal_batch(keys: Sequence[str], fsts: Sequence[K.Fst], num_frames: Sequence[int], seeds: Optional[Sequence[int]] = None,
                          num_retries: int = 10) -> List[Optional[Tuple[List[int], List[int]]]]:
    """Equal alignment of a batch of training graphs (mfa_equal_align).  None where Kaldi's EqualAlign fails."""
    batch = E.FstBatch.from_fsts(list(fsts))
    fo = np.zeros(len(fsts) + 1, np.int64)
    fo[1:] = np.cumsum(np.asarray(num_frames, np.int64))
    if seeds is None:
        seeds = [string_hash(k if k is not None else "") for k in keys]
    counts = np.asarray([int(np.count_nonzero(f.arc_olabel)) for f in fsts], np.int64)
    ali, words, wo, nw, st = batch.equal_align(fo, seeds, num_retries, counts)
    batch.close()
    out = []
    for u in range(len(fsts)):
        if st[u] != 0:
            out.append(None)
        else:
            out.append(([int(x) for x in ali[fo[u]:fo[u + 1]]], [int(x) for x in words[wo[u]:wo[u] + nw[u]]]))
    return out


def gmm_align_equal(decode_fst: K.Fst, feats: np.ndarray, utterance_id: Optional[str] = None):
    """kalpy.gmm.align.gmm_align_equal (acoustic_modeling/monophone.py:108): -> (alignment, words) or (None, None)."""
    r = gmm_align_equal_batch([utterance_id], [decode_fst], [feats.shape[0]])[0]
    return (None, None) if r is None else r


# ------------------------------------------------------------------------------------------------ statistics
class GmmStatsAccumulator:
    """kalpy.gmm.train.GmmStatsAccumulator (alignment/multiprocessing.py:652-666; acoustic_modeling/monophone.py:84,114-120)."""

    def __init__(self, acoustic_model_path):
        self.transition_model, self.acoustic_model = K.read_gmm_model(acoustic_model_path)
        self.engine = get_engine()
        self._dm = E.DeviceModel(self.engine, self.transition_model, self.acoustic_model)
        self._dm.acc_zero()
        self.transition_accs = self.transition_model.InitStats()
        self.gmm_accs = AccumAmDiagGmm.init(self.acoustic_model)
        self.num_done = 0    # utterances accumulated / skipped because the alignment length differs from the frame count
        self.num_error = 0

    def accumulate_stats(self, feature_archive: FeatureArchive, alignment_archive: AlignmentArchive, callback: Optional[Callable] = None,
                         batch_size: int = 512):
        keys = [k for k in feature_archive.keys if k in alignment_archive]
        for i in range(0, len(keys), batch_size):
            ks = keys[i:i + batch_size]
            feats, fo = feature_archive.batch(ks)
            ali = np.zeros(int(fo[-1]), np.int32)
            for j, k in enumerate(ks):
                a = np.asarray(alignment_archive[k].alignment, np.int32)
                if len(a) != int(fo[j + 1] - fo[j]):   # gmm-acc-stats-ali: "Alignments has wrong size" -> utterance skipped, error counted
                    self.num_error += 1
                    continue                           # transition-id 0 = frame ignored by the kernel
                ali[fo[j]:fo[j + 1]] = a
                self.num_done += 1
            self._dm.acc_stats(np.ascontiguousarray(feats, np.float32), ali)
            if callback:
                callback(len(ks))
        self.sync()

    def accumulate_batch(self, feats: np.ndarray, ali: np.ndarray):
        """AccumAmDiagGmm.acc_stats + TransitionModel.acc_stats for concatenated utterances (acoustic_modeling/monophone.py:114-120)."""
        self._dm.acc_stats(np.ascontiguousarray(feats, np.float32), np.ascontiguousarray(ali, np.int32))

    def device_tensor(self):
        """f64 device view of the accumulator block, for torch.distributed.all_reduce (NCCL)."""
        return self._dm.acc_tensor()

    def sync(self):
        d = self._dm.acc_read()
        self.gmm_accs = AccumAmDiagGmm.from_dict(d)
        self.transition_accs = np.array(d["trans"], dtype=np.float64)


from .fmllr import FmllrComputer, compose_transforms  # noqa: E402  (kalpy exposes FmllrComputer next to the aligner classes)


class MatrixArchive:
    """kalpy.util MatrixArchive (corpus/features.py:494,531): speaker -> float matrix (trans.ark / trans.scp)."""

    def __init__(self, file_name):
        self._m = {k: np.asarray(v, np.float32) for k, v in _read_table(file_name, "matrix").items()}

    def __getitem__(self, key):
        return self._m[str(key)]

    def __contains__(self, key):
        return str(key) in self._m

    def __iter__(self):
        return iter(self._m.items())
