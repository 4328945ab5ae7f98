"""ctypes wrapper over oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (montreal-forced-aligner_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
def _has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " avx2" in line
    except OSError:
        pass
    return False


# MFA_ORACLE_GENERIC=1 forces the generic build (tests compare the two builds)
VARIANT = "avx2" if (_has_avx2() and not os.environ.get("MFA_ORACLE_GENERIC")) else "generic"
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_avx2.so" if VARIANT == "avx2" else "liboracle.so")


def build(force: bool = False) -> str:
    """Both builds (so that a library built in one container still loads on a host without AVX2)."""
    src = os.path.join(_HERE, "oracle.c")
    mk = os.path.join(_HERE, "Makefile")
    libs = [os.path.join(_HERE, "_build", n) for n in ("liboracle.so", "liboracle_avx2.so")]
    newest = max(os.path.getmtime(src), os.path.getmtime(mk))
    if force or any(not os.path.exists(l) or os.path.getmtime(l) < newest for l in libs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class MfccOpts(C.Structure):
    _fields_ = [
        ("sample_frequency", C.c_float), ("frame_length_ms", C.c_float), ("frame_shift_ms", C.c_float),
        ("preemph_coeff", C.c_float), ("low_freq", C.c_float), ("high_freq", C.c_float),
        ("cepstral_lifter", C.c_float), ("energy_floor", C.c_float),
        ("num_mel_bins", C.c_int32), ("num_ceps", C.c_int32), ("use_energy", C.c_int32),
        ("raw_energy", C.c_int32), ("snip_edges", C.c_int32), ("remove_dc_offset", C.c_int32),
    ]


def mfcc_opts(**kw) -> MfccOpts:
    d = dict(sample_frequency=16000.0, frame_length_ms=25.0, frame_shift_ms=10.0, preemph_coeff=0.97,
             low_freq=20.0, high_freq=7800.0, cepstral_lifter=22.0, energy_floor=0.0, num_mel_bins=23,
             num_ceps=13, use_energy=0, raw_energy=1, snip_edges=1, remove_dc_offset=1)
    d.update(kw)
    return MfccOpts(**d)


class Gmm(C.Structure):
    _fields_ = [("dim", C.c_int32), ("num_pdfs", C.c_int32), ("off", C.c_void_p), ("gconsts", C.c_void_p),
                ("means_invvars", C.c_void_p), ("inv_vars", C.c_void_p)]


class FstS(C.Structure):
    _fields_ = [("num_states", C.c_int32), ("start", C.c_int32), ("arc_off", C.c_void_p), ("ilabel", C.c_void_p),
                ("olabel", C.c_void_p), ("nextstate", C.c_void_p), ("weight", C.c_void_p), ("final", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_mfcc_num_frames.restype = C.c_int64
        _lib.orc_mfcc_num_frames.argtypes = [C.c_void_p, C.c_int64]
        assert _lib.orc_sizeof_mfcc_opts() == C.sizeof(MfccOpts)
        assert _lib.orc_sizeof_gmm() == C.sizeof(Gmm)
        assert _lib.orc_sizeof_fst() == C.sizeof(FstS)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def mfcc(pcm: np.ndarray, opts: MfccOpts | None = None) -> np.ndarray:
    opts = opts or mfcc_opts()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    T = lib().orc_mfcc_num_frames(C.byref(opts), C.c_int64(pcm.shape[0]))
    out = np.zeros((T, opts.num_ceps), dtype=np.float32)
    lib().orc_mfcc(C.byref(opts), _p(pcm), C.c_int64(pcm.shape[0]), _p(out))
    return out


def cmvn_stats(feat_list) -> np.ndarray:
    dim = feat_list[0].shape[1]
    stats = np.zeros((2, dim + 1), dtype=np.float64)
    for f in feat_list:
        f = np.ascontiguousarray(f, dtype=np.float32)
        lib().orc_cmvn_acc(_p(f), C.c_int64(f.shape[0]), C.c_int(dim), _p(stats))
    return stats


def cmvn_apply(feats: np.ndarray, stats: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(feats, dtype=np.float32).copy()
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    lib().orc_cmvn_apply(_p(out), C.c_int64(out.shape[0]), C.c_int(out.shape[1]), _p(stats))
    return out


def add_deltas(feats: np.ndarray) -> np.ndarray:
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    out = np.zeros((feats.shape[0], 3 * feats.shape[1]), dtype=np.float32)
    lib().orc_add_deltas(_p(feats), C.c_int64(feats.shape[0]), C.c_int(feats.shape[1]), _p(out))
    return out


def splice(feats: np.ndarray, left=3, right=3) -> np.ndarray:
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    out = np.zeros((feats.shape[0], (left + right + 1) * feats.shape[1]), dtype=np.float32)
    lib().orc_splice(_p(feats), C.c_int64(feats.shape[0]), C.c_int(feats.shape[1]), C.c_int(left), C.c_int(right), _p(out))
    return out


def transform(feats: np.ndarray, M: np.ndarray) -> np.ndarray:
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    out = np.zeros((feats.shape[0], M.shape[0]), dtype=np.float32)
    lib().orc_transform(_p(feats), C.c_int64(feats.shape[0]), C.c_int(feats.shape[1]), _p(M), C.c_int(M.shape[0]),
                        C.c_int(M.shape[1]), _p(out))
    return out


class GmmModel:
    """Keeps the numpy arrays alive behind an ``orc_gmm`` struct."""

    def __init__(self, dim, offsets, gconsts, means_invvars, inv_vars):
        self.off = np.ascontiguousarray(offsets, dtype=np.int32)
        self.gconsts = np.ascontiguousarray(gconsts, dtype=np.float32)
        self.miv = np.ascontiguousarray(means_invvars, dtype=np.float32)
        self.iv = np.ascontiguousarray(inv_vars, dtype=np.float32)
        self.dim = int(dim)
        self.num_pdfs = self.off.shape[0] - 1
        self.s = Gmm(self.dim, self.num_pdfs, _p(self.off).value, _p(self.gconsts).value, _p(self.miv).value, _p(self.iv).value)

    @classmethod
    def from_am(cls, am, gconsts=None):
        return cls(am.dim, am.offsets, am.gconsts if gconsts is None else gconsts, am.means_invvars, am.inv_vars)


def gmm_loglikes(g: GmmModel, feats: np.ndarray) -> np.ndarray:
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    out = np.zeros((feats.shape[0], g.num_pdfs), dtype=np.float32)
    lib().orc_gmm_loglikes(C.byref(g.s), _p(feats), C.c_int64(feats.shape[0]), _p(out))
    return out


class FstCsr:
    """Arcs grouped by source state, original order kept (OpenFst iteration order)."""

    def __init__(self, fst):
        order = np.argsort(fst.arc_src, kind="stable")
        self.arc_off = np.searchsorted(fst.arc_src[order], np.arange(fst.num_states + 1)).astype(np.int32)
        self.ilabel = np.ascontiguousarray(fst.arc_ilabel[order], dtype=np.int32)
        self.olabel = np.ascontiguousarray(fst.arc_olabel[order], dtype=np.int32)
        self.next = np.ascontiguousarray(fst.arc_dst[order], dtype=np.int32)
        self.weight = np.ascontiguousarray(fst.arc_weight[order], dtype=np.float32)
        self.final = np.ascontiguousarray(fst.finals, dtype=np.float32)
        self.s = FstS(fst.num_states, fst.start, _p(self.arc_off).value, _p(self.ilabel).value, _p(self.olabel).value,
                      _p(self.next).value, _p(self.weight).value, _p(self.final).value)


STATUS = {0: "OK", 1: "RETRIED", 2: "NO_FINAL", 3: "EMPTY_GRAPH", 4: "ZERO_FRAMES"}


def align(fst, tid_cost: np.ndarray, g: GmmModel, tid2pdf: np.ndarray, feats: np.ndarray | None, T: int,
          acoustic_scale=0.1, beam=10.0, retry_beam=40.0, dense: np.ndarray | None = None, max_words=4096):
    """One utterance through AlignUtteranceWrapper semantics. Returns dict(status, ali, words, per_frame, like)."""
    csr = fst if isinstance(fst, FstCsr) else FstCsr(fst)
    tid_cost = np.ascontiguousarray(tid_cost, dtype=np.float32)
    tid2pdf = np.ascontiguousarray(tid2pdf, dtype=np.int32)
    ali = np.zeros(max(T, 1), dtype=np.int32)
    words = np.zeros(max_words, dtype=np.int32)
    nw = C.c_int32(0)
    pf = np.zeros(max(T, 1), dtype=np.float32)
    like = C.c_float(0)
    fp = _p(np.ascontiguousarray(feats, dtype=np.float32)) if feats is not None else None
    if feats is not None:
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        fp = _p(feats)
    dp = None
    if dense is not None:
        dense = np.ascontiguousarray(dense, dtype=np.float32)
        dp = _p(dense)
    st = lib().orc_align(C.byref(csr.s), _p(tid_cost), C.byref(g.s), _p(tid2pdf), fp, dp, C.c_int64(T),
                         C.c_float(acoustic_scale), C.c_float(beam), C.c_float(retry_beam), _p(ali), _p(words),
                         C.byref(nw), C.c_int32(max_words), _p(pf), C.byref(like))
    return dict(status=st, ali=ali[:T], words=words[: nw.value], per_frame=pf[:T], like=like.value)


def acc_stats(g: GmmModel, tid2pdf: np.ndarray, feats: np.ndarray, ali: np.ndarray, num_tids: int, accs=None):
    G = g.gconsts.shape[0]
    D = g.dim
    if accs is None:
        accs = dict(occ=np.zeros(G), mean=np.zeros((G, D)), var=np.zeros((G, D)), trans=np.zeros(num_tids + 1),
                    like=np.zeros(1), frames=0)
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    ali = np.ascontiguousarray(ali, dtype=np.int32)
    tid2pdf = np.ascontiguousarray(tid2pdf, dtype=np.int32)
    lib().orc_acc_stats(C.byref(g.s), _p(tid2pdf), _p(feats), _p(ali), C.c_int64(ali.shape[0]), _p(accs["occ"]),
                        _p(accs["mean"]), _p(accs["var"]), _p(accs["trans"]), _p(accs["like"]))
    accs["frames"] += int(ali.shape[0])
    return accs


def equal_align(fst, T: int, seed: int, num_retries: int = 10, max_words: int = 4096):
    """Kaldi EqualAlign + GetLinearSymbolSequence (libc srand/rand). Returns dict(status, ali, words)."""
    csr = fst if isinstance(fst, FstCsr) else FstCsr(fst)
    ali = np.zeros(max(T, 1), dtype=np.int32)
    words = np.zeros(max_words, dtype=np.int32)
    nw = C.c_int32(0)
    st = lib().orc_equal_align(C.byref(csr.s), C.c_int64(T), C.c_uint(seed & 0xFFFFFFFF), C.c_int(num_retries), _p(ali), _p(words),
                               C.byref(nw), C.c_int32(max_words))
    return dict(status=st, ali=ali[:T], words=words[: nw.value])


def fmllr_stats_size(D: int) -> int:
    return 1 + D * (D + 1) + D * ((D + 1) * (D + 2) // 2)


def fmllr_acc(g_post: GmmModel, g: GmmModel, tid2pdf: np.ndarray, tid_weight, feats: np.ndarray, ali: np.ndarray, stats=None):
    """FmllrDiagGmmAccs accumulation for one speaker's frames: stats = beta | K[D][D+1] | G[D][packed lower triangle]."""
    D = g.dim
    if stats is None:
        stats = np.zeros(fmllr_stats_size(D), dtype=np.float64)
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    ali = np.ascontiguousarray(ali, dtype=np.int32)
    tid2pdf = np.ascontiguousarray(tid2pdf, dtype=np.int32)
    tw = None if tid_weight is None else np.ascontiguousarray(tid_weight, dtype=np.float32)
    lib().orc_fmllr_acc(C.byref(g_post.s), C.byref(g.s), _p(tid2pdf), None if tw is None else _p(tw), _p(feats), _p(ali),
                        C.c_int64(ali.shape[0]), _p(stats))
    return stats


def fmllr_update(stats: np.ndarray, D: int, num_iters: int = 40, min_count: float = 500.0, W0=None):
    """ComputeFmllrMatrixDiagGmmFull from the unit transform (gmm-est-fmllr). Returns (W [D][D+1] f32, objf improvement)."""
    lib().orc_fmllr_update.restype = C.c_double
    W = np.ascontiguousarray(np.hstack([np.eye(D), np.zeros((D, 1))]) if W0 is None else W0, dtype=np.float32).copy()
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    impr = lib().orc_fmllr_update(_p(stats), C.c_int(D), C.c_int(num_iters), C.c_double(min_count), _p(W))
    return W, float(impr)
