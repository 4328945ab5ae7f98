"""Python host layer over the C ABI: engine, device-resident acoustic model, graph compiler, batched hot-path calls.

numpy arrays are passed as host buffers (MFA_HOST), torch CUDA tensors as device buffers (MFA_DEVICE).
torch is only used for device memory / streams / torch.distributed plumbing.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .kaldi_io import AmDiagGmm, ContextDependency, Fst, TransitionModel


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _buf(x, dtype, name="buffer"):
    """-> (keepalive, address, where). numpy -> host; torch cuda tensor -> device."""
    if x is None:
        return None, None, None
    if _is_torch(x):
        import torch
        want = {np.float32: torch.float32, np.int32: torch.int32, np.int16: torch.int16, np.float64: torch.float64,
                np.int64: torch.int64}[dtype]
        if x.dtype != want or not x.is_contiguous():
            raise TypeError(f"{name}: expected contiguous torch tensor of {want}")
        return x, C.c_void_p(x.data_ptr()), (L.MFA_DEVICE if x.is_cuda else L.MFA_HOST)
    a = np.ascontiguousarray(x, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p), L.MFA_HOST


def _host(x, dtype):
    a = np.ascontiguousarray(x, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


def mfcc_opts(**kw) -> L.MfccOpts:
    """kalpy/MFA option names (corpus/features.py:780-820) -> mfa_mfcc_opts."""
    d = dict(sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0, preemphasis_coefficient=0.97,
             low_frequency=20.0, high_frequency=7800.0, cepstral_lifter=22.0, energy_floor=0.0, num_mel_bins=23,
             num_coefficients=13, use_energy=False, raw_energy=True, snip_edges=True, remove_dc_offset=True, dither=0.0)
    for k, v in kw.items():
        if k in ("allow_downsample", "allow_upsample", "sample_frequency_in"):
            continue
        if k not in d:
            raise TypeError(f"unknown MFCC option {k!r}")
        d[k] = v
    if d["dither"] not in (0, 0.0):
        raise L.MfaError("dither != 0 is not supported by the B200 engine (parity runs force dither=0; SURVEY.md section 5)")
    return L.MfccOpts(float(d["sample_frequency"]), float(d["frame_length"]), float(d["frame_shift"]),
                      float(d["preemphasis_coefficient"]), float(d["low_frequency"]), float(d["high_frequency"]),
                      float(d["cepstral_lifter"]), float(d["energy_floor"]), int(d["num_mel_bins"]),
                      int(d["num_coefficients"]), int(bool(d["use_energy"])), int(bool(d["raw_energy"])),
                      int(bool(d["snip_edges"])), int(bool(d["remove_dc_offset"])))


def num_frames(opts: L.MfccOpts, n_samples: int) -> int:
    return int(L.lib().mfa_mfcc_num_frames(C.byref(opts), C.c_int64(int(n_samples))))


def frame_offsets(opts: L.MfccOpts, sample_off) -> np.ndarray:
    """Vectorised mfa_mfcc_num_frames over a batch: sample_off[n+1] -> frame_off[n+1]."""
    so = np.asarray(sample_off, dtype=np.int64)
    n = so[1:] - so[:-1]
    shift = int(np.float32(opts.sample_frequency) * np.float32(0.001) * np.float32(opts.frame_shift_ms))
    length = int(np.float32(opts.sample_frequency) * np.float32(0.001) * np.float32(opts.frame_length_ms))
    if opts.snip_edges:
        t = np.where(n < length, 0, 1 + (n - length) // shift)
    else:
        t = (n + shift // 2) // shift
    fo = np.zeros(so.shape[0], dtype=np.int64)
    fo[1:] = np.cumsum(t)
    return fo


class Engine:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L.check(L.lib().mfa_engine_create(C.c_int(device), C.byref(self._h)))
        self.device = device
        self._keepalive = []

    def close(self):
        if self._h:
            L.lib().mfa_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(L.lib().mfa_engine_sync(self._h))
        self._keepalive.clear()

    def _hold(self, *tensors):
        """Device-buffer calls return before their kernels have run (the Viterbi launch is not even joined on the main stream): the torch
        tensors they read and write stay referenced here until sync(), so that torch's caching allocator -- which knows nothing about the
        engine's streams -- cannot hand their memory to someone else meanwhile."""
        self._keepalive.append(tensors)

    @property
    def stream(self) -> int:
        return int(L.lib().mfa_engine_stream(self._h) or 0)

    @property
    def sm_count(self) -> int:
        return int(L.lib().mfa_engine_sm_count(self._h))

    def set_option(self, name: str, value: int):
        """mfa_engine_set_option: experiment / test switches (see include/mfa_b200.h); the environment is only read at creation."""
        L.check(L.lib().mfa_engine_set_option(self._h, name.encode(), C.c_int(int(value))))

    def get_option(self, name: str) -> int:
        v = C.c_int()
        L.check(L.lib().mfa_engine_get_option(self._h, name.encode(), C.byref(v)))
        return v.value

    def options(self, **kw):
        """Context manager: set the given options, restore the previous values on exit."""
        eng = self

        class _Ctx:
            def __enter__(self):
                self.old = {k: eng.get_option(k) for k in kw}
                for k, v in kw.items():
                    eng.set_option(k, v)
                return eng

            def __exit__(self, *a):
                for k, v in self.old.items():
                    eng.set_option(k, v)
                return False
        return _Ctx()

    @property
    def launch_count(self) -> int:
        return int(L.lib().mfa_engine_launch_count(self._h))

    @property
    def band_fallbacks(self) -> int:
        """Utterances the band Viterbi kernel handed to the sparse kernel so far (live window wider than the band)."""
        return int(L.lib().mfa_engine_band_fallbacks(self._h))

    def gmm_timing(self):
        ms, n, rows = C.c_float(), C.c_int64(), C.c_int64()
        L.check(L.lib().mfa_engine_gmm_timing(self._h, C.byref(ms), C.byref(n), C.byref(rows)))
        return ms.value, n.value, rows.value

    def stage_timing(self) -> dict:
        ms = (C.c_float * 4)()
        L.check(L.lib().mfa_engine_stage_timing(self._h, ms))
        return {"mfcc_cmvn": ms[0], "features": ms[1], "gmm": ms[2], "viterbi": ms[3]}

    def gmm_flops(self) -> float:
        f = C.c_double()
        L.check(L.lib().mfa_engine_gmm_flops(self._h, C.byref(f)))
        return f.value

    def gmm_issued_flops(self) -> float:
        """FLOPs the tensor-core launches of the last call issued (padding included); 0 for the CUDA-core kernels."""
        f = C.c_double()
        L.check(L.lib().mfa_engine_gmm_issued_flops(self._h, C.byref(f)))
        return f.value

    def fmllr_update(self, stats, dim: int, num_iters: int = 40, min_count: float = 500.0):
        """mfa_fmllr_update: per-speaker statistics [S, size] (numpy or torch cuda f64) -> (W [S, D, D+1] f32 of the same kind,
        objective improvement [S] numpy, count [S] numpy)."""
        k, sp, where = _buf(stats, np.float64, "stats")
        S = int(stats.shape[0])
        if where == L.MFA_DEVICE:
            import torch
            W = torch.empty((S, dim, dim + 1), dtype=torch.float32, device=stats.device)
        else:
            W = np.empty((S, dim, dim + 1), dtype=np.float32)
        k2, wp, _ = _buf(W, np.float32, "transforms")
        impr, cnt = np.zeros(max(S, 1), np.float64), np.zeros(max(S, 1), np.float64)
        L.check(L.lib().mfa_fmllr_update(self._h, sp, C.c_int32(dim), C.c_int32(S), C.c_int32(num_iters), C.c_double(min_count), wp,
                                         impr.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), C.c_int(where)))
        return W, impr[:S], cnt[:S]

    # ---- K1 / CMVN / features -----------------------------------------------------------------
    def mfcc(self, pcm, sample_off, opts: L.MfccOpts, out=None):
        so, sop = _host(sample_off, np.int64)
        n = so.shape[0] - 1
        fo = frame_offsets(opts, so)
        keep, pp, where = _buf(pcm, np.int16, "pcm")
        if out is None:
            if where == L.MFA_DEVICE:
                import torch
                out = torch.empty((int(fo[-1]), opts.num_ceps), dtype=torch.float32, device=pcm.device)
            else:
                out = np.empty((int(fo[-1]), opts.num_ceps), dtype=np.float32)
        ko, op, w2 = _buf(out, np.float32, "out")
        assert w2 == where
        L.check(L.lib().mfa_mfcc(self._h, C.byref(opts), pp, sop, C.c_int32(n), fo.ctypes.data_as(C.c_void_p), op, C.c_int(where)))
        return out, fo

    def cmvn_stats(self, feats, frame_off, utt2spk, n_spk: int):
        fo, fop = _host(frame_off, np.int64)
        us, usp = _host(utt2spk, np.int32)
        keep, fp, where = _buf(feats, np.float32, "feats")
        dim = feats.shape[1]
        if where == L.MFA_DEVICE:
            import torch
            stats = torch.zeros((n_spk, 2, dim + 1), dtype=torch.float64, device=feats.device)
        else:
            stats = np.zeros((n_spk, 2, dim + 1), dtype=np.float64)
        ks, sp, _ = _buf(stats, np.float64, "stats")
        L.check(L.lib().mfa_cmvn_stats(self._h, fp, C.c_int32(dim), fop, usp, C.c_int32(fo.shape[0] - 1), C.c_int32(n_spk), sp, C.c_int(where)))
        return stats

    def features(self, feats, frame_off, mode: str = "deltas", lda: Optional[np.ndarray] = None, splice_ctx: int = 3,
                 fmllr: Optional[np.ndarray] = None, cmvn_stats: Optional[np.ndarray] = None, utt2spk=None, n_spk: int = 0):
        fo, fop = _host(frame_off, np.int64)
        n = fo.shape[0] - 1
        o, keep = make_feat_opts(feats.shape[1], mode, lda, splice_ctx, fmllr, cmvn_stats, n_spk)
        od = int(L.lib().mfa_feat_out_dim(C.byref(o)))
        k1, ip, where = _buf(feats, np.float32, "feats")
        if where == L.MFA_DEVICE:
            import torch
            out = torch.empty((feats.shape[0], od), dtype=torch.float32, device=feats.device)
        else:
            out = np.empty((feats.shape[0], od), dtype=np.float32)
        k2, op, _ = _buf(out, np.float32)
        usp = None
        if utt2spk is not None:
            us, usp = _host(utt2spk, np.int32)
        L.check(L.lib().mfa_features(self._h, C.byref(o), ip, fop, usp, C.c_int32(n), op, C.c_int(where)))
        return out


def make_feat_opts(in_dim, mode, lda, splice_ctx, fmllr, cmvn_stats, n_spk):
    keep = []
    o = L.FeatOpts()
    o.mode = {"none": 0, "deltas": 1, "lda": 2, "splice_lda": 2}[mode]
    o.in_dim = int(in_dim)
    o.splice_ctx = int(splice_ctx)
    o.n_spk = int(n_spk)
    if o.mode == 2:
        if lda is None:
            raise ValueError("mode 'lda' needs the LDA matrix")
        a, p = _host(lda, np.float32)
        keep.append(a)
        o.lda = p.value
        o.lda_rows, o.lda_cols = a.shape
    if fmllr is not None:
        a, p = _host(fmllr, np.float32)
        keep.append(a)
        o.fmllr = p.value
        o.n_spk = a.shape[0]
    if cmvn_stats is not None:
        a, p = _host(cmvn_stats, np.float64)
        keep.append(a)
        o.cmvn_stats = p.value
        o.n_spk = a.shape[0]
    o._keep = keep
    return o, keep


class DeviceModel:
    """Device-resident AmDiagGmm + tid->pdf map (mfa_model)."""

    def __init__(self, engine: Engine, tm: Optional[TransitionModel], am: AmDiagGmm):
        """tm may be None for a bare AmDiagGmm (M-step of host-side accumulators): no transition-ids, nothing to align."""
        self.engine = engine
        if tm is None:
            class _NoTm:
                tid2pdf = np.zeros(1, np.int32)
                num_tids = 0
            tm = _NoTm()
        self.tm, self.am = tm, am
        self._arrs = [np.ascontiguousarray(am.offsets, dtype=np.int32), np.ascontiguousarray(am.gconsts, dtype=np.float32),
                      np.ascontiguousarray(am.means_invvars, dtype=np.float32), np.ascontiguousarray(am.inv_vars, dtype=np.float32),
                      np.ascontiguousarray(np.maximum(tm.tid2pdf, 0), dtype=np.int32), np.ascontiguousarray(am.weights, dtype=np.float32)]
        d = L.ModelDesc(am.dim, am.NumPdfs(), am.NumGauss(), tm.num_tids, *[a.ctypes.data_as(C.c_void_p).value for a in self._arrs])
        self._h = C.c_void_p()
        L.check(L.lib().mfa_model_create(engine._h, C.byref(d), C.byref(self._h)))
        self.dim, self.num_pdfs, self.num_gauss, self.num_tids = am.dim, am.NumPdfs(), am.NumGauss(), tm.num_tids
        self._trans_set = False

    def close(self):
        if self._h:
            L.lib().mfa_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def boost_pdfs(self, factor: float, pdfs: Sequence[int]):
        p, pp = _host(np.asarray(sorted(set(int(x) for x in pdfs)), dtype=np.int32), np.int32)
        L.check(L.lib().mfa_model_boost_pdfs(self._h, C.c_float(factor), pp, C.c_int32(p.shape[0])))

    def loglikes(self, feats, impl: int = 0):
        """[T, dim] -> [T, num_pdfs] unscaled log-likelihoods (frame-major, the kalpy gmm_compute_likes layout)."""
        k, fp, where = _buf(feats, np.float32, "feats")
        T = feats.shape[0]
        if where == L.MFA_DEVICE:
            import torch
            out = torch.empty((T, self.num_pdfs), dtype=torch.float32, device=feats.device)
        else:
            out = np.empty((T, self.num_pdfs), dtype=np.float32)
        k2, op, _ = _buf(out, np.float32)
        L.check(L.lib().mfa_gmm_loglikes(self.engine._h, self._h, fp, C.c_int64(T), op, C.c_int(where), C.c_int(impl)))
        return out

    # ---- N3: M-step on the device
    def set_transitions(self, tm: Optional[TransitionModel] = None):
        """mfa_model_set_transitions: transition-state tables + current log-probabilities (needed to re-estimate transitions on the
        device and to re-fold them into packed graphs)."""
        tm = tm or self.tm
        arrs = [np.ascontiguousarray(tm.state2id, np.int32), np.ascontiguousarray(tm.self_loop_tid, np.int32),
                np.ascontiguousarray(tm.log_probs, np.float32)]
        d = L.TransDesc(int(tm.tuples.shape[0]), *[a.ctypes.data_as(C.c_void_p).value for a in arrs])
        L.check(L.lib().mfa_model_set_transitions(self._h, C.byref(d)))
        self._trans_set = True

    def mle_update(self, mixup: int = 0, power: float = 0.25, min_gaussian_occupancy: float = 10.0, min_gaussian_weight: float = 1.0e-5,
                   min_variance: float = 0.001, remove_low_count_gaussians: bool = True, perturb_factor: float = 0.01, min_count: float = 20.0,
                   update_transitions: bool = False, transition_floor: float = 0.01, transition_mincount: float = 5.0, seed: int = 1234) -> dict:
        """mfa_model_mle_update: the model is re-estimated in place from its (all-reduced) device accumulators; returns the result
        struct as a dict.  ``self.am`` / ``self.tm`` are NOT touched: call ``read()`` for host copies of the new parameters."""
        if update_transitions and not self._trans_set:
            self.set_transitions()
        o = L.MleOpts(float(min_gaussian_occupancy), float(min_gaussian_weight), float(min_variance), int(bool(remove_low_count_gaussians)),
                      int(mixup or 0), float(power), float(min_count), float(perturb_factor), int(bool(update_transitions)),
                      float(transition_floor), float(transition_mincount), int(seed) & 0xFFFFFFFFFFFFFFFF)
        r = L.MleResult()
        L.check(L.lib().mfa_model_mle_update(self.engine._h, self._h, C.byref(o), C.byref(r)))
        self.num_gauss = int(r.num_gauss_after)
        return {k: getattr(r, k) for k, _ in L.MleResult._fields_}

    def reserve(self, max_gauss: int):
        """mfa_model_reserve: allocate now what M-steps and accumulators need for up to `max_gauss` Gaussians (training loops: pass current
        Gaussians + mix-up target + 1), so that no iteration allocates device memory."""
        L.check(L.lib().mfa_model_reserve(self.engine._h, self._h, C.c_int64(int(max_gauss))))

    def read(self, with_transitions: bool = False):
        """mfa_model_read: host AmDiagGmm of the current device parameters (+ the transition log-probabilities)."""
        G, D = int(L.lib().mfa_model_num_gauss(self._h)), self.dim
        off, w, gc = np.zeros(self.num_pdfs + 1, np.int32), np.zeros(G, np.float32), np.zeros(G, np.float32)
        miv, iv = np.zeros((G, D), np.float32), np.zeros((G, D), np.float32)
        lp = np.zeros(self.num_tids + 1, np.float32) if with_transitions else None
        L.check(L.lib().mfa_model_read(self.engine._h, self._h, *[a.ctypes.data_as(C.c_void_p) for a in (off, w, gc, miv, iv)],
                                       lp.ctypes.data_as(C.c_void_p) if lp is not None else None))
        am = AmDiagGmm(D, off, w, miv, iv)
        am.device_gconsts = gc
        return (am, lp) if with_transitions else am

    # ---- K4
    def acc_size(self) -> int:
        return int(L.lib().mfa_acc_size(self._h))

    def acc_zero(self):
        L.check(L.lib().mfa_acc_zero(self.engine._h, self._h))

    def acc_stats(self, feats, ali):
        k, fp, where = _buf(feats, np.float32, "feats")
        k2, ap, w2 = _buf(ali, np.int32, "ali")
        assert where == w2
        L.check(L.lib().mfa_acc_stats(self.engine._h, self._h, fp, ap, C.c_int64(feats.shape[0]), C.c_int(where)))

    # ---- K5
    def fmllr_stats_size(self) -> int:
        return int(L.lib().mfa_fmllr_stats_size(C.c_int32(self.dim)))

    def fmllr_acc(self, feats, ali, frame_off, utt2spk, n_spk: int, tid_weight=None, post_model: Optional["DeviceModel"] = None):
        """Per-speaker fMLLR statistics (mfa_fmllr_acc): [n_spk, 1 + D(D+1) + D(D+1)(D+2)/2] f64 (numpy in -> numpy out, torch in -> torch out)."""
        k, fp, where = _buf(feats, np.float32, "feats")
        k2, ap, w2 = _buf(ali, np.int32, "ali")
        assert where == w2
        fo, fop = _host(frame_off, np.int64)
        us, usp = _host(utt2spk, np.int32)
        twp = None
        if tid_weight is not None:
            tw, twp = _host(tid_weight, np.float32)
            assert tw.shape[0] == self.num_tids + 1
        n = self.fmllr_stats_size()
        if where == L.MFA_DEVICE:
            import torch
            stats = torch.zeros((n_spk, n), dtype=torch.float64, device=feats.device)
        else:
            stats = np.zeros((n_spk, n), dtype=np.float64)
        k3, sp, _ = _buf(stats, np.float64, "stats")
        L.check(L.lib().mfa_fmllr_acc(self.engine._h, post_model._h if post_model is not None else None, self._h, fp, ap, twp, fop, usp,
                                      C.c_int32(fo.shape[0] - 1), C.c_int32(n_spk), sp, C.c_int(where)))
        return stats

    def acc_device_ptr(self) -> int:
        return int(L.lib().mfa_acc_device_ptr(self.engine._h, self._h) or 0)

    def acc_tensor(self):
        """The device accumulator block as a torch f64 tensor view (for torch.distributed.all_reduce over NCCL)."""
        import torch
        n = self.acc_size()
        ptr = self.acc_device_ptr()
        if not ptr:
            raise L.MfaError("accumulators not allocated; call acc_zero() first")

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Holder(), device=f"cuda:{self.engine.device}")

    def acc_read(self) -> dict:
        out = np.zeros(self.acc_size(), dtype=np.float64)
        L.check(L.lib().mfa_acc_read(self.engine._h, self._h, out.ctypes.data_as(C.c_void_p)))
        return self.split_accs(out)

    def acc_write(self, occ, mean, var, trans=None, like: float = 0.0, frames: float = 0.0):
        """mfa_acc_write: host accumulators (kalpy-style objects summed on the host) -> the device block."""
        G, D, nt = self.num_gauss, self.dim, self.num_tids
        flat = np.zeros(self.acc_size(), np.float64)
        flat[:G] = occ
        flat[G:G + G * D] = np.asarray(mean, np.float64).reshape(-1)
        flat[G + G * D:G + 2 * G * D] = np.asarray(var, np.float64).reshape(-1)
        o = G + 2 * G * D
        if trans is not None:
            flat[o:o + nt + 1] = trans
        flat[o + nt + 1], flat[o + nt + 2] = like, frames
        L.check(L.lib().mfa_acc_write(self.engine._h, self._h, flat.ctypes.data_as(C.c_void_p)))

    def split_accs(self, flat: np.ndarray) -> dict:
        G, D, nt = self.num_gauss, self.dim, self.num_tids
        o = 0
        occ = flat[o:o + G]; o += G
        mean = flat[o:o + G * D].reshape(G, D); o += G * D
        var = flat[o:o + G * D].reshape(G, D); o += G * D
        trans = flat[o:o + nt + 1]; o += nt + 1
        return dict(occ=occ, mean=mean, var=var, trans=trans, like=float(flat[o]), frames=float(flat[o + 1]))


def hmm_desc(tm: TransitionModel, tree: ContextDependency):
    """Flatten topology + tuples + tree for mfa_graph_compiler_create. Returns (desc, keepalive list)."""
    topo = tm.topo
    nph = int(topo.phone2idx.shape[0])
    phone2entry = np.ascontiguousarray(topo.phone2idx, dtype=np.int32)
    eso = [0]
    fwd, slf, toff, tdst = [], [], [0], []
    for e in topo.entries:
        for s in e:
            has = len(s.transitions) > 0
            fwd.append(s.forward_pdf_class if has else -1)
            slf.append(s.self_loop_pdf_class if has else -1)
            for dst, _p in s.transitions:
                tdst.append(dst)
            toff.append(len(tdst))
        eso.append(len(fwd))
    node, aux_off, aux, root = tree.flatten()
    arrs = [phone2entry, np.asarray(eso, np.int32), np.asarray(fwd, np.int32), np.asarray(slf, np.int32), np.asarray(toff, np.int32),
            np.asarray(tdst if tdst else [0], np.int32), np.ascontiguousarray(tm.tuples, dtype=np.int32),
            np.ascontiguousarray(tm.state2id, dtype=np.int32), np.ascontiguousarray(node, np.int32),
            np.ascontiguousarray(aux_off, np.int32), np.ascontiguousarray(aux if aux.size else np.zeros(1, np.int32), np.int32)]
    p = [a.ctypes.data_as(C.c_void_p).value for a in arrs]
    d = L.HmmDesc(nph, p[0], len(topo.entries), p[1], p[2], p[3], p[4], p[5], tm.tuples.shape[0], p[6], p[7], tree.N, tree.P,
                  node.shape[0], root, p[8], p[9], p[10])
    return d, arrs


class FstBatch:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_fsts(cls, fsts: List[Fst]) -> "FstBatch":
        n = len(fsts)
        so = np.zeros(n + 1, np.int64)
        ao = np.zeros(n + 1, np.int64)
        for i, f in enumerate(fsts):
            so[i + 1] = so[i] + f.num_states
            ao[i + 1] = ao[i] + f.arc_src.shape[0]
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        start = np.asarray([f.start for f in fsts], np.int32)
        arrs = [so, ao, start, cat([f.finals for f in fsts], np.float32), cat([f.arc_src for f in fsts], np.int32),
                cat([f.arc_dst for f in fsts], np.int32), cat([f.arc_ilabel for f in fsts], np.int32),
                cat([f.arc_olabel for f in fsts], np.int32), cat([f.arc_weight for f in fsts], np.float32)]
        h = C.c_void_p()
        L.check(L.lib().mfa_fst_batch_create(C.c_int32(n), *[a.ctypes.data_as(C.c_void_p) for a in arrs], C.byref(h)))
        return cls(h)

    def sizes(self):
        n, s, a = C.c_int32(), C.c_int64(), C.c_int64()
        L.check(L.lib().mfa_fst_batch_sizes(self._h, C.byref(n), C.byref(s), C.byref(a)))
        return n.value, s.value, a.value

    def export(self) -> List[Fst]:
        n, S, A = self.sizes()
        so, ao = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
        start, finals = np.zeros(n, np.int32), np.zeros(S, np.float32)
        src, dst, il, ol = (np.zeros(A, np.int32) for _ in range(4))
        w = np.zeros(A, np.float32)
        L.check(L.lib().mfa_fst_batch_export(self._h, *[a.ctypes.data_as(C.c_void_p) for a in (so, ao, start, finals, src, dst, il, ol, w)]))
        out = []
        for u in range(n):
            a, b = ao[u], ao[u + 1]
            out.append(Fst(int(start[u]), int(so[u + 1] - so[u]), src[a:b].copy(), il[a:b].copy(), ol[a:b].copy(), dst[a:b].copy(),
                           w[a:b].copy(), finals[so[u]:so[u + 1]].copy()))
        return out

    def equal_align(self, frame_off, seeds, num_retries: int = 10, olabel_counts: Optional[np.ndarray] = None):
        """a11: Kaldi EqualAlign per graph (mfa_equal_align, host).  -> (ali[sum T], words, word_off, num_words, status)."""
        n = self.sizes()[0]
        fo, fop = _host(frame_off, np.int64)
        sd, sdp = _host(np.asarray(seeds, np.uint64) & 0xFFFFFFFF, np.uint32)
        if olabel_counts is None:   # a self-loop-free path of an acyclic training graph crosses each word arc at most once
            olabel_counts = np.asarray([int(np.count_nonzero(f.arc_olabel)) for f in self.export()], np.int64)
        wo = np.zeros(n + 1, np.int64)
        wo[1:] = np.cumsum(olabel_counts)
        ali = np.zeros(max(int(fo[-1]), 1), np.int32)
        words = np.zeros(max(int(wo[-1]), 1), np.int32)
        nw, st = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        L.check(L.lib().mfa_equal_align(self._h, fop, sdp, C.c_int32(num_retries), *[a.ctypes.data_as(C.c_void_p) for a in (ali, words, wo, nw, st)]))
        return ali[:int(fo[-1])], words, wo, nw[:n], st[:n]

    def close(self):
        if self._h:
            L.lib().mfa_fst_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GraphCompiler:
    """Host C++ training-graph compiler (mfa_graph_compiler)."""

    def __init__(self, tm: TransitionModel, tree: ContextDependency, lexicon):
        """lexicon: mfa_b200.lexicon.Lexicon"""
        self.tm, self.tree, self.lexicon = tm, tree, lexicon
        hd, self._keep1 = hmm_desc(tm, tree)
        ld, self._keep2 = lexicon.desc()
        self._h = C.c_void_p()
        L.check(L.lib().mfa_graph_compiler_create(C.byref(hd), C.byref(ld), C.byref(self._h)))

    def compile(self, word_id_seqs: Sequence[Sequence[int]], n_threads: int = 8) -> FstBatch:
        n = len(word_id_seqs)
        off = np.zeros(n + 1, np.int64)
        for i, w in enumerate(word_id_seqs):
            off[i + 1] = off[i] + len(w)
        words = np.ascontiguousarray(np.concatenate([np.asarray(w, np.int32) for w in word_id_seqs]) if n and off[-1] else np.zeros(0, np.int32), dtype=np.int32)
        h = C.c_void_p()
        L.check(L.lib().mfa_graph_compile(self._h, words.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), C.c_int32(n),
                                          C.c_int32(n_threads), C.byref(h)))
        return FstBatch(h)

    def close(self):
        if self._h:
            L.lib().mfa_graph_compiler_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graphs:
    """Decoder-ready packed graphs (mfa_graphs): AddTransitionProbs folded in."""

    def __init__(self, batch: FstBatch, tm: TransitionModel, transition_scale: float = 1.0, self_loop_scale: float = 0.1):
        tid_cost = np.ascontiguousarray(-tm.scaled_transition_log_probs(transition_scale, self_loop_scale), dtype=np.float32)
        tid2pdf = np.ascontiguousarray(np.maximum(tm.tid2pdf, 0), dtype=np.int32)
        self._h = C.c_void_p()
        L.check(L.lib().mfa_graphs_pack(batch._h, tid_cost.ctypes.data_as(C.c_void_p), tid2pdf.ctypes.data_as(C.c_void_p),
                                        C.c_int32(tm.num_tids), C.byref(self._h)))
        self.n_utts = batch.sizes()[0]

    def set_transitions(self, engine: "Engine", model: "DeviceModel", transition_scale: float = 1.0, self_loop_scale: float = 0.1):
        """mfa_graphs_set_transitions: AddTransitionProbs again, on the device, from the model's current transition log-probabilities."""
        L.check(L.lib().mfa_graphs_set_transitions(engine._h, self._h, model._h, C.c_float(transition_scale), C.c_float(self_loop_scale)))

    def offsets(self):
        """(state_off, arc_off, pdf_off): per-utterance prefix offsets of states / arcs / distinct pdfs."""
        so, ao, po = (np.zeros(self.n_utts + 1, np.int64) for _ in range(3))
        L.check(L.lib().mfa_graphs_offsets(self._h, *[a.ctypes.data_as(C.c_void_p) for a in (so, ao, po)]))
        return so, ao, po

    def band_view(self):
        """The band layout the primary Viterbi kernel runs on (mfa_graphs_band_view): dict of per-utterance / per-state / per-arc arrays."""
        so, ao, _ = self.offsets()
        S, A = int(so[-1]), int(ao[-1])
        ok, start, maxback = (np.zeros(self.n_utts, np.int32) for _ in range(3))
        stw, orig = np.zeros(S, np.uint32), np.zeros(S, np.uint16)
        apk, arcid = np.zeros(A, np.uint32), np.zeros(A, np.uint16)
        L.check(L.lib().mfa_graphs_band_view(self._h, *[a.ctypes.data_as(C.c_void_p) for a in (ok, start, maxback, stw, orig, apk, arcid)]))
        return dict(state_off=so, arc_off=ao, band_ok=ok, start=start, maxback=maxback, state_word=stw, orig_state=orig, arc_word=apk,
                    arc_index=arcid)

    def max_words(self) -> np.ndarray:
        out = np.zeros(self.n_utts, np.int32)
        L.check(L.lib().mfa_graphs_max_words(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if self._h:
            L.lib().mfa_graphs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_opts(acoustic_scale=0.1, beam=10.0, retry_beam=40.0, beam_delta=0.5, min_active=20) -> L.AlignOpts:
    return L.AlignOpts(float(acoustic_scale), float(beam), float(retry_beam), float(beam_delta), int(min_active))


class AlignResult:
    """Outputs of a batched alignment call (host numpy arrays or device torch tensors)."""

    def __init__(self, ali, per_frame, words, word_off, num_words, total_like, status, frame_off):
        self.ali, self.per_frame, self.words, self.word_off = ali, per_frame, words, word_off
        self.num_words, self.total_like, self.status, self.frame_off = num_words, total_like, status, frame_off

    def utterance(self, u: int):
        a, b = int(self.frame_off[u]), int(self.frame_off[u + 1])
        w0 = int(self.word_off[u])
        nw = int(self.num_words[u])
        return dict(status=int(self.status[u]), ali=self.ali[a:b], per_frame=self.per_frame[a:b], words=self.words[w0:w0 + nw],
                    like=float(self.total_like[u]))


def _alloc_outputs(n_frames, word_cap, n_utts, device=None):
    if device is None:
        return (np.zeros(n_frames, np.int32), np.zeros(n_frames, np.float32), np.zeros(max(word_cap, 1), np.int32),
                np.zeros(n_utts, np.int32), np.zeros(n_utts, np.float32), np.zeros(n_utts, np.int32))
    import torch
    z = lambda n, dt: torch.zeros(max(int(n), 1), dtype=dt, device=device)
    return (z(n_frames, torch.int32), z(n_frames, torch.float32), z(word_cap, torch.int32), z(n_utts, torch.int32),
            z(n_utts, torch.float32), z(n_utts, torch.int32))


def align_loglikes(engine: Engine, model: DeviceModel, graphs: Graphs, loglikes, frame_off, opts: L.AlignOpts) -> AlignResult:
    """K3 only: frame-major loglikes [sum T, num_pdfs] + packed graphs -> alignments."""
    fo, fop = _host(frame_off, np.int64)
    n = fo.shape[0] - 1
    wo = np.zeros(n + 1, np.int64)
    wo[1:] = np.cumsum(graphs.max_words())
    k, lp, where = _buf(loglikes, np.float32, "loglikes")
    dev = loglikes.device if where == L.MFA_DEVICE else None
    ali, pf, words, nw, tl, st = _alloc_outputs(int(fo[-1]), int(wo[-1]), n, dev)
    ptrs = [_buf(x, dt)[1] for x, dt in ((ali, np.int32), (pf, np.float32), (words, np.int32))]
    p2 = [_buf(x, dt)[1] for x, dt in ((nw, np.int32), (tl, np.float32), (st, np.int32))]
    L.check(L.lib().mfa_align(engine._h, model._h, graphs._h, C.byref(opts), lp, fop, C.c_int32(n), ptrs[0], ptrs[1], ptrs[2],
                              wo.ctypes.data_as(C.c_void_p), p2[0], p2[1], p2[2], C.c_int(where)))
    if where == L.MFA_DEVICE:
        engine._hold(k, ali, pf, words, nw, tl, st)
    return AlignResult(ali, pf, words, wo, nw, tl, st, fo)


def align_feats(engine: Engine, model: DeviceModel, graphs: Graphs, feats, frame_off, opts: L.AlignOpts, gmm_impl: int = 0,
                workspace_bytes: int = 0) -> AlignResult:
    """K2 (per-utterance pdf subsets) + K3 on final features [sum T, dim] (mfa_align_feats): what GmmAligner runs per batch."""
    fo, fop = _host(frame_off, np.int64)
    n = fo.shape[0] - 1
    wo = np.zeros(n + 1, np.int64)
    wo[1:] = np.cumsum(graphs.max_words())
    k, fp, where = _buf(feats, np.float32, "feats")
    dev = feats.device if where == L.MFA_DEVICE else None
    ali, pf, words, nw, tl, st = _alloc_outputs(int(fo[-1]), int(wo[-1]), n, dev)
    ptrs = [_buf(x, dt)[1] for x, dt in ((ali, np.int32), (pf, np.float32), (words, np.int32))]
    p2 = [_buf(x, dt)[1] for x, dt in ((nw, np.int32), (tl, np.float32), (st, np.int32))]
    L.check(L.lib().mfa_align_feats(engine._h, model._h, graphs._h, C.byref(opts), fp, fop, C.c_int32(n), C.c_int32(gmm_impl),
                                    C.c_int64(int(workspace_bytes)), ptrs[0], ptrs[1], ptrs[2], wo.ctypes.data_as(C.c_void_p), p2[0], p2[1], p2[2],
                                    C.c_int(where)))
    if where == L.MFA_DEVICE:
        engine._hold(k, ali, pf, words, nw, tl, st)
    return AlignResult(ali, pf, words, wo, nw, tl, st, fo)


def align_pcm(engine: Engine, model: DeviceModel, graphs: Graphs, pcm, sample_off, utt2spk, n_spk: int, mfcc: L.MfccOpts,
              feat_mode: str = "deltas", lda=None, splice_ctx: int = 3, fmllr=None, cmvn_stats=None, apply_cmvn: bool = True,
              align: Optional[L.AlignOpts] = None, gmm_impl: int = 0, workspace_bytes: int = 0, outputs=None) -> AlignResult:
    """Fused hot path (mfa_align_pcm): PCM -> MFCC -> CMVN -> features -> log-likelihoods -> Viterbi."""
    so, sop = _host(sample_off, np.int64)
    n = so.shape[0] - 1
    fo = frame_offsets(mfcc, so)
    us, usp = _host(utt2spk, np.int32)
    wo = np.zeros(n + 1, np.int64)
    wo[1:] = np.cumsum(graphs.max_words())
    fopts, keep = make_feat_opts(mfcc.num_ceps, feat_mode, lda, splice_ctx, fmllr, cmvn_stats, n_spk)
    po = L.PipelineOpts(mfcc, fopts, align or align_opts(), int(apply_cmvn), int(gmm_impl), int(workspace_bytes))
    k, pp, where = _buf(pcm, np.int16, "pcm")
    dev = pcm.device if where == L.MFA_DEVICE else None
    if outputs is None:
        outputs = _alloc_outputs(int(fo[-1]), int(wo[-1]), n, dev)
    ali, pf, words, nw, tl, st = outputs
    ptrs = [_buf(x, dt)[1] for x, dt in ((ali, np.int32), (pf, np.float32), (words, np.int32))]
    p2 = [_buf(x, dt)[1] for x, dt in ((nw, np.int32), (tl, np.float32), (st, np.int32))]
    L.check(L.lib().mfa_align_pcm(engine._h, model._h, graphs._h, C.byref(po), pp, sop, usp, C.c_int32(n), C.c_int32(n_spk),
                                  fo.ctypes.data_as(C.c_void_p), ptrs[0], ptrs[1], ptrs[2], wo.ctypes.data_as(C.c_void_p),
                                  p2[0], p2[1], p2[2], C.c_int(where)))
    if where == L.MFA_DEVICE:
        engine._hold(k, keep, ali, pf, words, nw, tl, st)
    return AlignResult(ali, pf, words, wo, nw, tl, st, fo)


def align_pcm_from_transcripts(engine: Engine, compiler: GraphCompiler, model: DeviceModel, transcripts, pcm, sample_off, utt2spk, n_spk: int,
                               mfcc: L.MfccOpts, feat_mode: str = "deltas", lda=None, splice_ctx: int = 3, align: Optional[L.AlignOpts] = None,
                               transition_scale: float = 1.0, self_loop_scale: float = 0.1, n_segments: int = 4, n_threads: int = 8,
                               workspace_bytes: int = 0):
    """Single-shot job from transcripts and HOST PCM (what CompileTrainGraphsFunction + AlignFunction do for one job,
    alignment/multiprocessing.py:489-574,791-863): the batch is cut at speaker boundaries into `n_segments` pieces (CMVN is per speaker, so
    a speaker never spans two pieces); the training graphs of piece k + 1 are compiled and packed on host threads WHILE piece k is
    uploaded and aligned -- graph compilation is the host-bound part of a fresh batch, the GPU work hides behind it.
    Utterances of a speaker must be contiguous (MFA orders a job by speaker); otherwise the batch runs as one piece.
    -> (AlignResult with host arrays for the whole batch, dict of wall-clock stage times in ms)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    so = np.ascontiguousarray(sample_off, np.int64)
    us = np.ascontiguousarray(utt2spk, np.int32)
    n = so.shape[0] - 1
    pcm = np.ascontiguousarray(pcm, np.int16)
    # ---- cuts: after a speaker change, near equal shares of the samples
    contiguous = True
    seen = set()
    for u in range(n):
        if u and us[u] != us[u - 1] and int(us[u]) in seen:
            contiguous = False
            break
        seen.add(int(us[u]))
    cuts = [0]
    if contiguous and n_segments > 1:
        for k in range(1, n_segments):
            u = int(np.searchsorted(so, so[-1] * k // n_segments))
            while 0 < u < n and us[u] == us[u - 1]:
                u += 1
            if cuts[-1] < u < n:
                cuts.append(u)
    cuts.append(n)
    tm = compiler.tm
    t = {"compile_pack_ms": [], "align_ms": [], "segments": len(cuts) - 1}

    def prepare(k):
        t0 = time.perf_counter()
        batch = compiler.compile(transcripts[cuts[k]:cuts[k + 1]], n_threads=n_threads)
        graphs = Graphs(batch, tm, transition_scale, self_loop_scale)
        return batch, graphs, 1e3 * (time.perf_counter() - t0)

    parts = []
    t_all = time.perf_counter()
    with ThreadPoolExecutor(max_workers=1) as ex:
        nxt = ex.submit(prepare, 0)
        for k in range(len(cuts) - 1):
            batch, graphs, ms = nxt.result()
            t["compile_pack_ms"].append(ms)
            if k + 2 < len(cuts):
                nxt = ex.submit(prepare, k + 1)
            u0, u1 = cuts[k], cuts[k + 1]
            uniq, local = np.unique(us[u0:u1], return_inverse=True)   # the piece's own speaker numbering
            t0 = time.perf_counter()
            r = align_pcm(engine, model, graphs, pcm[so[u0]:so[u1]], so[u0:u1 + 1] - so[u0], local.astype(np.int32), int(uniq.size), mfcc, feat_mode, lda=lda,
                          splice_ctx=splice_ctx, align=align, workspace_bytes=workspace_bytes)
            t["align_ms"].append(1e3 * (time.perf_counter() - t0))
            parts.append(r)
            t0 = time.perf_counter()
            graphs.close(); batch.close()
            t.setdefault("close_ms", []).append(1e3 * (time.perf_counter() - t0))
    t["loop_ms"] = 1e3 * (time.perf_counter() - t_all)
    t["total_ms"] = 1e3 * (time.perf_counter() - t_all)
    fo = np.zeros(n + 1, np.int64)
    wo = np.zeros(n + 1, np.int64)
    fo[1:] = np.cumsum(np.concatenate([np.diff(r.frame_off) for r in parts])) if parts else 0
    wo[1:] = np.cumsum(np.concatenate([np.diff(r.word_off) for r in parts])) if parts else 0
    cat = lambda name: np.concatenate([getattr(r, name) for r in parts])
    words = np.concatenate([r.words[:int(r.word_off[-1])] for r in parts]) if parts else np.zeros(0, np.int32)
    ali = np.concatenate([r.ali[:int(r.frame_off[-1])] for r in parts])
    pf = np.concatenate([r.per_frame[:int(r.frame_off[-1])] for r in parts])
    return AlignResult(ali, pf, words, wo, cat("num_words"), cat("total_like"), cat("status"), fo), t
