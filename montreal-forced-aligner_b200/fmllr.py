"""fMLLR estimation between the two alignment passes.

Mirrors kalpy's ``FmllrComputer`` as MFA uses it in CalcFmllrFunction._run (montreal_forced_aligner/corpus/features.py:460-548;
options ``fmllr_update_type / silence_weight / acoustic_scale`` from :759-766).  The O(frames) part -- posterior-weighted
per-speaker statistics beta, K, G_d -- runs on the GPU (``mfa_fmllr_acc``, csrc/fmllr.cu), and so does the O(speakers) transform
update (``mfa_fmllr_update``, one CTA per speaker).  The same update is restated here in float64 numpy, batched over speakers,
as the host cross-check of that kernel; it is Kaldi transform/fmllr-diag-gmm.cc ``ComputeFmllrMatrixDiagGmmFull``: 40 sweeps of
the row update  w_d = (alpha c_d + k_d) G_d^-1  where c_d is the cofactor row of A and alpha the better root of the quadratic
(``FmllrInnerUpdate``), accepted only if the auxiliary function did not decrease; speakers with beta <= min_count (500) keep the
unit transform (gmm-est-fmllr writes it).  The inverse of A is carried across row updates with the Sherman-Morrison identity and
recomputed once per sweep (Kaldi re-inverts for every row; the difference is rounding-level in f64).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import engine as E, kaldi_io as K
from ._lib import MfaError


def unpack_stats(stats: np.ndarray, D: int):
    """[S, size] -> beta [S], K [S, D, D+1], G [S, D, D+1, D+1] (symmetric, from the packed lower triangles)."""
    stats = np.asarray(stats, np.float64)
    S, D1 = stats.shape[0], D + 1
    NP = D1 * (D1 + 1) // 2
    beta = stats[:, 0]
    Kmat = stats[:, 1:1 + D * D1].reshape(S, D, D1)
    Gp = stats[:, 1 + D * D1:].reshape(S, D, NP)
    il, jl = np.tril_indices(D1)
    G = np.zeros((S, D, D1, D1))
    G[:, :, il, jl] = Gp
    G[:, :, jl, il] = Gp
    return beta, Kmat, G


def aux_function(W: np.ndarray, beta: np.ndarray, Kmat: np.ndarray, G: np.ndarray) -> np.ndarray:
    """FmllrAuxFuncDiagGmm per speaker: beta log|det A| + tr(W K^T) - 1/2 sum_d w_d G_d w_d^T."""
    D = W.shape[1]
    _, logdet = np.linalg.slogdet(W[:, :, :D])
    quad = np.einsum("sdi,sdij,sdj->s", W, G, W)
    return beta * logdet + np.einsum("sdi,sdi->s", W, Kmat) - 0.5 * quad


def compute_transforms(stats: np.ndarray, D: int, num_iters: int = 40, min_count: float = 500.0):
    """-> (W [S, D, D+1] float32, objective improvement [S], count [S]) from per-speaker statistics."""
    beta, Kmat, G = unpack_stats(stats, D)
    S, D1 = beta.shape[0], D + 1
    W_out = np.tile(np.eye(D, D1), (S, 1, 1))
    impr_out = np.zeros(S)
    ok = np.nonzero(beta > min_count)[0]
    if ok.size:
        b, Kk, Gg = beta[ok], Kmat[ok], G[ok]
        invG = np.ascontiguousarray(np.linalg.inv(Gg).transpose(1, 0, 2, 3))   # [D, S', D+1, D+1]
        Kd = np.ascontiguousarray(Kk.transpose(1, 0, 2))                        # [D, S', D+1]
        W0 = W_out[ok].copy()
        W = W0.copy()
        for _ in range(num_iters):
            Ainv = np.linalg.inv(W[:, :, :D])
            for d in range(D):
                c = np.zeros((ok.size, D1))
                c[:, :D] = Ainv[:, :, d]                 # row d of inv(A^T): the cofactor row up to the determinant
                cg = np.matmul(invG[d], c[:, :, None])[:, :, 0]
                e1 = (cg * c).sum(1)
                e2 = (cg * Kd[d]).sum(1)
                disc = np.sqrt(e2 * e2 + 4.0 * e1 * b)
                a1, a2 = (-e2 + disc) / (2.0 * e1), (-e2 - disc) / (2.0 * e1)
                f1 = b * np.log(np.abs(a1 * e1 + e2)) - 0.5 * a1 * a1 * e1
                f2 = b * np.log(np.abs(a2 * e1 + e2)) - 0.5 * a2 * a2 * e1
                alpha = np.where(f1 > f2, a1, a2)
                w_new = np.matmul(invG[d], (alpha[:, None] * c + Kd[d])[:, :, None])[:, :, 0]
                delta = w_new[:, :D] - W[:, d, :D]
                W[:, d] = w_new
                # Sherman-Morrison: A' = A + e_d delta^T
                u = Ainv[:, :, d].copy()                  # Ainv e_d
                v = np.matmul(delta[:, None, :], Ainv)[:, 0, :]   # delta^T Ainv
                denom = 1.0 + v[:, d]
                Ainv -= u[:, :, None] * (v / denom[:, None])[:, None, :]
        old = aux_function(W0, b, Kk, Gg).astype(np.float32).astype(np.float64)
        new = aux_function(W, b, Kk, Gg).astype(np.float32).astype(np.float64)
        impr = new - old
        approx_equal = np.abs(new - old) <= 0.001 * (np.abs(new) + np.abs(old))
        accept = ~((impr < 0.0) & ~approx_equal)
        W_out[ok[accept]] = W[accept]
        impr_out[ok[accept]] = impr[accept]
    return W_out.astype(np.float32), impr_out, beta


def compute_transforms_device(engine: "E.Engine", stats, D: int, num_iters: int = 40, min_count: float = 500.0):
    """The same update on the GPU (mfa_fmllr_update, one CTA per speaker); `stats` numpy or torch cuda f64."""
    return engine.fmllr_update(stats, D, num_iters, min_count)


def compose_transforms(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Kaldi ComposeTransforms(a, b, b_is_affine=true): apply b first, then a.  [D, D+1] x [D, D+1] -> [D, D+1]."""
    D = a.shape[0]
    c = np.zeros((D, D + 1), np.float32)
    c[:, :D] = a[:, :D] @ b[:, :D]
    c[:, D] = a[:, :D] @ b[:, D] + a[:, D]
    return c


class FmllrComputer:
    """kalpy.gmm... FmllrComputer as constructed at corpus/features.py:508-514:
    ``FmllrComputer(ali_model_path, model_path, silence_phones, spk2utt=..., fmllr_update_type=, silence_weight=, acoustic_scale=)``."""

    def __init__(self, alignment_model_path, acoustic_model_path, silence_phones: Sequence[int], spk2utt: Optional[Dict[str, List[str]]] = None,
                 fmllr_update_type: str = "full", silence_weight: float = 0.0, acoustic_scale: float = 0.1, min_count: float = 500.0,
                 num_iters: int = 40):
        from .kalpy_compat import get_engine
        if fmllr_update_type != "full":
            raise MfaError("only fmllr_update_type='full' is implemented (MFA's default, corpus/features.py:630)")
        self.alignment_model_path, self.acoustic_model_path = str(alignment_model_path), str(acoustic_model_path)
        self.silence_phones = sorted(int(p) for p in silence_phones)
        self.spk2utt = spk2utt or {}
        self.silence_weight, self.acoustic_scale, self.min_count, self.num_iters = silence_weight, acoustic_scale, min_count, num_iters
        self.engine = get_engine()
        self.transition_model, self.acoustic_model = K.read_gmm_model(self.acoustic_model_path)
        self._dm = E.DeviceModel(self.engine, self.transition_model, self.acoustic_model)
        self.two_models = self.alignment_model_path != self.acoustic_model_path
        self._dm_post = None
        if self.two_models:
            tm2, am2 = K.read_gmm_model(self.alignment_model_path)
            self._dm_post = E.DeviceModel(self.engine, tm2, am2)
        tm = self.transition_model
        sil = np.isin(tm.tid2phone, np.asarray(self.silence_phones, tm.tid2phone.dtype))
        self.tid_weight = np.where(sil, np.float32(silence_weight), np.float32(1.0)).astype(np.float32)
        self.tid_weight[0] = 0.0
        self.num_error = 0   # utterances skipped because the alignment length differs from the frame count

    def compute_stats(self, feats, ali, frame_off, utt2spk, n_spk: int):
        return self._dm.fmllr_acc(feats, ali, frame_off, utt2spk, n_spk, tid_weight=self.tid_weight, post_model=self._dm_post)

    def compute_transforms(self, stats, impl: str = "device"):
        """impl 'device': mfa_fmllr_update (CUDA, default); 'host': the float64 numpy restatement above (cross-check)."""
        if impl == "host":
            return compute_transforms(np.asarray(stats), self.acoustic_model.dim, self.num_iters, self.min_count)
        W, impr, count = compute_transforms_device(self.engine, stats, self.acoustic_model.dim, self.num_iters, self.min_count)
        return np.asarray(W), impr, count

    def export_transforms(self, file_name, feature_archive, alignment_archive, previous_transform_archive=None, callback: Optional[Callable] = None,
                          max_frames: int = 4_000_000):
        """trans.ark keyed by speaker.  Statistics over the archive's (already transformed, when it carries transforms) features;
        with ``previous_transform_archive`` the new estimate is composed on top of the previous one (Kaldi compose-transforms)."""
        spks = [s for s in self.spk2utt if any(u in alignment_archive for u in self.spk2utt[s])]
        have = set(feature_archive.keys)
        out: Dict[str, np.ndarray] = {}
        i = 0
        while i < len(spks):
            group, keys, u2s, total = [], [], [], 0
            while i < len(spks) and (not group or total < max_frames):
                s = spks[i]
                ks = [u for u in self.spk2utt[s] if u in alignment_archive and u in have]
                alis = [np.asarray(alignment_archive[u].alignment, np.int32) for u in ks]
                total += sum(len(a) for a in alis)
                keys.extend(zip(ks, alis)); u2s.extend([len(group)] * len(ks)); group.append(s)
                i += 1
            feats, fo = feature_archive.batch([k for k, _ in keys])
            ali = np.zeros(int(fo[-1]), np.int32)
            for j, (_, a) in enumerate(keys):
                if len(a) != int(fo[j + 1] - fo[j]):   # gmm-est-fmllr skips utterances whose alignment length is wrong
                    self.num_error += 1
                    continue                           # transition-id 0 = frame ignored by the kernel
                ali[fo[j]:fo[j + 1]] = a
            stats = self.compute_stats(np.ascontiguousarray(feats, np.float32), ali, fo, np.asarray(u2s, np.int32), len(group))
            W, impr, count = self.compute_transforms(stats)
            for j, s in enumerate(group):
                w = W[j]
                if previous_transform_archive is not None and s in previous_transform_archive:
                    w = compose_transforms(w, np.asarray(previous_transform_archive[s], np.float32))
                out[s] = w
                if callback:
                    callback((s, float(impr[j]), float(count[j])))
        with K.ArkWriter(file_name) as w:
            for s in out:
                w.write_matrix(str(s), out[s].astype(np.float32))
        return out
