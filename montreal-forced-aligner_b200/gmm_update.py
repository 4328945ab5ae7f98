"""Accumulator container and M-step entry point of the align -> acc-stats -> update loop (rows a10 / N3 of SURVEY.md section 8).

``am.mle_update(gmm_accs, mixup=current_gaussians, power=power)`` of the reference (montreal_forced_aligner/acoustic_modeling/base.py:
319-338 upstream, monophone.py:275-296) runs on the DEVICE here: the accumulators are written into the model's f64 block (or are already
there, all-reduced over NCCL, in the in-memory training loop), ``mfa_model_mle_update`` (csrc/mstep.cu: Kaldi MleDiagGmmUpdate /
MleAmDiagGmmUpdate, SplitByCount + Split, TransitionModel::MleUpdate) re-estimates the model in place and rebuilds the K2 operand images,
and the new parameters are read back only when a file ({it+1}.mdl) has to be written.  There is no host fallback; the numpy restatement
used to check the kernels lives in oracle/mstep_oracle.py (test infrastructure).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .kaldi_io import AmDiagGmm, TransitionModel


class AccumAmDiagGmm:
    """f64 accumulators for all pdfs, packed like the model: occ[G], mean_acc[G,D], var_acc[G,D]."""

    def __init__(self, num_gauss: int, dim: int):
        self.occ = np.zeros(num_gauss)
        self.mean = np.zeros((num_gauss, dim))
        self.var = np.zeros((num_gauss, dim))
        self.tot_like = 0.0
        self.tot_frames = 0.0

    @classmethod
    def init(cls, am: AmDiagGmm) -> "AccumAmDiagGmm":
        return cls(am.NumGauss(), am.dim)

    @classmethod
    def from_dict(cls, d: Dict) -> "AccumAmDiagGmm":
        a = cls(d["occ"].shape[0], d["mean"].shape[1])
        a.occ, a.mean, a.var = np.array(d["occ"], dtype=np.float64), np.array(d["mean"], dtype=np.float64), np.array(d["var"], dtype=np.float64)
        a.tot_like = float(np.asarray(d["like"]).reshape(-1)[0])
        a.tot_frames = float(d["frames"])
        return a

    def Add(self, scale: float, other: "AccumAmDiagGmm"):
        self.occ += scale * other.occ
        self.mean += scale * other.mean
        self.var += scale * other.var
        self.tot_like += scale * other.tot_like
        self.tot_frames += scale * other.tot_frames

    def TotLogLike(self) -> float:
        return self.tot_like

    def TotCount(self) -> float:
        return float(self.occ.sum())


def mle_update(am: AmDiagGmm, acc: AccumAmDiagGmm, mixup: int = 0, power: float = 0.25, min_gaussian_occupancy: float = 10.0,
               min_gaussian_weight: float = 1.0e-5, min_variance: float = 0.001, remove_low_count_gaussians: bool = True,
               perturb_factor: float = 0.01, min_count: float = 20.0, seed: int = 1234, tm: Optional[TransitionModel] = None,
               transition_accs: Optional[np.ndarray] = None, engine=None) -> Tuple[AmDiagGmm, float, float]:
    """kalpy ``AmDiagGmm.mle_update(accs, mixup=, power=, min_gaussian_occupancy=, ...)`` for host-resident accumulators (the file-based
    flow, where jobs hand their statistics back through callbacks): upload -> device M-step -> read back.  With ``tm`` and
    ``transition_accs`` the transition model is re-estimated in the same call (``tm.log_probs`` updated in place).
    Returns (new model, objective improvement, total count)."""
    from . import engine as E
    from .kalpy_compat import get_engine
    eng = engine or get_engine()
    dm = E.DeviceModel(eng, tm, am)
    try:
        dm.acc_zero()
        dm.acc_write(acc.occ, acc.mean, acc.var, trans=transition_accs if tm is not None else None, like=acc.tot_like, frames=acc.tot_frames)
        upd_t = tm is not None and transition_accs is not None
        r = dm.mle_update(mixup=mixup, power=power, min_gaussian_occupancy=min_gaussian_occupancy, min_gaussian_weight=min_gaussian_weight,
                          min_variance=min_variance, remove_low_count_gaussians=remove_low_count_gaussians, perturb_factor=perturb_factor,
                          min_count=min_count, update_transitions=upd_t, seed=seed)
        if upd_t:
            new_am, lp = dm.read(with_transitions=True)
            tm.log_probs = lp.astype(np.float32)
            tm._derive()
        else:
            new_am = dm.read()
    finally:
        dm.close()
    return new_am, float(r["gmm_objf_impr"]), float(r["gmm_count"])
