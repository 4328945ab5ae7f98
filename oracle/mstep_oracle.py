"""CPU restatement (numpy, float64) of the M-step of the align -> acc-stats -> update loop.  TEST INFRASTRUCTURE ONLY: imported by
tests/ as the checker of the device M-step (montreal-forced-aligner_b200/csrc/mstep.cu); the product path never imports it.

Restates Kaldi gmm/mle-diag-gmm.cc ``MleDiagGmmUpdate``, gmm/mle-am-diag-gmm.cc ``MleAmDiagGmmUpdate``, gmm/am-diag-gmm.cc
``SplitByCount`` / ``GetSplitTargets`` and gmm/diag-gmm.cc ``Split`` as called by the reference at
montreal_forced_aligner/acoustic_modeling/base.py:319-338 (upstream ``acc_stats``):
``am.mle_update(gmm_accs, mixup=current_gaussians, power=power)``.

``SplitByCount`` perturbs means with Gaussian noise, so parity of split models is statistical (hard part 8);
component COUNTS per pdf (GetSplitTargets) are deterministic and compared exactly.
Parity unpinned against real Kaldi (none installable here): pinned by closed-form checks in tests/test_host_logic.py.
"""
from __future__ import annotations

import heapq
import os
import sys
from typing import Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mfa_b200.kaldi_io import AmDiagGmm  # noqa: E402  (container type only)


def _gmm_objf(am: AmDiagGmm, acc) -> float:
    """MlObjective of mle-diag-gmm.cc summed over pdfs: sum_m occ*gconst + mean_acc.means_invvars - 0.5 var_acc.inv_vars."""
    return float((acc.occ * am.gconsts.astype(np.float64)).sum() + (acc.mean * am.means_invvars).sum() - 0.5 * (acc.var * am.inv_vars).sum())


def mle_update(am: AmDiagGmm, acc, mixup: int = 0, power: float = 0.25, min_gaussian_occupancy: float = 10.0,
               min_gaussian_weight: float = 1.0e-5, min_variance: float = 0.001, remove_low_count_gaussians: bool = True,
               perturb_factor: float = 0.01, min_count: float = 20.0, seed: int = 1234) -> Tuple[AmDiagGmm, float, float]:
    """Returns (new model, objective improvement, total count).  Updates means, variances and weights.
    Vectorised over all Gaussians (segment sums per pdf with np.add.reduceat); only mix-up walks pdfs one by one."""
    D, P = am.dim, am.NumPdfs()
    objf_before = _gmm_objf(am, acc)
    off = np.asarray(am.offsets, dtype=np.int64)
    n_per = np.diff(off)
    pdf_of = np.repeat(np.arange(P), n_per)
    occ = acc.occ
    state_occs = np.add.reduceat(occ, off[:-1]) if P else np.zeros(0)
    state_occs[n_per == 0] = 0.0
    prob = np.where(state_occs[pdf_of] > 0, occ / np.where(state_occs[pdf_of] > 0, state_occs[pdf_of], 1.0), 1.0 / n_per[pdf_of])
    upd = (occ > min_gaussian_occupancy) & (prob > min_gaussian_weight)
    w = am.weights.astype(np.float64).copy()
    mu, var = am.means().copy(), am.variances().copy()
    safe = np.where(upd, occ, 1.0)[:, None]
    m_new = acc.mean / safe
    v_new = np.maximum(acc.var / safe - m_new * m_new, min_variance)
    mu[upd], var[upd], w[upd] = m_new[upd], v_new[upd], prob[upd]
    if remove_low_count_gaussians:
        keep = upd.copy()
        none = np.add.reduceat(keep.astype(np.int64), off[:-1]) == 0
        for j in np.nonzero(none)[0]:   # MleDiagGmmUpdate walks the components in order and refuses to remove the only one left:
            keep[off[j + 1] - 1] = True   # the LAST index survives (un-updated)
    else:
        keep = np.ones_like(upd)
        w[~upd] = prob[~upd]
    w, mu, var = w[keep], mu[keep], var[keep]
    new_n = np.add.reduceat(keep.astype(np.int64), off[:-1])
    new_off = np.zeros(P + 1, dtype=np.int64)
    new_off[1:] = np.cumsum(new_n)
    w = w / np.repeat(np.add.reduceat(w, new_off[:-1]), new_n)
    out = AmDiagGmm(D, new_off.astype(np.int32), w.astype(np.float32), (mu / var).astype(np.float32), (1.0 / var).astype(np.float32))
    objf_after = None
    if out.NumGauss() == am.NumGauss():
        objf_after = _gmm_objf(out, acc)
    if mixup and mixup > out.NumGauss():
        out = split_by_count(out, state_occs, mixup, perturb_factor, power, min_count, seed)
    count = acc.TotCount()
    impr = (objf_after - objf_before) if objf_after is not None else float("nan")
    return out, impr, count


def get_split_targets(state_occs: np.ndarray, target_components: int, power: float, min_count: float) -> np.ndarray:
    """am-diag-gmm.cc GetSplitTargets: greedy allocation by occupancy^power / num_components."""
    n = state_occs.shape[0]
    comps = np.ones(n, dtype=np.int64)
    occp = np.power(np.maximum(state_occs, 0.0), power)
    heap = [(-occp[j] / (1 + 1.0e-10), j) for j in range(n)]
    heapq.heapify(heap)
    num = n
    dead = np.zeros(n, dtype=bool)
    while num < target_components and heap:
        negkey, j = heapq.heappop(heap)
        if negkey == 0.0 or dead[j]:
            break
        if (comps[j] + 1) * min_count >= state_occs[j]:
            dead[j] = True
            heapq.heappush(heap, (0.0, j))
        else:
            comps[j] += 1
            num += 1
            heapq.heappush(heap, (-occp[j] / (comps[j] + 1.0e-10), j))
    return comps


def split_by_count(am: AmDiagGmm, state_occs: np.ndarray, target_components: int, perturb_factor: float = 0.01, power: float = 0.25,
                   min_count: float = 20.0, seed: int = 1234) -> AmDiagGmm:
    """AmDiagGmm::SplitByCount -> DiagGmm::Split (heaviest component split, means perturbed by +-perturb*sqrt(var)*randn)."""
    rng = np.random.default_rng(seed)
    targets = get_split_targets(state_occs, target_components, power, min_count)
    mu_all, var_all = am.means(), am.variances()
    ws, mus, vars_, off = [], [], [], [0]
    for j in range(am.NumPdfs()):
        a, b = int(am.offsets[j]), int(am.offsets[j + 1])
        if targets[j] <= b - a:   # nothing to split: copy the pdf as it is (no random draws are consumed, as in DiagGmm::Split)
            ws.append(am.weights[a:b].astype(np.float64)); mus.append(mu_all[a:b]); vars_.append(var_all[a:b])
            off.append(off[-1] + (b - a))
            continue
        w = list(am.weights[a:b].astype(np.float64))
        mu = [m.copy() for m in mu_all[a:b]]
        var = [v.copy() for v in var_all[a:b]]
        while len(w) < targets[j]:
            k = int(np.argmax(w))
            w[k] *= 0.5
            w.append(w[k])
            r = rng.standard_normal(am.dim) * np.sqrt(var[k]) * perturb_factor
            mu.append(mu[k] + r)
            mu[k] = mu[k] - r
            var.append(var[k].copy())
        ws.append(np.asarray(w)); mus.append(np.asarray(mu)); vars_.append(np.asarray(var))
        off.append(off[-1] + len(w))
    w = np.concatenate(ws); mu = np.concatenate(mus); var = np.concatenate(vars_)
    return AmDiagGmm(am.dim, np.asarray(off, np.int32), w.astype(np.float32), (mu / var).astype(np.float32), (1.0 / var).astype(np.float32))
