"""Seeded synthetic corpora and acoustic models shaped like BASELINE.json's configurations.

There is no network for datasets or checkpoints, so bench.py and the large parity tests use:
  * audio: 16 kHz int16 "speech-shaped" signals whose spectrum depends on the phone being spoken
    (three formant-like sinusoid clusters + noise per phone, per-speaker formant scaling, 4 Hz envelope),
    LibriSpeech-like utterance lengths (clipped log-normal on [1, 30] s, mean ~12 s);
  * lexicon/text: ``n_phones`` phones (+ sil, spn), ``n_words`` words of 3-8 phones with 1-3 pronunciations;
  * model: monophone or triphone-tree GMM-HMM whose Gaussians are estimated from the synthetic features along the
    true phone segmentation, then split into ``gauss_per_pdf`` components (so beams prune as on real data).
Everything derives from one integer seed (default 1234 = MFA's config.SEED, montreal_forced_aligner/config.py:146).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from .kaldi_io import AmDiagGmm, ContextDependency, HmmState, Topology, TransitionModel
from .lexicon import Lexicon, Pron, make_phone_table

SEED = 1234


@dataclass
class SynthCorpus:
    lexicon: Lexicon
    phone_table: Dict[str, int]
    transcripts: List[List[int]]      # word ids per utterance
    truth: List[np.ndarray]           # per utterance [n_seg, 3] = (phone id, start sample, end sample)
    pcm: np.ndarray                   # int16, concatenated
    sample_off: np.ndarray            # int64 [n_utts+1]
    utt2spk: np.ndarray               # int32 [n_utts]
    n_spk: int
    n_phones: int

    @property
    def n_utts(self) -> int:
        return len(self.transcripts)

    @property
    def seconds(self) -> float:
        return float(self.sample_off[-1]) / 16000.0


def make_lexicon(rng: np.random.Generator, n_phones: int = 40, n_words: int = 2000, position_dependent: bool = False) -> Tuple[Lexicon, Dict[str, int]]:
    phones = [f"p{i:02d}" for i in range(n_phones)]
    pt = make_phone_table(phones, ("sil", "spn"), position_dependent)
    prons: Dict[str, List[Pron]] = {}
    for w in range(n_words):
        n_pr = int(rng.choice([1, 1, 1, 2, 2, 3]))
        base = [phones[int(x)] for x in rng.integers(0, n_phones, size=int(rng.integers(3, 9)))]
        lst = [Pron(list(base), 1.0)]
        for _ in range(n_pr - 1):
            alt = list(base)
            k = int(rng.integers(0, len(alt)))
            alt[k] = phones[int(rng.integers(0, n_phones))]
            if all(alt != p.phones for p in lst):
                lst.append(Pron(alt, float(rng.choice([0.5, 0.8, 1.0]))))
        prons[f"w{w:05d}"] = lst
    lex = Lexicon(prons, pt, position_dependent_phones=position_dependent)
    return lex, pt


def _phone_spectra(rng: np.random.Generator, n_ids: int) -> np.ndarray:
    """Per phone id: 3 formant frequencies (Hz), 3 amplitudes, noise gain -> [n_ids, 7]."""
    f1 = rng.uniform(250, 900, n_ids)
    f2 = rng.uniform(900, 2600, n_ids)
    f3 = rng.uniform(2600, 4200, n_ids)
    a = rng.uniform(0.3, 1.0, (n_ids, 3))
    nz = rng.uniform(0.05, 0.4, n_ids)
    return np.concatenate([np.stack([f1, f2, f3], 1), a, nz[:, None]], 1)


def make_corpus(seconds: float, seed: int = SEED, n_phones: int = 40, n_words: int = 2000, n_spk: Optional[int] = None,
                position_dependent: bool = False, mean_utt_s: float = 12.3, min_utt_s: float = 1.0, max_utt_s: float = 30.0,
                device=None, lexicon_seed: Optional[int] = None) -> SynthCorpus:
    """Generate ~``seconds`` of audio.  ``device`` (a torch device) moves the waveform synthesis onto the GPU.
    ``lexicon_seed`` (multi-rank runs): lexicon and phone spectra come from their own generator, so every rank shares one "language"
    while ``seed`` gives it its own utterances and speakers."""
    rng = np.random.default_rng(seed)
    rng_lex = rng if lexicon_seed is None else np.random.default_rng(lexicon_seed)
    lex, pt = make_lexicon(rng_lex, n_phones, n_words, position_dependent)
    word_ids = [lex.word_table[w] for w in lex.prons if w.startswith("w")]
    sil = pt["sil"]
    n_ids = max(pt.values()) + 1
    spectra = _phone_spectra(rng_lex, n_ids)
    spectra[sil] = [300, 1200, 3000, 0.01, 0.01, 0.01, 0.03]
    if "spn" in pt:
        spectra[pt["spn"]] = [500, 1500, 2500, 0.1, 0.1, 0.1, 0.5]
    # utterance lengths: clipped log-normal
    sigma = 0.6
    mu = math.log(mean_utt_s) - 0.5 * sigma * sigma
    transcripts, truth, lens = [], [], []
    total = 0.0
    sr = 16000
    while total < seconds:
        target = float(np.clip(rng.lognormal(mu, sigma), min_utt_s, max_utt_s))
        target = min(target, max(min_utt_s, seconds - total)) if seconds - total < max_utt_s else target
        segs = []
        t = 0
        def add(ph, dur_s):
            nonlocal t
            n = int(dur_s * sr)
            segs.append((ph, t, t + n))
            t += n
        add(sil, rng.uniform(0.1, 0.3))
        words = []
        while t / sr < target - 0.25 or not words:
            w = int(rng.choice(word_ids))
            words.append(w)
            prs = lex.word_prons_as_phone_ids(w)
            pr = prs[int(rng.integers(0, len(prs)))]
            for ph in pr:
                add(ph, rng.uniform(0.04, 0.14))
            if rng.random() < 0.3:
                add(sil, rng.uniform(0.08, 0.3))
            if len(words) > 200:
                break
        if segs[-1][0] != sil:
            add(sil, rng.uniform(0.1, 0.3))
        transcripts.append(words)
        truth.append(np.asarray(segs, dtype=np.int64))
        lens.append(t)
        total += t / sr
    n_utts = len(transcripts)
    if n_spk is None:
        n_spk = max(1, int(round(total / 3600.0 * 2.5)))
    n_spk = max(1, min(n_spk, n_utts))
    # MFA orders a job's utterances by their "{speaker}-{utterance}" key (corpus/multiprocessing.py:482-497): speakers are contiguous
    utt2spk = np.sort(rng.permutation(n_utts) % n_spk).astype(np.int32)
    spk_scale = rng.uniform(0.88, 1.12, n_spk)
    spk_f0 = rng.uniform(90, 220, n_spk)
    sample_off = np.zeros(n_utts + 1, dtype=np.int64)
    sample_off[1:] = np.cumsum(lens)
    # per-sample phone ids / speaker parameters -> waveform
    N = int(sample_off[-1])
    seg_all = np.concatenate([np.concatenate([s[:, :1], s[:, 1:] + sample_off[u]], 1) for u, s in enumerate(truth)])
    seg_len = (seg_all[:, 2] - seg_all[:, 1]).astype(np.int64)
    seg_spk = np.concatenate([np.full(len(s), utt2spk[u]) for u, s in enumerate(truth)])
    par = spectra[seg_all[:, 0]]  # [n_seg, 7]
    fr = par[:, :3] * spk_scale[seg_spk][:, None]
    if device is not None:
        import torch
        tdev = torch.device(device)
        rep = torch.from_numpy(seg_len).to(tdev)
        g = torch.Generator(device=tdev)
        g.manual_seed(seed)
        out = torch.zeros(N, dtype=torch.float32, device=tdev)
        tt = torch.arange(N, device=tdev, dtype=torch.float32) / sr
        for k in range(3):
            f = torch.repeat_interleave(torch.from_numpy(fr[:, k]).to(tdev, torch.float32), rep)
            a = torch.repeat_interleave(torch.from_numpy(par[:, 3 + k]).to(tdev, torch.float32), rep)
            phase = torch.remainder(torch.cumsum(f.double() * (2 * math.pi / sr), 0), 2 * math.pi).float()
            out += a * torch.sin(phase)
            del f, a, phase
        f0 = torch.repeat_interleave(torch.from_numpy(spk_f0[seg_spk]).to(tdev, torch.float32), rep)
        out *= 0.6 + 0.4 * torch.sin(torch.remainder(torch.cumsum(f0.double() * (2 * math.pi / sr), 0), 2 * math.pi).float())
        del f0
        nz = torch.repeat_interleave(torch.from_numpy(par[:, 6]).to(tdev, torch.float32), rep)
        out += nz * torch.randn(N, device=tdev, generator=g)
        del nz
        out *= 0.75 + 0.25 * torch.sin(2 * math.pi * 4.0 * tt)
        pcm = torch.clamp(out * (0.12 * 32768.0), -32767, 32767).round().to(torch.int16).cpu().numpy()
        del out, tt
    else:
        rep = seg_len
        out = np.zeros(N, dtype=np.float32)
        for k in range(3):
            f = np.repeat(fr[:, k], rep)
            a = np.repeat(par[:, 3 + k], rep).astype(np.float32)
            out += a * np.sin(np.cumsum(f * (2 * math.pi / sr))).astype(np.float32)
        f0 = np.repeat(spk_f0[seg_spk], rep)
        out *= (0.6 + 0.4 * np.sin(np.cumsum(f0 * (2 * math.pi / sr)))).astype(np.float32)
        out += np.repeat(par[:, 6], rep).astype(np.float32) * rng.standard_normal(N).astype(np.float32)
        out *= (0.75 + 0.25 * np.sin(2 * math.pi * 4.0 * np.arange(N) / sr)).astype(np.float32)
        pcm = np.clip(np.round(out * (0.12 * 32768.0)), -32767, 32767).astype(np.int16)
    return SynthCorpus(lex, pt, transcripts, truth, pcm, sample_off, utt2spk, n_spk, n_phones)


# ------------------------------------------------------------------------------------------------ model
def make_topology(phone_table: Dict[str, int], silence_names=("sil", "spn")) -> Topology:
    """MFA-style topology: 3-state Bakis for speech phones, 5-state silence model with skips
    (tests/data/dictionaries/expected/topo in the reference tree)."""
    nph = max(phone_table.values()) + 1
    phones = np.arange(1, nph, dtype=np.int32)
    sil_ids = {i for n, i in phone_table.items() if n.split("_")[0] in silence_names}
    phone2idx = np.full(nph, -1, dtype=np.int32)
    for p in phones:
        phone2idx[p] = 1 if int(p) in sil_ids else 0
    bakis = [HmmState(0, 0, [(0, 0.75), (1, 0.25)]), HmmState(1, 1, [(1, 0.75), (2, 0.25)]), HmmState(2, 2, [(2, 0.75), (3, 0.25)]),
             HmmState(-1, -1, [])]
    silt = [HmmState(0, 0, [(0, 0.25), (1, 0.25), (2, 0.25), (3, 0.25)])]
    for j in (1, 2, 3):
        silt.append(HmmState(j, j, [(1, 0.25), (2, 0.25), (3, 0.25), (4, 0.25)]))
    silt.append(HmmState(4, 4, [(4, 0.75), (5, 0.25)]))
    silt.append(HmmState(-1, -1, []))
    return Topology(phones, phone2idx, [bakis, silt])


def make_tree(rng: np.random.Generator, topo: Topology, triphone: bool, target_pdfs: int) -> Tuple[ContextDependency, int]:
    """N=1: one pdf per (phone, pdf_class).  N=3,P=1: TE on the centre phone -> TE on pdf-class -> random SE questions on the
    left/right phone, grown until ~target_pdfs leaves (silence phones stay context independent)."""
    nph = topo.phone2idx.shape[0]
    if not triphone:
        cd = ContextDependency(1, 0)
        cd.nodes.append((2, 0, 0, 0))
        children = [-1] * nph
        pdf = 0
        for ph in range(1, nph):
            states = topo.states_for(ph)
            tidx = len(cd.nodes)
            cd.nodes.append((2, -1, 0, 0))
            ch = []
            for s in states[:-1]:
                ch.append(len(cd.nodes))
                cd.nodes.append((0, 0, pdf, 0))
                pdf += 1
            cd.tables[tidx] = ch
            children[ph] = tidx
        cd.tables[0] = children
        return cd, pdf
    cd = ContextDependency(3, 1)
    cd.nodes.append((2, 1, 0, 0))
    children = [-1] * nph
    roots = []  # (node index of a CE leaf, is_speech)
    for ph in range(1, nph):
        states = topo.states_for(ph)
        tidx = len(cd.nodes)
        cd.nodes.append((2, -1, 0, 0))
        ch = []
        for s in states[:-1]:
            ch.append(len(cd.nodes))
            cd.nodes.append((0, 0, -1, 0))
            roots.append((len(cd.nodes) - 1, int(topo.phone2idx[ph]) == 0))
        cd.tables[tidx] = ch
        children[ph] = tidx
    cd.tables[0] = children
    leaves = [r for r, _ in roots]
    splittable = [r for r, sp in roots if sp]
    depth = {r: 0 for r in splittable}
    n_leaves = len(leaves)
    while n_leaves < target_pdfs and splittable:
        k = int(rng.integers(0, len(splittable)))
        node = splittable.pop(k)
        d = depth.pop(node)
        key = 0 if rng.random() < 0.5 else 2
        yes = np.sort(rng.choice(np.arange(0, nph), size=int(rng.integers(nph // 4, nph // 2 + 1)), replace=False)).astype(np.int32)
        y, n = len(cd.nodes), len(cd.nodes) + 1
        cd.nodes.append((0, 0, -1, 0))
        cd.nodes.append((0, 0, -1, 0))
        cd.nodes[node] = (1, key, y, n)
        cd.sets[node] = yes
        n_leaves += 1
        if d + 1 < 12:
            for c in (y, n):
                splittable.append(c)
                depth[c] = d + 1
    pdf = 0
    for i, nd in enumerate(cd.nodes):
        if nd[0] == 0:
            cd.nodes[i] = (0, 0, pdf, 0)
            pdf += 1
    return cd, pdf


def make_transition_model(topo: Topology, tree: ContextDependency, n_pdfs: int) -> TransitionModel:
    """All (phone, hmm_state, pdf) tuples the tree can produce, sorted like Kaldi's TransitionModel::ComputeTuples."""
    nph = topo.phone2idx.shape[0]
    tuples = set()

    def leaves_under(n, acc):
        t, key, a, b = tree.nodes[n]
        if t == 0:
            acc.add(a)
        elif t == 1:
            leaves_under(a, acc)
            leaves_under(b, acc)
        else:
            for c in tree.tables[n]:
                if c >= 0:
                    leaves_under(c, acc)

    key_phone = 0 if tree.N == 1 else 1
    root = tree.nodes[tree.root]
    assert root[0] == 2 and root[1] == key_phone
    for ph in range(1, nph):
        sub = tree.tables[tree.root][ph]
        if sub < 0:
            continue
        states = topo.states_for(ph)
        cls_node = tree.nodes[sub]
        assert cls_node[0] == 2 and cls_node[1] == -1
        for hs, st in enumerate(states[:-1]):
            acc = set()
            leaves_under(tree.tables[sub][st.forward_pdf_class], acc)
            for pdf in acc:
                tuples.add((ph, hs, pdf, pdf))
    tl = np.asarray(sorted(tuples), dtype=np.int32)
    ntid = 0
    for ph, hs, _a, _b in tl:
        ntid += len(topo.states_for(int(ph))[int(hs)].transitions)
    lp = np.zeros(ntid + 1, dtype=np.float32)
    k = 1
    for ph, hs, _a, _b in tl:
        for _dst, p in topo.states_for(int(ph))[int(hs)].transitions:
            lp[k] = math.log(p)
            k += 1
    return TransitionModel(topo, tl, lp)


def frame_pdfs_from_truth(corpus: SynthCorpus, topo: Topology, tree: ContextDependency, frame_off: np.ndarray, shift: int = 160,
                          win: int = 400) -> np.ndarray:
    """pdf id per frame from the true segmentation (each phone split evenly over three of its emitting HMM states;
    snip_edges framing).  Tree look-ups are done once per distinct (left, phone, right) triple."""
    out = np.zeros(int(frame_off[-1]), dtype=np.int32)
    tri = tree.N == 3
    nid = int(topo.phone2idx.shape[0])
    ph_all, l_all, r_all = [], [], []
    for segs in corpus.truth:
        ph = segs[:, 0]
        ph_all.append(ph)
        l_all.append(np.concatenate([[0], ph[:-1]]))
        r_all.append(np.concatenate([ph[1:], [0]]))
    ph_c, l_c, r_c = np.concatenate(ph_all), np.concatenate(l_all), np.concatenate(r_all)
    if not tri:
        l_c = np.zeros_like(l_c)
        r_c = np.zeros_like(r_c)
    key = (l_c * nid + ph_c) * nid + r_c
    uk, inv = np.unique(key, return_inverse=True)
    table = np.zeros((len(uk), 3), dtype=np.int32)
    for i, k in enumerate(uk):
        r = int(k % nid); ph = int((k // nid) % nid); l = int(k // (nid * nid))
        sts = topo.states_for(ph)[:-1]
        if len(sts) == 5:
            sts = [sts[0], sts[2], sts[4]]
        table[i] = [tree.lookup([l, ph, r] if tri else [ph], s.forward_pdf_class) for s in sts]
    pdfs_seg_all = table[inv]
    pos = 0
    for u, segs in enumerate(corpus.truth):
        n = len(segs)
        T = int(frame_off[u + 1] - frame_off[u])
        pdfs_seg = pdfs_seg_all[pos:pos + n]
        pos += n
        if T == 0:
            continue
        centers = np.arange(T) * shift + win // 2
        idx = np.clip(np.searchsorted(segs[:, 2], centers, side="right"), 0, n - 1)
        frac = (centers - segs[idx, 1]) / np.maximum(1, segs[idx, 2] - segs[idx, 1])
        sub = np.clip((frac * 3).astype(np.int64), 0, 2)
        out[frame_off[u]:frame_off[u + 1]] = pdfs_seg[idx, sub]
    return out


def estimate_gmms(feats: np.ndarray, frame_pdf: np.ndarray, n_pdfs: int, gauss_per_pdf: int, rng: np.random.Generator,
                  var_floor: float = 1e-2) -> AmDiagGmm:
    """Single-Gaussian ML estimate per pdf from (feats, frame->pdf), then split into ``gauss_per_pdf`` perturbed components."""
    D = feats.shape[1]
    f64 = feats.astype(np.float64)
    cnt = np.bincount(frame_pdf, minlength=n_pdfs).astype(np.float64)
    s1 = np.zeros((n_pdfs, D))
    s2 = np.zeros((n_pdfs, D))
    for d in range(D):
        s1[:, d] = np.bincount(frame_pdf, weights=f64[:, d], minlength=n_pdfs)
        s2[:, d] = np.bincount(frame_pdf, weights=f64[:, d] ** 2, minlength=n_pdfs)
    gmean = f64.mean(0)
    gvar = f64.var(0) + var_floor
    have = cnt >= 8
    mean = np.where(have[:, None], s1 / np.maximum(cnt, 1)[:, None], gmean[None, :] + 0.5 * np.sqrt(gvar)[None, :] * rng.standard_normal((n_pdfs, D)))
    var = np.where(have[:, None], s2 / np.maximum(cnt, 1)[:, None] - mean ** 2, gvar[None, :])
    var = np.maximum(var, var_floor * gvar[None, :])
    M = gauss_per_pdf
    ncomp = np.full(n_pdfs, M, dtype=np.int32) if M == 1 else np.clip(rng.integers(max(1, M // 2), M + M // 2 + 1, n_pdfs), 1, 120).astype(np.int32)
    off = np.zeros(n_pdfs + 1, dtype=np.int32)
    off[1:] = np.cumsum(ncomp)
    G = int(off[-1])
    pdf_of = np.repeat(np.arange(n_pdfs), ncomp)
    mu = mean[pdf_of] + (0.0 if M == 1 else 0.6) * np.sqrt(var[pdf_of]) * rng.standard_normal((G, D))
    vv = var[pdf_of] * (1.0 if M == 1 else rng.uniform(0.5, 1.0, (G, D)))
    w = rng.uniform(0.5, 1.5, G)
    wsum = np.bincount(pdf_of, weights=w, minlength=n_pdfs)
    w = w / wsum[pdf_of]
    inv = 1.0 / vv
    return AmDiagGmm(D, off, w.astype(np.float32), (mu * inv).astype(np.float32), inv.astype(np.float32))


def random_lda(rng: np.random.Generator, out_dim: int = 40, in_dim: int = 91) -> np.ndarray:
    """Random orthonormal-row projection standing in for lda.mat."""
    q, _ = np.linalg.qr(rng.standard_normal((in_dim, in_dim)))
    return np.ascontiguousarray(q[:out_dim], dtype=np.float32)
