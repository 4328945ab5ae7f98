"""Builds BASELINE.json-shaped workloads (synthetic corpus + acoustic model + compiled graphs) on one GPU.

Config 2: triphone LDA+MLLT-shaped GMM-HMM (D=40 via splice+-3 + 40x91 projection, ~4k pdfs, ~40k Gaussians) over
LibriSpeech-shaped 16 kHz audio.  The model's Gaussians are estimated from the engine's own features along the true
segmentation, so beams prune as they do on real data."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import engine as E, synth as SY
from .kaldi_io import AmDiagGmm, ContextDependency, TransitionModel


@dataclass
class Scenario:
    corpus: SY.SynthCorpus
    tm: TransitionModel
    am: AmDiagGmm
    tree: ContextDependency
    lda: Optional[np.ndarray]
    feat_mode: str
    frame_off: np.ndarray
    batch: E.FstBatch
    graphs: E.Graphs
    model: E.DeviceModel
    build_seconds: dict


def build(engine: E.Engine, seconds: float, seed: int = SY.SEED, triphone: bool = True, target_pdfs: int = 4000, gauss_per_pdf: int = 10,
          use_lda: bool = True, n_phones: int = 40, n_words: int = 2000, n_threads: int = 8, synth_device=None, log=None,
          model_seed: Optional[int] = None) -> Scenario:
    """model_seed (multi-rank runs): lexicon, phone spectra, tree, transition model and LDA are drawn from it and are therefore the
    same on every rank; `seed` then only selects the rank's own utterances / speakers.  The Gaussians are still estimated from this
    rank's features: a replicated model additionally needs rank 0's AmDiagGmm broadcast (bench.py does that)."""
    t = {}
    t0 = time.time()
    corpus = SY.make_corpus(seconds, seed=seed, n_phones=n_phones, n_words=n_words, device=synth_device, lexicon_seed=model_seed)
    t["corpus"] = time.time() - t0
    rng = np.random.default_rng((seed if model_seed is None else model_seed) + 1)
    topo = SY.make_topology(corpus.phone_table)
    tree, n_pdfs = SY.make_tree(rng, topo, triphone, target_pdfs)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    lda = SY.random_lda(rng) if use_lda else None
    mode = "lda" if use_lda else "deltas"
    t0 = time.time()
    mo = E.mfcc_opts()
    raw, frame_off = engine.mfcc(corpus.pcm, corpus.sample_off, mo)
    stats = engine.cmvn_stats(raw, frame_off, corpus.utt2spk, corpus.n_spk)
    feats = engine.features(raw, frame_off, mode, lda=lda, cmvn_stats=stats, utt2spk=corpus.utt2spk, n_spk=corpus.n_spk)
    t["features"] = time.time() - t0
    t0 = time.time()
    fp = SY.frame_pdfs_from_truth(corpus, topo, tree, frame_off)
    am = SY.estimate_gmms(feats, fp, n_pdfs, gauss_per_pdf, rng)
    t["model"] = time.time() - t0
    del feats, raw
    t0 = time.time()
    batch = E.GraphCompiler(tm, tree, corpus.lexicon).compile(corpus.transcripts, n_threads=n_threads)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    t["graphs"] = time.time() - t0
    model = E.DeviceModel(engine, tm, am)
    if log:
        log(f"scenario: {corpus.n_utts} utts, {corpus.seconds / 3600:.2f} h, {corpus.n_spk} spk, {n_pdfs} pdfs, {am.NumGauss()} Gaussians, "
            f"dim {am.dim}, graphs {batch.sizes()[1]} states / {batch.sizes()[2]} arcs; build {t}")
    return Scenario(corpus, tm, am, tree, lda, mode, frame_off, batch, graphs, model, t)
