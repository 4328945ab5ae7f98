#!/usr/bin/env python
"""End-to-end example on a synthetic corpus: what `mfa align` does on its hot path, through the MFA-shaped job functions of this repo.

  wav files -> MfccFunction -> calc_cmvn -> FinalFeatureFunction -> CompileTrainGraphsFunction
            -> align_utterances (pass 1, speaker independent) -> calc_fmllr -> align_utterances (pass 2, fMLLR features)
            -> export_textgrids

Needs a B200 (there is no CPU fallback).  Usage:  python examples/two_pass_alignment.py [output_dir] [seconds_of_audio]
"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mfa_b200 import kaldi_io as K, kalpy_compat as KC, mfa_functions as MF, export as X  # noqa: E402


def main():
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(tempfile.mkdtemp(prefix="mfa_b200_example_"))
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    run(out, seconds)


def run(out: Path, seconds: float, quiet: bool = False, n_jobs: int = 2, threads: bool = True) -> dict:
    """The flow on `seconds` of synthetic audio under `out`; returns wall time, x real time and per-stage seconds (bench.py's file_flow)."""
    say = (lambda *a: None) if quiet else print
    out.mkdir(parents=True, exist_ok=True)
    # a small synthetic "language", corpus and triphone LDA model; the model's Gaussians are estimated from the engine's own features
    from mfa_b200 import scenario as SC
    built = SC.build(KC.get_engine(), seconds, seed=3, triphone=True, target_pdfs=200, gauss_per_pdf=2, use_lda=True, n_phones=10, n_words=60)
    sc = {"corpus": built.corpus, "tm": built.tm, "am": built.am, "tree": built.tree, "lda": built.lda}
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    split, work, wavs = out / "split", out / "work", out / "wav"
    for d in (split, work, wavs):
        d.mkdir(exist_ok=True)
    utts = []
    for u in range(c.n_utts):
        wav = wavs / f"utt{u:04d}.wav"
        K.write_wav_int16(wav, c.pcm[c.sample_off[u]:c.sample_off[u + 1]])
        utts.append(MF.Utterance(u, int(c.utt2spk[u]), str(wav), " ".join(c.lexicon.id2word[w] for w in c.transcripts[u]),
                                 duration=(c.sample_off[u + 1] - c.sample_off[u]) / 16000.0))
    K.write_gmm_model(work / "final.mdl", tm, am)
    K.write_tree(work / "tree", sc["tree"])
    K.write_matrix_file(work / "lda.mat", sc["lda"])
    jobs = MF.assign_jobs(utts, n_jobs, split)
    nthr = n_jobs if threads else 1   # MFA's USE_THREADING: the jobs of a stage as threads, each on its own engine
    t0 = time.time()
    stages = {}

    def timed(name, fn):
        t = time.time()
        r = fn()
        stages[name] = round(time.time() - t, 3)
        return r
    mc = KC.MfccComputer(use_energy=False, dither=0.0, snip_edges=True)
    timed("mfcc", lambda: list(MF.run_kaldi_function(MF.MfccFunction, [MF.MfccArguments(j.id, j, None, split, mc) for j in jobs], num_threads=nthr)))
    timed("cmvn", lambda: MF.calc_cmvn(jobs, split))
    timed("final_features", lambda: list(MF.run_kaldi_function(MF.FinalFeatureFunction, [MF.FinalFeatureArguments(j.id, j, None, split) for j in jobs], num_threads=nthr)))
    lex = {1: c.lexicon}
    timed("compile_graphs", lambda: list(MF.run_kaldi_function(
        MF.CompileTrainGraphsFunction, [MF.CompileTrainGraphsArguments(j.id, j, None, work, lex, work / "tree", work / "final.mdl") for j in jobs], num_threads=nthr)))
    opts = dict(transition_scale=1.0, acoustic_scale=0.1, self_loop_scale=0.1, beam=10, retry_beam=40, boost_silence=1.0)
    score1, failed1 = timed("align_pass1", lambda: MF.align_utterances(jobs, work, work / "final.mdl", opts, num_threads=nthr))
    sil = [c.lexicon.phone_table["sil"]]
    fm = timed("fmllr", lambda: MF.calc_fmllr(jobs, work, work / "final.mdl", work / "final.mdl", dict(silence_weight=0.0), sil))
    score2, failed2 = timed("align_pass2", lambda: MF.align_utterances(jobs, work, work / "final.mdl", opts, num_threads=nthr))
    written = timed("textgrids", lambda: MF.export_textgrids(jobs, work, work / "final.mdl", lex, out / "aligned"))
    dt = time.time() - t0
    say(f"{c.n_utts} utterances / {c.seconds:.0f} s of audio, {c.n_spk} speakers; files under {out}")
    say(f"pass 1: mean log-likelihood per utterance {score1:.1f} ({failed1} failed); fMLLR for {len(fm)} speakers "
          f"(mean objective improvement per frame {np.mean([v[0] / max(v[1], 1) for v in fm.values()]):.3f}); "
          f"pass 2: {score2:.1f} ({failed2} failed)")
    first = sorted(p for p in written.values() if p is not None)[0]
    tiers = X.read_textgrid(first)
    say(f"{len(written)} TextGrids in {out / 'aligned'}; {first.name}: words = {[e[2] for e in tiers['words'] if e[2]][:8]} ...")
    say(f"wall time incl. file I/O and graph compilation: {dt:.2f} s = {c.seconds / dt:.0f} x real time; per stage (s): {stages}")
    return {"audio_s": float(c.seconds), "utterances": int(c.n_utts), "wall_s": dt, "xRT": c.seconds / dt, "stages_s": stages,
            "failed_pass2": int(failed2), "textgrids": len(written), "jobs": n_jobs, "jobs_as_threads": bool(threads)}


if __name__ == "__main__":
    main()
