"""Upgrade path of SURVEY.md section 8c: when the real reference stack (kalpy over Kaldi) is importable -- e.g. a baseline/_ref
install on a box that has it -- diff the oracle against it with dither = 0 and report *reference* parity.  In this image kalpy
cannot be installed (no network), so these tests skip and DESIGN.md says "parity unpinned"."""
import numpy as np
import pytest

kalpy = pytest.importorskip("kalpy", reason="kalpy / Kaldi are not installable offline: oracle parity only (DESIGN.md section 2)")

from helpers import gold  # noqa: E402
from oracle import oracle as O  # noqa: E402


def test_mfcc_against_kalpy():
    from kalpy.feat.mfcc import MfccComputer
    pcm = gold()["acoustic_corpus_pcm"]
    mc = MfccComputer(use_energy=False, dither=0.0, energy_floor=0.0, snip_edges=True, sample_frequency=16000, frame_length=25, frame_shift=10,
                      num_mel_bins=23, num_coefficients=13, low_frequency=20, high_frequency=7800, preemphasis_coefficient=0.97,
                      cepstral_lifter=22)
    ref = np.asarray(mc.compute_mfccs(pcm.astype(np.float32)).numpy())
    got = O.mfcc(pcm)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
