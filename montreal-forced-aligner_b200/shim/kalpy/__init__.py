"""Import shim: the kalpy module paths MFA imports (``from kalpy.gmm.align import GmmAligner`` ...) served by the B200 engine.

Put this directory's parent (``<repo>/montreal-forced-aligner_b200/shim``) in front of ``sys.path`` -- ``mfa_b200.install_kalpy_shim()``
does that -- INSTEAD of installing kalpy: MFA's hot-path modules (alignment/multiprocessing.py:30-60, corpus/features.py:13-23,
online/alignment.py:8-14, acoustic_modeling/monophone.py:11-15, command_line/align_one.py:8-11, db.py:15-18, models.py:25-29) then
resolve every kalpy name to mfa_b200.kalpy_compat.  Names outside the alignment path (pitch, VAD, i-vectors, transcription archives)
import fine and fail when constructed.
"""
__mfa_b200_shim__ = True
__version__ = "0.6.7+mfa_b200"
