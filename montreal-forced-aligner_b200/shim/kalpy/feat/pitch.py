from mfa_b200.kalpy_compat import PitchComputer  # noqa: F401  (imports; refuses construction: pitch is outside the hot path)
