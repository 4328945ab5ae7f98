from mfa_b200.kalpy_compat import MfccComputer  # noqa: F401
