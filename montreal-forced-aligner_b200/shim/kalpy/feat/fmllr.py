from mfa_b200.kalpy_compat import FmllrComputer  # noqa: F401
