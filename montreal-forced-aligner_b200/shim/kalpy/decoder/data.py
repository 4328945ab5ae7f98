from mfa_b200.kalpy_compat import FstArchive  # noqa: F401
