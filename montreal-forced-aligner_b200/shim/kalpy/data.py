from mfa_b200.kalpy_compat import KaldiMapping, MatrixArchive, Segment  # noqa: F401
