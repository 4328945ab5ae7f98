from mfa_b200.kalpy_compat import GmmStatsAccumulator, TwoFeatsStatsAccumulator  # noqa: F401
