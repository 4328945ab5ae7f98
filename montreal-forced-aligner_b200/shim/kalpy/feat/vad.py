from mfa_b200.kalpy_compat import VadComputer  # noqa: F401  (imports; refuses construction)
