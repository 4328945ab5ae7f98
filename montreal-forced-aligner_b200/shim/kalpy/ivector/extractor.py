from mfa_b200.kalpy_compat import IvectorExtractor  # noqa: F401  (imports; refuses construction)
