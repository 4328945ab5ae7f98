from mfa_b200.kalpy_compat import FeatureArchive  # noqa: F401
