from mfa_b200.kalpy_compat import CmvnComputer  # noqa: F401
