from mfa_b200.kalpy_compat import TrainingGraphCompiler  # noqa: F401
