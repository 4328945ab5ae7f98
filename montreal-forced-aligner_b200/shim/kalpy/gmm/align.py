from mfa_b200.kalpy_compat import GmmAligner, gmm_align_equal  # noqa: F401
