from mfa_b200.kalpy_compat import generate_read_specifier, generate_write_specifier, kalpy_logger, read_kaldi_object  # noqa: F401
