from mfa_b200.kalpy_compat import KalpyUtterance as Utterance, Segment  # noqa: F401
