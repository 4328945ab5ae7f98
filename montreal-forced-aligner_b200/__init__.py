"""B200-native engine for MFA's alignment hot path (MFCC+CMVN -> GMM log-likelihoods -> Viterbi,
plus GMM accumulator statistics).  Import as ``mfa_b200`` (see ../mfa_b200.py)."""
__version__ = "0.1.0"


def install_kalpy_shim():
    """Puts ``<package>/shim`` in front of ``sys.path`` so that ``import kalpy...`` (as MFA's modules do) resolves to this engine's
    kalpy-shaped classes.  Refuses when a real kalpy is already imported."""
    import os
    import sys
    shim = os.path.join(_pkg_dir if "_pkg_dir" in globals() else os.path.dirname(os.path.abspath(__file__)), "shim")
    mod = sys.modules.get("kalpy")
    if mod is not None and not getattr(mod, "__mfa_b200_shim__", False):
        raise ImportError("a real kalpy is already imported; start the interpreter with the shim directory on PYTHONPATH instead")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    return shim
