"""B200-native engine for MFA's alignment hot path (MFCC+CMVN -> GMM log-likelihoods -> Viterbi,
plus GMM accumulator statistics).  Import as ``mfa_b200`` (see ../mfa_b200.py)."""
__version__ = "0.1.0"
