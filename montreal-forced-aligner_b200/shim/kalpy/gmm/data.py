from mfa_b200.export import HierarchicalCtm  # noqa: F401
from mfa_b200.kalpy_compat import Alignment, AlignmentArchive, CtmInterval, TranscriptionArchive  # noqa: F401
