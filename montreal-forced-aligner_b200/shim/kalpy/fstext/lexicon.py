from mfa_b200.export import HierarchicalCtm  # noqa: F401
from mfa_b200.lexicon_compiler import LexiconCompiler, Pronunciation, SymbolTable  # noqa: F401
