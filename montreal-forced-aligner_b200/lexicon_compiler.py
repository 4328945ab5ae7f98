"""kalpy-shaped front end of the lexicon: ``LexiconCompiler`` as MFA constructs and drives it (reference:
montreal_forced_aligner/models.py:495-511, command_line/align_one.py:118-143, online/alignment.py:44-118,
alignment/multiprocessing.py:1543-1546), ``Pronunciation`` and a pywrapfst-like ``SymbolTable`` (member / find / add_symbol /
read_text / write_text), over mfa_b200.lexicon.Lexicon.  The lexicon FST is never materialised (csrc/graph.cc composes it with the
transcript in closed form), so ``fst`` / ``align_fst`` and the ``load_l_*`` caches of kalpy's class do not exist here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

from ._lib import MfaError
from .lexicon import Lexicon, Pron, make_phone_table, parse_dictionary


class SymbolTable:
    """The subset of pywrapfst.SymbolTable MFA touches: symbol <-> integer, text round trip."""

    def __init__(self, symbols: Optional[Dict[str, int]] = None):
        self._s2i: Dict[str, int] = dict(symbols or {})
        self._i2s: Dict[int, str] = {i: s for s, i in self._s2i.items()}

    def add_symbol(self, symbol: str, key: Optional[int] = None) -> int:
        if symbol in self._s2i:
            return self._s2i[symbol]
        if key is None:
            key = (max(self._i2s) + 1) if self._i2s else 0
        self._s2i[symbol] = key
        self._i2s[key] = symbol
        return key

    def member(self, x) -> bool:
        return (x in self._s2i) if isinstance(x, str) else (x in self._i2s)

    def find(self, x):
        """find(symbol) -> key (-1 if absent); find(key) -> symbol ('' if absent) -- pywrapfst's overload."""
        if isinstance(x, str):
            return self._s2i.get(x, -1)
        return self._i2s.get(int(x), "")

    def get(self, key, default=None):           # dict-style access by integer key (Alignment.generate_ctm)
        return self._i2s.get(key, default)

    def num_symbols(self) -> int:
        return len(self._s2i)

    def __iter__(self):
        return iter(sorted(self._i2s.items()))

    def __len__(self):
        return len(self._s2i)

    def as_dict(self) -> Dict[str, int]:
        return dict(self._s2i)

    def write_text(self, path):
        with open(path, "w", encoding="utf8") as f:
            for i, s in sorted(self._i2s.items()):
                f.write(f"{s} {i}\n")

    @classmethod
    def read_text(cls, path) -> "SymbolTable":
        t = cls()
        with open(path, "r", encoding="utf8") as f:
            for line in f:
                p = line.split()
                if len(p) == 2:
                    t.add_symbol(p[0], int(p[1]))
        return t


@dataclass
class Pronunciation:
    """kalpy.fstext.lexicon.Pronunciation (online/alignment.py:56-66)."""
    orthography: str
    pronunciation: str
    probability: Optional[float] = None
    silence_after_probability: Optional[float] = None
    silence_before_correction: Optional[float] = None
    non_silence_before_correction: Optional[float] = None
    disambiguation: Optional[int] = None


class LexiconCompiler:
    def __init__(self, disambiguation: bool = False, silence_probability: float = 0.5, initial_silence_probability: float = 0.5,
                 final_silence_correction: Optional[float] = None, final_non_silence_correction: Optional[float] = None,
                 silence_word: str = "<eps>", oov_word: str = "<unk>", silence_phone: str = "sil", oov_phone: str = "spn",
                 position_dependent_phones: bool = False, ignore_case: bool = True, phones: Optional[Iterable[str]] = None,
                 word_begin_label: str = "#1", word_end_label: str = "#2"):
        self.disambiguation = disambiguation
        self.silence_probability, self.initial_silence_probability = silence_probability, initial_silence_probability
        self.final_silence_correction, self.final_non_silence_correction = final_silence_correction, final_non_silence_correction
        self.silence_word, self.oov_word, self.silence_phone, self.oov_phone = silence_word, oov_word, silence_phone, oov_phone
        self.position_dependent_phones, self.ignore_case = position_dependent_phones, ignore_case
        self.phones = set(phones or ())
        self._prons: Dict[str, List[Pron]] = {}
        self._phone_table: Optional[SymbolTable] = None
        self._word_table: Optional[SymbolTable] = None
        self._lex: Optional[Lexicon] = None

    # ---- tables (assignable, as MFA does with symbol tables read from a model archive)
    @property
    def phone_table(self) -> SymbolTable:
        if self._phone_table is None:
            sil = [p for p in (self.silence_phone, self.oov_phone) if p]
            sil = list(dict.fromkeys(sil))
            non = sorted(p for p in self.phones if p not in sil)
            self._phone_table = SymbolTable(make_phone_table(non, sil, self.position_dependent_phones))
        return self._phone_table

    @phone_table.setter
    def phone_table(self, table):
        self._phone_table = table if isinstance(table, SymbolTable) else SymbolTable(dict(table))
        self._lex = None

    @property
    def word_table(self) -> SymbolTable:
        if self._word_table is None:
            self._word_table = SymbolTable(self.lexicon.word_table)
        return self._word_table

    @word_table.setter
    def word_table(self, table):
        self._word_table = table if isinstance(table, SymbolTable) else SymbolTable(dict(table))

    @property
    def silence_symbols(self) -> List[int]:
        """Integer ids of every form of the silence phones (GmmAligner.boost_silence, online/alignment.py:106)."""
        names = {self.silence_phone, self.oov_phone}
        return sorted(i for i, s in self.phone_table if s.split("_")[0] in names and s != "<eps>")

    # ---- pronunciations
    def load_pronunciations(self, file_name):
        for w, prons in parse_dictionary(file_name, self.ignore_case).items():
            for pr in prons:
                self._add(w, pr)
        self._lex = None
        self._word_table = None

    def add_pronunciation(self, pron: Pronunciation):
        w = pron.orthography.lower() if self.ignore_case else pron.orthography
        ph = pron.pronunciation.split() if isinstance(pron.pronunciation, str) else list(pron.pronunciation)
        self._add(w, Pron(ph, pron.probability if pron.probability is not None else 1.0, pron.silence_after_probability,
                          pron.silence_before_correction, pron.non_silence_before_correction))
        self._lex = None
        self._word_table = None

    def _add(self, word: str, pr: Pron):
        lst = self._prons.setdefault(word, [])
        if not any(p.phones == pr.phones for p in lst):
            lst.append(pr)
        if self._phone_table is None:
            self.phones.update(pr.phones)

    @property
    def lexicon(self) -> Lexicon:
        """The compiled form the graph compiler takes (rebuilt after pronunciations change)."""
        if self._lex is None:
            self._lex = Lexicon(self._prons, self.phone_table.as_dict(), silence_phone=self.silence_phone, oov_word=self.oov_word,
                                oov_phone=self.oov_phone, silence_probability=self.silence_probability,
                                initial_silence_probability=self.initial_silence_probability,
                                final_silence_correction=self.final_silence_correction,
                                final_non_silence_correction=self.final_non_silence_correction,
                                position_dependent_phones=self.position_dependent_phones, silence_word=self.silence_word)
        return self._lex

    def to_int(self, text: str) -> List[int]:
        return self.lexicon.to_int(text.lower() if self.ignore_case else text)

    def phones_to_pronunciations(self, words: Sequence[int], intervals, transcription: bool = False, text: Optional[str] = None):
        from .export import phones_to_pronunciations
        return phones_to_pronunciations(self.lexicon, words, intervals, transcription=transcription, text=text)

    def clear(self):
        """kalpy frees its FSTs here; there are none to free."""

    # ---- what does not exist in a closed-form compiler
    def _no_fst(self, *a, **k):
        raise MfaError("the B200 graph compiler composes the lexicon in closed form: there is no materialised L.fst to load, write or cache")

    load_l_from_file = load_l_align_from_file = _no_fst

    @property
    def fst(self):
        self._no_fst()

    @property
    def align_fst(self):
        self._no_fst()
