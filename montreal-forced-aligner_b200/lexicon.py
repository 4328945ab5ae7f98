"""Pronunciation lexicon for the training-graph compiler.

Stands in for kalpy ``LexiconCompiler`` as MFA constructs it (reference:
montreal_forced_aligner/dictionary/multispeaker.py:443-470): silence_probability,
initial_silence_probability, final corrections, optional per-pronunciation probabilities
(``word  prob  sil_after  sil_before_corr  nonsil_before_corr  phones...``), optional
position-dependent phones (``_B _E _I _S``), OOV word -> ``spn``.  The lexicon FST itself is never
materialised: the graph compiler (csrc/graph.cc) composes it with the linear transcript in closed form.
Layout reference: tests/data/dictionaries/expected/lexicon.text.fst in the reference tree.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


@dataclass
class Pron:
    phones: List[str]
    prob: float = 1.0
    sil_after: Optional[float] = None
    sil_before_corr: Optional[float] = None
    nonsil_before_corr: Optional[float] = None


def _cost(p: float) -> float:
    return -math.log(p) if p > 0 else 1.0e10


class Lexicon:
    def __init__(self, prons: Dict[str, List[Pron]], phone_table: Dict[str, int], silence_phone: str = "sil", oov_word: str = "<unk>",
                 oov_phone: str = "spn", silence_probability: float = 0.5, initial_silence_probability: float = 0.5,
                 final_silence_correction: Optional[float] = None, final_non_silence_correction: Optional[float] = None,
                 position_dependent_phones: bool = False, silence_word: str = "<eps>"):
        self.phone_table = dict(phone_table)
        self.position_dependent_phones = position_dependent_phones
        self.silence_phone, self.oov_word, self.oov_phone, self.silence_word = silence_phone, oov_word, oov_phone, silence_word
        self.silence_probability = silence_probability
        self.initial_silence_probability = initial_silence_probability
        self.final_silence_correction = final_silence_correction
        self.final_non_silence_correction = final_non_silence_correction
        # pronunciations using phones the model does not know are dropped (their words fall back to the OOV word)
        def known(pr):
            try:
                for ph in pr.phones:
                    self.phone_id(ph, "I")
                return True
            except KeyError:
                return False
        self.prons = {}
        for w, ps in prons.items():
            ps = [pr for pr in ps if known(pr)]
            if ps:
                self.prons[w] = ps
        if oov_word not in self.prons:
            self.prons[oov_word] = [Pron([oov_phone])]
        # word table: <eps>=0 then sorted words (MFA's words.txt convention: <eps> 0, specials, words...)
        self.word_table: Dict[str, int] = {silence_word: 0}
        for w in sorted(self.prons):
            if w not in self.word_table:
                self.word_table[w] = len(self.word_table)
        self.id2word = {i: w for w, i in self.word_table.items()}
        self._build()

    def phone_id(self, ph: str, pos: str) -> int:
        if self.position_dependent_phones:
            name = f"{ph}_{pos}"
            if name in self.phone_table:
                return self.phone_table[name]
            raise KeyError(f"phone {name!r} not in the phone table")
        if ph not in self.phone_table:
            raise KeyError(f"phone {ph!r} not in the phone table")
        return self.phone_table[ph]

    def _build(self):
        nw = len(self.word_table)
        wpo = [0]
        ppo = [0]
        phones: List[int] = []
        pcost, sa, nsa, sb, nsb = [], [], [], [], []
        psil = self.silence_probability
        for wid in range(nw):
            w = self.id2word[wid]
            for pr in self.prons.get(w, []):
                n = len(pr.phones)
                for k, ph in enumerate(pr.phones):
                    pos = "S" if n == 1 else ("B" if k == 0 else ("E" if k == n - 1 else "I"))
                    phones.append(self.phone_id(ph, pos))
                ppo.append(len(phones))
                pcost.append(_cost(pr.prob) if pr.prob is not None else 0.0)
                p_after = pr.sil_after if pr.sil_after is not None else psil
                sa.append(_cost(p_after))
                nsa.append(_cost(1.0 - p_after))
                sb.append(_cost(pr.sil_before_corr) if pr.sil_before_corr else 0.0)
                nsb.append(_cost(pr.nonsil_before_corr) if pr.nonsil_before_corr else 0.0)
            wpo.append(len(pcost))
        self._arrs = [np.asarray(wpo, np.int32), np.asarray(ppo, np.int32), np.asarray(phones if phones else [0], np.int32),
                      np.asarray(pcost if pcost else [0], np.float32), np.asarray(sa if sa else [0], np.float32),
                      np.asarray(nsa if nsa else [0], np.float32), np.asarray(sb if sb else [0], np.float32),
                      np.asarray(nsb if nsb else [0], np.float32)]

    def desc(self):
        p = [a.ctypes.data_as(C.c_void_p).value for a in self._arrs]
        pis = self.initial_silence_probability
        fs = _cost(self.final_silence_correction) if self.final_silence_correction else 0.0
        fn = _cost(self.final_non_silence_correction) if self.final_non_silence_correction else 0.0
        d = L.LexiconDesc(len(self.word_table), p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                          self.phone_id(self.silence_phone, "S") if False else self.phone_table[self.silence_phone],
                          _cost(self.silence_probability), _cost(1.0 - self.silence_probability), _cost(pis), _cost(1.0 - pis), fs, fn)
        return d, self._arrs

    def to_int(self, text: str) -> List[int]:
        """Transcript -> word ids (unknown words -> the OOV word)."""
        oov = self.word_table[self.oov_word]
        return [self.word_table.get(w, oov) if w in self.prons else oov for w in text.split()]

    def word_prons_as_phone_ids(self, wid: int) -> List[List[int]]:
        a = self._arrs
        out = []
        for pr in range(a[0][wid], a[0][wid + 1]):
            out.append([int(x) for x in a[2][a[1][pr]:a[1][pr + 1]]])
        return out


_NUM = re.compile(r"^-?\d+(\.\d+)?([eE]-?\d+)?$")


def parse_dictionary(path, ignore_case: bool = True) -> Dict[str, List[Pron]]:
    """MFA dictionary text: ``word [prob [sil_after sil_before_corr nonsil_before_corr]] phone...`` (tab or space separated)."""
    out: Dict[str, List[Pron]] = {}
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            parts = re.split(r"\s+", line)
            w = parts[0].lower() if ignore_case else parts[0]
            rest = parts[1:]
            nums = []
            while rest and _NUM.match(rest[0]) and len(nums) < 4:
                nums.append(float(rest.pop(0)))
            if not rest:
                continue
            pr = Pron(rest)
            if len(nums) >= 1:
                pr.prob = nums[0]
            if len(nums) == 4:
                pr.sil_after, pr.sil_before_corr, pr.nonsil_before_corr = nums[1], nums[2], nums[3]
            lst = out.setdefault(w, [])
            if not any(p.phones == pr.phones for p in lst):
                lst.append(pr)
    return out


def make_phone_table(non_silence_phones: Sequence[str], silence_phones: Sequence[str] = ("sil", "spn"),
                     position_dependent: bool = False) -> Dict[str, int]:
    """MFA's phones.txt order: <eps>, silence phones (+ _B _E _I _S variants), then sorted non-silence phones
    (each with _B _E _I _S when position dependent).  See tests/data/dictionaries/expected/phones.txt."""
    table = {"<eps>": 0}
    for p in silence_phones:
        table[p] = len(table)
        if position_dependent:
            for pos in "BEIS":
                table[f"{p}_{pos}"] = len(table)
    for p in sorted(non_silence_phones):
        if position_dependent:
            for pos in "BEIS":
                table[f"{p}_{pos}"] = len(table)
        else:
            table[p] = len(table)
    return table


def read_symbol_table(path) -> Dict[str, int]:
    out = {}
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            parts = line.split()
            if len(parts) == 2:
                out[parts[0]] = int(parts[1])
    return out
