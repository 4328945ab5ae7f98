"""ali -> CTM -> word / phone intervals -> TextGrid (the user-visible output of the path).

Mirrors, on the host, what the reference does after AlignFunction:
  * AlignmentExtractionFunction._run (montreal_forced_aligner/alignment/multiprocessing.py:1614-1862): per utterance
    ``alignment.generate_ctm`` -> ``lexicon_compiler.phones_to_pronunciations(words, intervals, text=...)`` ->
    ``ctm.update_utterance_boundaries(begin, end)`` -> ``fix_unk_words`` (helper.py:772);
  * export_textgrid (textgrid.py:463-572): words / phones tiers per speaker, the last interval snapped to the file duration when
    it ends within two frames of it, overlaps clipped, formats long_textgrid / short_textgrid / json / csv;
  * construct_textgrid_output's ``cleanup_textgrids`` (textgrid.py:279-330): silence words / phones are left blank.
praatio is not a dependency here: the TextGrid writer fills the blanks itself (praatio ``includeBlankSpaces=True``) and a small reader
for both TextGrid flavours is provided for the round-trip tests.
"""
from __future__ import annotations

import csv
import json
import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

from .kalpy_compat import Alignment, CtmInterval
from .lexicon import Lexicon

_POS = re.compile(r"_[BEIS]$")


@dataclass
class WordCtmInterval:
    """kalpy.gmm.data.WordCtmInterval: a word with the phone intervals of the pronunciation that was aligned."""
    label: str
    word_id: int
    phones: List[CtmInterval] = field(default_factory=list)
    pronunciation: str = ""

    @property
    def begin(self) -> float:
        return self.phones[0].begin

    @property
    def end(self) -> float:
        return self.phones[-1].end


@dataclass
class HierarchicalCtm:
    """kalpy.gmm.data.HierarchicalCtm: word intervals owning their phone intervals."""
    word_intervals: List[WordCtmInterval]
    text: Optional[str] = None
    likelihood: Optional[float] = None

    @property
    def phone_intervals(self) -> List[CtmInterval]:
        return [p for w in self.word_intervals for p in w.phones]

    def export_textgrid(self, output_path, file_duration: float, output_format: str = "long_textgrid", frame_shift: float = 0.01,
                        speaker: str = "speaker", silence_word: str = "<eps>", cleanup_textgrids: bool = True):
        """kalpy HierarchicalCtm.export_textgrid as align_one calls it (command_line/align_one.py:194-196): words / phones tiers."""
        words = [CtmInterval(w.begin, w.end, w.label) for w in self.word_intervals if not (cleanup_textgrids and w.label == silence_word)]
        phones = [CtmInterval(p.begin, p.end, p.label, p.confidence) for w in self.word_intervals
                  if not (cleanup_textgrids and w.label == silence_word) for p in w.phones]
        return export_textgrid({speaker: {"words": words, "phones": phones}}, output_path, file_duration, frame_shift, output_format)

    def update_utterance_boundaries(self, begin: Optional[float], end: Optional[float] = None):
        """Shift by the utterance's begin inside its file; the last phone is clipped to the utterance end."""
        b = float(begin or 0.0)
        for w in self.word_intervals:
            for p in w.phones:
                p.begin = round(p.begin + b, 6)
                p.end = round(p.end + b, 6)
        if end is not None and self.word_intervals:
            last = self.word_intervals[-1].phones[-1]
            if last.end > end:
                last.end = round(float(end), 6)


def _strip(label: str, position_dependent: bool) -> str:
    return _POS.sub("", label) if position_dependent else label


def phones_to_pronunciations(lexicon: Lexicon, words: Sequence[int], intervals: Sequence[CtmInterval], transcription: bool = False,
                             text: Optional[str] = None) -> HierarchicalCtm:
    """LexiconCompiler.phones_to_pronunciations as called at alignment/multiprocessing.py:1739-1744: group the phone intervals of an
    alignment into the words of its olabel sequence.  Optional-silence intervals between words become ``<eps>`` word intervals; each
    word takes the pronunciation (of that word) that matches the phones at its position -- a depth-first search over the word's
    pronunciations resolves the rare case where one pronunciation is a prefix of another."""
    pt = lexicon.phone_table
    sil_id = pt[lexicon.silence_phone]
    ids = [pt[iv.label] if isinstance(iv.label, str) else int(iv.label) for iv in intervals]
    n, nw = len(ids), len(words)
    prons = {w: sorted(lexicon.word_prons_as_phone_ids(int(w)), key=len, reverse=True) for w in set(int(w) for w in words)}
    dead = set()

    def search(i: int, k: int):
        """-> list of (start, length) per word from position i / word k, or None."""
        if (i, k) in dead:
            return None
        j = i
        while j < n and ids[j] == sil_id:
            j += 1
        if k == nw:
            return [] if j == n else None
        for pr in prons[int(words[k])]:
            if ids[j:j + len(pr)] == pr:
                rest = search(j + len(pr), k + 1)
                if rest is not None:
                    return [(j, len(pr))] + rest
        # a word whose own pronunciation is the silence phone (not produced by MFA's lexicons, kept for completeness)
        dead.add((i, k))
        return None

    import sys
    if nw + 50 > sys.getrecursionlimit():
        sys.setrecursionlimit(nw + 200)
    spans = search(0, 0)
    if spans is None:
        raise ValueError("phones_to_pronunciations: the phone sequence is not a concatenation of pronunciations of the aligned words")
    pd = lexicon.position_dependent_phones
    out: List[WordCtmInterval] = []
    pos = 0
    for k, (s, ln) in enumerate(spans + [(n, 0)]):
        for j in range(pos, s):   # silence run before the word (or trailing)
            iv = intervals[j]
            out.append(WordCtmInterval(lexicon.silence_word, 0, [CtmInterval(iv.begin, iv.end, _strip(str(iv.label), pd), iv.confidence)],
                                       lexicon.silence_phone))
        if ln:
            ph = [CtmInterval(iv.begin, iv.end, _strip(str(iv.label), pd), iv.confidence) for iv in intervals[s:s + ln]]
            out.append(WordCtmInterval(lexicon.id2word[int(words[k])], int(words[k]), ph, " ".join(p.label for p in ph)))
        pos = s + ln
    return HierarchicalCtm(out, text)


def fix_unk_words(ref_words: Sequence[str], word_intervals: List[WordCtmInterval], lexicon: Lexicon) -> List[WordCtmInterval]:
    """helper.py:772-851: the aligned ``<unk>`` intervals get the original (out-of-vocabulary) word of the transcript back.  The
    reference aligns the two sequences with an edit-distance alignment; word order is fixed in forced alignment, so a positional
    match over the non-silence intervals is the same thing whenever the counts agree (and nothing is touched when they do not)."""
    real = [w for w in word_intervals if w.label != lexicon.silence_word]
    if len(real) == len(ref_words):
        for w, r in zip(real, ref_words):
            if w.label == lexicon.oov_word:
                w.label = r
    return word_intervals


def alignment_to_ctm(alignment: Alignment, transition_model, lexicon: Lexicon, frame_shift: float = 0.01, begin: Optional[float] = None,
                     end: Optional[float] = None, text: Optional[str] = None) -> HierarchicalCtm:
    """The per-utterance body of AlignmentExtractionFunction._run (alignment/multiprocessing.py:1734-1751)."""
    id2ph = {v: k for k, v in lexicon.phone_table.items()}
    intervals = alignment.generate_ctm(transition_model, id2ph, frame_shift)
    ctm = phones_to_pronunciations(lexicon, alignment.words, intervals, transcription=False, text=text)
    ctm.likelihood = alignment.likelihood
    ctm.update_utterance_boundaries(begin, end)
    if text is not None:
        ctm.word_intervals = fix_unk_words(text.split(), ctm.word_intervals, lexicon)
    return ctm


# ------------------------------------------------------------------------------------------------ TextGrid writer / reader
def _tier_entries(intervals: Sequence[CtmInterval], duration: float, frame_shift: float) -> List[Tuple[float, float, str]]:
    """textgrid.py:540-556: sort, snap the last interval to the duration, clip overlaps; then fill blanks (includeBlankSpaces)."""
    ivs = sorted(intervals, key=lambda x: x.begin)
    ent: List[Tuple[float, float, str]] = []
    for i, a in enumerate(ivs):
        b, e = round(a.begin, 6), round(a.end, 6)
        if i == len(ivs) - 1 and duration - e < frame_shift * 2:
            e = duration
        if ent and ent[-1][1] > b:
            b = ent[-1][1]
        if e > duration:
            e = duration
        if e > b:
            ent.append((b, e, str(a.label)))
    out: List[Tuple[float, float, str]] = []
    t = 0.0
    for b, e, lab in ent:
        if b > t:
            out.append((t, b, ""))
        out.append((b, e, lab))
        t = e
    if t < duration:
        out.append((t, duration, ""))
    return out


def _q(s: str) -> str:
    return '"' + s.replace('"', '""') + '"'


def write_textgrid(path, tiers: Dict[str, List[Tuple[float, float, str]]], duration: float, short: bool = False):
    L = []
    if short:
        L += ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "0", f"{duration:.10g}", "<exists>", str(len(tiers))]
        for name, ent in tiers.items():
            L += ['"IntervalTier"', _q(name), "0", f"{duration:.10g}", str(len(ent))]
            for b, e, lab in ent:
                L += [f"{b:.10g}", f"{e:.10g}", _q(lab)]
    else:
        L += ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "xmin = 0 ", f"xmax = {duration:.10g} ", "tiers? <exists> ", f"size = {len(tiers)} ", "item []: "]
        for ti, (name, ent) in enumerate(tiers.items(), 1):
            L += [f"    item [{ti}]:", '        class = "IntervalTier" ', f"        name = {_q(name)} ", "        xmin = 0 ", f"        xmax = {duration:.10g} ",
                  f"        intervals: size = {len(ent)} "]
            for k, (b, e, lab) in enumerate(ent, 1):
                L += [f"        intervals [{k}]:", f"            xmin = {b:.10g} ", f"            xmax = {e:.10g} ", f"            text = {_q(lab)} "]
    with open(path, "w", encoding="utf8") as f:
        f.write("\n".join(L) + "\n")


def read_textgrid(path) -> Dict[str, List[Tuple[float, float, str]]]:
    """Interval tiers of a long or short TextGrid (enough for MFA's own outputs and the reference's test fixtures)."""
    raw = Path(path).read_bytes()
    txt = raw.decode("utf-16") if raw[:2] in (b"\xff\xfe", b"\xfe\xff") else raw.decode("utf8")
    tok = re.findall(r'"(?:[^"]|"")*"|-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?|<exists>', txt)
    # both flavours reduce to the same token stream once the keys of the long format are dropped
    tok = tok[2:]                                    # "ooTextFile", "TextGrid"
    i = 0
    nums = lambda s: float(s)
    i += 2                                           # xmin xmax
    assert tok[i] == "<exists>", "not an interval TextGrid"
    n_tiers = int(tok[i + 1]); i += 2
    tiers: Dict[str, List[Tuple[float, float, str]]] = {}
    unq = lambda s: s[1:-1].replace('""', '"')
    for _ in range(n_tiers):
        # the long format numbers its items ("item [1]:"): skip bare integers until the class string
        while not tok[i].startswith('"'):
            i += 1
        cls, name = unq(tok[i]), unq(tok[i + 1]); i += 2
        i += 2                                       # tier xmin xmax
        n = int(tok[i]); i += 1
        ent = []
        for _k in range(n):
            while not (re.match(r"-?\d", tok[i]) and re.match(r"-?\d", tok[i + 1]) and tok[i + 2].startswith('"')):
                i += 1                               # "intervals [k]:" index of the long format
            if cls == "IntervalTier":
                ent.append((nums(tok[i]), nums(tok[i + 1]), unq(tok[i + 2]))); i += 3
            else:
                i += 2
        tiers[name] = ent
    return tiers


def export_textgrid(speaker_data: Dict[str, Dict[str, List[CtmInterval]]], output_path, duration: float, frame_shift: float,
                    output_format: str = "long_textgrid"):
    """textgrid.py:463-572.  speaker_data: speaker -> {"words": [...], "phones": [...]}.  Nothing is written when there is no interval."""
    duration = round(duration, 6)
    has_data = any(len(iv) for data in speaker_data.values() for iv in data.values())
    if not has_data:
        return False
    multi = len(speaker_data) > 1
    if output_format == "csv":
        with open(output_path, "w", encoding="utf8", newline="") as f:
            w = csv.DictWriter(f, fieldnames=["Begin", "End", "Label", "Type", "Speaker"])
            w.writeheader()
            for spk, data in speaker_data.items():
                for kind, ivs in data.items():
                    for a in ivs:
                        e = duration if duration - a.end < frame_shift * 2 else a.end
                        w.writerow({"Begin": a.begin, "End": e, "Label": a.label, "Type": kind, "Speaker": spk})
        return True
    tiers: Dict[str, List[Tuple[float, float, str]]] = {}
    for spk, data in speaker_data.items():
        for kind, ivs in data.items():
            tiers[f"{spk} - {kind}" if multi else kind] = _tier_entries(ivs, duration, frame_shift)
    if output_format == "json":
        js = {"start": 0, "end": duration, "tiers": {k: {"type": "interval", "entries": [[b, e, lab] for b, e, lab in v if lab != ""]}
                                                       for k, v in tiers.items()}}
        with open(output_path, "w", encoding="utf8") as f:
            json.dump(js, f, indent=4, ensure_ascii=False)
        return True
    write_textgrid(output_path, tiers, duration, short=(output_format == "short_textgrid"))
    return True


def ctm_to_speaker_data(ctm: HierarchicalCtm, speaker: str, lexicon: Lexicon, cleanup_textgrids: bool = True):
    """words / phones interval lists of one utterance; silence left blank when cleanup_textgrids (textgrid.py:315-317)."""
    words, phones = [], []
    for w in ctm.word_intervals:
        sil = w.label == lexicon.silence_word
        if sil and cleanup_textgrids:
            continue
        words.append(CtmInterval(w.begin, w.end, w.label))
        phones.extend(CtmInterval(p.begin, p.end, p.label, p.confidence) for p in w.phones)
    return {speaker: {"words": words, "phones": phones}}
