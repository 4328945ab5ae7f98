"""Commits the reference's own dictionary fixtures (tests/data/dictionaries/expected/: the lexicon FST in OpenFst text form, the
Kaldi text topology, phones.txt, words.txt -- the files SURVEY.md section 8(f) N1 names) plus the dictionary they were made from as
tests/golden/dictionary_fixtures.npz, so the lexicon / topology / graph-compiler checks run on the GPU box without /root/reference.

  python tests/golden/make_dictionary_golden.py
"""
import os

import numpy as np

REF = "/root/reference/tests/data/dictionaries"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for key, rel in (("lexicon_text_fst", "expected/lexicon.text.fst"), ("topo", "expected/topo"), ("phones_txt", "expected/phones.txt"),
                     ("words_txt", "expected/words.txt"), ("abstract_dict", "test_abstract.txt")):
        out[key] = np.frombuffer(open(os.path.join(REF, rel), "rb").read(), dtype=np.uint8)
    path = os.path.join(OUT, "dictionary_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
