"""Writes tests/golden/kalpy_imports.json: every ``from kalpy... import ...`` of the reference's hot-path modules,
parsed with ``ast`` from /root/reference (read-only).  tests/test_kalpy_shim.py resolves each of them against the shim package
(montreal-forced-aligner_b200/shim/kalpy) -- and re-derives this file when the reference tree is present, so it cannot go stale.

  python tools/extract_kalpy_imports.py
"""
import ast
import json
import os

REF = "/root/reference/montreal_forced_aligner"
FILES = ["alignment/multiprocessing.py", "alignment/base.py", "alignment/mixins.py", "corpus/features.py", "corpus/acoustic_corpus.py",
         "online/alignment.py", "acoustic_modeling/monophone.py", "acoustic_modeling/trainer.py", "acoustic_modeling/sat.py",
         "acoustic_modeling/base.py", "command_line/align_one.py", "db.py", "models.py"]
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kalpy_imports.json")


def extract(ref=REF):
    out = {}
    for rel in FILES:
        path = os.path.join(ref, rel)
        if not os.path.exists(path):
            continue
        tree = ast.parse(open(path, encoding="utf8").read())
        rows = []
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and (node.module == "kalpy" or node.module.startswith("kalpy.")):
                rows.append([node.module, sorted(a.name for a in node.names), node.lineno])
        if rows:
            out[rel] = sorted(rows, key=lambda r: r[2])
    return out


if __name__ == "__main__":
    d = extract()
    json.dump(d, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, sum(len(v) for v in d.values()), "import statements from", len(d), "files")
