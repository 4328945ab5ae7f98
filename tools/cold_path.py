"""Host-side cost of a NEW batch: graph compilation + packing (+ phase times of the packer with MFA_PACK_TRACE=1) on a synthetic corpus.
Usage: python tools/cold_path.py [hours] [threads]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MFA_PACK_TRACE", "1")
import numpy as np
from mfa_b200 import engine as E, synth as SY
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
nt = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
corpus = SY.make_corpus(hours * 3600.0, seed=1234, n_phones=40, n_words=2000)
rng = np.random.default_rng(1235)
topo = SY.make_topology(corpus.phone_table)
tree, n_pdfs = SY.make_tree(rng, topo, True, 4000)
tm = SY.make_transition_model(topo, tree, n_pdfs)
gc = E.GraphCompiler(tm, tree, corpus.lexicon)
for rep in range(3):
    t0 = time.perf_counter(); b = gc.compile(corpus.transcripts, n_threads=nt); t1 = time.perf_counter()
    g = E.Graphs(b, tm, 1.0, 0.1); t2 = time.perf_counter()
    print(f"{corpus.n_utts} utterances, {nt} threads: compile {1e3 * (t1 - t0):.1f} ms, pack {1e3 * (t2 - t1):.1f} ms, sizes {b.sizes()}")
    g.close(); b.close()
