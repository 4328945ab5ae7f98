"""What the first alignment call after a model / graphs change pays on top of the steady-state step (K2 operand images, per-utterance
tile plan, graph upload): 10 h workload."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mfa_b200 import engine as E, scenario as SC
eng = E.Engine(0); dev = torch.device("cuda", 0)
sc = SC.build(eng, float(os.environ.get("HOURS", "10")) * 3600.0, seed=1234, target_pdfs=4000, gauss_per_pdf=10, synth_device=dev)
c = sc.corpus; mo = E.mfcc_opts()
d_pcm = torch.from_numpy(c.pcm).to(dev)
def step(model, graphs):
    eng.sync(); t0 = time.perf_counter()
    E.align_pcm(eng, model, graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda)
    eng.sync(); return 1e3 * (time.perf_counter() - t0)
print("first call (fresh model + graphs): %.1f ms" % step(sc.model, sc.graphs))
print("steady: %.1f %.1f ms" % (step(sc.model, sc.graphs), step(sc.model, sc.graphs)))
t0 = time.perf_counter(); m2 = E.DeviceModel(eng, sc.tm, sc.am); eng.sync(); print("DeviceModel(): %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print("new model, same graphs: %.1f ms, then %.1f" % (step(m2, sc.graphs), step(m2, sc.graphs)))
t0 = time.perf_counter(); g2 = E.Graphs(sc.batch, sc.tm, 1.0, 0.1); print("Graphs() pack: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print("same model, new graphs: %.1f ms, then %.1f" % (step(m2, g2), step(m2, g2)))
