"""Timeline of the end-to-end arm with several jobs (engines) on one GPU: per call, wall-clock start / end on the host and the
per-stage CUDA-event intervals, to see what overlaps between jobs.  Usage: python tools/e2e_timeline.py [jobs] [steps_per_job] [hours]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as G
G.build()
from mfa_b200 import engine as E, scenario as SC

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
hours = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
eng = E.Engine(0)
sc = SC.build(eng, hours * 3600.0, seed=1234, n_threads=os.cpu_count(), synth_device=dev)
c = sc.corpus
mo = E.mfcc_opts()
n_frames = int(sc.frame_off[-1])
wo_total = int(np.cumsum(sc.graphs.max_words())[-1])
h_pcm = torch.from_numpy(c.pcm).pin_memory()
jobs = []
for j in range(n_jobs):
    e2 = eng if j == 0 else E.Engine(0)
    outs = E._alloc_outputs(n_frames, wo_total, c.n_utts, dev)
    ho = tuple(torch.zeros(x.shape, dtype=x.dtype).pin_memory().numpy() for x in outs)
    jobs.append((e2, sc.model if j == 0 else E.DeviceModel(e2, sc.tm, sc.am), sc.graphs if j == 0 else E.Graphs(sc.batch, sc.tm, 1.0, 0.1), ho))
pcm_np = h_pcm.numpy()

def step(j):
    en, mdl, gr, ho = jobs[j]
    return E.align_pcm(en, mdl, gr, pcm_np, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, outputs=ho, workspace_bytes=24 << 30)

for j in range(n_jobs):
    step(j); step(j)
log = []
go = threading.Barrier(n_jobs + 1)

def work(j):
    go.wait()
    for i in range(steps):
        t0 = time.perf_counter()
        step(j)
        t1 = time.perf_counter()
        log.append((j, i, t0, t1, jobs[j][0].stage_timing()))

th = [threading.Thread(target=work, args=(j,)) for j in range(n_jobs)]
for t in th:
    t.start()
go.wait()
T0 = time.perf_counter()
for t in th:
    t.join()
T1 = time.perf_counter()
print(f"{n_jobs} jobs x {steps} steps: {1e3 * (T1 - T0) / (n_jobs * steps):.2f} ms per step, {c.seconds * n_jobs * steps / (T1 - T0):.0f} x RT")
for j, i, t0, t1, st in sorted(log, key=lambda x: x[2]):
    print(f"job {j} step {i}: host call {1e3 * (t0 - T0):8.2f} -> {1e3 * (t1 - T0):8.2f} ms ({1e3 * (t1 - t0):6.2f})  stages " +
          " ".join(f"{k}={v:.1f}" for k, v in st.items()))
