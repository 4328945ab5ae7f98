"""Host wall-clock latency of repeated device M-steps (many small synchronous driver calls): distribution over N iterations.
Usage: python tools/mstep_latency.py [hours] [iterations]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mfa_b200 import engine as E, scenario as SC
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
eng = E.Engine(0)
sc = SC.build(eng, hours * 3600.0, seed=1234, n_threads=os.cpu_count(), synth_device=dev)
c = sc.corpus; mo = E.mfcc_opts(); fo = sc.frame_off; T = int(fo[-1])
d_pcm = torch.from_numpy(c.pcm).to(dev)
model = E.DeviceModel(eng, sc.tm, sc.am); model.set_transitions(sc.tm)
graphs = E.Graphs(sc.batch, sc.tm, 1.0, 0.1)
raw, _ = eng.mfcc(d_pcm, c.sample_off, mo); stats = eng.cmvn_stats(raw, fo, c.utt2spk, c.n_spk); eng.sync()
feats = eng.features(raw, fo, sc.feat_mode, lda=sc.lda, cmvn_stats=stats.cpu().numpy(), utt2spk=c.utt2spk, n_spk=c.n_spk); eng.sync()
g0 = sc.am.NumGauss()
lat = {"align": [], "acc": [], "mstep": [], "refold": []}
for it in range(iters):
    t0 = time.perf_counter(); res = E.align_pcm(eng, model, graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda); eng.sync(); lat["align"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); model.acc_zero(); model.acc_stats(feats, res.ali[:T]); eng.sync(); lat["acc"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); model.mle_update(mixup=g0, update_transitions=True, seed=1234 + it); lat["mstep"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); graphs.set_transitions(eng, model, 1.0, 0.1); eng.sync(); lat["refold"].append(time.perf_counter() - t0)
print("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
for k, v in lat.items():
    v = 1e3 * np.asarray(v)
    print(f"{k:7s} median {np.median(v):8.2f} ms  p90 {np.percentile(v, 90):8.2f}  max {v.max():8.2f}  (top 3: {np.sort(v)[-3:].round(1).tolist()})")
