import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mfa_b200 import engine as E, scenario as SC
eng = E.Engine(0); dev = torch.device("cuda", 0)
sc = SC.build(eng, float(os.environ.get("HOURS", "2")) * 3600.0, seed=1234, target_pdfs=4000, gauss_per_pdf=10, synth_device=dev)
c = sc.corpus; mo = E.mfcc_opts()
d_pcm = torch.from_numpy(c.pcm).to(dev)
res = E.align_pcm(eng, sc.model, sc.graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda)
eng.sync()
fo = sc.frame_off; T = int(fo[-1])
raw, _ = eng.mfcc(d_pcm, c.sample_off, mo)
stats = eng.cmvn_stats(raw, fo, c.utt2spk, c.n_spk); eng.sync()
feats = eng.features(raw, fo, sc.feat_mode, lda=sc.lda, cmvn_stats=stats.cpu().numpy(), utt2spk=c.utt2spk, n_spk=c.n_spk)
ali = res.ali[:T].contiguous()
eng.sync()
print("sum per_frame / T", float(res.per_frame[:T].double().sum()) / T, "nm max", int(np.diff(sc.am.offsets).max()))
for impl in ("segmented", "atomic", "segmented"):
    eng.set_option("acc_impl", 1 if impl == "atomic" else 0)
    sc.model.acc_zero(); sc.model.acc_stats(feats, ali); a = sc.model.acc_read()
    print(impl, a["like"] / a["frames"], a["frames"], a["occ"].sum())
