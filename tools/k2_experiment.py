"""Where does K2's time go?  Builds timing variants of the library (-DMFA_TC_EXP=n, WRONG results, into /tmp) and measures the K2 stage of
the 10 h workload with each, in a subprocess per variant (MFA_B200_LIB selects the library).  Run on the GPU box:

    python tools/k2_experiment.py [--hours 10] [--variants 0,1,2,3,4,5,6]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from mfa_b200 import engine as E, scenario as SC
eng = E.Engine(0)
dev = torch.device("cuda", 0)
sc = SC.build(eng, %f * 3600.0, seed=1234, target_pdfs=4000, gauss_per_pdf=10, n_threads=os.cpu_count() or 8, synth_device=dev)
c = sc.corpus
d_pcm = torch.from_numpy(c.pcm).to(dev)
mo = E.mfcc_opts()
outs = E._alloc_outputs(int(sc.frame_off[-1]), int(np.cumsum(sc.graphs.max_words())[-1]), c.n_utts, dev)
ts = []
for i in range(8):
    E.align_pcm(eng, sc.model, sc.graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, workspace_bytes=100 << 30, outputs=outs)
    eng.sync()
    if i >= 3:
        ts.append(eng.stage_timing())
print("RESULT", json.dumps({k: float(np.median([t[k] for t in ts])) for k in ts[0]}))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=10.0)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6")
    args = ap.parse_args()
    import mfa_b200.build as B
    names = {0: "baseline", 1: "no stores", 2: "epilogue = TMEM loads + release only", 3: "epilogue releases unread", 4: "3 MMAs per tile (of 15)",
             5: "B tiles fetched once per item", 6: "2 + 5", 7: "correct results + cycle counters"}
    out = {}
    for v in [int(x) for x in args.variants.split(",")]:
        lib = f"/tmp/libmfa_exp{v}.so"
        B.build(force=True, extra_flags=[f"-DMFA_TC_EXP={v}"], out=lib, objdir=f"/tmp/mfa_exp_obj{v}")
        env = dict(os.environ, MFA_B200_LIB=lib)
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, args.hours)], env=env, capture_output=True, text=True, timeout=900)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
        out[names.get(v, str(v))] = json.loads(line[0][7:]) if line else {"failed": r.stderr[-400:]}
        dbg = [ln for ln in r.stderr.splitlines() if ln.startswith("[tc-exp7]")]
        if dbg:
            out["cycle counters (last launch)"] = dbg[-1]
            print(dbg[-1])
        print(v, names.get(v), out[names.get(v, str(v))], flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k2_experiment.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
