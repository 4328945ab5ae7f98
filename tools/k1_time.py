import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mfa_b200 import engine as E
eng = E.Engine(0)
rng = np.random.default_rng(0)
n_utts = 600
lens = rng.integers(16000 * 5, 16000 * 20, n_utts)
off = np.zeros(n_utts + 1, np.int64); off[1:] = np.cumsum(lens)
pcm = torch.from_numpy((rng.standard_normal(int(off[-1])) * 3000).astype(np.int16)).cuda()
mo = E.mfcc_opts()
for gen in (0, 1):
    with eng.options(mfcc_generic=gen):
        for _ in range(3): out, fo = eng.mfcc(pcm, off, mo)
        eng.sync(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): out, fo = eng.mfcc(pcm, off, mo)
        eng.sync(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("generic" if gen else "fast", f"{dt*1e3:.3f} ms for {int(fo[-1])} frames = {dt*1e3/(int(fo[-1])/3.6e6):.2f} ms per 10 h", float(out.float().abs().sum()))
