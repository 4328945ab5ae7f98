"""-m gpu, needs >= 2 GPUs (skips otherwise): SURVEY.md section 8e on hardware.  Two ranks (torchrun, NCCL) each align their own shard of one
synthetic corpus and accumulate K4 statistics; the NCCL all-reduce of the f64 accumulator blocks must equal the statistics a single
rank gets on the whole corpus (transition counts and frames exactly, sums to f64 rounding)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from helpers import build_synth_scenario
    from mfa_b200 import engine as E
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    sc = build_synth_scenario(seconds=40.0, seed=41, triphone=True, n_phones=8, n_words=30, gauss_per_pdf=2, n_spk=4)   # same on every rank
    c, tm, am = sc["corpus"], sc["tm"], sc["am"]
    eng = E.Engine(rank)
    dm = E.DeviceModel(eng, tm, am)
    def stats(utts):
        batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile([c.transcripts[u] for u in utts])
        graphs = E.Graphs(batch, tm, 1.0, 0.1)
        pcm = np.concatenate([c.pcm[c.sample_off[u]:c.sample_off[u + 1]] for u in utts])
        so = np.zeros(len(utts) + 1, np.int64); so[1:] = np.cumsum([c.sample_off[u + 1] - c.sample_off[u] for u in utts])
        spk = {{s: i for i, s in enumerate(sorted({{int(c.utt2spk[u]) for u in utts}}))}}
        u2s = np.asarray([spk[int(c.utt2spk[u])] for u in utts], np.int32)
        res = E.align_pcm(eng, dm, graphs, pcm, so, u2s, len(spk), E.mfcc_opts(), "deltas")
        raw, fo = eng.mfcc(pcm, so, E.mfcc_opts())
        cm = eng.cmvn_stats(raw, fo, u2s, len(spk))
        feats = eng.features(raw, fo, "deltas", cmvn_stats=cm, utt2spk=u2s, n_spk=len(spk))
        dm.acc_zero(); dm.acc_stats(feats, res.ali[: int(fo[-1])]); eng.sync()
    mine = [u for u in range(c.n_utts) if int(c.utt2spk[u]) % world == rank]   # a speaker never spans ranks
    stats(mine)
    t = dm.acc_tensor()
    dist.all_reduce(t)
    torch.cuda.synchronize()
    summed = t.cpu().numpy().copy()
    if rank == 0:
        by_spk = sorted(range(c.n_utts), key=lambda u: (int(c.utt2spk[u]) % world, u))
        stats(by_spk)
        whole = dm.acc_tensor().cpu().numpy()
        a, b = dm.split_accs(summed), dm.split_accs(whole)
        assert a["frames"] == b["frames"] > 0 and np.array_equal(a["trans"], b["trans"])
        assert abs(a["like"] - b["like"]) <= 1e-9 * abs(b["like"])
        assert np.allclose(a["occ"], b["occ"], rtol=1e-9, atol=1e-9) and np.allclose(a["mean"], b["mean"], rtol=1e-9, atol=1e-7)
        print("MULTI_OK", a["frames"])
    dist.barrier(); dist.destroy_process_group()
""")


def test_nccl_allreduce_of_accumulators_equals_single_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
