"""Host-side logic of the hot path's callers (no GPU): job partition, file naming, M-step, CTM extraction, and the N>1
reduction of accumulator statistics over torch.distributed (gloo, world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from helpers import load_model
from mfa_b200 import gmm_update as GU, kaldi_io as K, mfa_functions as MF
from oracle import mstep_oracle as MO


def test_job_assignment_and_paths(tmp_path):
    rng = np.random.default_rng(0)
    utts = []
    for spk in range(9):
        for k in range(int(rng.integers(1, 30))):
            utts.append(MF.Utterance(len(utts), spk, f"/x/{spk}_{k}.wav"))
    jobs = MF.assign_jobs(utts, 3, tmp_path)
    assert sum(len(j.utterances) for j in jobs) == len(utts)
    owner = {}
    for j in jobs:
        for u in j.utterances:
            assert owner.setdefault(u.speaker_id, j.id) == j.id      # a speaker never spans jobs
    loads = [len(j.utterances) for j in jobs]
    assert max(loads) - min(loads) <= max(np.bincount([u.speaker_id for u in utts]))   # LPT bound
    j = jobs[0]
    assert j.construct_path(tmp_path, "ali", "ark", 1).name == f"ali.1.{j.id}.ark"        # db.py:2212-2236
    assert j.construct_path(tmp_path, "feats", "scp").name == f"feats.{j.id}.scp"
    assert utts[5].kaldi_id == f"{utts[5].speaker_id}-{utts[5].id}"


def test_mle_update_matches_closed_form():
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(1)
    G, D = am.NumGauss(), am.dim
    acc = GU.AccumAmDiagGmm(G, D)
    acc.occ = rng.uniform(0, 60, G)
    mu = rng.standard_normal((G, D))
    var = rng.uniform(0.5, 2.0, (G, D))
    acc.mean = acc.occ[:, None] * mu
    acc.var = acc.occ[:, None] * (var + mu ** 2)
    new, impr, count = MO.mle_update(am, acc, mixup=0, min_gaussian_occupancy=10.0)
    assert abs(count - acc.occ.sum()) < 1e-6 and new.NumPdfs() == am.NumPdfs()
    k = 0
    for j in range(am.NumPdfs()):
        a, b = am.offsets[j], am.offsets[j + 1]
        keep = np.where(acc.occ[a:b] > 10.0)[0]
        if len(keep) == 0:
            keep = np.array([np.argmax(acc.occ[a:b])])
        na, nb = new.offsets[j], new.offsets[j + 1]
        assert nb - na == len(keep)
        if (acc.occ[a:b][keep] > 10.0).all():
            assert np.allclose(new.means()[na:nb], mu[a:b][keep], atol=1e-4)
            assert np.allclose(new.variances()[na:nb], var[a:b][keep], rtol=1e-4)
            w = acc.occ[a:b][keep] / acc.occ[a:b][keep].sum()
            # weights are occ/occ_sum over ALL components, renormalised over the survivors
            assert np.allclose(new.weights[na:nb], w, atol=1e-5)
        assert abs(new.weights[na:nb].sum() - 1.0) < 1e-5
    # gconsts of the written model are the closed form
    assert np.allclose(new.gconsts, new.compute_gconsts())
    # mix-up: total grows to the target, per-pdf targets follow occupancy^power
    state_occs = np.asarray([acc.occ[am.offsets[j]:am.offsets[j + 1]].sum() for j in range(am.NumPdfs())])
    t = MO.get_split_targets(state_occs, 600, 0.25, 20.0)
    assert t.sum() <= 600 and t.min() >= 1
    target = new.NumGauss() + 40
    up, _, _ = MO.mle_update(am, acc, mixup=target)
    tg = MO.get_split_targets(state_occs, target, 0.25, 20.0)
    # SplitByCount only ever splits: every pdf ends with max(its current size, its target)
    for j in range(up.NumPdfs()):
        assert up.offsets[j + 1] - up.offsets[j] == max(new.offsets[j + 1] - new.offsets[j], tg[j])
    assert up.NumGauss() > new.NumGauss()
    for j in range(up.NumPdfs()):
        assert abs(up.weights[up.offsets[j]:up.offsets[j + 1]].sum() - 1.0) < 1e-5


def test_transition_update_and_ctm():
    tm, am, _ = load_model("mono")
    stats = tm.InitStats()
    rng = np.random.default_rng(2)
    stats[1:] = rng.integers(0, 50, tm.num_tids)
    old = tm.log_probs.copy()
    impr, cnt = tm.mle_update(stats)
    assert cnt > 0 and np.isfinite(tm.log_probs[1:]).all()
    for ts in range(1, tm.tuples.shape[0] + 1):
        a, b = tm.state2id[ts], tm.state2id[ts + 1]
        assert abs(np.exp(tm.log_probs[a:b].astype(np.float64)).sum() - 1.0) < 1e-4
    # CTM: phone = [forward tids..., final-transition tid, trailing self-loops]
    from mfa_b200.kalpy_compat import Alignment
    tm2, _, _ = load_model("mono")
    ph = 20
    states = tm2.topo.states_for(ph)
    seq = []
    for hs in range(3):
        ts = [i + 1 for i, r in enumerate(tm2.tuples) if r[0] == ph and r[1] == hs][0]
        fwd = [tm2.state2id[ts] + k for k, (d, _) in enumerate(states[hs].transitions) if d != hs][0]
        seq += [fwd] + [tm2.self_loop_tid[ts]] * (hs + 1)
    ctm = Alignment("u", seq + seq, [], 0.0, np.zeros(2 * len(seq))).generate_ctm(tm2, {ph: "x"}, 0.01)
    assert [(c.begin, c.end, c.label) for c in ctm] == [(0.0, 0.09, "x"), (0.09, 0.18, "x")]


def test_accumulator_allreduce_two_ranks_gloo(tmp_path):
    """N>1 path on CPU: two processes hold different accumulator blocks; after all_reduce(SUM) over gloo both hold the total,
    and the M-step run on either rank gives the identical model."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {repr(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))})
        sys.path.insert(0, {repr(os.path.dirname(os.path.abspath(__file__)))})
        import numpy as np, torch, torch.distributed as dist
        from helpers import load_model
        from mfa_b200 import gmm_update as GU
        dist.init_process_group("gloo")
        r = dist.get_rank()
        tm, am, _ = load_model("g2p")
        G, D = am.NumGauss(), am.dim
        rng = np.random.default_rng(10 + r)
        flat = rng.uniform(0, 5, G + 2 * G * D + tm.num_tids + 1 + 2)
        np.save({repr(str(tmp_path))} + f"/in{{r}}.npy", flat)
        t = torch.from_numpy(flat.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        np.save({repr(str(tmp_path))} + f"/out{{r}}.npy", t.numpy())
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    i0, i1 = np.load(tmp_path / "in0.npy"), np.load(tmp_path / "in1.npy")
    o0, o1 = np.load(tmp_path / "out0.npy"), np.load(tmp_path / "out1.npy")
    assert np.array_equal(o0, o1) and np.allclose(o0, i0 + i1, rtol=1e-15)


def test_align_function_first_pass_then_second_pass_keeps_first_pass_archives(tmp_path, monkeypatch):
    """AlignFunction with final.alimdl writes *_first_pass archives and links ali/words/likelihoods to them; the pass with final.mdl
    must REPLACE those links by regular files and leave the first-pass archives untouched (the reference unlinks the three targets
    before every export, alignment/multiprocessing.py:836-839).  The aligner is stubbed: this is file-flow logic, no GPU."""
    from pathlib import Path
    from mfa_b200 import kalpy_compat as KC

    class FakeAligner:
        def __init__(self, model_path, **kw):
            self.tag = Path(str(model_path)).name.encode()

        def boost_silence(self, *a):
            pass

        def export_alignments(self, ali, graphs, feats, word_file_name=None, likelihood_file_name=None, callback=None):
            for p in (ali, word_file_name, likelihood_file_name):
                with open(p, "wb") as f:    # follows symlinks, like ArkWriter
                    f.write(self.tag)

    class FakeArchive:
        def __init__(self, *a, **k):
            pass

        def close(self):
            pass

    monkeypatch.setattr(KC, "GmmAligner", FakeAligner)
    monkeypatch.setattr(KC, "FstArchive", FakeArchive)
    job = MF.Job(1, [], tmp_path)
    monkeypatch.setattr(MF.Job, "construct_feature_archive", lambda self, wd, did=None, **k: FakeArchive())
    for model in ("final.alimdl", "final.mdl"):
        args = MF.AlignArguments(1, job, None, tmp_path, tmp_path / model, {"beam": 10}, False, False, ())
        MF.AlignFunction(args)._run()
        if model == "final.alimdl":
            for name in ("ali", "words", "likelihoods"):
                link = tmp_path / f"{name}.1.1.ark"
                assert link.is_symlink() and link.read_bytes() == b"final.alimdl"
    for name in ("ali", "words", "likelihoods"):
        out, first = tmp_path / f"{name}.1.1.ark", tmp_path / f"{name}_first_pass.1.1.ark"
        assert not out.is_symlink() and out.read_bytes() == b"final.mdl"
        assert first.read_bytes() == b"final.alimdl"         # the first pass survives the second


def test_transition_update_floor_is_the_last_step():
    """TransitionModel::MleUpdate renormalises, then floors, three times: a transition seen far less often than floor ends exactly at
    the floor (0.01), not below it."""
    tm, _, _ = load_model("mono")
    stats = tm.InitStats()
    ts = next(t for t in range(1, tm.tuples.shape[0] + 1) if tm.state2id[t + 1] - tm.state2id[t] == 2)
    a = int(tm.state2id[ts])
    stats[a], stats[a + 1] = 100000.0, 3.0
    tm.mle_update(stats)
    p = np.exp(tm.log_probs[a:a + 2].astype(np.float64))
    assert abs(p[1] - 0.01) < 1e-7 and p[0] < 1.0


def test_mle_update_keeps_the_last_gaussian_of_a_starved_pdf():
    tm, am, _ = load_model("g2p")
    acc = GU.AccumAmDiagGmm.init(am)
    pdf = int(np.argmax(np.diff(am.offsets)))
    a, b = int(am.offsets[pdf]), int(am.offsets[pdf + 1])
    assert b - a >= 2
    rng = np.random.default_rng(0)
    acc.occ[:] = 50.0
    acc.mean[:] = 50.0 * am.means()
    acc.var[:] = 50.0 * (am.variances() + am.means() ** 2)
    acc.occ[a:b] = np.linspace(5.0, 1.0, b - a)     # every component under min_gaussian_occupancy; the FIRST is the heaviest
    new, _, _ = MO.mle_update(am, acc)
    assert new.offsets[pdf + 1] - new.offsets[pdf] == 1
    k = int(new.offsets[pdf])
    assert np.allclose(new.means()[k], am.means()[b - 1], rtol=1e-5)    # Kaldi keeps the LAST index, un-updated


def test_bench_reference_arm_contract_without_cuda_library():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm): one JSON line with the contract's keys, the oracle
    as the thing measured, and libmfa_b200.so never mapped into the process."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--hours", "0.01"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0 and d["cuda_library_loaded"] is False
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["oracle_build"] in ("avx2", "generic")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["vs_baseline"] is None
