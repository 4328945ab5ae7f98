#!/usr/bin/env python
"""bench.py -- audio-seconds aligned per second (xRT) for MFA's alignment hot path on B200.

A "step" is one pass of the fused hot path (PCM -> MFCC -> CMVN -> splice+LDA -> all-pdf GMM log-likelihoods ->
beam Viterbi) over the whole per-GPU workload: BASELINE.json configs[1], a triphone LDA-shaped GMM-HMM (~4k pdfs,
~40k Gaussians, D=40) aligning 10 h of synthetic LibriSpeech-shaped 16 kHz audio (per GPU; weak scaling).

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the CPU oracle port on the host cores (reference arm)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before torch creates the CUDA context (see mfa_b200/_lib.py); reported in the line

METRIC = "audio-sec aligned/sec (xRT)"
UNIT = "audio-s/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0), tf_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def bind_to_gpu_numa(index: int):
    """Pin this process (and the pinned host buffers it allocates and fills from now on: first touch) to the CPUs of the NUMA node
    the GPU hangs off, so that the end-to-end arm's H2D traffic does not cross the socket interconnect.  Returns a description."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
        if node < 0:
            # virtualised boxes often hide the sysfs node; NVML may still know the GPU's ideal CPUs
            try:
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByPciBusId(pci.encode())
                words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
                cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & os.sched_getaffinity(0)
                if cpus and len(cpus) < len(os.sched_getaffinity(0)):
                    os.sched_setaffinity(0, cpus)
                    return {"pci": pci, "numa_node": node, "bound": True, "cpus": len(cpus), "via": "nvml"}
            except Exception:
                pass
            return {"pci": pci, "numa_node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"pci": pci, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"pci": pci, "numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as ex:   # containers without sysfs topology, older torch: run unbound
        return {"bound": False, "why": repr(ex)}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, every 50 ms -- eight ranks polling every 5 ms made the
    driver-side NVML lock a shared stall in round 1; falls back to nvidia-smi polling)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mask, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self.sm.append(float(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
                    try:
                        self.mask |= int(self._nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    except Exception:
                        self.mask |= int(self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self._stop.wait(0.05)
                else:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(out[0])); self.max_mhz = float(out[1]); self.mask |= int(out[2].strip(), 16)
                    self._stop.wait(0.05)
            except Exception:
                self._stop.wait(0.05)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        reasons = sorted(n for n, b in self.REASONS.items() if self.mask & b)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_pass(sc, utts, cores: int):
    """The oracle port (oracle/oracle.c: scalar restatement of the Kaldi path MFA calls through kalpy) over `utts`,
    `cores` host threads (ctypes releases the GIL).  Returns (wall seconds, audio seconds, n_ok, results) -- results[i] is the
    oracle's alignment of utts[i] (status, ali, words, like, per_frame), kept for the parity block of the JSON line."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    c = sc.corpus
    g = O.GmmModel.from_am(sc.am)
    tid_cost = -sc.tm.scaled_transition_log_probs(1.0, 0.1)
    fsts = sc._fsts
    csr = {u: O.FstCsr(fsts[u]) for u in utts}
    opts = O.mfcc_opts()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        raw = dict(zip(utts, ex.map(lambda u: O.mfcc(c.pcm[c.sample_off[u]:c.sample_off[u + 1]], opts), utts)))
        spk = sorted({int(c.utt2spk[u]) for u in utts})
        stats = {s: O.cmvn_stats([raw[u] for u in utts if c.utt2spk[u] == s]) for s in spk}

        def one(u):
            x = O.cmvn_apply(raw[u], stats[int(c.utt2spk[u])])
            x = O.transform(O.splice(x, 3, 3), sc.lda) if sc.feat_mode == "lda" else O.add_deltas(x)
            return O.align(csr[u], tid_cost, g, sc.tm.tid2pdf, x, x.shape[0], 0.1, 10.0, 40.0)
        res = list(ex.map(one, utts))
    dt = time.perf_counter() - t0
    secs = float(sum(c.sample_off[u + 1] - c.sample_off[u] for u in utts)) / 16000.0
    return dt, secs, sum(1 for r in res if r["status"] < 2), res


def parity_block(sc, utts, ref, gpu):
    """BASELINE.json's second metric: the GPU step's outputs against the oracle's on the CPU-baseline sample of the very same
    workload (north_star bars: >= 99.9 % identical transition-ids, boundaries within one 10 ms frame, log-likelihood 1e-4 relative).
    `gpu` = host copies (ali, per_frame, words, num_words, total_like, status) of the last timed device-resident step."""
    from mfa_b200 import kalpy_compat as KC
    ali, pf, words, nw, tl, st = gpu
    fo = sc.frame_off
    wo = np.zeros(len(fo), np.int64)
    wo[1:] = np.cumsum(sc.graphs.max_words())
    same = total = status_mis = word_mis = 0
    b_ok = b_tot = 0
    like_err = pf_err = 0.0
    worst = pf_worst = None
    pf_n_bad = 0
    for u, r in zip(utts, ref):
        if int(st[u]) != int(r["status"]):
            status_mis += 1
            continue
        if r["status"] >= 2:
            continue
        a = ali[fo[u]:fo[u + 1]]
        n_same = int((a == r["ali"]).sum())
        same += n_same; total += len(a)
        if worst is None or len(a) - n_same > worst[1]:
            dt = np.nonzero(a != r["ali"])[0][:6]
            worst = (int(u), len(a) - n_same, len(a), int(st[u]), [(int(t), int(a[t]), int(r["ali"][t]), int(sc.tm.tid2phone[a[t]]), int(sc.tm.tid2phone[r["ali"][t]]),
                                                                   int(sc.tm.tid2pdf[a[t]]), int(sc.tm.tid2pdf[r["ali"][t]])) for t in dt])
        if list(words[wo[u]:wo[u] + nw[u]]) != list(r["words"]):
            word_mis += 1
        like_err = max(like_err, abs(float(tl[u]) - r["like"]) / max(1e-30, abs(r["like"])))
        eq = a == r["ali"]
        d = np.abs(pf[fo[u]:fo[u + 1]] - r["per_frame"]) / np.maximum(1.0, np.abs(r["per_frame"]))
        d[~eq] = 0.0
        if d.size and float(d.max()) > pf_err:
            t = int(np.argmax(d))
            pf_err = float(d.max())
            pf_worst = {"utt": int(u), "frame": t, "oracle": float(r["per_frame"][t]), "gpu": float(pf[fo[u] + t]), "tid": int(a[t])}
        pf_n_bad += int((d > 1e-4).sum())
        cg = KC.Alignment(str(u), a, [], float(tl[u])).generate_ctm(sc.tm, None)
        cr = KC.Alignment(str(u), r["ali"], [], r["like"]).generate_ctm(sc.tm, None)
        b_tot += 2 * len(cr)
        if [x.label for x in cg] == [x.label for x in cr]:
            b_ok += sum(int(abs(x.begin - y.begin) <= 0.0101) + int(abs(x.end - y.end) <= 0.0101) for x, y in zip(cg, cr))
    return {"against": "oracle port (oracle/oracle.c), same utterances as cpu_baseline.sample; parity unpinned vs real Kaldi (DESIGN.md 2)",
            "utterances": len(utts), "frames": total, "frame_agreement_pct": 100.0 * same / max(1, total), "status_mismatches": status_mis,
            "word_sequence_mismatches": word_mis, "loglike_rel_err_max": like_err, "per_frame_loglike_rel_err_max": pf_err,
            "per_frame_loglike_worst": pf_worst, "per_frame_loglikes_beyond_1e-4": pf_n_bad,
            "boundary_within_1_frame_pct": 100.0 * b_ok / max(1, b_tot), "phone_boundaries": b_tot,
            "worst_utterance": None if worst is None else {"utt": worst[0], "differing_frames": worst[1], "frames": worst[2], "status": worst[3],
                                                                   "first_differences_t_gpuTid_oracleTid_gpuPhone_oraclePhone_gpuPdf_oraclePdf": worst[4]},
            "retried_utterances_in_sample": int(sum(1 for r in ref if r["status"] == 1))}


def reference_scenario(seconds: float, seed: int, target_pdfs: int, gauss_per_pdf: int, cores: int):
    """The CPU arm's own workload, built WITHOUT this repo's CUDA library (libmfa_b200.so is never loaded by `--impl reference`): the same
    synthetic corpus generator and model recipe as mfa_b200.scenario.build (numpy), features from the oracle, training graphs from the
    oracle's pure-Python restatement of the graph compiler (oracle/graph_oracle.py)."""
    from concurrent.futures import ThreadPoolExecutor
    from types import SimpleNamespace
    from mfa_b200 import synth as SY
    from oracle import graph_oracle as GO, oracle as O
    corpus = SY.make_corpus(seconds, seed=seed, n_phones=40, n_words=2000, device=None)
    rng = np.random.default_rng(seed + 1)
    topo = SY.make_topology(corpus.phone_table)
    tree, n_pdfs = SY.make_tree(rng, topo, True, target_pdfs)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    lda = SY.random_lda(rng)
    c = corpus
    opts = O.mfcc_opts()
    with ThreadPoolExecutor(cores) as ex:
        raw = list(ex.map(lambda u: O.mfcc(c.pcm[c.sample_off[u]:c.sample_off[u + 1]], opts), range(c.n_utts)))
        stats = [O.cmvn_stats([raw[u] for u in range(c.n_utts) if c.utt2spk[u] == s]) for s in range(c.n_spk)]
        feats = list(ex.map(lambda u: O.transform(O.splice(O.cmvn_apply(raw[u], stats[int(c.utt2spk[u])]), 3, 3), lda), range(c.n_utts)))
    frame_off = np.zeros(c.n_utts + 1, np.int64)
    frame_off[1:] = np.cumsum([f.shape[0] for f in feats])
    fp = SY.frame_pdfs_from_truth(corpus, topo, tree, frame_off)
    am = SY.estimate_gmms(np.concatenate(feats), fp, n_pdfs, gauss_per_pdf, rng)
    fsts = [GO.compile_fst(tm, tree, corpus.lexicon, w) for w in corpus.transcripts]
    return SimpleNamespace(corpus=corpus, tm=tm, am=am, tree=tree, lda=lda, feat_mode="lda", frame_off=frame_off, _fsts=fsts)


def pick_sample(sc, audio_seconds: float, also=()):
    """Whole speakers from the start of the corpus until `audio_seconds` are covered, plus the whole speakers of the utterances in
    `also`: per-speaker CMVN statistics need every utterance of a speaker, so a sample that cuts a speaker would give the CPU arm
    different features from the GPU arm's."""
    c = sc.corpus
    dur = (c.sample_off[1:] - c.sample_off[:-1]) / 16000.0
    spk, tot = [], 0.0
    for u in range(c.n_utts):
        s = int(c.utt2spk[u])
        if s not in spk:
            if tot >= audio_seconds:
                break
            spk.append(s)
        tot += float(dur[u])
    for u in also:
        if int(c.utt2spk[u]) not in spk:
            spk.append(int(c.utt2spk[u]))
    keep = set(spk)
    return [u for u in range(c.n_utts) if int(c.utt2spk[u]) in keep]


def train_extras(eng, sc, d_pcm, res, mo, dev, stream, pk, args):
    """K4 (GMM accumulator statistics) and K5 (per-speaker fMLLR statistics + transform update) on the 10 h workload."""
    import torch
    from mfa_b200 import engine as E, fmllr as F
    c = sc.corpus
    raw, fo = eng.mfcc(d_pcm, c.sample_off, mo)
    stats = eng.cmvn_stats(raw, fo, c.utt2spk, c.n_spk)
    eng.sync()
    feats = eng.features(raw, fo, sc.feat_mode, lda=sc.lda, cmvn_stats=stats.cpu().numpy(), utt2spk=c.utt2spk, n_spk=c.n_spk)
    eng.sync()   # the feature kernel reads `raw` on the engine stream: it must finish before torch may recycle that memory
    del raw
    ali = res.ali[: int(fo[-1])].contiguous()
    T, D, G = int(fo[-1]), sc.am.dim, sc.am.NumGauss()
    out = {}

    def timed(fn, reps):
        fn(); eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        eng.sync(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def k4():
        sc.model.acc_zero()
        sc.model.acc_stats(feats, ali)
    k4_ms = timed(k4, max(1, args.steps))
    k4_bytes = float(T * (4 * D + 4) + 8 * G * (1 + 2 * D))
    out["k4_acc_stats"] = {"kernel": "K4 acc_hist + acc_scan + acc_scatter + acc_items_kernel (one acc-stats pass over the step's alignments)",
                           "ms": k4_ms, "bound": "hbm", "algorithmic_bytes": k4_bytes, "achieved": k4_bytes / (k4_ms * 1e-3) / 1e9, "unit": "GB/s",
                           "peak": pk["hbm_gbs"], "frac": k4_bytes / (k4_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "xRT": c.seconds / (k4_ms * 1e-3)}
    with eng.options(acc_impl=1):
        out["k4_acc_stats"]["first_version_atomic_ms"] = timed(k4, 1)
    acc = sc.model.acc_read()
    out["k4_acc_stats"]["frames"] = acc["frames"]
    out["k4_acc_stats"]["avg_loglike_per_frame"] = acc["like"] / max(1.0, acc["frames"])
    sil = c.lexicon.phone_table.get("sil")
    tw = np.where(sc.tm.tid2phone == sil, np.float32(0), np.float32(1)).astype(np.float32)
    tw[0] = 0
    holder = {}

    def k5():
        holder["s"] = sc.model.fmllr_acc(feats, ali, fo, c.utt2spk, c.n_spk, tid_weight=tw)
    k5_ms = timed(k5, max(1, args.steps))
    D1 = D + 1
    weighted = float(holder["s"][:, 0].sum().item())
    dfma = weighted * D * (D1 * (D1 + 1) // 2)
    out["k5_fmllr_stats"] = {"kernel": "K5 fmllr_frame_kernel + fmllr_accum_kernel (per-speaker beta, K, G_d in f64)", "ms": k5_ms, "bound": "f64 FMA",
                             "speakers": int(c.n_spk), "weighted_frames": weighted, "algorithmic_dfma": dfma,
                             "achieved_tflops_f64": 2.0 * dfma / (k5_ms * 1e-3) / 1e12, "xRT": c.seconds / (k5_ms * 1e-3)}
    # config 3's flow on this shard: pass 1 -> fMLLR statistics + update -> pass 2 on the transformed features, all device-resident
    from mfa_b200 import mfa_functions as MF

    def sat():
        holder["sat"] = MF.two_pass_align_pcm(eng, sc.model, sc.model, sc.graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda,
                                              silence_phone_ids=[sil], workspace_bytes=int(args.workspace_gb * (1 << 30)))
    sat(); eng.sync()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        sat()
    eng.sync()
    sat_ms = 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    r1, Wh, (impr_s, cnt_s), r2 = holder["sat"]
    ok1, ok2 = int((r1.status.cpu().numpy() < 2).sum()), int((r2.status.cpu().numpy() < 2).sum())
    l1, l2 = float(r1.total_like.double().sum().item()), float(r2.total_like.double().sum().item())
    out["sat_two_pass"] = {"what": "pass 1 + features + K5 statistics + transform update + pass 2 (fMLLR features), host wall clock around synchronised calls",
                           "ms": sat_ms, "xRT": c.seconds / (sat_ms * 1e-3), "aligned_pass1": ok1, "aligned_pass2": ok2,
                           "sum_loglike_pass1": l1, "sum_loglike_pass2": l2, "speakers": int(c.n_spk)}
    upd = {}

    def k5u():
        upd["r"] = F.compute_transforms_device(eng, holder["s"], D)
    upd_ms = timed(k5u, max(1, args.steps))
    W, impr, cnt = upd["r"]
    out["k5_fmllr_update"] = {"kernel": "fmllr_update_kernel (one CTA per speaker, 40 sweeps of row updates in f64)", "ms": upd_ms,
                              "speakers_updated": int((np.asarray(cnt) > 500).sum()),
                              "mean_objf_impr_per_frame": float(np.sum(impr) / max(1.0, float(np.sum(cnt))))}
    return out


def mfcc_sweep(eng, dev, stream, pk, dist, sizes, seed=1234, chunk_utts=16384):
    """BASELINE config 5: MFCC + per-speaker CMVN statistics (K1) over N utterances of 1-30 s (uniform), N in `sizes`, split evenly over
    the ranks.  The PCM (up to ~0.5 TB for 1 M utterances) does not fit HBM, so it is generated ON THE DEVICE chunk by chunk (noise with a
    slow envelope: K1's cost does not depend on the signal) and only the kernel work is timed: CUDA events on the engine stream around
    mfa_mfcc + mfa_cmvn_stats of each chunk, summed; max over ranks.  GB/s uses SURVEY.md 8(d)'s algorithmic bytes: 2 per sample in,
    4 x 13 per frame out, + the CMVN second pass over the MFCCs (4 x 13 per frame read)."""
    import torch
    from mfa_b200 import engine as E
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mo = E.mfcc_opts()
    rows = []
    for n_total in sizes:
        n_mine = n_total // world + (1 if rank < n_total % world else 0)
        rng = np.random.default_rng(seed + 7919 * rank + n_total)
        ms = 0.0
        samples = frames = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for c0 in range(0, n_mine, chunk_utts):
            n = min(chunk_utts, n_mine - c0)
            lens = (rng.uniform(1.0, 30.0, n) * 16000).astype(np.int64)
            so = np.zeros(n + 1, np.int64)
            so[1:] = np.cumsum(lens)
            N = int(so[-1])
            g = torch.Generator(device=dev)
            g.manual_seed(int(seed + c0 + n_total))
            pcm = torch.empty(N, dtype=torch.int16, device=dev)
            step = 1 << 28
            for a in range(0, N, step):     # bounded temporaries: 1 GiB of float noise at a time
                b = min(N, a + step)
                x = torch.randn(b - a, device=dev, generator=g)
                x *= 3000.0 * (1.0 + 0.5 * torch.sin(torch.arange(a, b, device=dev, dtype=torch.float32) * (2 * np.pi * 4.0 / 16000.0)))
                pcm[a:b] = x.clamp_(-32767, 32767).to(torch.int16)
                del x
            u2s = (np.arange(n) // 64).astype(np.int32)      # ~64 utterances per speaker
            n_spk = int(u2s[-1]) + 1
            torch.cuda.synchronize(dev)
            fo = E.frame_offsets(mo, so)
            out = torch.empty((int(fo[-1]), mo.num_ceps), dtype=torch.float32, device=dev)
            if c0 == 0:   # first chunk of a size warms the tables / workspaces outside the timing
                eng.mfcc(pcm, so, mo, out=out); eng.cmvn_stats(out, fo, u2s, n_spk); eng.sync()
            e0.record(stream)
            eng.mfcc(pcm, so, mo, out=out)
            eng.cmvn_stats(out, fo, u2s, n_spk)
            e1.record(stream)
            eng.sync(); torch.cuda.synchronize(dev)
            ms += e0.elapsed_time(e1)
            samples += N; frames += int(fo[-1])
            del pcm, out
        t = torch.tensor([ms, float(samples), float(frames)], device=dev, dtype=torch.float64)
        if dist is not None:
            mx = t[:1].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t[1:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms_max, samples_all, frames_all = float(mx.item()), float(sm[0].item()), float(sm[1].item())
        else:
            ms_max, samples_all, frames_all = ms, float(samples), float(frames)
        byt = 2.0 * samples_all + 52.0 * frames_all + 52.0 * frames_all
        gbs = byt / (ms_max * 1e-3) / 1e9
        rows.append({"utterances": int(n_total), "audio_hours": samples_all / 16000.0 / 3600.0, "ms": ms_max, "algorithmic_GB": byt / 1e9,
                     "GB_per_s": gbs, "frac_of_hbm_peak_all_gpus": gbs / (pk["hbm_gbs"] * world), "fp32_tflops": 16.0e3 * frames_all / (ms_max * 1e-3) / 1e12,
                     "xRT": samples_all / 16000.0 / (ms_max * 1e-3)})
    return {"what": "config 5: K1 MFCC + CMVN statistics over N utterances of 1-30 s, PCM generated on the device in chunks; kernel time only "
                    "(CUDA events), max over ranks; K1 is fp32-issue bound, not HBM bound (roofline_k1)", "n_gpus": world, "sizes": rows}


def cold_single_shot(eng, sc, dev, args, cores):
    """What a user sees for ONE 10 h job that starts from transcripts and host PCM: training graphs compiled (host C++ threads) and packed,
    graphs + tile plan uploaded, the fused alignment with host buffers, results back on the host -- everything inside the timed region
    except the CUDA context and the model upload (a loaded aligner, as in MFA where the job function constructs GmmAligner once).
    Stages are host wall clock; the alignment call overlaps its own H2D with compute."""
    import torch
    from mfa_b200 import engine as E
    c = sc.corpus
    mo = E.mfcc_opts()
    model = E.DeviceModel(eng, sc.tm, sc.am)
    gc = E.GraphCompiler(sc.tm, sc.tree, c.lexicon)
    h_pcm = torch.from_numpy(c.pcm).pin_memory().numpy()
    eng.sync()
    t = {}
    t0 = time.perf_counter()
    batch = gc.compile(c.transcripts, n_threads=cores)
    t["graph_compile_ms"] = 1e3 * (time.perf_counter() - t0)
    t1 = time.perf_counter()
    graphs = E.Graphs(batch, sc.tm, 1.0, 0.1)
    t["graph_pack_ms"] = 1e3 * (time.perf_counter() - t1)
    t1 = time.perf_counter()
    res = E.align_pcm(eng, model, graphs, h_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda,
                      workspace_bytes=int(args.workspace_gb * (1 << 30)))
    t["first_align_call_ms"] = 1e3 * (time.perf_counter() - t1)
    total = time.perf_counter() - t0
    t["first_call_gpu_stages_ms"] = eng.stage_timing()
    ok = int((res.status < 2).sum())
    t1 = time.perf_counter()
    E.align_pcm(eng, model, graphs, h_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, workspace_bytes=int(args.workspace_gb * (1 << 30)))
    t["second_align_call_ms"] = 1e3 * (time.perf_counter() - t1)   # same graphs / model again: the difference is the one-time upload + planning
    graphs.close(); batch.close()
    # the same job with graph compilation pipelined against the alignment: piece k + 1 is compiled / packed while piece k is aligned
    piped = None
    try:
        eng.sync()
        t1 = time.perf_counter()
        rp, tp = E.align_pcm_from_transcripts(eng, gc, model, c.transcripts, h_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda,
                                              n_segments=4, n_threads=cores, workspace_bytes=int(args.workspace_gb * (1 << 30)))
        tot_p = time.perf_counter() - t1
        same = bool(np.array_equal(rp.status, res.status) and np.array_equal(rp.ali[:len(res.ali)], res.ali[:len(rp.ali)]))
        piped = {"ms": 1e3 * tot_p, "xRT": c.seconds / tot_p, "segments": tp["segments"], "compile_pack_ms": tp["compile_pack_ms"], "align_ms": tp["align_ms"], "close_ms": tp.get("close_ms"), "loop_ms": tp.get("loop_ms"),
                 "identical_to_one_call": same}
    except Exception as ex:
        piped = {"failed": repr(ex)}
    model.close(); gc.close()
    return {"what": "single-shot job: compile graphs + pack + first alignment call (host PCM in, host alignments out), graphs compiled inside "
                    "the timed region; `pipelined` = the same job through align_pcm_from_transcripts (graphs of the next piece compiled while "
                    "the current piece is aligned)", "ms": 1e3 * total, "xRT": c.seconds / total, "stages_ms": t, "aligned_utterances": ok,
            "host_threads": cores, "pipelined": piped}


def train_loop(eng, sc, d_pcm, dev, stream, dist, args):
    """Config 4's loop on this rank's shard, `--train-iters` iterations: align (the fused step, PCM in) -> K4 statistics -> NCCL
    all-reduce of the f64 accumulator block (N > 1) -> M-step ON THE DEVICE (csrc/mstep.cu: GMM update with low-count removal and mix-up
    back to the starting size, transition update, K2 operand images rebuilt) -> transition costs re-folded into the packed graphs.  No
    accumulator and no model parameter crosses PCIe; per iteration the host reads one result struct and the new pdf offsets (16 KB).
    Stage times are host wall clock around synchronised stages on THIS rank; `iteration_ms` is additionally reduced with MAX over ranks.
    A rank that fails tells the others before the next collective (MIN all-reduce of an ok flag), so nobody waits forever."""
    import torch
    from mfa_b200 import engine as E
    c = sc.corpus
    mo = E.mfcc_opts()
    fo = sc.frame_off
    T = int(fo[-1])
    model = E.DeviceModel(eng, sc.tm, sc.am)      # the loop updates its own copy in place; sc.model stays the benchmark's model
    graphs = E.Graphs(sc.batch, sc.tm, 1.0, 0.1)
    model.set_transitions(sc.tm)
    g0 = sc.am.NumGauss()
    model.reserve(2 * g0 + 1)   # both parameter sets + the accumulator block for any M-step with mix-up target g0: no iteration allocates
    raw, _ = eng.mfcc(d_pcm, c.sample_off, mo)
    stats = eng.cmvn_stats(raw, fo, c.utt2spk, c.n_spk)
    eng.sync()
    feats = eng.features(raw, fo, sc.feat_mode, lda=sc.lda, cmvn_stats=stats.cpu().numpy(), utt2spk=c.utt2spk, n_spk=c.n_spk)
    eng.sync()   # the feature kernel reads `raw` on the engine stream: it must finish before torch may recycle that memory
    del raw
    wo_total = int(np.cumsum(graphs.max_words())[-1])
    outs = E._alloc_outputs(T, wo_total, c.n_utts, dev)
    ws = int(args.workspace_gb * (1 << 30))

    def all_ok(ok: bool) -> bool:
        if dist is None:
            return ok
        f = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item() > 0.5)

    iters, failed = [], None
    E.align_pcm(eng, model, graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, workspace_bytes=ws, outputs=outs)
    eng.sync()   # first call with this model / these graphs: plans and uploads are paid here, like the warm-up steps of the main arm
    for it in range(args.train_iters):
        t = {}
        ok = True
        t_it = time.perf_counter()
        try:
            t0 = time.perf_counter()
            res = E.align_pcm(eng, model, graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, workspace_bytes=ws, outputs=outs)
            eng.sync(); t["align_ms"] = 1e3 * (time.perf_counter() - t0)
            t["align_stages_ms"] = eng.stage_timing()
            t["k2_useful_tflop"] = eng.gmm_flops() / 1e12; t["k2_issued_over_useful"] = eng.gmm_issued_flops() / max(1.0, eng.gmm_flops())
            st_it = res.status.cpu().numpy()
            t["retried"] = int((st_it == 1).sum()); t["failed_utts"] = int((st_it >= 2).sum()); t["band_fallbacks_total"] = int(eng.band_fallbacks)
            t0 = time.perf_counter()
            model.acc_zero()
            model.acc_stats(feats, res.ali[:T])
            eng.sync(); t["acc_stats_ms"] = 1e3 * (time.perf_counter() - t0)
        except Exception as ex:
            ok, failed = False, repr(ex)
        if not all_ok(ok):
            failed = failed or "another rank failed"
            break
        t0 = time.perf_counter()
        if dist is not None:
            acc_t = model.acc_tensor()
            dist.all_reduce(acc_t)
            torch.cuda.synchronize(dev)
            t["allreduce_bytes"] = int(acc_t.numel() * 8)
        t["allreduce_ms"] = 1e3 * (time.perf_counter() - t0)
        try:
            t0 = time.perf_counter()
            r = model.mle_update(mixup=g0, update_transitions=True, seed=1234 + it)
            t["mstep_device_ms"] = 1e3 * (time.perf_counter() - t0)
            t0 = time.perf_counter()
            graphs.set_transitions(eng, model, 1.0, 0.1)
            eng.sync(); t["refold_graphs_ms"] = 1e3 * (time.perf_counter() - t0)
            t["avg_loglike_per_frame"] = r["tot_like"] / max(1.0, r["tot_frames"])
            t["frames_all_ranks"] = r["tot_frames"]
            t["gaussians"] = r["num_gauss_after"]; t["removed"] = r["num_removed"]; t["split"] = r["num_split"]
            t["layout_changed"] = r["layout_changed"]
        except Exception as ex:
            ok, failed = False, repr(ex)
        if not all_ok(ok):
            failed = failed or "another rank failed"
            break
        t["iteration_ms"] = 1e3 * (time.perf_counter() - t_it)
        if dist is not None:
            m = torch.tensor([t["iteration_ms"]], device=dev, dtype=torch.float64)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            t["iteration_ms_max_over_ranks"] = float(m.item())
        iters.append(t)
    model.close()
    graphs.close()
    out = {"iterations": iters, "hours_all_ranks": c.seconds / 3600.0 * (dist.get_world_size() if dist is not None else 1),
           "note": "config 4: align -> K4 -> NCCL all-reduce -> device M-step (mix-up to the starting size) -> graph re-fold; wall clock per "
                   "synchronised stage on rank 0; avg_loglike_per_frame must not decrease"}
    if failed:
        out["failed"] = failed
    return out


def reference_arm(args, real_stdout):
    """`--impl reference`: the reference's CPU implementation of the path -- kalpy / Kaldi cannot be installed offline, so this is the
    oracle port (oracle/oracle.c, the -O3 -mavx2 build on AVX2 hosts, one utterance per host thread) -- on all host cores, on a bounded sample of the same workload
    shape (same corpus generator, same model recipe: 4 000 pdfs / ~40 k Gaussians / D = 40), each step one pass over the sample.
    Nothing of this repo's CUDA library is loaded or called: scenario, features, graphs and alignment all come from oracle/."""
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    seconds = min(args.hours * 3600.0, args.cpu_sample_seconds or 7200.0)   # 2 h: ~590 utterances of several speakers, ~0.5 s of 16-core work per step
    t0 = time.time()
    sc = reference_scenario(seconds, 1234, args.pdfs, args.gauss_per_pdf, cores)
    log(f"reference scenario ({sc.corpus.n_utts} utts, {sc.corpus.seconds:.0f} s, {sc.am.NumGauss()} Gaussians) built in {time.time() - t0:.1f}s without libmfa_b200.so")
    c = sc.corpus
    utts = list(range(c.n_utts))
    workload = (f"configs[1]: triphone LDA-shaped GMM-HMM ({sc.am.NumPdfs()} pdfs, {sc.am.NumGauss()} Gaussians, D={sc.am.dim}), "
                f"{c.seconds / 3600:.2f} h sample of the synthetic 16 kHz workload ({c.n_utts} utts, {c.n_spk} speakers), beam 10 / retry 40")
    config = {"workload": workload, "hours_per_gpu": round(c.seconds / 3600, 3), "utterances_per_gpu": c.n_utts, "pdfs": sc.am.NumPdfs(),
              "gaussians": sc.am.NumGauss(), "dim": sc.am.dim, "beam": 10, "retry_beam": 40, "parallelism": f"{cores} host threads"}
    for _ in range(args.warmup):
        cpu_reference_pass(sc, utts[: max(1, len(utts) // 8)], cores)
    times = []
    for _ in range(args.steps):
        dt, secs, ok, _ = cpu_reference_pass(sc, utts, cores)
        times.append(dt)
    T = float(np.sum(times))
    val = secs * args.steps / T
    sample = f"{len(utts)} utterances / {secs:.0f} audio-s of the same synthetic workload shape per step; oracle port (kalpy/Kaldi not installable offline)"
    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if "libmfa_b200" in ln]
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * T / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "impl": "reference",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "oracle_build": O.VARIANT},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "cuda_library_loaded": bool(loaded)}
    real_stdout.write(json.dumps(line) + "\n"); real_stdout.flush()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hours", type=float, default=10.0, help="audio hours per GPU (configs[1] = 10)")
    ap.add_argument("--pdfs", type=int, default=4000)
    ap.add_argument("--gauss-per-pdf", type=int, default=10)
    ap.add_argument("--gmm-impl", type=int, default=0)
    ap.add_argument("--cpu-sample-seconds", type=float, default=0.0, help="audio seconds for the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workspace-gb", type=float, default=100.0)
    ap.add_argument("--train-iters", type=int, default=4, help="iterations of the align -> acc-stats -> all-reduce -> update loop timed under extras.train_loop")
    ap.add_argument("--extras-dist", action="store_true", help="multi-rank runs: also time K4 / K5 / the SAT two-pass flow per rank (the training loop with its NCCL all-reduce always runs)")
    ap.add_argument("--same-shards", action="store_true", help="multi-rank runs: every rank gets the SAME corpus (seed 1234): separates data effects (a rank owning a slow utterance) from system effects in the per-rank table")
    ap.add_argument("--seed-offset", type=int, default=None, help="corpus seed = 1234 + this instead of 1234 + rank: reproduces a given rank's shard of a multi-GPU run on one GPU")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin this process to the CPUs of the GPU's NUMA node")
    ap.add_argument("--sweep-max", type=int, default=100000, help="largest utterance count of the config-5 MFCC sweep under extras (1000000 = BASELINE's full sweep; ~10 s more)")
    ap.add_argument("--no-extras", action="store_true", help="skip the K4 / K5 timings reported under 'extras'")
    ap.add_argument("--e2e-jobs", type=int, default=2, help="concurrent jobs (engines) per GPU in the end-to-end arm")
    args = ap.parse_args()

    # the ONE JSON line goes to the real stdout; anything a library prints on fd 1 meanwhile (NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0
    if args.impl == "reference":
        return reference_arm(args, real_stdout)
    import torch
    import __graft_entry__ as G
    if rank == 0:
        G.build()
    from mfa_b200 import engine as E, scenario as SC

    dist = None
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = {"bound": False} if (args.no_numa_bind or args.impl == "reference" or world == 1) else bind_to_gpu_numa(local_rank)
    eng = E.Engine(local_rank)
    cores = os.cpu_count() or 1
    seconds = args.hours * 3600.0
    t0 = time.time()
    sc = SC.build(eng, seconds, seed=1234 + (args.seed_offset if args.seed_offset is not None else (0 if args.same_shards else rank)), target_pdfs=args.pdfs, gauss_per_pdf=args.gauss_per_pdf, n_threads=max(1, cores // max(1, world)),
                  synth_device=dev, log=log if rank == 0 else None, model_seed=1234 if (world > 1 or args.seed_offset is not None) else None)
    if args.seed_offset:
        # reproduce rank `seed_offset` of a multi-GPU run: that rank aligns ITS shard with rank 0's acoustic model
        sc0 = SC.build(eng, seconds, seed=1234, target_pdfs=args.pdfs, gauss_per_pdf=args.gauss_per_pdf, n_threads=cores, synth_device=dev, model_seed=1234)
        sc.am = sc0.am
        sc.model.close()
        sc.model = E.DeviceModel(eng, sc.tm, sc.am)
        sc0.model.close(); sc0.graphs.close(); sc0.batch.close()
        del sc0
    if dist is not None:
        # one replicated acoustic model (rank 0's estimate), each rank its own shard of utterances: what MFA's jobs see
        box = [(sc.am.dim, sc.am.offsets, sc.am.weights, sc.am.means_invvars, sc.am.inv_vars) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            from mfa_b200.kaldi_io import AmDiagGmm
            sc.am = AmDiagGmm(*box[0])
            sc.model.close()
            sc.model = E.DeviceModel(eng, sc.tm, sc.am)
    c = sc.corpus
    audio_s = c.seconds
    n_frames = int(sc.frame_off[-1])
    workload = (f"configs[1]: triphone LDA-shaped GMM-HMM ({sc.am.NumPdfs()} pdfs, {sc.am.NumGauss()} Gaussians, D={sc.am.dim}), "
                f"{audio_s / 3600:.2f} h synthetic 16 kHz audio per GPU ({c.n_utts} utts), beam 10 / retry 40")
    config = {"workload": workload, "hours_per_gpu": round(audio_s / 3600, 3), "utterances_per_gpu": c.n_utts, "pdfs": sc.am.NumPdfs(),
              "gaussians": sc.am.NumGauss(), "dim": sc.am.dim, "beam": 10, "retry_beam": 40, "parallelism": f"utterance-sharded x{world}",
              "l2": "inputs larger than L2 (PCM >> 126 MB per step); no explicit flush"}
    log(f"setup {time.time() - t0:.1f}s")

    # ---- device-resident arm -----------------------------------------------------------------------------------
    mo = E.mfcc_opts()
    wo_total = int(np.cumsum(sc.graphs.max_words())[-1])
    d_pcm = torch.from_numpy(c.pcm).to(dev)
    outs = E._alloc_outputs(n_frames, wo_total, c.n_utts, dev)
    outs_b = E._alloc_outputs(n_frames, wo_total, c.n_utts, dev)   # consecutive steps write alternate output sets, as consecutive batches would:
    ws = int(args.workspace_gb * (1 << 30))                          # the engine lets step i+1's K1 / features start under step i's Viterbi tail
    step_no = [0]

    def step_device():
        step_no[0] += 1
        return E.align_pcm(eng, sc.model, sc.graphs, d_pcm, c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda, gmm_impl=args.gmm_impl,
                           workspace_bytes=ws, outputs=outs if step_no[0] & 1 else outs_b)

    stream = torch.cuda.ExternalStream(eng.stream, device=dev)

    def barrier():
        eng.sync()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()

    for _ in range(max(0, args.warmup)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count
    fb0 = eng.band_fallbacks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = step_device()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    gmm_ms, gmm_n, gmm_rows = eng.gmm_timing()   # K2 launches of the LAST step (event pairs on the engine stream)
    gmm_flops_last = eng.gmm_flops()
    gmm_issued_last = eng.gmm_issued_flops()
    stage_ms = eng.stage_timing()
    launches = eng.launch_count - l0
    fallbacks = eng.band_fallbacks - fb0
    gpu_host = tuple(x.cpu().numpy() for x in (res.ali, res.per_frame, res.words, res.num_words, res.total_like, res.status))
    other = outs_b if res.ali is outs[0] else outs          # the step before the last wrote the other set: same input, must be the same output
    steps_identical = bool(args.steps + args.warmup < 2 or all(bool((x == y).all().item()) for x, y in zip((res.ali, res.words, res.status), (other[0], other[2], other[5]))))
    import zlib
    out_crc = {k: zlib.crc32(np.ascontiguousarray(v).tobytes()) for k, v in zip(("ali", "per_frame", "words", "num_words", "total_like", "status"), gpu_host)}
    st = gpu_host[5]
    retried = np.nonzero(st == 1)[0]
    n_ok = int((st < 2).sum())
    if dist is not None:
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max = float(t.item())
        a = torch.tensor([audio_s, float(launches), float(n_ok), float(c.n_utts)], device=dev, dtype=torch.float64)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        audio_total, launches_total, ok_total, utts_total = (float(x) for x in a.tolist())
    else:
        dev_ms_max, audio_total, launches_total, ok_total, utts_total = dev_ms, audio_s, float(launches), float(n_ok), float(c.n_utts)
    value = audio_total * args.steps / (dev_ms_max / 1000.0)
    # per-rank view (every rank has its own corpus, so one rank may own a slow Viterbi tail or a band fallback): device time per step,
    # per-stage CUDA-event times of the last step, fallbacks, and the part of the step no stage accounts for
    mine = {"rank": rank, "dev_ms_per_step": dev_ms / args.steps, "stages_ms": stage_ms, "k3_band_fallbacks": int(fallbacks),
            "unaccounted_ms": dev_ms / args.steps - float(sum(stage_ms.values())), "utterances": int(c.n_utts), "frames": n_frames,
            "longest_utt_frames": int((sc.frame_off[1:] - sc.frame_off[:-1]).max()), "retried": int(retried.size), "numa": numa}
    per_rank = [mine]
    if dist is not None:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    def spread(key):
        xs = [float(key(r)) for r in per_rank]
        return {"min": min(xs), "median": float(np.median(xs)), "max": max(xs), "argmax_rank": int(np.argmax(xs))}
    rank_summary = {"dev_ms_per_step": spread(lambda r: r["dev_ms_per_step"]), "unaccounted_ms": spread(lambda r: r["unaccounted_ms"]),
                    "k3_band_fallbacks": spread(lambda r: r["k3_band_fallbacks"]), "longest_utt_frames": spread(lambda r: r["longest_utt_frames"]),
                    **{f"stage_{k}_ms": spread(lambda r, k=k: r["stages_ms"][k]) for k in stage_ms},
                    "same_shards": bool(args.same_shards)}

    # ---- end-to-end arm: host (pinned) PCM in, host results out, through the same C-ABI call ---------------------
    h_pcm = torch.from_numpy(c.pcm).pin_memory()
    h_outs = [torch.zeros(x.shape, dtype=x.dtype).pin_memory() for x in outs]
    h_np = tuple(x.numpy() for x in h_outs)

    # MFA runs several jobs per device, each with its own aligner (alignment/multiprocessing.py: one GmmAligner per job,
    # threads or processes); the end-to-end arm does the same: `--e2e-jobs` engines, one host thread each (ctypes releases the
    # GIL), every step a full pass over the per-GPU workload with its own H2D and D2H inside the timed region.  While one job's
    # PCM crosses PCIe the other job's kernels run.  The single-job figure is reported next to it.
    import threading
    n_jobs = max(1, args.e2e_jobs)
    jobs = [(eng, sc.model, sc.graphs, h_np)]
    for _ in range(1, n_jobs):
        e2 = E.Engine(local_rank)
        jobs.append((e2, E.DeviceModel(e2, sc.tm, sc.am), E.Graphs(sc.batch, sc.tm, 1.0, 0.1),
                     tuple(torch.zeros(x.shape, dtype=x.dtype).pin_memory().numpy() for x in outs)))

    def step_host(j=0):
        en, mdl, gr, ho = jobs[j]
        return E.align_pcm(en, mdl, gr, h_pcm.numpy(), c.sample_off, c.utt2spk, c.n_spk, mo, sc.feat_mode, lda=sc.lda,
                           gmm_impl=args.gmm_impl, workspace_bytes=ws, outputs=ho)

    def timed_host(n_threads, steps_total, stagger_s=0.0):
        per = [steps_total // n_threads + (1 if j < steps_total % n_threads else 0) for j in range(n_threads)]
        go = threading.Barrier(n_threads + 1)

        def work(j):
            go.wait()
            if j and stagger_s > 0:
                time.sleep(j * stagger_s / n_threads)   # inside the timed region: de-phases the jobs' uploads
            for _ in range(per[j]):
                step_host(j)   # returns after the D2H copies have landed (MFA_HOST contract)

        th = [threading.Thread(target=work, args=(j,)) for j in range(n_threads)]
        for t_ in th:
            t_.start()
        barrier()
        go.wait()
        t0 = time.perf_counter()
        for t_ in th:
            t_.join()
        for j in range(n_threads):
            jobs[j][0].sync()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt

    for j in range(n_jobs):
        step_host(j)
    e2e1_s = timed_host(1, args.steps)
    e2e_stage_ms = eng.stage_timing()   # last single-job host-buffer step: the MFCC interval includes waiting for the PCM pieces
    e2e_steps = (args.steps + n_jobs - 1) // n_jobs * n_jobs
    e2e_s = timed_host(n_jobs, e2e_steps, e2e1_s / args.steps) if n_jobs > 1 else e2e1_s
    e2e_val = audio_total * e2e_steps / e2e_s if n_jobs > 1 else audio_total * args.steps / e2e1_s
    e2e1_val = audio_total * args.steps / e2e1_s
    clocks = sampler.stop()   # sampled across BOTH timed regions (device-resident steps and the end-to-end steps)
    # what the host -> device link gives each rank while ALL ranks copy at once (the end-to-end arm's floor at N > 1: GPUs may share PCIe
    # switch uplinks, host DRAM and, across sockets, the interconnect): 4 copies of the step's pinned PCM, CUDA events, after a barrier
    probe_buf = torch.empty(h_pcm.numel(), dtype=h_pcm.dtype, device=dev)
    probe_buf.copy_(h_pcm, non_blocking=True)
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(4):
        probe_buf.copy_(h_pcm, non_blocking=True)
    pe1.record()
    torch.cuda.synchronize(dev)
    probe_gbs = 4.0 * c.pcm.nbytes / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
    del probe_buf
    probe_all = [probe_gbs]
    if dist is not None:
        probe_all = [None] * world
        dist.all_gather_object(probe_all, probe_gbs)
    h2d_probe = {"gbs_per_rank": [round(x, 2) for x in probe_all], "concurrent_ranks": world,
                 "floor_ms_per_step": c.pcm.nbytes / (min(probe_all) * 1e9) * 1e3,
                 "what": "pinned-host -> device bandwidth of every rank while all ranks copy at once; floor = this step's PCM bytes at the slowest rank's rate",
                 "host_cpus": len(os.sched_getaffinity(0)), "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")}
    clocks["window"] = "device-resident timed steps + end-to-end timed steps"
    h2d = int(c.pcm.nbytes)
    d2h = int(sum(x.numel() * x.element_size() for x in h_outs))

    # ---- rooflines: per-stage CUDA-event times of the last device-resident step; the dominant kernel goes into "roofline"
    pk = measured_peaks()
    stages = stage_ms
    step_ms = dev_ms / args.steps
    k2 = None
    if gmm_n > 0 and gmm_ms > 0:
        # useful FLOPs = 2*(2D+1) per (frame, Gaussian) actually scored: the fused pipeline scores, per utterance, only the pdfs
        # its graph references (what Kaldi's decodable evaluates lazily), not frames x all Gaussians
        per_launch_flops = gmm_flops_last / gmm_n
        avg_ms = gmm_ms / gmm_n
        achieved = per_launch_flops / (avg_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_gmm_tc_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            key = {0: "dram_bytes_per_flop_ragged", 2: "dram_bytes_per_flop_dense"}.get(args.gmm_impl)
            if key in tj:   # DRAM bytes per useful FLOP from the committed ncu --set full capture x FLOPs of this launch
                traffic = tj[key] * per_launch_flops
        k2 = {"kernel": "K2 gmm log-likelihoods (xsplit + gather_b + gmm_tc_kernel)", "bound": "tensor", "achieved": achieved,
              "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"], "traffic": traffic,
              "peak_source": pk["source"] + " bf16 sustained", "issued_over_useful_flops": gmm_issued_last / max(1.0, gmm_flops_last),
              "issued_over_useful_note": "3 fp16 products x K padding (80 / 81) x tile fill (columns of a 128-wide Gaussian tile and frames of a 256-frame pair that exist)",
              "scored": "per-utterance pdf subsets" if args.gmm_impl == 0 else "all pdfs", "launches_per_step": gmm_n, "avg_launch_ms": avg_ms,
              "algorithmic_flops_per_launch": per_launch_flops, "share_of_step": gmm_ms / step_ms}
    # K3 (band kernel): algorithmic bytes per utterance = T * (4 P_u log-likelihoods read once + 512 back-pointer row
    # written + 512 read by the back-trace + 8 outputs) + its graph (4 B per state + 8 B per arc, read through L1)
    so, ao, po = sc.graphs.offsets()
    T_u = (sc.frame_off[1:] - sc.frame_off[:-1]).astype(np.float64)
    S_u, A_u, P_u = np.diff(so).astype(np.float64), np.diff(ao).astype(np.float64), np.diff(po).astype(np.float64)
    k3_bytes = float((T_u * (4 * P_u + 512 + 512 + 8) + 8 * A_u + 4 * S_u).sum())
    k3 = None
    if stages["viterbi"] > 0:
        ach = k3_bytes / (stages["viterbi"] * 1e-3) / 1e9
        k3 = {"kernel": "K3 viterbi_band_kernel (all shared-memory classes on side streams, fork to join)", "bound": "hbm", "achieved": ach,
              "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"] + " HBM copy",
              "algorithmic_bytes_per_launch": k3_bytes, "avg_launch_ms": stages["viterbi"], "share_of_step": stages["viterbi"] / step_ms,
              "note": "latency-bound: a sequential per-frame recursion per utterance (2 warps each); bounded by resident utterances x "
                      "per-frame dependency chain, not by bytes -- see profiles/r1_viterbi_band_full.md for stall reasons"}
    # K1: bytes = 2 per sample + 4*13 per frame
    k1_bytes = float(2 * c.pcm.shape[0] + 52 * n_frames)
    k1 = None
    if stages["mfcc_cmvn"] > 0:
        ach = k1_bytes / (stages["mfcc_cmvn"] * 1e-3) / 1e9
        k1 = {"kernel": "K1 mfcc512_kernel + CMVN statistics", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
              "frac": ach / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": stages["mfcc_cmvn"],
              "share_of_step": stages["mfcc_cmvn"] / step_ms,
              "achieved_fp32_tflops": 16.0e3 * n_frames / (stages["mfcc_cmvn"] * 1e-3) / 1e12,   # ~16 kFLOP per frame
              "note": "fp32-ALU / issue bound (register FFT, ~1 080 warp instructions per frame, 38 % of them FADD/FFMA/FMUL), not HBM bound -- see profiles/r1_mfcc512_full.md"}
    cands = [x for x in (k2, k3, k1) if x]
    # the dominant kernel is K2 (the tensor-core scoring kernel: largest share of device work; K3's interval is as long but it is a
    # latency chain that overlaps the next step); `roofline` stays on it so that the figure does not flip between runs
    roof = k2 if k2 else (max(cands, key=lambda x: x["share_of_step"]) if cands else None)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches_total),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "jobs_per_gpu": n_jobs,
                    "steps": e2e_steps if n_jobs > 1 else args.steps, "single_job_value": e2e1_val, "single_job_stage_ms": e2e_stage_ms,
                    "h2d_probe": h2d_probe},
            "roofline": roof, "roofline_k2": k2, "roofline_k3": k3, "roofline_k1": k1, "stages_ms": stages,
            "aligned_utterances": int(ok_total), "utterances": int(utts_total),
            "k3_band_fallbacks_per_step": fallbacks / max(1, args.steps), "per_rank": rank_summary,
            "pipelining": {"what": "stages_ms are per-stage CUDA-event intervals of the LAST step; the Viterbi launch of step i is joined on its own stream, "
                                   "so K1 / features of step i+1 run under its tail and the intervals overlap (their sum exceeds ms_per_step); engine option "
                                   "k3_overlap = 0 serialises them", "k3_overlap": eng.get_option("k3_overlap"),
                           "consecutive_steps_identical_outputs": steps_identical, "output_crc32_rank0": out_crc,
                           "input_pcm_crc32_rank0": zlib.crc32(c.pcm.tobytes()),
                           "crc_note": "the synthetic waveform is generated with torch CUDA scans whose float summation order varies run to run, so "
                                       "input (and output) CRCs differ between processes; within a process the same input gives the same output"},
            "per_rank_rows": per_rank if world > 1 else None}
    # ---- training-loop / SAT stages next to the alignment path (config 4 and config 3's fMLLR pass), timed on their own with CUDA
    # events on the engine stream, OUTSIDE the timed alignment step: K4 accumulator statistics and K5 per-speaker fMLLR statistics
    # over the step's device-resident features and alignments.
    # (they run on every rank at every N: no collective inside, a failure is caught per rank; config 3 -- SAT two-pass on speaker shards --
    # is then reported per rank: every rank's shard holds its own speakers, the way MFA's assign_jobs splits a corpus)
    if not args.no_extras:
        line["extras"] = {}
        try:
            line["extras"].update(train_extras(eng, sc, d_pcm, res, mo, dev, stream, pk, args))
        except Exception as ex:
            line["extras"]["failed"] = repr(ex)
        if dist is not None:
            mine_sat = line["extras"].get("sat_two_pass", {}).get("ms")
            sat_all = [None] * world
            dist.all_gather_object(sat_all, (mine_sat, float(c.seconds), int(c.n_spk)))
            good = [x for x in sat_all if x[0] is not None]
            if good:
                ms = [x[0] for x in good]
                line["extras"]["sat_two_pass_all_ranks"] = {
                    "what": "config 3: two-pass SAT alignment, one speaker shard per GPU (speakers never span ranks, so fMLLR needs no collective); "
                            "job time = the slowest rank", "ranks": len(good), "hours_all_ranks": sum(x[1] for x in good) / 3600.0,
                    "speakers_all_ranks": sum(x[2] for x in good), "ms_min": min(ms), "ms_median": float(np.median(ms)), "ms_max": max(ms),
                    "xRT_job": sum(x[1] for x in good) / (max(ms) * 1e-3)}
        # config 4's loop runs at every N: its all-reduce is the one collective of the path (a rank that fails before a collective
        # tells the others through a MIN all-reduce of an ok flag, so nobody waits forever)
        try:
            line["extras"]["train_loop"] = train_loop(eng, sc, d_pcm, dev, stream, dist, args)
        except Exception as ex:
            line["extras"]["train_loop"] = {"failed": repr(ex)}
        try:
            del d_pcm
            torch.cuda.empty_cache()
            sizes = [n for n in (1000, 10000, 100000, 1000000) if n <= args.sweep_max]
            line["extras"]["mfcc_sweep"] = mfcc_sweep(eng, dev, stream, pk, dist, sizes)
        except Exception as ex:
            line["extras"]["mfcc_sweep"] = {"failed": repr(ex)}
        if world == 1:
            try:
                line["cold_e2e"] = cold_single_shot(eng, sc, dev, args, cores)
            except Exception as ex:
                line["cold_e2e"] = {"failed": repr(ex)}
            # the MFA-shaped FILE flow (wav files -> MfccFunction -> CMVN -> features -> graph archives -> pass 1 -> fMLLR -> pass 2 ->
            # TextGrids, all through the kalpy-compatible classes and Kaldi archives on disk) on 1 h of audio with a small model: what a
            # maintainer who only swaps the imports gets; bound by per-utterance Python / file work, not by the GPU
            try:
                import shutil
                import tempfile
                sys.path.insert(0, os.path.join(ROOT, "examples"))
                import two_pass_alignment as flow
                tmp = tempfile.mkdtemp(prefix="mfa_b200_flow_")
                try:
                    from pathlib import Path
                    line["file_flow"] = flow.run(Path(tmp), 3600.0, quiet=True, n_jobs=4, threads=True)
                    line["file_flow"]["what"] = ("examples/two_pass_alignment.py on 1 h of synthetic audio (200-pdf triphone LDA model, 4 jobs as threads of this "
                                                 "process, one engine per thread -- MFA's USE_THREADING mode): wall clock from wav files to TextGrids incl. every "
                                                 "archive written and read; first-call costs of the process included")
                finally:
                    shutil.rmtree(tmp, ignore_errors=True)
            except Exception as ex:
                line["file_flow"] = {"failed": repr(ex)}
    if rank == 0 and not args.no_cpu_baseline and world == 1:   # reported at N = 1 only (the reference arm times the CPU path at every N)
        try:
            from oracle import oracle as O
            sc._fsts = sc.batch.export()
            sample_s = args.cpu_sample_seconds or 7200.0
            utts = pick_sample(sc, sample_s, also=[int(u) for u in retried[:2]])   # a retry-beam utterance is part of the parity sample
            cpu_reference_pass(sc, utts[: max(1, len(utts) // 10)], cores)
            dt, secs, ok, ref = cpu_reference_pass(sc, utts, cores)
            line["cpu_baseline"] = {"value": secs / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{len(utts)} utterances / {secs:.0f} audio-s of the same workload, {dt:.1f} s wall; oracle port",
                                    "oracle_build": O.VARIANT}
            line["parity"] = parity_block(sc, utts, ref, gpu_host)
            line["frame_agreement_pct"] = line["parity"]["frame_agreement_pct"]
        except Exception as ex:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {ex}"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n"); real_stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
