"""Shared test utilities: golden-fixture loading, oracle-side pipelines, small synthetic scenarios."""
import json
import os

import numpy as np

from mfa_b200 import kaldi_io as K, lexicon as LX, synth as SY
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fixtures.npz")
_gold = None


def gold():
    global _gold
    if _gold is None:
        _gold = np.load(GOLD)
    return _gold


def _topo(js):
    d = json.loads(bytes(js).decode())
    entries = [[K.HmmState(s[0], s[1], [(int(a), float(b)) for a, b in s[2]]) for s in e] for e in d["entries"]]
    return K.Topology(np.asarray(d["phones"], np.int32), np.asarray(d["phone2idx"], np.int32), entries)


def _tree(g, prefix):
    N, P, root = (int(x) for x in g[f"{prefix}_tree_np"])
    node, aux_off, aux = g[f"{prefix}_tree_nodes"], g[f"{prefix}_tree_aux_off"], g[f"{prefix}_tree_aux"]
    cd = K.ContextDependency(N, P)
    for i, (t, key, a, b) in enumerate(node):
        cd.nodes.append((int(t), int(key), int(a), int(b)))
        if t == 1:
            cd.sets[i] = aux[aux_off[i]:aux_off[i + 1]]
        elif t == 2:
            cd.tables[i] = [int(x) for x in aux[aux_off[i]:aux_off[i + 1]]]
    cd.root = root
    return cd


def load_model(prefix):
    """prefix in {'mono','g2p'} -> (tm, am, tree) rebuilt from the derived golden arrays."""
    g = gold()
    tm = K.TransitionModel(_topo(g[f"{prefix}_topo"]), g[f"{prefix}_tuples"], g[f"{prefix}_log_probs"])
    am = K.AmDiagGmm(int(g[f"{prefix}_dim"][0]), g[f"{prefix}_offsets"], g[f"{prefix}_weights"], g[f"{prefix}_miv"], g[f"{prefix}_iv"],
                     g[f"{prefix}_stored_gconsts"])
    return tm, am, _tree(g, prefix)


def mono_sample_setup(tmp_path):
    """Config 1: the reference's sample utterance + its fixture monophone model + test_acoustic dictionary."""
    g = gold()
    tm, am, tree = load_model("mono")
    phones = json.loads(bytes(g["mono_phones"]).decode())
    pt = LX.make_phone_table(phones, ("sil", "sp", "spn"), True)
    dpath = os.path.join(str(tmp_path), "dict.txt")
    with open(dpath, "wb") as f:
        f.write(bytes(g["test_acoustic_dict"]))
    lex = LX.Lexicon(LX.parse_dictionary(dpath), pt, position_dependent_phones=True)
    text = bytes(g["acoustic_corpus_lab"]).decode().strip().lower()
    return dict(tm=tm, am=am, tree=tree, lex=lex, pt=pt, text=text, pcm=g["acoustic_corpus_pcm"], pcm2=g["cold_corpus_pcm"])


def oracle_features(pcm_list, utt2spk, n_spk, mode="deltas", lda=None, fmllr=None, opts=None, cmvn=True):
    """Oracle feature chain per utterance: MFCC -> per-speaker CMVN -> deltas | splice+LDA -> fMLLR."""
    opts = opts or O.mfcc_opts()
    raw = [O.mfcc(p, opts) for p in pcm_list]
    stats = []
    for s in range(n_spk):
        fl = [raw[u] for u in range(len(raw)) if utt2spk[u] == s and raw[u].shape[0] > 0]
        stats.append(O.cmvn_stats(fl) if fl else np.zeros((2, opts.num_ceps + 1)))
    out = []
    for u, m in enumerate(raw):
        if m.shape[0] == 0:
            D = 3 * m.shape[1] if mode == "deltas" else (lda.shape[0] if mode == "lda" else m.shape[1])
            out.append(np.zeros((0, D), np.float32))
            continue
        x = O.cmvn_apply(m, stats[utt2spk[u]]) if cmvn else m
        if mode == "deltas":
            x = O.add_deltas(x)
        elif mode == "lda":
            x = O.transform(O.splice(x, 3, 3), lda)
        if fmllr is not None:
            x = O.transform(x, fmllr[utt2spk[u]])
        out.append(x)
    return raw, np.stack(stats), out


def build_synth_scenario(seconds=40.0, seed=7, triphone=False, n_phones=12, n_words=60, target_pdfs=120, gauss_per_pdf=3,
                         use_lda=False, n_spk=3, position_dependent=False, mean_utt_s=4.0, min_utt_s=1.0, max_utt_s=8.0):
    """Small synthetic corpus + a model estimated from its oracle features (CPU)."""
    corpus = SY.make_corpus(seconds, seed=seed, n_phones=n_phones, n_words=n_words, n_spk=n_spk, position_dependent=position_dependent,
                            mean_utt_s=mean_utt_s, min_utt_s=min_utt_s, max_utt_s=max_utt_s)
    rng = np.random.default_rng(seed + 1)
    topo = SY.make_topology(corpus.phone_table)
    tree, n_pdfs = SY.make_tree(rng, topo, triphone, target_pdfs)
    tm = SY.make_transition_model(topo, tree, n_pdfs)
    pcm_list = [corpus.pcm[corpus.sample_off[u]:corpus.sample_off[u + 1]] for u in range(corpus.n_utts)]
    lda = SY.random_lda(rng) if use_lda else None
    raw, stats, feats = oracle_features(pcm_list, corpus.utt2spk, corpus.n_spk, "lda" if use_lda else "deltas", lda)
    frame_off = np.zeros(corpus.n_utts + 1, np.int64)
    frame_off[1:] = np.cumsum([f.shape[0] for f in feats])
    fp = SY.frame_pdfs_from_truth(corpus, topo, tree, frame_off)
    am = SY.estimate_gmms(np.concatenate(feats), fp, n_pdfs, gauss_per_pdf, rng)
    return dict(corpus=corpus, tm=tm, am=am, tree=tree, lda=lda, pcm_list=pcm_list, raw=raw, stats=stats, feats=feats, frame_off=frame_off)


def oracle_align_all(sc, fsts, beam=10.0, retry_beam=40.0, acoustic_scale=0.1, transition_scale=1.0, self_loop_scale=0.1, feats=None):
    tm, am = sc["tm"], sc["am"]
    g = O.GmmModel.from_am(am)
    tid_cost = -tm.scaled_transition_log_probs(transition_scale, self_loop_scale)
    feats = feats if feats is not None else sc["feats"]
    return [O.align(fsts[u], tid_cost, g, tm.tid2pdf, feats[u], feats[u].shape[0], acoustic_scale, beam, retry_beam) for u in range(len(fsts))]
