"""Generates the committed fixtures under tests/golden/ from the reference's own test data
(/root/reference/tests/data, read-only) and from torchaudio's independent Kaldi-MFCC port.
Run here (the build container); the GPU box only sees the committed .npz files.

  python tests/golden/make_golden.py
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mfa_b200 import kaldi_io as K  # noqa: E402

REF = "/root/reference/tests/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def topo_json(topo):
    return json.dumps({"phones": topo.phones.tolist(), "phone2idx": topo.phone2idx.tolist(),
                       "entries": [[[s.forward_pdf_class, s.self_loop_pdf_class, [[int(d), float(p)] for d, p in s.transitions]] for s in e]
                                   for e in topo.entries]})


def tree_arrays(tree):
    node, aux_off, aux, root = tree.flatten()
    return dict(tree_np=np.asarray([tree.N, tree.P, root], np.int32), tree_nodes=node, tree_aux_off=aux_off, tree_aux=aux)


def model_arrays(prefix, tm, am, tree):
    d = {f"{prefix}_topo": np.frombuffer(topo_json(tm.topo).encode(), dtype=np.uint8), f"{prefix}_tuples": tm.tuples,
         f"{prefix}_log_probs": tm.log_probs, f"{prefix}_dim": np.asarray([am.dim], np.int32), f"{prefix}_offsets": am.offsets,
         f"{prefix}_weights": am.weights, f"{prefix}_miv": am.means_invvars, f"{prefix}_iv": am.inv_vars,
         f"{prefix}_stored_gconsts": am.stored_gconsts}
    d.update({f"{prefix}_{k}": v for k, v in tree_arrays(tree).items()})
    return d


def main():
    import tempfile
    import torch
    import torchaudio.compliance.kaldi as TK
    tmp = tempfile.mkdtemp()
    for n in ("mono_model", "acoustic_g2p_output_model"):
        zipfile.ZipFile(f"{REF}/am/{n}.zip").extractall(tmp)
    out = {}
    tm, am = K.read_gmm_model(f"{tmp}/mono_model/final.mdl")
    out.update(model_arrays("mono", tm, am, K.read_tree(f"{tmp}/mono_model/tree")))
    tm2, am2 = K.read_gmm_model(f"{tmp}/acoustic_g2p_output_model/final.mdl")
    out.update(model_arrays("g2p", tm2, am2, K.read_tree(f"{tmp}/acoustic_g2p_output_model/tree")))
    out["g2p_lda"] = K.read_matrix_file(f"{tmp}/acoustic_g2p_output_model/lda.mat")
    import yaml
    meta = yaml.safe_load(open(f"{tmp}/mono_model/meta.yaml"))
    out["mono_phones"] = np.frombuffer(json.dumps(meta["phones"]).encode(), dtype=np.uint8)
    # sample corpus: PCM + transcript + dictionary text
    pcm, sr = K.read_wav_int16(f"{REF}/wav/acoustic_corpus.wav")
    assert sr == 16000
    out["acoustic_corpus_pcm"] = pcm
    pcm2, _ = K.read_wav_int16(f"{REF}/wav/cold_corpus.wav")
    out["cold_corpus_pcm"] = pcm2[: 16000 * 10]
    out["acoustic_corpus_lab"] = np.frombuffer(open(f"{REF}/lab/acoustic_corpus.lab", "rb").read(), dtype=np.uint8)
    out["test_acoustic_dict"] = np.frombuffer(open(f"{REF}/dictionaries/test_acoustic.txt", "rb").read(), dtype=np.uint8)
    # independent MFCC reference (torchaudio's port of Kaldi's compute-mfcc-feats), dither=0
    w = torch.from_numpy(pcm.astype(np.float32))[None]
    kw = dict(dither=0.0, energy_floor=0.0, sample_frequency=16000.0, num_mel_bins=23, num_ceps=13, low_freq=20.0, high_freq=7800.0)
    out["ta_mfcc_snip"] = TK.mfcc(w, use_energy=False, snip_edges=True, **kw).numpy()
    out["ta_mfcc_nosnip"] = TK.mfcc(w, use_energy=False, snip_edges=False, **kw).numpy()
    kw["energy_floor"] = 1.0
    out["ta_mfcc_energy"] = TK.mfcc(w, use_energy=True, snip_edges=True, **kw).numpy()
    # reference TextGrids of the sample corpus (short format, words + phones tiers; made upstream with a different English
    # model: a loose sanity bound for the exported boundaries, and a parser fixture) and one long-format file
    for key, name in (("acoustic_corpus_textgrid", "acoustic_corpus"), ("long_textgrid_fixture", "michaelandsickmichael")):
        out[key] = np.frombuffer(open(f"{REF}/textgrid/{name}.TextGrid", "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "reference_fixtures.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_fixtures.npz"), os.path.getsize(os.path.join(OUT, "reference_fixtures.npz")))


if __name__ == "__main__":
    main()
