"""Host-side format readers/writers against the reference's own Kaldi-format fixtures."""
import io
import os
import zipfile

import numpy as np
import pytest

from mfa_b200 import kaldi_io as K
from helpers import load_model

REF = "/root/reference/tests/data"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures not mounted")


@pytest.fixture(scope="module")
def ref_models(tmp_path_factory):
    d = tmp_path_factory.mktemp("ref")
    for n in ("mono_model", "acoustic_g2p_output_model"):
        zipfile.ZipFile(f"{REF}/am/{n}.zip").extractall(d)
    return str(d)


@needs_ref
def test_parse_fixture_models(ref_models):
    tm, am = K.read_gmm_model(f"{ref_models}/mono_model/final.mdl")
    assert (am.dim, am.NumPdfs(), am.NumGauss(), tm.num_tids, tm.tuples.shape[0]) == (39, 132, 132, 1206, 543)
    tm2, am2 = K.read_gmm_model(f"{ref_models}/acoustic_g2p_output_model/final.mdl")
    assert (am2.dim, am2.NumPdfs(), am2.NumGauss(), tm2.num_tids, tm2.tuples.shape[0]) == (40, 80, 500, 504, 244)
    # closed-form check: recomputed gconsts reproduce the stored <GCONSTS>
    assert np.abs(am.gconsts - am.stored_gconsts).max() < 1e-3
    assert np.abs(am2.gconsts - am2.stored_gconsts).max() < 1e-3
    # per transition-state probabilities sum to one
    for t in (tm, tm2):
        for ts in range(1, t.tuples.shape[0] + 1):
            a, b = t.state2id[ts], t.state2id[ts + 1]
            assert abs(np.exp(t.log_probs[a:b].astype(np.float64)).sum() - 1.0) < 1e-4
    lda = K.read_matrix_file(f"{ref_models}/acoustic_g2p_output_model/lda.mat")
    assert lda.shape == (40, 91)


@needs_ref
def test_tree_lookup_matches_tuples(ref_models):
    tm, _ = K.read_gmm_model(f"{ref_models}/mono_model/final.mdl")
    tree = K.read_tree(f"{ref_models}/mono_model/tree")
    assert (tree.N, tree.P) == (1, 0)
    assert [tree.lookup([1], c) for c in range(5)] == [10, 11, 12, 13, 14]
    for ph, hs, fpdf, _ in tm.tuples:
        pc = tm.topo.states_for(int(ph))[int(hs)].forward_pdf_class
        assert tree.lookup([int(ph)], pc) == fpdf
    t2 = K.read_tree(f"{ref_models}/acoustic_g2p_output_model/tree")
    assert (t2.N, t2.P) == (3, 1)


@needs_ref
def test_writers_roundtrip(ref_models, tmp_path):
    src = f"{ref_models}/acoustic_g2p_output_model"
    tm, am = K.read_gmm_model(f"{src}/final.mdl")
    K.write_gmm_model(tmp_path / "rt.mdl", tm, am)
    tm2, am2 = K.read_gmm_model(tmp_path / "rt.mdl")
    assert np.array_equal(tm.tuples, tm2.tuples) and np.array_equal(tm.log_probs, tm2.log_probs)
    assert np.array_equal(am.means_invvars, am2.means_invvars) and np.array_equal(am.inv_vars, am2.inv_vars)
    tree = K.read_tree(f"{src}/tree")
    K.write_tree(tmp_path / "rt.tree", tree)
    assert open(tmp_path / "rt.tree", "rb").read() == open(f"{src}/tree", "rb").read()
    with open(f"{src}/english_us_mfa.fst", "rb") as f:
        raw = f.read()
    fst = K.read_fst(io.BytesIO(raw))
    assert (fst.num_states, fst.arc_src.shape[0]) == (8478, 18602)
    buf = io.BytesIO()
    K.write_fst(buf, fst)
    fst2 = K.read_fst(io.BytesIO(buf.getvalue()))
    assert np.array_equal(fst.arc_dst, fst2.arc_dst) and np.array_equal(fst.arc_weight, fst2.arc_weight) and len(buf.getvalue()) == len(raw)


@needs_ref
def test_golden_is_derived_from_reference(ref_models):
    tm, am = K.read_gmm_model(f"{ref_models}/mono_model/final.mdl")
    tmg, amg, _ = load_model("mono")
    assert np.array_equal(tm.tuples, tmg.tuples) and np.array_equal(am.means_invvars, amg.means_invvars)
    assert np.array_equal(tm.tid2pdf, tmg.tid2pdf)


def test_compressed_matrix_codec():
    rng = np.random.default_rng(0)
    m = (rng.standard_normal((300, 13)) * np.linspace(1, 30, 13)).astype(np.float32)
    blob = K.compress_matrix(m)
    d = K.decompress_matrix(blob)
    assert d.shape == m.shape
    # one-byte codec: error bounded by the widest quantisation step of each column
    rngc = m.max(0) - m.min(0)
    assert np.all(np.abs(d - m).max(0) <= rngc / 64.0)
    # idempotence: re-compressing the decoded matrix is (nearly) a fixed point
    d2 = K.decompress_matrix(K.compress_matrix(d))
    assert np.abs(d2 - d).max() <= (rngc / 128.0).max()
    small = rng.standard_normal((4, 5)).astype(np.float32)  # rows <= 8 -> two-byte format
    ds = K.decompress_matrix(K.compress_matrix(small))
    assert np.abs(ds - small).max() < (small.max() - small.min()) / 60000.0


def test_ark_scp_tables(tmp_path):
    rng = np.random.default_rng(1)
    ark, scp = tmp_path / "x.ark", tmp_path / "x.scp"
    objs = {f"spk1-utt{i}": rng.integers(1, 500, size=50 + i).astype(np.int32) for i in range(5)}
    with K.ArkWriter(ark, scp) as w:
        for k, v in objs.items():
            w.write_int_vector(k, v)
    got = dict(K.read_ark(ark, "int_vector"))
    assert list(got) == list(objs) and all(np.array_equal(got[k], objs[k]) for k in objs)
    for key, path, off in K.read_scp(scp):
        assert np.array_equal(K.read_scp_object(path, off, "int_vector"), objs[key])
    ark2 = tmp_path / "m.ark"
    mats = {"a": rng.standard_normal((20, 13)).astype(np.float32), "b": rng.standard_normal((3, 14))}
    with K.ArkWriter(ark2) as w:
        w.write_matrix("a", mats["a"])
        w.write_matrix("b", mats["b"])
        w.write_matrix("c", mats["a"], compress=True)
        w.write_vector("v", mats["a"][0])
    it = K.read_ark(ark2, "matrix")
    k, m = next(it)
    assert k == "a" and np.array_equal(m, mats["a"])
    k, m = next(it)
    assert k == "b" and m.dtype == np.float64 and np.array_equal(m, mats["b"])
    k, m = next(it)
    assert k == "c" and np.abs(m - mats["a"]).max() < 0.1
