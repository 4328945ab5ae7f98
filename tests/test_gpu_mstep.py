"""N3 / a10: the device M-step (csrc/mstep.cu) against the numpy float64 restatement (oracle/mstep_oracle.py), the on-device rebuild of
the K2 operand images, the transition re-estimation and its re-folding into packed graphs, and a whole EM loop that never moves an
accumulator to the host."""
import numpy as np
import pytest

from helpers import build_synth_scenario, load_model
from mfa_b200 import engine as E, gmm_update as GU
from oracle import mstep_oracle as MO, oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


def _random_accs(am, rng, scale=60.0, starve=()):
    """Accumulators of a plausible E-step: occupancies, first / second order statistics around perturbed means."""
    G, D = am.NumGauss(), am.dim
    acc = GU.AccumAmDiagGmm.init(am)
    acc.occ = rng.uniform(0.2, 1.0, G) * scale
    for p in starve:
        acc.occ[am.offsets[p]:am.offsets[p + 1]] = rng.uniform(0.1, 2.0, am.offsets[p + 1] - am.offsets[p])
    mu = am.means() + 0.1 * rng.standard_normal((G, D)) * np.sqrt(am.variances())
    var = am.variances() * rng.uniform(0.5, 1.5, (G, D))
    var[::7, 0] = 1e-5     # exercises the variance floor
    acc.mean = acc.occ[:, None] * mu
    acc.var = acc.occ[:, None] * (var + mu * mu)
    acc.tot_like, acc.tot_frames = -1234.5, float(acc.occ.sum())
    return acc


def _assert_ll_close(a, b, rtol, atol, what=""):
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    i = np.unravel_index(int(np.argmax(err)), err.shape)
    assert err[i] <= 0, (what, i, float(a[i]), float(b[i]))


def _assert_models_close(got, ref, rtol=1e-6):
    assert np.array_equal(got.offsets, ref.offsets)
    assert np.allclose(got.weights, ref.weights, rtol=rtol, atol=1e-9)
    assert np.allclose(got.inv_vars, ref.inv_vars, rtol=rtol)
    assert np.allclose(got.means_invvars, ref.means_invvars, rtol=rtol, atol=1e-6)
    assert np.allclose(got.gconsts, ref.gconsts, rtol=rtol, atol=1e-5)


@pytest.mark.parametrize("which,min_occ", [("g2p", 10.0), ("mono", 10.0), ("g2p", 40.0)])
def test_device_mle_update_equals_numpy_restatement(eng, which, min_occ):
    tm, am, _ = load_model(which)
    rng = np.random.default_rng(3)
    P = am.NumPdfs()
    acc = _random_accs(am, rng, starve=(1, P // 2))          # two pdfs where every component is under the occupancy bar
    ref, impr_ref, cnt_ref = MO.mle_update(am, acc, mixup=0, min_gaussian_occupancy=min_occ)
    dm = E.DeviceModel(eng, tm, am)
    dm.acc_zero()
    dm.acc_write(acc.occ, acc.mean, acc.var, like=acc.tot_like, frames=acc.tot_frames)
    r = dm.mle_update(min_gaussian_occupancy=min_occ)
    got = dm.read()
    _assert_models_close(got, ref)
    assert np.allclose(got.device_gconsts, got.gconsts, rtol=1e-6, atol=1e-5)     # the kernel's gconsts == ComputeGconsts of its own rows
    assert r["num_gauss_after"] == ref.NumGauss() and r["num_removed"] == am.NumGauss() - ref.NumGauss()
    assert abs(r["gmm_count"] - cnt_ref) <= 1e-9 * cnt_ref and r["tot_frames"] == acc.tot_frames
    if ref.NumGauss() == am.NumGauss():
        assert abs(r["gmm_objf_impr"] - impr_ref) <= 1e-6 * abs(impr_ref) + 1e-3
    assert r["variance_floored"] > 0
    if which == "g2p" and min_occ == 10.0:
        # no component removed anywhere: Kaldi's objective change is defined for the whole model
        acc2 = _random_accs(am, np.random.default_rng(4))
        ref2, impr2, _ = MO.mle_update(am, acc2, min_gaussian_occupancy=5.0)
        assert ref2.NumGauss() == am.NumGauss()
        dm2 = E.DeviceModel(eng, tm, am)
        dm2.acc_zero()
        dm2.acc_write(acc2.occ, acc2.mean, acc2.var)
        r2 = dm2.mle_update(min_gaussian_occupancy=5.0)
        assert r2["layout_changed"] == 0 and abs(r2["gmm_objf_impr"] - impr2) <= 1e-5 * abs(impr2)
        _assert_models_close(dm2.read(), ref2)
        dm2.close()
    # the operand images rebuilt on the device score like a model created from the same parameters
    x = (ref.means()[rng.integers(0, ref.NumGauss(), 300)] + rng.standard_normal((300, am.dim))).astype(np.float32)
    fresh = E.DeviceModel(eng, tm, got)
    for impl in (0, 1):
        a, b = dm.loglikes(x, impl=impl), fresh.loglikes(x, impl=impl)
        assert np.allclose(a, b, rtol=2e-6, atol=2e-4), impl
    # against the oracle with the exact-order fp32 kernel: the accumulators above floor some variances at 1e-3 around means of
    # magnitude ~5, where the split-fp16 tensor-core path is only accurate relative to its (cancelling) terms -- K2's own precision is
    # tested in test_gpu_parity.py / test_gpu_config2.py, here the subject is the updated model
    _assert_ll_close(dm.loglikes(x, impl=1), O.gmm_loglikes(O.GmmModel.from_am(ref), x), 1e-4, 0.02, "ffma")
    fresh.close(); dm.close()


def test_device_mixup_split_targets_and_split_structure(eng):
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(5)
    acc = _random_accs(am, rng, scale=400.0)
    target = am.NumGauss() + 137
    dm = E.DeviceModel(eng, tm, am)
    dm.acc_zero()
    dm.acc_write(acc.occ, acc.mean, acc.var)
    r = dm.mle_update(mixup=target, power=0.25, min_count=20.0, seed=7)
    got = dm.read()
    base, _, _ = MO.mle_update(am, acc, mixup=0)
    state_occs = np.add.reduceat(acc.occ, am.offsets[:-1].astype(np.int64))
    tg = MO.get_split_targets(state_occs, target, 0.25, 20.0)
    want = np.maximum(np.diff(base.offsets), tg)             # SplitByCount only ever adds components
    assert np.array_equal(np.diff(got.offsets), want)
    assert r["num_split"] == int((want - np.diff(base.offsets)).sum()) and r["layout_changed"] == 1
    for p in range(am.NumPdfs()):
        a, b = got.offsets[p], got.offsets[p + 1]
        assert abs(float(got.weights[a:b].sum()) - 1.0) < 1e-5
        n0 = base.offsets[p + 1] - base.offsets[p]
        if b - a > n0:
            # mass is conserved by halving; the pdf's mean (sum w mu) is unchanged by the +-r perturbation of each split pair
            m_got = (got.weights[a:b, None] * got.means()[a:b]).sum(0)
            m_ref = (base.weights[base.offsets[p]:base.offsets[p + 1], None] * base.means()[base.offsets[p]:base.offsets[p + 1]]).sum(0)
            assert np.allclose(m_got, m_ref, rtol=1e-4, atol=1e-4)
            bv = base.variances()[base.offsets[p]:base.offsets[p + 1]]
            for row in got.variances()[a + n0:b]:     # a new component copies the variances of the one it was split from
                assert np.min(np.max(np.abs(bv - row) / bv, axis=1)) < 1e-5
    x = (got.means()[rng.integers(0, got.NumGauss(), 200)]).astype(np.float32)
    ref_ll = O.gmm_loglikes(O.GmmModel.from_am(got), x)
    # (variances floored at 1e-3 around means of magnitude ~5: fp32 itself cancels ~25 000-sized terms here, oracle and kernels alike)
    _assert_ll_close(dm.loglikes(x, impl=1), ref_ll, 1e-4, 0.02, "ffma")
    _assert_ll_close(dm.loglikes(x, impl=0), ref_ll, 1e-4, 0.1, "tc")
    dm.close()


def test_device_transition_update_and_graph_refold(eng):
    sc = build_synth_scenario(seconds=40.0, seed=11, triphone=True, n_phones=10, n_words=40, target_pdfs=80, gauss_per_pdf=2)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    dm = E.DeviceModel(eng, tm, am)
    feats = np.concatenate(sc["feats"]).astype(np.float32)
    res0 = E.align_feats(eng, dm, graphs, feats, sc["frame_off"], E.align_opts())
    ok = res0.status < 2
    assert ok.sum() >= 0.8 * len(ok)
    dm.acc_zero()
    dm.acc_stats(feats, res0.ali)
    trans = dm.acc_read()["trans"]
    r = dm.mle_update(update_transitions=True)
    _, lp = dm.read(with_transitions=True)
    import copy
    tm_ref = copy.deepcopy(tm)
    impr, cnt = tm_ref.mle_update(trans.copy())
    assert np.allclose(lp[1:], tm_ref.log_probs[1:], rtol=1e-6, atol=1e-7)
    assert abs(r["trans_objf_impr"] - impr) <= 1e-6 * abs(impr) + 1e-6 and abs(r["trans_count"] - cnt) < 1e-6
    # re-folding on the device == packing the graphs again with the new transition model
    graphs.set_transitions(eng, dm, 1.0, 0.1)
    res_dev = E.align_feats(eng, dm, graphs, feats, sc["frame_off"], E.align_opts())
    graphs2 = E.Graphs(batch, tm_ref, 1.0, 0.1)
    res_ref = E.align_feats(eng, dm, graphs2, feats, sc["frame_off"], E.align_opts())
    assert np.array_equal(res_dev.ali, res_ref.ali) and np.array_equal(res_dev.status, res_ref.status)
    assert np.allclose(res_dev.total_like, res_ref.total_like, rtol=1e-6)
    dm.close()


def test_em_loop_stays_on_device_and_likelihood_rises(eng):
    """align -> K4 -> device M-step (with mix-up back to the starting size) x 4, nothing read back but the result scalars; the
    average log-likelihood per frame of the accumulation passes must not decrease, and the final model must agree with the oracle."""
    sc = build_synth_scenario(seconds=120.0, seed=13, triphone=True, n_phones=12, n_words=60, target_pdfs=150, gauss_per_pdf=4)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    graphs = E.Graphs(batch, tm, 1.0, 0.1)
    dm = E.DeviceModel(eng, tm, am)
    feats = np.concatenate(sc["feats"]).astype(np.float32)
    avg = []
    for it in range(4):
        res = E.align_feats(eng, dm, graphs, feats, sc["frame_off"], E.align_opts())
        dm.acc_zero()
        dm.acc_stats(feats, res.ali)
        r = dm.mle_update(mixup=am.NumGauss(), update_transitions=True, seed=100 + it)
        graphs.set_transitions(eng, dm, 1.0, 0.1)
        avg.append(r["tot_like"] / r["tot_frames"])
        # SplitByCount never takes components away, so the total may exceed the target by the surplus of pdfs above their share
        assert r["num_gauss_after"] <= am.NumGauss() + am.NumPdfs() and r["tot_frames"] > 0
    assert all(b >= a - 1e-3 for a, b in zip(avg, avg[1:])), avg
    assert avg[-1] > avg[0]
    final, lp = dm.read(with_transitions=True)
    x = feats[::37][:256]
    assert np.allclose(dm.loglikes(x), O.gmm_loglikes(O.GmmModel.from_am(final), x), rtol=1e-4, atol=1e-3)
    dm.close()


def test_host_accumulator_entry_point_runs_on_the_device(eng):
    """gmm_update.mle_update (what the file-based acc_stats / mono_align_equal call with the jobs' summed accumulators)."""
    tm, am, _ = load_model("g2p")
    rng = np.random.default_rng(8)
    acc = _random_accs(am, rng)
    trans = rng.integers(0, 200, tm.num_tids + 1).astype(np.float64)
    trans[0] = 0
    import copy
    tm2 = copy.deepcopy(tm)
    l0 = eng.launch_count
    new, impr, cnt = GU.mle_update(am, acc, tm=tm2, transition_accs=trans, engine=eng)
    assert eng.launch_count - l0 >= 4
    ref, impr_ref, _ = MO.mle_update(am, acc)
    _assert_models_close(new, ref)
    tm3 = copy.deepcopy(tm)
    tm3.mle_update(trans.copy())
    assert np.allclose(tm2.log_probs[1:], tm3.log_probs[1:], rtol=1e-6, atol=1e-7)


def test_reserved_capacity_changes_nothing_but_the_allocations(eng):
    """mfa_model_reserve (training loops call it once so that no iteration allocates device memory): two EM iterations with and without
    the reservation end in bit-identical models."""
    sc = build_synth_scenario(seconds=60.0, seed=17, triphone=True, n_phones=10, n_words=50, target_pdfs=100, gauss_per_pdf=3)
    tm, am, c = sc["tm"], sc["am"], sc["corpus"]
    batch = E.GraphCompiler(tm, sc["tree"], c.lexicon).compile(c.transcripts)
    feats = np.concatenate(sc["feats"]).astype(np.float32)
    finals = []
    for reserve in (False, True):
        graphs = E.Graphs(batch, tm, 1.0, 0.1)
        dm = E.DeviceModel(eng, tm, am)
        if reserve:
            dm.reserve(2 * am.NumGauss() + 1)
            with pytest.raises(Exception):
                dm.reserve(am.NumGauss() - 1)          # below the current size: refused
        for it in range(2):
            res = E.align_feats(eng, dm, graphs, feats, sc["frame_off"], E.align_opts())
            dm.acc_zero()
            dm.acc_stats(feats, res.ali)
            dm.mle_update(mixup=am.NumGauss(), update_transitions=True, seed=7 + it)
            graphs.set_transitions(eng, dm, 1.0, 0.1)
        finals.append(dm.read(with_transitions=True))
        dm.close(); graphs.close()
    (a, lpa), (b, lpb) = finals
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.weights, b.weights) and np.array_equal(a.means_invvars, b.means_invvars)
    assert np.array_equal(a.inv_vars, b.inv_vars) and np.array_equal(lpa, lpb)
