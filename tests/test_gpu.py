"""Parity of the CUDA path (through the C ABI / the drop-in classes) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_torch
from oracle.ref_import import FakeTrajectory
from tests import _cases as C

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None
BASE = ref_torch.DIPEPTIDE_NM * 10.0


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def _pp_pair(c):
    """(pp_layer for the drop-in, closed-form oracle pre-processing) of a golden case."""
    from colvarsfinder import utils
    if c["pp_kind"] == "identity":
        return torch.nn.Identity(), cf.Preproc(identity=True)
    al = utils.Align(c["ref"], c["align_idx"]) if c["align_idx"] is not None else None
    fm = utils.FeatureMap(c["features"]) if c["features"] is not None else None
    return utils.Preprocessing(al, fm), cf.Preproc(align_idx=c["align_idx"], ref=c["ref"], feats=c["features"])


def _eigen_task(c, tmp, X=None, w=None, **kw):
    from colvarsfinder import core, nn
    model = nn.EigenFunctions(c["layer_dims"], c["k"])
    with torch.no_grad():
        for i in range(c["k"]):
            for p, v in zip(model.eigen_funcs[i].parameters(), c["params"][i]):
                p.copy_(torch.as_tensor(v))
    X = c["X"] if X is None else X
    w = c["w"] if w is None else w
    traj = FakeTrajectory(X, w.astype(np.float64), dt=1.0)
    pp, _ = _pp_pair(c)
    diag = None if c["diag_coeff"] is None else torch.as_tensor(c["diag_coeff"])
    task = core.EigenFunctionTask(traj, pp, model, str(tmp), c["alpha"], c["eig_w"], diag_coeff=diag, beta=c["beta"],
                                  sort_eigvals_in_training=c["sort"], k=c["k"], device=DEV, verbose=False, debug_mode=False,
                                  **kw)
    return task, model


# --------------------------------------------------------------------------------------------- alignment
@pytest.mark.parametrize("B", [1, 5, 128, 1000, 4099])
def test_align_fwd_matches_oracle(B):
    from colvarsfinder import utils
    X = ref_torch.synth_frames(BASE, B, seed=B)
    al = utils.Align(BASE, list(range(22))).to(DEV)
    y, R, c = al.rotation(torch.as_tensor(X, device=DEV))
    yo, Ro, co, _, _ = cf.kabsch(X.astype(np.float64), list(range(22)), BASE)
    assert np.abs(y.cpu().numpy() - yo).max() < 1e-6          # Angstrom (north_star: aligned coordinates to 1e-6 A)
    assert np.abs(R.cpu().numpy() - Ro).max() < 2e-7
    assert np.abs(c.cpu().numpy() - co[:, 0]).max() < 2e-6
    np.testing.assert_allclose(torch.linalg.det(R.double()).cpu().numpy(), 1.0, atol=1e-6)
    assert torch.equal(al(torch.as_tensor(X, device=DEV)), y)


def test_align_subset_large_molecule_and_invariance():
    from colvarsfinder import utils
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    X = ref_torch.synth_frames(BASE, 777, seed=2)
    y = utils.Align(BASE[heavy], heavy).to(DEV)(torch.as_tensor(X, device=DEV)).cpu().numpy()
    yo = cf.kabsch(X.astype(np.float64), heavy, BASE[heavy])[0]
    assert np.abs(y - yo).max() < 1e-6
    # 300-atom chain: a tile does not fit, warp-per-frame kernel
    chain = ref_torch.chain_structure(300, seed=4)
    Xc = ref_torch.synth_frames(chain, 257, seed=9)
    sel = list(range(0, 300, 7))
    al = utils.Align(chain[sel], sel).to(DEV)
    yc = al(torch.as_tensor(Xc, device=DEV)).cpu().numpy()
    yco = cf.kabsch(Xc.astype(np.float64), sel, chain[sel])[0]
    assert np.abs(yc - yco).max() < 5e-6                       # |y| up to ~40 A: float32 spacing is 4e-6 there
    # a rigid motion of the input does not change the aligned frame
    Q = np.linalg.qr(np.random.default_rng(0).normal(size=(3, 3)))[0]
    Q *= np.sign(np.linalg.det(Q))
    X2 = (X.astype(np.float64) @ Q + 3.0).astype(np.float32)
    y2 = utils.Align(BASE[heavy], heavy).to(DEV)(torch.as_tensor(X2, device=DEV)).cpu().numpy()
    assert np.abs(y2 - y).max() < 2e-5


def test_features_fwd_matches_oracle():
    from colvarsfinder import utils
    c = C.eigen_case("eigen_dipep_features")
    pp, ppo = _pp_pair(c)
    X = ref_torch.synth_frames(BASE, 1500, seed=12)
    r = pp.to(DEV)(torch.as_tensor(X, device=DEV)).cpu().numpy()
    ro = ppo.prepare(X.astype(np.float64))["r"]
    assert r.shape == ro.shape
    assert np.abs(r - ro).max() < 2e-5        # fp32 feature arithmetic on values up to ~10 (float32 eps * 10 * a few ops)
    c = C.eigen_case("eigen_dipep_invariant")
    pp, ppo = _pp_pair(c)
    r = pp.to(DEV)(torch.as_tensor(X, device=DEV)).cpu().numpy()
    assert np.abs(r - ppo.prepare(X.astype(np.float64))["r"]).max() < 2e-5



# --------------------------------------------------------------------------------------------- envelopes
# pass when |ours - gold64| <= max(1e-5 |gold64|, 1 x |ref32 - gold64|): gold64 = fp64 closed form, ref32 = the reference's own
# fp32 arithmetic (oracle/ref_torch.py: autograd through torch.linalg.svd, double backward) on the same fp32 inputs.  The
# achieved errors at BASELINE sizes are in profiles/r02_parity_errors.json (loss ~1e-8, eigenvalues ~6e-8, gradients ~1e-5).
def _ref32(X, w, nets, ppo, alpha, eig_w, diag=None, beta=1.0, sort=True):
    """(loss, eig, grads) of the reference's fp32 path; None when the batch is too large to be worth a CPU autograd run."""
    if len(X) > 25000:
        return None
    if ppo.identity:
        ppt = ref_torch.Preprocess()
    else:
        al = None if ppo.align_idx is None else ref_torch.Align(ppo.ref, ppo.align_idx)
        ppt = ref_torch.Preprocess(al, None if ppo.feats is None else ref_torch.FeatureMap(ppo.feats))
    tn = [[torch.as_tensor(p).float().requires_grad_() for p in n] for n in nets]
    Xt = torch.as_tensor(X).float().requires_grad_()
    a = None if diag is None else torch.as_tensor(diag).float()
    out = ref_torch.eigen_loss(Xt, torch.as_tensor(w).float(), tn, ppt, alpha, eig_w, a, beta, sort)
    out[0].backward()
    return float(out[0]), out[1].numpy().astype(np.float64), [[np.zeros(tuple(p.shape)) if p.grad is None else p.grad.numpy() for p in n] for n in tn]


def _assert_within_envelope(out, grads, comb, g64, ref32, tag=""):
    loss, eig, obj, pen, cvec = out
    bad = []
    if list(cvec.cpu().numpy()) != list(comb["cvec"]):
        # an ordering flip is legitimate only between eigenvalues closer than the fp32 floor
        bad.append(("cvec", list(cvec.cpu().numpy()), list(comb["cvec"])))
    l32 = ref32[0] if ref32 else comb["loss"]
    if not abs(float(loss) - comb["loss"]) <= max(1e-5 * abs(comb["loss"]), abs(l32 - comb["loss"])):
        bad.append(("loss", float(loss), comb["loss"], l32))
    e = eig.cpu().numpy().astype(np.float64)
    for i in range(len(e)):
        e32 = ref32[1][i] if ref32 else comb["eig"][i]
        if not abs(e[i] - comb["eig"][i]) <= max(1e-5 * abs(comb["eig"][i]), abs(e32 - comb["eig"][i])):
            bad.append(("eig", i, e[i], comb["eig"][i], e32))
    scale = max(np.abs(t).max() for n in g64 for t in n)
    for i in range(len(g64)):
        for j in range(len(g64[i])):
            if np.abs(g64[i][j]).max() < 1e-9 * scale:     # last-layer bias: zero up to rounding
                if not np.abs(grads[i][j]).max() < 1e-5 * scale:
                    bad.append(("zero-grad", i, j, float(np.abs(grads[i][j]).max())))
                continue
            floor = C.rel_l2(ref32[2][i][j], g64[i][j]) if ref32 else 0.0
            err = C.rel_l2(grads[i][j], g64[i][j])
            if not err <= max(2e-5, floor):
                bad.append(("grad", i, j, err, floor))
    assert not bad, (tag, bad)


# --------------------------------------------------------------------------------------------- eigenfunction loss
def _check_eigen(c, out, grads, gold, tag=""):
    loss, eig, obj, pen, cvec = out
    assert list(cvec.cpu().numpy()) == list(gold["cvec"])
    assert abs(float(loss) - gold["loss"]) <= C.tol(gold["loss"], gold.get("loss32", gold["loss"])), (tag, float(loss), gold["loss"])
    assert abs(float(obj) - gold["obj"]) <= C.tol(gold["obj"], gold.get("obj32", gold["obj"]))
    assert abs(float(pen) - gold["pen"]) <= C.tol(gold["pen"], gold.get("pen32", gold["pen"]), rel=2e-5)
    e32 = gold.get("eig32", gold["eig"])
    for i in range(len(gold["eig"])):
        assert abs(float(eig[i]) - gold["eig"][i]) <= C.tol(gold["eig"][i], e32[i]), (tag, i)
    for i in range(len(grads)):
        for j in range(len(grads[i])):
            g64 = gold["grads"][i][j]
            if np.abs(g64).max() < 1e-12:       # last-layer bias: exactly zero gradient in the generator loss
                assert np.abs(grads[i][j]).max() < 1e-4 * max(1.0, abs(gold["loss"]))
                continue
            ref_err = C.rel_l2(gold["grads32"][i][j], g64) if "grads32" in gold else 0.0
            assert C.rel_l2(grads[i][j], g64) <= max(2e-5, ref_err), (tag, i, j, C.rel_l2(grads[i][j], g64), ref_err)


@pytest.mark.parametrize("name", C.EIGEN_GENERATOR_CASES)
def test_eigen_loss_matches_reference_golden(name, tmp_path):
    """loss, eigenvalues, objective, penalty, cvec and every parameter gradient vs the reference (fp64 gold; the
    reference's own fp32 run sets the floor -- SURVEY 7.3-D)."""
    c = C.eigen_case(name)
    task, model = _eigen_task(c, tmp_path)
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    gold = dict(loss=float(c["g64_loss"]), obj=float(c["g64_obj"]), pen=float(c["g64_pen"]), eig=c["g64_eig"],
                cvec=c["g64_cvec"], grads=c["g64"], loss32=float(c["r32_loss"]), obj32=float(c["r32_obj"]),
                pen32=float(c["r32_pen"]), eig32=c["r32_eig"], grads32=c["g32"])
    _check_eigen(c, out, grads, gold, name)


@pytest.mark.parametrize("B", [1, 37, 128, 129, 1000, 20011])
@pytest.mark.parametrize("name", ["eigen_dipep_k3", "eigen_dipep_features", "eigen_2d_k3_diag"])
def test_eigen_loss_matches_oracle_ragged_sizes(name, B, tmp_path):
    """Same inputs through the CUDA path and the fp64 closed-form oracle at batch sizes that exercise the tail
    tile, a single frame, and many CTAs."""
    c = C.eigen_case(name)
    if c["pp_kind"] == "identity":
        X = ref_torch.ring_2d(B, 100 + B) if hasattr(ref_torch, "ring_2d") else None
        rng = np.random.default_rng(100 + B)
        X = rng.normal(size=(B, 2)).astype(np.float32)
    else:
        X = ref_torch.synth_frames(BASE, B, seed=100 + B)
    w = ref_torch.boltzmann_weights(B, seed=B)
    if B == 1:
        pytest.skip("variance of a single frame is zero: the reference loss is undefined (division by zero)")
    task, model = _eigen_task(c, tmp_path, X, w)
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    _, ppo = _pp_pair(c)
    comb, g64, _ = cf.eigen_loss_and_grads(X, w, c["params"], ppo, c["alpha"], c["eig_w"], c["diag_coeff"], c["beta"], c["sort"])
    ref32 = _ref32(X, w, c["params"], ppo, c["alpha"], c["eig_w"], c["diag_coeff"], c["beta"], c["sort"])
    _assert_within_envelope(out, grads, comb, g64, ref32, f"{name}/B={B}")


def _random_nets(dims, k, seed):
    torch.manual_seed(seed)
    return [[p.numpy() for p in ref_torch.init_mlp_params(dims)] for _ in range(k)]


FAST_CASES = {
    # name: (layer dims, k, pre-processing, diag)
    "c3_all_atoms": ([66, 20, 20, 20, 1], 3, "align_all", False),
    "align_subset_k2": ([66, 20, 20, 20, 1], 2, "align_subset", False),
    "two_hidden_k4": ([66, 20, 20, 1], 4, "align_all", False),
    "width16_k1": ([66, 16, 16, 16, 1], 1, "align_all", False),
    "width32_k2": ([5, 32, 32, 32, 1], 2, "identity", True),
    "identity_2d_k1": ([2, 20, 20, 20, 1], 1, "identity", False),
    "identity_5d_diag_k3": ([5, 20, 20, 20, 1], 3, "identity", True),
}


@pytest.mark.parametrize("B", [2, 37, 512, 1500])
@pytest.mark.parametrize("name", sorted(FAST_CASES))
def test_eigen_fast_path_matches_oracle(name, B, tmp_path):
    """The thread-private FFMA2 kernels (cvf_eigen_fast.cu) against the fp64 closed-form oracle: every network shape they are
    instantiated for, alignment on all atoms and on a subset, Identity pre-processing with and without diag_coeff, batch
    sizes below / at / above their 512-frame tile."""
    from colvarsfinder import core, nn, utils
    dims, k, ppk, use_diag = FAST_CASES[name]
    nets = _random_nets(dims, k, seed=len(name) + B)
    rng = np.random.default_rng(B + k)
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    if ppk == "identity":
        X = rng.normal(size=(B, dims[0])).astype(np.float32)
        pp, ppo = torch.nn.Identity(), cf.Preproc(identity=True)
    elif ppk == "align_all":
        X = ref_torch.synth_frames(BASE, B, seed=B + 1)
        pp, ppo = utils.Align(BASE, list(range(22))), cf.Preproc(align_idx=list(range(22)), ref=BASE)
    else:
        X = ref_torch.synth_frames(BASE, B, seed=B + 2)
        pp, ppo = utils.Align(BASE[heavy], heavy), cf.Preproc(align_idx=heavy, ref=BASE[heavy])
    w = ref_torch.boltzmann_weights(B, seed=B)
    diag = (0.5 + rng.random(dims[0])).astype(np.float32) if use_diag else None
    eig_w = [1.0, 0.6, 0.3, 0.2][:k]
    model = nn.EigenFunctions(dims, k)
    with torch.no_grad():
        for i in range(k):
            for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                p.copy_(torch.as_tensor(v))
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=1.0), pp, model, str(tmp_path), 20.0, eig_w,
                                  diag_coeff=None if diag is None else torch.as_tensor(diag), k=k, device=DEV, verbose=False,
                                  debug_mode=False)
    assert task._ctx.fast_path
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    comb, g64, _ = cf.eigen_loss_and_grads(X, w, nets, ppo, 20.0, eig_w, diag)
    _assert_within_envelope(out, grads, comb, g64, _ref32(X, w, nets, ppo, 20.0, eig_w, diag), f"{name}/B={B}")


@pytest.mark.parametrize("name", ["eigen_2d_k1", "eigen_dipep_k3"])
def test_eigen_general_kernels_on_fast_shapes(name, tmp_path):
    """cvf_eigen_set_path(1) forces the general row-engine kernels on the shapes the fast path normally takes: both
    implementations are held to the reference's golden vectors and to each other."""
    from colvarsfinder import _lib
    c = C.eigen_case(name)
    res = {}
    for mode in (1, 0):
        _lib.check(_lib.lib().cvf_eigen_set_path(mode), "cvf_eigen_set_path")
        try:
            task, model = _eigen_task(c, tmp_path)
            assert task._ctx.fast_path == (mode == 0)
            out = task.loss_func(task._traj, task._weights, None, None)
            out[0].backward()
            grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
        finally:
            _lib.lib().cvf_eigen_set_path(0)
        gold = dict(loss=float(c["g64_loss"]), obj=float(c["g64_obj"]), pen=float(c["g64_pen"]), eig=c["g64_eig"],
                    cvec=c["g64_cvec"], grads=c["g64"], loss32=float(c["r32_loss"]), obj32=float(c["r32_obj"]),
                    pen32=float(c["r32_pen"]), eig32=c["r32_eig"], grads32=c["g32"])
        _check_eigen(c, out, grads, gold, f"{name}/mode{mode}")
        res[mode] = (float(out[0]), grads)
    assert abs(res[0][0] - res[1][0]) <= 2e-5 * abs(res[1][0])


def test_eigen_invariant_features_with_alignment_matches_oracle(tmp_path):
    """C4-style pre-processing: Kabsch alignment followed by bond / angle / dihedral features only.  The step elides the
    alignment (the features are invariant under it); the oracle goes through the alignment and its Jacobian."""
    from colvarsfinder import core, nn, utils
    B, k, dims = 700, 2, [8, 16, 16, 1]
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    feats = [("bond", [1, 4]), ("bond", [4, 8]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]), ("dihedral", [6, 8, 14, 16]),
             ("bond", [10, 18])]
    X = ref_torch.synth_frames(BASE, B, seed=41)
    w = ref_torch.boltzmann_weights(B, seed=42)
    nets = _random_nets(dims, k, seed=43)
    pp = utils.Preprocessing(utils.Align(BASE[heavy], heavy), utils.FeatureMap(feats))
    ppo = cf.Preproc(align_idx=heavy, ref=BASE[heavy], feats=feats)
    model = nn.EigenFunctions(dims, k)
    with torch.no_grad():
        for i in range(k):
            for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                p.copy_(torch.as_tensor(v))
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=1.0), pp, model, str(tmp_path), 20.0, [1.0, 0.5], k=k,
                                  device=DEV, verbose=False, debug_mode=False)
    assert task._ctx.spec.alignment_elided
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    comb, g64, _ = cf.eigen_loss_and_grads(X, w, nets, ppo, 20.0, [1.0, 0.5])
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    _assert_within_envelope(out, grads, comb, g64, _ref32(X, w, nets, ppo, 20.0, [1.0, 0.5]), "invariant features after alignment")


FEATURE_FAST_CASES = {
    # name: (molecule, records, alignment atoms or None, diag_coeff, k)
    "dipep_invariant_aligned": ("dipeptide", [("bond", [1, 4]), ("bond", [4, 8]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]),
                                              ("dihedral", [6, 8, 14, 16]), ("bond", [10, 18]), ("angle", [14, 16, 18])],
                                [1, 4, 5, 6, 8, 10, 14, 15, 16, 18], False, 2),
    "dipep_positions_bonds_diag": ("dipeptide", [("position", [4]), ("bond", [4, 6]), ("dihedral", [4, 6, 8, 14]), ("position", [8]),
                                                 ("angle", [6, 8, 14]), ("bond", [1, 18])], None, True, 3),
    "c4_chain": ("chain166", None, list(range(0, 160, 4)), False, 3),
}


def _c4_records():
    sel = list(range(5, 166, 16))[:10]
    feats = [("bond", [a, b]) for i, a in enumerate(sel) for b in sel[i + 1:]]
    feats += [("dihedral", [s, s + 1, s + 2, s + 3]) for s in range(10, 10 + 18 * 8, 8)]
    return feats


@pytest.mark.parametrize("B", [2, 37, 512, 1500])
@pytest.mark.parametrize("name", sorted(FEATURE_FAST_CASES))
def test_eigen_fast_feature_path_matches_oracle(name, B, tmp_path):
    """Feature maps (position / bond / angle / dihedral records) on the thread-private kernels: features and gradient stencils
    from the raw frame (prep_feat), J diag(a) J^T applied per atom (jjt), then the common pass 2.  Checked against the fp64
    closed-form oracle (which goes through the alignment and its Jacobian where one is configured) and against the general
    row-engine kernels on the same inputs."""
    from colvarsfinder import _lib, core, nn, utils
    mol, feats, align_idx, use_diag, k = FEATURE_FAST_CASES[name]
    base = BASE if mol == "dipeptide" else ref_torch.chain_structure(166, seed=2026)
    if feats is None:
        feats = _c4_records()
    d_r = sum({"position": 3, "bond": 1, "angle": 1, "dihedral": 2}[t] for t, _ in feats)
    dims = [d_r, 20, 20, 20, 1]
    nets = _random_nets(dims, k, seed=len(name) + B)
    rng = np.random.default_rng(B + k)
    X = ref_torch.synth_frames(base, B, seed=B + 3)
    w = ref_torch.boltzmann_weights(B, seed=B)
    fmap = utils.FeatureMap(feats)
    if align_idx is None:
        pp, ppo = fmap, cf.Preproc(feats=feats)
    else:
        pp = utils.Preprocessing(utils.Align(base[align_idx], align_idx), fmap)
        ppo = cf.Preproc(align_idx=align_idx, ref=base[align_idx], feats=feats)
    diag = (0.5 + rng.random(3 * base.shape[0])).astype(np.float32) if use_diag else None
    eig_w = [1.0, 0.6, 0.3][:k]
    res = {}
    for mode in (0, 1):
        _lib.check(_lib.lib().cvf_eigen_set_path(mode), "cvf_eigen_set_path")
        try:
            model = nn.EigenFunctions(dims, k)
            with torch.no_grad():
                for i in range(k):
                    for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                        p.copy_(torch.as_tensor(v))
            task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=1.0), pp, model, str(tmp_path), 20.0, eig_w,
                                          diag_coeff=None if diag is None else torch.as_tensor(diag), k=k, device=DEV,
                                          verbose=False, debug_mode=False)
            assert task._ctx.fast_path == (mode == 0)
            out = task.loss_func(task._traj, task._weights, None, None)
            out[0].backward()
            res[mode] = (out, [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs])
        finally:
            _lib.lib().cvf_eigen_set_path(0)
    comb, g64, _ = cf.eigen_loss_and_grads(X, w, nets, ppo, 20.0, eig_w, diag)
    ref32 = _ref32(X, w, nets, ppo, 20.0, eig_w, diag)
    for mode in (0, 1):
        _assert_within_envelope(res[mode][0], res[mode][1], comb, g64, ref32, f"{name}/B={B}/mode{mode}")


def _feature_task(tmp_path, base, feats, align_idx, dims, k, X, w, diag=None, lag_tau=0, dt=1.0, eig_w=None):
    from colvarsfinder import core, nn, utils
    fmap = utils.FeatureMap(feats)
    pp = fmap if align_idx is None else utils.Preprocessing(utils.Align(base[align_idx], align_idx), fmap)
    nets = _random_nets(dims, k, seed=sum(dims) + k)
    model = nn.EigenFunctions(dims, k)
    with torch.no_grad():
        for i in range(k):
            for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                p.copy_(torch.as_tensor(v))
    eig_w = eig_w or [1.0, 0.6, 0.3, 0.2, 0.15, 0.1, 0.05][:k]
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=dt), pp, model, str(tmp_path), 20.0, eig_w,
                                  diag_coeff=None if diag is None else torch.as_tensor(diag), lag_tau=lag_tau, k=k, device=DEV,
                                  verbose=False, debug_mode=False)
    return task, model, nets, eig_w


@pytest.mark.parametrize("dims_k", [([9, 16, 16, 16, 1], 1), ([9, 32, 32, 32, 1], 2), ([9, 20, 20, 1], 4), ([9, 20, 20, 20, 1], 7)])
def test_eigen_fast_feature_path_network_shapes(dims_k, tmp_path):
    """Every instantiated network shape of the thread-private kernels on a feature map, and every warps-per-network setting of
    the J diag(a) J^T kernel (k = 1..3: four warps, k = 4..6: two, k = 7, 8: one)."""
    dims, k = dims_k
    feats = [("bond", [1, 4]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]), ("position", [8]), ("dihedral", [6, 8, 14, 16])]
    B = 333
    X = ref_torch.synth_frames(BASE, B, seed=21)
    w = ref_torch.boltzmann_weights(B, seed=22)
    task, model, nets, eig_w = _feature_task(tmp_path, BASE, feats, None, dims, k, X, w)
    assert task._ctx.fast_path
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    ppo = cf.Preproc(feats=feats)
    comb, g64, _ = cf.eigen_loss_and_grads(X, w, nets, ppo, 20.0, eig_w)
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    _assert_within_envelope(out, grads, comb, g64, _ref32(X, w, nets, ppo, 20.0, eig_w), f"{dims}/k={k}")


def test_eigen_feature_path_determinism_and_additivity(tmp_path):
    """C4 records on 2^18 frames of the 166-atom chain: two runs give bit-identical losses and gradients (every sum of the
    feature kernels has one owner), and the fp64 batch sums of a batch equal the sum over two unequal parts."""
    base = ref_torch.chain_structure(166, seed=2026)
    n = 1 << 18
    X = ref_torch.synth_frames(base, n, seed=9)
    w = ref_torch.boltzmann_weights(n, seed=9)
    task, model, nets, eig_w = _feature_task(tmp_path, base, _c4_records(), list(range(0, 160, 4)), [81, 20, 20, 20, 1], 3, X, w)
    assert task._ctx.fast_path and task._ctx.spec.alignment_elided
    outs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out = task.loss_func(task._traj, task._weights)
        out[0].backward()
        outs.append((out[0].clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()))
    assert torch.isfinite(outs[0][0]) and torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ctx, Xd, wd = task._ctx, task._traj, task._weights
    _, s_all = ctx.stats(Xd, wd)
    h = n // 2 + 77
    _, s1 = ctx.stats(Xd[:h].contiguous(), wd[:h].contiguous())
    _, s2 = ctx.stats(Xd[h:].contiguous(), wd[h:].contiguous())
    torch.testing.assert_close(s1 + s2, s_all, rtol=1e-12, atol=0)
    # a slice against the oracle (which goes through the alignment and its Jacobian)
    ppo = cf.Preproc(align_idx=list(range(0, 160, 4)), ref=base[list(range(0, 160, 4))], feats=_c4_records())
    _, _, S = cf.eigen_loss_and_grads(X[:1500], w[:1500], nets, ppo, 20.0, eig_w)
    _, s_small = ctx.stats(Xd[:1500].contiguous(), wd[:1500].contiguous())
    np.testing.assert_allclose(s_small.cpu().numpy()[-3:], S["SD"], rtol=1e-4)


def _lag_ref(X, w, nets, ppt, alpha, eig_w, lag, lag_time, sort, dtype):
    tn = [[torch.tensor(p, dtype=dtype, requires_grad=True) for p in net] for net in nets]
    Xt, wt = torch.tensor(X, dtype=dtype), torch.tensor(w, dtype=dtype)
    pp = ppt.double() if dtype == torch.float64 else ppt.float()
    ref = ref_torch.eigen_loss(Xt[:-lag], wt[:-lag], tn, pp, alpha, eig_w, sort=sort, X_lagged=Xt[lag:], weight_lagged=wt[lag:],
                               lag_time=lag_time)
    ref[0].backward()
    grads = [[np.zeros(tuple(t.shape)) if t.grad is None else t.grad.numpy().astype(np.float64) for t in n] for n in tn]
    return dict(loss=float(ref[0]), eig=np.asarray(ref[1], dtype=np.float64), cvec=[int(v) for v in ref[4]]), grads


def _assert_lag_within_envelope(out, model, X, w, nets, ppt, alpha, eig_w, lag, lag_time, sort=True, tag=""):
    """Transfer-operator loss against the reference formula in fp64, with the reference's fp32 run as the floor."""
    gold, g64 = _lag_ref(X, w, nets, ppt, alpha, eig_w, lag, lag_time, sort, torch.float64)
    r32, g32 = _lag_ref(X, w, nets, ppt, alpha, eig_w, lag, lag_time, sort, torch.float32)
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    _assert_within_envelope(out, grads, gold, g64, (r32["loss"], r32["eig"], g32), tag)


def test_eigen_lag_loss_on_feature_path(tmp_path):
    """Transfer-operator branch with a feature map as pre-processing: both forward passes and both backward passes run on the
    feature kernels (the lagged batch uses the second scratch slot)."""
    feats = [("bond", [1, 4]), ("dihedral", [4, 6, 8, 14]), ("angle", [6, 8, 14]), ("dihedral", [6, 8, 14, 16]), ("bond", [10, 18])]
    lag, B, k, dims = 2, 600, 2, [7, 20, 20, 20, 1]
    X = ref_torch.synth_frames(BASE, B + lag, seed=51)
    w = ref_torch.boltzmann_weights(B + lag, seed=52)
    task, model, nets, eig_w = _feature_task(tmp_path, BASE, feats, None, dims, k, X, w, lag_tau=1.0, dt=0.5)
    assert task.lag_idx == lag and task._ctx.fast_path
    Xd, wd = task._traj, task._weights
    out = task.loss_func(Xd[:-lag].contiguous(), wd[:-lag].contiguous(), Xd[lag:].contiguous(), wd[lag:].contiguous())
    out[0].backward()
    _assert_lag_within_envelope(out, model, X, w, nets, ref_torch.Preprocess(None, ref_torch.FeatureMap(feats)), 20.0, eig_w, lag, 1.0,
                                tag="lag loss on the feature path")


def test_eigen_feature_descriptor_too_small_poisons_instead_of_overrunning(tmp_path):
    """The three sizing fields of cvf_preproc are a contract: with n_shared_atoms understated the table builder refuses the
    record list and the Dirichlet sums come out NaN -- no out-of-bounds write, no silent wrong answer."""
    feats = [("bond", [1, 4]), ("bond", [4, 8]), ("bond", [1, 8]), ("dihedral", [4, 6, 8, 14])]
    X = ref_torch.synth_frames(BASE, 200, seed=3)
    w = ref_torch.boltzmann_weights(200, seed=3)
    task, model, nets, eig_w = _feature_task(tmp_path, BASE, feats, None, [5, 20, 20, 20, 1], 1, X, w)
    assert task._ctx.fast_path and task._ctx.spec.struct.n_shared_atoms == 3
    good = task.loss_func(task._traj, task._weights)[0]
    assert torch.isfinite(good)
    task._ctx.spec.struct.n_shared_atoms = 1
    task._ctx._ws.clear()
    bad = task.loss_func(task._traj, task._weights)[0]
    assert torch.isnan(bad)


def test_eigen_batch_sums_are_additive_at_full_size(tmp_path):
    """Size-independent property at BASELINE scale (2^20 frames of C3): the fp64 batch sums of a batch equal the sum over
    its halves, and the gradient sums of pass 2 (at fixed coefficients) are additive too."""
    c = C.eigen_case("eigen_dipep_k3")
    n = 1 << 20
    gen = torch.Generator(device="cpu").manual_seed(5)
    base = torch.as_tensor(BASE, dtype=torch.float32)
    small = torch.as_tensor(ref_torch.synth_frames(BASE, 4096, seed=77))
    X = (small[torch.randint(0, 4096, (n,), generator=gen)] + 0.05 * torch.randn(n, 22, 3, generator=gen)).numpy()
    w = ref_torch.boltzmann_weights(n, seed=3)
    task, model = _eigen_task(c, tmp_path, X, w)
    ctx = task._ctx
    Xd, wd = task._traj, task._weights
    y, s_all = ctx.stats(Xd, wd)
    h = n // 2 + 13
    y1, s1 = ctx.stats(Xd[:h], wd[:h])
    y2, s2 = ctx.stats(Xd[h:], wd[h:])
    torch.testing.assert_close(s1 + s2, s_all, rtol=1e-12, atol=0)
    assert torch.equal(torch.cat([y1, y2], 1), y)
    comb = ctx.combine(s_all)
    g_all = ctx.grads(Xd, wd, y, comb)
    g_sum = ctx.grads(Xd[:h], wd[:h], y1.contiguous(), comb) + ctx.grads(Xd[h:], wd[h:], y2.contiguous(), comb)
    scale = g_all.abs().max()
    assert (g_all - g_sum).abs().max() <= 2e-6 * scale      # fp32 within-tile sums regroup when the tile boundaries move
    # sample check against the oracle on a slice
    comb_o, _, S = cf.eigen_loss_and_grads(X[:3000], w[:3000], c["params"], _pp_pair(c)[1], c["alpha"], c["eig_w"])
    _, s_small = ctx.stats(Xd[:3000].contiguous(), wd[:3000].contiguous())
    np.testing.assert_allclose(s_small.cpu().numpy()[0], S["S0"], rtol=1e-6)
    np.testing.assert_allclose(s_small.cpu().numpy()[-3:], S["SD"], rtol=1e-4)


def test_eigen_full_comparison_at_large_batch(tmp_path):
    """Loss, eigenvalues, ordering and EVERY gradient of C3 at 2^18 frames against the fp64 closed form (evaluated in chunks);
    the same comparison at 2^20 frames for C1-C4 is profiles/parity_report.py -> profiles/r02_parity_errors.json."""
    from colvarsfinder import core, nn, utils
    n = 1 << 18
    X = ref_torch.synth_frames(BASE, n, seed=2026)
    w = ref_torch.boltzmann_weights(n, seed=2026)
    dims, k, eig_w = [66, 20, 20, 20, 1], 3, [1.0, 0.6, 0.3]
    nets = _random_nets(dims, k, seed=2026)
    model = nn.EigenFunctions(dims, k)
    with torch.no_grad():
        for i in range(k):
            for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                p.copy_(torch.as_tensor(v))
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), utils.Align(BASE, list(range(22))), model, str(tmp_path),
                                  20.0, eig_w, k=k, device=DEV, verbose=False, debug_mode=False)
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    comb, g64 = cf.eigen_loss_and_grads_chunked(X, w, nets, cf.Preproc(align_idx=list(range(22)), ref=BASE), 20.0, eig_w)
    assert list(out[4].cpu().numpy()) == list(comb["cvec"])
    assert abs(float(out[0]) - comb["loss"]) <= 1e-6 * abs(comb["loss"])          # achieved: ~1e-8
    np.testing.assert_allclose(out[1].cpu().numpy(), comb["eig"], rtol=2e-6)      # achieved: ~6e-8
    scale = max(np.abs(t).max() for net in g64 for t in net)
    for i in range(k):
        for p, t in zip(model.eigen_funcs[i].parameters(), g64[i]):
            if np.abs(t).max() < 1e-9 * scale:
                continue
            assert C.rel_l2(p.grad.cpu().numpy(), t) < 5e-5                       # achieved: max 1.2e-5, median 3e-6


def test_eigen_determinism(tmp_path):
    c = C.eigen_case("eigen_dipep_k3")
    X = ref_torch.synth_frames(BASE, 30000, seed=1)
    w = ref_torch.boltzmann_weights(30000, seed=1)
    task, model = _eigen_task(c, tmp_path, X, w)
    outs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out = task.loss_func(task._traj, task._weights)
        out[0].backward()
        outs.append((out[0].clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_eigen_second_backward_recomputes_consumed_scratch(tmp_path):
    """Pass 2 consumes pass 1's intermediates (it overwrites grad_r y and drops those lines from L2): a second backward() of the
    same loss must notice and run pass 1 again.  Deterministic kernels: the accumulated gradient is exactly twice the first."""
    c = C.eigen_case("eigen_dipep_k3")
    X = ref_torch.synth_frames(BASE, 5000, seed=3)
    w = ref_torch.boltzmann_weights(5000, seed=3)
    task, model = _eigen_task(c, tmp_path, X, w)
    out = task.loss_func(task._traj, task._weights)
    out[0].backward(retain_graph=True)
    g1 = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    out[0].backward()
    g2 = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.isfinite(g2).all() and torch.equal(g2, 2.0 * g1)


def test_eigen_train_matches_reference_run(tmp_path):
    """Whole train(): same split (numpy global RNG drawn twice), same batches, Adam -- per-iteration losses and the final
    parameters follow the reference run stored in tests/golden/train_eigen_2d.npz."""
    from colvarsfinder import core, nn
    d = C.load("train_eigen_2d")
    torch.manual_seed(11)
    model = nn.EigenFunctions([2, 20, 20, 20, 1], 2)
    for j, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.as_tensor(d[f"init_{j}"]))      # same init stream as the reference
    traj = FakeTrajectory(d["X"].astype(np.float64), d["w"].astype(np.float64), dt=0.1)
    task = core.EigenFunctionTask(traj, torch.nn.Identity(), model, str(tmp_path), 20.0, [1.0, 0.7], beta=1.0, lag_tau=0,
                                  learning_rate=0.005, k=2, batch_size=300, num_epochs=3, test_ratio=0.2,
                                  save_model_every_step=0, device=DEV, verbose=False, debug_mode=False)
    np.random.seed(77)
    task.train()
    tr = np.stack([l[0].numpy() for l in task.loss_list])
    te = np.stack([l[1].numpy() for l in task.loss_list])
    assert tr.shape == d["train_hist"].shape and te.shape == d["test_hist"].shape
    np.testing.assert_allclose(tr, d["train_hist"], rtol=2e-3)
    np.testing.assert_allclose(te, d["test_hist"], rtol=2e-3)
    np.testing.assert_allclose(task.train_loss_df.to_numpy(), d["train_df"], rtol=2e-3)
    assert list(task.train_loss_df.columns) == ['loss', 'eigen_non_penalty', 'eigen_penalty', 'eig_1', 'eig_2']
    for j, (name, p) in enumerate(model.named_parameters()):
        if name.endswith("4.bias"):
            # the last-layer bias has an exactly-zero gradient in the generator loss (it cancels from var, cov and
            # grad f); Adam turns the rounding noise of either implementation into +-lr steps, so it is not comparable
            continue
        np.testing.assert_allclose(p.detach().cpu().numpy(), d[f"final_{j}"], atol=2e-4)


# --------------------------------------------------------------------------------------------- transfer operator
def test_eigen_lag_loss_matches_reference_golden(tmp_path):
    """Transfer-operator branch (lag_tau > 0, reference core.py:412-416,428,440) against the reference's own golden run:
    loss, eigenvalues, objective, penalty, cvec and every parameter gradient."""
    from colvarsfinder import core, nn
    c = C.eigen_case("eigen_2d_lag")
    lag = int(round(c["lag_tau"] / c["dt"]))
    model = nn.EigenFunctions(c["layer_dims"], c["k"])
    with torch.no_grad():
        for i in range(c["k"]):
            for p, v in zip(model.eigen_funcs[i].parameters(), c["params"][i]):
                p.copy_(torch.as_tensor(v))
    traj = FakeTrajectory(c["X"], c["w"].astype(np.float64), dt=c["dt"])
    task = core.EigenFunctionTask(traj, torch.nn.Identity(), model, str(tmp_path), c["alpha"], c["eig_w"], beta=c["beta"],
                                  lag_tau=c["lag_tau"], sort_eigvals_in_training=c["sort"], k=c["k"], device=DEV, verbose=False,
                                  debug_mode=False)
    assert task.lag_idx == lag
    X, w = task._traj, task._weights
    out = task.loss_func(X[:-lag].contiguous(), w[:-lag].contiguous(), X[lag:].contiguous(), w[lag:].contiguous())
    out[0].backward()
    grads = [[p.grad.cpu().numpy() for p in f.parameters()] for f in model.eigen_funcs]
    gold = dict(loss=float(c["g64_loss"]), obj=float(c["g64_obj"]), pen=float(c["g64_pen"]), eig=c["g64_eig"],
                cvec=c["g64_cvec"], grads=c["g64"], loss32=float(c["r32_loss"]), obj32=float(c["r32_obj"]),
                pen32=float(c["r32_pen"]), eig32=c["r32_eig"], grads32=c["g32"])
    _check_eigen(c, out, grads, gold, "eigen_2d_lag")


@pytest.mark.parametrize("case", ["dipeptide_fast_k3", "ring_general_k2_nosort"])
def test_eigen_lag_loss_matches_autograd_oracle(case, tmp_path):
    """Same branch on the fast kernels (aligned dipeptide frames, k = 3) and the general kernels (unsorted, k = 2) against
    the autograd restatement of the reference formula."""
    from colvarsfinder import core, nn, utils
    lag, B = 3, 1000
    if case == "dipeptide_fast_k3":
        dims, k, sort = [66, 20, 20, 20, 1], 3, True
        X = ref_torch.synth_frames(BASE, B + lag, seed=31)
        pp, ppo = utils.Align(BASE, list(range(22))), ref_torch.Preprocess(ref_torch.Align(BASE, list(range(22))), None)
    else:
        dims, k, sort = [2, 9, 7, 1], 2, False
        X = np.random.default_rng(5).normal(size=(B + lag, 2)).astype(np.float32)
        pp, ppo = torch.nn.Identity(), ref_torch.Preprocess()
    w = ref_torch.boltzmann_weights(B + lag, seed=8)
    nets = _random_nets(dims, k, seed=17)
    eig_w = [1.0, 0.6, 0.3][:k]
    model = nn.EigenFunctions(dims, k)
    with torch.no_grad():
        for i in range(k):
            for p, v in zip(model.eigen_funcs[i].parameters(), nets[i]):
                p.copy_(torch.as_tensor(v))
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=0.5), pp, model, str(tmp_path), 15.0, eig_w,
                                  lag_tau=1.5, sort_eigvals_in_training=sort, k=k, device=DEV, verbose=False, debug_mode=False)
    assert task.lag_idx == lag and task._ctx.fast_path == (case == "dipeptide_fast_k3")
    Xd, wd = task._traj, task._weights
    out = task.loss_func(Xd[:-lag].contiguous(), wd[:-lag].contiguous(), Xd[lag:].contiguous(), wd[lag:].contiguous())
    out[0].backward()
    _assert_lag_within_envelope(out, model, X, w, nets, ppo, 15.0, eig_w, lag, 1.5, sort=sort, tag=case)


def test_eigen_lag_train_follows_oracle_loop(tmp_path):
    """train() with a time lag: the split is drawn on the first n - lag frames, X_lagged = traj[index + lag]
    (reference core.py:461-512); per-iteration losses follow a CPU autograd loop on the same batches."""
    from colvarsfinder import core, nn
    n, lag, bs = 900, 2, 200
    rng = np.random.default_rng(3)
    X = np.cumsum(rng.normal(scale=0.2, size=(n, 2)), 0).astype(np.float32)
    X -= X.mean(0)
    w = ref_torch.boltzmann_weights(n, seed=4)
    torch.manual_seed(21)
    model = nn.EigenFunctions([2, 8, 8, 1], 2)
    nets = [[p.detach().clone().double().requires_grad_() for p in f.parameters()] for f in model.eigen_funcs]
    task = core.EigenFunctionTask(FakeTrajectory(X.astype(np.float64), w.astype(np.float64), dt=0.1), torch.nn.Identity(), model,
                                  str(tmp_path), 10.0, [1.0, 0.5], lag_tau=0.2, learning_rate=0.01, k=2, batch_size=bs,
                                  num_epochs=2, test_ratio=0.25, save_model_every_step=0, device=DEV, verbose=False,
                                  debug_mode=False)
    np.random.seed(5)
    task.train()
    got = np.concatenate([l[0].numpy()[:, 0] for l in task.loss_list])
    # oracle loop
    np.random.seed(5)
    tr, te = ref_torch.split_indices(n - lag, 0.25, draws=2)
    Xt, wt = torch.tensor(X, dtype=torch.float64), torch.tensor(w, dtype=torch.float64)
    opt = torch.optim.Adam([p for net in nets for p in net], lr=0.01)
    want = []
    spans, _ = ref_torch.batches(len(tr), bs)
    for epoch in range(2):
        for a, b in spans:
            idx = torch.as_tensor(tr[a:b])
            opt.zero_grad()
            out = ref_torch.eigen_loss(Xt[idx], wt[idx], nets, ref_torch.Preprocess(), 10.0, [1.0, 0.5], X_lagged=Xt[idx + lag],
                                       weight_lagged=wt[idx + lag], lag_time=0.2)
            out[0].backward()
            opt.step()
            want.append(float(out[0]))
    assert got.shape == (len(want),)
    np.testing.assert_allclose(got, want, rtol=2e-3)


# --------------------------------------------------------------------------------------------- autoencoder
def _ae_task(c, tmp, F=None, w=None):
    from colvarsfinder import core, nn
    model = nn.AutoEncoder(c["e_dims"], c["d_dims"])
    with torch.no_grad():
        for p, v in zip(model.encoder.parameters(), c["enc"]):
            p.copy_(torch.as_tensor(v))
        for p, v in zip(model.decoder.parameters(), c["dec"]):
            p.copy_(torch.as_tensor(v))
    F = c["F"] if F is None else F
    w = c["w"] if w is None else w
    task = core.AutoEncoderTask(FakeTrajectory(F, w.astype(np.float64)), torch.nn.Identity(), model, str(tmp), device=DEV,
                                verbose=False, debug_mode=False)
    return task, model


@pytest.mark.parametrize("name", C.AE_CASES)
def test_ae_loss_matches_reference_golden(name, tmp_path):
    c = C.ae_case(name)
    task, model = _ae_task(c, tmp_path)
    loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
    loss.backward()
    assert abs(float(loss) - c["g64_loss"]) <= C.tol(c["g64_loss"], c["r32_loss"])
    got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
    for g, g64, g32 in zip(got, c["g64_enc"] + c["g64_dec"], c["g32_enc"] + c["g32_dec"]):
        assert C.rel_l2(g, g64) <= max(1e-5, 2 * C.rel_l2(g32, g64)), (C.rel_l2(g, g64), C.rel_l2(g32, g64))
    with torch.no_grad():
        l2 = task.weighted_MSE_loss(task._feature_traj, task._weights)
    assert torch.equal(l2, loss.detach())


@pytest.mark.parametrize("B", [1, 100, 128, 130, 5000, 40001])
def test_ae_loss_matches_oracle_ragged_sizes(B, tmp_path):
    c = C.ae_case("ae_dipep")
    rng = np.random.default_rng(B)
    F = (c["F"][rng.integers(0, len(c["F"]), B)] + rng.normal(scale=0.05, size=(B, 66))).astype(np.float32)
    w = ref_torch.boltzmann_weights(B, seed=B)
    task, model = _ae_task(c, tmp_path, F, w)
    loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
    loss.backward()
    lo, genc, gdec = cf.ae_loss_and_grads(F, w, c["enc"], c["dec"])
    assert abs(float(loss) - lo) <= 1e-5 * abs(lo)
    got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
    for g, go in zip(got, genc + gdec):
        assert C.rel_l2(g, go) < 2e-5


@pytest.mark.parametrize("d,e", [(30, 2), (66, 2), (71, 2), (30, 1), (66, 3)])
@pytest.mark.parametrize("B", [3, 700, 20000])
def test_ae_fast_and_general_kernels_match_oracle(B, d, e, tmp_path):
    """The notebook-sized chain [d,20,20,20,e] + [e,10,10,d] (examples/dipeptide/main.ipynb:434 uses d = 30, e = 2) on the thread-private
    kernels (cvf_ae_fast.cu) and, with cvf_ae_set_fast_path(1), on the general row-engine kernel: both against the fp64 oracle,
    and against each other."""
    from colvarsfinder import _lib, core, nn
    rng = np.random.default_rng(B + d)
    torch.manual_seed(B + d)
    e_dims, d_dims = [d, 20, 20, 20, e], [e, 10, 10, d]
    enc = [p.numpy() for p in ref_torch.init_mlp_params(e_dims)]
    dec = [p.numpy() for p in ref_torch.init_mlp_params(d_dims)]
    F = rng.normal(scale=1.5, size=(B, d)).astype(np.float32)
    w = ref_torch.boltzmann_weights(B, seed=B)
    lo, genc, gdec = cf.ae_loss_and_grads(F, w, enc, dec)
    res = {}
    for mode in (0, 1):
        _lib.check(_lib.lib().cvf_ae_set_fast_path(mode), "cvf_ae_set_fast_path")
        try:
            task, model = _ae_task(dict(e_dims=e_dims, d_dims=d_dims, enc=enc, dec=dec, F=F, w=w), tmp_path)
            _lib.profile_read(reset=True)
            loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
            launched = _lib.profile_read(reset=True)
            assert (launched["ae_fast_main"][2] > 0) == (mode == 0) and (launched["ae_step"][2] > 0) == (mode == 1)
            loss.backward()
            with torch.no_grad():
                l_eval = task.weighted_MSE_loss(task._feature_traj, task._weights)
        finally:
            _lib.lib().cvf_ae_set_fast_path(0)
        assert torch.equal(l_eval, loss.detach())
        assert abs(float(loss) - lo) <= 1e-5 * abs(lo), (mode, float(loss), lo)
        got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
        for i, (g, go) in enumerate(zip(got, genc + gdec)):
            assert C.rel_l2(g, go) < 2e-5, (mode, i, C.rel_l2(g, go))
        res[mode] = float(loss)
    assert abs(res[0] - res[1]) <= 2e-6 * abs(res[1])


@pytest.mark.parametrize("mode", ["tensor_cores", "simt"])
@pytest.mark.parametrize("B", [130, 1000, 33000])
def test_ae_wide_layers_match_oracle(B, mode, tmp_path):
    """Networks whose weights do not fit shared memory take the layer-wise path (cvf_ae_wide.cu): dense products with fused
    epilogues and split-K weight gradients, either on the tensor cores (tcgen05, 3 x TF32 split, cvf_gemm_tc.cu) or as fp32
    SIMT products; 33000 frames span two chunks."""
    from colvarsfinder import core, nn, _lib
    _lib.check(_lib.lib().cvf_ae_set_wide_path(0 if mode == "tensor_cores" else 1), "cvf_ae_set_wide_path")
    e_dims, d_dims = [150, 260, 200, 2], [2, 200, 260, 150]
    torch.manual_seed(B)
    enc = [p.numpy() for p in ref_torch.init_mlp_params(e_dims)]
    dec = [p.numpy() for p in ref_torch.init_mlp_params(d_dims)]
    rng = np.random.default_rng(B)
    F = rng.normal(size=(B, 150)).astype(np.float32)
    w = ref_torch.boltzmann_weights(B, seed=B)
    model = nn.AutoEncoder(e_dims, d_dims)
    with torch.no_grad():
        for p, v in zip(model.encoder.parameters(), enc):
            p.copy_(torch.as_tensor(v))
        for p, v in zip(model.decoder.parameters(), dec):
            p.copy_(torch.as_tensor(v))
    task = core.AutoEncoderTask(FakeTrajectory(F, w.astype(np.float64)), torch.nn.Identity(), model, str(tmp_path), device=DEV,
                                verbose=False, debug_mode=False)
    loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
    loss.backward()
    lo, genc, gdec = cf.ae_loss_and_grads(F, w, enc, dec)
    assert abs(float(loss) - lo) <= 1e-5 * abs(lo)
    got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
    for g, go in zip(got, genc + gdec):
        assert C.rel_l2(g, go) < 2e-5, C.rel_l2(g, go)
    with torch.no_grad():
        assert torch.equal(task.weighted_MSE_loss(task._feature_traj, task._weights), loss.detach())


@pytest.mark.parametrize("mode", ["tensor_cores", "simt"])
@pytest.mark.parametrize("B,data", [(4096, "unit"), (33000, "unit"), (4096, "angstrom")])
def test_ae_c5_shape_matches_oracle(B, data, mode, tmp_path):
    """BASELINE config 5 at its real shape, AutoEncoder([3000,512,512,2],[2,512,512,3000]): the 3000-term accumulations are where
    the dropped lo*lo term of the 3 x TF32 split and the fp32 accumulation in tensor memory would show first.
    "unit": standardised features (what bench.py feeds): loss 1e-5, gradients 2e-5 rel-L2 against the fp64 closed form, as for
    the narrow cases.  "angstrom": raw coordinates of a 1000-atom chain (|x| up to tens of Angstrom) with torch's default
    initialisation saturate the first tanh layer, 1 - a^2 cancels in fp32 and the reference's own fp32 run loses digits: there
    the reference's fp32 error is the floor.  The achieved errors are printed."""
    from colvarsfinder import core, nn, _lib
    _lib.check(_lib.lib().cvf_ae_set_wide_path(0 if mode == "tensor_cores" else 1), "cvf_ae_set_wide_path")
    try:
        e_dims, d_dims = [3000, 512, 512, 2], [2, 512, 512, 3000]
        torch.manual_seed(B)
        enc = [p.numpy() for p in ref_torch.init_mlp_params(e_dims)]
        dec = [p.numpy() for p in ref_torch.init_mlp_params(d_dims)]
        rng = np.random.default_rng(B)
        if data == "unit":
            F = rng.normal(size=(B, 3000)).astype(np.float32)
        else:
            base = np.cumsum(rng.normal(scale=1.5 / np.sqrt(3), size=(1000, 3)), 0)
            base -= base.mean(0)
            F = (base.reshape(1, 3000) + rng.normal(scale=0.3, size=(B, 3000))).astype(np.float32)
        w = ref_torch.boltzmann_weights(B, seed=B)
        model = nn.AutoEncoder(e_dims, d_dims)
        with torch.no_grad():
            for p_, v in zip(model.encoder.parameters(), enc):
                p_.copy_(torch.as_tensor(v))
            for p_, v in zip(model.decoder.parameters(), dec):
                p_.copy_(torch.as_tensor(v))
        task = core.AutoEncoderTask(FakeTrajectory(F, w.astype(np.float64)), torch.nn.Identity(), model, str(tmp_path), device=DEV,
                                    verbose=False, debug_mode=False)
        loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
        loss.backward()
        lo, genc, gdec = cf.ae_loss_and_grads(F, w, enc, dec)
        got = [p_.grad.cpu().numpy() for p_ in model.encoder.parameters()] + [p_.grad.cpu().numpy() for p_ in model.decoder.parameters()]
        errs = [C.rel_l2(g, go) for g, go in zip(got, genc + gdec)]
        floor_loss, floors = 0.0, [0.0] * len(errs)
        if data == "angstrom":
            te = [torch.as_tensor(p_).requires_grad_() for p_ in enc]
            td = [torch.as_tensor(p_).requires_grad_() for p_ in dec]
            l32 = ref_torch.ae_loss(torch.as_tensor(F), torch.as_tensor(w), te, td)
            l32.backward()
            floor_loss = abs(float(l32) - lo)
            floors = [C.rel_l2(t.grad.numpy(), go) for t, go in zip(te + td, genc + gdec)]
        print(f"C5 shape, B={B}, {data}, {mode}: loss rel err {abs(float(loss) - lo) / abs(lo):.2e} (ref32 {floor_loss / abs(lo):.2e}), "
              f"gradient rel-L2 max {max(errs):.2e} (ref32 max {max(floors):.2e}); per tensor (W1 b1 W2 b2 ...): " +
              " ".join(f"{e:.1e}" for e in errs))
        assert abs(float(loss) - lo) <= max(1e-5 * abs(lo), floor_loss)
        for e, f in zip(errs, floors):
            assert e <= max(2e-5, f), (errs, floors)
    finally:
        _lib.check(_lib.lib().cvf_ae_set_wide_path(0), "cvf_ae_set_wide_path")


def test_ae_prepass_and_train_match_reference_run(tmp_path):
    from colvarsfinder import core, nn, utils
    d = C.load("train_ae_2d")
    torch.manual_seed(12)
    model = nn.AutoEncoder([2, 12, 12, 1], [1, 12, 2])
    for j, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.as_tensor(d[f"init_{j}"]))
    traj = FakeTrajectory(d["X"].astype(np.float64), d["w"].astype(np.float64), dt=0.1)
    task = core.AutoEncoderTask(traj, torch.nn.Identity(), model, str(tmp_path), learning_rate=0.005, batch_size=200,
                                num_epochs=3, test_ratio=0.2, save_model_every_step=0, device=DEV, verbose=False,
                                debug_mode=False)
    np.random.seed(78)
    task.train()
    tr = np.stack([l[0].numpy() for l in task.loss_list])
    np.testing.assert_allclose(tr, d["train_hist"], rtol=1e-3)
    np.testing.assert_allclose(np.stack([l[1].numpy() for l in task.loss_list]), d["test_hist"], rtol=1e-3)
    for j, p in enumerate(model.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), d[f"final_{j}"], atol=1e-4)
    # molecular pre-pass (core.py:635): features of the whole trajectory = aligned coordinates
    X = ref_torch.synth_frames(BASE, 3000, seed=4)
    ae = nn.AutoEncoder([66, 20, 20, 20, 2], [2, 10, 10, 66])
    t2 = core.AutoEncoderTask(FakeTrajectory(X, np.ones(3000)), utils.Align(BASE, list(range(22))), ae, str(tmp_path),
                              device=DEV, verbose=False, debug_mode=False)
    yo = cf.kabsch(X.astype(np.float64), list(range(22)), BASE)[0].reshape(3000, 66)
    assert np.abs(t2._feature_traj.cpu().numpy() - yo).max() < 1e-6
    assert t2.k == 2 and isinstance(t2.colvar_model(), torch.nn.Sequential)


def test_save_model_and_unsupported(tmp_path):
    from colvarsfinder import core, nn
    c = C.eigen_case("eigen_2d_k2_nosort")
    task, model = _eigen_task(c, tmp_path)
    task.save_model(0)
    import os
    assert os.path.isfile(tmp_path / "latest" / "model.pt") and os.path.isfile(tmp_path / "latest" / "0_1_weight.txt")
    sd = torch.load(tmp_path / "latest" / "model.pt")
    assert "eigen_funcs.1.2.bias" in sd
    traj = FakeTrajectory(c["X"], c["w"].astype(np.float64), dt=0.1)
    with pytest.raises(AssertionError, match="not divisable"):
        core.EigenFunctionTask(traj, torch.nn.Identity(), nn.EigenFunctions([2, 4, 1], 1), str(tmp_path), 1.0, [1.0],
                               lag_tau=0.25, device=DEV, verbose=False)
    with pytest.raises(RuntimeError, match="Tanh"):
        core.EigenFunctionTask(traj, torch.nn.Identity(), nn.EigenFunctions([2, 4, 1], 1, torch.nn.GELU()), str(tmp_path), 1.0,
                               [1.0], device=DEV, verbose=False)


@pytest.mark.gpu
def test_save_model_exports_scripted_collective_variables(tmp_path):
    """save_model writes scripted_cv_gpu.pt / scripted_cv_cpu.pt (reference core.py:212-227); the exported module -- stock torch
    ops only -- reproduces colvar_model(), which runs the CUDA pre-processing kernels."""
    from colvarsfinder import core, nn, utils
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    feats = [("bond", [1, 4]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]), ("position", [8]), ("dihedral", [6, 8, 14, 16])]
    X = ref_torch.synth_frames(BASE, 300, seed=11)
    w = ref_torch.boltzmann_weights(300, seed=12)
    for pp, d_r in ((utils.Align(BASE, list(range(22))), 66),
                    (utils.Preprocessing(utils.Align(BASE[heavy], heavy), utils.FeatureMap(feats)), 9)):
        torch.manual_seed(3)
        out_dir = tmp_path / f"d{d_r}"
        model = nn.EigenFunctions([d_r, 20, 20, 20, 1], 2)
        task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64), dt=1.0), pp, model, str(out_dir), 20.0, [1.0, 0.5], k=2,
                                      device=DEV, verbose=False, debug_mode=False)
        task.save_model(0)
        Xd = torch.as_tensor(X, device=DEV)
        want = task.colvar_model()(Xd)
        got_gpu = torch.jit.load(str(out_dir / "latest" / "scripted_cv_gpu.pt"))(Xd)
        got_cpu = torch.jit.load(str(out_dir / "latest" / "scripted_cv_cpu.pt"))(torch.as_tensor(X))
        assert want.shape == (300, 2)
        assert torch.allclose(got_gpu, want, atol=2e-5, rtol=1e-4), (got_gpu - want).abs().max()
        assert torch.allclose(got_cpu, want.cpu(), atol=2e-5, rtol=1e-4), (got_cpu - want.cpu()).abs().max()
        # the task still trains after the export (the parameters must still alias the flat buffer the kernels read)
        loss = task.loss_func(task._traj, task._weights, None, None)[0]
        loss.backward()
        assert all(p.grad is not None for p in model.parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["eigen", "eigen_lag", "ae"])
def test_train_loop_cuda_graph_replays_the_eager_steps(kind, tmp_path, monkeypatch):
    """train() replays its iteration as a CUDA graph after three eager iterations (core._GraphedStep).  Every iteration is a real
    step on its own mini-batch, so the per-iteration log and the final parameters must equal those of the eager loop
    (CVF_CUDA_GRAPH=0) bit for bit -- the kernels are deterministic and the graph launches the same ones."""
    from colvarsfinder import core, nn
    n = 2600
    rng = np.random.default_rng(12)
    X = np.cumsum(rng.normal(scale=0.2, size=(n, 2)), 0).astype(np.float32)
    X -= X.mean(0)
    w = ref_torch.boltzmann_weights(n, seed=13)
    results = []
    monkeypatch.setenv("CVF_CUDA_GRAPH_MIN_STEPS", "1")
    for graph in ("1", "0"):
        monkeypatch.setenv("CVF_CUDA_GRAPH", graph)
        torch.manual_seed(5)
        np.random.seed(6)
        traj = FakeTrajectory(X.astype(np.float64), w.astype(np.float64), dt=0.1)
        if kind == "ae":
            model = nn.AutoEncoder([2, 8, 1], [1, 8, 2])
            task = core.AutoEncoderTask(traj, torch.nn.Identity(), model, str(tmp_path / graph), learning_rate=0.01, batch_size=250,
                                        num_epochs=3, test_ratio=0.2, save_model_every_step=0, device=DEV, verbose=False,
                                        debug_mode=False)
        else:
            model = nn.EigenFunctions([2, 20, 20, 20, 1], 2)
            task = core.EigenFunctionTask(traj, torch.nn.Identity(), model, str(tmp_path / graph), 10.0, [1.0, 0.5],
                                          lag_tau=0.2 if kind == "eigen_lag" else 0, learning_rate=0.01, k=2, batch_size=250,
                                          num_epochs=3, test_ratio=0.2, save_model_every_step=0, device=DEV, verbose=False,
                                          debug_mode=False)
        task.train()
        gs = task._graphed_step
        assert (gs.replays > 0) == (graph == "1") and gs.eager_steps + gs.replays == 3 * (2080 // 250)
        hist = torch.cat([l[0].reshape(len(l[0]), -1) for l in task.loss_list])
        results.append((hist, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu(), task.train_loss_df.to_numpy()))
    assert torch.equal(results[0][0], results[1][0])
    assert torch.equal(results[0][1], results[1][1])
    assert np.array_equal(results[0][2], results[1][2])
