"""RegAutoEncoderTask (reference core.py:746-1217) on the CUDA passes, against golden vectors written by the
unmodified reference (oracle/gen_golden_regae.py): every loss term, the eigenvalues and their ordering, every parameter gradient,
and a whole train() history.  Also: the torch.library operators behind all three tasks pass torch.library.opcheck."""
import numpy as np
import pytest
import torch

from oracle import ref_torch
from oracle.ref_import import FakeTrajectory
from tests import _cases as C

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def _load(seq, params):
    with torch.no_grad():
        for p, v in zip(seq.parameters(), params):
            p.copy_(torch.as_tensor(v))


def _task(c, tmp, **kw):
    from colvarsfinder import core, nn, utils
    model = nn.RegAutoEncoder(c["e_dims"], c["d_dims"], c["r_dims"], c["K"])
    _load(model.encoder, c["enc"])
    _load(model.decoder, c["dec"])
    for i in range(c["K"]):
        _load(model.reg[i], c["reg"][i])
    pp = torch.nn.Identity() if c["pp_kind"] == "identity" else utils.Align(c["ref"], c["align_idx"])
    task = core.RegAutoEncoderTask(FakeTrajectory(c["X"], c["w"].astype(np.float64), dt=c["dt"]), pp, model, str(tmp), c["eig_w"],
                                   alpha=c["alpha"], gamma=c["gamma"], eta=c["eta"], lag_tau_ae=c["lag_tau_ae"],
                                   lag_tau_reg=c["lag_tau_reg"], beta=c["beta"], freeze_encoder=c["freeze"], device=DEV,
                                   verbose=False, debug_mode=False, **kw)
    return task, model


@pytest.mark.parametrize("name", C.REGAE_CASES)
def test_regae_loss_terms_and_gradients_match_reference_golden(name, tmp_path):
    c = C.regae_case(name)
    task, model = _task(c, tmp_path)
    halo = max(task.lag_idx, task.lag_ae_idx)
    n = c["X"].shape[0] - halo
    idx = torch.arange(n, device=DEV)
    X, w = task._traj[:n], task._weights[:n]
    Xa = task._traj[idx + task.lag_ae_idx] if task.lag_ae_idx > 0 else None
    Xr = task._traj[idx + task.lag_idx] if task.lag_idx > 0 else None
    wr = task._weights[idx + task.lag_idx] if task.lag_idx > 0 else None
    if c["freeze"]:
        for p in model.encoder.parameters():
            p.requires_grad = False
    loss, row, cvec = task._total_loss(X, w, Xa, Xr, wr)
    loss.backward()
    row = row.cpu().numpy()
    K = c["K"]
    assert list(cvec.cpu().numpy()) == list(c["g64_cvec"])
    names = ["loss", "ae", "g0", "g1"]
    for j, f in enumerate(names):
        assert abs(row[j] - c[f"g64_{f}"]) <= C.tol(c[f"g64_{f}"], c[f"r32_{f}"], slack=1.0) + 1e-7, (f, row[j], c[f"g64_{f}"])
    for i in range(K):
        assert abs(row[4 + i] - c["g64_eig"][i]) <= C.tol(c["g64_eig"][i], c["r32_eig"][i], slack=1.0), ("eig", i)
    for j, f in enumerate(["e0", "e1", "e2"]):
        assert abs(row[4 + K + j] - c[f"g64_{f}"]) <= C.tol(c[f"g64_{f}"], c[f"r32_{f}"], slack=1.0) + 1e-7, (f, row[4 + K + j])

    def grads(seq):
        return [np.zeros(tuple(p.shape)) if p.grad is None else p.grad.cpu().numpy() for p in seq.parameters()]
    got = grads(model.encoder) + grads(model.decoder) + [g for i in range(K) for g in grads(model.reg[i])]
    g64 = c["g64_enc"] + c["g64_dec"] + [g for i in range(K) for g in c["g64_reg"][i]]
    g32 = c["g32_enc"] + c["g32_dec"] + [g for i in range(K) for g in c["g32_reg"][i]]
    scale = max(np.abs(g).max() for g in g64)
    for j, (g, a, b) in enumerate(zip(got, g64, g32)):
        if np.abs(a).max() < 1e-9 * scale:      # frozen encoder / last-layer biases of the regularisers: zero gradient
            assert np.abs(g).max() < 1e-5 * scale, j
            continue
        assert C.rel_l2(g, a) <= max(2e-5, C.rel_l2(b, a)), (j, C.rel_l2(g, a), C.rel_l2(b, a))


def test_regae_public_loss_methods(tmp_path):
    """The reference's per-term methods (core.py:876-973) with their own signatures."""
    c = C.regae_case("regae_2d_generator")
    task, model = _task(c, tmp_path)
    X, w = task._traj, task._weights
    assert abs(float(task.weighted_MSE_loss(X, X, w)) - c["g64_ae"]) <= C.tol(c["g64_ae"], c["r32_ae"])
    assert abs(float(task.reg_enc_grad_loss(X, w)) - c["g64_e0"]) <= C.tol(c["g64_e0"], c["r32_e0"])
    assert abs(float(task.reg_enc_norm_loss(X, w)) - c["g64_e1"]) <= C.tol(c["g64_e1"], c["r32_e1"])
    assert abs(float(task.reg_enc_orthognal_loss(X, w)) - c["g64_e2"]) <= C.tol(c["g64_e2"], c["r32_e2"])
    eig, obj, pen, cvec = task.reg_eigen_loss(X, w, None, None)
    assert list(cvec.cpu().numpy()) == list(c["g64_cvec"])
    assert abs(float(obj) - c["g64_g0"]) <= C.tol(c["g64_g0"], c["r32_g0"])
    assert abs(float(pen) - c["g64_g1"]) <= C.tol(c["g64_g1"], c["r32_g1"])
    cv, reg = task.colvar_model(), task.reg_model()
    assert cv(X[:5]).shape == (5, 2) and reg(X[:5]).shape == (5, 2)


def test_regae_train_reproduces_reference_history(tmp_path):
    """train() on the same data, initial weights and numpy RNG state as the reference run (core.py:1039-1217)."""
    from colvarsfinder import core, nn
    d = C.load("train_regae_2d")
    torch.manual_seed(13)
    model = nn.RegAutoEncoder([2, 20, 20, 20, 1], [1, 20, 20, 2], [1, 20, 20, 1], 1)
    for j, p in enumerate(model.parameters()):
        np.testing.assert_array_equal(p.detach().numpy(), d[f"init_{j}"])      # same construction order, same init stream
    traj = FakeTrajectory(d["X"].astype(np.float64), d["w"].astype(np.float64), dt=0.1)
    task = core.RegAutoEncoderTask(traj, torch.nn.Identity(), model, str(tmp_path), [1.0], gamma=[1, 20], eta=[0, 0, 0],
                                   lag_tau_ae=0.1, lag_tau_reg=0.1, learning_rate=0.005, test_ratio=0.2, batch_size=240,
                                   num_epochs=3, save_model_every_step=0, device=DEV, verbose=False, debug_mode=False)
    np.random.seed(79)
    task.train()
    tr = np.stack([l[0].numpy() for l in task.loss_list])
    te = np.stack([l[1].numpy() for l in task.loss_list])
    assert tr.shape == d["train_hist"].shape and te.shape == d["test_hist"].shape
    np.testing.assert_allclose(tr, d["train_hist"], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(te, d["test_hist"], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(task.train_loss_df.to_numpy(), d["train_df"], rtol=2e-3, atol=1e-5)
    # The last-layer bias of a regulariser has an exactly zero gradient in the eigenfunction loss (it cancels from the variance,
    # the covariance and y' - y); the reference's autograd leaves rounding noise there, which Adam's normalisation turns into
    # full-size steps of arbitrary sign.  The parameter does not enter the loss: it is left out of the comparison.
    skip = {id(list(f.parameters())[-1]) for f in model.reg}
    for j, p in enumerate(model.parameters()):
        if id(p) in skip:
            continue
        np.testing.assert_allclose(p.detach().cpu().numpy(), d[f"final_{j}"], rtol=5e-3, atol=2e-4)


def test_time_lagged_target_of_the_reconstruction_loss(tmp_path):
    """cvf_ae_step_target against plain torch on the same networks (fp64), several batch sizes."""
    from colvarsfinder import _ops, nn
    torch.manual_seed(3)
    model = nn.AutoEncoder([6, 14, 2], [2, 9, 6]).to(DEV)
    actx = _ops.AEContext(model, DEV)
    for B in (1, 37, 128, 1000, 5003):
        F = torch.randn(B, 6, device=DEV)
        T = F + 0.3 * torch.randn(B, 6, device=DEV)
        w = torch.rand(B, device=DEV) + 0.5
        model.zero_grad(set_to_none=True)
        loss = _ops.ae_loss(actx, F, w, target=T)
        loss.backward()
        m64 = nn.AutoEncoder([6, 14, 2], [2, 9, 6]).to(DEV).double()
        m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
        ref = (w.double() * ((m64(F.double()) - T.double()) ** 2).sum(1)).sum() / w.double().sum()
        ref.backward()
        assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
        for p, q in zip(model.parameters(), m64.parameters()):
            assert C.rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 2e-5


def test_custom_operators_pass_opcheck(tmp_path):
    """The cvf:: operators are registered with torch.library (schema, fake kernels, autograd formulas): torch.library.opcheck
    exercises all three on real inputs."""
    from colvarsfinder import _ops, nn, utils
    torch.manual_seed(0)
    model = nn.EigenFunctions([66, 20, 20, 20, 1], 2).to(DEV)
    base = ref_torch.DIPEPTIDE_NM * 10.0
    ectx = _ops.EigenContext(model, utils.Align(base, list(range(22))).to(DEV), (22, 3), DEV, 20.0, [1.0, 0.5], 1.0, None, True)
    X = torch.as_tensor(ref_torch.synth_frames(base, 300, seed=1), device=DEV)
    w = torch.as_tensor(ref_torch.boltzmann_weights(300, seed=1), device=DEV)
    params = ectx.flat.flat.clone().requires_grad_()
    tests = ("test_schema", "test_autograd_registration", "test_faketensor")
    torch.library.opcheck(torch.ops.cvf.eigen_stats.default, (X, w, params, ectx.handle, 0), test_utils=tests)
    y, stats = torch.ops.cvf.eigen_stats(X, w, params.detach(), ectx.handle, 0)
    torch.library.opcheck(torch.ops.cvf.eigen_combine.default, (stats.clone().requires_grad_(), ectx.handle), test_utils=tests)
    torch.library.opcheck(torch.ops.cvf.eigen_loss.default, (X, w, params, ectx.handle), test_utils=tests)
    # the fused operator and the two-operator composition give the same loss and gradients
    la = _ops.eigen_loss(ectx, X, w)[0]
    ga = torch.autograd.grad(la, list(model.parameters()))
    y2, st2 = torch.ops.cvf.eigen_stats(X, w, ectx.packed_params(), ectx.handle, 0)
    lb = torch.ops.cvf.eigen_combine(st2, ectx.handle)[0].to(torch.float32)
    gb = torch.autograd.grad(lb, list(model.parameters()))
    assert torch.equal(la.detach(), lb.detach())
    for a_, b_ in zip(ga, gb):
        assert torch.allclose(a_, b_, rtol=1e-5, atol=1e-7 * float(a_.abs().max()))
    coef = torch.zeros(ectx.n_comb, dtype=torch.float64, device=DEV)
    torch.library.opcheck(torch.ops.cvf.eigen_grad.default, (X, w, y, params.detach(), coef, None, ectx.handle, 0, -1), test_utils=tests)
    yl = y + 0.1
    torch.library.opcheck(torch.ops.cvf.eigen_tlag_sx.default, (y.clone().requires_grad_(), yl, w, ectx.handle), test_utils=tests)
    ae = nn.AutoEncoder([66, 20, 20, 20, 2], [2, 10, 10, 66]).to(DEV)
    actx = _ops.AEContext(ae, DEV)
    F = X.reshape(300, 66).contiguous()
    torch.library.opcheck(torch.ops.cvf.ae_sums.default, (F, None, w, actx.flat.flat.clone().requires_grad_(), True, actx.handle),
                          test_utils=tests)
    al = utils.Align(base, list(range(22))).to(DEV)
    torch.library.opcheck(torch.ops.cvf.align_fwd.default, (X, al.ref_pos, al.align_idx), test_utils=("test_schema", "test_faketensor"))
    # the operators compose under autograd exactly like the task's loss_func
    loss = _ops.eigen_loss(ectx, X, w)[0]
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("n", [1, 5, 2048, 2049, 100003])
def test_weight_filter_on_device_matches_reference_semantics(n, tmp_path):
    """cvf_weights_filter against the reference's host-side arithmetic (utils.py:140-169): normalise, select (min_w, max_w),
    renormalise; kept indices ascending; and WeightedTrajectory(device=...) feeding a task directly."""
    from colvarsfinder import core, nn, utils
    rng = np.random.default_rng(n)
    w = np.exp(rng.normal(size=n))
    lo, hi = (0.2, 3.0) if n > 1 else (0.0, float("inf"))
    idx, wk = utils.filter_weights(torch.as_tensor(w, device=DEV), lo, hi)
    wn = w / w.mean()
    keep = (wn > lo) & (wn < hi)
    assert np.array_equal(idx.cpu().numpy(), np.nonzero(keep)[0])
    np.testing.assert_allclose(wk.cpu().numpy(), wn[keep] / wn[keep].mean(), rtol=1e-13)
    if n == 2049:
        t = 0.5 * np.arange(n)
        X = rng.normal(size=(n, 2))
        np.savetxt(tmp_path / "traj.txt", np.column_stack([t, X]))
        np.savetxt(tmp_path / "w.txt", w)
        host = utils.WeightedTrajectory(traj_filename=str(tmp_path / "traj.txt"), weight_filename=str(tmp_path / "w.txt"), min_w=lo,
                                        max_w=hi, verbose=False)
        devt = utils.WeightedTrajectory(traj_filename=str(tmp_path / "traj.txt"), weight_filename=str(tmp_path / "w.txt"), min_w=lo,
                                        max_w=hi, verbose=False, device=DEV)
        assert devt.trajectory.is_cuda and devt.weights.is_cuda
        np.testing.assert_array_equal(devt.trajectory.cpu().numpy(), host.trajectory)
        np.testing.assert_allclose(devt.weights.cpu().numpy(), host.weights, rtol=1e-13)
        torch.manual_seed(2)
        model = nn.EigenFunctions([2, 20, 20, 20, 1], 1)
        task = core.EigenFunctionTask(devt, torch.nn.Identity(), model, str(tmp_path), 20.0, [1.0], k=1, device=DEV, verbose=False,
                                      debug_mode=False)
        assert task._traj.shape == (int(keep.sum()), 2)
        loss = task.loss_func(task._traj, task._weights)[0]
        assert torch.isfinite(loss)


@pytest.mark.parametrize("act", ["sigmoid", "softplus", "elu", "relu"])
def test_other_activations_match_reference_golden(act, tmp_path):
    """Activations other than Tanh (the reference takes any module, nn.py:29-59): generator loss on 2-d data and on aligned
    frames with a feature map, autoencoder loss -- against the unmodified reference (oracle/gen_golden_activations.py), fp64 gold
    with the reference's fp32 run as the floor.  These run on the general kernels (the thread-private ones are tanh-only)."""
    from colvarsfinder import core, nn, utils
    d = C.load("activations")
    make = {"sigmoid": torch.nn.Sigmoid, "softplus": torch.nn.Softplus, "elu": torch.nn.ELU, "relu": torch.nn.ReLU}[act]
    feats = [(str(t), [int(a) for a in row if a >= 0]) for t, row in zip(d["feat_types"], d["feat_atoms"])]
    cases = (("e2d", [2, 12, 12, 1], torch.nn.Identity(), 10.0, [1.0, 0.5]),
             ("emol", [12, 14, 14, 1], utils.Preprocessing(utils.Align(d["emol_ref"], d["emol_align"]), utils.FeatureMap(feats)), 10.0, [1.0, 0.4]))
    for tag, dims, pp, alpha, eig_w in cases:
        model = nn.EigenFunctions(dims, 2, make())
        with torch.no_grad():
            for i in range(2):
                for j, p in enumerate(model.eigen_funcs[i].parameters()):
                    p.copy_(torch.as_tensor(d[f"{tag}_p_{i}_{j}"]))
        X, w = d[f"{tag}_X"], d[f"{tag}_w"]
        task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), pp, model, str(tmp_path), alpha, eig_w, k=2, device=DEV,
                                      verbose=False, debug_mode=False)
        assert not task._ctx.fast_path
        loss, eig, obj, pen, cvec = task.loss_func(task._traj, task._weights)
        loss.backward()
        g64l, r32l = float(d[f"{act}_{tag}_g64_loss"]), float(d[f"{act}_{tag}_r32_loss"])
        assert list(cvec.cpu().numpy()) == list(d[f"{act}_{tag}_g64_cvec"]), (act, tag)
        assert abs(float(loss) - g64l) <= C.tol(g64l, r32l), (act, tag, float(loss), g64l, r32l)
        for i in range(2):
            ge, re_ = d[f"{act}_{tag}_g64_eig"][i], d[f"{act}_{tag}_r32_eig"][i]
            assert abs(float(eig[i]) - ge) <= C.tol(ge, re_), (act, tag, "eig", i)
            for j, p in enumerate(model.eigen_funcs[i].parameters()):
                a, b = d[f"{act}_{tag}_g64_g_{i}_{j}"], d[f"{act}_{tag}_r32_g_{i}_{j}"]
                if np.abs(a).max() < 1e-12:
                    continue
                assert C.rel_l2(p.grad.cpu().numpy(), a) <= max(2e-5, C.rel_l2(b, a)), (act, tag, i, j, C.rel_l2(p.grad.cpu().numpy(), a), C.rel_l2(b, a))
    ae = nn.AutoEncoder([2, 16, 16, 1], [1, 16, 2], make())
    with torch.no_grad():
        for j, p in enumerate(list(ae.encoder.parameters()) + list(ae.decoder.parameters())):
            p.copy_(torch.as_tensor(d[f"ae_p_{j}"]))
    task = core.AutoEncoderTask(FakeTrajectory(d["ae_F"], d["ae_w"].astype(np.float64)), torch.nn.Identity(), ae, str(tmp_path), device=DEV,
                                verbose=False, debug_mode=False)
    loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
    loss.backward()
    assert abs(float(loss) - float(d[f"{act}_ae_g64_loss"])) <= C.tol(float(d[f"{act}_ae_g64_loss"]), float(d[f"{act}_ae_r32_loss"]))
    for j, p in enumerate(list(ae.encoder.parameters()) + list(ae.decoder.parameters())):
        a, b = d[f"{act}_ae_g64_g_{j}"], d[f"{act}_ae_r32_g_{j}"]
        assert C.rel_l2(p.grad.cpu().numpy(), a) <= max(2e-5, C.rel_l2(b, a)), (act, "ae", j)
