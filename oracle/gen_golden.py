"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

Every case feeds seeded fp32 inputs and explicit network parameters to the reference's own
``EigenFunctionTask.loss_func`` + ``backward`` (core.py:387-457,517) or
``AutoEncoderTask.weighted_MSE_loss`` + ``backward`` (core.py:652-666,708) and stores
what came out, once with the reference's default dtype float32 ("ref32") and once with
float64 on the same fp32-rounded inputs ("gold64").  Two further cases
store whole ``train()`` runs (split sizes, per-iteration losses, final parameters).

The pre-processing layer used by the molecular cases is oracle/ref_torch.py's Align/FeatureMap
(molann, the reference's own choice, is not vendored or installed: PARITY UNPINNED for that layer);
the loss, Jacobian, penalty and backward are the reference's code.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_import, ref_torch  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _mlp_params(layer_dims, rng, scale=1.0):
    ps = []
    for i in range(len(layer_dims) - 1):
        bound = scale / np.sqrt(layer_dims[i])
        ps.append(rng.uniform(-bound, bound, size=(layer_dims[i + 1], layer_dims[i])).astype(np.float32))
        ps.append(rng.uniform(-bound, bound, size=(layer_dims[i + 1],)).astype(np.float32))
    return ps


def _load_seq(seq, params, dtype):
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for j, m in enumerate(lin):
            m.weight.copy_(torch.as_tensor(params[2 * j]).to(dtype))
            m.bias.copy_(torch.as_tensor(params[2 * j + 1]).to(dtype))


def build_pp(spec):
    """spec: dict(kind='identity') or dict(kind='mol', ref=[n_a,3], align_idx=[n_a] or None, features=[(type, idx)] or None)."""
    if spec["kind"] == "identity":
        return torch.nn.Identity()
    align = None
    if spec.get("align_idx") is not None:
        align = ref_torch.Align(spec["ref"], spec["align_idx"])
    fmap = None
    if spec.get("features") is not None:
        fmap = ref_torch.FeatureMap(spec["features"])
    return ref_torch.Preprocess(align, fmap)


def run_eigen(case, dtype):
    core, nn, _ = ref_import.load()
    torch.set_default_dtype(dtype)
    try:
        k = case["k"]
        model = nn.EigenFunctions(case["layer_dims"], k)
        for i in range(k):
            _load_seq(model.eigen_funcs[i], case["params"][i], dtype)
        X = case["X"]
        traj = ref_import.FakeTrajectory(X, case["w"].astype(np.float64), dt=case.get("dt", 1.0))
        with tempfile.TemporaryDirectory() as tmp:
            task = core.EigenFunctionTask(
                traj, build_pp(case["pp"]), model, tmp, case["alpha"], case["eig_w"],
                diag_coeff=None if case.get("diag_coeff") is None else torch.as_tensor(case["diag_coeff"]).to(dtype),
                beta=case.get("beta", 1.0), lag_tau=case.get("lag_tau", 0), k=k,
                sort_eigvals_in_training=case.get("sort", True), verbose=False, debug_mode=False)
            Xt = torch.as_tensor(X).to(dtype)
            wt = torch.as_tensor(case["w"]).to(dtype)
            if task.lag_idx == 0:
                Xt.requires_grad_()
                out = task.loss_func(Xt, wt, None, None)
            else:
                lag = task.lag_idx
                out = task.loss_func(Xt[:-lag], wt[:-lag], Xt[lag:], wt[lag:])
            loss, eig, obj, pen, cvec = out
            loss.backward()
            grads = [[p.grad.detach().double().numpy() if p.grad is not None else np.zeros(tuple(p.shape))
                      for p in model.eigen_funcs[i].parameters()] for i in range(k)]
        return dict(loss=float(loss), eig=eig.detach().double().numpy(), obj=float(obj), pen=float(pen),
                    cvec=np.asarray(cvec, dtype=np.int64), grads=grads)
    finally:
        torch.set_default_dtype(torch.float32)


def run_ae(case, dtype):
    core, nn, _ = ref_import.load()
    torch.set_default_dtype(dtype)
    try:
        model = nn.AutoEncoder(case["e_dims"], case["d_dims"])
        _load_seq(model.encoder, case["enc"], dtype)
        _load_seq(model.decoder, case["dec"], dtype)
        F = case["F"]
        traj = ref_import.FakeTrajectory(F, case["w"].astype(np.float64))
        with tempfile.TemporaryDirectory() as tmp:
            task = core.AutoEncoderTask(traj, torch.nn.Identity(), model, tmp, verbose=False, debug_mode=False)
            loss = task.weighted_MSE_loss(torch.as_tensor(F).to(dtype), torch.as_tensor(case["w"]).to(dtype))
            loss.backward()
            genc = [p.grad.detach().double().numpy() for p in model.encoder.parameters()]
            gdec = [p.grad.detach().double().numpy() for p in model.decoder.parameters()]
        return dict(loss=float(loss), genc=genc, gdec=gdec)
    finally:
        torch.set_default_dtype(torch.float32)


def save_eigen(name, case):
    r32 = run_eigen(case, torch.float32)
    g64 = run_eigen(case, torch.float64)
    d = dict(X=case["X"], w=case["w"], k=case["k"], layer_dims=np.asarray(case["layer_dims"]),
             alpha=case["alpha"], eig_w=np.asarray(case["eig_w"], dtype=np.float64),
             beta=case.get("beta", 1.0), lag_tau=case.get("lag_tau", 0), dt=case.get("dt", 1.0),
             sort=case.get("sort", True), pp_kind=case["pp"]["kind"])
    if case.get("diag_coeff") is not None:
        d["diag_coeff"] = case["diag_coeff"]
    pp = case["pp"]
    if pp["kind"] == "mol":
        if pp.get("align_idx") is not None:
            d["ref"] = np.asarray(pp["ref"], dtype=np.float64)
            d["align_idx"] = np.asarray(pp["align_idx"], dtype=np.int64)
        if pp.get("features") is not None:
            d["feat_types"] = np.asarray([t for t, _ in pp["features"]])
            d["feat_atoms"] = np.asarray([list(a) + [-1] * (4 - len(a)) for _, a in pp["features"]], dtype=np.int64)
    for i in range(case["k"]):
        for j, p in enumerate(case["params"][i]):
            d[f"p_{i}_{j}"] = p
            d[f"g32_{i}_{j}"] = r32["grads"][i][j]
            d[f"g64_{i}_{j}"] = g64["grads"][i][j]
    for tag, r in (("r32", r32), ("g64", g64)):
        d[f"{tag}_loss"], d[f"{tag}_eig"], d[f"{tag}_obj"], d[f"{tag}_pen"], d[f"{tag}_cvec"] = \
            r["loss"], r["eig"], r["obj"], r["pen"], r["cvec"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: loss32={r32['loss']:.8g} loss64={g64['loss']:.12g} eig64={g64['eig']} cvec={g64['cvec']}")


def save_ae(name, case):
    r32 = run_ae(case, torch.float32)
    g64 = run_ae(case, torch.float64)
    d = dict(F=case["F"], w=case["w"], e_dims=np.asarray(case["e_dims"]), d_dims=np.asarray(case["d_dims"]),
             r32_loss=r32["loss"], g64_loss=g64["loss"])
    for j, p in enumerate(case["enc"]):
        d[f"enc_{j}"], d[f"g32_enc_{j}"], d[f"g64_enc_{j}"] = p, r32["genc"][j], g64["genc"][j]
    for j, p in enumerate(case["dec"]):
        d[f"dec_{j}"], d[f"g32_dec_{j}"], d[f"g64_dec_{j}"] = p, r32["gdec"][j], g64["gdec"][j]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: loss32={r32['loss']:.8g} loss64={g64['loss']:.12g}")


def ring_2d(n, seed):
    """2-d points on a noisy ring."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi, np.pi, n)
    r = rng.normal(1.0, 0.25, n)
    return np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)


def save_train_eigen(name):
    """Whole reference EigenFunctionTask.train() run on 2-d data (C1 shape, reduced n)."""
    core, nn, _ = ref_import.load()
    X = ring_2d(1200, 30)
    w = ref_torch.boltzmann_weights(1200, seed=5)
    torch.manual_seed(11)
    model = nn.EigenFunctions([2, 20, 20, 20, 1], 2)
    init = [p.detach().numpy().copy() for p in model.parameters()]
    traj = ref_import.FakeTrajectory(X.astype(np.float64), w.astype(np.float64), dt=0.1)
    with tempfile.TemporaryDirectory() as tmp:
        task = core.EigenFunctionTask(traj, torch.nn.Identity(), model, tmp, 20.0, [1.0, 0.7], beta=1.0, lag_tau=0,
                                      learning_rate=0.005, k=2, batch_size=300, num_epochs=3, test_ratio=0.2,
                                      save_model_every_step=0, verbose=False, debug_mode=False)
        np.random.seed(77)
        task.train()
    d = dict(X=X, w=w, train_hist=np.stack([l[0].numpy() for l in task.loss_list]),
             test_hist=np.stack([l[1].numpy() for l in task.loss_list]),
             train_df=task.train_loss_df.to_numpy(), test_df=task.test_loss_df.to_numpy())
    for j, p in enumerate(init):
        d[f"init_{j}"] = p
    for j, p in enumerate(model.parameters()):
        d[f"final_{j}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "train_hist", d["train_hist"].shape, "test_hist", d["test_hist"].shape, d["train_df"][-1])


def save_train_ae(name):
    core, nn, _ = ref_import.load()
    X = ring_2d(1000, 31)
    w = ref_torch.boltzmann_weights(1000, seed=6)
    torch.manual_seed(12)
    model = nn.AutoEncoder([2, 12, 12, 1], [1, 12, 2])
    init = [p.detach().numpy().copy() for p in model.parameters()]
    traj = ref_import.FakeTrajectory(X.astype(np.float64), w.astype(np.float64), dt=0.1)
    with tempfile.TemporaryDirectory() as tmp:
        task = core.AutoEncoderTask(traj, torch.nn.Identity(), model, tmp, learning_rate=0.005, batch_size=200,
                                    num_epochs=3, test_ratio=0.2, save_model_every_step=0, verbose=False,
                                    debug_mode=False)
        np.random.seed(78)
        task.train()
    d = dict(X=X, w=w, train_hist=np.stack([l[0].numpy() for l in task.loss_list]),
             test_hist=np.stack([l[1].numpy() for l in task.loss_list]),
             train_df=task.train_loss_df.to_numpy(), test_df=task.test_loss_df.to_numpy())
    for j, p in enumerate(init):
        d[f"init_{j}"] = p
    for j, p in enumerate(model.parameters()):
        d[f"final_{j}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "train_hist", d["train_hist"].shape, d["train_df"][-1])


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2026)

    # --- E1: C1 shape. 2-d, Identity, k=1, generator (examples/2d/2d.ipynb:487-490,614-622)
    dims = [2, 20, 20, 20, 1]
    save_eigen("eigen_2d_k1", dict(
        X=ring_2d(1000, 30), w=np.ones(1000, np.float32), k=1, layer_dims=dims,
        params=[_mlp_params(dims, rng, 2.0)], alpha=20.0, eig_w=[1.0], pp=dict(kind="identity")))

    # --- E2: 2-d, k=3, diag_coeff, beta != 1, Boltzmann weights, sorting on
    save_eigen("eigen_2d_k3_diag", dict(
        X=ring_2d(777, 31), w=ref_torch.boltzmann_weights(777, seed=1), k=3, layer_dims=[2, 16, 8, 1],
        params=[_mlp_params([2, 16, 8, 1], rng, 3.0) for _ in range(3)], alpha=7.5, eig_w=[1.0, 0.6, 0.3],
        diag_coeff=np.array([0.5, 2.0], np.float32), beta=1.7, pp=dict(kind="identity")))

    # --- E3: no sorting
    save_eigen("eigen_2d_k2_nosort", dict(
        X=ring_2d(500, 32), w=ref_torch.boltzmann_weights(500, seed=2), k=2, layer_dims=[2, 10, 1],
        params=[_mlp_params([2, 10, 1], rng, 3.0) for _ in range(2)], alpha=3.0, eig_w=[1.0, 0.5], sort=False,
        pp=dict(kind="identity")))

    # --- E4: C3 shape. 22 atoms, Kabsch on all atoms, position features (d_r = 66), k=3, generator
    base = ref_torch.DIPEPTIDE_NM * 10.0
    Xm = ref_torch.synth_frames(base, 768, seed=3)
    dims = [66, 20, 20, 20, 1]
    save_eigen("eigen_dipep_k3", dict(
        X=Xm, w=ref_torch.boltzmann_weights(768, seed=3), k=3, layer_dims=dims,
        params=[_mlp_params(dims, rng, 1.0) for _ in range(3)], alpha=20.0, eig_w=[1.0, 0.6, 0.3],
        pp=dict(kind="mol", ref=base, align_idx=list(range(22)), features=[("position", list(range(22)))])))

    # --- E5: align on a subset (heavy atoms), position features of a subset, diag_coeff over all 3N coordinates
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    dims = [30, 12, 12, 1]
    save_eigen("eigen_dipep_subset_diag", dict(
        X=Xm[:400], w=ref_torch.boltzmann_weights(400, seed=4), k=2, layer_dims=dims,
        params=[_mlp_params(dims, rng, 1.0) for _ in range(2)], alpha=5.0, eig_w=[1.0, 0.5],
        diag_coeff=rng.uniform(0.5, 2.0, 66).astype(np.float32), beta=2.5,
        pp=dict(kind="mol", ref=base[heavy], align_idx=heavy, features=[("position", heavy)])))

    # --- E6: internal-coordinate features (bond, angle, dihedral) after alignment, k=2
    feats = [("bond", [1, 4]), ("bond", [4, 6]), ("bond", [8, 14]), ("angle", [1, 4, 6]), ("angle", [6, 8, 14]),
             ("dihedral", [4, 6, 8, 14]), ("dihedral", [6, 8, 14, 16]), ("dihedral", [1, 4, 6, 8]),
             ("position", [8, 10])]
    dims = [3 + 2 + 6 + 6, 16, 16, 1]
    save_eigen("eigen_dipep_features", dict(
        X=Xm[:512], w=ref_torch.boltzmann_weights(512, seed=5), k=2, layer_dims=dims,
        params=[_mlp_params(dims, rng, 1.5) for _ in range(2)], alpha=10.0, eig_w=[1.0, 0.4],
        diag_coeff=rng.uniform(0.5, 2.0, 66).astype(np.float32),
        pp=dict(kind="mol", ref=base[heavy], align_idx=heavy, features=feats)))

    # --- E7: invariant features only, no alignment layer at all
    feats = [("bond", [0, 1]), ("bond", [4, 5]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]),
             ("dihedral", [6, 8, 14, 16])]
    dims = [2 + 1 + 4, 10, 1]
    save_eigen("eigen_dipep_invariant", dict(
        X=Xm[:300], w=np.ones(300, np.float32), k=2, layer_dims=dims,
        params=[_mlp_params(dims, rng, 1.5) for _ in range(2)], alpha=10.0, eig_w=[1.0, 0.4],
        pp=dict(kind="mol", align_idx=None, features=feats)))

    # --- E8: transfer operator (lag 2 frames), 2-d, k=2  (core.py:412-416,428,440)
    Xl = np.cumsum(np.random.default_rng(8).normal(scale=0.1, size=(600, 2)), 0).astype(np.float32)
    save_eigen("eigen_2d_lag", dict(
        X=Xl, w=ref_torch.boltzmann_weights(600, seed=8), k=2, layer_dims=[2, 12, 12, 1],
        params=[_mlp_params([2, 12, 12, 1], rng, 2.0) for _ in range(2)], alpha=10.0, eig_w=[1.0, 0.5],
        lag_tau=0.2, dt=0.1, pp=dict(kind="identity")))

    # --- A1/A2: autoencoder losses
    save_ae("ae_2d", dict(F=ring_2d(640, 33), w=ref_torch.boltzmann_weights(640, seed=9),
                          e_dims=[2, 20, 20, 1], d_dims=[1, 20, 20, 2],
                          enc=_mlp_params([2, 20, 20, 1], rng, 2.0), dec=_mlp_params([1, 20, 20, 2], rng, 2.0)))
    al = ref_torch.Align(base, list(range(22)))
    with torch.no_grad():
        Fm = al(torch.as_tensor(Xm[:512]).double()).reshape(512, 66).float().numpy()
    save_ae("ae_dipep", dict(F=Fm, w=ref_torch.boltzmann_weights(512, seed=10),
                             e_dims=[66, 20, 20, 20, 2], d_dims=[2, 10, 10, 66],
                             enc=_mlp_params([66, 20, 20, 20, 2], rng, 1.0),
                             dec=_mlp_params([2, 10, 10, 66], rng, 1.0)))

    # --- T1/T2: whole train() runs
    save_train_eigen("train_eigen_2d")
    save_train_ae("train_ae_2d")


if __name__ == "__main__":
    main()
