"""Golden vectors of RegAutoEncoderTask (reference core.py:746-1217) from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden_regae.py

Loss cases: seeded fp32 inputs and explicit parameters go through the reference's own loss methods
(``weighted_MSE_loss``, ``reg_enc_grad_loss``, ``reg_enc_norm_loss``, ``reg_enc_orthognal_loss``, ``reg_eigen_loss``), combined
exactly as the loop body of ``train()`` does (core.py:1066-1113), then ``loss.backward()``; stored once in the reference's
default float32 ("r32") and once in float64 on the same fp32-rounded inputs ("g64").  One further case stores a whole
``train()`` run (2-d data, the configuration of examples/2d/2d.ipynb:716-721 with smaller n).
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
from oracle.gen_golden import OUT, _load_seq, _mlp_params, build_pp, ring_2d  # noqa: E402


def run_case(case, dtype):
    core, nn, _ = ref_import.load()
    torch.set_default_dtype(dtype)
    try:
        K = case["K"]
        model = nn.RegAutoEncoder(case["e_dims"], case["d_dims"], case["r_dims"], K)
        _load_seq(model.encoder, case["enc"], dtype)
        _load_seq(model.decoder, case["dec"], dtype)
        for i in range(K):
            _load_seq(model.reg[i], case["reg"][i], dtype)
        X = case["X"]
        traj = ref_import.FakeTrajectory(X, case["w"].astype(np.float64), dt=case["dt"])
        with tempfile.TemporaryDirectory() as tmp:
            task = core.RegAutoEncoderTask(traj, build_pp(case["pp"]), model, tmp, case["eig_w"], alpha=case["alpha"],
                                           gamma=case["gamma"], eta=case["eta"], lag_tau_ae=case["lag_tau_ae"],
                                           lag_tau_reg=case["lag_tau_reg"], beta=case["beta"],
                                           freeze_encoder=case.get("freeze", False), verbose=False, debug_mode=False)
            halo = max(task.lag_idx, task.lag_ae_idx)
            n = X.shape[0] - halo
            idx = torch.arange(n)
            Xt, wt = task._traj[:n], task._weights[:n]
            if task.freeze_encoder:
                for p in model.encoder.parameters():
                    p.requires_grad = False
            eps = task._eps
            # ---- the loop body of train(), core.py:1066-1113
            ae = task.weighted_MSE_loss(Xt, task._traj[idx + task.lag_ae_idx] if task.lag_ae_idx > 0 else Xt, wt) \
                if task.alpha > eps else 0.0
            e0 = task.reg_enc_grad_loss(Xt, wt) if task.eta[0] > eps else 0.0
            e1 = task.reg_enc_norm_loss(Xt, wt) if task.eta[1] > eps else 0.0
            e2 = task.reg_enc_orthognal_loss(Xt, wt) if task.eta[2] > eps else 0.0
            if task.gamma[0] + task.gamma[1] > eps:
                Xl = task._traj[idx + task.lag_idx] if task.lag_idx > 0 else None
                wl = task._weights[idx + task.lag_idx] if task.lag_idx > 0 else None
                eig, g0, g1, cvec = task.reg_eigen_loss(Xt.clone(), wt, Xl, wl)
            else:
                g0 = g1 = 0.0
                eig, cvec = torch.zeros(K), np.arange(K)
            loss = task.alpha * ae + task.gamma[0] * g0 + task.gamma[1] * g1 + task.eta[0] * e0 + task.eta[1] * e1 + task.eta[2] * e2
            loss.backward()

            def grads(seq):
                return [p.grad.detach().double().numpy() if p.grad is not None else np.zeros(tuple(p.shape)) for p in seq.parameters()]
            out = dict(loss=float(loss), ae=float(ae), g0=float(g0), g1=float(g1), e0=float(e0), e1=float(e1), e2=float(e2),
                       eig=np.asarray(eig.detach().double().numpy()), cvec=np.asarray(cvec, dtype=np.int64),
                       genc=grads(model.encoder), gdec=grads(model.decoder), greg=[grads(model.reg[i]) for i in range(K)])
        return out
    finally:
        torch.set_default_dtype(torch.float32)


def save_case(name, case):
    r32, g64 = run_case(case, torch.float32), run_case(case, torch.float64)
    d = dict(X=case["X"], w=case["w"], K=case["K"], e_dims=np.asarray(case["e_dims"]), d_dims=np.asarray(case["d_dims"]),
             r_dims=np.asarray(case["r_dims"]), eig_w=np.asarray(case["eig_w"], dtype=np.float64), alpha=case["alpha"],
             gamma=np.asarray(case["gamma"], dtype=np.float64), eta=np.asarray(case["eta"], dtype=np.float64),
             lag_tau_ae=case["lag_tau_ae"], lag_tau_reg=case["lag_tau_reg"], beta=case["beta"], dt=case["dt"],
             freeze=case.get("freeze", False), pp_kind=case["pp"]["kind"])
    pp = case["pp"]
    if pp["kind"] == "mol":
        d["ref"] = np.asarray(pp["ref"], dtype=np.float64)
        d["align_idx"] = np.asarray(pp["align_idx"], dtype=np.int64)
    for j, p in enumerate(case["enc"]):
        d[f"enc_{j}"], d[f"g32_enc_{j}"], d[f"g64_enc_{j}"] = p, r32["genc"][j], g64["genc"][j]
    for j, p in enumerate(case["dec"]):
        d[f"dec_{j}"], d[f"g32_dec_{j}"], d[f"g64_dec_{j}"] = p, r32["gdec"][j], g64["gdec"][j]
    for i in range(case["K"]):
        for j, p in enumerate(case["reg"][i]):
            d[f"reg_{i}_{j}"], d[f"g32_reg_{i}_{j}"], d[f"g64_reg_{i}_{j}"] = p, r32["greg"][i][j], g64["greg"][i][j]
    for tag, r in (("r32", r32), ("g64", g64)):
        for key in ("loss", "ae", "g0", "g1", "e0", "e1", "e2", "eig", "cvec"):
            d[f"{tag}_{key}"] = r[key]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: loss32={r32['loss']:.8g} loss64={g64['loss']:.12g} parts64={[g64[k] for k in ('ae', 'g0', 'g1', 'e0', 'e1', 'e2')]} "
          f"eig64={g64['eig']} cvec={g64['cvec']}")


def save_train(name):
    """Whole reference RegAutoEncoderTask.train() run: examples/2d/2d.ipynb:716-721 (time-lagged autoencoder + transfer-operator
    regulariser) on a short random walk."""
    core, nn, _ = ref_import.load()
    X = np.cumsum(np.random.default_rng(21).normal(scale=0.15, size=(900, 2)), 0).astype(np.float32)
    X -= X.mean(0)
    w = ref_torch.boltzmann_weights(900, seed=21)
    torch.manual_seed(13)
    model = nn.RegAutoEncoder([2, 20, 20, 20, 1], [1, 20, 20, 2], [1, 20, 20, 1], 1)
    init = [p.detach().numpy().copy() for p in model.parameters()]
    traj = ref_import.FakeTrajectory(X.astype(np.float64), w.astype(np.float64), dt=0.1)
    with tempfile.TemporaryDirectory() as tmp:
        task = core.RegAutoEncoderTask(traj, torch.nn.Identity(), model, tmp, [1.0], gamma=[1, 20], eta=[0, 0, 0], lag_tau_ae=0.1,
                                       lag_tau_reg=0.1, learning_rate=0.005, test_ratio=0.2, batch_size=240, num_epochs=3,
                                       save_model_every_step=0, verbose=False, debug_mode=False)
        np.random.seed(79)
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


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(4052)

    # --- R1: every term switched on, generator regulariser, encoded dimension 2 (orthogonality term), K = 2, 2-d identity
    e, dd, r = [2, 12, 12, 2], [2, 12, 2], [2, 10, 10, 1]
    save_case("regae_2d_generator", dict(
        X=ring_2d(700, 41), w=ref_torch.boltzmann_weights(700, seed=41), K=2, e_dims=e, d_dims=dd, r_dims=r,
        enc=_mlp_params(e, rng, 2.0), dec=_mlp_params(dd, rng, 2.0), reg=[_mlp_params(r, rng, 2.5) for _ in range(2)],
        eig_w=[1.0, 0.5], alpha=1.5, gamma=[1.0, 20.0], eta=[0.3, 0.5, 0.7], lag_tau_ae=0, lag_tau_reg=0, beta=1.3, dt=0.1,
        pp=dict(kind="identity")))

    # --- R2: the notebook's configuration (2d.ipynb:716-721): time-lagged reconstruction + transfer-operator regulariser
    e, dd, r = [2, 20, 20, 20, 1], [1, 20, 20, 2], [1, 20, 20, 1]
    Xw = np.cumsum(np.random.default_rng(42).normal(scale=0.1, size=(640, 2)), 0).astype(np.float32)
    save_case("regae_2d_lagged", dict(
        X=Xw, w=ref_torch.boltzmann_weights(640, seed=42), K=1, e_dims=e, d_dims=dd, r_dims=r,
        enc=_mlp_params(e, rng, 2.0), dec=_mlp_params(dd, rng, 2.0), reg=[_mlp_params(r, rng, 2.5)],
        eig_w=[1.0], alpha=1.0, gamma=[1.0, 20.0], eta=[0.0, 0.0, 0.0], lag_tau_ae=0.1, lag_tau_reg=0.2, beta=1.0, dt=0.1,
        pp=dict(kind="identity")))

    # --- R3: K = 2 transfer operator with sorting, different lags, encoder frozen, encoder penalties on
    e, dd, r = [2, 10, 2], [2, 10, 2], [2, 8, 1]
    save_case("regae_2d_lagged_k2_frozen", dict(
        X=Xw[:500], w=ref_torch.boltzmann_weights(500, seed=43), K=2, e_dims=e, d_dims=dd, r_dims=r,
        enc=_mlp_params(e, rng, 2.0), dec=_mlp_params(dd, rng, 2.0), reg=[_mlp_params(r, rng, 2.5) for _ in range(2)],
        eig_w=[1.0, 0.6], alpha=2.0, gamma=[0.7, 5.0], eta=[0.0, 0.4, 0.2], lag_tau_ae=0.75, lag_tau_reg=0.25, beta=1.0, dt=0.25,
        freeze=True, pp=dict(kind="identity")))

    # --- R4: molecular frames: 22 atoms aligned on all atoms, position features (d_r = 66 = tot_dim), generator regulariser
    base = ref_torch.DIPEPTIDE_NM * 10.0
    Xm = ref_torch.synth_frames(base, 384, seed=44)
    e, dd, r = [66, 20, 20, 2], [2, 16, 66], [2, 12, 1]
    save_case("regae_dipep_generator", dict(
        X=Xm, w=ref_torch.boltzmann_weights(384, seed=44), K=2, e_dims=e, d_dims=dd, r_dims=r,
        enc=_mlp_params(e, rng, 1.0), dec=_mlp_params(dd, rng, 1.0), reg=[_mlp_params(r, rng, 2.0) for _ in range(2)],
        eig_w=[1.0, 0.5], alpha=1.0, gamma=[1.0, 10.0], eta=[0.2, 0.3, 0.4], lag_tau_ae=0, lag_tau_reg=0, beta=2.0, dt=1.0,
        pp=dict(kind="mol", ref=base, align_idx=list(range(22)), features=[("position", list(range(22)))])))

    save_train("train_regae_2d")


if __name__ == "__main__":
    main()
