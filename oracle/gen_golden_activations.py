"""Golden vectors for activations other than Tanh (reference nn.py:29-59 takes any module) from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden_activations.py

Per activation (Sigmoid, Softplus, ELU, ReLU): the generator loss of EigenFunctionTask on 2-d data (k = 2) and on aligned dipeptide
frames with a feature map (k = 2), and the weighted MSE of AutoEncoderTask, each with its backward pass, in float32 ("r32") and in
float64 on the same fp32-rounded inputs ("g64").
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

ACTS = {"sigmoid": torch.nn.Sigmoid, "softplus": torch.nn.Softplus, "elu": torch.nn.ELU, "relu": torch.nn.ReLU}


def run_eigen(case, act, dtype):
    core, nn, _ = ref_import.load()
    torch.set_default_dtype(dtype)
    try:
        k = case["k"]
        model = nn.EigenFunctions(case["layer_dims"], k, ACTS[act]())
        for i in range(k):
            _load_seq(model.eigen_funcs[i], case["params"][i], dtype)
        traj = ref_import.FakeTrajectory(case["X"], case["w"].astype(np.float64), dt=1.0)
        with tempfile.TemporaryDirectory() as tmp:
            task = core.EigenFunctionTask(traj, build_pp(case["pp"]), model, tmp, case["alpha"], case["eig_w"], k=k, verbose=False,
                                          debug_mode=False)
            Xt = torch.as_tensor(case["X"]).to(dtype).requires_grad_()
            loss, eig, obj, pen, cvec = task.loss_func(Xt, torch.as_tensor(case["w"]).to(dtype), None, None)
            loss.backward()
            grads = [[p.grad.detach().double().numpy() if p.grad is not None else np.zeros(tuple(p.shape))
                      for p in model.eigen_funcs[i].parameters()] for i in range(k)]
        return dict(loss=float(loss), eig=eig.detach().double().numpy(), cvec=np.asarray(cvec, dtype=np.int64), grads=grads)
    finally:
        torch.set_default_dtype(torch.float32)


def run_ae(case, act, dtype):
    core, nn, _ = ref_import.load()
    torch.set_default_dtype(dtype)
    try:
        model = nn.AutoEncoder(case["e_dims"], case["d_dims"], ACTS[act]())
        _load_seq(model.encoder, case["enc"], dtype)
        _load_seq(model.decoder, case["dec"], dtype)
        traj = ref_import.FakeTrajectory(case["F"], case["w"].astype(np.float64))
        with tempfile.TemporaryDirectory() as tmp:
            task = core.AutoEncoderTask(traj, torch.nn.Identity(), model, tmp, verbose=False, debug_mode=False)
            loss = task.weighted_MSE_loss(torch.as_tensor(case["F"]).to(dtype), torch.as_tensor(case["w"]).to(dtype))
            loss.backward()
            g = [p.grad.detach().double().numpy() for p in list(model.encoder.parameters()) + list(model.decoder.parameters())]
        return dict(loss=float(loss), grads=g)
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    rng = np.random.default_rng(777)
    base = ref_torch.DIPEPTIDE_NM * 10.0
    heavy = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]
    feats = [("bond", [1, 4]), ("angle", [4, 6, 8]), ("dihedral", [4, 6, 8, 14]), ("position", [8, 10]), ("dihedral", [6, 8, 14, 16])]
    e2d = dict(X=ring_2d(600, 51), w=ref_torch.boltzmann_weights(600, seed=51), k=2, layer_dims=[2, 12, 12, 1],
               params=[_mlp_params([2, 12, 12, 1], rng, 2.5) for _ in range(2)], alpha=10.0, eig_w=[1.0, 0.5], pp=dict(kind="identity"))
    dims = [1 + 1 + 2 + 6 + 2, 14, 14, 1]
    emol = dict(X=ref_torch.synth_frames(base, 400, seed=52), w=ref_torch.boltzmann_weights(400, seed=52), k=2, layer_dims=dims,
                params=[_mlp_params(dims, rng, 1.5) for _ in range(2)], alpha=10.0, eig_w=[1.0, 0.4],
                pp=dict(kind="mol", ref=base[heavy], align_idx=heavy, features=feats))
    ae = dict(F=ring_2d(500, 53), w=ref_torch.boltzmann_weights(500, seed=53), e_dims=[2, 16, 16, 1], d_dims=[1, 16, 2],
              enc=_mlp_params([2, 16, 16, 1], rng, 2.0), dec=_mlp_params([1, 16, 2], rng, 2.0))
    d = dict(e2d_X=e2d["X"], e2d_w=e2d["w"], emol_X=emol["X"], emol_w=emol["w"], ae_F=ae["F"], ae_w=ae["w"], emol_ref=base[heavy],
             emol_align=np.asarray(heavy), feat_types=np.asarray([t for t, _ in feats]),
             feat_atoms=np.asarray([list(a) + [-1] * (4 - len(a)) for _, a in feats], dtype=np.int64))
    for tag, case in (("e2d", e2d), ("emol", emol)):
        for i in range(case["k"]):
            for j, p in enumerate(case["params"][i]):
                d[f"{tag}_p_{i}_{j}"] = p
    for j, p in enumerate(ae["enc"] + ae["dec"]):
        d[f"ae_p_{j}"] = p
    for act in ACTS:
        for tag, case in (("e2d", e2d), ("emol", emol)):
            r32, g64 = run_eigen(case, act, torch.float32), run_eigen(case, act, torch.float64)
            for name, r in (("r32", r32), ("g64", g64)):
                d[f"{act}_{tag}_{name}_loss"], d[f"{act}_{tag}_{name}_eig"], d[f"{act}_{tag}_{name}_cvec"] = r["loss"], r["eig"], r["cvec"]
                for i in range(case["k"]):
                    for j, g in enumerate(r["grads"][i]):
                        d[f"{act}_{tag}_{name}_g_{i}_{j}"] = g
            print(act, tag, "loss32", r32["loss"], "loss64", g64["loss"], "eig64", g64["eig"], "cvec", g64["cvec"])
        r32, g64 = run_ae(ae, act, torch.float32), run_ae(ae, act, torch.float64)
        for name, r in (("r32", r32), ("g64", g64)):
            d[f"{act}_ae_{name}_loss"] = r["loss"]
            for j, g in enumerate(r["grads"]):
                d[f"{act}_ae_{name}_g_{j}"] = g
        print(act, "ae loss32", r32["loss"], "loss64", g64["loss"])
    np.savez_compressed(os.path.join(OUT, "activations.npz"), **d)


if __name__ == "__main__":
    main()
