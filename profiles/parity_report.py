"""Achieved parity errors of the CUDA step against the fp64 closed form (oracle/closed_form.py) at BASELINE batch sizes, and
against the reference's golden vectors -- the numbers the test envelopes are set from ("parity holes" 1b).

    python profiles/parity_report.py [--frames 1048576] [--out gpurun_out/r02_parity_errors.json]

For every workload: |loss - gold| / |gold|, max_i |eig_i - gold_i| / |gold_i|, max over parameter tensors of the rel-L2 gradient
error (tensors whose exact gradient is zero -- the last-layer biases of the generator loss -- excluded), and cvec equality.
The closed form is evaluated in chunks (two passes, fp64 numpy) so that 2^20 frames fit the host.  TEST INFRASTRUCTURE: the
oracle is the checker here, never the thing measured."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "colvars-finder_b200"))

import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import closed_form as cf, ref_torch  # noqa: E402
from oracle.ref_import import FakeTrajectory  # noqa: E402
from tests import _cases as C  # noqa: E402
from colvarsfinder import core, nn, utils  # noqa: E402

DEV = torch.device("cuda", 0)
try:      # small fp64 products: OpenBLAS with a thread per core only spins
    import threadpoolctl
    _BLAS_LIMIT = threadpoolctl.threadpool_limits(limits=4, user_api="blas")
except Exception:
    pass


def chunked_closed_form(X, w, nets, pp, alpha, eig_w):
    return cf.eigen_loss_and_grads_chunked(X, w, nets, pp, alpha, eig_w)


def eigen_errors(task, model, X, w, nets, ppo, alpha, eig_w):
    out = task.loss_func(task._traj, task._weights, None, None)
    out[0].backward()
    torch.cuda.synchronize()
    t0 = time.time()
    comb, g64 = chunked_closed_form(X, w, nets, ppo, alpha, eig_w)
    loss, eig = float(out[0]), out[1].cpu().numpy()
    gerr = []
    scale = max(np.abs(t).max() for n in g64 for t in n)
    for i, f in enumerate(model.eigen_funcs):
        for p, t in zip(f.parameters(), g64[i]):
            if np.abs(t).max() < 1e-9 * scale:
                continue
            gerr.append(C.rel_l2(p.grad.cpu().numpy(), t))
    return dict(frames=int(len(X)), loss=loss, loss_gold=float(comb["loss"]), loss_rel_err=abs(loss - comb["loss"]) / abs(comb["loss"]),
                eig=[float(v) for v in eig], eig_gold=[float(v) for v in comb["eig"]],
                eig_rel_err_max=float(np.max(np.abs(eig - comb["eig"]) / np.abs(comb["eig"]))), grad_rel_l2_max=float(max(gerr)),
                grad_rel_l2_median=float(np.median(gerr)), cvec_equal=list(out[4].cpu().numpy()) == list(comb["cvec"]),
                oracle_seconds=round(time.time() - t0, 1))


def run_c3(n, tmp):
    base = ref_torch.DIPEPTIDE_NM * 10.0
    X = ref_torch.synth_frames(base, n, seed=2026)
    w = ref_torch.boltzmann_weights(n, seed=2026)
    torch.manual_seed(2026)
    dims, k, eig_w = [66, 20, 20, 20, 1], 3, [1.0, 0.6, 0.3]
    model = nn.EigenFunctions(dims, k)
    nets = [[p.detach().numpy().copy() for p in f.parameters()] for f in model.eigen_funcs]
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), utils.Align(base, list(range(22))), model, tmp, 20.0, eig_w,
                                  k=k, device=DEV, verbose=False, debug_mode=False)
    return eigen_errors(task, model, X, w, nets, cf.Preproc(align_idx=list(range(22)), ref=base), 20.0, eig_w)


def run_c4(n, tmp):
    import bench_data as bd
    base = bd.chain_structure(166, seed=2026)
    feats, align = bd.c4_features()
    X = ref_torch.synth_frames(base, n, seed=2027)
    w = ref_torch.boltzmann_weights(n, seed=2027)
    torch.manual_seed(2027)
    dims, k, eig_w = [81, 20, 20, 20, 1], 3, [1.0, 0.6, 0.3]
    model = nn.EigenFunctions(dims, k)
    nets = [[p.detach().numpy().copy() for p in f.parameters()] for f in model.eigen_funcs]
    pp = utils.Preprocessing(utils.Align(base[align], align), utils.FeatureMap(feats))
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), pp, model, tmp, 20.0, eig_w, k=k, device=DEV, verbose=False,
                                  debug_mode=False)
    return eigen_errors(task, model, X, w, nets, cf.Preproc(align_idx=align, ref=base[align], feats=feats), 20.0, eig_w)


def run_c1(n, tmp):
    rng = np.random.default_rng(30)
    th, r = rng.uniform(-np.pi, np.pi, n), rng.normal(1.0, 0.25, n)
    X = np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
    w = np.ones(n, np.float32)
    torch.manual_seed(30)
    model = nn.EigenFunctions([2, 20, 20, 20, 1], 1)
    nets = [[p.detach().numpy().copy() for p in f.parameters()] for f in model.eigen_funcs]
    task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), torch.nn.Identity(), model, tmp, 20.0, [1.0], k=1, device=DEV,
                                  verbose=False, debug_mode=False)
    return eigen_errors(task, model, X, w, nets, cf.Preproc(identity=True), 20.0, [1.0])


def run_c2(n, tmp):
    base = ref_torch.DIPEPTIDE_NM * 10.0
    X = ref_torch.synth_frames(base, n, seed=2028)
    w = np.ones(n, np.float32)
    torch.manual_seed(2028)
    model = nn.AutoEncoder([66, 20, 20, 20, 2], [2, 10, 10, 66])
    enc = [p.detach().numpy().copy() for p in model.encoder.parameters()]
    dec = [p.detach().numpy().copy() for p in model.decoder.parameters()]
    task = core.AutoEncoderTask(FakeTrajectory(X, w.astype(np.float64)), utils.Align(base, list(range(22))), model, tmp, device=DEV,
                                verbose=False, debug_mode=False)
    loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
    loss.backward()
    F = task._feature_traj.cpu().numpy()
    y64 = cf.Preproc(align_idx=list(range(22)), ref=base).prepare(X[:65536].astype(np.float64))["r"]
    lo, genc, gdec = cf.ae_loss_and_grads(F, w, enc, dec)
    got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
    errs = [C.rel_l2(a, b) for a, b in zip(got, genc + gdec)]
    return dict(frames=int(n), loss=float(loss), loss_gold=float(lo), loss_rel_err=abs(float(loss) - lo) / abs(lo),
                grad_rel_l2_max=float(max(errs)), aligned_coordinates_max_abs_err_A=float(np.abs(F[:65536] - y64).max()))


def golden_errors(tmp):
    """ours vs the reference's fp64 run next to the reference's own fp32 run vs its fp64 run (tests/golden/*.npz)."""
    out = {}
    for name in C.EIGEN_GENERATOR_CASES:
        c = C.eigen_case(name)
        model = nn.EigenFunctions(c["layer_dims"], c["k"])
        with torch.no_grad():
            for i in range(c["k"]):
                for p, v in zip(model.eigen_funcs[i].parameters(), c["params"][i]):
                    p.copy_(torch.as_tensor(v))
        if c["pp_kind"] == "identity":
            pp = torch.nn.Identity()
        else:
            al = utils.Align(c["ref"], c["align_idx"]) if c["align_idx"] is not None else None
            fm = utils.FeatureMap(c["features"]) if c["features"] is not None else None
            pp = utils.Preprocessing(al, fm)
        diag = None if c["diag_coeff"] is None else torch.as_tensor(c["diag_coeff"])
        task = core.EigenFunctionTask(FakeTrajectory(c["X"], c["w"].astype(np.float64)), pp, model, tmp, c["alpha"], c["eig_w"],
                                      diag_coeff=diag, beta=c["beta"], sort_eigvals_in_training=c["sort"], k=c["k"], device=DEV,
                                      verbose=False, debug_mode=False)
        o = task.loss_func(task._traj, task._weights, None, None)
        o[0].backward()
        ours, ref = [], []
        for i, f in enumerate(model.eigen_funcs):
            for j, p in enumerate(f.parameters()):
                if np.abs(c["g64"][i][j]).max() < 1e-12:
                    continue
                ours.append(C.rel_l2(p.grad.cpu().numpy(), c["g64"][i][j]))
                ref.append(C.rel_l2(c["g32"][i][j], c["g64"][i][j]))
        gl = float(c["g64_loss"])
        out[name] = dict(loss_rel_err=abs(float(o[0]) - gl) / abs(gl), ref32_loss_rel_err=abs(float(c["r32_loss"]) - gl) / abs(gl),
                         eig_rel_err_max=float(np.max(np.abs(o[1].cpu().numpy() - c["g64_eig"]) / np.abs(c["g64_eig"]))),
                         ref32_eig_rel_err_max=float(np.max(np.abs(c["r32_eig"] - c["g64_eig"]) / np.abs(c["g64_eig"]))),
                         grad_rel_l2_max=float(max(ours)), ref32_grad_rel_l2_max=float(max(ref)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1 << 20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity_errors.json"))
    a = ap.parse_args()
    res = {"note": "errors of the CUDA step against the fp64 closed form on the same fp32 inputs; golden_cases: against the "
                   "reference's own fp64 run, next to the reference's fp32 run", "source_hash": g.library_hash()}
    with tempfile.TemporaryDirectory() as tmp:
        res["golden_cases"] = golden_errors(tmp)
        for name, fn, n in (("C1", run_c1, a.frames), ("C2", run_c2, a.frames), ("C3", run_c3, a.frames), ("C4", run_c4, a.frames // 4)):
            t0 = time.time()
            res[name] = fn(n, tmp)
            print(name, json.dumps(res[name]), f"({time.time() - t0:.0f} s)", flush=True)
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
