"""The oracle against the reference's golden vectors (tests/golden, made by oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_torch
from tests import _cases as C


def _pp_torch(c):
    if c["pp_kind"] == "identity":
        return ref_torch.Preprocess()
    al = ref_torch.Align(c["ref"], c["align_idx"]) if c["align_idx"] is not None else None
    fm = ref_torch.FeatureMap(c["features"]) if c["features"] is not None else None
    return ref_torch.Preprocess(al, fm)


def _pp_cf(c):
    if c["pp_kind"] == "identity":
        return cf.Preproc(identity=True)
    return cf.Preproc(align_idx=c["align_idx"], ref=c["ref"], feats=c["features"])


@pytest.mark.parametrize("name", C.EIGEN_GENERATOR_CASES + ["eigen_2d_lag"])
def test_autograd_restatement_matches_reference(name):
    c = C.eigen_case(name)
    dt = torch.float64
    nets = [[torch.as_tensor(p).to(dt).requires_grad_() for p in n] for n in c["params"]]
    X = torch.as_tensor(c["X"]).to(dt)
    w = torch.as_tensor(c["w"]).to(dt)
    a = None if c["diag_coeff"] is None else torch.as_tensor(c["diag_coeff"]).to(dt)
    if c["lag_tau"] == 0:
        X.requires_grad_()
        out = ref_torch.eigen_loss(X, w, nets, _pp_torch(c), c["alpha"], c["eig_w"], a, c["beta"], c["sort"])
    else:
        lag = int(round(c["lag_tau"] / c["dt"]))
        out = ref_torch.eigen_loss(X[:-lag], w[:-lag], nets, _pp_torch(c), c["alpha"], c["eig_w"], sort=c["sort"],
                                   X_lagged=X[lag:], weight_lagged=w[lag:], lag_time=c["dt"] * lag)
    loss, eig, obj, pen, cvec = out
    loss.backward()
    assert abs(float(loss) - c["g64_loss"]) <= 1e-10 * abs(c["g64_loss"])
    np.testing.assert_allclose(eig.numpy(), c["g64_eig"], rtol=1e-10)
    assert abs(float(obj) - c["g64_obj"]) <= 1e-10 * abs(c["g64_obj"])
    assert abs(float(pen) - c["g64_pen"]) <= 1e-10 * abs(c["g64_pen"])
    assert list(cvec) == list(c["g64_cvec"])
    for i in range(c["k"]):
        for j, p in enumerate(nets[i]):
            g = np.zeros(p.shape) if p.grad is None else p.grad.numpy()
            assert C.rel_l2(g, c["g64"][i][j]) < 1e-9 or np.abs(c["g64"][i][j]).max() < 1e-12


@pytest.mark.parametrize("name", C.EIGEN_GENERATOR_CASES)
def test_closed_form_matches_reference(name):
    c = C.eigen_case(name)
    comb, grads, _ = cf.eigen_loss_and_grads(c["X"], c["w"], c["params"], _pp_cf(c), c["alpha"], c["eig_w"],
                                             c["diag_coeff"], c["beta"], c["sort"])
    assert abs(comb["loss"] - c["g64_loss"]) <= 1e-10 * abs(c["g64_loss"])
    np.testing.assert_allclose(comb["eig"], c["g64_eig"], rtol=1e-9)
    assert abs(comb["obj"] - c["g64_obj"]) <= 1e-9 * abs(c["g64_obj"])
    assert abs(comb["pen"] - c["g64_pen"]) <= 1e-9 * abs(c["g64_pen"])
    assert list(comb["cvec"]) == list(c["g64_cvec"])
    for i in range(c["k"]):
        for j in range(len(grads[i])):
            gold = c["g64"][i][j]
            if np.abs(gold).max() < 1e-12:      # last-layer bias: exactly zero in the generator loss
                assert np.abs(grads[i][j]).max() < 1e-9
            else:
                assert C.rel_l2(grads[i][j], gold) < 1e-8, (i, j)


@pytest.mark.parametrize("name", C.AE_CASES)
def test_ae_oracles_match_reference(name):
    c = C.ae_case(name)
    loss, genc, gdec = cf.ae_loss_and_grads(c["F"], c["w"], c["enc"], c["dec"])
    assert abs(loss - c["g64_loss"]) <= 1e-11 * abs(c["g64_loss"])
    for g, gold in zip(genc + gdec, c["g64_enc"] + c["g64_dec"]):
        assert C.rel_l2(g, gold) < 1e-9
    enc = [torch.as_tensor(p).double().requires_grad_() for p in c["enc"]]
    dec = [torch.as_tensor(p).double().requires_grad_() for p in c["dec"]]
    l2 = ref_torch.ae_loss(torch.as_tensor(c["F"]).double(), torch.as_tensor(c["w"]).double(), enc, dec)
    assert abs(float(l2) - c["g64_loss"]) <= 1e-11 * abs(c["g64_loss"])


def test_split_arithmetic_pinned_by_notebooks():
    # examples/2d/2d.ipynb:518-523 (5000 -> 4000/1000, 4 iters at bs 1000), :651-656 (4998 -> 3998/1000, 3 iters),
    # examples/dipeptide/main.ipynb:481-486 (150000 -> 120000/30000, 6 iters at bs 20000)
    for n, bs, ntr, nte, iters in [(5000, 1000, 4000, 1000, 4), (4998, 1000, 3998, 1000, 3),
                                   (150000, 20000, 120000, 30000, 6)]:
        tr, te = ref_torch.split_indices(n, 0.2)
        assert (len(tr), len(te)) == (ntr, nte)
        b, _ = ref_torch.batches(len(tr), bs)
        assert len(b) == iters


def test_align_properties():
    base = ref_torch.DIPEPTIDE_NM * 10
    X = ref_torch.synth_frames(base, 64, seed=1).astype(np.float64)
    y, R, c, Kinv, refc = cf.kabsch(X, list(range(22)), base)
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-12)
    al = ref_torch.Align(base, list(range(22)))
    np.testing.assert_allclose(al(torch.as_tensor(X)).numpy(), y, atol=1e-11)
    # rotating + translating the input does not change the aligned frame
    Q = np.linalg.qr(np.random.default_rng(0).normal(size=(3, 3)))[0]
    Q *= np.sign(np.linalg.det(Q))
    y2 = cf.kabsch(X @ Q + 3.0, list(range(22)), base)[0]
    np.testing.assert_allclose(y2, y, atol=1e-10)
