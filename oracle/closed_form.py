"""Closed-form (autograd-free) numpy/float64 restatement of the training step.  TEST INFRASTRUCTURE ONLY.

``oracle/ref_torch.py`` restates the reference with autograd (double backward through an SVD).  This
module restates the SAME quantities the way the CUDA kernels compute them, so that every formula the
kernels rely on is pinned on the CPU against the reference's golden vectors before any GPU code runs:

* Kabsch rotation + the closed-form transpose-Jacobian / Jacobian of the alignment
  (replaces autograd through torch.linalg.svd, core.py:424);
* analytic gradients of bond / angle / dihedral features;
* the Dirichlet-energy parameter gradient from ONE tangent direction per frame and eigenfunction
  (replaces the double backward of core.py:517);
* the two-pass split of EigenFunctionTask.loss_func (core.py:387-457): pass 1 = batch sums
  (S0, S1_i, S2_ij, SD_i), ``combine`` = loss / eigenvalues / cvec and the per-frame seed coefficients,
  pass 2 = parameter gradients;
* AutoEncoderTask.weighted_MSE_loss (core.py:652-666) forward + backward.

Nothing under colvars-finder_b200/ imports this file.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- alignment
def kabsch(x, align_idx, ref):
    """x [B,N,3]; ref [n_a,3] (centred here).  Returns y, R [B,3,3], c [B,1,3], Kinv [B,3,3]."""
    ref = ref - ref.mean(0, keepdims=True)
    xa = x[:, align_idx]
    c = xa.mean(1, keepdims=True)
    H = np.einsum("bna,nc->bac", xa - c, ref)
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    D = np.ones((x.shape[0], 3))
    D[:, 2] = d
    R = (U * D[:, None, :]) @ Vt
    y = (x - c) @ R
    M = np.einsum("bna,nc->bac", y[:, align_idx], ref)
    M = 0.5 * (M + M.transpose(0, 2, 1))
    K = np.trace(M, axis1=1, axis2=2)[:, None, None] * np.eye(3) - M
    return y, R, c, np.linalg.inv(K), ref


def align_vjp(G, y, R, Kinv, align_idx, ref):
    """G = df/dy [B,N,3]  ->  df/dx [B,N,3]   (SURVEY 7.3-A)."""
    tau = np.cross(G, y).sum(1)
    q = np.einsum("bij,bj->bi", Kinv, tau)
    Gy = G.copy()
    Gy[:, align_idx] -= G.sum(1, keepdims=True) / len(align_idx)
    Gy[:, align_idx] -= np.cross(ref[None], q[:, None, :])
    return Gy @ R.transpose(0, 2, 1)


def align_jvp(dx, y, R, Kinv, align_idx, ref):
    """dx [B,N,3] -> dy [B,N,3]."""
    dc = dx[:, align_idx].mean(1, keepdims=True)
    u = (dx - dc) @ R
    om = np.einsum("bij,bj->bi", Kinv, np.cross(u[:, align_idx], ref[None]).sum(1))
    return u + np.cross(om[:, None, :], y)


# ----------------------------------------------------------------------------- features
def _unit(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return v / n, n


def feature_stencil(y, ftype, idx):
    """Values and per-atom gradient stencils of one feature.
    Returns (vals [B,m], grads list over outputs of [B,n_atoms_in_feature,3])."""
    if ftype == "position":
        B = y.shape[0]
        vals = y[:, idx].reshape(B, -1)
        grads = []
        for a in range(len(idx)):
            for cdim in range(3):
                g = np.zeros((B, len(idx), 3))
                g[:, a, cdim] = 1.0
                grads.append(g)
        return vals, grads
    if ftype == "bond":
        i, j = idx
        dh, n = _unit(y[:, j] - y[:, i])
        return n, [np.stack([-dh, dh], 1)]
    if ftype == "angle":
        i, j, k = idx
        ah, na = _unit(y[:, i] - y[:, j])
        bh, nb = _unit(y[:, k] - y[:, j])
        cs = (ah * bh).sum(-1, keepdims=True)
        ga = (bh - cs * ah) / na
        gb = (ah - cs * bh) / nb
        return cs, [np.stack([ga, -(ga + gb), gb], 1)]
    if ftype == "dihedral":
        i, j, k, l = idx
        r12, r23, r34 = y[:, j] - y[:, i], y[:, k] - y[:, j], y[:, l] - y[:, k]
        n1, n2 = np.cross(r12, r23), np.cross(r23, r34)
        n1s, n2s = (n1 * n1).sum(-1, keepdims=True), (n2 * n2).sum(-1, keepdims=True)
        l23 = np.linalg.norm(r23, axis=-1, keepdims=True)
        den = np.sqrt(n1s * n2s)
        cs = (n1 * n2).sum(-1, keepdims=True) / den
        sn = (n1 * r34).sum(-1, keepdims=True) * l23 / den
        # gradient of the angle phi = atan2(sn, cs)
        gi = -l23 / n1s * n1
        gl = l23 / n2s * n2
        p = (r12 * r23).sum(-1, keepdims=True) / (l23 * l23)
        q = (r34 * r23).sum(-1, keepdims=True) / (l23 * l23)
        gj = -gi - p * gi + q * gl
        gk = -gl + p * gi - q * gl
        gphi = np.stack([gi, gj, gk, gl], 1)
        return np.concatenate([cs, sn], 1), [-sn[:, :, None] * gphi, cs[:, :, None] * gphi]
    raise ValueError(ftype)


def features_fwd(y, feats):
    return np.concatenate([feature_stencil(y, t, a)[0] for t, a in feats], 1)


def features_vjp(u, y, feats):
    """u [B,d_r] -> G [B,N,3]."""
    G = np.zeros_like(y)
    col = 0
    for t, a in feats:
        if t == "position":      # unit stencils: the gradient is u itself (same result as the generic loop, without 3|a| dense arrays)
            G[:, a] += u[:, col:col + 3 * len(a)].reshape(u.shape[0], len(a), 3)
            col += 3 * len(a)
            continue
        _, grads = feature_stencil(y, t, a)
        for g in grads:
            for s, atom in enumerate(a):
                G[:, atom] += u[:, col:col + 1] * g[:, s]
            col += 1
    return G


def features_jvp(dy, y, feats):
    out = []
    for t, a in feats:
        if t == "position":
            out.extend(dy[:, a].reshape(dy.shape[0], -1).T)
            continue
        _, grads = feature_stencil(y, t, a)
        for g in grads:
            out.append((g * dy[:, a]).sum((1, 2)))
    return np.stack(out, 1)


class Preproc:
    """r(x) with its transpose-Jacobian and Jacobian.  kind: identity | mol."""

    def __init__(self, align_idx=None, ref=None, feats=None, identity=False):
        self.identity = identity
        self.align_idx = None if align_idx is None else list(align_idx)
        self.ref = None if ref is None else np.asarray(ref, dtype=np.float64)
        self.feats = feats

    def prepare(self, x):
        if self.identity:
            return dict(r=x.reshape(x.shape[0], -1))
        st = dict()
        if self.align_idx is not None:
            y, R, c, Kinv, refc = kabsch(x, self.align_idx, self.ref)
            st.update(y=y, R=R, Kinv=Kinv, refc=refc)
        else:
            st.update(y=x)
        feats = self.feats if self.feats is not None else [("position", list(range(x.shape[1])))]
        st["feats"] = feats
        st["r"] = features_fwd(st["y"], feats)
        return st

    def vjp(self, st, u):
        """u = dg/dr [B,d_r] -> dg/dx flattened [B,3N]."""
        if self.identity:
            return u
        G = features_vjp(u, st["y"], st["feats"])
        if self.align_idx is not None:
            G = align_vjp(G, st["y"], st["R"], st["Kinv"], self.align_idx, st["refc"])
        return G.reshape(G.shape[0], -1)

    def jvp(self, st, dx):
        if self.identity:
            return dx
        dx = dx.reshape(st["y"].shape)
        if self.align_idx is not None:
            dx = align_jvp(dx, st["y"], st["R"], st["Kinv"], self.align_idx, st["refc"])
        return features_jvp(dx, st["y"], st["feats"])


# ----------------------------------------------------------------------------- small MLP, all passes
def mlp_forward(params, r, acts=None):
    """params [W1,b1,...]; acts[l] True where layer l is followed by tanh (default: all but last)."""
    L = len(params) // 2
    acts = [l < L - 1 for l in range(L)] if acts is None else acts
    a = [r]
    for l in range(L):
        z = a[-1] @ params[2 * l].T + params[2 * l + 1]
        a.append(np.tanh(z) if acts[l] else z)
    return a


def mlp_input_grad(params, a, acts=None):
    """Reverse sweep with seed 1 on the (scalar) output.  Returns u = dy/dr and gbar_l (adjoints of z_l)."""
    L = len(params) // 2
    acts = [l < L - 1 for l in range(L)] if acts is None else acts
    gbar = [None] * (L + 1)
    gbar[L] = np.ones_like(a[L])
    for l in range(L, 0, -1):
        p = gbar[l] @ params[2 * (l - 1)]
        if l - 1 >= 1:
            gbar[l - 1] = p * (1 - a[l - 1] ** 2) if acts[l - 2] else p
        else:
            u = p
    return u, gbar


def eigen_net_grads(params, a, gbar, v, seed_y):
    """Parameter gradient of  sum_f [ seed_y_f * y_f + ydot_f ]  where ydot is the tangent of y along v
    (v already carries the per-frame factor 2 w c_D).  Single reverse sweep over (primal, tangent)."""
    L = len(params) // 2
    # tangent forward
    T = [v]
    E = [None]
    for l in range(1, L):
        zd = T[-1] @ params[2 * (l - 1)].T
        T.append((1 - a[l] ** 2) * zd)
        E.append(-2 * a[l] * gbar[l] * zd)
    grads = [None] * (2 * L)
    s = seed_y[:, None]                      # adjoint of z_L
    for l in range(L, 0, -1):
        q = gbar[l]                          # adjoint of zdot_l (per unit seed; factor lives in T)
        grads[2 * (l - 1)] = s.T @ a[l - 1] + q.T @ T[l - 1]
        grads[2 * (l - 1) + 1] = s.sum(0)
        if l - 1 >= 1:
            s = (s @ params[2 * (l - 1)]) * (1 - a[l - 1] ** 2) + E[l - 1]
    return grads


# ----------------------------------------------------------------------------- eigenfunction loss, two passes
def eigen_stats(X, w, nets, pp, diag_coeff):
    """Pass 1.  Returns dict of batch sums (float64) and the per-frame state reused by pass 2."""
    k = len(nets)
    st = pp.prepare(X)
    r = st["r"]
    A, GB, U, V, Y, Dl = [], [], [], [], [], []
    for p in nets:
        a = mlp_forward(p, r)
        u, gbar = mlp_input_grad(p, a)
        gx = pp.vjp(st, u)                         # grad wrt raw coordinates, [B, tot_dim]
        h = gx * diag_coeff[None, :]
        Dl.append((gx * h).sum(1))
        V.append(pp.jvp(st, h))
        A.append(a), GB.append(gbar), U.append(u), Y.append(a[-1][:, 0])
    Y = np.stack(Y, 1)
    S = dict(S0=w.sum(), S1=(w[:, None] * Y).sum(0), S2=np.einsum("b,bi,bj->ij", w, Y, Y),
             SD=np.array([(w * Dl[i]).sum() for i in range(k)]))
    return S, dict(A=A, GB=GB, V=V, Y=Y, r=r)


def eigen_combine(S, alpha, eig_w, beta=1.0, sort=True):
    """loss, eigenvalues, cvec + seed coefficients  (core.py:406-410,426-455)."""
    k = len(S["S1"])
    S0 = S["S0"]
    mean = S["S1"] / S0
    cov = S["S2"] / S0 - np.outer(mean, mean)
    var = np.diag(cov).copy()
    E = S["SD"] / (beta * S0)
    eig = E / var
    cvec = np.argsort(eig) if sort else np.arange(k)
    omega = np.zeros(k)
    omega[cvec] = np.asarray(eig_w, dtype=np.float64)
    obj = (omega * eig).sum()
    pen = ((var - 1) ** 2).sum() + sum(cov[i, j] ** 2 for i in range(k) for j in range(i + 1, k))
    loss = obj + alpha * pen
    cD = omega / (beta * S0 * var)
    cv = -omega * E / var ** 2 + 2 * alpha * (var - 1)
    C2 = 2 * alpha * cov / S0
    C2[np.arange(k), np.arange(k)] = 2 * cv / S0
    return dict(loss=loss, eig=eig[cvec], obj=obj, pen=pen, cvec=cvec, mean=mean, cD=cD, C2=C2)


def eigen_grads(w, nets, state, comb):
    """Pass 2: parameter gradients of the loss."""
    k = len(nets)
    Yc = state["Y"] - comb["mean"][None, :]
    out = []
    for i in range(k):
        seed = w * (Yc @ comb["C2"][i])
        v = state["V"][i] * (2 * w * comb["cD"][i])[:, None]
        out.append(eigen_net_grads(nets[i], state["A"][i], state["GB"][i], v, seed))
    return out


def eigen_loss_and_grads(X, w, nets, pp, alpha, eig_w, diag_coeff=None, beta=1.0, sort=True):
    X = np.asarray(X, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    nets = [[np.asarray(p, dtype=np.float64) for p in n] for n in nets]
    tot_dim = X[0].size
    a = np.ones(tot_dim) if diag_coeff is None else np.asarray(diag_coeff, dtype=np.float64)
    S, state = eigen_stats(X, w, nets, pp, a)
    comb = eigen_combine(S, alpha, eig_w, beta, sort)
    return comb, eigen_grads(w, nets, state, comb), S


def eigen_loss_and_grads_chunked(X, w, nets, pp, alpha, eig_w, diag_coeff=None, beta=1.0, sort=True, chunk=32768):
    """eigen_loss_and_grads for batches whose per-frame state does not fit the host: pass 1 over chunks (batch sums add),
    combine, pass 2 over chunks with pass 1 recomputed (gradient sums add).  Returns (comb, grads)."""
    nets = [[np.asarray(p, dtype=np.float64) for p in n] for n in nets]
    a = np.ones(np.asarray(X[0]).size) if diag_coeff is None else np.asarray(diag_coeff, dtype=np.float64)
    S = None
    for s in range(0, len(X), chunk):
        Sc, _ = eigen_stats(np.asarray(X[s:s + chunk], dtype=np.float64), np.asarray(w[s:s + chunk], dtype=np.float64), nets, pp, a)
        S = Sc if S is None else {key: S[key] + Sc[key] for key in S}
    comb = eigen_combine(S, alpha, eig_w, beta, sort)
    grads = None
    for s in range(0, len(X), chunk):
        wc = np.asarray(w[s:s + chunk], dtype=np.float64)
        _, st = eigen_stats(np.asarray(X[s:s + chunk], dtype=np.float64), wc, nets, pp, a)
        gc = eigen_grads(wc, nets, st, comb)
        grads = gc if grads is None else [[p + q for p, q in zip(gi, gj)] for gi, gj in zip(grads, gc)]
    return comb, grads


# ----------------------------------------------------------------------------- autoencoder loss
def ae_loss_and_grads(F, w, enc, dec):
    """weighted MSE (core.py:666) and its parameter gradients; encoder's last layer is linear."""
    F = np.asarray(F, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    params = [np.asarray(p, dtype=np.float64) for p in list(enc) + list(dec)]
    Le, Ld = len(enc) // 2, len(dec) // 2
    acts = [l < Le - 1 for l in range(Le)] + [l < Ld - 1 for l in range(Ld)]
    a = mlp_forward(params, F, acts)
    diff = a[-1] - F
    S0 = w.sum()
    loss = (w * (diff ** 2).sum(1)).sum() / S0
    s = 2 * w[:, None] * diff / S0
    L = Le + Ld
    grads = [None] * (2 * L)
    for l in range(L, 0, -1):
        grads[2 * (l - 1)] = s.T @ a[l - 1]
        grads[2 * (l - 1) + 1] = s.sum(0)
        if l - 1 >= 1:
            s = s @ params[2 * (l - 1)]
            if acts[l - 2]:
                s = s * (1 - a[l - 1] ** 2)
    return loss, grads[:2 * Le], grads[2 * Le:]
