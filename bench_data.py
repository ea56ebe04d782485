"""Synthetic workloads of bench.py -- data only, no arithmetic of the training step.

Kept apart from ``oracle/`` so that the GPU arm of bench.py never imports the checker: base structures, the frame / weight
generators (on the device for the GPU arm) and a duck-typed trajectory holder (the tasks only read ``.trajectory``,
``.weights``, ``.dt`` -- reference core.py:329,343-346).
"""
from __future__ import annotations

import numpy as np
import torch

# 22 atoms of alanine dipeptide, nm (reference examples/dipeptide/top.gro:3-24); x10 -> Angstrom
DIPEPTIDE_NM = np.array([
    [0.200, 0.100, -0.000], [0.200, 0.209, 0.000], [0.149, 0.245, 0.089], [0.149, 0.245, -0.089],
    [0.343, 0.264, -0.000], [0.439, 0.188, -0.000], [0.356, 0.397, -0.000], [0.273, 0.456, -0.000],
    [0.485, 0.461, -0.000], [0.541, 0.432, 0.089], [0.566, 0.422, -0.123], [0.512, 0.452, -0.213],
    [0.663, 0.472, -0.121], [0.581, 0.314, -0.124], [0.471, 0.613, 0.000], [0.360, 0.665, 0.000],
    [0.585, 0.683, 0.000], [0.674, 0.636, -0.000], [0.585, 0.828, 0.000], [0.482, 0.865, 0.000],
    [0.636, 0.865, 0.089], [0.636, 0.865, -0.089]])


class SyntheticTrajectory:
    """What the task constructors read from utils.WeightedTrajectory."""

    def __init__(self, trajectory, weights, dt=1.0):
        self.trajectory, self.weights, self.dt = trajectory, weights, dt
        self.n_frames = trajectory.shape[0]


def chain_structure(n_atoms: int, seed: int = 2026, bond: float = 1.5) -> np.ndarray:
    """Random-walk chain with fixed bond length, centred (C4's 166-atom base structure)."""
    rng = np.random.default_rng(seed)
    steps = rng.normal(size=(n_atoms, 3))
    steps *= bond / np.linalg.norm(steps, axis=1, keepdims=True)
    pos = np.cumsum(steps, axis=0)
    return pos - pos.mean(0)


def c4_features():
    """45 pair distances among 10 designated atoms + 18 backbone dihedrals of the 166-atom chain, and the 40 alignment atoms."""
    sel = list(range(5, 166, 16))[:10]
    feats = [("bond", [a, b]) for i, a in enumerate(sel) for b in sel[i + 1:]]
    feats += [("dihedral", [s, s + 1, s + 2, s + 3]) for s in range(10, 10 + 18 * 8, 8)]
    return feats, list(range(0, 160, 4))


def frames(base, n, device, seed):
    """frame = base Q + t + eps: Haar rotation from a random unit quaternion, t ~ N(0,5^2) A, eps ~ N(0,0.3^2) A; float32."""
    g = torch.Generator(device=device).manual_seed(seed)
    base_t = torch.as_tensor(np.asarray(base), dtype=torch.float32, device=device)
    out = torch.empty(n, base_t.shape[0], 3, dtype=torch.float32, device=device)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        q = torch.randn(m, 4, generator=g, device=device)
        q = q / q.norm(dim=1, keepdim=True)
        w, x, y, z = q.unbind(1)
        Q = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(m, 3, 3)
        fr = torch.einsum("ni,bij->bnj", base_t, Q)
        fr += 5.0 * torch.randn(m, 1, 3, generator=g, device=device)
        fr += 0.3 * torch.randn(m, base_t.shape[0], 3, generator=g, device=device)
        out[s:s + m] = fr
    return out


def boltzmann_weights(n, device, seed, dbeta=0.5):
    """w = exp(-dbeta (E - mean E)), E ~ N(0,1), normalised to mean 1 (formula of reference utils.py:411-412,145)."""
    g = torch.Generator(device=device).manual_seed(seed + 1)
    E = torch.randn(n, generator=g, device=device)
    w = torch.exp(-dbeta * (E - E.mean()))
    return (w / w.mean()).contiguous()


def ring_2d(n, device, seed):
    """C1: points on a noisy ring, theta ~ U(-pi,pi), r ~ N(1,0.25^2) (examples/2d/2d.ipynb:148)."""
    g = torch.Generator(device=device).manual_seed(seed)
    th = (torch.rand(n, generator=g, device=device) * 2 - 1) * np.pi
    r = 1.0 + 0.25 * torch.randn(n, generator=g, device=device)
    return torch.stack([r * torch.cos(th), r * torch.sin(th)], 1).contiguous()
