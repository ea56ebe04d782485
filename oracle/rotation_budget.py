"""How many Jacobi sweeps / Newton evaluations does cvf_rotation (csrc/cvf_math.cuh) need?  TEST INFRASTRUCTURE ONLY.

Compiles the header for the host with every (sweeps, evaluations) combination and reports the largest error of the fp64 rotation
against numpy's SVD Kabsch over easy frames (thermal noise), hard frames (large noise, small / nearly planar alignment subsets)
and reflection-prone ones.  The kernels need the rotation to ~1e-8 so that aligned coordinates hold 1e-6 A after the fp32
rounding of the output.

    python oracle/rotation_budget.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import closed_form as cf, ref_torch  # noqa: E402


def cases():
    base = ref_torch.DIPEPTIDE_NM * 10
    out = []
    for name, idx, noise, n in (("all atoms, noise 0.3", list(range(22)), 0.3, 20000), ("all atoms, noise 1.0", list(range(22)), 1.0, 20000),
                                ("4 atoms nearly planar", [1, 4, 6, 8], 0.5, 20000), ("3 atoms", [4, 6, 8], 0.5, 20000),
                                ("10 heavy atoms, noise 2.0", [1, 4, 5, 6, 8, 10, 14, 15, 16, 18], 2.0, 20000)):
        X = ref_torch.synth_frames(base, n, seed=len(name), noise_sd=noise).astype(np.float64)
        y, R, c, Kinv, refc = cf.kabsch(X, idx, base[idx])
        H = np.ascontiguousarray(np.einsum("bna,nc->bac", X[:, idx] - c, refc))
        out.append((name, H, R, Kinv))
    return out


def main():
    cs = cases()
    with tempfile.TemporaryDirectory() as tmp:
        for sweeps in (2, 3, 4):
            for evals in (2, 3):
                so = os.path.join(tmp, f"h_{sweeps}_{evals}.so")
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-DCVF_JACOBI_SWEEPS={sweeps}",
                                       f"-DCVF_NEWTON_EVALS={evals}", "-o", so, os.path.join(HERE, "host_math.cpp")])
                h = C.CDLL(so)
                row = []
                for name, H, R, Kinv in cs:
                    Rd = np.zeros((len(H), 9))
                    Kf = np.zeros((len(H), 6), np.float32)
                    h.host_rotation_d(H.ctypes.data_as(C.c_void_p), len(H), Rd.ctypes.data_as(C.c_void_p), Kf.ctypes.data_as(C.c_void_p))
                    err = np.abs(Rd.reshape(-1, 3, 3) - R).max(axis=(1, 2))
                    Kfull = np.stack([Kf[:, 0], Kf[:, 1], Kf[:, 2], Kf[:, 1], Kf[:, 3], Kf[:, 4], Kf[:, 2], Kf[:, 4], Kf[:, 5]], 1).reshape(-1, 3, 3)
                    kerr = (np.abs(Kfull - Kinv).max(axis=(1, 2)) / np.abs(Kinv).max(axis=(1, 2))).max()
                    row.append(f"{name}: R max {err.max():.1e} (99.9% {np.quantile(err, 0.999):.1e}), K^-1 rel {kerr:.1e}")
                print(f"sweeps {sweeps}, evaluations {evals}:\n   " + "\n   ".join(row))


if __name__ == "__main__":
    main()
