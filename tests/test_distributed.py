"""Data-parallel algebra of the step on 2 CPU ranks (gloo): frame shards + one all-reduce of the fp64 batch sums
after pass 1 and one of the fp64 gradient sums after pass 2 reproduce the single-rank result.  The per-rank
arithmetic is the oracle (no GPU here); sharding and the collectives are the package's own (_ops.shard_range,
_ops.allreduce_sum_), i.e. exactly what the CUDA path calls between its kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as cf
from tests import _cases as C


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from colvarsfinder import _ops
        c = C.eigen_case(name)
        pp = cf.Preproc(align_idx=c["align_idx"], ref=c["ref"], feats=c["features"]) if c["pp_kind"] == "mol" \
            else cf.Preproc(identity=True)
        X, w = c["X"].astype(np.float64), c["w"].astype(np.float64)
        nets = [[np.asarray(p, np.float64) for p in n] for n in c["params"]]
        a = np.ones(X[0].size) if c["diag_coeff"] is None else c["diag_coeff"].astype(np.float64)
        assert _ops.world_size() == world and _ops.rank() == rank
        lo, hi = _ops.shard_range(len(X), rank, world)
        S, state = cf.eigen_stats(X[lo:hi], w[lo:hi], nets, pp, a)                     # pass 1 on the shard
        k = c["k"]
        packed = torch.from_numpy(np.concatenate([[S["S0"]], S["S1"], S["S2"].ravel(), S["SD"]]))
        _ops.allreduce_sum_(packed)                                                    # collective 1
        v = packed.numpy()
        Sg = dict(S0=v[0], S1=v[1:1 + k], S2=v[1 + k:1 + k + k * k].reshape(k, k), SD=v[1 + k + k * k:])
        comb = cf.eigen_combine(Sg, c["alpha"], c["eig_w"], c["beta"], c["sort"])       # identical on every rank
        grads = cf.eigen_grads(w[lo:hi], nets, state, comb)                            # pass 2 on the shard
        flat = torch.from_numpy(np.concatenate([g.ravel() for net in grads for g in net]))
        _ops.allreduce_sum_(flat)                                                      # collective 2
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=comb["loss"], eig=comb["eig"], cvec=comb["cvec"],
                 grad=flat.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_step_matches_single_rank(tmp_path):
    name = "eigen_dipep_subset_diag"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    c = C.eigen_case(name)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for key in ("loss", "eig", "cvec", "grad"):
        assert np.array_equal(r0[key], r1[key]), key                                  # ranks agree bit for bit
    assert abs(float(r0["loss"]) - float(c["g64_loss"])) <= 1e-10 * abs(float(c["g64_loss"]))
    np.testing.assert_allclose(r0["eig"], c["g64_eig"], rtol=1e-9)
    gold = np.concatenate([g.ravel() for net in c["g64"] for g in net])
    assert C.rel_l2(r0["grad"], gold) < 1e-8


def _plan_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from colvarsfinder import _ops, core
        from sklearn.model_selection import train_test_split
        # 100 000 frames with lag 1 on 2 ranks: shards of 50 000 and 49 999 frames -> 40 000 / 39 999 training frames
        n = 100000 - 1
        lo, hi = _ops.shard_range(n, rank, world)
        tr, te = train_test_split(np.arange(hi - lo), test_size=0.2)
        own = (min(1000, len(tr)), min(1000, len(te)), len(tr) // 1000, len(te) // 1000)
        plan = core._iteration_plan(len(tr), len(te), 1000)
        np.savez(os.path.join(out_dir, f"plan{rank}.npz"), own=np.asarray(own), plan=np.asarray(plan))
    finally:
        dist.destroy_process_group()


def test_iteration_plan_is_identical_on_all_ranks(tmp_path):
    """Shards that differ by one frame give different per-rank iteration counts (40 vs 39 here); every iteration ends in
    collectives, so the loop must run the collective minimum on every rank (multi-rank train() could hang)."""
    port = _free_port()
    mp.spawn(_plan_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "plan0.npz"), np.load(tmp_path / "plan1.npz")
    assert p0["own"][2] != p1["own"][2]                       # the situation the plan exists for
    assert np.array_equal(p0["plan"], p1["plan"])
    assert p0["plan"][2] == min(p0["own"][2], p1["own"][2]) and p0["plan"][3] == min(p0["own"][3], p1["own"][3])
