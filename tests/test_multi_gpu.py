"""Hardware data-parallel parity (skips below 2 GPUs): two ranks over NCCL on shards of one batch reproduce the single-rank
loss, eigenvalues and gradients, and a 2-rank train() with shards of unequal length (one frame apart) finishes."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from colvarsfinder import _ops, core, nn, utils
        from oracle import ref_torch
        from oracle.ref_import import FakeTrajectory
        base = ref_torch.DIPEPTIDE_NM * 10.0
        n = 20001                                     # odd: the shards differ by one frame
        X = ref_torch.synth_frames(base, n, seed=5)
        w = ref_torch.boltzmann_weights(n, seed=5)
        torch.manual_seed(1)
        model = nn.EigenFunctions([66, 20, 20, 20, 1], 3)
        task = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), utils.Align(base, list(range(22))), model,
                                      os.path.join(out_dir, f"r{rank}"), 20.0, [1.0, 0.6, 0.3], k=3, batch_size=3000, num_epochs=2,
                                      learning_rate=1e-3, save_model_every_step=0, device=dev, verbose=False, debug_mode=False)
        lo, hi = _ops.shard_range(n, rank, world)
        assert task._traj.shape[0] == hi - lo
        Xd, wd = torch.as_tensor(X, device=dev), torch.as_tensor(w, device=dev)

        def run(Xb, wb):
            model.zero_grad(set_to_none=True)
            out = task.loss_func(Xb, wb, None, None)
            out[0].backward()
            return out[0].detach().double(), out[1].double(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).double()
        with _ops.no_collectives():
            l1, e1, g1 = run(Xd, wd)                                      # the whole batch on one rank
        l2, e2, g2 = run(task._traj, task._weights)                       # this rank's shard + NCCL all-reduces
        res = dict(loss=float((l1 - l2).abs() / l1.abs()), eig=float(((e1 - e2).abs() / e1.abs()).max()),
                   grad=float((g1 - g2).norm() / g1.norm()))
        model.zero_grad(set_to_none=True)
        np.random.seed(3)
        task.train()                                                     # unequal shards: must not hang (collective plan)
        res["iters"] = int(task.loss_list[0][0].shape[0])
        res["final"] = float(task.loss_list[-1][0][-1, 0])
        # the same run with the iteration captured as a CUDA graph (NCCL all-reduces inside) replays the eager steps
        hist = {}
        for mode in ("0", "1"):
            os.environ["CVF_CUDA_GRAPH"] = mode
            os.environ["CVF_CUDA_GRAPH_MIN_STEPS"] = "1"
            torch.manual_seed(1)
            m2 = nn.EigenFunctions([66, 20, 20, 20, 1], 3)
            t2 = core.EigenFunctionTask(FakeTrajectory(X, w.astype(np.float64)), utils.Align(base, list(range(22))), m2,
                                        os.path.join(out_dir, f"g{mode}r{rank}"), 20.0, [1.0, 0.6, 0.3], k=3, batch_size=1500,
                                        num_epochs=3, learning_rate=1e-3, save_model_every_step=0, device=dev, verbose=False,
                                        debug_mode=False)
            np.random.seed(4)
            t2.train()
            hist[mode] = torch.stack([l[0] for l in t2.loss_list])
            if mode == "1":
                res["replays"] = int(t2._graphed_step.replays)
        res["graph_max_abs_diff"] = float((hist["0"] - hist["1"]).abs().max())
        np.savez(os.path.join(out_dir, f"dp{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "dp0.npz"), np.load(tmp_path / "dp1.npz")
    for r in (r0, r1):
        assert r["loss"] < 1e-6 and r["eig"] < 1e-6 and r["grad"] < 1e-5, dict(r)
    assert int(r0["iters"]) == int(r1["iters"]) and float(r0["final"]) == float(r1["final"])
    assert int(r0["replays"]) > 0 and int(r1["replays"]) == int(r0["replays"])
    assert float(r0["graph_max_abs_diff"]) == 0.0 and float(r1["graph_max_abs_diff"]) == 0.0
