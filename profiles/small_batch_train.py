"""Wall time of train() at notebook batch sizes with and without the CUDA-graph step (CVF_CUDA_GRAPH=0)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "colvars-finder_b200"))
import __graft_entry__ as g; g.build()
from colvarsfinder import core, nn, utils
import bench_data as bd
dev = torch.device("cuda", 0)
def run(kind, graph):
    os.environ["CVF_CUDA_GRAPH"] = graph
    os.environ["CVF_CUDA_GRAPH_MIN_STEPS"] = "1"
    torch.manual_seed(0); np.random.seed(0)
    if kind == "c1":
        n, bs = 100000, 1000
        X = bd.ring_2d(n, dev, 1).cpu().numpy().astype(np.float64)
        traj = bd.SyntheticTrajectory(X, np.ones(n), dt=0.1)
        task = core.EigenFunctionTask(traj, torch.nn.Identity(), nn.EigenFunctions([2, 20, 20, 20, 1], 1), "/tmp/sbt", 20.0, [1.0], k=1,
                                      learning_rate=0.005, batch_size=bs, num_epochs=10, test_ratio=0.2, save_model_every_step=0,
                                      device=dev, verbose=False, debug_mode=False)
    else:
        n, bs = 1000000, 20000
        base = bd.DIPEPTIDE_NM * 10.0
        X = bd.frames(base, n, dev, 1).cpu().numpy()
        traj = bd.SyntheticTrajectory(X, np.ones(n), dt=1.0)
        task = core.EigenFunctionTask(traj, utils.Align(base, list(range(22))), nn.EigenFunctions([66, 20, 20, 20, 1], 3), "/tmp/sbt", 20.0,
                                      [1.0, 0.6, 0.3], k=3, learning_rate=0.001, batch_size=bs, num_epochs=10, test_ratio=0.2,
                                      save_model_every_step=0, device=dev, verbose=False, debug_mode=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        task.train()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tr0 = time.perf_counter()
    steps = task._graphed_step.eager_steps + task._graphed_step.replays
    print(kind, "graph" if graph == "1" else "eager", "train() %.3f s" % (t1 - t0), steps, "steps", "%.0f us/step incl. test loop" % ((t1 - t0) / steps * 1e6),
          "replays", task._graphed_step.replays, "capture s", round(getattr(task._graphed_step, "capture_seconds", 0.0), 3), round(getattr(task._graphed_eval, "capture_seconds", 0.0), 3), "final loss", float(task.train_loss_df["loss"].iloc[-1]), flush=True)
import gc
for kind in ("c1", "c3"):
    for graph in ("0", "0", "1", "1", "1"):     # the first run of a kind also pays one-time costs (module loading, allocator growth)
        run(kind, graph)
        t0 = time.perf_counter(); gc.collect(); torch.cuda.synchronize(); print("   teardown s", round(time.perf_counter() - t0, 3), flush=True)
