"""Kernel list of ONE training iteration at the notebook batch size (C3 shape, 20 000 frames), for `ncu --metrics gpu__time_duration.sum`:
which launches an iteration consists of and how long each runs (serialised, under the profiler: shares, not bench values)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "colvars-finder_b200"))
import __graft_entry__ as g; g.build()
from colvarsfinder import core, nn, utils
import bench_data as bd
dev = torch.device("cuda", 0)
n = 20000
base = bd.DIPEPTIDE_NM * 10.0
X = bd.frames(base, n, dev, 1)
w = bd.boltzmann_weights(n, dev, 1)
traj = bd.SyntheticTrajectory(X[:1024].cpu().numpy(), np.ones(1024), dt=1.0)
task = core.EigenFunctionTask(traj, utils.Align(base, list(range(22))), nn.EigenFunctions([66, 20, 20, 20, 1], 3), "/tmp/sbk", 20.0,
                              [1.0, 0.6, 0.3], k=3, learning_rate=0.001, device=dev, verbose=False, debug_mode=False)
for it in range(4):
    if it == 3:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    task.optimizer.zero_grad(set_to_none=True)
    loss = task.loss_func(X, w)[0]
    loss.backward()
    task.optimizer.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
