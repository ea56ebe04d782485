# per-tensor gradient errors of the layer-wise autoencoder path against the fp64 closed form (development aid)
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'colvars-finder_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from colvarsfinder import core, nn, _lib
from oracle import ref_torch, closed_form as cf
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'tests'))
import _cases as C
class FakeTrajectory:
    def __init__(self, traj, weights, dt=1.0):
        self.trajectory, self.weights, self.dt, self.n_frames = traj, weights, dt, len(traj)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 130
e_dims, d_dims = [150, 260, 200, 2], [2, 200, 260, 150]
torch.manual_seed(B)
enc = [p.numpy() for p in ref_torch.init_mlp_params(e_dims)]
dec = [p.numpy() for p in ref_torch.init_mlp_params(d_dims)]
rng = np.random.default_rng(B)
F = rng.normal(size=(B, 150)).astype(np.float32)
w = ref_torch.boltzmann_weights(B, seed=B)
model = nn.AutoEncoder(e_dims, d_dims)
with torch.no_grad():
    for p, v in zip(model.encoder.parameters(), enc): p.copy_(torch.as_tensor(v))
    for p, v in zip(model.decoder.parameters(), dec): p.copy_(torch.as_tensor(v))
task = core.AutoEncoderTask(FakeTrajectory(F, w.astype(np.float64)), torch.nn.Identity(), model, "/tmp/wd", device=torch.device('cuda'), verbose=False, debug_mode=False)
loss = task.weighted_MSE_loss(task._feature_traj, task._weights)
loss.backward()
lo, genc, gdec = cf.ae_loss_and_grads(F, w, enc, dec)
print('loss', float(loss), lo)
got = [p.grad.cpu().numpy() for p in model.encoder.parameters()] + [p.grad.cpu().numpy() for p in model.decoder.parameters()]
for i, (g, go) in enumerate(zip(got, genc + gdec)):
    print(i, g.shape, 'rel', C.rel_l2(g, go), 'norm got', np.linalg.norm(g), 'ref', np.linalg.norm(go), 'nan', np.isnan(g).sum())
    if g.ndim == 2 and C.rel_l2(g, go) > 1e-3:
        r = np.abs(g - go) / (np.abs(go).max())
        print('   bad rows', np.where(r.max(1) > 1e-3)[0][:20], 'bad cols', np.where(r.max(0) > 1e-3)[0][:40])
