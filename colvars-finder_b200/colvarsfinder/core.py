"""Training tasks with the reference's public API (reference colvarsfinder/core.py) on top of the CUDA step.

``EigenFunctionTask`` and ``AutoEncoderTask`` keep the constructor signatures, attributes, return values and
logging names of the reference; what changes is where the arithmetic runs:

* the trajectory shard lives in device memory (this rank's contiguous slice of the frames when launched
  under ``torch.distributed``); the train / test split consumes numpy's global RNG exactly like the
  reference's ``train_test_split`` calls (core.py:465,468 / 672) and mini-batches are slices of the
  device-resident shuffled copy (``shuffle=False``, ``drop_last=True`` semantics of core.py:472-481);
* ``loss_func`` / ``weighted_MSE_loss`` return tensors whose ``backward()`` fills ``param.grad`` from the
  fused CUDA passes (``colvarsfinder._ops``); ``optimizer.step()`` is stock ``torch.optim``.

Outside the envelope (non-Tanh activations, arbitrary pp_layer modules, non-CUDA devices) the constructors
raise: there is no PyTorch fallback.
"""
from __future__ import annotations

import copy
import itertools
import os
import time
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _ops
from . import utils as _utils
from .nn import AutoEncoder, EigenFunctions, RegAutoEncoder, RegModel, chain_spec


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass


_WRITER_WARNED = False


def _make_writer(path):
    """tensorboardX.SummaryWriter of the reference (core.py:143); torch's own tensorboard writer when that package is not
    installed.  With neither, the scalar logs are dropped -- said once, not silently."""
    global _WRITER_WARNED
    errors = []
    for modname in ('tensorboardX', 'torch.utils.tensorboard'):
        try:
            mod = __import__(modname, fromlist=['SummaryWriter'])
            return mod.SummaryWriter(path)
        except ImportError as exc:
            errors.append(f'{modname}: {exc}')
    if not _WRITER_WARNED:
        _WRITER_WARNED = True
        print('[Warning] no tensorboard writer available (%s): the scalar logs (loss, eig_i, Loss/train, ...) of this run are '
              'not written' % '; '.join(errors), flush=True)
    return _NullWriter()


def _iteration_plan(n_train, n_test, batch_size):
    """Batch sizes and iterations per epoch, identical on every rank.

    Every mini-batch ends in collectives (``_ops.allreduce_sum_``), so all ranks must run the same number of iterations;
    shards differ by one frame whenever the frame count is not a multiple of the world size, and ``n // bs`` can then differ
    across ranks.  The plan is the MIN over ranks of each rank's own ``min(batch_size, n_split)`` and ``n_split // bs``
    (reference core.py:470-481 on one process)."""
    bs_train, bs_test = min(batch_size, n_train), min(batch_size, n_test)
    if _ops.world_size() > 1:
        import torch.distributed as dist
        dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
        t = torch.tensor([bs_train, bs_test], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        bs_train, bs_test = int(t[0]), int(t[1])
        t = torch.tensor([n_train // max(bs_train, 1), n_test // max(bs_test, 1)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bs_train, bs_test, int(t[0]), int(t[1])
    return bs_train, bs_test, n_train // max(bs_train, 1), n_test // max(bs_test, 1)


def _shard_to_device(data, lo, hi, device):
    """Rows [lo, hi) of a trajectory / weight array -- numpy (reference WeightedTrajectory) or torch (WeightedTrajectory(device=...))
    -- as a contiguous float32 tensor on `device`."""
    part = data[lo:hi]
    t = part if torch.is_tensor(part) else torch.as_tensor(np.asarray(part))
    return t.to(device=device, dtype=torch.float32).contiguous()


def _split(n, test_ratio, draws):
    """Index split drawn like the reference: sklearn on numpy's global RNG, `draws` calls, last one kept."""
    from sklearn.model_selection import train_test_split
    for _ in range(draws):
        tr, te = train_test_split(np.arange(n), test_size=test_ratio)
    return tr, te


class _GraphedStep:
    """One training iteration -- loss, backward, optimizer step -- replayed as a CUDA graph.

    The loops of the reference (core.py:498-522, 699-712) run many small steps: at the notebook batch sizes (1 000 - 20 000
    states) an iteration is about twenty kernel launches worth 0.1 ms of GPU time behind 0.6 - 0.9 ms of Python, autograd and
    launch overhead.  After WARMUP eager iterations the iteration is captured once -- capture records the launches, it does not
    run them -- and every later iteration is a device-to-device copy of the mini-batch into static buffers plus one graph
    launch.  Every iteration, eager or replayed, is a real training step on its own mini-batch, so the sequence of parameter
    updates is exactly the one of the eager loop.  Any failure to capture falls back to the eager loop.

    ``step_fn(*batch) -> tuple of tensors``: evaluates the loss, calls backward(), returns what the loop logs."""

    WARMUP = 3
    MIN_STEPS = 1000     # a capture costs 2 ms - 0.7 s (the first one in a process is the slow one) and a replayed iteration saves
                         # 0.3 - 0.4 ms: shorter runs stay eager

    def __init__(self, task, step_fn, contexts, planned_steps, train=True):
        self.task, self.step_fn, self.contexts, self.train = task, step_fn, contexts, train   # train=False: evaluation only
        opt = task.optimizer
        ok_opt = isinstance(opt, torch.optim.SGD) or (isinstance(opt, torch.optim.Adam) and
                                                      all(g.get('capturable', False) for g in opt.param_groups))
        # multi-rank runs: NCCL collectives are capturable, every rank captures at the same iteration (the plan is collective)
        ok_dist = task._world == 1 or _ops.backend() == 'nccl'
        self.enabled = (getattr(task, 'use_cuda_graph', True) and os.environ.get('CVF_CUDA_GRAPH', '1') != '0'
                        and ok_dist and ok_opt and
                        planned_steps >= int(os.environ.get('CVF_CUDA_GRAPH_MIN_STEPS', self.MIN_STEPS)))
        self.eager_steps, self.replays = 0, 0
        self.graph, self.static_in, self.static_out, self._keep = None, None, None, None
        self.shapes = None

    def _eager(self, batch):
        if self.train:
            self.task.optimizer.zero_grad(set_to_none=True)
        outs = self.step_fn(*batch)
        if self.train:
            self.task.optimizer.step()
        self.eager_steps += 1
        return outs

    def __call__(self, *batch):
        if not self.enabled or self.eager_steps < self.WARMUP:
            return self._eager(batch)
        shapes = tuple(None if t is None else tuple(t.shape) for t in batch)
        if self.graph is None:
            try:
                _t0 = time.perf_counter()
                self.static_in = [None if t is None else t.detach().clone() for t in batch]
                if self.train:
                    self.task.optimizer.zero_grad(set_to_none=True)
                # capture_begin / capture_end on a side stream by hand: torch.cuda.graph() would also empty the allocator's
                # cache first, which costs up to a second when gigabytes of trajectory temporaries are cached
                graph = torch.cuda.CUDAGraph()
                cur, side = torch.cuda.current_stream(), torch.cuda.Stream(device=self.task.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    graph.capture_begin()
                    try:
                        outs = self.step_fn(*self.static_in)
                        if self.train:
                            self.task.optimizer.step()
                    finally:
                        graph.capture_end()
                cur.wait_stream(side)
                self.graph, self.static_out, self.shapes = graph, outs, shapes
                self.capture_seconds = time.perf_counter() - _t0
                # the graph writes into the scratch buffers that existed at capture: keep them alive even if a context later
                # grows its scratch for a larger batch
                self._keep = []
                for c in self.contexts:
                    self._keep += [w['buf'] for w in getattr(c, '_ws', {}).values() if w]
                    if getattr(c, 'workspace', None) is not None:
                        self._keep.append(c.workspace)
            except Exception as exc:   # capture is an optimisation: any failure means the eager loop
                if self.task._world > 1:
                    raise               # ranks must not diverge (one replaying, one eager): a failed capture ends the run
                self.enabled, self.graph = False, None
                if self.task.verbose:
                    print(f'[Info] CUDA-graph capture of the training step failed ({exc}); continuing with the eager loop', flush=True)
                torch.cuda.synchronize()
                return self._eager(batch)
        elif shapes != self.shapes:
            return self._eager_after_graph(batch)
        else:
            for dst, src in zip(self.static_in, batch):
                if dst is not None:
                    dst.copy_(src)
        self.graph.replay()
        self.replays += 1
        return tuple(o.clone() for o in self.static_out)

    def release(self):
        """Drop the captured graph and its static buffers (the counters stay).  train() calls this when its loops are done: a
        graph that captured NCCL all-reduces keeps a persistent reference on the communicator, and
        torch.distributed.destroy_process_group() then waits for it forever (seen as a 2-rank run that never exits)."""
        if self.graph is not None:
            torch.cuda.synchronize(self.task.device)
        self.graph, self.static_in, self.static_out, self._keep, self.shapes = None, None, None, None, None

    def _eager_after_graph(self, batch):
        # a batch of another size: run it eagerly on gradients of its own (the graph owns the static .grad tensors)
        if not self.train:
            return self._eager(batch)
        saved = [p.grad for g in self.task.optimizer.param_groups for p in g['params']]
        outs = self._eager(batch)
        for p, g in zip([p for gr in self.task.optimizer.param_groups for p in gr['params']], saved):
            p.grad = g
        return outs


class TrainingTask(ABC):
    """Common state of the training tasks (reference core.py:60-249)."""

    def __init__(self, traj_obj, pp_layer, model, model_path, learning_rate, load_model_filename, save_model_every_step, k,
                 batch_size, num_epochs, test_ratio, optimizer_name, device, plot_class, plot_frequency, verbose, debug_mode):
        device = torch.device(device)
        if device.type != 'cuda':      # the signature default stays the reference's torch.device('cpu') (core.py:311,621)
            raise RuntimeError(f"device={device}: this build of colvarsfinder runs the training step in sm_100a CUDA "
                               "kernels only; pass device=torch.device('cuda')")
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.traj_obj = traj_obj
        self.preprocessing_layer = pp_layer.to(device)
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.num_epochs = num_epochs
        self.test_ratio = test_ratio
        self.k = k
        self.model = model.to(device)
        self.load_model_filename = load_model_filename
        self.save_model_every_step = save_model_every_step
        self.model_path = model_path
        self.optimizer_name = optimizer_name
        self.device = device
        self.plot_class = plot_class
        self.plot_frequency = plot_frequency
        self.verbose = verbose
        self.debug_mode = debug_mode
        self.model_name = type(self).__name__
        self._rank, self._world = _ops.rank(), _ops.world_size()
        if self.verbose:
            print('\n[Info] Log directory: {}\n'.format(self.model_path), flush=True)
        self.writer = _make_writer(self.model_path) if self._rank == 0 else _NullWriter()

    def init_model_and_optimizer(self):
        """Optionally load weights (strict=False), then Adam ('adam', any case) or SGD (reference core.py:145-166)."""
        if self.load_model_filename:
            if os.path.isfile(self.load_model_filename):
                self.model.load_state_dict(torch.load(self.load_model_filename, map_location=self.device), strict=False)
                if self.verbose:
                    print(f'model parameters loaded from: {self.load_model_filename}')
            elif self.verbose:
                print(f'model file not found: {self.load_model_filename}')
        if self.optimizer_name.lower() == 'adam':
            # capturable: the step count lives on the device, so that optimizer.step() can be part of a CUDA graph (_GraphedStep)
            # and fused: one kernel for all parameters instead of a dozen foreach launches per step
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.learning_rate, capturable=True, fused=True)
        else:
            self.optimizer = torch.optim.SGD(self.model.parameters(), lr=self.learning_rate)

    def save_model(self, epoch, description="latest"):
        """model.pt + per-CV parameter text files (reference core.py:168-208).  Rank 0 only."""
        if self._rank != 0:
            return
        if self.verbose:
            print(f"\n\nEpoch={epoch}:")
        if self.debug_mode is True:
            d = f'{self.model_path}/models'
            os.makedirs(d, exist_ok=True)
            torch.save(self.model.state_dict(), f'{d}/model_{epoch}.pt')
        d = f'{self.model_path}/{description}'
        os.makedirs(d, exist_ok=True)
        torch.save(self.model.state_dict(), f'{d}/model.pt')
        for idx in range(self.k):
            for name, param in self.model.get_params_of_cv(idx):
                np.savetxt('%s/%d_' % (d, idx) + name.replace('.', '_') + '.txt', param.detach().cpu().numpy())
        if self.verbose:
            print(f'  trained model saved at:\n\t{d}/model.pt')
        # TorchScript export of the collective variables (reference core.py:210-227): scripted_cv_gpu.pt / scripted_cv_cpu.pt.
        # The pre-processing layers of this package are exported as their stock-torch restatement (utils.scriptable), so the
        # files load in libtorch (PLUMED, Colvars) without libcvf.
        cv = self.colvar_model()
        if isinstance(cv, torch.nn.Sequential) and len(cv) == 2:
            cv = torch.nn.Sequential(_utils.scriptable(cv[0]).to(self.device), cv[1])
        try:
            if self.device.type == 'cuda':
                torch.jit.script(cv).save(f'{d}/scripted_cv_gpu.pt')
                if self.verbose:
                    print(f'  script (GPU) model for CVs saved at:\n\t{d}/scripted_cv_gpu.pt\n', flush=True)
                torch.jit.script(copy.deepcopy(cv).to('cpu')).save(f'{d}/scripted_cv_cpu.pt')
            else:
                torch.jit.script(cv).save(f'{d}/scripted_cv_cpu.pt')
            if self.verbose:
                print(f'  script (CPU) model for CVs saved at:\n\t{d}/scripted_cv_cpu.pt\n', flush=True)
        except Exception as exc:   # a user-supplied pp_layer that TorchScript cannot compile: the reference would raise here
            raise RuntimeError(f"TorchScript export of the collective-variable model failed: {exc}") from exc

    def _cv_preprocessing(self):
        """pp_layer as the first stage of colvar_model(): a bare utils.Align returns [B,N,3] frames, the networks take [B,3N]."""
        pp = self.preprocessing_layer
        return _utils.Preprocessing(pp, None) if isinstance(pp, _utils.Align) else pp

    def _shard(self, n):
        return _ops.shard_range(n, self._rank, self._world)

    @abstractmethod
    def train(self):
        pass

    @abstractmethod
    def colvar_model(self):
        pass

    @abstractmethod
    def reg_model(self):
        pass


class EigenFunctionTask(TrainingTask):
    """Eigenfunctions of the generator by the variational (Rayleigh-quotient) loss (reference core.py:251-566)."""

    def __init__(self, traj_obj, pp_layer, model, model_path, alpha, eig_weights, diag_coeff=None, beta=1.0, lag_tau=0,
                 learning_rate=0.01, load_model_filename=None, save_model_every_step=10, sort_eigvals_in_training=True, k=1,
                 batch_size=1000, num_epochs=10, test_ratio=0.2, optimizer_name='Adam', device=torch.device('cpu'),
                 plot_class=None, plot_frequency=0, verbose=True, debug_mode=True):
        super().__init__(traj_obj, pp_layer, model, model_path, learning_rate, load_model_filename, save_model_every_step, k,
                         batch_size, num_epochs, test_ratio, optimizer_name, device, plot_class, plot_frequency, verbose,
                         debug_mode)
        assert isinstance(model, EigenFunctions), 'model must be an object of the class EigenFunctions'
        assert k == len(model.eigen_funcs), \
            f'number of cv ({k}) must equal the number of eigenfunctions ({len(model.eigen_funcs)})'
        self._alpha = alpha
        self._sort_eigvals_in_training = sort_eigvals_in_training
        self._eig_w = eig_weights
        self._cvec = None
        self.traj_dt = traj_obj.dt
        lag_idx = lag_tau / self.traj_dt
        assert abs(lag_idx - int(lag_idx)) < 1e-6, \
            f'lag-time ({lag_tau}) not divisable by the timestep {self.traj_dt} of the trajectory'
        self.lag_idx = int(lag_idx)
        self._ij_list = list(itertools.combinations(range(self.k), 2))
        self._num_ij_pairs = len(self._ij_list)
        if self.verbose:
            print('\nEigenfunctions:\n', self.model, flush=True)
        self.init_model_and_optimizer()
        traj, weights = traj_obj.trajectory, traj_obj.weights
        self.tot_dim = int(np.prod(traj.shape[1:]))
        self._beta = beta
        if diag_coeff is not None:
            assert diag_coeff.dim() == 1 and diag_coeff.size(dim=0) == self.tot_dim, \
                f'diag_coeff should be a 1d tensor of length {self.tot_dim}, current shape: {diag_coeff}'
            self._diag_coeff = diag_coeff
        else:
            self._diag_coeff = torch.ones(self.tot_dim)
        # this rank's contiguous shard of the frames, resident in HBM; with a time lag, the shard carries the lag_idx frames
        # that follow it so that index + lag_idx (reference core.py:511) stays local
        lo, hi = self._shard(traj.shape[0] - self.lag_idx)
        hi += self.lag_idx
        self._traj = _shard_to_device(traj, lo, hi, self.device)
        self._weights = _shard_to_device(weights, lo, hi, self.device)
        self._ctx = _ops.EigenContext(self.model, self.preprocessing_layer, traj.shape[1:], self.device, alpha, eig_weights,
                                      beta, diag_coeff, sort_eigvals_in_training)

    def get_reordered_eigenfunctions(self, model, cvec):
        """Deep copy of `model` with its eigenfunctions permuted by cvec (reference core.py:356-370)."""
        new = copy.deepcopy(model)
        new.eigen_funcs = torch.nn.ModuleList([copy.deepcopy(model.eigen_funcs[int(i)]) for i in cvec])
        return new

    def colvar_model(self):
        if self._cvec is None:
            self._cvec = torch.arange(self.k)
        return torch.nn.Sequential(self._cv_preprocessing(), self.get_reordered_eigenfunctions(self.model, self._cvec))

    def reg_model(self):
        return None

    def loss_func(self, X, weight, X_lagged=None, weight_lagged=None):
        """Total loss, eigenvalues (sorted when sort_eigvals_in_training), variational objective, penalty and
        the ordering cvec -- reference core.py:387-457: generator branch when lag_tau = 0, transfer-operator branch
        (X_lagged, weight_lagged used) otherwise.  ``loss.backward()`` runs the backward pass(es)."""
        if self.lag_idx == 0:
            return _ops.eigen_loss(self._ctx, X, weight)
        if X_lagged is None or weight_lagged is None:
            raise RuntimeError("lag_tau > 0: loss_func needs the time-lagged batch and its weights")
        return _ops.eigen_lag_loss(self._ctx, self.traj_dt * self.lag_idx, X, weight, X_lagged, weight_lagged)

    def _epoch_batches(self, X, w, bs, n_it, Xl=None, wl=None):
        for s in range(0, n_it * bs, bs):
            if Xl is None:
                yield X[s:s + bs], w[s:s + bs], None, None
            else:
                yield X[s:s + bs], w[s:s + bs], Xl[s:s + bs], wl[s:s + bs]

    def train(self):
        """Epoch / mini-batch loop of reference core.py:459-566 on device-resident shards."""
        ll = self._traj.shape[0] - self.lag_idx
        idx_train, idx_test = _split(ll, self.test_ratio, draws=2)      # the reference splits twice (core.py:465,468)
        it, ie = torch.as_tensor(idx_train, device=self.device), torch.as_tensor(idx_test, device=self.device)
        X_train, w_train = self._traj[it], self._weights[it]
        X_test, w_test = self._traj[ie], self._weights[ie]
        Xl_train = wl_train = Xl_test = wl_test = None
        if self.lag_idx > 0:      # X_lagged = traj[index + lag_idx] (reference core.py:510-512,546-547)
            Xl_train, wl_train = self._traj[it + self.lag_idx], self._weights[it + self.lag_idx]
            Xl_test, wl_test = self._traj[ie + self.lag_idx], self._weights[ie + self.lag_idx]
        bs_train, bs_test, n_it_train, n_it_test = _iteration_plan(X_train.shape[0], X_test.shape[0], self.batch_size)
        self.loss_list = []
        min_loss = float("inf")
        if self._rank == 0:
            print("\nTraining starts.\n%d epochs in total, batch sizes (train/test): %d/%d" % (self.num_epochs, bs_train, bs_test))
            print("\nTrain set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_train), n_it_train, n_it_train * self.num_epochs), flush=True)
            print("Test set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_test), n_it_test, n_it_test * self.num_epochs), flush=True)
        loss_names = ['loss', 'eigen_non_penalty', 'eigen_penalty'] + ['eig_%d' % (i + 1) for i in range(self.k)]

        def one_step(X, weight, Xl, wl):
            loss, eig_vals, non_penalty_loss, penalty, cvec = self.loss_func(X, weight, Xl, wl)
            loss.backward()
            return torch.cat([torch.stack([loss.detach(), non_penalty_loss, penalty]), eig_vals]), cvec

        def one_eval(X, weight, Xl, wl):
            loss, eig_vals, non_penalty_loss, penalty, _ = self.loss_func(X, weight, Xl, wl)
            return (torch.cat([torch.stack([loss.detach(), non_penalty_loss, penalty]), eig_vals]),)

        graphed = self._graphed_step = _GraphedStep(self, one_step, [self._ctx], n_it_train * self.num_epochs)
        graphed_test = self._graphed_eval = _GraphedStep(self, one_eval, [self._ctx], n_it_test * self.num_epochs, train=False)
        for epoch in range(self.num_epochs):
            self.model.train()
            train_rows = []
            loss = None
            for X, weight, Xl, wl in self._epoch_batches(X_train, w_train, bs_train, n_it_train, Xl_train, wl_train):
                row, self._cvec = graphed(X, weight, Xl, wl)      # zero_grad, loss_func, backward, optimizer.step
                train_rows.append(row)
                loss = row[0]
            if self.save_model_every_step > 0 and epoch % self.save_model_every_step == self.save_model_every_step - 1:
                self.save_model(epoch)
                if loss is not None and loss < min_loss:
                    min_loss = loss
                    self.save_model(epoch, 'best')
            if self.plot_frequency > 0 and epoch % self.plot_frequency == self.plot_frequency - 1:
                if self.plot_class is not None:
                    self.plot_class.plot(self.colvar_model(), epoch=epoch)
            test_rows = []
            for X, weight, Xl, wl in self._epoch_batches(X_test, w_test, bs_test, n_it_test, Xl_test, wl_test):
                test_rows.append(graphed_test(X, weight, Xl, wl)[0])
            # one device->host transfer per epoch for the whole log
            tr = torch.stack(train_rows).cpu() if train_rows else torch.zeros(0, 3 + self.k)
            te = torch.stack(test_rows).cpu() if test_rows else torch.zeros(0, 3 + self.k)
            self.loss_list.append([tr, te])
            mean_tr, mean_te = torch.mean(tr, 0), torch.mean(te, 0)
            for i, name in enumerate(loss_names):
                self.writer.add_scalar('%s/train' % name, mean_tr[i], epoch)
                self.writer.add_scalar('%s/test' % name, mean_te[i], epoch)
        graphed.release(), graphed_test.release()
        import pandas as pd
        self.train_loss_df = pd.DataFrame(torch.cat([torch.mean(l[0], dim=0, keepdim=True) for l in self.loss_list]).numpy(),
                                          columns=loss_names)
        self.test_loss_df = pd.DataFrame(torch.cat([torch.mean(l[1], dim=0, keepdim=True) for l in self.loss_list]).numpy(),
                                         columns=loss_names)


class AutoEncoderTask(TrainingTask):
    """Autoencoder with the weighted reconstruction loss (reference core.py:569-744)."""

    def __init__(self, traj_obj, pp_layer, model, model_path, learning_rate=0.01, load_model_filename=None,
                 save_model_every_step=10, batch_size=1000, num_epochs=10, test_ratio=0.2, optimizer_name='Adam',
                 device=torch.device('cpu'), plot_class=None, plot_frequency=0, verbose=True, debug_mode=True):
        super().__init__(traj_obj, pp_layer, model, model_path, learning_rate, load_model_filename, save_model_every_step,
                         model.encoded_dim, batch_size, num_epochs, test_ratio, optimizer_name, device, plot_class,
                         plot_frequency, verbose, debug_mode)
        assert isinstance(model, AutoEncoder), 'model must be an object of the class AutoEncoder'
        self.init_model_and_optimizer()
        traj = traj_obj.trajectory
        lo, hi = self._shard(traj.shape[0])
        self._weights = _shard_to_device(traj_obj.weights, lo, hi, self.device)
        x = _shard_to_device(traj, lo, hi, self.device)
        # whole-trajectory pre-pass (reference core.py:635) on the device
        self._feature_traj = self.preprocessing_layer(x)
        if self._feature_traj.dim() != 2:
            self._feature_traj = self._feature_traj.reshape(self._feature_traj.shape[0], -1)
        self._feature_traj = self._feature_traj.contiguous()
        if self.verbose:
            print('\nShape of trajectory data array:\n {}'.format(self._feature_traj.shape), flush=True)
        self._ctx = _ops.AEContext(self.model, self.device)

    def colvar_model(self):
        # a deep copy, as EigenFunctionTask.colvar_model hands out: a callback that moves the module (cv.to('cpu'), as
        # save_model does) must not re-point the live parameters away from the flat buffer the kernels and the graph read
        return torch.nn.Sequential(self._cv_preprocessing(), copy.deepcopy(self.model.encoder))

    def reg_model(self):
        return None

    def weighted_MSE_loss(self, X, weight):
        """sum_l w_l |dec(enc(X_l)) - X_l|^2 / sum_l w_l  (reference core.py:652-666)."""
        return _ops.ae_loss(self._ctx, X, weight)

    def train(self):
        """Loop of reference core.py:668-744."""
        n = self._feature_traj.shape[0]
        idx_train, idx_test = _split(n, self.test_ratio, draws=1)
        it, ie = torch.as_tensor(idx_train, device=self.device), torch.as_tensor(idx_test, device=self.device)
        X_train, w_train = self._feature_traj[it], self._weights[it]
        X_test, w_test = self._feature_traj[ie], self._weights[ie]
        bs_train, bs_test, n_it_train, n_it_test = _iteration_plan(X_train.shape[0], X_test.shape[0], self.batch_size)
        self.loss_list = []
        min_loss = float("inf")
        if self._rank == 0:
            print("\nTraining starts.\n%d epochs in total, batch sizes (train/test): %d/%d" % (self.num_epochs, bs_train, bs_test))
            print("\nTrain set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_train), n_it_train, n_it_train * self.num_epochs), flush=True)
            print("Test set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_test), n_it_test, n_it_test * self.num_epochs), flush=True)

        def one_step(X, weight):
            loss = self.weighted_MSE_loss(X, weight)
            loss.backward()
            return (loss.detach(),)

        def one_eval(X, weight):
            return (self.weighted_MSE_loss(X, weight).detach(),)

        graphed = self._graphed_step = _GraphedStep(self, one_step, [self._ctx], n_it_train * self.num_epochs)
        graphed_test = self._graphed_eval = _GraphedStep(self, one_eval, [self._ctx], n_it_test * self.num_epochs, train=False)
        for epoch in range(self.num_epochs):
            self.model.train()
            train_loss = []
            loss = None
            for s in range(0, n_it_train * bs_train, bs_train):
                loss, = graphed(X_train[s:s + bs_train], w_train[s:s + bs_train])   # zero_grad, loss, backward, optimizer.step
                train_loss.append(loss)
            if self.save_model_every_step > 0 and epoch % self.save_model_every_step == self.save_model_every_step - 1:
                self.save_model(epoch)
                if loss is not None and loss < min_loss:
                    min_loss = loss
                    self.save_model(epoch, 'best')
            if self.plot_frequency > 0 and epoch % self.plot_frequency == self.plot_frequency - 1:
                if self.plot_class is not None:
                    self.plot_class.plot(self.colvar_model(), epoch=epoch)
            self.model.eval()
            with torch.no_grad():
                test_loss = [graphed_test(X_test[s:s + bs_test], w_test[s:s + bs_test])[0]
                             for s in range(0, n_it_test * bs_test, bs_test)]
            tr = torch.stack(train_loss).cpu() if train_loss else torch.zeros(0)
            te = torch.stack(test_loss).cpu() if test_loss else torch.zeros(0)
            self.loss_list.append([tr, te])
            self.writer.add_scalar('Loss/train', torch.mean(tr), epoch)
            self.writer.add_scalar('Loss/test', torch.mean(te), epoch)
        graphed.release(), graphed_test.release()
        import pandas as pd
        self.train_loss_df = pd.DataFrame(torch.stack([torch.mean(l[0]) for l in self.loss_list]).numpy(), columns=['loss'])
        self.test_loss_df = pd.DataFrame(torch.stack([torch.mean(l[1]) for l in self.loss_list]).numpy(), columns=['loss'])


class _Moments:
    """Weighted moments of k scalar functions from the batch sums of ``cvf::eigen_stats`` (fp64, on the device, differentiable):
    ``S0`` = sum w, ``mean[k]``, ``cov[k,k]`` (``var`` = its diagonal), ``SD[k]`` = sum w |grad f_i|^2."""

    def __init__(self, stats, k):
        self.S0 = stats[0]
        self.mean = stats[1:1 + k] / self.S0
        self.cov = stats[1 + k:1 + k + k * k].view(k, k) / self.S0 - torch.outer(self.mean, self.mean)
        self.var = torch.diagonal(self.cov)
        self.SD = stats[1 + k + k * k:1 + 2 * k + k * k]


class RegAutoEncoderTask(TrainingTask):
    """Regularised autoencoder (reference core.py:746-1217): weighted (time-lagged) reconstruction loss, the eigenfunction loss
    of the regularisers ``reg_i(encoder(r(x)))`` (generator or transfer operator) and three penalties on the encoder.

    Every term is a function of a handful of batch sums, each produced by the CUDA passes behind ``cvf::eigen_stats`` /
    ``cvf::ae_sums``; the combination, its ordering logic (core.py:1017-1037) and the chain rule down to the shared encoder
    parameters are a few scalar-sized torch operations on the device:

    * reconstruction (core.py:876-887): ``cvf::ae_sums`` on ``r(X)`` with target ``r(X_lagged)``;
    * encoder penalties (core.py:889-973): the k encoder components seen as k networks that share their hidden layers
      (the packed parameter vector repeats them), pre-processing = identity on ``r(X)``: ``SD`` gives the gradient norm, the
      second moments the variance and orthogonality terms;
    * eigenfunction loss (core.py:975-1037): network i = encoder followed by regulariser i, with the encoder's last (linear)
      layer folded into the regulariser's first one; the gradient is taken w.r.t. the raw state X, through ``r``."""

    def __init__(self, traj_obj, pp_layer, model, model_path, eig_weights=[], learning_rate=0.01, load_model_filename=None,
                 save_model_every_step=10, batch_size=1000, num_epochs=10, test_ratio=0.2, optimizer_name='Adam', alpha=1.0,
                 gamma=[0.0, 0.0], eta=[0.0, 0.0, 0.0], lag_tau_ae=0, lag_tau_reg=0, beta=1.0, device=torch.device('cpu'),
                 plot_class=None, plot_frequency=0, freeze_encoder=False, verbose=True, debug_mode=True):
        super().__init__(traj_obj, pp_layer, model, model_path, learning_rate, load_model_filename, save_model_every_step,
                         model.encoded_dim, batch_size, num_epochs, test_ratio, optimizer_name, device, plot_class,
                         plot_frequency, verbose, debug_mode)
        self.init_model_and_optimizer()
        assert isinstance(model, RegAutoEncoder), 'model must be an object of the class RegAutoEncoder'
        assert model.num_reg == len(eig_weights), 'number of weights does not match the number of eigenfunctions!'
        traj, weights = traj_obj.trajectory, traj_obj.weights
        self.alpha = alpha
        self.gamma = gamma
        self.eta = eta
        self.num_reg = model.num_reg
        self.tot_dim = int(np.prod(traj.shape[1:]))
        self._eps = 1e-5
        self._eig_w = eig_weights
        self._cvec = None
        self.freeze_encoder = freeze_encoder
        self.traj_dt = traj_obj.dt
        lag_ae_idx = lag_tau_ae / self.traj_dt
        lag_idx = lag_tau_reg / self.traj_dt
        assert abs(lag_ae_idx - int(lag_ae_idx)) < 1e-6 and abs(lag_idx - int(lag_idx)) < 1e-6, \
            f'lag-times ({lag_tau_ae}, {lag_tau_reg}) not divisable by the timestep {self.traj_dt} of the trajectory'
        self.lag_ae_idx = int(lag_ae_idx)
        self.lag_idx = int(lag_idx)
        self._beta = beta
        if self.gamma[0] + self.gamma[1] > self._eps:
            assert self.num_reg > 0, 'number of eigenfunctions must be positive!'
            self._ij_list = list(itertools.combinations(range(self.num_reg), 2))
            self._num_ij_pairs = len(self._ij_list)
            if self.lag_idx == 0:
                self._diag_coeff = torch.ones(self.tot_dim)      # only the identity matrix, as in the reference
        if self.eta[2] > self._eps:
            self._enc_ij_list = list(itertools.combinations(range(self.k), 2))
            self._enc_num_ij_pairs = len(self._enc_ij_list)
        # this rank's shard plus the frames the two lag times reach into
        halo = max(self.lag_idx, self.lag_ae_idx)
        lo, hi = self._shard(traj.shape[0] - halo)
        hi += halo
        self._halo = halo
        self._traj = _shard_to_device(traj, lo, hi, self.device)
        self._weights = _shard_to_device(weights, lo, hi, self.device)
        if self.verbose:
            print('\nShape of trajectory data array:\n {}'.format(self._traj.shape), flush=True)
        # contexts of the CUDA passes
        e_dims, e_acts, self._enc_lins = chain_spec(self.model.encoder)
        self._ae_ctx = _ops.AEContext(self.model, self.device)
        self._enc_ctx = self._reg_ctx = None
        if max(self.eta) > self._eps:
            self._enc_ctx = _ops.EigenContext(None, torch.nn.Identity(), (e_dims[0],), self.device, 0.0, [], 1.0, None, False,
                                              dims=list(e_dims[:-1]) + [1], acts=list(e_acts), k=self.k)
        if self.gamma[0] + self.gamma[1] > self._eps:
            r_specs = [chain_spec(f) for f in self.model.reg]
            r_dims, r_acts, _ = r_specs[0]
            self._reg_lins = [spec[2] for spec in r_specs]
            self._reg_ctx = _ops.EigenContext(None, self.preprocessing_layer, traj.shape[1:], self.device, 0.0, eig_weights, beta,
                                              None, True, dims=list(e_dims[:-1]) + list(r_dims[1:]),
                                              acts=list(e_acts[:-1]) + list(r_acts), k=self.num_reg)

    # ---- models handed to callbacks / save_model (deep copies: see AutoEncoderTask.colvar_model)
    def colvar_model(self):
        return torch.nn.Sequential(self._cv_preprocessing(), copy.deepcopy(self.model.encoder))

    def reg_model(self):
        if self._cvec is None:
            self._cvec = torch.arange(self.model.num_reg)
        return torch.nn.Sequential(self._cv_preprocessing(), RegModel(copy.deepcopy(self.model), [int(c) for c in self._cvec]))

    # ---- packed parameter vectors of the two network families (differentiable w.r.t. the model's parameters)
    def _enc_component_params(self):
        trunk = [t for lin in self._enc_lins[:-1] for t in (lin.weight.reshape(-1), lin.bias)]
        last = self._enc_lins[-1]
        parts = []
        for i in range(self.k):
            parts += trunk + [last.weight[i].reshape(-1), last.bias[i:i + 1]]
        return torch.cat(parts)

    def _reg_chain_params(self):
        trunk = [t for lin in self._enc_lins[:-1] for t in (lin.weight.reshape(-1), lin.bias)]
        last = self._enc_lins[-1]
        parts = []
        for lins in self._reg_lins:
            first = lins[0]      # first regulariser layer o last (linear) encoder layer = one linear layer
            parts += trunk + [(first.weight @ last.weight).reshape(-1), first.weight @ last.bias + first.bias]
            parts += [t for lin in lins[1:] for t in (lin.weight.reshape(-1), lin.bias)]
        return torch.cat(parts)

    def _features(self, X):
        """r(X) as a flat [B, d_r] array (the CUDA pre-pass for the molecular layers)."""
        if isinstance(self.preprocessing_layer, torch.nn.Identity):
            return X
        F = self.preprocessing_layer(X)
        return F.reshape(F.shape[0], -1)

    def _enc_moments(self, X, weight):
        F, weight = _ops._check_batch(self._features(X), weight, "RegAutoEncoderTask (encoder penalties)")
        _, stats = _ops.eigen_stats_op(F, weight, self._enc_component_params(), self._enc_ctx.handle, 0)
        return _Moments(stats, self.k)

    # ---- the loss terms (same names and return values as the reference)
    def weighted_MSE_loss(self, X, X_lagged, weight):
        """sum_l w_l |dec(enc(r(X_l))) - r(X_lagged_l)|^2 / sum_l w_l  (reference core.py:876-887)."""
        F = self._features(X)
        T = None if X_lagged is X else self._features(X_lagged)
        return _ops.ae_loss(self._ae_ctx, F, weight, target=T)

    def reg_enc_grad_loss(self, X, weight, moments=None):
        """sum_i E_w |grad_r enc_i|^2  (reference core.py:889-910)."""
        m = moments or self._enc_moments(X, weight)
        return (m.SD.sum() / m.S0).to(torch.float32)

    def reg_enc_norm_loss(self, X, weight, moments=None):
        """sum_i (var_w enc_i - 1)^2  (reference core.py:912-935)."""
        m = moments or self._enc_moments(X, weight)
        return ((m.var - 1.0) ** 2).sum().to(torch.float32)

    def reg_enc_orthognal_loss(self, X, weight, moments=None):
        """sum_{i<j} cov_w(enc_i, enc_j)^2  (reference core.py:937-962)."""
        m = moments or self._enc_moments(X, weight)
        return (torch.triu(m.cov, diagonal=1) ** 2).sum().to(torch.float32)

    def reg_eigen_loss(self, X, weight, X_lagged, weight_lagged):
        """Eigenvalues (sorted), variational objective, penalty and ordering of the regularisers (reference core.py:964-1037)."""
        ectx, K = self._reg_ctx, self.num_reg
        X, weight = _ops._check_batch(X, weight, "RegAutoEncoderTask.reg_eigen_loss")
        params = self._reg_chain_params()
        omega = torch.as_tensor([float(v) for v in self._eig_w], dtype=torch.float64, device=self.device)
        y, stats = _ops.eigen_stats_op(X, weight, params, ectx.handle, 0)
        m = _Moments(stats, K)
        if self.lag_idx == 0:      # generator: Rayleigh quotients (core.py:1008-1011)
            ratio = m.SD / (m.S0 * self._beta) / m.var
            cvec = torch.argsort(ratio.detach(), stable=True)
            non_penalty = (omega * ratio[cvec]).sum()
            eig = ratio.detach()[cvec]
        else:                      # transfer operator (core.py:1012-1013,1023-1024)
            X_lagged, weight_lagged = _ops._check_batch(X_lagged, weight_lagged, "RegAutoEncoderTask.reg_eigen_loss (time-lagged data)")
            yl, stats_l = _ops.eigen_stats_op(X_lagged, weight_lagged, params, ectx.handle, 1)
            ml = _Moments(stats_l, K)
            sx = _ops.eigen_tlag_sx_op(y, yl, weight, ectx.handle)
            tau = self.traj_dt * self.lag_idx
            denom = ml.var + m.var
            eig_all = (sx / (tau * m.S0) / denom).detach()
            cvec = torch.argsort(eig_all, stable=True)
            # the reference pairs the numerator of term idx with the denominator of network cvec[idx] (core.py:1024)
            non_penalty = (omega * sx / denom[cvec]).sum() / (tau * m.S0)
            eig = eig_all[cvec]
        penalty = ((m.var - 1.0) ** 2).sum() + (torch.triu(m.cov, diagonal=1) ** 2).sum()
        return eig.to(torch.float32), non_penalty.to(torch.float32), penalty.to(torch.float32), cvec

    def _total_loss(self, X, weight, X_ae, X_reg, w_reg):
        """The combination of reference core.py:1066-1113; returns (loss, the logged row, cvec or None)."""
        zero = torch.zeros((), dtype=torch.float32, device=self.device)
        ae_loss = self.weighted_MSE_loss(X, X if X_ae is None else X_ae, weight) if self.alpha > self._eps else zero
        enc = [zero, zero, zero]
        if max(self.eta) > self._eps:
            m = self._enc_moments(X, weight)
            fns = (self.reg_enc_grad_loss, self.reg_enc_norm_loss, self.reg_enc_orthognal_loss)
            enc = [fn(X, weight, m) if self.eta[i] > self._eps else zero for i, fn in enumerate(fns)]
        cvec = None
        if self.gamma[0] + self.gamma[1] > self._eps:
            eig_vals, eigen_0, eigen_1, cvec = self.reg_eigen_loss(X, weight, X_reg, w_reg)
        else:
            eigen_0 = eigen_1 = zero
            eig_vals = torch.zeros(self.num_reg, dtype=torch.float32, device=self.device)
        loss = self.alpha * ae_loss + self.gamma[0] * eigen_0 + self.gamma[1] * eigen_1 \
            + self.eta[0] * enc[0] + self.eta[1] * enc[1] + self.eta[2] * enc[2]
        row = torch.cat([torch.stack([loss.detach(), ae_loss.detach(), eigen_0.detach(), eigen_1.detach()]), eig_vals.detach(),
                         torch.stack([e.detach() for e in enc])])
        return loss, row, cvec

    def train(self):
        """Loop of reference core.py:1039-1217 on device-resident shards."""
        ll = self._traj.shape[0] - self._halo
        idx_train, idx_test = _split(ll, self.test_ratio, draws=1)
        it, ie = torch.as_tensor(idx_train, device=self.device), torch.as_tensor(idx_test, device=self.device)

        def gather(idx):
            X, w = self._traj[idx], self._weights[idx]
            Xa = self._traj[idx + self.lag_ae_idx] if self.lag_ae_idx > 0 and self.alpha > self._eps else None
            use_reg = self.lag_idx > 0 and self.gamma[0] + self.gamma[1] > self._eps
            Xr = self._traj[idx + self.lag_idx] if use_reg else None
            wr = self._weights[idx + self.lag_idx] if use_reg else None
            return X, w, Xa, Xr, wr

        train_set, test_set = gather(it), gather(ie)
        bs_train, bs_test, n_it_train, n_it_test = _iteration_plan(train_set[0].shape[0], test_set[0].shape[0], self.batch_size)
        self.loss_list = []
        min_loss = float("inf")
        if self._rank == 0:
            print("\nTraining starts.\n%d epochs in total, batch sizes (train/test): %d/%d" % (self.num_epochs, bs_train, bs_test))
            print("\nTrain set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_train), n_it_train, n_it_train * self.num_epochs), flush=True)
            print("Test set:\n\t%d data, %d iterations per epoch, %d iterations in total." %
                  (len(idx_test), n_it_test, n_it_test * self.num_epochs), flush=True)
        loss_names = ['loss', 'ae_loss', 'eigen_non_penalty', 'eigen_penalty'] + ['eig_%d' % i for i in range(self.num_reg)] \
            + ['encoder_gradient', 'encoder_norm', 'encoder_orthogonality']
        no_cvec = torch.full((max(self.num_reg, 1),), -1, dtype=torch.int64, device=self.device)

        def one_step(X, weight, Xa, Xr, wr):
            if self.freeze_encoder is True:
                for param in self.model.encoder.parameters():
                    param.requires_grad = False
            loss, row, cvec = self._total_loss(X, weight, Xa, Xr, wr)
            loss.backward()
            if self.freeze_encoder is True:
                for param in self.model.encoder.parameters():
                    param.requires_grad = True
            return row, no_cvec if cvec is None else cvec

        def one_eval(X, weight, Xa, Xr, wr):
            _, row, cvec = self._total_loss(X, weight, Xa, Xr, wr)
            return row, no_cvec if cvec is None else cvec

        contexts = [c for c in (self._ae_ctx, self._enc_ctx, self._reg_ctx) if c is not None]
        graphed = self._graphed_step = _GraphedStep(self, one_step, contexts, n_it_train * self.num_epochs)
        graphed_test = self._graphed_eval = _GraphedStep(self, one_eval, contexts, n_it_test * self.num_epochs, train=False)

        def batches(data, bs, n_it):
            for s in range(0, n_it * bs, bs):
                yield tuple(None if t is None else t[s:s + bs] for t in data)

        for epoch in range(self.num_epochs):
            self.model.train()
            train_rows, loss = [], None
            for batch in batches(train_set, bs_train, n_it_train):
                row, cvec = graphed(*batch)            # zero_grad, loss, backward, optimizer.step
                if self._reg_ctx is not None:
                    self._cvec = cvec
                train_rows.append(row)
                loss = row[0]
            if self.save_model_every_step > 0 and epoch % self.save_model_every_step == self.save_model_every_step - 1:
                self.save_model(epoch)
                if loss is not None and loss < min_loss:
                    min_loss = loss
                    self.save_model(epoch, 'best')
            if self.plot_frequency > 0 and epoch % self.plot_frequency == self.plot_frequency - 1:
                if self.plot_class is not None:
                    self.plot_class.plot(self.colvar_model(), self.reg_model(), epoch=epoch)
            test_rows = []
            with torch.no_grad():
                for batch in batches(test_set, bs_test, n_it_test):
                    row, cvec = graphed_test(*batch)
                    if self._reg_ctx is not None:
                        self._cvec = cvec
                    test_rows.append(row)
            width = 7 + self.num_reg
            tr = torch.stack(train_rows).cpu() if train_rows else torch.zeros(0, width)
            te = torch.stack(test_rows).cpu() if test_rows else torch.zeros(0, width)
            self.loss_list.append([tr, te])
            mean_tr, mean_te = torch.mean(tr, 0), torch.mean(te, 0)
            for i, name in enumerate(loss_names):
                self.writer.add_scalar('%s/train' % name, mean_tr[i], epoch)
                self.writer.add_scalar('%s/test' % name, mean_te[i], epoch)
        graphed.release(), graphed_test.release()
        import pandas as pd
        self.train_loss_df = pd.DataFrame(torch.cat([torch.mean(l[0], dim=0, keepdim=True) for l in self.loss_list]).numpy(),
                                          columns=loss_names)
        self.test_loss_df = pd.DataFrame(torch.cat([torch.mean(l[1], dim=0, keepdim=True) for l in self.loss_list]).numpy(),
                                         columns=loss_names)
