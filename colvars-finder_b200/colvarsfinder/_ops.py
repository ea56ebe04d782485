"""Glue between the task classes and libcvf_sm100.so: descriptors, flat parameter storage, the autograd
nodes that stand where the reference builds its autograd graph, and the one collective per pass.

Data-parallel layout: every rank holds a shard of the frames in its own HBM and the full (replicated)
parameters.  The loss couples frames only through a handful of batch sums, so each pass ends with ONE
``all_reduce(SUM)`` over NCCL -- fp64 batch sums after pass 1, fp64 gradient sums after pass 2 -- and every
rank then runs the same tiny combine kernel / optimizer step.  With one process the collective is skipped.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .nn import chain_spec


# ------------------------------------------------------------------------------------------ distributed
def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks on the current stream (no-op for a single process)."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_range(n: int, r: int, w: int):
    """Contiguous, near-equal frame shard [lo, hi) of rank r out of w."""
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


# ------------------------------------------------------------------------------------------ descriptors
class PreprocSpec:
    """``cvf_preproc`` for a pp_layer: Identity, utils.Align, utils.FeatureMap or utils.Preprocessing."""

    def __init__(self, pp_layer, frame_shape, device, diag_coeff):
        from . import utils
        self.device = torch.device(device)
        self._keep = []
        self.alignment_elided = False
        s = _lib.Preproc()
        frame_shape = tuple(int(v) for v in frame_shape)
        tot_dim = int(np.prod(frame_shape))
        diag = None
        if diag_coeff is not None:
            diag = torch.as_tensor(diag_coeff).detach().to("cpu", torch.float32).reshape(-1)
            if diag.numel() != tot_dim:
                raise RuntimeError(f"diag_coeff has {diag.numel()} entries, the state has {tot_dim} coordinates")
            if bool((diag == 1.0).all()):
                diag = None
        if isinstance(pp_layer, torch.nn.Identity):
            if len(frame_shape) != 1:
                raise RuntimeError("Identity pre-processing needs a flat [n, d] trajectory")
            s.kind, s.dim, s.d_r = 0, tot_dim, tot_dim
            if diag is not None:
                s.diag = self._dev(diag)
            self.d_r = tot_dim
        else:
            if isinstance(pp_layer, utils.Align):
                align, fmap = pp_layer, None
            elif isinstance(pp_layer, utils.FeatureMap):
                align, fmap = None, pp_layer
            elif isinstance(pp_layer, utils.Preprocessing):
                align, fmap = pp_layer.align, pp_layer.feature_mapper
            else:
                raise RuntimeError(
                    f"pp_layer of type {type(pp_layer).__name__} is outside the supported envelope of the CUDA step: use "
                    "torch.nn.Identity, colvarsfinder.utils.Align, FeatureMap or Preprocessing (no PyTorch fallback)")
            if len(frame_shape) != 2 or frame_shape[1] != 3:
                raise RuntimeError(f"molecular pre-processing needs [n, N, 3] frames, got frame shape {frame_shape}")
            n_atoms = frame_shape[0]
            if fmap is None:
                records = [(_lib.FEAT_POSITION, [a]) for a in range(n_atoms)]
            else:
                records = fmap.records()
            used, pos = [], {}
            for _, atoms in records:
                for a in atoms:
                    if not 0 <= a < n_atoms:
                        raise RuntimeError(f"feature atom {a} outside the {n_atoms} input atoms")
                    if a not in pos:
                        pos[a] = len(used)
                        used.append(a)
            n_feat_atoms = len(used)
            positions_only = all(t == _lib.FEAT_POSITION for t, _ in records) and len(records) == n_feat_atoms
            align_idx = [] if align is None else [int(a) for a in align.align_idx.cpu().tolist()]
            # Bond lengths, angles and dihedrals do not change under the rigid motion the alignment applies, and with
            # diag_coeff = 1 neither does |grad_x f|^2: r(x) and its Jacobian products are evaluated on the raw frame and
            # the Kabsch step (with its Jacobian) drops out of the training step.  (pp_layer(x) itself still aligns.)
            self.alignment_elided = bool(align_idx) and diag is None and all(t != _lib.FEAT_POSITION for t, _ in records)
            if self.alignment_elided:
                align_idx = []
            for a in align_idx:
                if not 0 <= a < n_atoms:
                    raise RuntimeError(f"alignment atom {a} outside the {n_atoms} input atoms")
                if a not in pos:
                    pos[a] = len(used)
                    used.append(a)
            positions_only = positions_only and len(used) == n_feat_atoms
            s.kind, s.n_atoms, s.n_used = 1, n_atoms, len(used)
            s.used_atoms = self._dev(torch.tensor(used, dtype=torch.int32))
            s.n_align = len(align_idx)
            if align_idx:
                s.align_used = self._dev(torch.tensor([pos[a] for a in align_idx], dtype=torch.int32))
                s.ref = self._dev(align.ref_pos.detach().to("cpu", torch.float32).reshape(-1))
            table = []
            d_r = 0
            for t, atoms in records:
                table.append([t] + [pos[a] for a in atoms] + [0] * (4 - len(atoms)))
                d_r += {_lib.FEAT_POSITION: 3, _lib.FEAT_BOND: 1, _lib.FEAT_ANGLE: 1, _lib.FEAT_DIHEDRAL: 2}[t]
            s.n_feat, s.d_r = len(records), d_r
            for t, _ in records:
                s.n_feat_by_type[t] += 1
            reads = {}
            for t, atoms in records:
                for a in atoms:
                    reads[a] = reads.get(a, 0) + (2 if t == _lib.FEAT_POSITION else 1)
            s.n_self_records = sum(1 for t, atoms in records
                                   if t != _lib.FEAT_POSITION and any(reads[a] == 1 for a in atoms))
            s.n_shared_atoms = sum(1 for c in reads.values() if c > 1)
            s.feat = self._dev(torch.tensor(table, dtype=torch.int32).reshape(-1))
            s.positions_only = 1 if positions_only else 0
            s.used_identity = 1 if used == list(range(n_atoms)) else 0
            if diag is not None:
                g = diag.reshape(n_atoms, 3)[torch.tensor(used, dtype=torch.long)].reshape(-1).contiguous()
                s.diag = self._dev(g)
            self.d_r = d_r
        self.struct = s

    def _dev(self, t):
        t = t.contiguous().to(self.device)
        self._keep.append(t)
        return t.data_ptr()

    def struct_ptr(self):
        return C.byref(self.struct)


class FlatParams:
    """All parameters of a set of Linear stacks in ONE contiguous fp32 device buffer, in torch's
    ``parameters()`` order; the modules' parameters are re-pointed to views of it, so the optimizer updates the
    buffer the kernels read and nothing is packed per step."""

    def __init__(self, chains, device):
        self.params = []
        self.dims, self.acts = [], []
        for seq in chains:
            dims, acts, lins = chain_spec(seq)
            self.dims.append(dims), self.acts.append(acts)
            for lin in lins:
                self.params += [lin.weight, lin.bias]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=device)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        self.n = n
        self.rebind()

    def rebind(self):
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view(p.shape)
                if p.data.data_ptr() != view.data_ptr() or p.data.device != self.flat.device:
                    view.copy_(p.data.to(device=self.flat.device, dtype=torch.float32))
                    p.data = view

    def check(self):
        """Parameters must still alias the flat buffer (e.g. model.to() would break it): re-bind if not."""
        base = self.flat.data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.data.data_ptr() != base + 4 * off:
                self.rebind()
                return

    def split(self, g):
        return tuple(g[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_batch(X, weight, what):
    if not (X.is_cuda and weight.is_cuda):
        raise RuntimeError(f"{what}: tensors must live on a CUDA device (this build has no CPU path)")
    if X.dtype != torch.float32:
        raise RuntimeError(f"{what}: data must be float32, got {X.dtype}")
    X = X.detach().contiguous()
    weight = weight.detach().to(torch.float32).contiguous()
    if weight.numel() != X.shape[0]:
        raise RuntimeError(f"{what}: {weight.numel()} weights for {X.shape[0]} states")
    return X, weight


# ------------------------------------------------------------------------------------------ eigenfunctions
class EigenContext:
    """Everything constant across steps for one EigenFunctionTask."""

    def __init__(self, model, pp_layer, frame_shape, device, alpha, eig_w, beta, diag_coeff, sort):
        self.device = torch.device(device)
        self.k = len(model.eigen_funcs)
        if self.k > _lib.MAX_K:
            raise RuntimeError(f"k = {self.k} eigenfunctions is outside the supported envelope (<= {_lib.MAX_K})")
        self.flat = FlatParams(list(model.eigen_funcs), self.device)
        if any(d != self.flat.dims[0] for d in self.flat.dims):
            raise RuntimeError("all eigenfunction networks must share one architecture")
        self.mlp = _lib.make_mlp(self.flat.dims[0], self.flat.acts[0])
        self.spec = PreprocSpec(pp_layer, frame_shape, self.device, diag_coeff)
        if self.spec.d_r != self.flat.dims[0][0]:
            raise RuntimeError(f"pre-processing layer outputs {self.spec.d_r} features, the networks take {self.flat.dims[0][0]}")
        self.alpha, self.beta, self.sort = float(alpha), float(beta), 1 if sort else 0
        self.eig_w = (C.c_double * self.k)(*[float(v) for v in eig_w])
        L = _lib.lib()
        self.n_stats = L.cvf_eigen_num_stats(self.k)
        self.n_comb = L.cvf_eigen_num_combine(self.k)
        self.n_per_net = int(L.cvf_mlp_param_count(C.byref(self.mlp)))
        self.fast_path = bool(L.cvf_eigen_path(self.spec.struct_ptr(), C.byref(self.mlp), self.k) == 1)
        # scratch slots: 0 for X, 1 for the time-lagged X of the transfer-operator loss (both must survive until backward)
        self._ws = {}

    def _workspace_for(self, B, slot=0):
        """Scratch of the step kernels, grown to the largest batch seen (PyTorch's caching allocator owns it)."""
        ws = self._ws.get(slot)
        if ws is None or B > ws["frames"]:
            need = int(_lib.lib().cvf_eigen_workspace_bytes(self.spec.struct_ptr(), C.byref(self.mlp), self.k, B))
            if need <= 0:
                raise RuntimeError("cvf_eigen_workspace_bytes: unsupported configuration")
            self._ws[slot] = None
            ws = {"buf": torch.empty(need, dtype=torch.uint8, device=self.device), "bytes": need, "frames": B, "key": None}
            self._ws[slot] = ws
        return ws

    def _key(self, X, weight):
        return (X.data_ptr(), X.shape[0], X._version, weight.data_ptr(), weight._version, self.flat.flat._version)

    def stats(self, X, weight, slot=0):
        """Pass 1 on this rank's frames -> (y [k,B] fp32, stats fp64)."""
        L = _lib.lib()
        B = X.shape[0]
        ws = self._workspace_for(B, slot)
        y = torch.empty(self.k, B, dtype=torch.float32, device=self.device)
        stats = torch.empty(self.n_stats, dtype=torch.float64, device=self.device)
        _lib.check(L.cvf_eigen_stats(X.data_ptr(), weight.data_ptr(), B, self.spec.struct_ptr(), C.byref(self.mlp), self.k,
                                     self.flat.flat.data_ptr(), y.data_ptr(), stats.data_ptr(), ws["buf"].data_ptr(), ws["bytes"],
                                     _stream()), "cvf_eigen_stats")
        ws["key"] = self._key(X, weight)
        return y, stats

    def combine(self, stats):
        comb = torch.empty(self.n_comb, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().cvf_eigen_combine(stats.data_ptr(), self.k, self.alpha, self.eig_w, self.beta, self.sort,
                                                comb.data_ptr(), _stream()), "cvf_eigen_combine")
        return comb

    def grads(self, X, weight, y, comb, seed_extra=None, slot=0):
        """Pass 2 on this rank's frames -> fp64 gradient sums [k * n_per_net]."""
        g = torch.empty(self.k * self.n_per_net, dtype=torch.float64, device=self.device)
        ws = self._workspace_for(X.shape[0], slot)
        # the scratch still holds pass 1's intermediates iff the last stats() call saw these very tensors and parameters
        valid = 1 if ws["key"] is not None and ws["key"] == self._key(X, weight) else 0
        _lib.check(_lib.lib().cvf_eigen_grad(X.data_ptr(), weight.data_ptr(), X.shape[0], self.spec.struct_ptr(),
                                             C.byref(self.mlp), self.k, self.flat.flat.data_ptr(), y.data_ptr(),
                                             comb.data_ptr(), None if seed_extra is None else seed_extra.data_ptr(),
                                             g.data_ptr(), ws["buf"].data_ptr(), ws["bytes"], valid, _stream()),
                   "cvf_eigen_grad")
        # pass 2 consumes the scratch (it leaves v = J J^T u where pass 1 left u): a second backward recomputes it
        ws["key"] = None
        return g

    def tlag_terms(self, y, y_lag, weight, coef=None):
        """sum_f w (y' - y)^2 per network (coef None), or the per-frame seed term coef_i w (y_i - y'_i) of the backward pass."""
        L, B = _lib.lib(), y.shape[1]
        if coef is None:
            sx = torch.empty(self.k, dtype=torch.float64, device=self.device)
            ws = self._workspace_for(B, 0)
            _lib.check(L.cvf_eigen_tlag_terms(y.data_ptr(), y_lag.data_ptr(), weight.data_ptr(), B, self.k, None, sx.data_ptr(),
                                              None, ws["buf"].data_ptr(), ws["bytes"], _stream()), "cvf_eigen_tlag_terms")
            return sx
        extra = torch.empty(self.k, B, dtype=torch.float32, device=self.device)
        _lib.check(L.cvf_eigen_tlag_terms(y.data_ptr(), y_lag.data_ptr(), weight.data_ptr(), B, self.k, coef.data_ptr(), None,
                                          extra.data_ptr(), None, 0, _stream()), "cvf_eigen_tlag_terms")
        return extra

    def tlag_combine(self, stats, stats_lag, sx, tau):
        comb = torch.empty(self.n_comb + self.k, dtype=torch.float64, device=self.device)
        comb_lag = torch.empty(self.n_comb, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().cvf_eigen_tlag_combine(stats.data_ptr(), stats_lag.data_ptr(), sx.data_ptr(), self.k, self.alpha,
                                                     self.eig_w, float(tau), self.sort, comb.data_ptr(), comb_lag.data_ptr(),
                                                     _stream()), "cvf_eigen_tlag_combine")
        return comb, comb_lag


class _EigenLoss(torch.autograd.Function):
    """loss = EigenFunctionTask.loss_func (reference core.py:387-457); backward = reference core.py:517."""

    @staticmethod
    def forward(ctx, ectx, X, weight, *params):
        with torch.cuda.device(ectx.device):
            y, stats = ectx.stats(X, weight)
            allreduce_sum_(stats)
            comb = ectx.combine(stats)
        ctx.ectx, ctx.X, ctx.weight, ctx.y, ctx.comb = ectx, X, weight, y, comb
        k = ectx.k
        out32 = comb[:3 + k].to(torch.float32)
        loss, obj, pen, eig = out32[0], out32[1], out32[2], out32[3:3 + k]
        cvec = comb[3 + k:3 + 2 * k].to(torch.int64)
        ctx.mark_non_differentiable(obj, pen, eig, cvec)
        return loss, eig, obj, pen, cvec

    @staticmethod
    def backward(ctx, g_loss, *unused):
        ectx = ctx.ectx
        with torch.cuda.device(ectx.device):
            g = ectx.grads(ctx.X, ctx.weight, ctx.y, ctx.comb)
            allreduce_sum_(g)
            g32 = (g * g_loss.to(torch.float64)).to(torch.float32)
        return (None, None, None) + ectx.flat.split(g32)


def eigen_loss(ectx: EigenContext, X, weight):
    X, weight = _check_batch(X, weight, "EigenFunctionTask.loss_func")
    ectx.flat.check()
    return _EigenLoss.apply(ectx, X, weight, *ectx.flat.params)


class _EigenLagLoss(torch.autograd.Function):
    """Transfer-operator branch of EigenFunctionTask.loss_func (reference core.py:412-416,428,440): forward passes on X and
    on the time-lagged X, ONE all-reduce of {sums of y, sums of y', sum w (y'-y)^2}; backward = two first-order passes."""

    @staticmethod
    def forward(ctx, ectx, tau, X, weight, Xl, wl, *params):
        k = ectx.k
        with torch.cuda.device(ectx.device):
            y, st = ectx.stats(X, weight, slot=0)
            yl, stl = ectx.stats(Xl, wl, slot=1)
            sx = ectx.tlag_terms(y, yl, weight)
            packed = torch.cat([st, stl, sx])
            allreduce_sum_(packed)
            st, stl, sx = packed[:ectx.n_stats], packed[ectx.n_stats:2 * ectx.n_stats], packed[2 * ectx.n_stats:]
            comb, comb_lag = ectx.tlag_combine(st.contiguous(), stl.contiguous(), sx.contiguous(), tau)
        ctx.ectx, ctx.saved = ectx, (X, weight, Xl, wl, y, yl, comb, comb_lag)
        out32 = comb[:3 + k].to(torch.float32)
        loss, obj, pen, eig = out32[0], out32[1], out32[2], out32[3:3 + k]
        cvec = comb[3 + k:3 + 2 * k].to(torch.int64)
        ctx.mark_non_differentiable(obj, pen, eig, cvec)
        return loss, eig, obj, pen, cvec

    @staticmethod
    def backward(ctx, g_loss, *unused):
        ectx = ctx.ectx
        X, weight, Xl, wl, y, yl, comb, comb_lag = ctx.saved
        with torch.cuda.device(ectx.device):
            extra = ectx.tlag_terms(y, yl, weight, coef=comb[ectx.n_comb:])
            g = ectx.grads(X, weight, y, comb, seed_extra=extra, slot=0)
            g += ectx.grads(Xl, wl, yl, comb_lag, seed_extra=extra.neg_(), slot=1)
            allreduce_sum_(g)
            g32 = (g * g_loss.to(torch.float64)).to(torch.float32)
        return (None, None, None, None, None, None) + ectx.flat.split(g32)


def eigen_lag_loss(ectx: EigenContext, tau, X, weight, X_lagged, weight_lagged):
    X, weight = _check_batch(X, weight, "EigenFunctionTask.loss_func")
    X_lagged, weight_lagged = _check_batch(X_lagged, weight_lagged, "EigenFunctionTask.loss_func (time-lagged data)")
    if X_lagged.shape != X.shape:
        raise RuntimeError(f"time-lagged batch has shape {tuple(X_lagged.shape)}, the batch {tuple(X.shape)}")
    ectx.flat.check()
    return _EigenLagLoss.apply(ectx, tau, X, weight, X_lagged, weight_lagged, *ectx.flat.params)


# ------------------------------------------------------------------------------------------ autoencoder
class AEContext:
    def __init__(self, model, device):
        self.device = torch.device(device)
        self.flat = FlatParams([model.encoder, model.decoder], self.device)
        e_dims, d_dims = self.flat.dims
        self.mlp = _lib.make_mlp(list(e_dims) + list(d_dims[1:]), list(self.flat.acts[0]) + list(self.flat.acts[1]))
        L = _lib.lib()
        self.n_params = int(L.cvf_mlp_param_count(C.byref(self.mlp)))
        assert self.n_params == self.flat.n
        self.ws_bytes, self.ws_frames, self.workspace = 0, 0, None
        self.d_r = e_dims[0]

    def _workspace_for(self, B):
        if self.workspace is None or B > self.ws_frames:
            need = int(_lib.lib().cvf_ae_workspace_bytes(C.byref(self.mlp), B))
            if need <= 0:
                raise RuntimeError("cvf_ae_workspace_bytes: " + _lib.lib().cvf_last_error_string().decode("utf-8", "replace"))
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.ws_bytes, self.ws_frames = need, B
        return self.workspace.data_ptr()

    def step(self, F, weight, want_grad):
        """One pass over this rank's frames -> fp64 [2 + n_params]: sum w|e|^2, sum w, gradient of the first sum."""
        buf = torch.empty(2 + (self.n_params if want_grad else 0), dtype=torch.float64, device=self.device)
        gptr = buf.data_ptr() + 16 if want_grad else None
        ws = self._workspace_for(F.shape[0])
        _lib.check(_lib.lib().cvf_ae_step(F.data_ptr(), weight.data_ptr(), F.shape[0], C.byref(self.mlp),
                                          self.flat.flat.data_ptr(), buf.data_ptr(), gptr, ws, self.ws_bytes, _stream()),
                   "cvf_ae_step")
        return buf


class _AELoss(torch.autograd.Function):
    """weighted MSE (reference core.py:652-666) with its parameter gradient computed in the same pass."""

    @staticmethod
    def forward(ctx, actx, F, weight, want_grad, *params):
        with torch.cuda.device(actx.device):
            buf = actx.step(F, weight, want_grad)
            allreduce_sum_(buf)
        ctx.actx, ctx.buf, ctx.want_grad = actx, buf, want_grad
        return (buf[0] / buf[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g_loss):
        if not ctx.want_grad:
            raise RuntimeError("weighted_MSE_loss was evaluated without gradients (torch.no_grad or frozen parameters)")
        buf = ctx.buf
        g32 = (buf[2:] * (g_loss.to(torch.float64) / buf[1])).to(torch.float32)
        return (None, None, None, None) + ctx.actx.flat.split(g32)


def ae_loss(actx: AEContext, F, weight):
    F, weight = _check_batch(F, weight, "AutoEncoderTask.weighted_MSE_loss")
    if F.dim() != 2 or F.shape[1] != actx.d_r:
        raise RuntimeError(f"autoencoder input must be [B,{actx.d_r}], got {tuple(F.shape)}")
    actx.flat.check()
    want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in actx.flat.params)
    return _AELoss.apply(actx, F, weight, want_grad, *actx.flat.params)
