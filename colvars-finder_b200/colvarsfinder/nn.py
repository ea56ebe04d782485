"""Network containers with the reference's module layout (colvarsfinder/nn.py of the reference).

They stay ordinary ``torch.nn.Module`` objects so that ``state_dict`` keys (``eigen_funcs.{i}.{j}.weight``,
``encoder.{j}.weight``, ``decoder.{j}.weight``; j counts from 1 -- reference nn.py:55-57,84-85,272),
``torch.manual_seed`` reproducibility of the initial weights, checkpoints and ``save_model`` keep working.
Training never calls their ``forward``: the task classes hand the parameters to the CUDA step kernels.
``forward`` itself is what ``torch.nn.Sequential`` gives (used for inference / export of the trained CVs).
"""
from __future__ import annotations

import re

import numpy as np
import torch


def create_sequential_nn(layer_dims, activation=torch.nn.Tanh()):
    """Feed-forward stack: Linear layers named '1'..'m', the SAME activation instance between them
    (named 'activation i'), no activation after the last layer (reference nn.py:29-59)."""
    assert len(layer_dims) >= 2, \
        'Error: at least 2 layers are needed to define a neural network (length={})!'.format(len(layer_dims))
    net = torch.nn.Sequential()
    n_linear = len(layer_dims) - 1
    for j in range(1, n_linear + 1):
        net.add_module(str(j), torch.nn.Linear(layer_dims[j - 1], layer_dims[j]))
        if j < n_linear:
            net.add_module('activation %d' % j, activation)
    return net


def activation_kind(m) -> int:
    """CVF_ACT_* (include/cvf.h) of an activation module; raises for modules outside the kernel envelope.  The kernels keep only a
    layer's output, so they take the activations whose first two derivatives are functions of the output: Tanh (thread-private and
    tensor-core kernels), Sigmoid, Softplus (beta 1, threshold 20), ELU (alpha 1), ReLU (general kernels)."""
    if isinstance(m, torch.nn.Tanh):
        return 1
    if isinstance(m, torch.nn.Sigmoid):
        return 2
    if isinstance(m, torch.nn.Softplus) and m.beta == 1 and m.threshold == 20:
        return 3
    if isinstance(m, torch.nn.ELU) and m.alpha == 1.0:
        return 4
    if isinstance(m, torch.nn.ReLU):
        return 5
    raise RuntimeError(
        f"activation {type(m).__name__} is outside the supported envelope of the CUDA step (torch.nn.Tanh, Sigmoid, "
        "Softplus(beta=1), ELU(alpha=1), ReLU); there is no PyTorch fallback")


def chain_spec(seq: torch.nn.Sequential):
    """(dims, acts, linear modules) of a stack built by create_sequential_nn -- acts[l] is the CVF_ACT_* kind after layer l (0:
    none); raises outside the kernel envelope."""
    dims, acts, lins = [], [], []
    mods = list(seq._modules.values())   # children() would drop the repeated (shared) activation instance
    i = 0
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, torch.nn.Linear):
            raise RuntimeError(f"unsupported module {type(m).__name__} in network (Linear/Tanh stacks only)")
        if m.bias is None:
            raise RuntimeError("Linear layers without bias are outside the supported envelope")
        if not dims:
            dims.append(m.in_features)
        dims.append(m.out_features)
        lins.append(m)
        act = 0
        if i + 1 < len(mods) and not isinstance(mods[i + 1], torch.nn.Linear):
            act = activation_kind(mods[i + 1])
            i += 1
        acts.append(act)
        i += 1
    return dims, acts, lins


class AutoEncoder(torch.nn.Module):
    """encoder / decoder stacks; ``forward = decoder(encoder(x))`` (reference nn.py:61-114)."""

    def __init__(self, e_layer_dims, d_layer_dims, activation=torch.nn.Tanh()):
        super().__init__()
        assert e_layer_dims[-1] == d_layer_dims[0], "ouput dimension of encoder and input dimension of decoder do not match!"
        self.encoder = create_sequential_nn(e_layer_dims, activation)
        self.decoder = create_sequential_nn(d_layer_dims, activation)
        self.encoded_dim = e_layer_dims[-1]
        self._num_encoder_layer = len(e_layer_dims) - 1

    def get_params_of_cv(self, cv_idx):
        """[name, parameter] pairs of the encoder; the last layer is cut down to row cv_idx."""
        assert 0 <= cv_idx < self.encoded_dim, f"index {cv_idx} exceeded the range [0, {self.encoded_dim-1}]!"
        out = []
        for name, param in self.encoder.named_parameters():
            if int(re.search(r'\d+', name).group()) < self._num_encoder_layer:
                out.append([name, param])
            else:
                out.append([name, param[cv_idx:cv_idx + 1, ...]])
        return out

    def forward(self, inp):
        return self.decoder(self.encoder(inp))


class RegAutoEncoder(torch.nn.Module):
    """Regularised autoencoder: encoder, decoder and K regulariser networks on the encoded variables
    (reference nn.py:116-203).  ``state_dict`` keys ``encoder.{j}.*``, ``decoder.{j}.*``, ``reg.{i}.{j}.*``."""

    def __init__(self, e_layer_dims, d_layer_dims, reg_layer_dims, K, activation=torch.nn.Tanh()):
        super().__init__()
        assert e_layer_dims[-1] == d_layer_dims[0], "ouput dimension of encoder and input dimension of decoder do not match!"
        self.num_reg = K
        assert self.num_reg == 0 or e_layer_dims[-1] == reg_layer_dims[0], \
            "ouput dimension of encoder and input dimension of regulator part do not match!"
        self.encoder = create_sequential_nn(e_layer_dims, activation)
        self.decoder = create_sequential_nn(d_layer_dims, activation)
        self.encoded_dim = e_layer_dims[-1]
        self._num_encoder_layer = len(e_layer_dims) - 1
        if self.num_reg > 0:
            self.reg = torch.nn.ModuleList([create_sequential_nn(reg_layer_dims, activation) for _ in range(self.num_reg)])
        else:
            self.reg = None

    def get_params_of_cv(self, cv_idx):
        """[name, parameter] pairs of the encoder; the last layer is cut down to row cv_idx."""
        assert 0 <= cv_idx < self.encoded_dim, f"index {cv_idx} exceeded the range [0, {self.encoded_dim-1}]!"
        out = []
        for name, param in self.encoder.named_parameters():
            if int(re.search(r'\d+', name).group()) < self._num_encoder_layer:
                out.append([name, param])
            else:
                out.append([name, param[cv_idx:cv_idx + 1, ...]])
        return out

    def forward_ae(self, inp):
        return self.decoder(self.encoder(inp))

    def forward_reg(self, inp):
        assert self.num_reg > 0, 'number of regularizers is not positive.'
        encoded = self.encoder(inp)
        return torch.cat([f(encoded) for f in self.reg], dim=1)

    def forward(self, inp):
        encoded = self.encoder(inp)
        return torch.cat((self.decoder(encoded), torch.cat([f(encoded) for f in self.reg], dim=1)), dim=1)


class RegModel(torch.nn.Module):
    """The regularisers of a :class:`RegAutoEncoder` as functions of the input, reordered by ``cvec``
    (reference nn.py:205-239).  Shares the encoder / regulariser modules with ``reg_ae``."""

    def __init__(self, reg_ae, cvec):
        super().__init__()
        assert reg_ae.num_reg > 0, 'number of regularizers is not positive.'
        assert len(cvec) == reg_ae.num_reg, 'length of cvec doesn\'t equal to number of regularizers'
        assert (np.sort(np.asarray([int(c) for c in cvec])) == np.arange(reg_ae.num_reg)).all(), \
            f'cvec should be a permutation of 0,1,...,{len(cvec)-1}.'
        self.encoder = reg_ae.encoder
        self.reg = reg_ae.reg
        self.cvec = cvec
        self.encoded_dim = reg_ae.encoded_dim
        self.num_reg = reg_ae.num_reg

    def forward(self, inp):
        encoded = self.encoder(inp)
        return torch.cat([self.reg[int(idx)](encoded) for idx in self.cvec], dim=1)


class EigenFunctions(torch.nn.Module):
    """k independent scalar networks of one architecture; ``forward`` concatenates them to [l, k]
    (reference nn.py:242-293)."""

    def __init__(self, layer_dims, k, activation=torch.nn.Tanh()):
        super().__init__()
        assert layer_dims[-1] == 1, "each eigenfunction must be scalar-valued"
        self.eigen_funcs = torch.nn.ModuleList([create_sequential_nn(layer_dims, activation) for _ in range(k)])

    def get_params_of_cv(self, cv_idx):
        return [[name, param] for name, param in self.eigen_funcs[cv_idx].named_parameters()]

    def forward(self, inp):
        return torch.cat([f(inp) for f in self.eigen_funcs], dim=1)
