"""Parameter containers for the MADE conditioner (host side, PyTorch = plumbing).

Mirrors what naz gets from pyro (`pyro.nn.{Conditional}AutoRegressiveNN`, `MaskedLinear`) closely
enough that the reference's weight formats keep working: parameters are named
``nn.layers.{k}.weight`` / ``.bias`` under each transform (src/naz/trainers/train_flows.py:20-71),
``arn.masks`` / ``arn.permutation`` / ``arn.layers`` are what ``torch_to_jax`` reads
(src/naz/flows/bflow_jax_maf.py:26-46).  These modules hold weights only — they have no forward():
evaluation happens in libnazb (CUDA), never in PyTorch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn


def sample_mask_indices(input_dim: int, hidden_dim: int) -> torch.Tensor:
    """pyro `sample_mask_indices(simple=True)`; twin: bflow_jax_maf.py:48-50."""
    return torch.round(torch.linspace(1, input_dim, steps=hidden_dim, device="cpu"))


def create_mask(input_dim: int, context_dim: int, hidden_dims: Sequence[int], permutation: torch.Tensor,
                output_dim_multiplier: int):
    """MADE masks, pyro `create_mask`; twin: bflow_jax_maf.py:52-72."""
    permutation = permutation.cpu()
    var_index = torch.empty(permutation.shape, dtype=torch.float32)
    var_index[permutation] = torch.arange(input_dim, dtype=torch.float32)
    input_indices = torch.cat((torch.zeros(context_dim), 1 + var_index))
    if context_dim > 0:
        hidden_indices = [sample_mask_indices(input_dim, h) - 1 for h in hidden_dims]
    else:
        hidden_indices = [sample_mask_indices(input_dim - 1, h) for h in hidden_dims]
    output_indices = (var_index + 1).repeat(output_dim_multiplier)
    mask_skip = (output_indices.unsqueeze(-1) > input_indices.unsqueeze(0)).float()
    masks = [(hidden_indices[0].unsqueeze(-1) >= input_indices.unsqueeze(0)).float()]
    for i in range(1, len(hidden_dims)):
        masks.append((hidden_indices[i].unsqueeze(-1) >= hidden_indices[i - 1].unsqueeze(0)).float())
    masks.append((output_indices.unsqueeze(-1) > hidden_indices[-1].unsqueeze(0)).float())
    return masks, mask_skip


class MaskedLinear(nn.Linear):
    """Weight container with its MADE mask (pyro.nn.MaskedLinear).  The mask is folded into the packed
    weights once per pack (libnazb), not re-multiplied on every call."""

    def __init__(self, in_features: int, out_features: int, mask: torch.Tensor, bias: bool = True):
        super().__init__(in_features, out_features, bias)
        self.register_buffer("mask", mask.data)

    def forward(self, _input):  # pragma: no cover - deliberate
        raise RuntimeError("naz_b200 conditioners are evaluated by libnazb (CUDA); there is no PyTorch forward path")


class ConditionalAutoRegressiveNN(nn.Module):
    def __init__(self, input_dim: int, context_dim: int, hidden_dims: List[int], param_dims: List[int] = [1, 1],
                 permutation: Optional[torch.Tensor] = None, skip_connections: bool = False, nonlinearity=None,
                 dropout_p: Optional[float] = None):
        super().__init__()
        if skip_connections:
            raise NotImplementedError("skip connections are never enabled by naz (transforms.py:142-152)")
        if nonlinearity is not None and not isinstance(nonlinearity, nn.Tanh):
            raise NotImplementedError("libnazb implements naz's default activation, nn.Tanh() (transforms.py:133,165)")
        if min(hidden_dims) < input_dim:
            raise ValueError("Hidden dimension must not be less than input dimension.")   # pyro's check
        self.input_dim, self.context_dim = input_dim, context_dim
        self.hidden_dims = list(hidden_dims)
        self.param_dims = list(param_dims)
        self.count_params = len(param_dims)
        self.output_multiplier = sum(param_dims)
        self.all_ones = all(p == 1 for p in param_dims)
        self.dropout_p = dropout_p
        if permutation is None:
            P = torch.randperm(input_dim, device="cpu").to(torch.int64)
        else:
            P = permutation.type(dtype=torch.int64).cpu()
        self.register_buffer("permutation", P)
        self.masks, self.mask_skip = create_mask(input_dim, context_dim, hidden_dims, self.permutation,
                                                 self.output_multiplier)
        layers = [MaskedLinear(input_dim + context_dim, hidden_dims[0], self.masks[0])]
        for i in range(1, len(hidden_dims)):
            layers.append(MaskedLinear(hidden_dims[i - 1], hidden_dims[i], self.masks[i]))
        layers.append(MaskedLinear(hidden_dims[-1], input_dim * self.output_multiplier, self.masks[-1]))
        self.layers = nn.ModuleList(layers)
        self.skip_layer = None

    def get_permutation(self):
        return self.permutation

    def forward(self, *a, **k):  # pragma: no cover - deliberate
        raise RuntimeError("naz_b200 conditioners are evaluated by libnazb (CUDA); there is no PyTorch forward path")


class AutoRegressiveNN(ConditionalAutoRegressiveNN):
    def __init__(self, input_dim, hidden_dims, **kw):
        super().__init__(input_dim, 0, hidden_dims, **kw)


class ConditionalDenseNN(nn.Module):
    """Weight container of pyro's `(Conditional)DenseNN` — the hyper-network of the coupling flow (src/naz/flows/transforms.py:
    219-224): plain `nn.Linear` layers `[context | x1] -> hidden... -> sum(param_dims)`, named `layers.{k}` as upstream.  No
    forward(): `NormalizingFlow` maps the coupling layer onto the engine's masked-conditioner format when it packs
    (`flows/transforms.py::SplineCoupling.as_made`)."""

    def __init__(self, input_dim: int, context_dim: int, hidden_dims: List[int], param_dims: List[int] = [1, 1], nonlinearity=None,
                 dropout_p: Optional[float] = None):
        super().__init__()
        if nonlinearity is not None and not isinstance(nonlinearity, nn.Tanh):
            raise NotImplementedError("libnazb implements naz's default activation, nn.Tanh() (transforms.py:201)")
        self.input_dim, self.context_dim = input_dim, context_dim
        self.hidden_dims, self.param_dims = list(hidden_dims), list(param_dims)
        self.dropout_p = dropout_p
        dims = [input_dim + context_dim] + self.hidden_dims + [sum(self.param_dims)]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])

    def forward(self, *a, **k):  # pragma: no cover - deliberate
        raise RuntimeError("naz_b200 conditioners are evaluated by libnazb (CUDA); there is no PyTorch forward path")


class DenseNN(ConditionalDenseNN):
    def __init__(self, input_dim, hidden_dims, **kw):
        super().__init__(input_dim, 0, hidden_dims, **kw)
