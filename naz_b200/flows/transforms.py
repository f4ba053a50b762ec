"""Factories mirroring src/naz/flows/transforms.py: same names, arguments and return triple
``(flow, transforms, nets)``; the transforms are weight containers whose arithmetic runs in libnazb."""
from __future__ import annotations

import torch
import torch.nn as nn

from .made import AutoRegressiveNN, ConditionalAutoRegressiveNN


def bounding_transform(x, low, high):
    """transforms.py:20-23: box -> unbounded by a logit of the box coordinate; returns (y, log|dy/dx| summed over dims).
    Host-side helper for API compatibility; inside log_prob the same map is fused into the kernels (transforms.cuh)."""
    width = high - low
    u = (x - low) / width                               # box coordinate in (0, 1); low / high broadcast over the batch
    log_jac = -(u.log() + torch.log1p(-u)).sum(-1) - width.log().sum()
    return torch.logit(u), log_jac


def inverse_bounding_transform(y, low, high):
    """transforms.py:25-27: unbounded -> box."""
    return low + (high - low) * torch.sigmoid(y)


class _ARTransform(nn.Module):
    """One flow layer: holds the conditioner as ``.nn`` (pyro's ConditionalAffineAutoregressive /
    ConditionalSplineAutoregressive keep it under that name; bflow_jax_maf.py:34 reads ``flow_layer.nn``)."""
    kind = "maf"

    def __init__(self, arn, **meta):
        super().__init__()
        self.nn = arn
        self.arn = arn
        self.meta = meta


class AffineAutoregressive(_ARTransform):
    kind = "maf"
    log_scale_min_clip, log_scale_max_clip = -5.0, 3.0


class SplineAutoregressive(_ARTransform):
    kind = "nsa"


class ComposeTransformModule(nn.ModuleList):
    """`flow` in the returned triple; `.parts` as torch's ComposeTransform exposes (mcdpflow.py:15)."""

    @property
    def parts(self):
        return list(self)


def _hidden_list(hidden_dim):
    return list(hidden_dim) if isinstance(hidden_dim, (list, tuple)) else [hidden_dim]


class Permute(nn.Module):
    """pyro's `T.Permute` as naz appends it after each flow layer when `random_perm=True` (transforms.py:155-156, :193-194):
    `y[..., k] = x[..., permutation[k]]`, log-det 0.  No kernel runs for it: `NormalizingFlow` folds the re-labelling
    into the neighbouring conditioners' weight columns / output rows / MADE order when it packs (flows/flow.py)."""
    kind = "permute"

    def __init__(self, permutation):
        super().__init__()
        self.register_buffer("permutation", torch.as_tensor(permutation).to(torch.int64).cpu())


class BatchNorm(nn.Module):
    """pyro's `T.BatchNorm` as naz appends it when `use_batchnorm=True` (transforms.py:157-158, :195-196), in its eval()
    form: sampling direction `y = (x - beta) / gamma_c * sqrt(moving_variance + eps) + moving_mean`, gamma_c = relu(gamma) +
    1e-6, log|dy/dx| = 0.5 log(moving_variance + eps) - log gamma_c per dimension.  Evaluated inside the SIMT flow kernel as a
    per-layer element-wise affine (`nazb_set_layer_affine`).  The train()-mode batch statistics belong to the training
    loop (out of scope): `log_prob` in train() mode raises."""
    kind = "batchnorm"

    def __init__(self, input_dim, momentum=0.1, epsilon=1e-5):
        super().__init__()
        self.input_dim, self.momentum, self.epsilon = input_dim, momentum, epsilon
        self.gamma = nn.Parameter(torch.ones(input_dim))
        self.beta = nn.Parameter(torch.zeros(input_dim))
        self.register_buffer("moving_mean", torch.zeros(input_dim))
        self.register_buffer("moving_variance", torch.ones(input_dim))

    @property
    def constrained_gamma(self):
        return torch.relu(self.gamma) + 1e-6

    def affine(self):
        """(a, b) of the sampling direction y = a x + b, float32 [D] each."""
        with torch.no_grad():
            a = torch.sqrt(self.moving_variance + self.epsilon) / self.constrained_gamma
            return a.float(), (self.moving_mean - self.beta * a).float()


def _extras(transforms, theta_dim, use_batchnorm, random_perm):
    if random_perm:
        transforms.append(Permute(torch.randperm(theta_dim)))
    if use_batchnorm:
        transforms.append(BatchNorm(theta_dim))


def masked_affine_autoregressive(theta_dim, condition_dim, hidden_dim, num_layers, activation=None, use_batchnorm=False,
                                 random_mask=True, random_perm=False, dropout_p=None):
    """transforms.py:133-160."""
    transforms, nets = [], []
    for _ in range(num_layers):
        perm = None if random_mask else torch.arange(theta_dim)
        arn = ConditionalAutoRegressiveNN(theta_dim, condition_dim, _hidden_list(hidden_dim), nonlinearity=activation,
                                          permutation=perm, dropout_p=dropout_p)
        nets.append(arn)
        transforms.append(AffineAutoregressive(arn))
        _extras(transforms, theta_dim, use_batchnorm, random_perm)
    return ComposeTransformModule(transforms), transforms, nets


def neural_spline_autoregressive(theta_dim, condition_dim, hidden_dim, num_layers, count_bins, order="quadratic",
                                 activation=None, use_batchnorm=False, random_mask=True, random_perm=False,
                                 dropout_p=None):
    """transforms.py:165-198."""
    if order == "linear":
        paramdim = [count_bins, count_bins, count_bins - 1, count_bins]
    elif order == "quadratic":
        paramdim = [count_bins, count_bins, count_bins - 1]
    else:
        raise ValueError(order)
    transforms, nets = [], []
    for _ in range(num_layers):
        perm = None if random_mask else torch.arange(theta_dim)
        arn = ConditionalAutoRegressiveNN(theta_dim, condition_dim, _hidden_list(hidden_dim), param_dims=paramdim,
                                          nonlinearity=activation, permutation=perm, dropout_p=dropout_p)
        nets.append(arn)
        transforms.append(SplineAutoregressive(arn, count_bins=count_bins, order=order, bound=3.0))
        _extras(transforms, theta_dim, use_batchnorm, random_perm)
    return ComposeTransformModule(transforms), transforms, nets


def neural_spline_coupling(*args, **kwargs):
    """transforms.py:201-236 cannot be constructed upstream (undefined names); out of scope."""
    raise NotImplementedError("'nsc' is unconstructible in the reference (transforms.py:201-236) and is out of scope")
