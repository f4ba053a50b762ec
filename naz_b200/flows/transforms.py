"""Factories mirroring src/naz/flows/transforms.py: same names, arguments and return triple
``(flow, transforms, nets)``; the transforms are weight containers whose arithmetic runs in libnazb."""
from __future__ import annotations

import torch
import torch.nn as nn

from .made import AutoRegressiveNN, ConditionalAutoRegressiveNN, ConditionalDenseNN


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


class _LowerSpline(nn.Module):
    """Free parameters of the element-wise spline on the first `split_dim` coordinates (pyro `T.Spline`, the `lower_spline`
    of `T.SplineCoupling`): same parameter names and initialisers as upstream."""

    def __init__(self, input_dim, count_bins, order):
        super().__init__()
        self.unnormalized_widths = nn.Parameter(torch.randn(input_dim, count_bins))
        self.unnormalized_heights = nn.Parameter(torch.randn(input_dim, count_bins))
        self.unnormalized_derivatives = nn.Parameter(torch.randn(input_dim, count_bins - 1))
        if order == "linear":
            self.unnormalized_lambdas = nn.Parameter(torch.rand(input_dim, count_bins))

    def groups(self, order):
        g = [self.unnormalized_widths, self.unnormalized_heights, self.unnormalized_derivatives]
        return g + ([self.unnormalized_lambdas] if order == "linear" else [])


class SplineCoupling(nn.Module):
    """`ConditionalSplineCoupling(...).condition(context)` / `T.SplineCoupling` (transforms.py:113-129, :226): the first
    `split_dim` coordinates pass through an element-wise spline with free parameters, the others through a spline whose
    parameters a dense hyper-network computes from [context | first part].

    No kernel of its own: a coupling layer IS a masked conditioner with a single hidden degree — every hidden unit sees the
    context and the first `split_dim` inputs, the outputs of the remaining coordinates see every hidden unit, the outputs of the
    first part see none (their biases are the free spline parameters).  `as_made` writes it in that form (identity MADE order,
    hyper-network rows re-indexed from pyro's dimension-major `[.., D - split, K]` to the engine's slot-major `m D + d`), and
    the incremental inverse then costs ONE conditioner pass per layer (all hidden units become final at stage `split_dim`)."""
    kind = "nsc"

    def __init__(self, input_dim, split_dim, dense_nn, count_bins=8, bound=3.0, order="quadratic"):
        super().__init__()
        if not 0 < split_dim < input_dim:
            raise ValueError("split_dim must lie strictly between 0 and input_dim")
        self.input_dim, self.split_dim, self.count_bins, self.bound, self.order = input_dim, split_dim, count_bins, bound, order
        self.nn = dense_nn
        self.lower_spline = _LowerSpline(split_dim, count_bins, order)

    def slot_groups(self):
        K = self.count_bins
        return [K, K, K - 1] + ([K] if self.order == "linear" else [])

    def made_masks(self, context_dim):
        D, s, C, H = self.input_dim, self.split_dim, context_dim, self.nn.hidden_dims
        M = sum(self.slot_groups())
        m0 = torch.zeros(H[0], C + D)
        m0[:, :C + s] = 1.0
        masks = [m0] + [torch.ones(H[i], H[i - 1]) for i in range(1, len(H))]
        mo = torch.zeros(M, D, H[-1])
        mo[:, s:, :] = 1.0
        return masks + [mo.reshape(M * D, H[-1])]

    def as_made(self, lins, lower_groups, context_dim):
        """lins: the hyper-network's [(W, b)] (optionally with leading draw axes), lower_groups: the free parameters
        [widths, heights, derivatives(, lambdas)] each [.., split_dim, K_g]  ->  [(W, b)] in the engine's conditioner format."""
        D, s, C = self.input_dim, self.split_dim, context_dim
        out = [(W, b) for (W, b) in lins]
        W0, b0 = lins[0]
        W0e = W0.new_zeros(W0.shape[:-1] + (C + D,))
        W0e[..., :C + s] = W0
        out[0] = (W0e, b0)
        Wl, bl = lins[-1]
        H = Wl.shape[-1]
        lead_w, lead_b = Wl.shape[:-2], bl.shape[:-1]
        rows_W, rows_b, off = [], [], 0
        for Kg, low in zip(self.slot_groups(), lower_groups):
            n = (D - s) * Kg
            up_W = Wl[..., off:off + n, :].reshape(lead_w + (D - s, Kg, H)).transpose(-3, -2)          # [.., Kg, D - s, H]
            up_b = bl[..., off:off + n].reshape(lead_b + (D - s, Kg)).transpose(-2, -1)                  # [.., Kg, D - s]
            lo_b = low.to(bl.dtype).transpose(-2, -1)                                                    # [.., Kg, s]
            lo_b = lo_b.expand(lead_b + lo_b.shape[-2:]) if lo_b.dim() < up_b.dim() else lo_b
            rows_W.append(torch.cat([up_W.new_zeros(lead_w + (Kg, s, H)), up_W], dim=-2))                # [.., Kg, D, H]
            rows_b.append(torch.cat([lo_b, up_b], dim=-1))                                               # [.., Kg, D]
            off += n
        We = torch.cat(rows_W, dim=-3)
        be = torch.cat(rows_b, dim=-2)
        out[-1] = (We.reshape(lead_w + (We.shape[-3] * D, H)), be.reshape(lead_b + (be.shape[-2] * D,)))
        return out

    def grads_from_made(self, gW, gb, context_dim):
        return _coupling_grads(self, gW, gb, context_dim)


def _coupling_grads(t, gW, gb, context_dim):
    """Gradients in the engine's conditioner format (one draw: gW / gb [n_lin] of [out, in] / [out]) -> gradients of the
    coupling layer's own parameters, in `t.parameters()` order: hyper-network (W, b) per linear, then the free spline
    parameters.  The exact transpose of `SplineCoupling.as_made` (an index re-arrangement)."""
    D, s, C = t.input_dim, t.split_dim, context_dim
    out = []
    n_lin = len(gW)
    for j in range(n_lin - 1):
        out += [gW[j][:, :C + s] if j == 0 else gW[j], gb[j]]
    H = gW[-1].shape[-1]
    M = sum(t.slot_groups())
    gWl = gW[-1].reshape(M, D, H)
    gbl = gb[-1].reshape(M, D)
    up_W, up_b, low, m0 = [], [], [], 0
    for Kg in t.slot_groups():
        up_W.append(gWl[m0:m0 + Kg, s:, :].transpose(0, 1).reshape((D - s) * Kg, H))      # dimension-major rows, as pyro's reshape
        up_b.append(gbl[m0:m0 + Kg, s:].transpose(0, 1).reshape((D - s) * Kg))
        low.append(gbl[m0:m0 + Kg, :s].transpose(0, 1).contiguous())                      # [split_dim, K_g]
        m0 += Kg
    if n_lin == 1:
        out += [torch.cat(up_W, 0)[:, :C + s], torch.cat(up_b, 0)]
    else:
        out += [torch.cat(up_W, 0), torch.cat(up_b, 0)]
    return out + low


def neural_spline_coupling(theta_dim, condition_dim, hidden_dim, num_layers, count_bins, split_dim, order="quadratic",
                           activation=None, use_batchnorm=False, random_perm=False, dropout_p=None):
    """transforms.py:201-236, built to its evident intent.  Upstream the function cannot run: it reads the undefined names
    `input_dim`, `paramdim` and `condotion_dim`, and its unconditional branch constructs `T.SplineAutoregressive` with coupling
    arguments.  With those read as `theta_dim`, `param_dims`, `condition_dim` and `T.SplineCoupling` it is:
    `num_layers` coupling layers, hyper-network `(Conditional)DenseNN(split_dim[, condition_dim], hidden_dims, param_dims)`,
    optional Permute / BatchNorm behind each layer.  Semantics restated from pyro (unpinned: pyro is absent and no upstream
    output can exist)."""
    n = theta_dim - split_dim
    param_dims = [n * count_bins, n * count_bins, n * (count_bins - 1)]
    if order == "linear":
        param_dims.append(n * count_bins)
    elif order != "quadratic":
        raise ValueError(order)
    transforms, nets = [], []
    for _ in range(num_layers):
        arn = ConditionalDenseNN(split_dim, max(condition_dim, 0), _hidden_list(hidden_dim), param_dims=param_dims,
                                 nonlinearity=activation, dropout_p=dropout_p)
        nets.append(arn)
        transforms.append(SplineCoupling(theta_dim, split_dim, arn, count_bins=count_bins, order=order))
        _extras(transforms, theta_dim, use_batchnorm, random_perm)
    return ComposeTransformModule(transforms), transforms, nets
