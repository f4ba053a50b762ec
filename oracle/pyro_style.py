"""Second, independent CPU restatement of the reference's torch path.  TEST INFRASTRUCTURE ONLY.

Where oracle/flow_oracle.py is functional numpy following the in-tree JAX twin, this file rebuilds
the *module* structure the reference actually executes through pyro-ppl (absent here):
``MaskedLinear`` -> ``(Conditional)AutoRegressiveNN`` -> ``(Conditional)AffineAutoregressive`` /
``(Conditional)SplineAutoregressive`` (plus ``Permute``, eval-mode ``BatchNorm``, ``(Conditional)DenseNN`` and
``SplineCoupling`` for the factories' optional layers and the coupling flow) -> the real
``torch.distributions.TransformedDistribution``,
assembled the way naz does it (src/naz/flows/transforms.py:133-198, src/naz/flows/flow.py:26-129).
It keeps the reference's cost structure on purpose — ``W*mask`` re-multiplied on every call,
D full conditioner passes per layer in ``_inverse``, a Python loop over draws with an in-place
``param.copy_`` per draw (src/naz/trainers/train_flows.py:47-71,414-420) — so it doubles as the
"reference CPU path" timed by ``bench.py`` (``cpu_baseline.kind == "port"``).

PARITY UNPINNED (see oracle/flow_oracle.py header): pyro is not importable in this image.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Normal, TransformedDistribution, Transform, constraints


def hidden_degrees(n_degrees, width):
    """Degrees of `width` hidden units spread evenly over 1..n_degrees (pyro `sample_mask_indices`, simple=True): an fp32
    linspace rounded half-to-even, as torch / jax round."""
    return torch.round(torch.linspace(1.0, float(n_degrees), steps=width))


def create_mask(input_dim, context_dim, hidden_dims, permutation, output_dim_multiplier):
    """MADE masks from the DEFINITION (Germain et al. 2015, as pyro arranges it), written independently of
    naz_b200/flows/made.py: every unit carries a degree; a connection in -> out exists iff degree(out) >= degree(in) for
    hidden targets and degree(out) > degree(in) for output targets.  Context inputs have degree 0, x_d has degree
    1 + rank(d) under the layer's permutation, hidden units of a CONDITIONAL net take degrees 0 .. D-1 (so that degree-0
    units see the context only), of an unconditional one 1 .. D-1."""
    D = int(input_dim)
    rank = torch.zeros(D)
    for position, dim in enumerate(permutation.tolist()):
        rank[dim] = position
    deg_in = torch.cat([torch.zeros(int(context_dim)), rank + 1.0])
    if context_dim > 0:
        deg_hidden = [hidden_degrees(D, h) - 1.0 for h in hidden_dims]
    else:
        deg_hidden = [hidden_degrees(D - 1, h) for h in hidden_dims]
    deg_out = (rank + 1.0).repeat(output_dim_multiplier)

    def connect(deg_to, deg_from, strict):
        to, frm = deg_to[:, None], deg_from[None, :]
        return (to > frm).float() if strict else (to >= frm).float()

    masks = [connect(deg_hidden[0], deg_in, strict=False)]
    for prev, cur in zip(deg_hidden[:-1], deg_hidden[1:]):
        masks.append(connect(cur, prev, strict=False))
    masks.append(connect(deg_out, deg_hidden[-1], strict=True))
    return masks, connect(deg_out, deg_in, strict=True)


class MaskedLinear(nn.Linear):
    def __init__(self, in_features, out_features, mask, bias=True):
        super().__init__(in_features, out_features, bias)
        self.register_buffer("mask", mask.data)

    def forward(self, x):
        return F.linear(x, self.weight * self.mask, self.bias)     # re-masked on every call, as upstream


class ConditionalAutoRegressiveNN(nn.Module):
    """pyro.nn.ConditionalAutoRegressiveNN with naz's dropout placement (transforms.py:38-43):
    ``h = drop(f(layer(h)))`` on every hidden layer, none on the output layer.  ``keep`` (optional,
    [n_hidden, H] 0/1) replaces nn.Dropout's per-element Philox mask by an explicit per-unit mask."""

    def __init__(self, input_dim, context_dim, hidden_dims, param_dims=(1, 1), permutation=None,
                 nonlinearity=nn.Tanh(), dropout_p=None):
        super().__init__()
        self.input_dim, self.context_dim, self.hidden_dims = input_dim, context_dim, list(hidden_dims)
        self.param_dims = list(param_dims)
        self.count_params = len(param_dims)
        self.output_multiplier = sum(param_dims)
        self.all_ones = all(p == 1 for p in param_dims)
        ends = torch.cumsum(torch.tensor(self.param_dims), dim=0)
        starts = torch.cat((torch.zeros(1).type_as(ends), ends[:-1]))
        self.param_slices = [slice(s.item(), e.item()) for s, e in zip(starts, ends)]
        if permutation is None:
            permutation = torch.randperm(input_dim)
        self.register_buffer("permutation", permutation.to(torch.int64))
        self.masks, self.mask_skip = create_mask(input_dim, context_dim, hidden_dims, self.permutation,
                                                 self.output_multiplier)
        layers = [MaskedLinear(input_dim + context_dim, hidden_dims[0], self.masks[0])]
        for i in range(1, len(hidden_dims)):
            layers.append(MaskedLinear(hidden_dims[i - 1], hidden_dims[i], self.masks[i]))
        layers.append(MaskedLinear(hidden_dims[-1], input_dim * self.output_multiplier, self.masks[-1]))
        self.layers = nn.ModuleList(layers)
        self.f = nonlinearity
        self.dropout_p = dropout_p
        self.keep = None

    def forward(self, x, context=None):
        if self.context_dim > 0:
            context = context.expand(x.size()[:-1] + (context.size(-1),))
            x = torch.cat([context, x], dim=-1)
        return self._forward(x)

    def _forward(self, x):
        h = x
        for i, layer in enumerate(self.layers[:-1]):
            h = self.f(layer(h))
            if self.dropout_p:
                if self.keep is not None:
                    h = h * (self.keep[i] / (1.0 - self.dropout_p))
                else:
                    h = F.dropout(h, self.dropout_p, self.training)
        h = self.layers[-1](h)
        if self.output_multiplier == 1:
            return h
        h = h.reshape(list(x.size()[:-1]) + [self.output_multiplier, self.input_dim])
        if self.count_params == 1:
            return h
        if self.all_ones:
            return torch.unbind(h, dim=-2)
        return tuple(h[..., s, :] for s in self.param_slices)


class _ARTransform(Transform):
    domain = constraints.real_vector
    codomain = constraints.real_vector
    bijective = True
    sign = +1

    def __init__(self, nn_fn, arn):
        super().__init__(cache_size=1)
        self.nn_fn, self.arn = nn_fn, arn
        self._cached_ld = None

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class AffineAutoregressive(_ARTransform):
    """pyro AffineAutoregressive (stable=False, clip -5/3)."""

    def __init__(self, nn_fn, arn, lo=-5.0, hi=3.0):
        super().__init__(nn_fn, arn)
        self.lo, self.hi = lo, hi

    def _call(self, x):
        mean, log_scale = self.nn_fn(x)
        log_scale = log_scale.clamp(self.lo, self.hi)
        self._cached_ld = log_scale
        return torch.exp(log_scale) * x + mean

    def _inverse(self, y):
        input_dim = y.size(-1)
        x = [torch.zeros(y.size()[:-1])] * input_dim
        for idx in self.arn.permutation:
            mean, log_scale = self.nn_fn(torch.stack(x, dim=-1))
            inverse_scale = torch.exp(-log_scale[..., idx].clamp(self.lo, self.hi))
            x[idx] = (y[..., idx] - mean[..., idx]) * inverse_scale
        self._cached_ld = log_scale.clamp(self.lo, self.hi)
        return torch.stack(x, dim=-1)

    def log_abs_det_jacobian(self, x, y):
        return self._cached_ld.sum(-1)


def _calculate_knots(lengths, lower, upper):
    knots = torch.cumsum(lengths, dim=-1)
    knots = F.pad(knots, pad=(1, 0), mode="constant", value=0.0)
    knots = (upper - lower) * knots + lower
    knots[..., 0] = lower
    knots[..., -1] = upper
    return knots[..., 1:] - knots[..., :-1], knots


def _select_bins(x, idx):
    idx = idx.clamp(min=0, max=x.size(-1) - 1)
    return x.gather(-1, idx).squeeze(-1)


def monotonic_rational_spline(inputs, widths, heights, derivatives, lambdas=None, inverse=False, bound=3.0,
                              min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3, min_lambda=0.025,
                              eps=1e-6):
    left, right, bottom, top = -bound, bound, -bound, bound
    inside = (inputs >= left) & (inputs <= right)
    outside = ~inside
    num_bins = widths.size(-1)
    widths = min_bin_width + (1.0 - min_bin_width * num_bins) * widths
    heights = min_bin_height + (1.0 - min_bin_height * num_bins) * heights
    derivatives = min_derivative + derivatives
    widths, cumwidths = _calculate_knots(widths, left, right)
    heights, cumheights = _calculate_knots(heights, bottom, top)
    derivatives = F.pad(derivatives, pad=(1, 1), mode="constant", value=1.0 - min_derivative)
    seq = cumheights + eps if inverse else cumwidths + eps
    bin_idx = (torch.sum(inputs[..., None] >= seq, dim=-1) - 1).unsqueeze(-1)
    input_widths = _select_bins(widths, bin_idx)
    input_cumwidths = _select_bins(cumwidths, bin_idx)
    input_cumheights = _select_bins(cumheights, bin_idx)
    input_delta = _select_bins(heights / widths, bin_idx)
    input_derivatives = _select_bins(derivatives, bin_idx)
    input_derivatives_plus_one = _select_bins(derivatives[..., 1:], bin_idx)
    input_heights = _select_bins(heights, bin_idx)
    if lambdas is not None:
        lambdas = (1 - 2 * min_lambda) * lambdas + min_lambda
        lam = _select_bins(lambdas, bin_idx)
        wa = 1.0
        wb = torch.sqrt(input_derivatives / input_derivatives_plus_one) * wa
        wc = (lam * wa * input_derivatives + (1 - lam) * wb * input_derivatives_plus_one) / input_delta
        ya = input_cumheights
        yb = input_heights + input_cumheights
        yc = ((1.0 - lam) * wa * ya + lam * wb * yb) / ((1.0 - lam) * wa + lam * wb)
        if inverse:
            le, gt = (inputs <= yc).float(), (inputs > yc).float()
            numerator = (lam * wa * (ya - inputs)) * le + ((wc - lam * wb) * inputs + lam * wb * yb - wc * yc) * gt
            denominator = ((wc - wa) * inputs + wa * ya - wc * yc) * le + ((wc - wb) * inputs + wb * yb - wc * yc) * gt
            theta = numerator / denominator
            outputs = theta * input_widths + input_cumwidths
            dnum = (wa * wc * lam * (yc - ya) * le + wb * wc * (1 - lam) * (yb - yc) * gt) * input_widths
            logabsdet = torch.log(dnum) - 2 * torch.log(torch.abs(denominator))
        else:
            theta = (inputs - input_cumwidths) / input_widths
            le, gt = (theta <= lam).float(), (theta > lam).float()
            numerator = (wa * ya * (lam - theta) + wc * yc * theta) * le + (wc * yc * (1 - theta) + wb * yb * (theta - lam)) * gt
            denominator = (wa * (lam - theta) + wc * theta) * le + (wc * (1 - theta) + wb * (theta - lam)) * gt
            outputs = numerator / denominator
            dnum = (wa * wc * lam * (yc - ya) * le + wb * wc * (1 - lam) * (yb - yc) * gt) / input_widths
            logabsdet = torch.log(dnum) - 2 * torch.log(torch.abs(denominator))
    else:
        if inverse:
            a = (inputs - input_cumheights) * (input_derivatives + input_derivatives_plus_one - 2 * input_delta) \
                + input_heights * (input_delta - input_derivatives)
            b = input_heights * input_derivatives - (inputs - input_cumheights) * (
                input_derivatives + input_derivatives_plus_one - 2 * input_delta)
            c = -input_delta * (inputs - input_cumheights)
            discriminant = b.pow(2) - 4 * a * c
            discriminant = discriminant.masked_fill(outside, 0)
            root = (2 * c) / (-b - torch.sqrt(discriminant))
            outputs = root * input_widths + input_cumwidths
            tomt = root * (1 - root)
            denominator = input_delta + (input_derivatives + input_derivatives_plus_one - 2 * input_delta) * tomt
            dnum = input_delta.pow(2) * (input_derivatives_plus_one * root.pow(2) + 2 * input_delta * tomt
                                         + input_derivatives * (1 - root).pow(2))
            logabsdet = -(torch.log(dnum) - 2 * torch.log(denominator))
        else:
            theta = (inputs - input_cumwidths) / input_widths
            tomt = theta * (1 - theta)
            numerator = input_heights * (input_delta * theta.pow(2) + input_derivatives * tomt)
            denominator = input_delta + (input_derivatives + input_derivatives_plus_one - 2 * input_delta) * tomt
            outputs = input_cumheights + numerator / denominator
            dnum = input_delta.pow(2) * (input_derivatives_plus_one * theta.pow(2) + 2 * input_delta * tomt
                                         + input_derivatives * (1 - theta).pow(2))
            logabsdet = torch.log(dnum) - 2 * torch.log(denominator)
    outputs = torch.where(outside, inputs, outputs)
    logabsdet = torch.where(outside, torch.zeros_like(logabsdet), logabsdet)
    return outputs, logabsdet


class SplineAutoregressive(_ARTransform):
    """pyro SplineAutoregressive: one conditioner pass in _call, D Jacobi sweeps in _inverse."""

    def __init__(self, nn_fn, arn, count_bins=8, bound=3.0, order="quadratic"):
        super().__init__(nn_fn, arn)
        self.count_bins, self.bound, self.order = count_bins, bound, order

    def _params(self, x):
        if self.order == "linear":
            w, h, d, l = self.nn_fn(x)
            l = torch.sigmoid(l.transpose(-1, -2))
        else:
            w, h, d = self.nn_fn(x)
            l = None
        w = F.softmax(w.transpose(-1, -2), dim=-1)
        h = F.softmax(h.transpose(-1, -2), dim=-1)
        d = F.softplus(d.transpose(-1, -2))
        return w, h, d, l

    def _call(self, x):
        w, h, d, l = self._params(x)
        y, ld = monotonic_rational_spline(x, w, h, d, l, bound=self.bound)
        self._cached_ld = ld
        return y

    def _inverse(self, y):
        x = torch.zeros_like(y)
        for _ in range(y.size(-1)):
            w, h, d, l = self._params(x)
            x, ld = monotonic_rational_spline(y, w, h, d, l, bound=self.bound, inverse=True)
        self._cached_ld = -ld
        return x

    def log_abs_det_jacobian(self, x, y):
        return self._cached_ld.sum(-1)


class ConditionalDenseNN(nn.Module):
    """pyro.nn.ConditionalDenseNN / DenseNN (restated): plain Linear layers, context concatenated IN FRONT of the input,
    nonlinearity on every hidden layer, output split by `param_dims` (each part flat, e.g. (input_dim - split_dim) * count_bins).
    naz builds it as the hyper-network of its coupling flow (src/naz/flows/transforms.py:219-224)."""

    def __init__(self, input_dim, context_dim, hidden_dims, param_dims, nonlinearity=nn.Tanh()):
        super().__init__()
        self.input_dim, self.context_dim, self.hidden_dims, self.param_dims = input_dim, context_dim, list(hidden_dims), list(param_dims)
        dims = [input_dim + context_dim] + list(hidden_dims) + [sum(param_dims)]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
        self.f = nonlinearity

    def forward(self, x, context=None):
        if self.context_dim > 0:
            context = context.expand(x.size()[:-1] + (context.size(-1),))
            x = torch.cat([context, x], dim=-1)
        h = x
        for layer in self.layers[:-1]:
            h = self.f(layer(h))
        h = self.layers[-1](h)
        out, off = [], 0
        for pd in self.param_dims:
            out.append(h[..., off:off + pd])
            off += pd
        return tuple(out)


class SplineCoupling(Transform):
    """pyro `T.SplineCoupling` (pyro/distributions/transforms/spline_coupling.py, restated; with `ConditionalSplineCoupling`
    of transforms.py:113-129 the hyper-network also sees the context): the first `split_dim` coordinates go through an
    element-wise spline with free parameters (`lower`: unnormalized widths / heights [split_dim, K], derivatives
    [split_dim, K-1], lambdas [split_dim, K] for the linear order), the remaining ones through a spline whose parameters the
    hyper-network computes from the first part, reshaped [..., input_dim - split_dim, K]; widths / heights by softmax,
    derivatives by softplus, lambdas by sigmoid, as in SplineAutoregressive.  PARITY UNPINNED (pyro absent), and the
    reference's own factory for it cannot be constructed (undefined names)."""
    domain = constraints.real_vector
    codomain = constraints.real_vector
    bijective = True

    def __init__(self, input_dim, split_dim, hypernet_fn, lower, count_bins=8, bound=3.0, order="quadratic"):
        super().__init__(cache_size=1)
        self.input_dim, self.split_dim, self.nn_fn, self.lower = input_dim, split_dim, hypernet_fn, lower
        self.count_bins, self.bound, self.order = count_bins, bound, order
        self._cached_ld = None

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    @staticmethod
    def _normalise(w, h, d, l):
        return F.softmax(w, dim=-1), F.softmax(h, dim=-1), F.softplus(d), (None if l is None else torch.sigmoid(l))

    def _upper_params(self, x1):
        out = self.nn_fn(x1)
        n, K = self.input_dim - self.split_dim, self.count_bins
        w = out[0].reshape(out[0].shape[:-1] + (n, K))
        h = out[1].reshape(out[1].shape[:-1] + (n, K))
        d = out[2].reshape(out[2].shape[:-1] + (n, K - 1))
        l = out[3].reshape(out[3].shape[:-1] + (n, K)) if self.order == "linear" else None
        return self._normalise(w, h, d, l)

    def _lower_params(self):
        return self._normalise(self.lower[0], self.lower[1], self.lower[2], self.lower[3] if self.order == "linear" else None)

    def _both(self, v, inverse):
        v1, v2 = v[..., :self.split_dim], v[..., self.split_dim:]
        w, h, d, l = [None if t is None else t.expand(v1.shape + t.shape[-1:]) for t in self._lower_params()]
        o1, ld1 = monotonic_rational_spline(v1, w, h, d, l, bound=self.bound, inverse=inverse)
        w, h, d, l = self._upper_params(o1 if inverse else v1)     # the hyper-network always sees the INPUT-side first part
        o2, ld2 = monotonic_rational_spline(v2, w, h, d, l, bound=self.bound, inverse=inverse)
        return torch.cat([o1, o2], dim=-1), torch.cat([ld1, ld2], dim=-1)

    def _call(self, x):
        y, ld = self._both(x, False)
        self._cached_ld = ld
        return y

    def _inverse(self, y):
        x, ld = self._both(y, True)
        self._cached_ld = -ld
        return x

    def log_abs_det_jacobian(self, x, y):
        return self._cached_ld.sum(-1)


class Permute(Transform):
    """pyro `T.Permute` (pyro/distributions/transforms/permute.py, restated): y = x.index_select(-1, permutation), inverse
    by the inverse permutation, log|det J| = 0.  naz appends one per flow layer with random_perm=True (transforms.py:155-156)."""
    domain = constraints.real_vector
    codomain = constraints.real_vector
    bijective = True

    def __init__(self, permutation):
        super().__init__(cache_size=0)
        self.permutation = torch.as_tensor(permutation, dtype=torch.int64)
        self.inv_permutation = torch.argsort(self.permutation)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def _call(self, x):
        return x.index_select(-1, self.permutation)

    def _inverse(self, y):
        return y.index_select(-1, self.inv_permutation)

    def log_abs_det_jacobian(self, x, y):
        return torch.zeros(x.shape[:-1], dtype=x.dtype)


class BatchNormEval(Transform):
    """pyro `T.BatchNorm` (pyro/distributions/transforms/batchnorm.py, restated) in eval() mode — the training-mode batch
    statistics are the training loop's business:  _call(x) = (x - beta) / gamma_c * sqrt(moving_variance + eps) + moving_mean,
    _inverse(y) = (y - moving_mean) * gamma_c / sqrt(moving_variance + eps) + beta, gamma_c = relu(gamma) + 1e-6,
    log|dy/dx| = -log gamma_c + 0.5 log(moving_variance + eps) per dimension (summed here over the event dimension, as the
    TransformedDistribution does for pyro's element-wise version).  naz appends one per flow layer with use_batchnorm=True
    (transforms.py:157-158)."""
    domain = constraints.real_vector
    codomain = constraints.real_vector
    bijective = True

    def __init__(self, gamma, beta, moving_mean, moving_variance, epsilon=1e-5):
        super().__init__(cache_size=0)
        self.gamma, self.beta = torch.as_tensor(gamma), torch.as_tensor(beta)
        self.moving_mean, self.moving_variance, self.epsilon = torch.as_tensor(moving_mean), torch.as_tensor(moving_variance), epsilon

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    @property
    def constrained_gamma(self):
        return F.relu(self.gamma) + 1e-6

    def _call(self, x):
        return (x - self.beta) / self.constrained_gamma * torch.sqrt(self.moving_variance + self.epsilon) + self.moving_mean

    def _inverse(self, y):
        return (y - self.moving_mean) * self.constrained_gamma / torch.sqrt(self.moving_variance + self.epsilon) + self.beta

    def log_abs_det_jacobian(self, x, y):
        ld = -self.constrained_gamma.log() + 0.5 * torch.log(self.moving_variance + self.epsilon)
        return ld.expand(x.shape).sum(-1)


def bounding_transform(x, low, high):
    """src/naz/flows/transforms.py:20-23."""
    y = (x - low.expand(x.shape)) / ((high - low).expand(x.shape))
    log_jac = -torch.sum(torch.log(y) + torch.log1p(-y), axis=-1) - torch.sum(torch.log(high - low))
    return torch.logit(y), log_jac


def inverse_bounding_transform(y, low, high):
    """src/naz/flows/transforms.py:25-27."""
    return torch.sigmoid(y) * ((high - low).expand(y.shape)) + low.expand(y.shape)


class PyroStyleFlow(nn.Module):
    """naz ``NormalizingFlow`` (flow.py:24-129) for flow_type in {"maf", "nsa"} on the torch stack."""

    def __init__(self, flow_type, bounds, theta_dim, condition_dim, hidden_dim, num_layers, count_bins=8,
                 order="quadratic", permutations=None, dropout_p=None, extras=None):
        """extras: optional [L] of lists of ready-made transforms (Permute / BatchNormEval) placed behind flow layer l, in the
        order naz appends them (transforms.py:155-158: Permute, then BatchNorm)."""
        super().__init__()
        self.extras = extras
        assert flow_type in ("maf", "nsa")
        self.flow_type, self.bounds = flow_type, bounds
        self.theta_dim, self.condition_dim = theta_dim, condition_dim
        self.conditional = condition_dim > 0
        self.count_bins, self.order = count_bins, order
        hidden_dims = hidden_dim if isinstance(hidden_dim, list) else [hidden_dim]
        if flow_type == "maf":
            param_dims = [1, 1]
        else:
            param_dims = [count_bins, count_bins, count_bins - 1] + ([count_bins] if order == "linear" else [])
        self.nets = nn.ModuleList()
        for l in range(num_layers):
            perm = None if permutations is None else torch.as_tensor(permutations[l])
            self.nets.append(ConditionalAutoRegressiveNN(theta_dim, condition_dim, hidden_dims, param_dims, perm,
                                                         nn.Tanh(), dropout_p))
        self.base_dist = Normal(torch.zeros(theta_dim), torch.ones(theta_dim))

    def _transforms(self, condition):
        ts, ts_ar = [], []
        for arn in self.nets:
            fn = partial(arn, context=condition) if self.conditional else arn
            if self.flow_type == "maf":
                ts.append(AffineAutoregressive(fn, arn))
            else:
                ts.append(SplineAutoregressive(fn, arn, self.count_bins, 3.0, self.order))
            if self.extras is not None:
                ts.extend(self.extras[len(ts_ar)])
            ts_ar.append(arn)
        return ts

    def _pdf(self, condition):
        if self.conditional:
            assert condition is not None
        return TransformedDistribution(self.base_dist, self._transforms(condition))

    def log_prob(self, x, condition=None):
        if self.bounds is None:
            y, log_jac = x, 0.0
        else:
            y, log_jac = bounding_transform(x, self.bounds["low"], self.bounds["high"])
        return self._pdf(condition).log_prob(y) + log_jac

    def sample(self, shape, condition=None, base_noise=None):
        pdf = self._pdf(condition)
        with torch.no_grad():
            x = pdf.base_dist.sample(shape) if base_noise is None else base_noise
            for t in pdf.transforms:
                x = t(x)
        return x if self.bounds is None else inverse_bounding_transform(x, self.bounds["low"], self.bounds["high"])

    # --- reference-format weight exchange (train_flows.py:20-71) ---
    def set_from_pytree(self, params):
        """params: [L][n_lin](W[out,in], b[out]) numpy / tensors — one draw."""
        with torch.no_grad():
            for arn, layer in zip(self.nets, params):
                for lin, (W, b) in zip(arn.layers, layer):
                    lin.weight.copy_(torch.as_tensor(W))
                    lin.bias.copy_(torch.as_tensor(b))

    def set_keep(self, keep):
        """keep: [L, n_hidden, H] 0/1 (one MC-dropout draw) or None."""
        for l, arn in enumerate(self.nets):
            arn.keep = None if keep is None else torch.as_tensor(keep[l], dtype=torch.float32)


def log_prob_draws_reference_loop(flow: PyroStyleFlow, draws, x, condition=None, keep=None):
    """The loop the CUDA path replaces: per draw, copy weights in place, evaluate, collect
    (train_flows.py:414-420; compute_bic_simpler.py:116-120)."""
    S = draws[0][0][0].shape[0]
    out = []
    with torch.no_grad():
        for s in range(S):
            flow.set_from_pytree([[(W[s], b[s]) for (W, b) in layer] for layer in draws])
            if keep is not None:
                flow.set_keep(keep[s])
            out.append(flow.log_prob(x, condition=condition))
    return torch.stack(out)


def sample_draws_reference_loop(flow: PyroStyleFlow, draws, z, condition=None, keep=None):
    S = draws[0][0][0].shape[0]
    out = []
    with torch.no_grad():
        for s in range(S):
            flow.set_from_pytree([[(W[s], b[s]) for (W, b) in layer] for layer in draws])
            if keep is not None:
                flow.set_keep(keep[s])
            zz = z[s] if z.dim() == 3 else z
            out.append(flow.sample(None, condition=condition, base_noise=zz))
    return torch.stack(out)
