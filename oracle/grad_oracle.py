"""Gradient oracle for SURVEY §8 row f1.  TEST INFRASTRUCTURE ONLY (same rules as flow_oracle.py: imported by tests/,
smoke() and bench.py's CPU legs, never by naz_b200/).

The reference has no hand-written backward pass: its NUTS / SVI / MLE drivers differentiate the scalar
``log_prob(theta) = sum_n lp_n(theta)`` with jax.grad (src/naz/flows/bflow_jax_maf.py:233-246, :277-287, :321-327,
:344-348) or torch autograd (src/naz/trainers/train_flows.py:195-213).  The oracle therefore is *automatic differentiation
of the restated forward function*: the masked-affine twin below follows bflow_jax_maf.py line by line
(masked_linear :74-77, nn_fn :135-165 with the context first and the [.., M, D] reshape, inverse_fn :181-194 with its D
sequential passes and clip(-5, 3), log_prob assembly :210-212) in torch float64, and torch.autograd supplies the
gradients.  PINNED BY DERIVATIVES OF THE REFERENCE'S OWN CODE: tools/make_reference_goldens.py::grad_fixture executes the
reference's bflow_jax_maf.py (make_conditional_autoregressive_nn / inverse_fn / make_normalizing_flow log_prob, its bytes)
with torch standing in for jax.numpy and lets torch.autograd play jax.grad; the resulting d(sum lp)/d(W, b) and d lp/dx are
committed as tests/golden/ref_twin_grad_*.npz, and tests/test_grad.py requires this oracle to reproduce them to 1e-8 (and
its forward values to equal the reference-executed ref_twin_*.npz and oracle/flow_oracle.py); central finite differences
of the pinned forward are checked as well.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _nn(xin, ctx, layer, masks_l, D):
    h = xin if ctx is None else torch.cat([ctx, xin], -1)          # context first (bflow_jax_maf.py:141-142)
    for j, (W, b) in enumerate(layer):
        h = h @ (W * masks_l[j]).T + b                             # masked_linear (:74-77)
        if j < len(layer) - 1:
            h = torch.tanh(h)
    o = h.reshape(h.shape[:-1] + (2, D))                           # [.., M, D] (:161-163)
    return o[..., 0, :], o[..., 1, :]


def _inverse_layer(y, ctx, layer, masks_l, perm, D, clip):
    x = torch.zeros_like(y)
    for idx in perm:                                               # D sequential passes (:185-190)
        mu, s = _nn(x, ctx, layer, masks_l, D)
        s = s.clamp(clip[0], clip[1])
        x = x.clone()
        x[..., idx] = (y[..., idx] - mu[..., idx]) * torch.exp(-s[..., idx])
    mu, s = _nn(x, ctx, layer, masks_l, D)
    return x, s.clamp(clip[0], clip[1]).sum(-1)                    # log-det from the last pass (:191-193)


def log_prob(params, masks, perms, x, ctx=None, bounds=None, clip=(-5.0, 3.0)):
    """lp [N] of one draw.  params: [L][n_lin] of (W, b) torch float64 (may require grad); bounds follow the torch
    path's sign convention (flow.py:66-79: + log|d logit / dx|), which is what nazb_inverse computes."""
    D = x.shape[-1]
    y = x
    lj = 0.0
    if bounds is not None:
        lo, hi = bounds
        u = (x - lo) / (hi - lo)
        y = torch.log(u) - torch.log1p(-u)
        lj = -(torch.log(u) + torch.log1p(-u)).sum(-1) - torch.log(hi - lo).sum()
    ld = 0.0
    for l in reversed(range(len(params))):                         # reduce(inverse_fn, reversed(layers)) (:210-211)
        y, ldl = _inverse_layer(y, ctx, params[l], masks[l], [int(i) for i in perms[l]], D, clip)
        ld = ld + ldl
    return -0.5 * (y * y).sum(-1) - 0.5 * D * math.log(2.0 * math.pi) - ld + lj


def value_and_grad(params_np, masks_np, perms, x_np, ctx_np=None, bounds_np=None, want_dx=False):
    """-> (sum_n lp, gW [L][n_lin], gb [L][n_lin], dx [N, D] or None), float64 numpy, for ONE draw."""
    P = [[(torch.tensor(np.asarray(W, np.float64), requires_grad=True), torch.tensor(np.asarray(b, np.float64), requires_grad=True))
          for (W, b) in layer] for layer in params_np]
    M = [[torch.tensor(np.asarray(m, np.float64)) for m in ml] for ml in masks_np]
    x = torch.tensor(np.asarray(x_np, np.float64), requires_grad=want_dx)
    ctx = None
    if ctx_np is not None:
        ctx = torch.tensor(np.asarray(ctx_np, np.float64))
        if ctx.dim() == 1:
            ctx = ctx.unsqueeze(0).expand(x.shape[0], -1)
    bounds = None if bounds_np is None else tuple(torch.tensor(np.asarray(b, np.float64)) for b in bounds_np)
    lp = log_prob(P, M, perms, x, ctx, bounds)
    tot = lp.sum()
    tot.backward()
    gW = [[W.grad.numpy() for (W, _) in layer] for layer in P]
    gb = [[b.grad.numpy() for (_, b) in layer] for layer in P]
    return float(tot.detach()), gW, gb, (x.grad.numpy() if want_dx else None), lp.detach().numpy()


def value_and_grad_flow(spec, params_np, x_np, ctx_np=None, bounds_np=None, want_dx=False):
    """Gradient oracle for ANY flow kind the forward oracle covers (maf, nsa quadratic / linear): torch autograd (float64)
    through the second restatement oracle/pyro_style.py (torch modules on torch.distributions.TransformedDistribution), i.e.
    what the reference's torch path does in `train` (train_flows.py:195-213: loss = -flow.log_prob(...).mean(); backward()).
    The spline branch of that restatement is UNPINNED (pyro-ppl is absent); this function is groundwork for the spline
    backward kernel, checked in tests/ against finite differences and, for maf, against the pinned twin above.
    -> (sum_n lp, gW [L][n_lin], gb [L][n_lin]) float64 numpy, ONE draw; with want_dx also d lp_n / d x_n [N, D]."""
    from oracle import pyro_style as ps
    torch.set_default_dtype(torch.float64)
    try:
        bounds = None if bounds_np is None else {"low": torch.tensor(np.asarray(bounds_np[0], np.float64)),
                                                 "high": torch.tensor(np.asarray(bounds_np[1], np.float64))}
        flow = ps.PyroStyleFlow(spec.kind, bounds, spec.D, spec.C, list(spec.hidden), spec.L, spec.count_bins, spec.order,
                                permutations=spec.perms)
        flow.set_from_pytree([[(np.asarray(W, np.float64), np.asarray(b, np.float64)) for (W, b) in layer] for layer in params_np])
        x = torch.tensor(np.asarray(x_np, np.float64), requires_grad=bool(want_dx))
        ctx = None
        if ctx_np is not None:
            ctx = torch.tensor(np.asarray(ctx_np, np.float64))
            if ctx.dim() == 1:
                ctx = ctx.unsqueeze(0).expand(x.shape[0], -1)
        tot = flow.log_prob(x, ctx).sum()
        tot.backward()
        gW = [[lin.weight.grad.numpy().copy() for lin in arn.layers] for arn in flow.nets]
        gb = [[lin.bias.grad.numpy().copy() for lin in arn.layers] for arn in flow.nets]
        if want_dx:
            return float(tot.detach()), gW, gb, x.grad.numpy().copy()
        return float(tot.detach()), gW, gb
    finally:
        torch.set_default_dtype(torch.float32)
