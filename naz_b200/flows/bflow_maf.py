"""Functional twin API — drop-in for the pieces of src/naz/flows/bflow_jax_maf.py on the hot path:
`torch_to_jax` (weight export, :26-46), `make_normalizing_flow` -> {"lp", "sampler"} (:196-225), the
draw map theta = theta_MLE (1 + scale u) (:239-240) and `compute_bic` (:474-475).  Arrays are torch
tensors instead of jnp arrays; "lp"/"sampler" accept ONE draw (the reference call) or the whole batched
pytree `[L][n_lin](W[S,..], b[S,..])` that the paper scripts slice in a Python loop (calibrate.py:147-150)."""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from ..engine import FlowEngine, FlowShape


def torch_to_jax(torch_maf):
    masks, mask_skips, permutations, params, param_shapes = [], [], [], [], []
    for flow_layer in torch_maf.flow_dist.transforms:
        arn = flow_layer.nn
        this_params, this_shape = [], []
        for layer in arn.layers:
            this_params.append((layer.weight.detach().clone(), layer.bias.detach().clone()))
            this_shape.append((tuple(layer.weight.shape), tuple(layer.bias.shape)))
        param_shapes.append(this_shape)
        params.append(this_params)
        masks.append([m.detach().clone() for m in arn.masks])
        mask_skips.append(arn.mask_skip.detach().clone())
        permutations.append(arn.permutation.detach().clone())
    return params, param_shapes, masks, mask_skips, permutations


def make_conditional_autoregressive_nn(input_dim, context_dim, hidden_dims, param_dims=[1, 1], **_):
    """bflow_jax_maf.py:107-167 — returns the static description the transform needs."""
    return {"input_dim": input_dim, "context_dim": context_dim, "hidden_dims": list(hidden_dims), "param_dims": list(param_dims)}


def make_masked_affine_autoregressive_transform(nn_fn, input_dim, context=None):
    """bflow_jax_maf.py:169-194 — forward/inverse live in libnazb; the handle is the description."""
    return nn_fn


def make_normalizing_flow(transform, x, masks, mask_skips, perms, bounds=None, context=None, device=None, engine="auto"):
    desc = transform
    D, C, hidden = desc["input_dim"], desc["context_dim"], desc["hidden_dims"]
    L = len(masks)
    shape = FlowShape("maf", D, C, hidden, L)
    dev = torch.device(device or (x.device if isinstance(x, torch.Tensor) and x.is_cuda else "cuda"))
    x_dev = torch.as_tensor(x, dtype=torch.float32).to(dev)
    ctx = None if context is None else torch.as_tensor(context, dtype=torch.float32).to(dev)
    bnd = None if bounds is None else (bounds["low"], bounds["high"])
    if bounds is not None:
        lo, hi = torch.as_tensor(bounds["low"]).to(dev), torch.as_tensor(bounds["high"]).to(dev)
        inside = ((x_dev > lo) & (x_dev < hi)).all(dim=-1)
        x_dev = x_dev[inside]                                  # bflow_jax_maf.py:198
    cache = {}

    def _engine(params) -> FlowEngine:
        W0 = params[0][0][0]
        S = W0.shape[0] if W0.dim() == 3 else 1
        key = (S,) + tuple(t.data_ptr() for layer in params for pair in layer for t in pair)
        if cache.get("key") != key:
            if cache.get("eng") is None or cache["eng"].S != S:
                cache["eng"] = FlowEngine(shape, S, device=dev, engine=engine)
            cache["eng"].pack(params, masks, torch.stack([torch.as_tensor(p) for p in perms]))
            cache["key"] = key
        return cache["eng"]

    def log_prob(params):
        eng = _engine(params)
        lp = eng.inverse(x_dev, ctx, None if bnd is None else {"low": bnd[0], "high": bnd[1]}, want_lp=True)["lp"]
        if bnd is not None:
            # the twin SUBTRACTS the bounding log-Jacobian (bflow_jax_maf.py:199,211-212) where flow.py:79 adds it;
            # reproduce the twin's convention for its own API: lp_twin = lp_torch - 2 log_jac
            u = (x_dev - lo) / (hi - lo)
            log_jac = -(torch.log(u) + torch.log1p(-u)).sum(-1) - torch.log(hi - lo).sum()
            lp = lp - 2.0 * log_jac
        return lp[0] if params[0][0][0].dim() == 2 else lp

    def sample(params, rng_key, size):
        eng = _engine(params)
        gen = rng_key if isinstance(rng_key, torch.Generator) else None
        if gen is None and rng_key is not None:
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(rng_key))
        S = eng.S
        z = torch.randn((S, size, D), device=dev, generator=gen)
        assert ctx is None or ctx.dim() == 1                    # bflow_jax_maf.py:217
        y, ld = eng.forward(z, ctx, None if bnd is None else {"low": bnd[0], "high": bnd[1]}, want_logdet=True)
        # twin quirk kept: second output = base log-prob + forward log-dets (bflow_jax_maf.py:219)
        log_j = -(0.5 * z * z).sum(-1) - 0.5 * D * math.log(2 * math.pi) + ld
        if params[0][0][0].dim() == 2:
            return y[0], log_j[0]
        return y, log_j

    def _engine_standard(best_params, standard_params, scale) -> FlowEngine:
        """Engine packed straight from the reference's posterior format {"standard_params": [S, P], "scale"}:
        the draw map of bflow_jax_maf.py:239-240 runs inside the pack kernels (nazb_pack_draw_map)."""
        u = torch.as_tensor(standard_params, dtype=torch.float32)
        S = u.shape[0]
        sc_key = ("t", scale.data_ptr()) if isinstance(scale, torch.Tensor) and scale.numel() > 1 else float(scale)   # posterior["scale"] is [S]
        key = ("std", S, u.data_ptr(), sc_key) + tuple(t.data_ptr() for layer in best_params for pair in layer for t in pair)
        if cache.get("key") != key:
            if cache.get("eng") is None or cache["eng"].S != S:
                cache["eng"] = FlowEngine(shape, S, device=dev, engine=engine)
            cache["eng"].pack_draw_map(best_params, u, scale, masks, torch.stack([torch.as_tensor(p) for p in perms]))
            cache["key"] = key
        return cache["eng"]

    def log_prob_standard(best_params, standard_params, scale):
        """[S, N] log-densities of all posterior draws given as standard parameters (additive entry point: replaces the
        `for i in range(S): lp(unravel(params[i]))` loops of calibrate.py:144-150 / compute_bic_simpler.py:116-120)."""
        assert bnd is None, "bounded flows: use lp(draw_params(...))"
        return _engine_standard(best_params, standard_params, scale).inverse(x_dev, ctx, None, want_lp=True)["lp"]

    def sample_standard(best_params, standard_params, scale, rng_key, size):
        eng = _engine_standard(best_params, standard_params, scale)
        gen = rng_key if isinstance(rng_key, torch.Generator) else None
        if gen is None and rng_key is not None:
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(rng_key))
        z = torch.randn((eng.S, size, D), device=dev, generator=gen)
        assert ctx is None or ctx.dim() == 1
        return eng.forward(z, ctx)

    gcache = {}

    def value_and_grad(params, mean: bool = False):
        """(sum_n lp(params), d/d params) — what the reference's drivers obtain from jax.value_and_grad of
        `log_prob(theta) = jnp.sum(lp(unravel(theta)))` (or jnp.mean) (bflow_jax_maf.py:233-235; NUTS :321-327, SVI :344-348,
        MLE :277-287).  params: one draw `[L][n_lin](W[out,in], b[out])` or a batch `W[S,out,in]` (one gradient per draw /
        chain).  Returns (value [S] float64 or a 0-d tensor, grads with the structure and shapes of params)."""
        W0 = params[0][0][0]
        single = W0.dim() == 2
        S = 1 if single else W0.shape[0]
        if gcache.get("eng") is None or gcache["eng"].S != S:
            gcache["eng"] = FlowEngine(shape, S, device=dev, engine="simt")
        eng = gcache["eng"]
        eng.pack(params, masks, torch.stack([torch.as_tensor(p) for p in perms]))
        r = eng.inverse_grad(x_dev, ctx, None if bnd is None else {"low": bnd[0], "high": bnd[1]})
        val = r["sum_n"]
        if bnd is not None:   # the twin's sign convention for the bounding log-Jacobian (see log_prob); parameter-free
            u = (x_dev - lo) / (hi - lo)
            log_jac = -(torch.log(u) + torch.log1p(-u)).sum(-1) - torch.log(hi - lo).sum()
            val = val - 2.0 * log_jac.double().sum()
        sc = 1.0 / x_dev.shape[0] if mean else 1.0
        grads = [[((gw[0] if single else gw) * sc, (gb[0] if single else gb) * sc) for gw, gb in zip(lw, lb)]
                 for lw, lb in zip(r["gW"], r["gb"])]
        val = val * sc
        return (val[0] if single else val), grads

    return {"lp": log_prob, "sampler": sample, "lp_standard": log_prob_standard, "sampler_standard": sample_standard,
            "value_and_grad": value_and_grad}


def draw_params(best_params, standard_params: torch.Tensor, scale: float):
    """random_params = flat_params * (1 + scale * standard_params) (bflow_jax_maf.py:239-240), un-ravelled
    in pytree order.  standard_params: [S, P] in [-1, 1]."""
    out, off = [], 0
    S = standard_params.shape[0]
    for layer in best_params:
        lay = []
        for (W, b) in layer:
            nW, nb = W.numel(), b.numel()
            uW = standard_params[:, off:off + nW].reshape((S,) + tuple(W.shape)); off += nW
            ub = standard_params[:, off:off + nb].reshape((S,) + tuple(b.shape)); off += nb
            lay.append((W.unsqueeze(0) * (1.0 + scale * uW), b.unsqueeze(0) * (1.0 + scale * ub)))
        out.append(lay)
    return out


def compute_bic(log_ls, N, complexity):
    return complexity * math.log(N) - 2.0 * float(torch.as_tensor(log_ls).max())
