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


def sample_mask_indices(input_dim: int, hidden_dim: int, simple: bool = True):
    """bflow_jax_maf.py:48-50 (simple masking only: the random variant draws from numpy's global RNG upstream)."""
    from .made import sample_mask_indices as _smi
    if not simple:
        raise NotImplementedError("simple_masking=False draws Bernoulli degrees from np.random upstream; not on the hot path")
    return _smi(input_dim, hidden_dim)


def create_mask(input_dim, context_dim, hidden_dims, permutation, output_dim_multiplier):
    """bflow_jax_maf.py:52-72 -> (masks, mask_skip)."""
    from .made import create_mask as _cm
    return _cm(input_dim, context_dim, list(hidden_dims), torch.as_tensor(permutation).long(), output_dim_multiplier)


def tanh(args):
    """bflow_jax_maf.py:94-96: the default `activation_fn` (and the only one the kernels implement)."""
    return torch.tanh(args)


class _ConditionerFn:
    """First element of what `make_conditional_autoregressive_nn` returns (the reference's jitted `nn_fn`,
    bflow_jax_maf.py:135-165).  In libnazb the masked MLP is fused with the transform that consumes it (tensor-core pushes +
    transform epilogue), so this object is the conditioner's DESCRIPTION; the transforms built from it evaluate it on the
    GPU.  Calling it on its own would need the raw (mean, log_scale) pair, which the fused kernels never materialise."""

    def __init__(self, input_dim, context_dim, hidden_dims, param_dims):
        self.input_dim, self.context_dim = int(input_dim), int(context_dim)
        self.hidden_dims, self.param_dims = [int(h) for h in hidden_dims], [int(p) for p in param_dims]

    def __call__(self, x, params, masks, mask_skip, context=None):
        raise RuntimeError("the conditioner is fused into the libnazb transform kernels: evaluate it through "
                           "make_masked_affine_autoregressive_transform(...)[0 / 1] or make_normalizing_flow(...)")

    # legacy accessors (round-1 callers indexed the description like a dict)
    def __getitem__(self, k):
        return getattr(self, k)


def make_conditional_autoregressive_nn(input_dim: int, context_dim: int, hidden_dims: List[int], param_dims: List[int] = [1, 1],
                                       permutation=None, skip_connections: bool = False, activation_fn=tanh,
                                       simple_masking: bool = True):
    """bflow_jax_maf.py:107-167 -> (nn_fn, param_shapes, generate_mask), unpacked by every reference call site
    (calibrate.py:91, hmc_maf_exact.py:105, plot_svi.py:83)."""
    if skip_connections:
        raise NotImplementedError("skip_connections=True: the libnazb conditioner has no skip path (upstream default: False)")
    if activation_fn is not tanh and activation_fn is not torch.tanh:
        raise NotImplementedError("the libnazb conditioner implements tanh (upstream default) only")
    if not simple_masking:
        raise NotImplementedError("simple_masking=False (random degrees from np.random) is not supported")
    hidden_dims = [int(h) for h in hidden_dims]
    output_multiplier = sum(param_dims)
    default_perm = permutation

    def generate_mask(permutation=default_perm):
        if permutation is None:
            permutation = torch.randperm(input_dim)     # upstream: np.random.permutation(input_dim)
        permutation = torch.as_tensor(permutation).long()
        masks, mask_skip = create_mask(input_dim, context_dim, hidden_dims, permutation, output_multiplier)
        return masks, mask_skip, permutation

    # same (weight shape, bias shape) entries as upstream, bias "shapes" being bare ints there too (bflow_jax_maf.py:130-133)
    param_shapes = [((hidden_dims[0], input_dim + context_dim), (hidden_dims[0]))]
    for i in range(1, len(hidden_dims)):
        param_shapes.append(((hidden_dims[i], hidden_dims[i - 1]), (hidden_dims[i])))
    param_shapes.append(((input_dim * output_multiplier, hidden_dims[-1]), (input_dim * output_multiplier)))
    return _ConditionerFn(input_dim, context_dim, hidden_dims, param_dims), param_shapes, generate_mask


def _perm_from_mask_skip(mask_skip, C: int) -> torch.Tensor:
    """mask_skip[o, C + i] = (var_index[o] > var_index[i]) (bflow_jax_maf.py:66): the row sum over the x columns of output
    o < D is the rank of dimension o, so the layer's permutation can be read back from the masks alone."""
    ms = torch.as_tensor(mask_skip).float().cpu()
    D = ms.shape[1] - C
    rank = ms[:D, C:].sum(1).round().long()
    perm = torch.empty(D, dtype=torch.long)
    perm[rank] = torch.arange(D)
    return perm


class _LayerFn:
    """forward_fn / inverse_fn of ONE masked-affine autoregressive layer (bflow_jax_maf.py:173-193) on the libnazb path:
    a one-layer FlowEngine, repacked when the parameters change."""

    def __init__(self, desc: _ConditionerFn, context, inverse: bool, engine: str = "auto"):
        self.desc, self.context, self.inverse, self.engine = desc, context, inverse, engine
        self._eng, self._key = None, None

    def _get(self, params, masks, perm, dev):
        key = tuple((t.data_ptr(), getattr(t, "_version", 0)) for pair in params for t in pair) + (tuple(perm.tolist()),)
        if self._key != key:
            d = self.desc
            if self._eng is None:
                self._eng = FlowEngine(FlowShape("maf", d.input_dim, d.context_dim, d.hidden_dims, 1), 1, device=dev, engine=self.engine)
            self._eng.pack([list(params)], [list(masks)], perm.reshape(1, -1))
            self._key = key
        return self._eng

    def __call__(self, xj, args, context=None):
        context = self.context if context is None else context
        x, log_det_j = xj
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError("naz_b200 transforms evaluate on CUDA tensors only (no CPU fallback)")
        D, C = self.desc.input_dim, self.desc.context_dim
        if self.inverse:
            params, masks, mask_skip, perm = args
            perm = torch.as_tensor(perm).long().cpu()
        else:
            params, masks, mask_skip = args
            perm = _perm_from_mask_skip(mask_skip, C)
        eng = self._get(params, masks, perm, x.device)
        ctx = None if context is None else torch.as_tensor(context, dtype=torch.float32, device=x.device)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, D).float()
        if ctx is not None and ctx.dim() > 1:
            ctx = ctx.reshape(-1, C)
        if self.inverse:
            out = eng.inverse(x2, ctx, None, want_z=True, want_lp=True)
            z, lp = out["z"][0], out["lp"][0]
            # lp = -sum z^2/2 - D/2 log 2pi - sum log_scale  =>  the layer's log-det (what upstream adds, :192)
            ld = -(lp.double() + 0.5 * (z.double() ** 2).sum(-1) + 0.5 * D * math.log(2 * math.pi)).float()
            return z.reshape(*lead, D), log_det_j + ld.reshape(lead)
        y, ld = eng.forward(x2.unsqueeze(0), ctx, None, want_logdet=True)
        return y[0].reshape(*lead, D), log_det_j + ld[0].reshape(lead)


def make_masked_affine_autoregressive_transform(nn_fn, input_dim: int, context=None):
    """bflow_jax_maf.py:169-194 -> (forward_fn, inverse_fn).  forward_fn((x, log_det_j), (params, masks, mask_skip)[, context]),
    inverse_fn((y, log_det_j), (params, masks, mask_skip, perm)[, context]) — one flow layer each, usable on their own or
    reduced over the layers as upstream does; `make_normalizing_flow` recognises the pair and runs ALL layers in one launch."""
    if not isinstance(nn_fn, _ConditionerFn):
        raise TypeError("nn_fn must come from make_conditional_autoregressive_nn")
    assert int(input_dim) == nn_fn.input_dim
    return (_LayerFn(nn_fn, context, inverse=False), _LayerFn(nn_fn, context, inverse=True))


def make_normalizing_flow(transform, x, masks, mask_skips, perms, bounds=None, context=None, device=None, engine="auto"):
    """bflow_jax_maf.py:196-225 -> {"lp", "sampler"} (+ the draw-batched entry points documented in INTEGRATION.md)."""
    desc = transform[0].desc if isinstance(transform, (tuple, list)) else transform
    D, C, hidden = desc.input_dim, desc.context_dim, desc.hidden_dims
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
        # (address, in-place version) of every leaf: an optimizer step / W.mul_() / refilled buffer repacks; tensors that
        # are not fp32 CUDA are copied by the engine, so their addresses mean nothing and they are repacked every call
        leaves = [t for layer in params for pair in layer for t in pair]
        stable = all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in leaves)
        key = (S,) + tuple((t.data_ptr(), t._version) for t in leaves) if stable else None
        if key is None or cache.get("key") != key:
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
        if bnd is not None:
            # inverse bounding (bflow_jax_maf.py:220-222): + sigmoid_log_abs_det_jacobian(y_pre) + sum log(high - low),
            # written in terms of the bounded output: u = (y - low) / (high - low) = sigmoid(y_pre), log sigmoid' = log u (1 - u)
            u = (y - lo) / (hi - lo)
            log_j = log_j + (torch.log(u) + torch.log1p(-u)).sum(-1) + torch.log(hi - lo).sum()
        if params[0][0][0].dim() == 2:
            return y[0], log_j[0]
        return y, log_j

    def _engine_standard(best_params, standard_params, scale) -> FlowEngine:
        """Engine packed straight from the reference's posterior format {"standard_params": [S, P], "scale"}:
        the draw map of bflow_jax_maf.py:239-240 runs inside the pack kernels (nazb_pack_draw_map)."""
        u = torch.as_tensor(standard_params, dtype=torch.float32)
        S = u.shape[0]
        sc_key = ("t", scale.data_ptr(), scale._version) if isinstance(scale, torch.Tensor) and scale.numel() > 1 else float(scale)   # posterior["scale"] is [S]
        leaves = [t for layer in best_params for pair in layer for t in pair]
        stable = u.is_cuda and all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in leaves)
        key = (("std", S, u.data_ptr(), u._version, sc_key) + tuple((t.data_ptr(), t._version) for t in leaves)) if stable else None
        if key is None or cache.get("key") != key:
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


# ------------------------------------------------------------------------------------------------
# Bayesian flow over the MLE weights (bflow_jax_maf.py:227-268) without numpyro: the same three callables and unravel_fn,
# as plain objects whose draws / densities are torch tensors on the GPU.
# ------------------------------------------------------------------------------------------------
def ravel_pytree(params):
    """jax.flatten_util.ravel_pytree for the `[L][n_lin](W, b)` pytrees of this API -> (flat [P], unravel_fn).
    unravel_fn accepts [P] (one draw) or [S, P] (what `jax.vmap(unravel_fn)` is used for upstream, :332,:341,:349,:402)."""
    leaves = [t for layer in params for pair in layer for t in pair]
    shapes = [tuple(t.shape) for t in leaves]
    sizes = [int(t.numel()) for t in leaves]
    structure = [len(layer) for layer in params]
    flat = torch.cat([torch.as_tensor(t).reshape(-1) for t in leaves])

    def unravel_fn(v):
        v = torch.as_tensor(v)
        lead = tuple(v.shape[:-1])
        out, off, it = [], 0, iter(range(len(leaves)))
        for n_lin in structure:
            lay = []
            for _ in range(n_lin):
                iw = next(it); W = v[..., off:off + sizes[iw]].reshape(lead + shapes[iw]); off += sizes[iw]
                ib = next(it); b = v[..., off:off + sizes[ib]].reshape(lead + shapes[ib]); off += sizes[ib]
                lay.append((W, b))
            out.append(lay)
        return out

    return flat, unravel_fn


class _BayesModel:
    """`model(scale_max=..., prior=False)` of bflow_jax_maf.py:237-247.  Sites: "scale" (deterministic scale_max, or
    Uniform(0, scale_max) per draw / per parameter), "standard_params" ~ Uniform(-1, 1)^P, "params" = theta_MLE (1 + scale u),
    factor "log_l" = log_prob(params) (0 when prior=True).  Calling it draws ONE trace and returns its sites; `draw(S)` draws S
    at once in the posterior-file format {"scale", "standard_params"} that `lp_standard` / `pack_draw_map` consume directly."""

    def __init__(self, flat_params, log_prob, scale_max, multi_scale, fixed_scale):
        self.flat_params, self.log_prob = flat_params, log_prob
        self.scale_max, self.multi_scale, self.fixed_scale = scale_max, multi_scale, fixed_scale

    def draw(self, num_samples: int, scale_max=None, generator=None):
        sm = self.scale_max if scale_max is None else scale_max
        fp = self.flat_params
        P = fp.numel()
        if self.fixed_scale:
            scale = torch.full((num_samples,), float(sm), device=fp.device)
        elif self.multi_scale:
            scale = torch.rand((num_samples, P), device=fp.device, generator=generator) * float(sm)
        else:
            scale = torch.rand((num_samples,), device=fp.device, generator=generator) * float(sm)
        u = torch.rand((num_samples, P), device=fp.device, generator=generator) * 2.0 - 1.0
        return {"scale": scale, "standard_params": u}

    def params_of(self, sites):
        sc = sites["scale"]
        sc = sc if sc.dim() == 2 else sc.reshape(-1, 1)
        return self.flat_params * (1.0 + sc * sites["standard_params"])      # :240, same operation order

    def __call__(self, scale_max=None, prior: bool = False, anealed: bool = False, generator=None):
        sites = self.draw(1, scale_max, generator)
        sites["params"] = self.params_of(sites)[0]
        sites["log_l"] = torch.zeros((), device=self.flat_params.device) if prior else self.log_prob(sites["params"])
        return sites


class _BayesGuide:
    """`guide(scale_max=..., svi_params=None)` of bflow_jax_maf.py:249-260: standard_params ~ TruncatedNormal(mu_param_q,
    sigma_param_q, -1, 1) with mu in (-0.95, 0.95), sigma in (0, 1) (initial values 0 and 1).  `draw(S, uniform)` returns
    the draws and log q(u) from the libnazb guide kernel (nazb_truncnorm_sample)."""

    def __init__(self, flat_params, scale_max, multi_scale, fixed_scale):
        self.flat_params, self.scale_max, self.multi_scale, self.fixed_scale = flat_params, scale_max, multi_scale, fixed_scale
        self.params = {"mu_param_q": torch.zeros_like(flat_params), "sigma_param_q": torch.ones_like(flat_params)}

    def draw(self, num_samples: int, svi_params=None, uniform=None, generator=None):
        from ..stats import truncnorm_sample
        q = self.params if svi_params is None else svi_params
        mu = torch.as_tensor(q["mu_param_q"]).clamp(-0.95, 0.95)
        sg = torch.as_tensor(q["sigma_param_q"]).clamp(1e-6, 1.0)
        fp = self.flat_params
        if uniform is None:
            uniform = torch.rand((num_samples, fp.numel()), device=fp.device, generator=generator).clamp_(1e-7, 1 - 1e-7)
        u, log_q = truncnorm_sample(uniform, mu, sg, -1.0, 1.0)
        return {"scale": torch.full((num_samples,), float(self.scale_max), device=fp.device), "standard_params": u, "log_q": log_q}

    def __call__(self, scale_max=None, svi_params=None, generator=None):
        sites = self.draw(1, svi_params, generator=generator)
        sm = self.scale_max if scale_max is None else scale_max
        sites["params"] = self.flat_params * (1.0 + float(sm) * sites["standard_params"][0])
        return sites


def bayesian_normalizing_flow(flow_lp, best_params, scale_max=1.0, multi_scale=False, avg=False, fixed_scale=True,
                              return_log_l=False):
    """bflow_jax_maf.py:227-268 -> (model, guide, guided_model, unravel_fn[, log_prob]).
    `log_prob(flat) = flow_lp(unravel(flat)).sum()` (or .mean() with avg=True) — `flow_lp` is `make_normalizing_flow(...)["lp"]`,
    i.e. the libnazb log_prob kernel.  numpyro is not part of this stack: model / guide are plain callables (see their
    docstrings) whose sites are torch tensors; NUTS / SVI drivers stay upstream (out of scope)."""
    leaves = [t for layer in best_params for pair in layer for t in pair]
    dev = next((t.device for t in leaves if isinstance(t, torch.Tensor) and t.is_cuda), torch.device("cpu"))
    flat_params, unravel_fn = ravel_pytree([[(torch.as_tensor(W, dtype=torch.float32).to(dev), torch.as_tensor(b, dtype=torch.float32).to(dev))
                                             for (W, b) in layer] for layer in best_params])
    print(f"model complexity: {flat_params.numel() * (1 if not multi_scale else 2)}")

    def log_prob(params):
        lp = flow_lp(unravel_fn(params))
        return lp.sum(-1) if not avg else lp.mean(-1)

    model = _BayesModel(flat_params, log_prob, scale_max, multi_scale, fixed_scale)
    guide = _BayesGuide(flat_params, scale_max, multi_scale, fixed_scale)

    def guided_model(scale_max=scale_max, scale_sigma_init=None, scale_mean_init=None, svi_params=None, guide_fn=None):
        assert svi_params is not None and guide_fn is not None
        random_params = guide_fn(scale_max=scale_max, svi_params=svi_params)
        random_params = random_params["params"] if isinstance(random_params, dict) else random_params
        return {"params": random_params, "log_l": log_prob(random_params)}

    if not return_log_l:
        return model, guide, guided_model, unravel_fn
    return model, guide, guided_model, unravel_fn, log_prob


def quantile_bin_edges(theta_true, bin_count):
    """Equal-probability bin edges per axis: what physt's "quantile" binning (h2 / h in bflow_jax_maf.py:409,419) produces —
    np.percentile of each coordinate at linspace(0, 100, bins + 1).  physt is not part of this image; restated here."""
    import numpy as np
    th = np.asarray(theta_true, dtype=np.float64)
    return [np.percentile(th[:, i], np.linspace(0.0, 100.0, int(bin_count[i]) + 1)) for i in range(th.shape[1])]


def calibrate(ppds, theta_true, nq, cs, fthin=10, itype="hpd", twod=True, ranges=None, generator=None):
    """bflow_jax_maf.py:406-465: empirical coverage of the per-bin credible intervals of the posterior-predictive draws.
    ppds [S, N, D] (numpy or torch; moved to the GPU once), theta_true [Nt, D].  The per-draw N-D histograms and the HPD /
    equal-tailed intervals across draws run in libnazb (nazb_histogramdd, nazb_hpd) instead of one np.histogram2d per draw.
    Returns the coverage per credible level in `cs` (numpy array)."""
    import numpy as np
    from ..stats import histogramdd_draws, hpd_draws
    theta_true = np.asarray(theta_true, dtype=np.float64)
    Dd = theta_true.shape[-1]
    if twod:
        nbins = int(np.sqrt(nq))
        assert Dd == 2
    else:
        nbins = int(nq ** (1.0 / Dd))
        assert ranges is not None
        bad = np.zeros(len(theta_true), dtype=bool)
        for i in range(Dd):
            bad |= (theta_true[:, i] < ranges[i][0]) | (theta_true[:, i] > ranges[i][1])
        theta_true = theta_true[~bad, :]
    edges = quantile_bin_edges(theta_true, [nbins] * Dd)
    counts_true, _ = np.histogramdd(theta_true, bins=edges)
    vol = np.ones_like(counts_true)
    for i, e in enumerate(edges):
        shp = [1] * Dd
        shp[i] = -1
        vol = vol * np.diff(e).reshape(shp)
    den = counts_true / vol / len(theta_true)              # hist.densities / len(theta_true)  (:413,:422)
    n_empty = int((den.flatten() <= 0).sum())
    dev = torch.device("cuda")
    pp = torch.as_tensor(ppds, dtype=torch.float32).to(dev)
    S = pp.shape[0]
    den_t = torch.as_tensor(den, dtype=torch.float32, device=dev)
    ez = torch.zeros((len(cs),) + tuple(den.shape), dtype=torch.float64, device=dev)
    gen = generator
    for _ in range(fthin):
        idx = torch.randint(0, S, (int(S / fthin),), device=dev, generator=gen)          # np.random.choice(len(ppds), size=...)
        _, dens = histogramdd_draws(pp[idx], edges, density=True)
        for ci, c in enumerate(cs):
            if itype == "hpd":
                lo_hi = hpd_draws(dens, 1.0 - float(c))
            elif itype == "eqt":
                q = torch.tensor([0.5 - float(c) / 2.0, 0.5 + float(c) / 2.0], device=dev, dtype=dens.dtype)
                lo_hi = torch.quantile(dens, q, dim=0)
            else:
                raise ValueError(itype)
            ez[ci] += ((den_t < lo_hi[1]) & (den_t > lo_hi[0])).double() / fthin
    cov = ez.reshape(len(cs), -1).sum(-1) / (nq - n_empty)
    return cov.cpu().numpy()
