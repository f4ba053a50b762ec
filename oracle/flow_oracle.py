"""CPU oracle for naz's draw-batched flow-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``naz_b200/`` imports it; the product path fails loudly when the CUDA library is missing.

PARITY: MAF BRANCH PINNED BY REFERENCE OUTPUTS, SPLINE BRANCH UNPINNED.  The reference (AnaryaRay1/naz) ships
no tests, golden vectors or fixtures for this path, and its arithmetic lives partly in ``pyro-ppl`` (un-vendored,
unpinned, not installed here; neither is jax).  What pins this oracle:
  * MAF branch — outputs of the reference itself: ``tools/make_reference_goldens.py`` EXECUTES the reference's own
    src/naz/flows/bflow_jax_maf.py (create_mask :52-72, masked_linear :74-77, nn_fn :135-165, forward_fn / inverse_fn
    :173-194, make_normalizing_flow log_prob / sample :196-225) on CPU with numpy standing in for jax.numpy, and commits
    masks, log_prob and sampler outputs as tests/golden/ref_twin_*.npz; tests/test_oracle_cpu.py requires this oracle to
    reproduce them to 1e-10 (masks bit-identical) and tests/test_gpu_parity.py checks the CUDA path against them;
  * the spline / coupling / Permute / BatchNorm branches restate pyro-ppl 1.9's published
    algorithm (pyro/distributions/transforms/spline.py ``_monotonic_rational_spline``,
    ``SplineAutoregressive``; pyro/nn/auto_reg_nn.py ``create_mask``), anchored on the reference's
    call sites src/naz/flows/transforms.py:142-159,180-197 and src/naz/flows/flow.py:37-79 — pyro is absent, so there
    is nothing of the reference to execute for them: PARITY UNPINNED for splines;
  * an independent second restatement (oracle/pyro_style.py, torch modules on the real
    ``torch.distributions.TransformedDistribution``) must agree with this one to fp64 round-off;
  * mathematical identities checked in tests/: round trip, log-det == autograd slogdet,
    triangular Jacobian in permutation order, spline derivative == autograd, 2-D density
    integrating to 1.

Everything is plain numpy, dtype-generic (float32 mirrors the reference's fp32 path,
src/naz/utils.py:7; float64 is the "truth" copy parity tests compare against).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# MADE degrees and masks   (bflow_jax_maf.py:48-72; pyro.nn.auto_reg_nn.create_mask)
# --------------------------------------------------------------------------------------
def sample_mask_indices(input_dim: int, hidden_dim: int) -> np.ndarray:
    """round(linspace(1, input_dim, hidden_dim)) with torch's fp32 linspace + half-to-even
    rounding (bflow_jax_maf.py:48-50, simple=True branch; pyro ``sample_mask_indices``)."""
    import torch  # torch.linspace's fp32 symmetric formula is what pyro executes

    idx = torch.linspace(1, input_dim, steps=hidden_dim, dtype=torch.float32)
    return torch.round(idx).numpy().astype(np.float64)


def made_degrees(D: int, C: int, hidden: Sequence[int], perm: np.ndarray, M: int):
    """Degrees of input / hidden / output units (bflow_jax_maf.py:57-66)."""
    perm = np.asarray(perm, dtype=np.int64)
    var_index = np.empty(D, dtype=np.float64)
    var_index[perm] = np.arange(1, D + 1, dtype=np.float64)      # 1 + rank
    in_deg = np.concatenate([np.zeros(C), var_index])
    if C > 0:
        hid_deg = [sample_mask_indices(D, h) - 1 for h in hidden]
    else:
        hid_deg = [sample_mask_indices(D - 1, h) for h in hidden]
    out_deg = np.tile(var_index, M)                              # slot-major: index m*D + d
    return in_deg, hid_deg, out_deg


def create_masks(D: int, C: int, hidden: Sequence[int], perm: np.ndarray, M: int):
    """masks[k] has shape [out_k, in_k]; mask_skip [M*D, C+D] (bflow_jax_maf.py:52-72)."""
    in_deg, hid_deg, out_deg = made_degrees(D, C, hidden, perm, M)
    masks = [(hid_deg[0][:, None] >= in_deg[None, :]).astype(np.float32)]
    for i in range(1, len(hidden)):
        masks.append((hid_deg[i][:, None] >= hid_deg[i - 1][None, :]).astype(np.float32))
    masks.append((out_deg[:, None] > hid_deg[-1][None, :]).astype(np.float32))
    mask_skip = (out_deg[:, None] > in_deg[None, :]).astype(np.float32)
    return masks, mask_skip


# --------------------------------------------------------------------------------------
# Flow description + weights
# --------------------------------------------------------------------------------------
@dataclass
class FlowSpec:
    """Static description of one flow (mirrors the arguments of naz's factories,
    transforms.py:133 ``masked_affine_autoregressive`` / :165 ``neural_spline_autoregressive``)."""
    kind: str                      # "maf" | "nsa"
    D: int
    C: int
    hidden: List[int]
    L: int
    perms: np.ndarray              # int64 [L, D]
    count_bins: int = 8
    order: str = "quadratic"       # naz default (transforms.py:165); pyro's own default is "linear"
    bound: float = 3.0
    clip: Tuple[float, float] = (-5.0, 3.0)
    activation: str = "tanh"       # naz default nn.Tanh() (transforms.py:133,165)
    masks_override: Optional[list] = None   # [L][n_lin] masks that are not the canonical MADE masks of `perms` (coupling layers)

    @property
    def M(self) -> int:
        if self.kind == "maf":
            return 2
        K = self.count_bins
        return 3 * K - 1 if self.order == "quadratic" else 4 * K - 1

    def masks(self):
        if self.masks_override is not None:
            return self.masks_override
        return [create_masks(self.D, self.C, self.hidden, self.perms[l], self.M)[0] for l in range(self.L)]

    def n_params(self) -> int:
        dims = [self.D + self.C] + list(self.hidden) + [self.M * self.D]
        return self.L * sum(dims[i + 1] * dims[i] + dims[i + 1] for i in range(len(dims) - 1))


def init_weights(spec: FlowSpec, rng: np.random.Generator, dtype=np.float32):
    """Synthetic weights per SURVEY §8(d): W ~ N(0, 1/sqrt(fan_in)), b ~ N(0, 0.01).
    Returns [L][n_lin] of (W[out,in], b[out]) — the per-draw slice of the reference's pytree
    (bflow_jax_maf.py:26-46)."""
    dims = [spec.D + spec.C] + list(spec.hidden) + [spec.M * spec.D]
    params = []
    for _ in range(spec.L):
        layer = []
        for i in range(len(dims) - 1):
            W = rng.normal(0.0, 1.0 / math.sqrt(dims[i]), size=(dims[i + 1], dims[i])).astype(dtype)
            b = rng.normal(0.0, 0.01, size=(dims[i + 1],)).astype(dtype)
            layer.append((W, b))
        params.append(layer)
    return params


def perturb_draws(params, S: int, scale: float, rng: np.random.Generator, dtype=np.float32,
                  u: Optional[np.ndarray] = None):
    """theta_s = theta_0 * (1 + scale * u_s), u ~ U(-1, 1)  (bflow_jax_maf.py:239-240).
    Returns the batched pytree [L][n_lin] of (W[S,out,in], b[S,out])."""
    out = []
    for layer in params:
        lay = []
        for (W, b) in layer:
            uW = rng.uniform(-1, 1, size=(S,) + W.shape)
            ub = rng.uniform(-1, 1, size=(S,) + b.shape)
            lay.append(((W[None] * (1.0 + scale * uW)).astype(dtype), (b[None] * (1.0 + scale * ub)).astype(dtype)))
        out.append(lay)
    return out


def slice_draw(draws, s: int):
    """this_p = [[(W[i], b[i]) ...] ...]  (calibrate.py:148)."""
    return [[(W[s], b[s]) for (W, b) in layer] for layer in draws]


# --------------------------------------------------------------------------------------
# Conditioner  (bflow_jax_maf.py:74-77 masked_linear, :135-165 nn_fn; dropout transforms.py:38-43)
# --------------------------------------------------------------------------------------
def _act(h, name):
    if name == "tanh":
        return np.tanh(h)
    if name == "relu":
        return np.maximum(h, 0)
    raise ValueError(name)


def conditioner(x, layer_params, masks, context=None, keep=None, p_drop=0.0, activation="tanh"):
    """MADE pass. ``x`` [N, D]; ``context`` [N, C] | [C] | None, concatenated FIRST
    (bflow_jax_maf.py:141-142).  ``keep`` [n_hidden, H] 0/1 per-unit inverted-dropout masks for
    this flow layer (explicit per-draw masks).  Returns out [N, M, D]."""
    dt = x.dtype
    if context is not None:
        ctx = np.broadcast_to(np.asarray(context, dtype=dt), x.shape[:-1] + (np.shape(context)[-1],))
        h = np.concatenate([ctx, x], axis=-1)
    else:
        h = x
    n_lin = len(layer_params)
    for k in range(n_lin - 1):
        W, b = layer_params[k]
        h = _act(h @ (W.astype(dt) * masks[k].astype(dt)).T + b.astype(dt), activation)
        if keep is not None:
            h = h * (keep[k].astype(dt) / dt.type(1.0 - p_drop))
    W, b = layer_params[-1]
    out = h @ (W.astype(dt) * masks[-1].astype(dt)).T + b.astype(dt)
    D = x.shape[-1]
    return out.reshape(x.shape[:-1] + (out.shape[-1] // D, D))       # [..., M, D], flat index m*D + d


# --------------------------------------------------------------------------------------
# Affine autoregressive  (bflow_jax_maf.py:173-194; pyro AffineAutoregressive)
# --------------------------------------------------------------------------------------
def affine_forward(x, layer_params, masks, context=None, clip=(-5.0, 3.0), **kw):
    out = conditioner(x, layer_params, masks, context, **kw)
    mean, log_scale = out[..., 0, :], out[..., 1, :]
    log_scale = np.clip(log_scale, x.dtype.type(clip[0]), x.dtype.type(clip[1]))
    y = mean + x * np.exp(log_scale)
    return y, log_scale.sum(-1)


def affine_inverse(y, layer_params, masks, perm, context=None, clip=(-5.0, 3.0), **kw):
    """D sequential full conditioner passes in ``perm`` order (bflow_jax_maf.py:181-194)."""
    x = np.zeros_like(y)
    lo, hi = y.dtype.type(clip[0]), y.dtype.type(clip[1])
    log_scale = None
    for idx in np.asarray(perm):
        out = conditioner(x, layer_params, masks, context, **kw)
        mean, log_scale = out[..., 0, :], out[..., 1, :]
        inv_scale = np.exp(-np.clip(log_scale[..., idx], lo, hi))
        x[..., idx] = (y[..., idx] - mean[..., idx]) * inv_scale
    log_scale = np.clip(log_scale, lo, hi)
    return x, log_scale.sum(-1)


# --------------------------------------------------------------------------------------
# Monotone rational spline  (pyro-ppl 1.9 spline.py::_monotonic_rational_spline)
# --------------------------------------------------------------------------------------
def _softmax(a):
    a = a - a.max(-1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(-1, keepdims=True)


def _softplus(a):
    # torch.nn.functional.softplus: beta=1, threshold=20
    return np.where(a > 20, a, np.log1p(np.exp(np.minimum(a, 20))))


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def _calculate_knots(lengths, lower, upper):
    dt = lengths.dtype
    knots = np.cumsum(lengths, axis=-1, dtype=dt)
    knots = np.concatenate([np.zeros(knots.shape[:-1] + (1,), dtype=dt), knots], axis=-1)
    knots = dt.type(upper - lower) * knots + dt.type(lower)
    knots[..., 0] = lower
    knots[..., -1] = upper
    lengths = knots[..., 1:] - knots[..., :-1]
    return lengths, knots


def _select(a, idx):
    idx = np.clip(idx, 0, a.shape[-1] - 1)
    return np.take_along_axis(a, idx[..., None], axis=-1)[..., 0]


def rational_spline(inputs, widths, heights, derivatives, lambdas=None, inverse=False, bound=3.0,
                    min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3, min_lambda=0.025, eps=1e-6):
    """``inputs`` [...]; widths/heights [..., K] (already softmaxed), derivatives [..., K-1]
    (already softplus'd), lambdas [..., K] (already sigmoided) or None (quadratic order)."""
    dt = inputs.dtype
    t = dt.type
    left, right, bottom, top = -bound, bound, -bound, bound
    inside = (inputs >= t(left)) & (inputs <= t(right))
    K = widths.shape[-1]
    widths = t(min_bin_width) + t(1.0 - min_bin_width * K) * widths
    heights = t(min_bin_height) + t(1.0 - min_bin_height * K) * heights
    derivatives = t(min_derivative) + derivatives
    widths, cumwidths = _calculate_knots(widths, left, right)
    heights, cumheights = _calculate_knots(heights, bottom, top)
    pad = np.full(derivatives.shape[:-1] + (1,), 1.0 - min_derivative, dtype=dt)
    derivatives = np.concatenate([pad, derivatives, pad], axis=-1)
    knots = (cumheights if inverse else cumwidths) + t(eps)
    bin_idx = (inputs[..., None] >= knots).sum(-1) - 1

    in_w = _select(widths, bin_idx)
    in_cw = _select(cumwidths, bin_idx)
    in_ch = _select(cumheights, bin_idx)
    in_delta = _select(heights / widths, bin_idx)
    in_d = _select(derivatives, bin_idx)
    in_d1 = _select(derivatives[..., 1:], bin_idx)
    in_h = _select(heights, bin_idx)

    with np.errstate(all="ignore"):
        if lambdas is not None:
            lambdas = t(1 - 2 * min_lambda) * lambdas + t(min_lambda)
            lam = _select(lambdas, bin_idx)
            wa = t(1.0)
            wb = np.sqrt(in_d / in_d1) * wa
            wc = (lam * wa * in_d + (1 - lam) * wb * in_d1) / in_delta
            ya = in_ch
            yb = in_h + in_ch
            yc = ((1 - lam) * wa * ya + lam * wb * yb) / ((1 - lam) * wa + lam * wb)
            if inverse:
                le = (inputs <= yc).astype(dt)
                gt = (inputs > yc).astype(dt)
                numerator = (lam * wa * (ya - inputs)) * le + ((wc - lam * wb) * inputs + lam * wb * yb - wc * yc) * gt
                denominator = ((wc - wa) * inputs + wa * ya - wc * yc) * le + ((wc - wb) * inputs + wb * yb - wc * yc) * gt
                theta = numerator / denominator
                outputs = theta * in_w + in_cw
                dnum = (wa * wc * lam * (yc - ya) * le + wb * wc * (1 - lam) * (yb - yc) * gt) * in_w
                logabsdet = np.log(dnum) - 2 * np.log(np.abs(denominator))
            else:
                theta = (inputs - in_cw) / in_w
                le = (theta <= lam).astype(dt)
                gt = (theta > lam).astype(dt)
                numerator = (wa * ya * (lam - theta) + wc * yc * theta) * le + (wc * yc * (1 - theta) + wb * yb * (theta - lam)) * gt
                denominator = (wa * (lam - theta) + wc * theta) * le + (wc * (1 - theta) + wb * (theta - lam)) * gt
                outputs = numerator / denominator
                dnum = (wa * wc * lam * (yc - ya) * le + wb * wc * (1 - lam) * (yb - yc) * gt) / in_w
                logabsdet = np.log(dnum) - 2 * np.log(np.abs(denominator))
        else:
            if inverse:
                a = (inputs - in_ch) * (in_d + in_d1 - 2 * in_delta) + in_h * (in_delta - in_d)
                b = in_h * in_d - (inputs - in_ch) * (in_d + in_d1 - 2 * in_delta)
                c = -in_delta * (inputs - in_ch)
                disc = b * b - 4 * a * c
                disc = np.where(inside, disc, t(0))
                root = (2 * c) / (-b - np.sqrt(disc))
                outputs = root * in_w + in_cw
                tomt = root * (1 - root)
                den = in_delta + (in_d + in_d1 - 2 * in_delta) * tomt
                dnum = in_delta ** 2 * (in_d1 * root ** 2 + 2 * in_delta * tomt + in_d * (1 - root) ** 2)
                logabsdet = -(np.log(dnum) - 2 * np.log(den))
            else:
                theta = (inputs - in_cw) / in_w
                tomt = theta * (1 - theta)
                num = in_h * (in_delta * theta ** 2 + in_d * tomt)
                den = in_delta + (in_d + in_d1 - 2 * in_delta) * tomt
                outputs = in_ch + num / den
                dnum = in_delta ** 2 * (in_d1 * theta ** 2 + 2 * in_delta * tomt + in_d * (1 - theta) ** 2)
                logabsdet = np.log(dnum) - 2 * np.log(den)
    outputs = np.where(inside, outputs, inputs)
    logabsdet = np.where(inside, logabsdet, t(0))
    return outputs.astype(dt), logabsdet.astype(dt)


def spline_params(out, K, order):
    """Split conditioner output [..., M, D] into (w, h, d, lambda) each [..., D, .]
    (pyro SplineAutoregressive._params: transpose(-1,-2), softmax/softmax/softplus/sigmoid)."""
    w = np.swapaxes(out[..., 0:K, :], -1, -2)
    h = np.swapaxes(out[..., K:2 * K, :], -1, -2)
    d = np.swapaxes(out[..., 2 * K:3 * K - 1, :], -1, -2)
    lam = None
    if order == "linear":
        lam = _sigmoid(np.swapaxes(out[..., 3 * K - 1:4 * K - 1, :], -1, -2))
    return _softmax(w), _softmax(h), _softplus(d), lam


def spline_forward(x, layer_params, masks, context=None, K=8, order="quadratic", bound=3.0, **kw):
    out = conditioner(x, layer_params, masks, context, **kw)
    w, h, d, lam = spline_params(out, K, order)
    y, ld = rational_spline(x, w, h, d, lam, inverse=False, bound=bound)
    return y, ld.sum(-1)


def spline_inverse(y, layer_params, masks, context=None, K=8, order="quadratic", bound=3.0, **kw):
    """D Jacobi sweeps x <- spline^-1(y; params(x)) from x = 0 (pyro SplineAutoregressive._inverse);
    cached log-det is the FORWARD log-det of the last sweep (ConditionedSpline._inverse negates)."""
    x = np.zeros_like(y)
    ld = None
    for _ in range(y.shape[-1]):
        out = conditioner(x, layer_params, masks, context, **kw)
        w, h, d, lam = spline_params(out, K, order)
        x, ld_inv = rational_spline(y, w, h, d, lam, inverse=True, bound=bound)
        ld = -ld_inv
    return x, ld.sum(-1)


# --------------------------------------------------------------------------------------
# Bounding transform  (transforms.py:20-27; twin bflow_jax_maf.py:95-105)
# --------------------------------------------------------------------------------------
def bounding_transform(x, low, high):
    u = (x - low) / (high - low)
    with np.errstate(all="ignore"):
        log_jac = -np.sum(np.log(u) + np.log1p(-u), axis=-1) - np.sum(np.log(high - low))
        y = np.log(u) - np.log1p(-u)          # torch.logit
    return y.astype(x.dtype), log_jac.astype(x.dtype)


def inverse_bounding_transform(y, low, high):
    return (_sigmoid(y) * (high - low) + low).astype(y.dtype)


# --------------------------------------------------------------------------------------
# Whole-flow log_prob / sample for ONE weight set  (flow.py:45-129; bflow_jax_maf.py:210-223)
# --------------------------------------------------------------------------------------
def _layer_kw(spec: FlowSpec, keep, p_drop, l):
    kw = dict(activation=spec.activation)
    if keep is not None:
        kw.update(keep=keep[l], p_drop=p_drop)
    return kw


def flow_inverse(spec: FlowSpec, params, x, context=None, bounds=None, keep=None, p_drop=0.0, layer_affine=None):
    """Reference ``log_prob`` direction: returns (z, lp).  Layers inverted in reverse order
    (torch TransformedDistribution.log_prob; bflow_jax_maf.py:211).  ``layer_affine`` = (a, b) [L, D]: the element-wise
    step x <- a[l] x + b[l] behind flow layer l in the sampling direction (eval-mode T.BatchNorm, transforms.py:157-158)."""
    dt = x.dtype
    masks = spec.masks()
    if bounds is not None:
        y, log_jac = bounding_transform(x, np.asarray(bounds[0], dt), np.asarray(bounds[1], dt))
    else:
        y, log_jac = x, dt.type(0)
    ld_total = np.zeros(x.shape[:-1], dtype=dt)
    for l in reversed(range(spec.L)):
        kw = _layer_kw(spec, keep, p_drop, l)
        if layer_affine is not None:
            a, b = np.asarray(layer_affine[0][l], dt), np.asarray(layer_affine[1][l], dt)
            y = (y - b) / a
            ld_total = ld_total + np.log(a).sum().astype(dt)
        if spec.kind == "maf":
            y, ld = affine_inverse(y, params[l], masks[l], spec.perms[l], context, clip=spec.clip, **kw)
        else:
            y, ld = spline_inverse(y, params[l], masks[l], context, K=spec.count_bins, order=spec.order,
                                   bound=spec.bound, **kw)
        ld_total = ld_total + ld
    z = y
    lp = -(0.5 * z * z).sum(-1) - dt.type(0.5 * spec.D * LOG_2PI) - ld_total + log_jac
    return z, lp.astype(dt)


def flow_forward(spec: FlowSpec, params, z, context=None, bounds=None, keep=None, p_drop=0.0, layer_affine=None):
    """Reference ``sample`` direction (one conditioner pass per layer): returns (x, sum log-det)."""
    dt = z.dtype
    masks = spec.masks()
    x = z
    ld_total = np.zeros(z.shape[:-1], dtype=dt)
    for l in range(spec.L):
        kw = _layer_kw(spec, keep, p_drop, l)
        if spec.kind == "maf":
            x, ld = affine_forward(x, params[l], masks[l], context, clip=spec.clip, **kw)
        else:
            x, ld = spline_forward(x, params[l], masks[l], context, K=spec.count_bins, order=spec.order,
                                   bound=spec.bound, **kw)
        ld_total = ld_total + ld
        if layer_affine is not None:
            a, b = np.asarray(layer_affine[0][l], dt), np.asarray(layer_affine[1][l], dt)
            x = a * x + b
            ld_total = ld_total + np.log(a).sum().astype(dt)
    if bounds is not None:
        x = inverse_bounding_transform(x, np.asarray(bounds[0], dt), np.asarray(bounds[1], dt))
    return x, ld_total


# --------------------------------------------------------------------------------------
# Draw-batched drivers — the Python loops the CUDA path replaces
# (train_flows.py:414-420 predict; calibrate.py:147-150; compute_bic_simpler.py:116-120; mcdpflow.py:47-54)
# --------------------------------------------------------------------------------------
def log_prob_draws(spec, draws, x, context=None, bounds=None, keep=None, p_drop=0.0):
    """-> lp [S, N], z [S, N, D].  ``draws`` is [L][n_lin](W[S,..], b[S,..]); ``keep`` [S, L, n_hidden, H]."""
    S = draws[0][0][0].shape[0]
    lps, zs = [], []
    for s in range(S):
        z, lp = flow_inverse(spec, slice_draw(draws, s), x, context, bounds,
                             None if keep is None else keep[s], p_drop)
        lps.append(lp)
        zs.append(z)
    return np.stack(lps), np.stack(zs)


def sample_draws(spec, draws, z, context=None, bounds=None, keep=None, p_drop=0.0):
    """-> x [S, N, D], logdet [S, N].  ``z`` is [S, N, D] or [N, D] shared base noise."""
    S = draws[0][0][0].shape[0]
    xs, lds = [], []
    for s in range(S):
        zz = z[s] if z.ndim == 3 else z
        x, ld = flow_forward(spec, slice_draw(draws, s), zz, context, bounds,
                             None if keep is None else keep[s], p_drop)
        xs.append(x)
        lds.append(ld)
    return np.stack(xs), np.stack(lds)


# --------------------------------------------------------------------------------------
# Cross-draw reductions  (plot.py:272-275; bflow_jax_maf.py:474-475; pyro Importance)
# --------------------------------------------------------------------------------------
def logsumexp(a, axis=0):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    return (np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def posterior_predictive(lp, log_w=None):
    """log p_hat(x_n) = logsumexp_s(lp[s,n] + log w_s); w = 1/S by default (host mean of exp, plot.py:272-275)."""
    S = lp.shape[0]
    if log_w is None:
        log_w = np.full((S,), -math.log(S))
    return logsumexp(lp.astype(np.float64) + np.asarray(log_w, np.float64)[:, None], axis=0)


def importance(sum_lp, log_prior, log_q):
    """log w_s = log p(theta_s) + sum_n lp[s,n] - log q(theta_s); log Z = lse(log w) - log S;
    ESS = exp(2 lse(log w) - lse(2 log w))   (pyro Importance; train_flows.py:360-378)."""
    log_w = np.asarray(log_prior, np.float64) + np.asarray(sum_lp, np.float64) - np.asarray(log_q, np.float64)
    S = log_w.shape[0]
    log_z = logsumexp(log_w) - math.log(S)
    ess = math.exp(2 * logsumexp(log_w) - logsumexp(2 * log_w))
    return log_w, float(log_z), float(ess)


def compute_bic(log_ls, N, complexity):
    """bflow_jax_maf.py:474-475."""
    return complexity * math.log(N) - 2.0 * float(np.max(log_ls))
