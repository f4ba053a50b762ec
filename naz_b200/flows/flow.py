"""`NormalizingFlow` — drop-in for src/naz/flows/flow.py:24-129 on the libnazb CUDA path, plus the
draw-batched entry points (`log_prob_draws`, `sample_draws`) that replace the reference's per-draw
Python loops (train_flows.py:414-420, calibrate.py:147-150, compute_bic_simpler.py:116-120)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ..engine import FlowEngine, FlowShape
from .transforms import masked_affine_autoregressive, neural_spline_autoregressive, neural_spline_coupling


def _continuous_free_form(*a, **k):
    raise NotImplementedError("'cnf' (FFJORD) is a continuous flow; BASELINE.json's north_star covers discrete flows only")


flow_makers = {"maf": masked_affine_autoregressive, "nsa": neural_spline_autoregressive,
               "nsc": neural_spline_coupling, "cnf": _continuous_free_form}


class _FlowDist:
    """Stand-in for pyro's (Conditional)TransformedDistribution: the reference only touches
    `.transforms` (train_flows.py:38,65) and `.condition(...)` (flow.py:76)."""

    def __init__(self, base_dist, transforms):
        self.base_dist = base_dist
        self.transforms = transforms

    def condition(self, context):
        return self


def draws_from_posterior_samples(posterior_samples: Dict[str, torch.Tensor], L: int, n_lin: int):
    """Reference format `"flow_{i}_{name}" -> tensor[S, ...]` (train_flows.py:71, bflow.py:72) -> pytree."""
    return [[(posterior_samples[f"flow_{i}_nn.layers.{j}.weight"], posterior_samples[f"flow_{i}_nn.layers.{j}.bias"])
             for j in range(n_lin)] for i in range(L)]


class NormalizingFlow(nn.Module):
    def __init__(self, flow_type, bounds, *flow_maker_args, embedding_net=None, engine: str = "auto", **flow_maker_kwargs):
        super().__init__()
        assert flow_type in list(flow_makers.keys())
        flow_maker = flow_makers[flow_type]
        self.flow_type = flow_type
        self.conditional = True if flow_maker_args[1] > 0 else False
        if embedding_net is not None:
            assert self.conditional
            self.embedding_net = embedding_net
        else:
            self.embedding_net = nn.Identity()
        self.bounds = bounds
        self.theta_dim, self.condition_dim = int(flow_maker_args[0]), int(flow_maker_args[1])
        self.flow, self.transforms, self.nets = flow_maker(*flow_maker_args, **flow_maker_kwargs)
        self.base_dist = torch.distributions.Normal(torch.zeros(self.theta_dim), torch.ones(self.theta_dim))
        self.flow_dist = _FlowDist(self.base_dist, self.transforms)
        hidden = self.nets[0].hidden_dims
        order = flow_maker_kwargs.get("order", "quadratic")
        count_bins = 8
        if flow_type == "nsa":
            count_bins = flow_maker_kwargs.get("count_bins", flow_maker_args[4] if len(flow_maker_args) > 4 else 8)
        kind = "maf" if flow_type == "maf" else ("nsa" if order == "quadratic" else "nsa_linear")
        self.shape = FlowShape(kind, self.theta_dim, self.condition_dim, list(hidden), len(self.nets), count_bins)
        self.dropout_p = flow_maker_kwargs.get("dropout_p", None)
        self._engine_kind = engine
        self._eng1: Optional[FlowEngine] = None
        self._eng1_key = None

    # ------------------------------------------------------------------ packing helpers
    def masks(self):
        return [[lin.mask for lin in arn.layers] for arn in self.nets]

    def perms(self):
        return torch.stack([arn.permutation.cpu() for arn in self.nets])

    def current_draw(self):
        return [[(lin.weight.detach(), lin.bias.detach()) for lin in arn.layers] for arn in self.nets]

    def _device(self):
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("naz_b200.NormalizingFlow evaluates on CUDA only (no CPU fallback): call flow.cuda() first")
        return p.device

    def _single_engine(self) -> FlowEngine:
        """Engine holding the module's current parameters (S = 1), re-packed when they change."""
        dev = self._device()
        key = (dev,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._eng1 is None or self._eng1_key != key:
            if self._eng1 is None or self._eng1.device != dev:
                self._eng1 = FlowEngine(self.shape, 1, device=dev, engine=self._engine_kind)
            self._eng1.pack(self.current_draw(), self.masks(), self.perms())
            self._eng1_key = key
        return self._eng1

    def make_engine(self, draws, keep=None, p_drop: float = 0.0, device=None) -> FlowEngine:
        """Engine holding S draws given as the reference pytree `[L][n_lin](W[S,out,in], b[S,out])`
        or as the `"flow_{i}_{name}"` dict."""
        if isinstance(draws, dict):
            draws = draws_from_posterior_samples(draws, len(self.nets), len(self.nets[0].layers))
        S = 1
        for layer in draws:
            for (W, b) in layer:
                if W.dim() == 3:
                    S = W.shape[0]
        if keep is not None:
            S = keep.shape[0]
        eng = FlowEngine(self.shape, S, device=device or self._device(), engine=self._engine_kind)
        eng.pack(draws, self.masks(), self.perms(), keep, p_drop)
        return eng

    def _cond(self, condition):
        if self.conditional:
            assert condition is not None
            return self.embedding_net(condition)
        return None

    # ------------------------------------------------------------------ reference API (flow.py:45-129)
    def log_prob(self, x, *args, condition=None, **kwargs):
        """flow.py:66-79.  Evaluated by libnazb on the module's CURRENT weights; the result is a plain tensor (no autograd
        graph — gradients come from FlowEngine.inverse_grad / the twin's value_and_grad).  A dropout flow in train() mode
        would apply a fresh nn.Dropout mask upstream: that stochastic path is `MCDPNormalizingFlow.sample_uncertain` /
        `log_prob_draws(keep=...)` here, so calling log_prob in that state raises instead of silently ignoring dropout."""
        if self.training and self.dropout_p not in (None, 0.0):
            raise RuntimeError("log_prob on a dropout flow in train() mode: call flow.eval() for the deterministic density, or "
                               "log_prob_draws(..., keep=masks, p_drop=p) for explicit MC-dropout masks")
        eng = self._single_engine()
        out = eng.inverse(x, self._cond(condition), self.bounds, want_lp=True)
        return out["lp"][0]

    def bounded_log_prob(self, x, *args, condition=None, **kwargs):
        if self.bounds is None:
            return self.log_prob(x, *args, condition=condition, **kwargs)
        lp = torch.full_like(x[:, 0], -math.inf)
        low, high = self.bounds["low"].to(x.device), self.bounds["high"].to(x.device)
        valid = ((x > low.expand(x.shape)) & (x < high.expand(x.shape))).all(dim=1)
        if valid.any():
            cond = condition
            if condition is not None and condition.dim() == 2 and condition.shape[0] == x.shape[0]:
                cond = condition[valid]
            lp[valid] = self.log_prob(x[valid, :], *args, condition=cond, **kwargs)
        return lp

    def average_log_prob(self, x, *args, condition=None, **kwargs):
        return torch.mean(self.bounded_log_prob(x, *args, condition=condition, **kwargs))

    def sample(self, *args, condition=None, base_noise=None, **kwargs):
        """`flow.sample([N], condition=c)`; `predict` upstream passes `(cond, [N])` positionally
        (train_flows.py:419) — both spellings are accepted."""
        shape = None
        for a in args:
            if isinstance(a, torch.Tensor) and condition is None and self.conditional:
                condition = a
            elif isinstance(a, (list, tuple, torch.Size)):
                shape = list(a)
        eng = self._single_engine()
        if base_noise is None:
            assert shape is not None, "sample shape required"
            n = int(torch.Size(shape).numel())
            base_noise = torch.randn((n, self.theta_dim), device=eng.device)
        x = eng.forward(base_noise.reshape(-1, self.theta_dim), self._cond(condition), self.bounds)[0]
        return x.reshape(*(shape if shape is not None else [x.shape[0]]), self.theta_dim)

    # ------------------------------------------------------------------ draw-batched entry points
    def log_prob_draws(self, x, draws, condition=None, reduce: Optional[str] = None, log_w=None, keep=None,
                       p_drop: float = 0.0, engine: Optional[FlowEngine] = None):
        """log p(x_n | theta_s) for all draws.  reduce=None -> [S,N]; "lse" -> posterior predictive
        log (1/S) sum_s p(x_n|theta_s) [N] (or weighted by log_w); "sum" -> sum_n lp[s,n] [S] (float64)."""
        eng = engine or self.make_engine(draws, keep, p_drop)
        cond = self._cond(condition)
        if reduce is None:
            return eng.inverse(x, cond, self.bounds, want_lp=True)["lp"]
        if reduce == "lse":
            out = eng.inverse(x, cond, self.bounds, want_lp=False, want_lse=True, log_w=log_w)
            return eng.lse_finish(out["lse_max"], out["lse_sum"], 0.0 if log_w is not None else -math.log(eng.S))
        if reduce == "sum":
            return eng.inverse(x, cond, self.bounds, want_lp=False, want_sum=True)["sum_n"]
        raise ValueError(reduce)

    def sample_draws(self, draws, n_or_noise, condition=None, keep=None, p_drop: float = 0.0,
                     engine: Optional[FlowEngine] = None, generator=None):
        """x[s, n, :] ~ p(. | theta_s).  `n_or_noise` is a sample count (fresh N(0,I) noise per draw) or
        the base noise itself, [S,N,D] or [N,D] (shared)."""
        eng = engine or self.make_engine(draws, keep, p_drop)
        if isinstance(n_or_noise, int):
            z = torch.randn((eng.S, n_or_noise, self.theta_dim), device=eng.device, generator=generator)
        else:
            z = n_or_noise
        return eng.forward(z, self._cond(condition), self.bounds)
