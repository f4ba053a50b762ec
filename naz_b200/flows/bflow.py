"""`BayesianNormalizingFlow` — src/naz/flows/bflow.py:20-137 on the libnazb path, without pyro.

Upstream the class is a pyro model: `prior_model` walks the parameters of every transform, samples each from a prior
centred on the MLE weights with width sigma = scale * |theta_MLE| (scale ~ Uniform(0, scale_max)), copies the sample
into the module, and `model` scores the data with the freshly set weights; pyro's MCMC / SVI / Importance drive it one
draw at a time.  Here the same prior (all four kinds: 'Uniform', 'Normal', 'TruncNorm', 'StandardNormal') is drawn for
S weight sets at once as the batched pytree the engine packs, and the data are scored for all S draws in one launch
(`log_prob_draws`).  `draw_param` / `set_param` / `prior_model` / `model` keep their reference signatures and meaning for
one draw; `prior_draws` / `log_joint_draws` are the batched forms the inference drivers should call.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from .flow import NormalizingFlow

PRIOR_KINDS = ("Uniform", "Normal", "TruncNorm", "StandardNormal")


def _std_normal_cdf(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _truncnorm_icdf(u, loc, scale, low, high):
    """priors/TruncatedNormal.py:41-47: uniform -> truncated normal through the normal CDF of the bounds."""
    cl, ch = _std_normal_cdf((low - loc) / scale), _std_normal_cdf((high - loc) / scale)
    p = u * (ch - cl) + cl
    return loc + scale * math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)


def _truncnorm_log_prob(y, loc, scale, low, high):
    cl, ch = _std_normal_cdf((low - loc) / scale), _std_normal_cdf((high - loc) / scale)
    return -0.5 * ((y - loc) / scale) ** 2 - torch.log(scale * math.sqrt(2.0 * math.pi)) - torch.log(ch - cl)


class BayesianNormalizingFlow(NormalizingFlow):
    def __init__(self, mle_flow, *args, prior_dist="Uniform", scale_max=0.1, set_grad=True, **kwargs):
        super().__init__(*args, **kwargs)
        if prior_dist not in PRIOR_KINDS:
            raise ValueError(f"prior_dist must be one of {PRIOR_KINDS}")
        self.mle_flow = mle_flow
        self.prior_dist = prior_dist
        self.scale_max = float(scale_max)
        self.set_grad = set_grad
        self.n_params = None
        self.param_bounds: Dict[str, tuple] = {}

    # ------------------------------------------------------------------ one parameter tensor
    def _prior_sample(self, mean, sigma, shape, generator=None):
        """-> (sample, log_prior summed over the parameter's own dims).  mean / sigma broadcast against `shape`."""
        a, b = mean - sigma, mean + sigma
        dev = mean.device
        ev = tuple(range(len(shape) - mean.dim(), len(shape)))          # the parameter's own (event) dims
        if self.prior_dist == "StandardNormal":
            x = torch.randn(shape, device=dev, generator=generator)
            lp = (-0.5 * x * x - 0.5 * math.log(2.0 * math.pi)).sum(ev)
        elif self.prior_dist == "Normal":
            x = mean + sigma * torch.randn(shape, device=dev, generator=generator)
            lp = (-0.5 * ((x - mean) / sigma) ** 2 - torch.log(sigma * math.sqrt(2.0 * math.pi))).sum(ev)
        elif self.prior_dist == "Uniform":
            x = a + (b - a) * torch.rand(shape, device=dev, generator=generator)
            lp = (-torch.log(b - a)).expand(shape).sum(ev)
        else:  # TruncNorm on [mean - sigma, mean + sigma]
            u = torch.rand(shape, device=dev, generator=generator).clamp_(1e-7, 1.0 - 1e-7)
            x = _truncnorm_icdf(u, mean, sigma, a, b)
            lp = _truncnorm_log_prob(x, mean, sigma, a, b).sum(ev)
        return x, lp

    def draw_param(self, name, param, mean, sigma, guide=False, guide_params=None, generator=None):
        """bflow.py:30-47 -> (sampled_param, a, b).  With guide=True the sample comes from the variational family
        TruncatedNormal(mean_q, sigma, a, b), mean_q = guide_params[f"{name}_mean_q"] (default: the prior mean)."""
        a, b = mean - sigma, mean + sigma
        if guide:
            mq = mean if guide_params is None else torch.as_tensor(guide_params[f"{name}_mean_q"]).to(mean.device)
            u = torch.rand(mean.shape, device=mean.device, generator=generator).clamp_(1e-7, 1.0 - 1e-7)
            return _truncnorm_icdf(u, mq, sigma, a, b), a, b
        x, _ = self._prior_sample(mean, sigma, tuple(mean.shape), generator)
        return x, a, b

    def set_param(self, param, new_param, grad=None):
        """bflow.py:49-55"""
        with torch.no_grad():
            param.copy_(new_param.to(param.device))
        if grad is not None and self.set_grad:
            param.requires_grad_(True)
            if param.grad is None:
                param.grad = torch.zeros_like(param)
            param.grad.copy_(grad.detach().clone().to(param.device))

    # ------------------------------------------------------------------ walks over all parameters
    def _named_pairs(self):
        """(site name, this flow's parameter, the MLE flow's parameter) in the reference's order and naming
        (`flow_{i}_{name}`, bflow.py:66-72; `embedding_0_{name}`, :80-86)."""
        for i, (t1, t2) in enumerate(zip(self.flow_dist.transforms, self.mle_flow.flow_dist.transforms)):
            mle = dict(t2.named_parameters())
            for name, param in t1.named_parameters():
                yield f"flow_{i}_{name}", param, mle[name]
        if not (self.embedding_net is None or isinstance(self.embedding_net, nn.Identity)):
            mle = dict(self.mle_flow.embedding_net.named_parameters())
            for name, param in self.embedding_net.named_parameters():
                yield f"embedding_0_{name}", param, mle[name]

    def prior_model(self, guide=False, set_param_bounds=False, guide_params=None, generator=None):
        """bflow.py:57-94: scale ~ Uniform(0, scale_max) (guide: TruncatedNormal(scale_mu_q, scale_sigma_q, 0, scale_max)),
        then every parameter from its prior with sigma = scale |theta_MLE|, copied into this module (prior draws only).
        Returns the sites {"scale", name: sample, ...}."""
        dev = next(self.parameters()).device
        if guide:
            mu = self.scale_max / 2 if guide_params is None else float(guide_params["scale_mu_q"])
            sg = self.scale_max / 4 if guide_params is None else float(guide_params["scale_sigma_q"])
            u = torch.rand((), device=dev, generator=generator).clamp_(1e-7, 1.0 - 1e-7)
            z = torch.zeros((), device=dev)
            scale = _truncnorm_icdf(u, z + mu, z + sg, z, z + self.scale_max)
        else:
            scale = torch.rand((), device=dev, generator=generator) * self.scale_max
        sites, n = {"scale": scale}, 0
        for name, param, p_mle in self._named_pairs():
            mean = p_mle.data.to(dev)
            n += mean.numel()
            sigma = scale.expand(param.shape) * mean.abs()
            sampled, a, b = self.draw_param(name, param, mean, sigma, guide=guide, guide_params=guide_params, generator=generator)
            if set_param_bounds:
                self.param_bounds[name] = (a, b)
                continue
            sites[name] = sampled
            if not guide:
                self.set_param(param, sampled, grad=p_mle.grad)
        if self.n_params is None:
            self.n_params = n
        return sites

    def model(self, theta, condition=None, generator=None):
        """bflow.py:96-111: one prior draw of the weights, then the data log-likelihood under it (libnazb log_prob).
        Returns sum_n log p(theta_n | weights) (the "log_l" site plus the "log_jac_bounding" factor)."""
        self.prior_model(guide=False, set_param_bounds=False, generator=generator)
        return self.log_prob(theta, condition=condition).sum()

    def svi_guide(self, guide_params=None, generator=None):
        return self.prior_model(guide=True, set_param_bounds=False, guide_params=guide_params, generator=generator)

    def param_transforms(self):
        """bflow.py:116-119: per-site bijection unconstrained -> (a, b): sigmoid then affine, as (forward, inverse) callables."""
        if self.param_bounds == {}:
            self.prior_model(guide=False, set_param_bounds=True)

        def mk(a, b):
            return (lambda u: torch.sigmoid(u) * (b - a) + a, lambda x: torch.logit((x - a) / (b - a)))
        return {k: mk(a, b) for k, (a, b) in self.param_bounds.items()}

    # ------------------------------------------------------------------ batched forms (what the drivers should call)
    def prior_draws(self, S: int, scale: Optional[torch.Tensor] = None, generator=None):
        """S prior draws at once -> (posterior_samples dict `"flow_{i}_{name}" -> [S, ...]` plus "scale" [S] — the format
        train_flows.set_params / predict and `make_engine` consume —, log_prior [S])."""
        dev = next(self.parameters()).device
        if any(name.startswith("embedding_") for name, _, _ in self._named_pairs()):
            raise NotImplementedError("batched prior draws cover the flow weights; a Bayesian embedding net is evaluated per draw upstream")
        if scale is None:
            scale = torch.rand((S,), device=dev, generator=generator) * self.scale_max
        out = {"scale": scale}
        logp = torch.full((S,), -math.log(self.scale_max), device=dev)
        for name, param, p_mle in self._named_pairs():
            mean = p_mle.data.to(dev)
            sc = scale.reshape((S,) + (1,) * mean.dim())
            x, lp = self._prior_sample(mean, sc * mean.abs() + 0.0 * mean, (S,) + tuple(mean.shape), generator)
            out[name] = x
            logp = logp + lp
        return out, logp

    def log_joint_draws(self, theta, draws, log_prior, condition=None):
        """log p(weights_s) + sum_n log p(theta_n | weights_s)  [S] (float64) — the unnormalised posterior pyro's MCMC /
        Importance evaluate one draw at a time through `model`."""
        return self.log_prob_draws(theta, draws, condition=condition, reduce="sum") + log_prior.double()
