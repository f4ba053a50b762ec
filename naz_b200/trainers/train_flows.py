"""Hot-path parts of src/naz/trainers/train_flows.py: `get_params` (:20-45), `set_params` (:47-71),
`predict` (:384-422) and the importance-weight reduction behind `train_importance` (:358-380)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import importance


def get_params(flow):
    """train_flows.py:20-45 -> one {parameter name: detached copy} dict per flow transform, in transform order."""
    return [{name: param.detach().clone().requires_grad_(param.requires_grad) for name, param in t.named_parameters()}
            for t in flow.flow_dist.transforms]


def set_params(flow, params, sample_idx=None):
    """train_flows.py:47-71.  `params` is either the list `get_params` returns, or (with `sample_idx`) the posterior-sample
    dict `"flow_{i}_{name}" -> tensor[S, ...]` from which draw `sample_idx` is copied in."""
    with torch.no_grad():
        for i, t in enumerate(flow.flow_dist.transforms):
            for name, param in t.named_parameters():
                src = params[i][name] if sample_idx is None else params[f"flow_{i}_{name}"][sample_idx]
                param.copy_(src)


def predict(flow, cond, posterior_samples, Nsamples, base_noise=None):
    """np.ndarray [S, Nsamples, D]: one batched launch over all posterior draws instead of the reference's
    per-draw `set_params` + `flow.sample` + host copy loop (train_flows.py:414-420)."""
    x = flow.sample_draws(posterior_samples, Nsamples if base_noise is None else base_noise, condition=cond)
    return x.cpu().detach().numpy()


def importance_weights(flow, theta_train, condition_train, draws, log_prior=None, log_q=None):
    """log w_s = log p(theta_s) + sum_n log p(x_n|theta_s) - log q(theta_s); returns
    (log_w [S] float64, log_evidence, ESS) — what pyro's Importance + posterior.ESS() yield upstream."""
    sum_n = flow.log_prob_draws(theta_train, draws, condition=condition_train, reduce="sum")
    lw, log_z, ess, _ = importance(sum_n, log_prior, log_q)
    return lw, float(log_z), float(ess)


def svi_importance(engine, base_params, masks, perms, x, condition, mu_q, sigma_q, scale, uniform, low=-1.0, high=1.0):
    """Importance-weighted evidence from the variational guide, entirely on the device:
      u_s   ~ TruncatedNormal(mu_q, sigma_q, low, high)          bflow_jax_maf.py:251-257 / priors/TruncatedNormal.py
      theta = theta_0 * (1 + scale * u_s)                        bflow_jax_maf.py:239-240  (inside the pack kernels)
      log w = log p(u_s) + sum_n log p(x_n | theta_s) - log q(u_s), p(u) = Uniform(-1, 1)^P   (train_flows.py:358-378)
    `uniform` is the caller's [S, P] U(0,1) noise.  Returns (log_w [S] float64, log_evidence, ESS, u [S, P])."""
    import math
    from ..engine import importance
    from ..stats import truncnorm_sample
    u, log_q = truncnorm_sample(uniform, mu_q, sigma_q, low, high)
    engine.pack_draw_map(base_params, u, scale, masks, perms)
    sum_n = engine.inverse(x, condition, want_lp=False, want_sum=True)["sum_n"]
    P = u.shape[1]
    log_prior = torch.full((u.shape[0],), -P * math.log(2.0), dtype=torch.float32, device=u.device)
    lw, log_z, ess, _ = importance(sum_n, log_prior, log_q.to(torch.float32))
    return lw, float(log_z), float(ess), u
