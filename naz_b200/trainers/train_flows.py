"""Hot-path parts of src/naz/trainers/train_flows.py: `get_params` (:20-45), `set_params` (:47-71),
`predict` (:384-422) and the importance-weight reduction behind `train_importance` (:358-380)."""
from __future__ import annotations

import copy

import numpy as np
import torch

from ..engine import importance


def get_params(flow):
    params = []
    for t in flow.flow_dist.transforms:
        this_params = {}
        for name, param in t.named_parameters():
            this_params[name] = copy.deepcopy(param)
        params.append(this_params)
    return params


def set_params(flow, params, sample_idx=None):
    for i, t in enumerate(flow.flow_dist.transforms):
        for name, param in t.named_parameters():
            with torch.no_grad():
                if sample_idx is None:
                    param.copy_(params[i][name])
                else:
                    param.copy_(params[f"flow_{i}_{name}"][sample_idx])


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
