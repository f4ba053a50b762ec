"""Hot-path parts of src/naz/trainers/train_flows.py: `get_params` (:20-45), `set_params` (:47-71), the MLE driver `train`
(:73-242) around the differentiable `log_prob`, `predict` (:384-422) and the importance-weight reduction behind
`train_importance` (:358-380)."""
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


def _loader(n: int, batch: int):
    order = torch.randperm(n)
    return [order[i:i + batch] for i in range(0, n, batch)]


def train(flow, x, y, opt=torch.optim.Adam, lr=0.001, num_epochs=1024, train_frac=0.7, batch_frac=0.005, lambda_l1=0.,
          lambda_l2=0., patience=32, min_epochs=128, clip_val=1.0, lr_decay=0.5, min_lr=None, return_final=False, verbose=True):
    """Maximum-likelihood training, train_flows.py:73-242: same arguments, same stopping rule and the same return tuple
    `(flow, history, history_val, best_mse, best_epoch)`.  Every `flow.log_prob(...)` is the CUDA path: the value from the
    forward engine, the gradient from ONE nazb_inverse_vjp launch per backward (`flows/flow.py::_LogProbFn`) instead of
    autograd through D conditioner passes per layer.  Minibatch loss = -mean log p(x | y) (+ lambda_l1 |W|_1; lambda_l2 is the
    optimiser's weight decay), gradient-norm clipping at `clip_val`, ReduceLROnPlateau(factor=lr_decay, patience=patience/2)
    on the validation loss, best-validation weights restored unless `return_final`; training ends once, past `min_epochs`,
    more than `patience` epochs brought no improvement and the learning rate has fallen below `min_lr` (default 1e-3 lr).
    The train / validation split is a random permutation (upstream: sklearn's shuffled `train_test_split`)."""
    params = [p for t in flow.flow_dist.transforms for p in t.parameters()]
    optimizer = opt(params, lr=lr, weight_decay=lambda_l2)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=lr_decay, patience=int(patience / 2))
    flow.to(x.device)
    split = torch.randperm(len(x), device=x.device)
    n_train = int(round(train_frac * len(x)))
    x_train, y_train = x[split[:n_train]], y[split[:n_train]]
    x_val, y_val = x[split[n_train:]], y[split[n_train:]]
    batch = max(1, int(n_train * batch_frac))
    floor_lr = lr * 1e-3 if min_lr is None else min_lr
    history, history_val = [], []
    best_mse, best_epoch, best_weights, stale = float("inf"), 0, None, 0
    for epoch in range(num_epochs):
        flow.train()
        running, batches = 0.0, _loader(n_train, batch)
        for idx in batches:
            idx = idx.to(x.device)
            optimizer.zero_grad()
            loss = -flow.log_prob(x_train[idx], condition=y_train[idx]).mean()
            if lambda_l1 > 0.:
                loss = loss + lambda_l1 * sum(p.abs().sum() for name, p in flow.named_parameters() if name.endswith("weight"))
            loss.backward()
            if clip_val is not None:
                torch.nn.utils.clip_grad_norm_(flow.parameters(), clip_val)
            optimizer.step()
            running += float(loss.detach())
        flow.flow_dist.clear_cache()
        flow.eval()
        with torch.no_grad():
            mse = float(-flow.log_prob(x_val, condition=y_val).mean())
        current_lr = optimizer.param_groups[0]["lr"]
        scheduler.step(mse)
        history.append(running / len(batches))
        history_val.append(mse)
        if verbose:
            print(f"epoch: {epoch}, validation_loss: {mse}, best validation_loss:{best_mse}, training_loss: {running}, "
                  f"learning_rate: {current_lr}, min_lr: {floor_lr}, no improvement for {stale}")
        if mse < best_mse:
            best_mse, best_epoch, best_weights, stale = mse, epoch, get_params(flow), 0
        elif epoch > min_epochs:
            stale += 1
        if epoch > min_epochs and stale > patience and current_lr < floor_lr:
            if verbose:
                print(f"network converged after {epoch} epochs")
            break
    if not return_final and best_weights is not None:
        set_params(flow, best_weights)
    return flow, history, history_val, best_mse, best_epoch


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
