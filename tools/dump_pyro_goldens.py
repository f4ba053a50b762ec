"""Pin the spline branch of the oracle against REAL pyro-ppl.

pyro-ppl is not installable in the build container (no network), so the rational-spline arithmetic of
oracle/flow_oracle.py is a restatement of pyro's published `_monotonic_rational_spline` /
`SplineAutoregressive`.  Run this script on any machine that has pyro-ppl:

    python tools/dump_pyro_goldens.py            # writes tests/golden/pyro_nsa_*.npz

It builds the transforms exactly as the reference's factory does (src/naz/flows/transforms.py:165-198:
ConditionalAutoRegressiveNN(theta_dim, condition_dim, hidden_dims, nonlinearity=Tanh(), param_dims=[K, K, K-1(, K)])
+ T.ConditionalSplineAutoregressive(theta_dim, arn, count_bins=K, order=...)), wraps them the way
src/naz/flows/flow.py:37-79 does (Normal(0,1) base, ConditionalTransformedDistribution), evaluates
log_prob / sample on seeded inputs and stores weights, masks, permutations, inputs and pyro's outputs.
tests/test_oracle_cpu.py::test_oracle_matches_pyro_goldens (CPU) and
tests/test_gpu_parity.py::test_cuda_matches_real_pyro (GPU, importorskip) consume them; when the files exist the
oracle header may drop the words "parity unpinned" for the spline branch.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, D, C, hidden, L, K, order, N
    ("pyro_nsa_cond_4d", 4, 2, [150, 150, 150], 16, 8, "quadratic", 256),      # the benchmarked architecture
    ("pyro_nsa_cond_3d_small", 3, 2, [32, 48], 3, 8, "quadratic", 256),
    ("pyro_nsa_uncond_2d_k5", 2, 0, [64, 64], 5, 5, "quadratic", 256),
    ("pyro_nsa_linear_3d", 3, 2, [32, 40], 3, 8, "linear", 256),
]


def build_pyro_flow(D, C, hidden, L, K, order, seed):
    """-> (list of pyro transforms, list of arns) exactly as transforms.py:165-198 builds them (random_mask=True path)."""
    import torch
    import torch.nn as nn
    import pyro.distributions.transforms as T
    from pyro.nn import AutoRegressiveNN, ConditionalAutoRegressiveNN
    torch.manual_seed(seed)
    paramdim = [K, K, K - 1, K] if order == "linear" else [K, K, K - 1]
    transforms, nets = [], []
    for _ in range(L):
        arn = (ConditionalAutoRegressiveNN(D, C, hidden, nonlinearity=nn.Tanh(), param_dims=paramdim) if C > 0
               else AutoRegressiveNN(D, hidden, nonlinearity=nn.Tanh(), param_dims=paramdim))
        nets.append(arn)
        transforms.append(T.ConditionalSplineAutoregressive(D, arn, count_bins=K, order=order) if C > 0
                          else T.SplineAutoregressive(D, arn, count_bins=K, order=order))
    return transforms, nets


def export_case(name, D, C, hidden, L, K, order, N, seed=0):
    import torch
    import pyro.distributions as dist
    import pyro.distributions.transforms as T
    transforms, nets = build_pyro_flow(D, C, hidden, L, K, order, seed)
    # perturb the (zero-mean, small) default initialisation so bins / derivatives are not near-uniform
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for arn in nets:
            for lin in arn.layers:
                lin.weight.mul_(3.0)
                lin.bias.add_(0.3 * torch.randn(lin.bias.shape, generator=g))
    base = dist.Normal(torch.zeros(D), torch.ones(D))
    x = (torch.randn((N, D), generator=g) * 1.5)
    ctx = torch.rand((N, C), generator=g) if C > 0 else None
    zin = torch.randn((N, D), generator=g)
    with torch.no_grad():
        if C > 0:
            fd = dist.ConditionalTransformedDistribution(base, transforms).condition(ctx)
        else:
            fd = dist.TransformedDistribution(base, transforms)
        lp = fd.log_prob(x)                                   # flow.py:79 (bounds=None)
        # sample direction with supplied base noise: y = T_L(...T_1(z)) (TransformedDistribution.rsample order)
        y = zin
        ld = torch.zeros(N)
        for tr in (fd.transforms if hasattr(fd, "transforms") else transforms):
            y_new = tr(y)
            ld = ld + tr.log_abs_det_jacobian(y, y_new)
            y = y_new
    out = {"D": D, "C": C, "L": L, "K": K, "hidden": np.asarray(hidden), "order": order, "x": x.numpy(), "zin": zin.numpy(),
           "lp": lp.numpy().astype(np.float64), "ys": y.numpy().astype(np.float64), "ld": ld.numpy().astype(np.float64),
           "perms": np.stack([arn.permutation.numpy() for arn in nets])}
    if C > 0:
        out["ctx"] = ctx.numpy()
    for l, arn in enumerate(nets):
        for j, lin in enumerate(arn.layers):
            out[f"W_{l}_{j}"] = lin.weight.detach().numpy()
            out[f"b_{l}_{j}"] = lin.bias.detach().numpy()
            out[f"mask_{l}_{j}"] = lin.mask.detach().numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return out


# ---- optional layers of the factories (Permute, BatchNorm in eval mode: transforms.py:155-158) and the coupling flow (:201-236) ----
EXTRA_CASES = [
    # name, kind, D, C, hidden, L, K, order, split_dim, N
    ("pyro_extra_maf_perm_bn_3d", "maf", 3, 2, [32, 32], 3, 8, "quadratic", 0, 200),
    ("pyro_extra_nsa_perm_bn_4d", "nsa", 4, 2, [48, 48], 3, 8, "quadratic", 0, 200),
    ("pyro_extra_nsc_uncond_5d", "nsc", 5, 0, [48, 48], 3, 8, "quadratic", 2, 200),
    ("pyro_extra_nsc_cond_5d_linear", "nsc", 5, 2, [48, 48], 3, 8, "linear", 2, 200),
]


def export_extra_case(name, kind, D, C, hidden, L, K, order, split, N, seed=0):
    """Flows with T.Permute + T.BatchNorm (eval mode, perturbed statistics) behind every autoregressive layer, and coupling flows
    (T.SplineCoupling over a (Conditional)DenseNN, the hyper-network conditioned the way transforms.py:113-129 does it), evaluated
    by real pyro.  tests/test_oracle_cpu.py::test_restatement_matches_pyro_extras rebuilds them from oracle/pyro_style.py."""
    from functools import partial
    import torch
    import torch.nn as nn
    import pyro.distributions as dist
    import pyro.distributions.transforms as T
    from pyro.nn import AutoRegressiveNN, ConditionalAutoRegressiveNN, ConditionalDenseNN, DenseNN
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((N, D), generator=g) * 1.2
    ctx = torch.rand((N, C), generator=g) if C > 0 else None
    zin = torch.randn((N, D), generator=g)
    out = {"kind": kind, "D": D, "C": C, "L": L, "K": K, "hidden": np.asarray(hidden), "order": order, "split": split,
           "x": x.numpy(), "zin": zin.numpy()}
    if C > 0:
        out["ctx"] = ctx.numpy()
    transforms = []
    for l in range(L):
        if kind == "nsc":
            n = D - split
            pd = [n * K, n * K, n * (K - 1)] + ([n * K] if order == "linear" else [])
            net = ConditionalDenseNN(split, C, hidden, param_dims=pd, nonlinearity=nn.Tanh()) if C > 0 else \
                DenseNN(split, hidden, param_dims=pd, nonlinearity=nn.Tanh())
            with torch.no_grad():
                for lin in net.layers:
                    lin.weight.mul_(2.0)
            hyper = partial(net, context=ctx) if C > 0 else net
            tr = T.SplineCoupling(D, split, hyper, count_bins=K, order=order)
            for j, lin in enumerate(net.layers):
                out[f"W_{l}_{j}"] = lin.weight.detach().numpy()
                out[f"b_{l}_{j}"] = lin.bias.detach().numpy()
            low = tr.lower_spline
            names = ["unnormalized_widths", "unnormalized_heights", "unnormalized_derivatives"] + (["unnormalized_lambdas"] if order == "linear" else [])
            for gi, nm in enumerate(names):
                out[f"low_{l}_{gi}"] = getattr(low, nm).detach().numpy()
            transforms.append(tr)
        else:
            pdm = [1, 1] if kind == "maf" else [K, K, K - 1]
            arn = ConditionalAutoRegressiveNN(D, C, hidden, nonlinearity=nn.Tanh(), param_dims=pdm)
            with torch.no_grad():
                for lin in arn.layers:
                    lin.weight.mul_(2.0)
            tr = (T.ConditionalAffineAutoregressive(arn) if kind == "maf" else
                  T.ConditionalSplineAutoregressive(D, arn, count_bins=K, order=order)).condition(ctx)
            out[f"perm_arn_{l}"] = arn.permutation.numpy()
            for j, lin in enumerate(arn.layers):
                out[f"W_{l}_{j}"] = lin.weight.detach().numpy()
                out[f"b_{l}_{j}"] = lin.bias.detach().numpy()
            transforms.append(tr)
        perm = torch.randperm(D, generator=g)
        transforms.append(T.Permute(perm))
        out[f"perm_{l}"] = perm.numpy()
        if kind != "nsc":
            bn = T.BatchNorm(D)
            with torch.no_grad():
                bn.gamma.copy_(0.5 + torch.rand(D, generator=g)); bn.beta.copy_(0.3 * torch.randn(D, generator=g))
                bn.moving_mean.copy_(0.2 * torch.randn(D, generator=g)); bn.moving_variance.copy_(0.5 + torch.rand(D, generator=g))
            bn.eval()
            for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
                out[f"bn_{l}_{nm}"] = getattr(bn, nm).detach().numpy()
            out[f"bn_{l}_eps"] = float(bn.epsilon)
            transforms.append(bn)
    with torch.no_grad():
        fd = dist.TransformedDistribution(dist.Normal(torch.zeros(D), torch.ones(D)), transforms)
        out["lp"] = fd.log_prob(x).numpy().astype(np.float64)
        y = zin
        for tr in transforms:
            y = tr(y)
        out["ys"] = y.numpy().astype(np.float64)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return out


def main():
    try:
        import pyro  # noqa: F401
    except Exception as e:  # pragma: no cover
        print(f"pyro-ppl is not importable here ({e!r}); nothing written.  Run this script where pyro-ppl is installed.")
        return 2
    for c in CASES:
        o = export_case(*c)
        print(f"wrote {c[0]}.npz  lp[:3] = {o['lp'][:3]}")
    for c in EXTRA_CASES:
        o = export_extra_case(*c)
        print(f"wrote {c[0]}.npz  lp[:3] = {o['lp'][:3]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
