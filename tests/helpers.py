"""Shared builders for the parity tests (oracle = checker; naz_b200 = the thing checked)."""
import numpy as np
import torch

from oracle import flow_oracle as fo


def make_case(kind, D, C, hidden, L, S, seed, count_bins=8, order="quadratic", scale=0.25, dropout_p=None):
    """Synthetic weights per SURVEY §8(d): one MLE-like weight set, S perturbed draws
    theta_s = theta_0 (1 + scale u_s) (bflow_jax_maf.py:239-240), random permutation per layer."""
    rng = np.random.default_rng(seed)
    perms = np.stack([rng.permutation(D) for _ in range(L)])
    spec = fo.FlowSpec(kind, D, C, list(hidden), L, perms, count_bins=count_bins, order=order)
    p0 = fo.init_weights(spec, rng, np.float32)
    keep = None
    if dropout_p:
        draws = [[(W[None].repeat(S, 0), b[None].repeat(S, 0)) for (W, b) in layer] for layer in p0]
        keep = (rng.uniform(size=(S, L, len(hidden), max(hidden))) > dropout_p).astype(np.float32)
    else:
        draws = fo.perturb_draws(p0, S, scale, rng, np.float32)
    return spec, draws, keep, rng


def engine_for(spec, draws, keep=None, p_drop=0.0, engine="auto", inverse_mode="incremental", device="cuda:0", options=None):
    from naz_b200 import FlowEngine, FlowShape
    kind = spec.kind if spec.order == "quadratic" or spec.kind == "maf" else "nsa_linear"
    shape = FlowShape(kind, spec.D, spec.C, list(spec.hidden), spec.L, spec.count_bins, spec.bound, spec.clip)
    S = draws[0][0][0].shape[0]
    eng = FlowEngine(shape, S, device=device, engine=engine, inverse_mode=inverse_mode)
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    masks = spec.masks()
    tdraws = [[(torch.from_numpy(W), torch.from_numpy(b)) for (W, b) in layer] for layer in draws]
    tmasks = [[torch.from_numpy(m) for m in ml] for ml in masks]
    eng.pack(tdraws, tmasks, torch.from_numpy(spec.perms), None if keep is None else torch.from_numpy(keep), p_drop)
    return eng


def to64(draws):
    return [[(W.astype(np.float64), b.astype(np.float64)) for (W, b) in layer] for layer in draws]


def tol_report(got, ref64, rtol=1e-4, atol=1e-5):
    """Fraction of entries outside |got-ref| <= atol + rtol |ref|, and the worst error / tolerance."""
    got = np.asarray(got, np.float64)
    err = np.abs(got - ref64)
    tol = atol + rtol * np.abs(ref64)
    bad = ~(err <= tol)
    return float(bad.mean()), float(np.nanmax(err / tol)) if err.size else 0.0


def load_golden(name):
    """-> (spec, draws [L][n_lin](W[S,..], b[S,..]) fp32, arrays dict)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"), allow_pickle=False)
    kind, D, C, L, K = str(g["kind"]), int(g["D"]), int(g["C"]), int(g["L"]), int(g["K"])
    hidden = [int(h) for h in g["hidden"]]
    spec = fo.FlowSpec(kind, D, C, hidden, L, g["perms"], count_bins=K, order=str(g["order"]))
    draws = [[(g[f"W_{l}_{j}"], g[f"b_{l}_{j}"]) for j in range(len(hidden) + 1)] for l in range(L)]
    return spec, draws, g


GOLDEN = ["maf_uncond_2d", "maf_cond_3d", "nsa_cond_4d", "nsa_uncond_2d_k5", "nsa_linear_3d"]

# REFERENCE outputs: produced by executing the reference's own bflow_jax_maf.py (tools/make_reference_goldens.py)
REF_TWIN = ["ref_twin_maf_cond_3d", "ref_twin_maf_cond_6d", "ref_twin_maf_uncond_2d", "ref_twin_maf_bcast_ctx_2d",
            # bench depth ([150] x 3 hidden, 16 flow layers); weights regenerated from the generator's seed (see load_ref_twin)
            "ref_twin_deep_maf_2d", "ref_twin_deep_maf_6d", "ref_twin_deep_maf_8d_bcast"]


def load_ref_twin(name):
    """-> (spec, params [L][n_lin](W, b) fp32 single draw, arrays dict incl. the reference's masks / lp / samples)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"), allow_pickle=False)
    D, C, L = int(g["D"]), int(g["C"]), int(g["L"])
    hidden = [int(h) for h in g["hidden"]]
    spec = fo.FlowSpec("maf", D, C, hidden, L, g["perms"])
    if "weights_regenerated" in g.files:
        # replay tools/make_reference_goldens.py::main statement by statement (same Generator, same draw order)
        rng = np.random.default_rng(sum(map(ord, name)))
        perms = np.stack([rng.permutation(D) for _ in range(L)])
        assert np.array_equal(perms, g["perms"])
        dims = [D + C] + list(hidden) + [2 * D]
        params = []
        for l in range(L):
            lay = []
            for j in range(len(dims) - 1):
                W = (rng.normal(size=(dims[j + 1], dims[j])) / np.sqrt(dims[j])).astype(np.float32)
                b = (rng.normal(size=(dims[j + 1],)) * 0.1).astype(np.float32)
                lay.append((W, b))
            params.append(lay)
        chk = float(sum(np.abs(W.astype(np.float64)).sum() + np.abs(b.astype(np.float64)).sum() for lay in params for (W, b) in lay))
        assert abs(chk - float(g["weights_checksum"])) <= 1e-9 * chk, "regenerated weights differ from the generator's"
        return spec, params, g
    params = [[(g[f"W_{l}_{j}"], g[f"b_{l}_{j}"]) for j in range(len(hidden) + 1)] for l in range(L)]
    return spec, params, g


def pyro_case_to_spec(o):
    """dict written by tools/dump_pyro_goldens.py (or loaded from its .npz) -> (spec, draws [L][n_lin](W[1,..], b[1,..]) fp32)"""
    D, C, L, K = int(o["D"]), int(o["C"]), int(o["L"]), int(o["K"])
    hidden = [int(h) for h in o["hidden"]]
    spec = fo.FlowSpec("nsa", D, C, hidden, L, np.asarray(o["perms"]), count_bins=K, order=str(o["order"]))
    draws = [[(np.asarray(o[f"W_{l}_{j}"], np.float32)[None], np.asarray(o[f"b_{l}_{j}"], np.float32)[None])
              for j in range(len(hidden) + 1)] for l in range(L)]
    return spec, draws


def pyro_golden_files():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyro_nsa_*.npz")))


def explicit_coupling_flow(flow, order, dtype=torch.float64):
    """Module-structured restatement of a product 'nsc' flow: explicit SplineCoupling (+ Permute) transforms on torch's
    TransformedDistribution, sharing the product's parameter values.  `build.leaves` lists the restatement's parameters in
    the order of `flow._flat_params()` (for gradient comparisons)."""
    from functools import partial
    from torch.distributions import Normal, TransformedDistribution
    from oracle import pyro_style as ps
    D, C = flow.theta_dim, flow.shape.C
    nets, leaves = [], []
    for t in flow.transforms:
        if t.kind == "nsc":
            net = ps.ConditionalDenseNN(t.split_dim, C, t.nn.hidden_dims, t.nn.param_dims).to(dtype)
            with torch.no_grad():
                for a, b in zip(net.layers, t.nn.layers):
                    a.weight.copy_(b.weight.detach().cpu()); a.bias.copy_(b.bias.detach().cpu())
            lower = [g.detach().cpu().to(dtype).clone().requires_grad_(True) for g in t.lower_spline.groups(order)]
            nets.append((t, net, lower))
            leaves += [q for lin in net.layers for q in (lin.weight, lin.bias)] + lower
        elif t.kind == "permute":
            nets.append((t, None, None))

    def build(ctx):
        out = []
        for t, net, lower in nets:
            if net is None:
                out.append(ps.Permute(t.permutation.cpu()))
            else:
                fn = partial(net, context=ctx) if C > 0 else net
                out.append(ps.SplineCoupling(D, t.split_dim, fn, lower, t.count_bins, t.bound, order))
        return TransformedDistribution(Normal(torch.zeros(D, dtype=dtype), torch.ones(D, dtype=dtype)), out)
    build.leaves = leaves
    return build
