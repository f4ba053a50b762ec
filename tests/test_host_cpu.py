"""CPU tests of the host side: the C-ABI library builds, loads and exports every symbol include/nazb.h
declares (no compute without a GPU), the Python mirror of the reference interface, and the draw-sharding
logic under a world_size-2 gloo group."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    import ctypes
    from naz_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nazb.h")).read()
    declared = set(re.findall(r"\b(nazb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"nazb_desc", "nazb_handle"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().nazb_strerror(0) == b"ok"
    assert b"sm_100" in _lib.lib().nazb_strerror(-5)


def test_header_is_plain_c_and_a_c_caller_links(built_lib, tmp_path):
    """The drop-in boundary is a C ABI: include/nazb.h must compile as C99 (what cgo / JNI / ctypes-style bindings parse)
    and a C program must link against libnazb.so and reach the error paths that need no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "c_caller.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "nazb.h"\n'
        "int main(void) {\n"
        "  nazb_handle* h = (nazb_handle*)0;\n"
        "  if (strcmp(nazb_strerror(NAZB_OK), \"ok\") != 0) return 1;\n"
        "  if (nazb_create(&h, (const nazb_desc*)0) != NAZB_ERR_BAD_ARG) return 2;   /* null descriptor */\n"
        "  if (nazb_inverse_grad(h, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) != NAZB_ERR_BAD_ARG) return 3;\n"
        "  if (nazb_launch_count() != 0) return 4;\n"
        '  puts("c-abi ok");\n  return 0;\n}\n')
    exe = tmp_path / "c_caller"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(built_lib)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", f"-I{inc}", str(src)], check=True)
    subprocess.run(["gcc", "-std=c99", f"-I{inc}", str(src), "-o", str(exe), f"-L{libdir}", "-lnazb", f"-Wl,-rpath,{libdir}"],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "c-abi ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_sass_is_blackwell_native(built_lib):
    """The tensor-core engine must contain tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM) and TMA bulk copies (UBLKCP)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnem in sass, mnem
    assert "HMMA." not in sass.replace("UTCHMMA", "")      # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu(built_lib):
    from naz_b200 import FlowEngine, FlowShape
    from naz_b200.flows import NormalizingFlow
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlowEngine(FlowShape("maf", 2, 0, [16, 16], 2), 1)
    flow = NormalizingFlow("maf", None, 2, 0, [16, 16], 2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        flow.log_prob(torch.zeros(4, 2))
    with pytest.raises(RuntimeError):
        flow.nets[0](torch.zeros(4, 2))                      # conditioners have no PyTorch forward


def test_no_product_module_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "naz_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f


def test_flow_api_mirrors_reference_formats():
    from naz_b200.flows import MCDPNormalizingFlow, NormalizingFlow
    from naz_b200.flows.bflow_maf import draw_params, torch_to_jax
    from naz_b200.flows.flow import draws_from_posterior_samples
    from naz_b200.trainers import get_params, set_params
    flow = NormalizingFlow("nsa", None, 4, 2, [150, 150, 150], 3, 8)
    names = [n for n, _ in flow.flow_dist.transforms[0].named_parameters()]
    assert names[:2] == ["nn.layers.0.weight", "nn.layers.0.bias"] and len(names) == 8
    params, shapes, masks, mask_skips, perms = torch_to_jax(flow)
    assert len(params) == 3 and params[0][0][0].shape == (150, 6) and params[0][-1][0].shape == (23 * 4, 150)
    assert masks[0][1].shape == (150, 150) and mask_skips[0].shape == (92, 6) and perms[0].shape == (4,)
    assert flow.shape.flops_per_eval() * 16 // 3 == 1_910_400        # SURVEY §8(d) F1 for config 3 (L=16)
    # get/set_params round trip and the "flow_{i}_{name}" posterior-sample dict (train_flows.py:65-71)
    saved = get_params(flow)
    post = {}
    for i, t in enumerate(flow.flow_dist.transforms):
        for n, p in t.named_parameters():
            post[f"flow_{i}_{n}"] = torch.stack([p.detach() * (1 + 0.1 * s) for s in range(5)])
    set_params(flow, post, sample_idx=3)
    assert torch.allclose(flow.nets[1].layers[2].weight, saved[1]["nn.layers.2.weight"] * 1.3)
    set_params(flow, saved)
    assert torch.equal(flow.nets[1].layers[2].weight, saved[1]["nn.layers.2.weight"])
    draws = draws_from_posterior_samples(post, 3, 4)
    assert draws[2][3][0].shape == (5, 92, 150)
    # draw map theta = theta_0 (1 + scale u)   (bflow_jax_maf.py:239-240)
    P = sum(W.numel() + b.numel() for layer in params for (W, b) in layer)
    u = torch.rand(4, P) * 2 - 1
    d = draw_params(params, u, 0.25)
    assert torch.allclose(d[0][0][0][2], params[0][0][0] * (1 + 0.25 * u[2, :900].reshape(150, 6)))
    with pytest.raises(AssertionError):
        MCDPNormalizingFlow("maf", None, 2, 0, [16], 2, dropout_p=None)
    with pytest.raises(NotImplementedError):
        NormalizingFlow("cnf", None, 4, 2, [32], 2)                  # continuous flows are outside the north star
    nsc = NormalizingFlow("nsc", None, 4, 2, [32], 2, 8, 2)          # coupling flows: built
    assert [n for n, _ in nsc.flow_dist.transforms[0].named_parameters()][:3] == [
        "nn.layers.0.weight", "nn.layers.0.bias", "nn.layers.1.weight"]


def test_hidden_degrees_recovered_from_masks():
    from naz_b200.engine import hidden_degrees_from_masks
    from naz_b200.flows.made import create_mask, sample_mask_indices
    for D, C, hidden in [(2, 0, [64, 64]), (4, 2, [150] * 3), (6, 4, [150] * 3), (16, 4, [150] * 3), (3, 0, [7, 9])]:
        perm = torch.randperm(D)
        masks, _ = create_mask(D, C, hidden, perm, 2)
        degs = hidden_degrees_from_masks(masks, perm, D, C)
        assert degs is not None
        for h, dg in zip(hidden, degs):
            ref = (sample_mask_indices(D, h) - 1) if C > 0 else sample_mask_indices(D - 1, h)
            assert dg == [int(v) for v in ref.tolist()]
        bad = [m.clone() for m in masks]
        bad[1][0, -1] = 1 - bad[1][0, -1]
        assert hidden_degrees_from_masks(bad, perm, D, C) is None


def test_shard_range_is_balanced_and_covering():
    from naz_b200.parallel import shard_range
    for S, W in [(1000, 8), (256, 8), (100, 8), (7, 3), (3, 8)]:
        spans = [shard_range(S, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == S
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, S, N, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from naz_b200.parallel import all_gather_draw_sums, all_gather_lse, shard_range
    g = torch.Generator().manual_seed(0)
    lp = torch.randn(S, N, generator=g) * 4           # same on both ranks
    b, e = shard_range(S, rank, world)
    loc = lp[b:e]
    m = loc.max(dim=0).values
    s = torch.exp(loc - m).sum(dim=0)
    ppd = all_gather_lse(m, s, S)
    sums = all_gather_draw_sums(loc.double().sum(dim=1), S)
    ref = torch.logsumexp(lp, dim=0) - np.log(S)
    ok = torch.allclose(ppd, ref, atol=1e-5) and torch.allclose(sums, lp.double().sum(dim=1))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_draw_sharded_reduction_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), 7, 33, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


class _FakeGradEngine:
    """Stands in for FlowEngine.inverse_grad on CPU: value and 'gradients' that are sums over points of simple functions."""

    def inverse_grad(self, x, ctx=None, bounds=None):
        c = 0.0 if ctx is None else (ctx.sum(-1, keepdim=True) if ctx.dim() == 2 and ctx.shape[0] == x.shape[0] else ctx.sum())
        f = (x + c)
        S = 3
        sc = torch.arange(1, S + 1, dtype=torch.float32)
        gW = [[(sc[:, None, None] * (f.T @ f.pow(2))[None]).contiguous()], [(sc[:, None, None] * f.pow(3).sum().reshape(1, 1, 1)).contiguous()]]
        gb = [[(sc[:, None] * f.sum(0)[None]).contiguous()], [(sc[:, None] * f.abs().sum().reshape(1, 1)).contiguous()]]
        return {"sum_n": (sc.double() * f.double().sum()), "gW": gW, "gb": gb}


def _gloo_grad_worker(rank, world, port, N, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from naz_b200.parallel import inverse_grad_point_sharded
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, 2, generator=g)
    ctx = torch.rand(N, 3, generator=g)
    eng = _FakeGradEngine()
    got = inverse_grad_point_sharded(eng, x, ctx)
    ref = eng.inverse_grad(x, ctx)
    ok = torch.allclose(got["sum_n"], ref["sum_n"], rtol=1e-6)
    for a, b in zip([t for l in got["gW"] + got["gb"] for t in l], [t for l in ref["gW"] + ref["gb"] for t in l]):
        ok = ok and a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    # broadcast context is not sliced
    got1 = inverse_grad_point_sharded(eng, x, ctx[0])
    ok = ok and torch.allclose(got1["sum_n"], eng.inverse_grad(x, ctx[0])["sum_n"], rtol=1e-6)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_point_sharded_gradient_allreduce_world2_gloo():
    """f1 multi-GPU: points sharded, one all-reduce of the gradient buffer (ragged split: 37 points over 2 ranks)"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_grad_worker, args=(world, _free_port(), 37, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_empty_batch_returns_empty_results_without_touching_the_library():
    """`flow.log_prob(x[0:0])` / `flow.sample([0])` are legal upstream (pyro / jnp return empty arrays).  The engine answers
    them on the host: exercised here on a bare FlowEngine object (no handle, no CUDA) so any library call would fail."""
    from naz_b200.engine import FlowEngine, FlowShape
    eng = object.__new__(FlowEngine)
    eng.shape = FlowShape("maf", 3, 2, [8, 8], 2)
    eng.S = 4
    eng.device = torch.device("cpu")
    eng._h = None
    eng._lib = None
    eng._keepalive = ([], [], [torch.ones(8, 5), torch.ones(8, 8), torch.ones(6, 8)] * 2, None, None)
    x = torch.zeros((0, 3))
    ctx = torch.zeros((0, 2))
    out = eng.inverse(x, ctx, want_z=True, want_lp=True, want_lse=True, want_sum=True, n_groups=2)
    assert out["z"].shape == (4, 0, 3) and out["lp"].shape == (4, 0)
    assert out["lse_max"].shape == (2, 0) and out["lse_sum"].shape == (2, 0)
    assert out["sum_n"].shape == (4,) and out["sum_n"].dtype == torch.float64 and float(out["sum_n"].abs().sum()) == 0.0
    assert eng.lse_finish(out["lse_max"], out["lse_sum"], 0.0).shape == (0,)
    xs, ld = eng.forward(torch.zeros((0, 3)), torch.zeros((2,)), want_logdet=True)
    assert xs.shape == (4, 0, 3) and ld.shape == (4, 0)
    assert eng.forward(torch.zeros((4, 0, 3)), torch.zeros((2,))).shape == (4, 0, 3)
    g = eng.inverse_grad(x, ctx, want_dx=True, want_lp=True)
    assert g["dx"].shape == (4, 0, 3) and g["lp"].shape == (4, 0) and float(g["sum_n"].abs().sum()) == 0.0
    assert all(float(t.abs().sum()) == 0.0 for layer in g["gW"] + g["gb"] for t in layer)
    assert g["gW"][1][2].shape == (4, 6, 8)
    eng._h = None   # nothing to destroy


def test_twin_factories_unpack_like_the_reference_call_sites():
    """calibrate.py:91-92 / hmc_maf_exact.py:105 / plot_svi.py:83 replayed verbatim: a 3-tuple and a (forward, inverse) pair."""
    import torch
    from naz_b200.flows.bflow_maf import (create_mask, make_conditional_autoregressive_nn,
                                          make_masked_affine_autoregressive_transform, _perm_from_mask_skip)
    theta_dim, lambda_dim, hidden_dims = 2, 2, [150, 150, 150]
    nn, param_shape, mask_generator = make_conditional_autoregressive_nn(theta_dim, lambda_dim, hidden_dims)
    transform = make_masked_affine_autoregressive_transform(nn, theta_dim)
    assert len(transform) == 2 and callable(transform[0]) and callable(transform[1])
    # param_shapes exactly as upstream builds them (bflow_jax_maf.py:130-133; bias "shapes" are bare ints there)
    assert param_shape == [((150, 4), 150), ((150, 150), 150), ((150, 150), 150), ((4, 150), 4)]
    # train_maf (:271) initialises from them: np.random.normal(size=shape[0]) / size=shape[1] must both be valid sizes
    for ws, bs in param_shape:
        assert np.random.normal(size=ws).shape == tuple(ws) and np.random.normal(size=bs).shape == (bs,)
    masks, mask_skip, perm = mask_generator(torch.tensor([1, 0]))
    assert [tuple(m.shape) for m in masks] == [(150, 4), (150, 150), (150, 150), (4, 150)] and tuple(mask_skip.shape) == (4, 4)
    m2, ms2 = create_mask(2, 2, hidden_dims, torch.tensor([1, 0]), 2)
    assert all(torch.equal(a, b) for a, b in zip(masks, m2)) and torch.equal(mask_skip, ms2)
    for D, C in ((2, 2), (6, 4), (5, 0)):
        for _ in range(3):
            pm = torch.randperm(D)
            _, ms, _ = make_conditional_autoregressive_nn(D, C, [32, 32])[2](pm)
            assert torch.equal(_perm_from_mask_skip(ms, C), pm)
    # the conditioner is fused into the transform kernels: calling it alone, or a transform on host tensors, fails loudly
    with pytest.raises(RuntimeError):
        nn(torch.zeros(3, 2), None, masks, mask_skip)
    with pytest.raises(RuntimeError):
        transform[0]((torch.zeros(3, 2), torch.zeros(3)), (None, masks, mask_skip))
    with pytest.raises(NotImplementedError):
        make_conditional_autoregressive_nn(2, 2, [8], skip_connections=True)


def test_ravel_pytree_and_bayesian_sites_host_side():
    import torch
    from naz_b200.flows.bflow_maf import bayesian_normalizing_flow, draw_params, ravel_pytree
    torch.manual_seed(0)
    best = [[(torch.randn(5, 3), torch.randn(5)), (torch.randn(4, 5), torch.randn(4))] for _ in range(2)]
    flat, unravel = ravel_pytree(best)
    assert flat.shape == (2 * (15 + 5 + 20 + 4),)
    back = unravel(flat)
    assert all(torch.equal(a, c) and torch.equal(b, d) for la, lb in zip(best, back) for (a, b), (c, d) in zip(la, lb))
    calls = []
    model, guide, guided_model, unravel_fn, log_prob = bayesian_normalizing_flow(
        lambda p: (calls.append(p), torch.arange(7.0))[1], best, scale_max=0.25, return_log_l=True)
    sites = model()
    assert float(sites["log_l"]) == 21.0 and sites["params"].shape == flat.shape
    assert float(sites["scale"][0]) == 0.25 and float(sites["standard_params"].abs().max()) <= 1.0
    assert float(model(prior=True)["log_l"]) == 0.0
    # params = theta_MLE (1 + scale u), un-ravelled in pytree order == draw_params on the batched standard parameters
    post = model.draw(3)
    a = unravel_fn(model.params_of(post))
    b = draw_params(best, post["standard_params"], 0.25)
    assert all(torch.allclose(x1, x2) and torch.allclose(y1, y2) for la, lb in zip(a, b) for (x1, y1), (x2, y2) in zip(la, lb))
    m2, *_ = bayesian_normalizing_flow(lambda p: torch.zeros(1), best, scale_max=0.5, fixed_scale=False, multi_scale=True)
    s2 = m2.draw(4)
    assert s2["scale"].shape == (4, flat.numel()) and float(s2["scale"].max()) <= 0.5


def test_bayesian_normalizing_flow_class_priors_host_side():
    """bflow.py:30-47,57-94: sigma = scale |theta_MLE|, the bounded priors stay inside mean +- sigma, sites are named
    flow_{i}_{name}, the sampled weights are copied into the module."""
    import torch
    from naz_b200.flows import BayesianNormalizingFlow, NormalizingFlow
    torch.manual_seed(0)
    mle = NormalizingFlow("maf", None, 2, 2, [16, 16], 2)
    for kind in ("Uniform", "TruncNorm", "Normal", "StandardNormal"):
        bf = BayesianNormalizingFlow(mle, "maf", None, 2, 2, [16, 16], 2, prior_dist=kind, scale_max=0.1)
        sites = bf.prior_model()
        assert bf.n_params == sum(p.numel() for p in mle.parameters())
        assert set(sites) == {"scale"} | {f"flow_{i}_{n}" for i, t in enumerate(bf.flow_dist.transforms) for n, _ in t.named_parameters()}
        W_mle, W_new = mle.nets[0].layers[0].weight.data, bf.nets[0].layers[0].weight.data
        assert torch.equal(W_new, sites["flow_0_nn.layers.0.weight"])
        if kind in ("Uniform", "TruncNorm"):
            assert bool(((W_new - W_mle).abs() <= sites["scale"] * W_mle.abs() * (1 + 1e-5)).all())
        draws, logp = bf.prior_draws(4)
        assert draws["flow_1_nn.layers.2.bias"].shape == (4, 4) and logp.shape == (4,) and bool(torch.isfinite(logp).all())
    tr = bf.param_transforms()
    fwd, inv = tr["flow_0_nn.layers.0.weight"]
    a, b = bf.param_bounds["flow_0_nn.layers.0.weight"]
    u = torch.randn_like(a)
    x = fwd(u)
    assert bool(((x >= torch.minimum(a, b)) & (x <= torch.maximum(a, b))).all()) and torch.allclose(inv(x), u, atol=1e-3)
    with pytest.raises(ValueError):
        BayesianNormalizingFlow(mle, "maf", None, 2, 2, [16, 16], 2, prior_dist="Cauchy")


@pytest.mark.parametrize("flow_type", ["maf", "nsa"])
def test_permute_and_batchnorm_layers_fold_into_the_packed_flow(flow_type):
    """random_perm / use_batchnorm of the reference's factories (transforms.py:155-158, :193-196).  Three independent
    statements must agree: (1) the module-structured restatement with EXPLICIT Permute / BatchNorm transforms on torch's
    TransformedDistribution; (2) the numpy oracle evaluated on what the product hands to libnazb — conditioner weights with
    the permutations folded in (`Relabelling`), re-labelled inputs / bounds, a per-layer affine; (3) the folded MADE masks
    are exactly the masks the folded MADE orders generate."""
    from naz_b200.flows.flow import NormalizingFlow
    from naz_b200.flows.transforms import BatchNorm, Permute
    from oracle import flow_oracle as fo, pyro_style as ps
    torch.manual_seed(3)
    D, C, hidden, L, K = 4, 2, [24, 24], 3, 6
    bounds = {"low": torch.tensor([-1.0, 0.0, -2.0, 0.5]), "high": torch.tensor([1.0, 3.0, 2.0, 4.5])}
    args = (D, C, hidden, L) + ((K,) if flow_type == "nsa" else ())
    flow = NormalizingFlow(flow_type, bounds, *args, random_perm=True, use_batchnorm=True)
    assert [type(t).__name__ for t in flow.transforms][1:3] == ["Permute", "BatchNorm"] and len(flow.transforms) == 3 * L
    assert not flow.relabel.trivial and flow.relabel.has_bn and flow.relabel.ar_pos == [0, 3, 6]
    with torch.no_grad():
        for t in flow.transforms:
            if isinstance(t, BatchNorm):
                t.gamma.copy_(0.5 + torch.rand(D)); t.beta.copy_(0.3 * torch.randn(D))
                t.moving_mean.copy_(0.2 * torch.randn(D)); t.moving_variance.copy_(0.5 + torch.rand(D))
    # (1) explicit layers
    torch.set_default_dtype(torch.float64)
    try:
        extras = []
        for l in range(L):
            pm, bn = flow.transforms[3 * l + 1], flow.transforms[3 * l + 2]
            extras.append([ps.Permute(pm.permutation), ps.BatchNormEval(bn.gamma.detach().double(), bn.beta.detach().double(),
                                                                        bn.moving_mean.double(), bn.moving_variance.double(), bn.epsilon)])
        b64 = {k: v.double() for k, v in bounds.items()}
        ref = ps.PyroStyleFlow(flow_type, b64, D, C, hidden, L, K, "quadratic", permutations=flow.perms().numpy(), extras=extras)
        ref.set_from_pytree([[(W.double().numpy(), b.double().numpy()) for (W, b) in layer] for layer in flow.current_draw()])
        x = torch.rand(64, D, dtype=torch.float64) * (b64["high"] - b64["low"]) * 0.96 + b64["low"] + 0.02 * (b64["high"] - b64["low"])
        ctx = torch.randn(64, C, dtype=torch.float64)
        with torch.no_grad():
            lp_ref = ref.log_prob(x, ctx).numpy()
            z = torch.randn(64, D, dtype=torch.float64)
            xs_ref = ref.sample(None, ctx, base_noise=z).numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    # (2) what the product packs
    perms_e = flow._packed_perms().numpy()
    spec = fo.FlowSpec(flow_type, D, C, hidden, L, perms_e, count_bins=K)
    params_e = [[(W.double().numpy(), b.double().numpy()) for (W, b) in layer] for layer in flow._fold_draws(flow.current_draw())]
    aff = tuple(t.double().numpy() for t in flow.relabel.layer_affine())
    be = flow._bounds_e()
    xe = flow.relabel.to_engine(x).numpy()
    _, lp_e = fo.flow_inverse(spec, params_e, xe, ctx.numpy(), (be["low"].double().numpy(), be["high"].double().numpy()), layer_affine=aff)
    # the product's affine table is fp32 (what libnazb takes): agreement to fp32 round-off of (a, b); a folding mistake is O(1)
    np.testing.assert_allclose(lp_e, lp_ref, rtol=2e-6, atol=2e-5)
    xs_e, _ = fo.flow_forward(spec, params_e, z.numpy(), ctx.numpy(), (be["low"].double().numpy(), be["high"].double().numpy()), layer_affine=aff)
    np.testing.assert_allclose(flow.relabel.from_engine(torch.as_tensor(xs_e)).numpy(), xs_ref, rtol=2e-6, atol=2e-5)
    # (3) folded masks == masks of the folded orders
    for ml_e, ml in zip(flow._packed_masks(), spec.masks()):
        for me, m in zip(ml_e, ml):
            assert np.array_equal(me.numpy(), m)
    # posterior-sample dicts index EVERY transform (train_flows.py:71): flow_0, flow_3, flow_6 are the conditioners
    from naz_b200.trainers.train_flows import get_params
    gp = get_params(flow)
    assert len(gp) == 3 * L and set(gp[1]) == set() and set(gp[2]) == {"gamma", "beta"}


@pytest.mark.parametrize("order,random_perm,C", [("quadratic", False, 2), ("quadratic", True, 2), ("linear", True, 0)])
def test_coupling_flow_is_a_single_degree_masked_conditioner(order, random_perm, C):
    """'nsc' (transforms.py:201-236, built to its evident intent).  The product hands a coupling layer to libnazb as a masked
    conditioner with one hidden degree (`SplineCoupling.as_made`).  Checked here without a GPU: the numpy oracle evaluated on
    exactly those weights / masks / orders (+ the Permute re-labelling) equals the explicit-transform restatement, in both
    directions; the masks are canonical MADE masks for the engine's incremental inverse with all hidden degrees = split_dim."""
    from naz_b200.engine import hidden_degrees_from_masks
    from naz_b200.flows.flow import NormalizingFlow
    from oracle import flow_oracle as fo
    torch.manual_seed(7)
    D, s, hidden, L, K = 5, 2, [20, 20], 3, 6
    flow = NormalizingFlow("nsc", None, D, C, hidden, L, K, s, order=order, random_perm=random_perm)
    assert flow.shape.M == (4 * K - 1 if order == "linear" else 3 * K - 1)
    from helpers import explicit_coupling_flow
    build = explicit_coupling_flow(flow, order)
    N = 50
    x = torch.randn(N, D, dtype=torch.float64) * 1.5
    ctx = torch.randn(N, C, dtype=torch.float64) if C else None
    z = torch.randn(N, D, dtype=torch.float64)
    with torch.no_grad():
        pdf = build(ctx)
        lp_ref = pdf.log_prob(x).numpy()
        xs = z
        for t in pdf.transforms:
            xs = t(xs)
        xs_ref = xs.numpy()
    masks_e = [[m.numpy() for m in ml] for ml in flow._packed_masks()]
    perms_e = flow._packed_perms().numpy()
    kind = "nsa"
    spec = fo.FlowSpec(kind, D, C, hidden, L, perms_e, count_bins=K, order=order, masks_override=masks_e)
    params_e = [[(W.double().numpy(), b.double().numpy()) for (W, b) in layer] for layer in flow._fold_draws(flow.current_draw())]
    cn = None if ctx is None else ctx.numpy()
    _, lp_e = fo.flow_inverse(spec, params_e, flow.relabel.to_engine(x).numpy(), cn)
    np.testing.assert_allclose(lp_e, lp_ref, rtol=1e-5, atol=1e-5)          # parameters are fp32 on both sides; fp64 arithmetic
    xs_e, _ = fo.flow_forward(spec, params_e, z.numpy(), cn)
    np.testing.assert_allclose(flow.relabel.from_engine(torch.as_tensor(xs_e)).numpy(), xs_ref, rtol=1e-5, atol=1e-5)
    for l in range(L):
        degs = hidden_degrees_from_masks([torch.as_tensor(m) for m in masks_e[l]], torch.as_tensor(perms_e[l]), D, C)
        assert degs is not None and all(set(d) == {s if C else s} for d in degs), degs


def test_gradient_unmapping_is_the_exact_transpose_of_the_packing_maps():
    """Gradients come back from libnazb in the engine's conditioner format.  Two host-side index maps take them to the
    reference's parameters: `Relabelling.unfold_layer_grads` (Permute layers) and `SplineCoupling.grads_from_made` (coupling
    layers).  Each must be the exact transpose of the map that packed the weights:  <pack(theta), G> == <theta, unpack(G)>."""
    from naz_b200.flows.flow import NormalizingFlow
    torch.manual_seed(5)
    # Permute folding
    D, C, hidden, L, K = 5, 3, [12, 12], 3, 4
    flow = NormalizingFlow("nsa", None, D, C, hidden, L, K, random_perm=True)
    M = flow.shape.M
    for l in range(L):
        W0, Wl, bl = torch.randn(hidden[0], C + D), torch.randn(M * D, hidden[-1]), torch.randn(M * D)
        packed = flow.relabel.fold_layer(l, [(W0, torch.zeros(hidden[0])), (Wl, bl)], C)
        G0, Gl, gl = torch.randn_like(W0), torch.randn_like(Wl), torch.randn_like(bl)
        u0, ul, ub = flow.relabel.unfold_layer_grads(l, G0, Gl, gl, C)
        lhs = (packed[0][0] * G0).sum() + (packed[1][0] * Gl).sum() + (packed[1][1] * gl).sum()
        rhs = (W0 * u0).sum() + (Wl * ul).sum() + (bl * ub).sum()
        assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))
    # coupling layers, both spline orders
    for order in ("quadratic", "linear"):
        flow = NormalizingFlow("nsc", None, D, C, hidden, 2, K, 2, order=order)
        t = flow._layers[0]
        lins = [(lin.weight.detach(), lin.bias.detach()) for lin in t.nn.layers]
        lower = [g.detach() for g in t.lower_spline.groups(order)]
        made = t.as_made(lins, lower, C)
        gW = [torch.randn_like(W) * m for (W, _), m in zip(made, t.made_masks(C))]      # the kernel never touches masked entries
        gb = [torch.randn_like(b) for (_, b) in made]
        back = t.grads_from_made(gW, gb, C)
        theta = [q for (W, b) in lins for q in (W, b)] + lower
        assert len(back) == len(theta) and all(a.shape == b.shape for a, b in zip(back, theta))
        lhs = sum(float((W * g).sum()) for (W, _), g in zip(made, gW)) + sum(float((b * g).sum()) for (_, b), g in zip(made, gb))
        rhs = sum(float((a * b).sum()) for a, b in zip(theta, back))
        assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs)), (order, lhs, rhs)
