"""value and gradient of the summed log-likelihood (masked-affine and quadratic neural-spline flows).
CPU: the gradient oracle (torch autograd of the restated twin) is pinned on the forward side by the reference-executed
fixtures and checked against finite differences.  GPU: nazb_inverse_grad against the oracle."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from oracle import grad_oracle as go
from helpers import make_case, load_ref_twin, REF_TWIN, to64


def _single(draws, s):
    return [[(W[s], b[s]) for (W, b) in layer] for layer in draws]


# ----------------------------------------------------------------------------------------------------------------
# CPU
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", REF_TWIN)
def test_grad_oracle_forward_matches_reference_outputs(name):
    """the differentiated function IS the reference's log_prob: its values equal the reference-executed fixtures"""
    spec, params, g = load_ref_twin(name)
    masks = spec.masks() if "weights_regenerated" in g.files else [[g[f"mask_{l}_{j}"] for j in range(len(spec.hidden) + 1)] for l in range(spec.L)]
    ctx = g["ctx"] if spec.C else None
    _, _, _, _, lp = go.value_and_grad(to64([[(W, b) for (W, b) in layer] for layer in params]), masks, spec.perms, g["x"], ctx)
    np.testing.assert_allclose(lp, g["lp"], rtol=1e-9, atol=1e-9)


REF_GRAD = ["ref_twin_maf_cond_3d", "ref_twin_maf_cond_6d", "ref_twin_maf_bcast_ctx_2d"]


def load_ref_grad(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name.replace("ref_twin_", "ref_twin_grad_") + ".npz"))


@pytest.mark.parametrize("name", REF_GRAD)
def test_grad_oracle_matches_autodiff_of_the_reference_code(name):
    """tests/golden/ref_twin_grad_*.npz = reverse-mode derivatives of the REFERENCE's own log_prob (its bytes executed with
    torch standing in for jax.numpy, tools/make_reference_goldens.py::grad_fixture): the gradient oracle must reproduce them"""
    spec, params, g = load_ref_twin(name)
    gr = load_ref_grad(name)
    masks = [[g[f"mask_{l}_{j}"] for j in range(len(spec.hidden) + 1)] for l in range(spec.L)]
    ctx = g["ctx"] if spec.C else None
    val, gW, gb, dx, _ = go.value_and_grad(to64(params), masks, spec.perms, g["x"], ctx, want_dx=True)
    assert abs(val - float(gr["sum_lp"])) < 1e-9 * max(1.0, abs(val))
    for l in range(spec.L):
        for j in range(len(spec.hidden) + 1):
            np.testing.assert_allclose(gW[l][j], gr[f"gW_{l}_{j}"], rtol=1e-8, atol=1e-9)
            np.testing.assert_allclose(gb[l][j], gr[f"gb_{l}_{j}"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dx, gr["dx"], rtol=1e-8, atol=1e-9)


def test_grad_oracle_matches_flow_oracle_and_finite_differences():
    spec, draws, _, rng = make_case("maf", 3, 2, [12, 12], 2, 1, seed=5)
    p = to64(_single(draws, 0))
    masks = spec.masks()
    x = rng.normal(size=(7, 3))
    ctx = rng.uniform(size=(7, 2))
    bounds = (np.full(3, -6.0), np.full(3, 6.0))
    val, gW, gb, dx, lp = go.value_and_grad(p, masks, spec.perms, x, ctx, bounds, want_dx=True)
    ref = fo.log_prob_draws(spec, [[(W[None], b[None]) for (W, b) in layer] for layer in p], x, ctx, bounds)[0][0]
    np.testing.assert_allclose(lp, ref, rtol=1e-10, atol=1e-10)

    def f(pp, xx=x):
        return fo.log_prob_draws(spec, [[(W[None], b[None]) for (W, b) in layer] for layer in pp], xx, ctx, bounds)[0][0].sum()

    eps = 1e-6
    for (l, j, idx) in [(0, 0, (3, 2)), (1, 1, (5, 4)), (0, 2, (1, 7)), (1, 2, (4, 0))]:
        if masks[l][j][idx] == 0:
            assert gW[l][j][idx] == 0.0
            continue
        pp = [[(W.copy(), b.copy()) for (W, b) in layer] for layer in p]
        pp[l][j][0][idx] += eps
        up = f(pp)
        pp[l][j][0][idx] -= 2 * eps
        dn = f(pp)
        assert abs((up - dn) / (2 * eps) - gW[l][j][idx]) < 1e-5 * max(1.0, abs(gW[l][j][idx]))
    pp = [[(W.copy(), b.copy()) for (W, b) in layer] for layer in p]
    pp[1][0][1][3] += eps
    up = f(pp)
    pp[1][0][1][3] -= 2 * eps
    dn = f(pp)
    assert abs((up - dn) / (2 * eps) - gb[1][0][3]) < 1e-5 * max(1.0, abs(gb[1][0][3]))
    xp = x.copy(); xp[2, 1] += eps
    xm = x.copy(); xm[2, 1] -= eps
    assert abs((f(p, xp) - f(p, xm)) / (2 * eps) - dx[2, 1]) < 1e-5 * max(1.0, abs(dx[2, 1]))


# ----------------------------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


GRAD_CASES = [
    # kind D C hidden L S N bounds ctx_rows
    (2, 0, [16, 16], 3, 2, 37, False, 0),
    (3, 2, [24, 24], 4, 3, 100, True, "N"),
    (6, 4, [150, 150, 150], 2, 2, 70, False, "N"),
    (4, 2, [150, 150, 150], 3, 1, 33, False, 1),
    (3, 14, [12, 12], 2, 2, 50, False, "N"),      # C + D = 17 > every hidden width: the zero-bias buffer of the transposed products
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GRAD_CASES)
def test_inverse_grad_matches_autograd_oracle(case):
    from helpers import engine_for
    D, C, hidden, L, S, N, bounded, crow = case
    spec, draws, _, rng = make_case("maf", D, C, hidden, L, S, seed=11)
    x = (rng.normal(size=(N, D)) * 1.2).astype(np.float32)
    ctx = None
    if C:
        ctx = rng.uniform(size=(N if crow == "N" else 1, C)).astype(np.float32)
    bounds = (np.full(D, -7.0, np.float32), np.full(D, 7.0, np.float32)) if bounded else None
    eng = engine_for(spec, draws, engine="simt")
    r = eng.inverse_grad(torch.from_numpy(x), None if ctx is None else torch.from_numpy(ctx),
                         None if bounds is None else (torch.from_numpy(bounds[0]), torch.from_numpy(bounds[1])),
                         want_dx=True, want_lp=True)
    torch.cuda.synchronize()
    masks = spec.masks()
    for s in range(S):
        p = to64(_single(draws, s))
        c64 = None if ctx is None else (ctx[0] if ctx.shape[0] == 1 else ctx).astype(np.float64)
        b64 = None if bounds is None else tuple(b.astype(np.float64) for b in bounds)
        val, gW, gb, dx, lp = go.value_and_grad(p, masks, spec.perms, x.astype(np.float64), c64, b64, want_dx=True)
        np.testing.assert_allclose(r["lp"][s].cpu().numpy(), lp, rtol=1e-4, atol=1e-4)
        assert abs(float(r["sum_n"][s]) - val) <= 1e-4 * max(1.0, abs(val))
        # tolerance: relative to the largest entry of each gradient array (fp32 products and fp32 atomic sums over N)
        for l in range(L):
            for j in range(len(hidden) + 1):
                assert _rel(r["gW"][l][j][s].cpu().numpy(), gW[l][j]) < 2e-4, (s, l, j, "W")
                assert _rel(r["gb"][l][j][s].cpu().numpy(), gb[l][j]) < 2e-4, (s, l, j, "b")
                # masked entries are never touched
                assert np.all(r["gW"][l][j][s].cpu().numpy()[masks[l][j] == 0] == 0.0)
        assert _rel(r["dx"][s].cpu().numpy(), dx) < 2e-4


@pytest.mark.gpu
def test_inverse_grad_draw_range_and_unsupported():
    from helpers import engine_for
    from naz_b200 import _lib
    spec, draws, _, rng = make_case("maf", 3, 1, [16, 16], 2, 4, seed=3)
    x = torch.from_numpy(rng.normal(size=(50, 3)).astype(np.float32))
    ctx = torch.from_numpy(rng.uniform(size=(1, 1)).astype(np.float32))
    eng = engine_for(spec, draws, engine="simt")
    full = eng.inverse_grad(x, ctx)
    part = eng.inverse_grad(x, ctx, s_begin=1, s_count=2)
    torch.cuda.synchronize()
    assert torch.allclose(part["sum_n"], full["sum_n"][1:3], rtol=1e-6)
    g_f, g_p = full["gW"][0][1], part["gW"][0][1]
    assert torch.all(g_p[0] == 0) and torch.all(g_p[3] == 0)
    assert _rel(g_p[1:3].cpu().numpy(), g_f[1:3].cpu().numpy()) < 1e-4
    # the value agrees with the log_prob entry point
    lp = eng.inverse(x, ctx, want_lp=True)["lp"].double().sum(-1)
    assert torch.allclose(lp, full["sum_n"], rtol=1e-5)
    # tensor-core handles are not served: loud error, no fallback
    spec2, draws2, _, _ = make_case("maf", 2, 1, [32, 32], 2, 2, seed=3)
    eng2 = engine_for(spec2, draws2, engine="tcgen05")
    x = x[:, :2].contiguous()
    with pytest.raises(_lib.NazbError):
        eng2.inverse_grad(x, ctx)


@pytest.mark.gpu
def test_twin_value_and_grad_single_draw_against_reference_fixture():
    """make_normalizing_flow(...)["value_and_grad"] on a reference-executed fixture: value = sum of the reference's own lp"""
    from naz_b200.flows import bflow_maf as bm
    spec, params, g = load_ref_twin("ref_twin_maf_cond_3d")
    masks = [[torch.from_numpy(g[f"mask_{l}_{j}"]) for j in range(len(spec.hidden) + 1)] for l in range(spec.L)]
    tparams = [[(torch.from_numpy(W), torch.from_numpy(b)) for (W, b) in layer] for layer in params]
    nn, _, _ = bm.make_conditional_autoregressive_nn(spec.D, spec.C, spec.hidden)
    transform = bm.make_masked_affine_autoregressive_transform(nn, spec.D)
    flow = bm.make_normalizing_flow(transform, torch.from_numpy(g["x"]), masks, None, [torch.from_numpy(p) for p in spec.perms],
                                    context=torch.from_numpy(g["ctx"]))
    val, grads = flow["value_and_grad"](tparams)
    torch.cuda.synchronize()
    assert abs(float(val) - float(g["lp"].astype(np.float64).sum())) < 1e-4 * max(1.0, abs(float(g["lp"].sum())))
    m64 = [[g[f"mask_{l}_{j}"] for j in range(len(spec.hidden) + 1)] for l in range(spec.L)]
    _, gW, gb, _, _ = go.value_and_grad(to64(params), m64, spec.perms, g["x"], g["ctx"])
    for l in range(spec.L):
        for j in range(len(spec.hidden) + 1):
            assert _rel(grads[l][j][0].cpu().numpy(), gW[l][j]) < 2e-4
            assert _rel(grads[l][j][1].cpu().numpy(), gb[l][j]) < 2e-4


@pytest.mark.gpu
def test_inverse_grad_directional_derivative_at_bench_size():
    """Size-independent property at the size bench.py times (config-4 architecture, 100 000 points): the central
    difference of the VALUE — computed by the tensor-core log_prob engine, an independent code path — along the
    normalised gradient equals the gradient's norm."""
    from helpers import engine_for
    S, N = 2, 100_000
    spec, draws, _, rng = make_case("maf", 2, 2, [150] * 3, 16, S, seed=21)
    x = torch.from_numpy((rng.normal(size=(N, 2)) * 1.5).astype(np.float32))
    ctx = torch.from_numpy(rng.uniform(size=(N, 2)).astype(np.float32))
    r = engine_for(spec, draws, engine="simt").inverse_grad(x, ctx)
    torch.cuda.synchronize()
    gW = [[t.cpu().numpy().astype(np.float64) for t in layer] for layer in r["gW"]]
    gb = [[t.cpu().numpy().astype(np.float64) for t in layer] for layer in r["gb"]]
    norm = np.sqrt(sum((g ** 2).reshape(S, -1).sum(1) for layer in gW for g in layer) +
                   sum((g ** 2).reshape(S, -1).sum(1) for layer in gb for g in layer))          # [S]
    assert np.all(np.isfinite(norm)) and np.all(norm > 0)
    eps = np.clip(20.0 / norm, 1e-5, 1e-3)                                                      # value step ~ +-20

    def value(sign):
        pert = [[((W + sign * (eps / norm)[:, None, None] * gW[l][j]).astype(np.float32),
                  (b + sign * (eps / norm)[:, None] * gb[l][j]).astype(np.float32))
                 for j, (W, b) in enumerate(layer)] for l, layer in enumerate(draws)]
        e = engine_for(spec, pert, engine="auto")
        v = e.inverse(x, ctx, want_lp=False, want_sum=True)["sum_n"].cpu().numpy()
        return v, e.engine_for("inverse")

    vp, used = value(+1.0)
    vm, _ = value(-1.0)
    assert used == "tcgen05"
    fd = (vp - vm) / (2 * eps)
    assert np.all(np.abs(fd - norm) <= 3e-2 * norm), (fd, norm, eps)


@pytest.mark.gpu
@pytest.mark.parametrize("name", REF_GRAD)
def test_cuda_gradient_matches_autodiff_of_the_reference_code(name):
    """nazb_inverse_grad against the derivatives of the reference's own code (ref_twin_grad_*.npz)"""
    from helpers import engine_for
    spec, params, g = load_ref_twin(name)
    gr = load_ref_grad(name)
    draws = [[(W[None], b[None]) for (W, b) in layer] for layer in params]
    eng = engine_for(spec, draws, engine="simt")
    ctx = torch.from_numpy(g["ctx"]) if spec.C else None
    r = eng.inverse_grad(torch.from_numpy(g["x"]), ctx, want_dx=True)
    torch.cuda.synchronize()
    assert abs(float(r["sum_n"][0]) - float(gr["sum_lp"])) < 1e-4 * max(1.0, abs(float(gr["sum_lp"])))
    for l in range(spec.L):
        for j in range(len(spec.hidden) + 1):
            assert _rel(r["gW"][l][j][0].cpu().numpy(), gr[f"gW_{l}_{j}"]) < 2e-4, (l, j, "W")
            assert _rel(r["gb"][l][j][0].cpu().numpy(), gr[f"gb_{l}_{j}"]) < 2e-4, (l, j, "b")
    assert _rel(r["dx"][0].cpu().numpy(), gr["dx"]) < 2e-4


def test_flow_level_grad_oracle_maf_equals_twin_and_spline_matches_finite_differences():
    """value_and_grad_flow (autograd through the pyro-style restatement): equals the pinned twin oracle for maf; for the
    spline flow (unpinned branch, groundwork for the spline backward kernel) it matches central finite differences of the
    numpy forward oracle."""
    spec, draws, _, rng = make_case("maf", 3, 2, [12, 12], 2, 1, seed=9)
    p = to64(_single(draws, 0))
    x = rng.normal(size=(9, 3))
    ctx = rng.uniform(size=(9, 2))
    v1, gW1, gb1, _, _ = go.value_and_grad(p, spec.masks(), spec.perms, x, ctx)
    v2, gW2, gb2 = go.value_and_grad_flow(spec, p, x, ctx)
    assert abs(v1 - v2) < 1e-9 * max(1.0, abs(v1))
    for l in range(spec.L):
        for j in range(3):
            np.testing.assert_allclose(gW2[l][j], gW1[l][j], rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(gb2[l][j], gb1[l][j], rtol=1e-8, atol=1e-10)
    spec, draws, _, rng = make_case("nsa", 2, 1, [10, 10], 2, 1, seed=10)
    p = to64(_single(draws, 0))
    x = rng.normal(size=(6, 2)) * 1.2
    ctx = rng.uniform(size=(6, 1))
    v, gW, gb = go.value_and_grad_flow(spec, p, x, ctx)

    def f(pp):
        return fo.log_prob_draws(spec, [[(W[None], b[None]) for (W, b) in layer] for layer in pp], x, ctx)[0][0].sum()

    assert abs(f(p) - v) < 1e-8 * max(1.0, abs(v))
    masks = spec.masks()
    eps, checked = 1e-6, 0
    for (l, j) in [(0, 0), (1, 1), (0, 2), (1, 2)]:
        idxs = np.argwhere(masks[l][j] != 0)
        for idx in (tuple(idxs[0]), tuple(idxs[len(idxs) // 2]), tuple(idxs[-1])):
            pp = [[(W.copy(), b.copy()) for (W, b) in layer] for layer in p]
            pp[l][j][0][idx] += eps
            up = f(pp)
            pp[l][j][0][idx] -= 2 * eps
            dn = f(pp)
            assert abs((up - dn) / (2 * eps) - gW[l][j][idx]) < 2e-5 * max(1.0, abs(gW[l][j][idx])), (l, j, idx)
            checked += 1
    assert checked == 12


# ----------------------------------------------------------------------------------------------------------------
# neural-spline flows
# ----------------------------------------------------------------------------------------------------------------
def test_device_spline_derivative_routine_on_the_host_matches_autograd(built_lib):
    """naz_b200/csrc/spline_grad.cuh compiled for the CPU (nazb_host_spline_grad; no GPU needed): 1/T', d ld/dx and the raw-slot
    coefficients ca = -(dT/draw)/T', cb = -d ld/draw against torch autograd (fp64) of the oracle's spline, both orders, incl.
    the identity region outside [-B, B], the first / last bins (fixed end derivatives) and K != 8."""
    import ctypes as C
    import torch.nn.functional as F
    from naz_b200 import _lib
    from oracle import pyro_style as ps
    L = _lib.lib()
    rng = np.random.default_rng(1)
    errs, bins = [], set()
    for trial in range(600):
        K = (8, 8, 5, 12)[trial % 4]
        linear = (trial // 4) % 2
        M, B = (4 if linear else 3) * K - 1, 3.0
        raw = (rng.normal(size=M) * (0.5 + 1.5 * rng.uniform())).astype(np.float32)
        x = np.float32(rng.uniform(-3.4, 3.4) if trial % 7 else rng.choice([-2.999, 2.999, -3.2, 0.0]))
        ca, cb = np.zeros(M, np.float32), np.zeros(M, np.float32)
        itx, ldx = C.c_float(), C.c_float()
        assert L.nazb_host_spline_grad(float(x), K, B, linear, raw.ctypes.data, ca.ctypes.data, cb.ctypes.data, C.addressof(itx),
                                       C.addressof(ldx)) == 0
        r = torch.tensor(raw.astype(np.float64), requires_grad=True)
        xt = torch.tensor(float(x), dtype=torch.float64, requires_grad=True)
        y, ld = ps.monotonic_rational_spline(xt[None], F.softmax(r[:K], -1)[None], F.softmax(r[K:2 * K], -1)[None],
                                             F.softplus(r[2 * K:3 * K - 1])[None],
                                             torch.sigmoid(r[3 * K - 1:])[None] if linear else None, bound=B)
        gy = torch.autograd.grad(y.sum(), [r, xt], retain_graph=True, allow_unused=True)
        gl = torch.autograd.grad(ld.sum(), [r, xt], allow_unused=True)
        z = lambda t, n: np.zeros(n) if t is None else t.numpy()
        tx = float(z(gy[1], ())) if abs(float(x)) <= B else 1.0
        ca_ref, cb_ref = -z(gy[0], M) / tx, -z(gl[0], M)
        if abs(float(x)) > B:
            assert itx.value == 1.0 and ldx.value == 0.0 and not ca.any() and not cb.any()
            continue
        bins.add(int(np.argmax(np.abs(ca_ref[:K]) > 0)) if np.any(ca_ref[:K]) else -1)
        errs.append([np.abs(ca - ca_ref).max() / max(1.0, np.abs(ca_ref).max()), np.abs(cb - cb_ref).max() / max(1.0, np.abs(cb_ref).max()),
                     abs(itx.value - 1.0 / tx) * tx, abs(ldx.value - float(z(gl[1], ()))) / max(1.0, abs(float(z(gl[1], ()))))])
    e = np.array(errs)
    # fp32 against fp64: round-off in the median; the tail are narrow bins (width ~ 1e-3 of the box: (x - X0) / W cancels)
    assert len(e) > 400 and np.median(e) < 5e-6 and np.percentile(e, 99) < 5e-4 and e.max() < 3e-2, (np.median(e, 0), np.percentile(e, 99, 0), e.max(0))


SPLINE_GRAD_CASES = [
    # D C hidden L K S N bounds ctx_rows order
    (2, 0, [16, 16], 3, 8, 2, 37, False, 0, "quadratic"),
    (3, 2, [24, 24], 4, 8, 2, 100, True, "N", "quadratic"),
    (4, 2, [150, 150, 150], 2, 8, 1, 70, False, 1, "quadratic"),
    (3, 1, [32, 32], 2, 5, 2, 45, False, "N", "quadratic"),
    (3, 2, [24, 24], 3, 8, 2, 90, True, "N", "linear"),
    (2, 1, [32, 32], 2, 6, 1, 50, False, 1, "linear"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", SPLINE_GRAD_CASES)
def test_inverse_grad_spline_matches_autograd_oracle(case):
    """nazb_inverse_grad on quadratic neural-spline flows against torch autograd (fp64) through the module-structured
    restatement — what the reference's torch path differentiates in `train` (train_flows.py:195-213)."""
    from helpers import engine_for
    D, C, hidden, L, K, S, N, bounded, crow, order = case
    spec, draws, _, rng = make_case("nsa", D, C, hidden, L, S, seed=13, count_bins=K, order=order)
    x = (rng.normal(size=(N, D)) * 1.3).astype(np.float32)
    x[0, 0] = 3.5                                   # one coordinate in the identity region of the last layer's spline
    ctx = None
    if C:
        ctx = rng.uniform(size=(N if crow == "N" else 1, C)).astype(np.float32)
    bounds = (np.full(D, -7.0, np.float32), np.full(D, 7.0, np.float32)) if bounded else None
    eng = engine_for(spec, draws, engine="simt")
    r = eng.inverse_grad(torch.from_numpy(x), None if ctx is None else torch.from_numpy(ctx),
                         None if bounds is None else (torch.from_numpy(bounds[0]), torch.from_numpy(bounds[1])),
                         want_dx=True, want_lp=True)
    torch.cuda.synchronize()
    masks = spec.masks()
    for s in range(S):
        p = to64(_single(draws, s))
        c64 = None if ctx is None else (ctx[0] if ctx.shape[0] == 1 else ctx).astype(np.float64)
        b64 = None if bounds is None else tuple(b.astype(np.float64) for b in bounds)
        val, gW, gb, dx = go.value_and_grad_flow(spec, p, x.astype(np.float64), c64, b64, want_dx=True)
        assert abs(float(r["sum_n"][s]) - val) <= 2e-4 * max(1.0, abs(val))
        for l in range(L):
            for j in range(len(hidden) + 1):
                assert _rel(r["gW"][l][j][s].cpu().numpy(), gW[l][j]) < 5e-4, (s, l, j, "W")
                assert _rel(r["gb"][l][j][s].cpu().numpy(), gb[l][j]) < 5e-4, (s, l, j, "b")
                assert np.all(r["gW"][l][j][s].cpu().numpy()[masks[l][j] == 0] == 0.0)
        assert _rel(r["dx"][s].cpu().numpy(), dx) < 5e-4


@pytest.mark.gpu
def test_inverse_grad_spline_directional_derivative_at_bench_shape():
    """Size-independent property on the benchmarked architecture (config 3: 4|2, [150] x 3, 16 layers, K = 8): the central
    difference of the VALUE — computed by the tensor-core log_prob engine, an independent code path — along the normalised
    gradient equals the gradient's norm."""
    from helpers import engine_for
    S, N = 1, 20_000
    spec, draws, _, rng = make_case("nsa", 4, 2, [150] * 3, 16, S, seed=23)
    x = torch.from_numpy((rng.normal(size=(N, 4)) * 1.2).astype(np.float32))
    ctx = torch.from_numpy(rng.uniform(size=(N, 2)).astype(np.float32))
    r = engine_for(spec, draws, engine="simt").inverse_grad(x, ctx)
    torch.cuda.synchronize()
    gW = [[t.cpu().numpy().astype(np.float64) for t in layer] for layer in r["gW"]]
    gb = [[t.cpu().numpy().astype(np.float64) for t in layer] for layer in r["gb"]]
    norm = np.sqrt(sum((g ** 2).reshape(S, -1).sum(1) for layer in gW for g in layer) +
                   sum((g ** 2).reshape(S, -1).sum(1) for layer in gb for g in layer))
    assert np.all(np.isfinite(norm)) and np.all(norm > 0)
    eps = np.clip(20.0 / norm, 1e-5, 1e-3)

    def value(sign):
        pert = [[((W + sign * (eps / norm)[:, None, None] * gW[l][j]).astype(np.float32),
                  (b + sign * (eps / norm)[:, None] * gb[l][j]).astype(np.float32))
                 for j, (W, b) in enumerate(layer)] for l, layer in enumerate(draws)]
        e = engine_for(spec, pert, engine="auto")
        return e.inverse(x, ctx, want_lp=False, want_sum=True)["sum_n"].cpu().numpy(), e.engine_for("inverse")

    vp, used = value(+1.0)
    vm, _ = value(-1.0)
    assert used == "tcgen05"
    fd = (vp - vm) / (2 * eps)
    assert np.all(np.abs(fd - norm) <= 3e-2 * norm), (fd, norm, eps)


@pytest.mark.gpu
@pytest.mark.parametrize("flow_type,random_perm", [("maf", False), ("nsa", False), ("nsa", True), ("maf", True)])
def test_reference_training_step_runs_through_log_prob_backward(flow_type, random_perm):
    """The reference's MLE loop body (train_flows.py:195-213): `loss = -flow.log_prob(x, condition=y).mean(); loss.backward();
    optimizer.step()` on the product's NormalizingFlow.  The parameter gradients autograd receives from `_LogProbFn`
    (one nazb_inverse_vjp launch) must equal fp64 autograd through the module-structured restatement — also with Permute
    layers between the flow layers (their re-labelling is undone for the gradients) and with a bounded flow."""
    from naz_b200.flows import NormalizingFlow
    from oracle import pyro_style as ps
    torch.manual_seed(5)
    D, C, hidden, L, K = 3, 2, [32, 32], 3, 6
    bounds = {"low": torch.tensor([-6.0, -5.0, -7.0]).cuda(), "high": torch.tensor([6.0, 7.0, 5.0]).cuda()}
    args = (D, C, hidden, L) + ((K,) if flow_type == "nsa" else ())
    flow = NormalizingFlow(flow_type, bounds, *args, random_perm=random_perm).cuda()
    N = 257
    x = (torch.randn(N, D) * 1.2).cuda()
    y = torch.rand(N, C).cuda()
    parameters = []
    for t in flow.flow_dist.transforms:
        parameters.extend(list(t.parameters()))
    optimizer = torch.optim.Adam(parameters, lr=1e-3)
    flow.train()
    optimizer.zero_grad()
    loss = -flow.log_prob(x, condition=y).mean()
    loss.backward()
    got = [[(lin.weight.grad.detach().cpu().numpy().astype(np.float64), lin.bias.grad.detach().cpu().numpy().astype(np.float64))
            for lin in arn.layers] for arn in flow.nets]
    # checker
    torch.set_default_dtype(torch.float64)
    try:
        step = 2 if random_perm else 1
        extras = [[ps.Permute(flow.transforms[step * l + 1].permutation.cpu())] for l in range(L)] if random_perm else None
        b64 = {k: v.double().cpu() for k, v in bounds.items()}
        ref = ps.PyroStyleFlow(flow_type, b64, D, C, hidden, L, K, "quadratic", permutations=flow.perms().numpy(), extras=extras)
        ref.set_from_pytree([[(W.double().cpu().numpy(), b.double().cpu().numpy()) for (W, b) in layer] for layer in flow.current_draw()])
        ref_loss = -ref.log_prob(x.double().cpu(), y.double().cpu()).mean()
        ref_loss.backward()
        want = [[(lin.weight.grad.numpy(), lin.bias.grad.numpy()) for lin in arn.layers] for arn in ref.nets]
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item()))
    for l in range(L):
        for j in range(len(hidden) + 1):
            assert _rel(got[l][j][0], want[l][j][0]) < 5e-4, (l, j, "W")
            assert _rel(got[l][j][1], want[l][j][1]) < 5e-4, (l, j, "b")
    # a few optimizer steps of the reference loop lower the loss; the validation pass runs under no_grad (train_flows.py:222-223)
    first = loss.item()
    for _ in range(20):
        optimizer.zero_grad()
        loss = -flow.log_prob(x, condition=y).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(flow.parameters(), 5.0)
        optimizer.step()
    flow.flow_dist.clear_cache()
    flow.eval()
    with torch.no_grad():
        val = -flow.log_prob(x, condition=y).mean()
    assert not val.requires_grad and float(val) < first
    # d lp / d x through the same node
    xg = x.clone().requires_grad_(True)
    flow.log_prob(xg, condition=y).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0


@pytest.mark.gpu
def test_train_driver_fits_a_conditional_gaussian():
    """naz_b200.trainers.train (signature and return tuple of train_flows.py:73-242) on x | y ~ N(A y + b, diag(s^2)): the
    validation loss must approach the entropy of the conditional, -E log p = sum log s + D/2 log(2 pi e)."""
    from naz_b200.flows import NormalizingFlow
    from naz_b200.trainers import train
    torch.manual_seed(2)
    D, C, N = 2, 2, 6000
    y = torch.rand(N, C).cuda()
    A = torch.tensor([[1.0, -0.5], [0.3, 0.8]]).cuda()
    s = torch.tensor([0.3, 0.6]).cuda()
    x = y @ A.T + 0.2 + s * torch.randn(N, D).cuda()
    flow = NormalizingFlow("maf", None, D, C, [32, 32], 3).cuda()
    with torch.no_grad():
        untrained = float(-flow.log_prob(x, condition=y).mean())
    flow, history, history_val, best_mse, best_epoch = train(flow, x, y, lr=5e-3, num_epochs=25, batch_frac=0.1, min_epochs=5,
                                                             patience=4, verbose=False)
    entropy = float(s.log().sum()) + 0.5 * D * np.log(2 * np.pi * np.e)
    assert len(history) == len(history_val) and 0 <= best_epoch < len(history_val)
    assert untrained - best_mse > 0.3, (untrained, best_mse)
    assert best_mse < entropy + 0.1, (best_mse, entropy)          # measured: entropy + 0.009
    # the best-validation weights were restored
    with torch.no_grad():
        assert not flow.log_prob(x[:10], condition=y[:10]).requires_grad


@pytest.mark.gpu
@pytest.mark.parametrize("flow_type,bcast", [("maf", False), ("nsa", False), ("nsa", True)])
def test_log_prob_backward_reaches_a_trainable_embedding_net(flow_type, bcast):
    """NormalizingFlow(embedding_net=...) (flow.py:30-36, :75): the context is the embedding of the raw condition, so training it
    needs d lp / d ctx.  nazb_inverse_vjp returns it per point (`dctx`); here the gradients autograd delivers to the embedding
    net's parameters, to the flow's own parameters and to x must equal fp64 autograd through the restatement — per-point
    conditions and one broadcast condition vector (whose cotangent is the sum over the points)."""
    from naz_b200.flows import NormalizingFlow
    from oracle import pyro_style as ps
    torch.manual_seed(9)
    D, C, Craw, hidden, L, K = 3, 2, 5, [32, 32], 3, 6
    emb = torch.nn.Sequential(torch.nn.Linear(Craw, 8), torch.nn.Tanh(), torch.nn.Linear(8, C))
    args = (D, C, hidden, L) + ((K,) if flow_type == "nsa" else ())
    flow = NormalizingFlow(flow_type, None, *args, embedding_net=emb).cuda()
    N = 130
    x = (torch.randn(N, D) * 1.2).cuda().requires_grad_(True)
    y = torch.randn(1 if bcast else N, Craw).cuda()
    wts = torch.rand(N).cuda() + 0.5                                   # non-uniform cotangents of lp
    loss = -(flow.log_prob(x, condition=y) * wts).sum()
    loss.backward()
    got_emb = [p.grad.detach().cpu().double().numpy() for p in emb.parameters()]
    got_w0 = flow.nets[0].layers[0].weight.grad.detach().cpu().double().numpy()
    got_x = x.grad.detach().cpu().double().numpy()
    torch.set_default_dtype(torch.float64)
    try:
        import copy
        emb64 = copy.deepcopy(emb).cpu().double()
        for p in emb64.parameters():
            p.grad = None
        ref = ps.PyroStyleFlow(flow_type, None, D, C, hidden, L, K, "quadratic", permutations=flow.perms().numpy())
        ref.set_from_pytree([[(W.double().cpu().numpy(), b.double().cpu().numpy()) for (W, b) in layer] for layer in flow.current_draw()])
        x64 = x.detach().cpu().double().requires_grad_(True)
        c64 = emb64(y.cpu().double())
        if bcast:
            c64 = c64.expand(N, C)
        ref_loss = -(ref.log_prob(x64, c64) * wts.cpu().double()).sum()
        ref_loss.backward()
        want_emb = [p.grad.numpy() for p in emb64.parameters()]
        want_w0 = ref.nets[0].layers[0].weight.grad.numpy()
        want_x = x64.grad.numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(loss.item() - ref_loss.item()) < 2e-4 * max(1.0, abs(ref_loss.item()))
    for g, w in zip(got_emb, want_emb):
        assert _rel(g, w) < 1e-3, (g, w)
    assert _rel(got_w0, want_w0) < 5e-4 and _rel(got_x, want_x) < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("flow_type", ["maf", "nsa"])
def test_log_prob_backward_through_eval_mode_batchnorm(flow_type):
    """use_batchnorm=True flows in eval() mode: the per-layer affine is a constant of the adjoint recursion (cotangents pass
    through y_hat = a y + b as 1/a).  Gradients of the conditioner parameters and of x against fp64 autograd of the restatement
    with explicit BatchNorm (eval) and Permute transforms."""
    from naz_b200.flows import NormalizingFlow
    from naz_b200.flows.transforms import BatchNorm
    from oracle import pyro_style as ps
    torch.manual_seed(4)
    D, C, hidden, L, K = 3, 2, [32, 32], 3, 6
    args = (D, C, hidden, L) + ((K,) if flow_type == "nsa" else ())
    flow = NormalizingFlow(flow_type, None, *args, random_perm=True, use_batchnorm=True)
    with torch.no_grad():
        for t in flow.transforms:
            if isinstance(t, BatchNorm):
                t.gamma.copy_(0.5 + torch.rand(D)); t.beta.copy_(0.3 * torch.randn(D))
                t.moving_mean.copy_(0.2 * torch.randn(D)); t.moving_variance.copy_(0.5 + torch.rand(D))
    flow = flow.cuda().eval()
    N = 200
    x = (torch.randn(N, D) * 1.1).cuda().requires_grad_(True)
    y = torch.rand(N, C).cuda()
    (-flow.log_prob(x, condition=y).mean()).backward()
    torch.set_default_dtype(torch.float64)
    try:
        extras = []
        for l in range(L):
            pm, bn = flow.transforms[3 * l + 1], flow.transforms[3 * l + 2]
            extras.append([ps.Permute(pm.permutation.cpu()),
                           ps.BatchNormEval(bn.gamma.detach().double().cpu(), bn.beta.detach().double().cpu(),
                                            bn.moving_mean.double().cpu(), bn.moving_variance.double().cpu(), bn.epsilon)])
        ref = ps.PyroStyleFlow(flow_type, None, D, C, hidden, L, K, "quadratic", permutations=flow.perms().numpy(), extras=extras)
        ref.set_from_pytree([[(W.double().cpu().numpy(), b.double().cpu().numpy()) for (W, b) in layer] for layer in flow.current_draw()])
        x64 = x.detach().cpu().double().requires_grad_(True)
        (-ref.log_prob(x64, y.cpu().double()).mean()).backward()
    finally:
        torch.set_default_dtype(torch.float32)
    for arn, rarn in zip(flow.nets, ref.nets):
        for lin, rlin in zip(arn.layers, rarn.layers):
            assert _rel(lin.weight.grad.cpu().double().numpy(), rlin.weight.grad.numpy()) < 5e-4
            assert _rel(lin.bias.grad.cpu().double().numpy(), rlin.bias.grad.numpy()) < 5e-4
    assert _rel(x.grad.cpu().double().numpy(), x64.grad.numpy()) < 5e-4
    flow.train()
    with pytest.raises(RuntimeError):
        flow.log_prob(x, condition=y)
