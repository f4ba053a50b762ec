"""Generate tests/golden/ref_twin_*.npz by EXECUTING the reference's own MAF code
(/root/reference/src/naz/flows/bflow_jax_maf.py) on CPU.

jax / numpyro / optax / h5py / physt are not installed here, so the reference module is loaded from its
file with a minimal stand-in: `jax.numpy` -> numpy (float32 arrays that carry the `.at[idx].set()` idiom the
reference uses), `jax.jit` -> identity, `jax.nn` -> the three elementwise functions, `random.normal(key, shape)`
-> the array passed as `key`, everything else the module merely imports -> empty stubs.  No reference source
is copied: the functions that run (`create_mask`, `masked_linear`, `make_conditional_autoregressive_nn.nn_fn`,
`make_masked_affine_autoregressive_transform.{forward_fn,inverse_fn}`, `make_normalizing_flow.{log_prob,sample}`)
are the reference's bytes, executed as they are.  The outputs are therefore REFERENCE outputs for the
masked-affine (MAF) branch of the hot path; the spline branch lives in pyro-ppl and stays unpinned.

Run in the build container only (needs /root/reference):   python tools/make_reference_goldens.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/naz/flows/bflow_jax_maf.py"
F64 = os.environ.get("REF_DTYPE", "float64") == "float64"   # jax would run fp32; fp64 gives the checker more digits


class JArr(np.ndarray):
    """ndarray with jax's functional-update idiom: a.at[idx].set(v) -> updated copy."""

    class _At:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            a = self.a

            class _Set:
                def set(self_inner, v):
                    out = np.array(a, copy=True).view(JArr)
                    out[idx] = v
                    return out

            return _Set()

    @property
    def at(self):
        return JArr._At(self)


def _wrap(fn):
    def inner(*a, **k):
        out = fn(*a, **k)
        if isinstance(out, np.ndarray):
            return out.view(JArr)
        if isinstance(out, (list, tuple)):
            return type(out)(o.view(JArr) if isinstance(o, np.ndarray) else o for o in out)
        return out

    return inner


def make_shim():
    dt = np.float64 if F64 else np.float32
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            setattr(jnp, name, _wrap(obj))
        else:
            setattr(jnp, name, obj)
    jnp.array = lambda a, dtype=None: np.array(a, dtype=dtype if dtype is not None else (dt if np.asarray(a).dtype.kind == "f" else None)).view(JArr)
    jnp.float32 = dt                       # the reference asks for float32 explicitly in create_mask
    jnp.linspace = lambda a, b, n: np.linspace(a, b, n, dtype=np.float32).astype(dt).view(JArr)   # jax: fp32 linspace
    jnp.ndarray = np.ndarray
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    nn = types.ModuleType("jax.nn")
    nn.sigmoid = _wrap(lambda x: 1.0 / (1.0 + np.exp(-x)))
    nn.softplus = _wrap(lambda x: np.logaddexp(x, 0.0))
    nn.tanh = _wrap(np.tanh)
    jax.nn = nn
    rnd = types.ModuleType("jax.random")
    rnd.normal = lambda key, shape=None: np.asarray(key).reshape(shape).view(JArr)   # "key" IS the base noise
    rnd.PRNGKey = lambda s: s
    jax.random = rnd
    jax.value_and_grad = lambda f, **k: f
    jax.lax = types.ModuleType("jax.lax")
    fu = types.ModuleType("jax.flatten_util")
    fu.ravel_pytree = lambda t: (None, None)
    jax.flatten_util = fu
    mods = {"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.random": rnd, "jax.lax": jax.lax, "jax.flatten_util": fu}
    for name in ("optax", "numpyro", "numpyro.distributions", "numpyro.infer", "h5py", "physt"):
        mods[name] = types.ModuleType(name)
    for a in ("MCMC", "NUTS", "Predictive", "SVI", "Trace_ELBO"):
        setattr(mods["numpyro.infer"], a, object)
    mods["physt"].h2 = mods["physt"].h = None
    mods["numpyro"].distributions = mods["numpyro.distributions"]
    mods["numpyro"].infer = mods["numpyro.infer"]
    # package context for the relative import `from ..statutils import ...`
    naz = types.ModuleType("naz"); naz.__path__ = []
    flows = types.ModuleType("naz.flows"); flows.__path__ = []
    st = types.ModuleType("naz.statutils"); st.hpd_vectorized = st.equal_quantile_binning_nd = None
    mods.update({"naz": naz, "naz.flows": flows, "naz.statutils": st})
    return mods


def load_reference():
    saved = {k: sys.modules.get(k) for k in make_shim()}
    sys.modules.update(make_shim())
    try:
        spec = importlib.util.spec_from_file_location("naz.flows.bflow_jax_maf", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


CASES = {
    # name: D, C, hidden, L, N, context per point?
    "ref_twin_maf_cond_3d": (3, 2, [16, 12], 3, 64, True),
    "ref_twin_maf_cond_6d": (6, 4, [40, 40, 40], 4, 48, True),
    "ref_twin_maf_uncond_2d": (2, 0, [16, 16], 3, 64, False),
    "ref_twin_maf_bcast_ctx_2d": (2, 2, [24, 24, 24], 5, 40, False),   # one context vector for all points (calibrate.py:85,126)
}


# Bench-depth cases ([150] x 3 hidden, 16 flow layers: where the fp16 hi/lo error of the tensor-core path accumulates).  The
# weights (~3 MB per case) are NOT stored: they are regenerated from the seeded numpy Generator by the same statements below
# (tests/helpers.py::load_ref_twin replays them and verifies a checksum); inputs, permutations and the REFERENCE outputs are.
DEEP_CASES = {
    "ref_twin_deep_maf_2d": (2, 2, [150, 150, 150], 16, 256, True),          # cfg 4 architecture
    "ref_twin_deep_maf_6d": (6, 4, [150, 150, 150], 16, 256, True),          # cfg 2 architecture
    "ref_twin_deep_maf_8d_bcast": (8, 4, [150, 150, 150], 16, 256, False),   # cfg 5 architecture, one context vector (+ sampler)
}


def main(out_dir=None):
    ref = load_reference()
    dt = np.float64 if F64 else np.float32
    out_dir = out_dir or os.path.join(ROOT, "tests", "golden")
    for name, (D, C, hidden, L, N, per_point) in list(CASES.items()) + list(DEEP_CASES.items()):
        deep = name in DEEP_CASES
        rng = np.random.default_rng(sum(map(ord, name)))
        perms = np.stack([rng.permutation(D) for _ in range(L)])
        # the reference's own conditioner / transform factories and masks
        nn_fn, param_shapes, _gen = ref.make_conditional_autoregressive_nn(D, C, hidden)
        transform = ref.make_masked_affine_autoregressive_transform(nn_fn, D)
        masks, mask_skips = [], []
        for l in range(L):
            m, ms = ref.create_mask(D, C, hidden, np.asarray(perms[l]).view(JArr), 2)
            masks.append([np.asarray(a, dtype=dt).view(JArr) for a in m])
            mask_skips.append(np.asarray(ms, dtype=dt).view(JArr))
        params = []
        dims = [D + C] + list(hidden) + [2 * D]
        for l in range(L):
            lay = []
            for j in range(len(dims) - 1):
                W = (rng.normal(size=(dims[j + 1], dims[j])) / np.sqrt(dims[j])).astype(np.float32)
                b = (rng.normal(size=(dims[j + 1],)) * 0.1).astype(np.float32)
                lay.append((W.astype(dt).view(JArr), b.astype(dt).view(JArr)))
            params.append(lay)
        x = (rng.normal(size=(N, D)) * 1.5).astype(np.float32)
        ctx = None
        if C:
            ctx = rng.uniform(size=(N, C) if per_point else (C,)).astype(np.float32)
        flow = ref.make_normalizing_flow(transform, x.astype(dt).view(JArr), masks, mask_skips,
                                         [np.asarray(p).view(JArr) for p in perms], bounds=None,
                                         context=None if ctx is None else ctx.astype(dt).view(JArr))
        lp = np.asarray(flow["lp"](params), dtype=np.float64)
        arrs = dict(kind="maf", D=D, C=C, hidden=np.array(hidden), L=L, perms=perms, x=x, lp=lp, dtype=str(np.dtype(dt)))
        if ctx is not None:
            arrs["ctx"] = ctx
        # sampler: only defined upstream for a 1-D context (bflow_jax_maf.py:216-218)
        if ctx is not None and ctx.ndim == 1:
            zin = rng.normal(size=(N, D)).astype(np.float32)
            y, log_j = flow["sampler"](params, zin.astype(dt), N)
            arrs.update(zin=zin, ys=np.asarray(y, np.float64), log_j=np.asarray(log_j, np.float64))
        if deep:
            arrs["weights_regenerated"] = True
            arrs["weights_checksum"] = float(sum(np.abs(np.asarray(W, np.float64)).sum() + np.abs(np.asarray(b, np.float64)).sum()
                                                 for lay in params for (W, b) in lay))
        for l in range(L):
            if deep:
                continue
            for j, m in enumerate(masks[l]):
                arrs[f"mask_{l}_{j}"] = np.asarray(m, np.float32)
            for j, (W, b) in enumerate(params[l]):
                arrs[f"W_{l}_{j}"] = np.asarray(W, np.float32)
                arrs[f"b_{l}_{j}"] = np.asarray(b, np.float32)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
        print(name, "lp", lp.min(), lp.max(), "sampler" if "ys" in arrs else "", os.path.getsize(os.path.join(out_dir, name + ".npz")))


def bounded_sampler_fixture(out_dir=None):
    """The reference sampler WITH bounds (bflow_jax_maf.py:213-223: inverse_bounding_transform after the flow, its
    log-Jacobian added to the second output).  Only the sampler is stored: upstream's bounded `lp` filters the points with
    an integer 0/1 index (:198, `x[jnp.prod(...)]`), which gathers rows 0 / 1 instead of masking — not a behaviour to pin."""
    ref = load_reference()
    dt = np.float64 if F64 else np.float32
    out_dir = out_dir or os.path.join(ROOT, "tests", "golden")
    name, D, C, hidden, L, N = "ref_twin_maf_bounded_sampler_3d", 3, 2, [24, 24], 4, 64
    rng = np.random.default_rng(sum(map(ord, name)))
    perms = np.stack([rng.permutation(D) for _ in range(L)])
    nn_fn, _, _ = ref.make_conditional_autoregressive_nn(D, C, hidden)
    transform = ref.make_masked_affine_autoregressive_transform(nn_fn, D)
    masks, mask_skips = [], []
    for l in range(L):
        m, ms = ref.create_mask(D, C, hidden, np.asarray(perms[l]).view(JArr), 2)
        masks.append([np.asarray(a, dtype=dt).view(JArr) for a in m])
        mask_skips.append(np.asarray(ms, dtype=dt).view(JArr))
    dims = [D + C] + list(hidden) + [2 * D]
    params = [[((rng.normal(size=(dims[j + 1], dims[j])) / np.sqrt(dims[j])).astype(np.float32).astype(dt).view(JArr),
                (rng.normal(size=(dims[j + 1],)) * 0.1).astype(np.float32).astype(dt).view(JArr)) for j in range(len(dims) - 1)] for _ in range(L)]
    low, high = np.array([-1.0, 0.0, 2.0], dtype=np.float32), np.array([3.0, 1.0, 7.0], dtype=np.float32)
    x = (low + (high - low) * rng.uniform(0.05, 0.95, size=(N, D))).astype(np.float32)
    ctx = rng.uniform(size=(C,)).astype(np.float32)
    flow = ref.make_normalizing_flow(transform, x.astype(dt).view(JArr), masks, mask_skips, [np.asarray(p).view(JArr) for p in perms],
                                     bounds={"low": low.astype(dt).view(JArr), "high": high.astype(dt).view(JArr)}, context=ctx.astype(dt).view(JArr))
    zin = rng.normal(size=(N, D)).astype(np.float32)
    y, log_j = flow["sampler"](params, zin.astype(dt), N)
    arrs = dict(kind="maf", D=D, C=C, hidden=np.array(hidden), L=L, perms=perms, x=x, ctx=ctx, low=low, high=high, zin=zin,
                ys=np.asarray(y, np.float64), log_j=np.asarray(log_j, np.float64), dtype=str(np.dtype(dt)))
    for l in range(L):
        for j, m in enumerate(masks[l]):
            arrs[f"mask_{l}_{j}"] = np.asarray(m, np.float32)
        for j, (W, b) in enumerate(params[l]):
            arrs[f"W_{l}_{j}"] = np.asarray(W, np.float32)
            arrs[f"b_{l}_{j}"] = np.asarray(b, np.float32)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
    print(name, "ys in bounds:", bool(((np.asarray(y) > low) & (np.asarray(y) < high)).all()), os.path.getsize(os.path.join(out_dir, name + ".npz")))


def stats_fixture(out_dir=None):
    """hpd_vectorized of the REAL reference module (src/naz/statutils.py is plain numpy + pandas, importable as is)."""
    out_dir = out_dir or os.path.join(ROOT, "tests", "golden")
    spec = importlib.util.spec_from_file_location("ref_statutils", "/root/reference/src/naz/statutils.py")
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    rng = np.random.default_rng(7)
    arrs = {}
    for i, (S, nx, ny, alpha) in enumerate([(200, 5, 7, 0.1), (37, 4, 4, 0.32), (1000, 3, 2, 0.05), (64, 6, 6, 0.5)]):
        v = rng.gamma(2.0, 1.0, size=(S, nx, ny)).astype(np.float32)
        v[:, 0, 0] = np.round(v[:, 0, 0])             # ties: argmin must take the first minimum
        out = st.hpd_vectorized(v, alpha)
        arrs[f"v_{i}"] = v; arrs[f"alpha_{i}"] = alpha; arrs[f"hpd_{i}"] = np.asarray(out)
    np.savez_compressed(os.path.join(out_dir, "ref_stats_hpd.npz"), n=4, **arrs)
    print("ref_stats_hpd", os.path.getsize(os.path.join(out_dir, "ref_stats_hpd.npz")))


def truncnorm_fixture(out_dir=None):
    """TruncatedNormalTransform of the REAL reference module (src/naz/priors/TruncatedNormal.py: pure torch math; only its
    `pyro.distributions` import and `utils.set_device` are stubbed), executed on CPU in fp32 as the reference would."""
    import torch
    out_dir = out_dir or os.path.join(ROOT, "tests", "golden")
    stubs = {"pyro": types.ModuleType("pyro"), "pyro.distributions": types.ModuleType("pyro.distributions"),
             "utils": types.ModuleType("utils")}
    stubs["pyro"].distributions = stubs["pyro.distributions"]
    stubs["utils"].set_device = lambda t, *a, **k: t
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_truncnorm", "/root/reference/src/naz/priors/TruncatedNormal.py")
        tn = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tn)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(11)
    S, P = 6, 500
    arrs = {}
    # (0) the torch-path guide: per-parameter bounds mean -/+ sigma (bflow.py:30-35); (1) the twin's guide: [-1, 1] (bflow_jax_maf.py:251-257)
    mean = (rng.normal(size=P) * 0.5).astype(np.float32)
    sigma = (np.abs(mean) * 0.25 + 0.05).astype(np.float32)
    cases = [(mean, sigma, mean - sigma, mean + sigma),
             (rng.uniform(-0.5, 0.5, size=P).astype(np.float32), rng.uniform(0.3, 1.0, size=P).astype(np.float32),
              np.float32(-1.0), np.float32(1.0))]
    for i, (loc, scale, low, high) in enumerate(cases):
        x = rng.uniform(0.02, 0.98, size=(S, P)).astype(np.float32)
        tr = tn.TruncatedNormalTransform(torch.from_numpy(np.asarray(loc)), torch.from_numpy(np.asarray(scale)),
                                         torch.as_tensor(low), torch.as_tensor(high))
        xt = torch.from_numpy(x)
        y = tr(xt)
        log_q = -(tr.log_abs_det_jacobian(xt, y)).sum(-1)          # Uniform(0,1) base: log q(y) = 0 - log|dy/dx|
        arrs.update({f"x_{i}": x, f"loc_{i}": np.asarray(loc), f"scale_{i}": np.asarray(scale), f"low_{i}": np.asarray(low),
                     f"high_{i}": np.asarray(high), f"y_{i}": y.numpy(), f"log_q_{i}": log_q.numpy()})
    np.savez_compressed(os.path.join(out_dir, "ref_truncnorm.npz"), n=len(cases), **arrs)
    print("ref_truncnorm", os.path.getsize(os.path.join(out_dir, "ref_truncnorm.npz")))


# ----------------------------------------------------------------------------------------------------------------------
# Gradient fixtures: the reference differentiates its OWN log_prob with jax.grad / jax.value_and_grad
# (bflow_jax_maf.py:233-246, :277-287).  Here the same bytes run with torch standing in for jax.numpy (float64), and
# torch.autograd plays jax.grad: the gradients are reverse-mode derivatives of the reference's code, not of a restatement.
# ----------------------------------------------------------------------------------------------------------------------
def make_torch_shim():
    import torch

    class JT(torch.Tensor):
        """torch tensor carrying jax's functional update  a.at[idx].set(v)  (out-of-place, differentiable)."""

        class _At:
            def __init__(self, a):
                self.a = a

            def __getitem__(self, idx):
                a = self.a

                class _Set:
                    def set(self_inner, v):
                        out = a.clone()
                        out[idx] = v
                        return out

                return _Set()

        @property
        def at(self):
            return JT._At(self)

    def T(a):
        if isinstance(a, torch.Tensor):
            return a.as_subclass(JT)
        return torch.as_tensor(np.asarray(a)).as_subclass(JT)

    jnp = types.ModuleType("jax.numpy")
    jnp.pi = np.pi
    jnp.ndarray = torch.Tensor
    jnp.array = lambda a, dtype=None: T(a)
    jnp.broadcast_to = lambda a, shape: T(a).expand(*shape)
    jnp.concatenate = lambda xs, axis=0: torch.cat([T(x) for x in xs], dim=axis)
    jnp.dot = lambda a, b: T(a) @ T(b)
    jnp.exp = lambda a: torch.exp(T(a))
    jnp.log = lambda a: torch.log(torch.as_tensor(a, dtype=torch.float64))
    jnp.clip = lambda a, lo, hi: torch.clamp(T(a), lo, hi)
    jnp.squeeze = lambda a: torch.squeeze(T(a))
    jnp.zeros_like = lambda a: torch.zeros_like(T(a)).as_subclass(JT)
    jnp.sum = lambda a, axis=None: T(a).sum() if axis is None else T(a).sum(dim=axis)
    jnp.split = lambda a, n, axis=0: torch.chunk(T(a), n, dim=axis)
    jnp.cumsum = lambda a: np.cumsum(np.asarray(a))          # construction-time integer bookkeeping only
    jnp.ones = lambda n: torch.ones(n, dtype=torch.float64).as_subclass(JT)
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    nn = types.ModuleType("jax.nn")
    nn.tanh = lambda a: torch.tanh(T(a))
    nn.sigmoid = lambda a: torch.sigmoid(T(a))
    nn.softplus = lambda a: torch.nn.functional.softplus(T(a))
    jax.nn = nn
    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey = lambda s: s
    jax.random = rnd
    jax.value_and_grad = lambda f, **k: f
    jax.lax = types.ModuleType("jax.lax")
    fu = types.ModuleType("jax.flatten_util")
    fu.ravel_pytree = lambda t: (None, None)
    jax.flatten_util = fu
    mods = make_shim()                                        # the empty stubs of everything the module merely imports
    mods.update({"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.random": rnd, "jax.lax": jax.lax, "jax.flatten_util": fu})
    return mods, T


def grad_fixture(out_dir=None):
    """tests/golden/ref_twin_grad_*.npz: d(sum_n lp)/d(W, b) and d lp / d x of the reference's own make_normalizing_flow
    log_prob, at the parameters / points of the committed ref_twin_* fixtures (which also pins the value: the lp this run
    computes must equal the lp stored there by the numpy-backed run)."""
    import torch
    out_dir = out_dir or os.path.join(ROOT, "tests", "golden")
    mods, T = make_torch_shim()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("naz.flows.bflow_jax_maf", REF)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for name in ("ref_twin_maf_cond_3d", "ref_twin_maf_cond_6d", "ref_twin_maf_bcast_ctx_2d"):
        g = np.load(os.path.join(out_dir, name + ".npz"))
        D, C, L = int(g["D"]), int(g["C"]), int(g["L"])
        hidden = [int(h) for h in g["hidden"]]
        nn_fn, _, _ = ref.make_conditional_autoregressive_nn(D, C, hidden)
        transform = ref.make_masked_affine_autoregressive_transform(nn_fn, D)
        n_lin = len(hidden) + 1
        masks = [[T(g[f"mask_{l}_{j}"].astype(np.float64)) for j in range(n_lin)] for l in range(L)]
        params = [[(T(g[f"W_{l}_{j}"].astype(np.float64)).requires_grad_(True), T(g[f"b_{l}_{j}"].astype(np.float64)).requires_grad_(True))
                   for j in range(n_lin)] for l in range(L)]
        x = T(g["x"].astype(np.float64)).requires_grad_(True)
        ctx = T(g["ctx"].astype(np.float64)) if C else None
        flow = ref.make_normalizing_flow(transform, x, masks, [None] * L, [np.asarray(p) for p in g["perms"]], bounds=None, context=ctx)
        lp = flow["lp"](params)
        assert np.allclose(lp.detach().numpy(), g["lp"], rtol=1e-12, atol=1e-12), "torch-backed run must reproduce the stored lp"
        lp.sum().backward()
        arrs = {"sum_lp": float(lp.detach().sum()), "dx": x.grad.numpy()}
        for l in range(L):
            for j in range(n_lin):
                arrs[f"gW_{l}_{j}"] = params[l][j][0].grad.numpy()
                arrs[f"gb_{l}_{j}"] = params[l][j][1].grad.numpy()
        np.savez_compressed(os.path.join(out_dir, name.replace("ref_twin_", "ref_twin_grad_") + ".npz"), **arrs)
        print(name, "grad", arrs["sum_lp"], os.path.getsize(os.path.join(out_dir, name.replace("ref_twin_", "ref_twin_grad_") + ".npz")))


if __name__ == "__main__":
    main()
    stats_fixture()
    truncnorm_fixture()
    grad_fixture()
    bounded_sampler_fixture()
