"""Generate tests/golden/*.npz from the oracle (fp64).  The reference ships no golden vectors and cannot be
imported here (pyro / jax absent), so these fixtures pin the ORACLE (and through it the CUDA path) against
regressions — they are not reference outputs ("parity unpinned", see oracle/flow_oracle.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as fo

CASES = {
    # name: kind, D, C, hidden, L, K, order, S, N, bounds
    "maf_uncond_2d": ("maf", 2, 0, [16, 16], 3, 8, "quadratic", 2, 64, False),
    "maf_cond_3d": ("maf", 3, 2, [24, 20], 3, 8, "quadratic", 3, 50, True),
    "nsa_cond_4d": ("nsa", 4, 2, [24, 24, 24], 3, 8, "quadratic", 2, 48, False),
    "nsa_uncond_2d_k5": ("nsa", 2, 0, [16, 16], 2, 5, "quadratic", 1, 40, False),
    "nsa_linear_3d": ("nsa", 3, 1, [20, 16], 2, 6, "linear", 2, 40, False),
}
out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)
for name, (kind, D, C, hidden, L, K, order, S, N, bounded) in CASES.items():
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    perms = np.stack([rng.permutation(D) for _ in range(L)])
    spec = fo.FlowSpec(kind, D, C, hidden, L, perms, count_bins=K, order=order)
    p0 = fo.init_weights(spec, rng, np.float32)
    draws = fo.perturb_draws(p0, S, 0.25, rng, np.float32)
    draws64 = [[(W.astype(np.float64), b.astype(np.float64)) for (W, b) in l] for l in draws]
    bounds = None
    if bounded:
        bounds = (np.full(D, -6.0), np.full(D, 6.0))
        x = rng.uniform(-5.5, 5.5, size=(N, D))
    else:
        x = rng.normal(size=(N, D)) * 1.5
    x = x.astype(np.float32).astype(np.float64)
    ctx = rng.uniform(size=(N, C)).astype(np.float32).astype(np.float64) if C else None
    lp, z = fo.log_prob_draws(spec, draws64, x, ctx, bounds)
    zin = rng.normal(size=(S, N, D)).astype(np.float32).astype(np.float64)
    xs, ld = fo.sample_draws(spec, draws64, zin, ctx, bounds)
    arrs = dict(kind=kind, D=D, C=C, hidden=np.array(hidden), L=L, K=K, order=order, perms=perms, x=x,
                lp=lp, z=z, zin=zin, xs=xs, ld=ld, ppd=fo.posterior_predictive(lp), bounded=bounded)
    if C:
        arrs["ctx"] = ctx
    for l in range(L):
        for j, (W, b) in enumerate(draws[l]):
            arrs[f"W_{l}_{j}"] = W
            arrs[f"b_{l}_{j}"] = b
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
    print(name, "lp range", lp.min(), lp.max(), os.path.getsize(os.path.join(out_dir, name + ".npz")))
