"""Dev tool: violators of the strict tolerance vs the conditioning of the point (sensitivity of the fp64 oracle to a 1-ulp(fp32)
relative perturbation of the inputs)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for, to64
from oracle import flow_oracle as fo
for kind, D, C, hidden, L, S, N, bcast, seed in [("nsa", 4, 2, [150]*3, 16, 3, 4000, True, 77), ("nsa", 4, 2, [150]*3, 16, 3, 4000, False, 78), ("nsa", 4, 2, [150]*3, 16, 3, 4000, True, 79)]:
    spec, draws, _, rng = make_case(kind, D, C, hidden, L, S, seed=seed)
    x = (rng.normal(size=(N, D)) * 1.5).astype(np.float32)
    ctx = rng.uniform(size=(C,) if bcast else (N, C)).astype(np.float32)
    d64 = to64(draws)
    lp64, z64 = fo.log_prob_draws(spec, d64, x.astype(np.float64), ctx.astype(np.float64))
    eps = 2.0 ** -23
    sens = np.zeros_like(lp64); sensz = np.zeros_like(z64)
    for sgn in (+1, -1):
        lpp, zp = fo.log_prob_draws(spec, d64, x.astype(np.float64) * (1 + sgn * eps), ctx.astype(np.float64) * (1 + sgn * eps))
        sens = np.maximum(sens, np.abs(lpp - lp64)); sensz = np.maximum(sensz, np.abs(zp - z64))
    lp32, z32 = fo.log_prob_draws(spec, draws, x, ctx)
    for engine in ("tcgen05", "simt"):
        eng = engine_for(spec, draws, engine=engine)
        out = eng.inverse(torch.from_numpy(x), torch.from_numpy(ctx), want_lp=True, want_z=True)
        for name, got, r64, r32, sn in (("lp", out["lp"], lp64, lp32, sens), ("z", out["z"], z64, z32, sensz)):
            got = got.cpu().numpy().astype(np.float64)
            err = np.abs(got - r64); tol = 1e-5 + 1e-4 * np.abs(r64); bad = err > tol
            e32 = np.abs(r32.astype(np.float64) - r64)
            print(f"{kind} bcast={bcast} seed={seed} {engine} {name}: viol={bad.mean():.5f} worst={np.max(err/tol):.2f}x; sens/tol quantiles (all points) 50%={np.quantile(sn/tol,0.5):.4f} 99%={np.quantile(sn/tol,0.99):.4f} 99.9%={np.quantile(sn/tol,0.999):.4f}")
            for i in np.argwhere(bad)[:12]:
                i = tuple(i)
                print(f"     violator err/tol={err[i]/tol[i]:.2f} e32/tol={e32[i]/tol[i]:.2f} sens/tol={sn[i]/tol[i]:.3f} err/sens={err[i]/max(sn[i],1e-300):.1f}")
            good = sn <= tol / 20
            print(f"     well-conditioned points (sens <= tol/20): {good.mean():.4f} of all; violators among them: {int((bad & good).sum())}; worst err/tol among them: {np.max((err/tol)[good]):.2f}", flush=True)
