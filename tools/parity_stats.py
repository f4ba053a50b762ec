"""Dev tool: how the CUDA log_prob errors relate to the reference's own fp32 error (oracle run in fp32) per element."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for, to64
from oracle import flow_oracle as fo
shapes = [("nsa", 4, 2, [150]*3, 16, 3, 3000, False), ("nsa", 4, 2, [150]*3, 16, 3, 3000, True), ("maf", 2, 2, [150]*3, 16, 3, 3000, False),
          ("maf", 8, 4, [150]*3, 16, 2, 1500, True), ("maf", 6, 4, [150]*3, 16, 2, 1500, False)]
for kind, D, C, hidden, L, S, N, bcast in shapes:
    spec, draws, _, rng = make_case(kind, D, C, hidden, L, S, seed=77)
    x = (rng.normal(size=(N, D)) * 1.5).astype(np.float32)
    ctx = rng.uniform(size=(C,) if bcast else (N, C)).astype(np.float32)
    lp64, z64 = fo.log_prob_draws(spec, to64(draws), x.astype(np.float64), ctx.astype(np.float64))
    lp32, z32 = fo.log_prob_draws(spec, draws, x, ctx)
    for engine in ("tcgen05", "simt"):
        eng = engine_for(spec, draws, engine=engine)
        out = eng.inverse(torch.from_numpy(x), torch.from_numpy(ctx), want_lp=True, want_z=True)
        for name, got, r64, r32 in (("lp", out["lp"], lp64, lp32), ("z", out["z"], z64, z32)):
            got = got.cpu().numpy().astype(np.float64)
            err = np.abs(got - r64); tol = 1e-5 + 1e-4 * np.abs(r64); bad = err > tol
            e32 = np.abs(r32.astype(np.float64) - r64)
            d32 = np.abs(got - r32.astype(np.float64)); tol32 = 1e-5 + 1e-4 * np.abs(r32)
            ratio = err[bad] / np.maximum(e32[bad], 1e-30)
            print(f"{kind} {D}|{C} bcast={bcast} {engine:8s} {name}: viol={bad.mean():.5f} worst={np.max(err/tol):.2f}x  fp32-oracle viol={np.mean(e32>tol):.5f} worst={np.max(e32/tol):.2f}x  "
                  f"violators: n={bad.sum()} err/e32 median={np.median(ratio) if bad.any() else 0:.2f} max={ratio.max() if bad.any() else 0:.2f} "
                  f"frac(err<=3 e32)={np.mean(ratio<=3) if bad.any() else 1:.3f}  e32>0.5tol at violators={np.mean(e32[bad]>0.5*tol[bad]) if bad.any() else 1:.3f}  vs-fp32-oracle viol={np.mean(d32>tol32):.5f} worst={np.max(d32/tol32):.2f}x", flush=True)
