import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from oracle import flow_oracle as fo
from helpers import make_case, engine_for, to64, tol_report
torch.manual_seed(0)
import os
ENG = os.environ.get("ENG", "auto")
for (kind, D, C, hidden, L, S, N, order, mode) in [
    ("maf", 2, 0, [64, 64], 5, 1, 1000, "quadratic", "incremental"),
    ("nsa", 2, 0, [64, 64], 5, 1, 1000, "quadratic", "incremental"),
    ("maf", 2, 2, [150]*3, 16, 3, 777, "quadratic", "incremental"),
    ("maf", 6, 4, [150]*3, 16, 2, 500, "quadratic", "jacobi"),
    ("nsa", 4, 2, [150]*3, 16, 2, 500, "quadratic", "incremental"),
    ("nsa", 4, 2, [150]*3, 4, 2, 300, "quadratic", "jacobi"),
    ("nsa", 3, 2, [32, 40], 3, 2, 300, "linear", "incremental"),
]:
    spec, draws, keep, rng = make_case(kind, D, C, hidden, L, S, seed=1, order=order)
    x = (rng.normal(size=(N, D)) * 1.5).astype(np.float32)
    ctx = rng.uniform(size=(N, C)).astype(np.float32) if C else None
    eng = engine_for(spec, draws, inverse_mode=mode, engine=ENG)
    t0=time.time()
    out = eng.inverse(torch.from_numpy(x), None if ctx is None else torch.from_numpy(ctx), want_z=True, want_lp=True, want_lse=True, want_sum=True)
    torch.cuda.synchronize()
    lp_ref, z_ref = fo.log_prob_draws(spec, to64(draws), x.astype(np.float64), None if ctx is None else ctx.astype(np.float64))
    lp = out["lp"].cpu().numpy(); z = out["z"].cpu().numpy()
    print(kind, D, C, mode, eng.engine_for("inverse"), eng.engine_for("forward"), "lp viol/worst", tol_report(lp, lp_ref), "z", tol_report(z, z_ref, 1e-4, 1e-5), "sum", out["sum_n"].cpu().numpy()[:2], lp_ref.sum(1)[:2])
    pp = eng.lse_finish(out["lse_max"], out["lse_sum"], -np.log(S)).cpu().numpy()
    print("    ppd", tol_report(pp, fo.posterior_predictive(lp_ref)))
    zin = rng.normal(size=(S, N, D)).astype(np.float32)
    xs, ld = eng.forward(torch.from_numpy(zin), None if ctx is None else torch.from_numpy(ctx), want_logdet=True)
    xs_ref, ld_ref = fo.sample_draws(spec, to64(draws), zin.astype(np.float64), None if ctx is None else ctx.astype(np.float64))
    print("    fwd x", tol_report(xs.cpu().numpy(), xs_ref), "ld", tol_report(ld.cpu().numpy(), ld_ref))
