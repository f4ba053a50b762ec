"""Timing of the f1 row (nazb_inverse_grad): value + gradient of sum_n lp per draw, masked-affine and quadratic-spline flows.
1 grad-eval = one (draw, point) pair pushed through the value AND the full parameter gradient.
CPU leg (CPU=1): torch autograd of the oracle's restated twin (oracle/grad_oracle.py) in fp32 with all host threads —
the stand-in for the reference's jax.value_and_grad / autograd call (neither jax nor pyro exists here)."""
import json, os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for

cases = {
    "cfg4": ("maf", 2, 2, [150] * 3, 16),
    "cfg2": ("maf", 6, 4, [150] * 3, 16),
    "cfg5a": ("maf", 8, 4, [150] * 3, 16),
    "cfg3": ("nsa", 4, 2, [150] * 3, 16),        # the benchmarked spline architecture (K = 8, quadratic)
    "cfg5b": ("nsa", 8, 4, [150] * 3, 16),
}
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
kind, D, C, hidden, L = cases[which]
spec, draws, keep, rng = make_case(kind, D, C, hidden, L, S, seed=1)
x_np = (rng.normal(size=(N, D)) * 1.5).astype(np.float32)
c_np = rng.uniform(size=(N, C)).astype(np.float32)
out = {"case": which, "S": S, "N": N}
if torch.cuda.is_available():
    x, ctx = torch.from_numpy(x_np).cuda(), torch.from_numpy(c_np).cuda()
    eng = engine_for(spec, draws, engine="simt")
    eng.inverse_grad(x, ctx); torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get("REPS", "3"))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = eng.inverse_grad(x, ctx); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    out["gpu_ms"] = ms
    out["gpu_grad_evals_per_s"] = S * N / (ms * 1e-3)
    # forward-only (value) on the same engine for scale
    eng.inverse(x, ctx, want_lp=False, want_sum=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.inverse(x, ctx, want_lp=False, want_sum=True); e1.record(); torch.cuda.synchronize()
    out["gpu_value_only_ms_same_engine"] = e0.elapsed_time(e1)
if os.environ.get("CPU", "0") == "1":
    from oracle import grad_oracle as go
    torch.set_num_threads(os.cpu_count())
    n = int(os.environ.get("CPU_N", "20000"))
    P = [[(torch.tensor(W[0], requires_grad=True), torch.tensor(b[0], requires_grad=True)) for (W, b) in layer] for layer in draws]
    M = [[torch.tensor(m.astype(np.float32)) for m in ml] for ml in spec.masks()]
    xs, cs = torch.from_numpy(x_np[:n]), torch.from_numpy(c_np[:n])
    def step():
        for layer in P:
            for (W, b) in layer:
                W.grad = None; b.grad = None
        go.log_prob(P, M, spec.perms, xs, cs).sum().backward()
    step()
    t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
    out["cpu_grad_evals_per_s"] = n / dt
    out["cpu_threads"] = os.cpu_count()
    out["cpu_sample"] = f"1 draw x {n} points, fp32 torch autograd of the restated twin, {dt:.1f} s"
print(json.dumps(out))
