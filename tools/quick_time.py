"""Quick engine timing at config-3/4/5-like shapes (dev tool; bench.py is the contract)."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
cases = {
 "cfg3": ("nsa", 4, 2, [150]*3, 16),
 "cfg4": ("maf", 2, 2, [150]*3, 16),
 "cfg2": ("maf", 6, 4, [150]*3, 16),
 "cfg5a": ("maf", 8, 4, [150]*3, 16),
}
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 148*128*2
kind, D, C, hidden, L = cases[which]
spec, draws, keep, rng = make_case(kind, D, C, hidden, L, S, seed=1)
x = torch.from_numpy((rng.normal(size=(N, D)) * 1.5).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, C)).astype(np.float32)).cuda() if C else None
z = torch.from_numpy(rng.normal(size=(N, D)).astype(np.float32)).cuda()
for engname in os.environ.get("ENGS", "tcgen05,simt").split(","):
    eng = engine_for(spec, draws, engine=engname)
    for direction in os.environ.get("DIRS", "inverse,forward").split(","):
        def run():
            if direction == "inverse":
                return eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=1)
            return eng.forward(z, ctx)
        run(); torch.cuda.synchronize()
        times = []
        for _ in range(int(os.environ.get("REPS", "3"))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        ev = S * N / (ms * 1e-3)
        fl = eng.shape.flops_per_eval() * ev
        print(f"{which} {engname:8s} {eng.engine_for(direction):8s} {direction:8s} S={S} N={N}: {ms:9.2f} ms  {ev/1e6:8.2f} Mevals/s  {fl/1e12:7.2f} TF/s algorithmic  -> 1e9 evals in {1e9/ev:6.1f} s")
    del eng
