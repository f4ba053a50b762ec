"""Dev tool: sampling sweep timing (cfg 5 shape: S draws x N points per draw)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
kind, D, C, S, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
spec, draws, _, rng = make_case(kind, D, C, [150] * 3, 16, S, seed=4, scale=0.1)
eng = engine_for(spec, draws)
z = torch.randn((S, N, D), device="cuda")
ctx = torch.rand((C,)) if C else None
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x = eng.forward(z, ctx); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"{kind} {D}|{C} S={S} N={N} forward[{eng.engine_for('forward')}]: {ms:.2f} ms  {S*N/ms/1e3:.1f} M samples/s")
