"""Dev tool: HBM roofline of the sample-consumer kernels (histogramdd per draw, HPD across draws)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from naz_b200.stats import histogramdd_draws, hpd_draws
S, N, D, nb = 1000, 100_000, 4, 8
x = torch.randn((S, N, D), device="cuda") * 1.5
edges = [np.linspace(-3, 3, nb + 1) for _ in range(D)]
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: histogramdd_draws(x, edges, density=False))
gb = S * N * D * 4 / 1e9
print(f"histogramdd S={S} N={N} D={D} bins={nb}^{D}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s read ({gb:.2f} GB algorithmic)")
ms = t(lambda: histogramdd_draws(x, edges, density=True))
print(f"  + density: {ms:.3f} ms")
v = torch.rand((4400, 64 * 64), device="cuda")
ms = t(lambda: hpd_draws(v, 0.1))
gb = v.numel() * 4 / 1e9
print(f"hpd S=4400 M=4096: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s read ({gb:.3f} GB algorithmic)")
