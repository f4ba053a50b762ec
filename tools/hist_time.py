"""Timing of nazb_histogramdd (C ABI called directly: counts memset + histdd_kernel): 1000 draws x 100 k points x 4-D, 8^4 bins
(calibrate_4p) and 2-D 32 x 32."""
import ctypes as C
import sys
sys.path.insert(0, "/root/repo")
import torch
from naz_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
for (S, N, D, nb) in [(1000, 100_000, 4, 8), (1000, 100_000, 2, 32), (4400, 100_000, 4, 8)]:
    x = torch.randn(S, N, D, device="cuda")
    edges = torch.cat([torch.linspace(-3, 3, nb + 1, dtype=torch.float64) for _ in range(D)]).cuda()
    counts = torch.empty((S, nb ** D), dtype=torch.int32, device="cuda")
    nba = (C.c_int32 * D)(*([nb] * D))
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: L.nazb_histogramdd(x.data_ptr(), S, N, D, edges.data_ptr(), nba, counts.data_ptr(), None, st)
    assert run() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"S={S} N={N} D={D} bins={nb}^{D}: {ms:.3f} ms = {S * N * D * 4 / ms / 1e6:.0f} GB/s of samples; total counted {int(counts.sum())}")
