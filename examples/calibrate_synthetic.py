"""End-to-end data flow of the paper scripts on the CUDA path, with synthetic inputs (no dataset / checkpoint is needed).

What examples/papers/2506.05657/calibrate.py does upstream, step by step, and what runs here instead:

  upstream (calibrate.py:85-155, bflow_jax_maf.py:405-460)            here
  ------------------------------------------------------------------  -----------------------------------------------------
  load the MLE flow + posterior {"standard_params", "scale"}           synthetic theta_0 and u ~ U(-1, 1) of the same shapes
  theta_s = theta_0 (1 + scale u_s), unravel, loop over the draws      FlowEngine.pack_draw_map: applied while packing
  per draw: sampler(params_s, key, N)  -> [S, N, D] on the host        ONE nazb_forward launch, the tensor stays on the device
  per draw: np.histogram2d / jnp.histogramdd(density=True)             nazb_histogramdd (numpy-exact counts)
  hpd_vectorized across draws (statutils.py:22-46)                     nazb_hpd
  posterior predictive log p(x_n) = log mean_s p(x_n | theta_s)        ONE nazb_inverse launch with the fused (max, sum exp)

Run on a B200:  python examples/calibrate_synthetic.py [S] [N]
"""
import math
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from naz_b200.flows import NormalizingFlow          # noqa: E402
from naz_b200.stats import histogramdd_draws, hpd_draws   # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    D, C = 2, 2
    flow = NormalizingFlow("maf", None, D, C, [150, 150, 150], 16).to(dev).eval()     # the twin's architecture (config 4)
    base = flow.current_draw()
    P = sum(W.numel() + b.numel() for layer in base for (W, b) in layer)
    u = torch.rand(S, P, device=dev) * 2 - 1                                            # posterior["standard_params"]
    scale = 0.05                                                                         # posterior["scale"]
    from naz_b200 import FlowEngine
    eng = FlowEngine(flow.shape, S, device=dev)                                          # one handle holds all S draws
    t0 = time.perf_counter()
    eng.pack_draw_map(base, u, scale, flow.masks(), flow.perms())
    cond = torch.tensor([0.3, 0.7], device=dev)                                          # one condition vector (calibrate.py:85)
    z = torch.randn(S, N, D, device=dev)
    x = eng.forward(z, cond)                                                             # [S, N, D]
    edges = [torch.linspace(-4, 4, 33, dtype=torch.float64) for _ in range(D)]
    counts, dens = histogramdd_draws(x, edges, density=True)                             # [S, 32, 32] each
    lo, hi = hpd_draws(dens.reshape(S, -1), 0.1)                                         # 90 % band across draws, per bin
    # posterior predictive of held-out points
    pts = x[0, :4096].contiguous()
    out = eng.inverse(pts, cond, want_lp=False, want_lse=True)
    ppd = eng.lse_finish(out["lse_max"], out["lse_sum"], -math.log(S))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"S = {S} draws x N = {N} samples: pack + sample + histogram + HPD + posterior predictive in {dt * 1e3:.1f} ms")
    print(f"  sample tensor {tuple(x.shape)}, density {tuple(dens.shape)}, band width (median bin) {float((hi - lo).median()):.4f}")
    print(f"  mean posterior-predictive log density of 4096 points: {float(ppd.mean()):.4f}")
    print(f"  engines: sample on {eng.engine_for('forward')}, log_prob on {eng.engine_for('inverse')}")


if __name__ == "__main__":
    main()
