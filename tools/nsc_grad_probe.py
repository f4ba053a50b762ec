"""Probe: gradient error (CUDA vs fp64 autograd of the restatement) against the spread of the data, coupling and autoregressive."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import explicit_coupling_flow
from naz_b200.flows import NormalizingFlow
from oracle import pyro_style as ps
D, s, C, hidden, K, N = 5, 2, 2, [48, 48], 8, 600
for kind in ("nsc", "nsa"):
    for (scale, L) in [(0.8, 1), (1.4, 1), (2.5, 1), (1.4, 3), (2.5, 3)]:
        torch.manual_seed(17)
        args = (D, C, hidden, L, K) + ((s,) if kind == "nsc" else ())
        flow = NormalizingFlow(kind, None, *args, engine="simt").cuda()
        x = (torch.randn(N, D) * scale).double(); ctx = torch.randn(N, C).double()
        flow.train(); flow.zero_grad()
        xg = x.float().cuda().requires_grad_(True)
        lp = flow.log_prob(xg, condition=ctx.float().cuda())
        (-lp.mean()).backward()
        x64 = x.clone().requires_grad_(True)
        if kind == "nsc":
            build = explicit_coupling_flow(flow, "quadratic")
            lp64 = build(ctx).log_prob(x64); leaves = build.leaves
        else:
            torch.set_default_dtype(torch.float64)
            ref = ps.PyroStyleFlow("nsa", None, D, C, hidden, L, K, "quadratic", permutations=flow.perms().numpy())
            ref.set_from_pytree([[(W.double().cpu().numpy(), b.double().cpu().numpy()) for (W, b) in layer] for layer in flow.current_draw()])
            lp64 = ref.log_prob(x64, ctx); leaves = [q for arn in ref.nets for lin in arn.layers for q in (lin.weight, lin.bias)]
            torch.set_default_dtype(torch.float32)
        (-lp64.mean()).backward()
        errs = [float((g.grad.detach().cpu().double() - w.grad).abs().max() / max(1e-12, float(w.grad.abs().max()))) for g, w in zip(flow._flat_params(), leaves)]
        elp = float(((lp.detach().cpu().double() - lp64.detach()).abs() / (1e-5 + 1e-4 * lp64.detach().abs())).max())
        edx = (xg.grad.detach().cpu().double() - x64.grad).abs()
        i = int(edx.max(1).values.argmax())
        print(f"{kind} scale={scale} L={L}: worst param err {max(errs):.1e} (#{int(np.argmax(errs))}), lp worst {elp:.2f} tol, dx err {float(edx.max() / x64.grad.abs().max()):.1e} at point {i} x={x[i].numpy().round(3)}")
