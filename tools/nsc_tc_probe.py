"""Probe: the tensor-core INVERSE program for the single-degree ladder of a coupling flow ('nsc'), engine option inv_gaps = 1."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from helpers import explicit_coupling_flow
from naz_b200.flows import NormalizingFlow
for (D, C, s, hidden, L, N) in [(5, 2, 2, [48, 48], 4, 600), (4, 2, 2, [150, 150, 150], 16, 20000), (5, 0, 2, [64, 64], 3, 3000)]:
    torch.manual_seed(1)
    flow = NormalizingFlow("nsc", None, D, C, hidden, L, 8, s).cuda().eval()
    x = (torch.randn(N, D) * 1.3); c = torch.rand(N, C) if C else None
    with torch.no_grad():
        ref = explicit_coupling_flow(flow, "quadratic")(None if c is None else c.double()).log_prob(x.double()) if N <= 3000 else None
        cg = None if c is None else c.cuda()
        lp_s = flow.log_prob(x.cuda(), condition=cg)
        eng = flow._single_engine()
        e_s = eng.engine_for("inverse")
        eng.set_option("inv_gaps", 1)
        eng.pack(flow._fold_draws(flow.current_draw()), flow._packed_masks(), flow._packed_perms())
        lp_t = eng.inverse(flow.relabel.to_engine(x.cuda()), cg, None, want_lp=True)["lp"][0]
        e_t = eng.engine_for("inverse")
        torch.cuda.synchronize()
        def worst(a, b):
            return float(((a.double().cpu() - b.double().cpu()).abs() / (1e-5 + 1e-4 * b.double().cpu().abs())).max())
        msg = f"D={D} C={C} hidden={hidden} L={L}: default inverse on {e_s}, inv_gaps=1 on {e_t}; tc vs simt worst {worst(lp_t, lp_s):.2f} tol"
        if ref is not None:
            msg += f"; tc vs fp64 {worst(lp_t, ref):.2f} tol, simt vs fp64 {worst(lp_s, ref):.2f} tol"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.inverse(flow.relabel.to_engine(x.cuda()), cg, None, want_lp=True); e1.record(); torch.cuda.synchronize()
        msg += f"; {e0.elapsed_time(e1):.3f} ms, watchdog {eng.get_option('watchdog')}, opts {eng.get_option('inv_kernel_in_use')}/{eng.get_option('inv_a_tmem_in_use')}"
        print(msg)
