"""Probe: does the tensor-core engine accept the single-degree masked-conditioner form of a coupling flow ('nsc')?"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from naz_b200.flows import NormalizingFlow
torch.manual_seed(1)
D, C, s = 4, 2, 2
flow = NormalizingFlow("nsc", None, D, C, [150, 150, 150], 16, 8, s).cuda().eval()
x = (torch.randn(20000, D) * 1.3).cuda(); c = torch.rand(20000, C).cuda(); z = torch.randn(20000, D).cuda()
lp_s = flow.log_prob(x, condition=c); xs_s = flow.sample(condition=c, base_noise=z)
flow._engine_kind = "auto"; flow._eng1 = None
try:
    lp_t = flow.log_prob(x, condition=c); xs_t = flow.sample(condition=c, base_noise=z)
    e = flow._single_engine()
    print("engines:", e.engine_for("inverse"), e.engine_for("forward"), e.options())
    print("lp  max |tc - simt| / (1e-5 + 1e-4 |.|):", float(((lp_t - lp_s).abs() / (1e-5 + 1e-4 * lp_s.abs())).max()))
    print("x   max |tc - simt| / (1e-5 + 1e-4 |.|):", float(((xs_t - xs_s).abs() / (2e-5 + 1e-4 * xs_s.abs())).max()))
    for name, f in (("simt", "simt"), ("auto", "auto")):
        flow._engine_kind = f; flow._eng1 = None
        flow.log_prob(x, condition=c); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); flow.log_prob(x, condition=c); e1.record(); torch.cuda.synchronize()
        print(name, "log_prob ms", e0.elapsed_time(e1))
except Exception as ex:
    print("tc engine refused:", repr(ex))
