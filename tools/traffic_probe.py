"""Dev tool: one full-shape log_prob launch (cfg3 flow, broadcast context) for DRAM-traffic measurements under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`.  usage: traffic_probe.py S N draws_per_group gate   (gate = the inv_gate option: 2 = distance 1, 3 = distance 2, 1 = auto;
the r2 logs under gpurun_out/ were taken when 1 meant distance 2 and 2 meant distance 1)"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from naz_b200.flows.flow import NormalizingFlow
S, N, dpg, gate = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda")
gen = torch.Generator(device=dev); gen.manual_seed(0)
torch.manual_seed(0)
fl = NormalizingFlow("nsa", None, 4, 2, [150] * 3, 16, 8).to(dev)
draws = [[(lin.weight.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((S,) + tuple(lin.weight.shape), device=dev, generator=gen) * 2 - 1)),
           lin.bias.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((S,) + tuple(lin.bias.shape), device=dev, generator=gen) * 2 - 1)))
          for lin in arn.layers] for arn in fl.nets]
eng = fl.make_engine(draws, device=dev)
del draws
eng.set_option("inv_gate", gate)
x = torch.randn((N, 4), device=dev, generator=gen) * 1.5
ctx = torch.rand((2,), device=dev, generator=gen)
G = max(1, S // dpg)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.inverse(x, ctx, None, want_lp=False, want_lse=True, n_groups=G); e1.record()
    torch.cuda.synchronize()
print(f"S={S} N={N} draws/group={dpg} groups={G} gate={gate}: {e0.elapsed_time(e1):.1f} ms  {S*N/e0.elapsed_time(e1)/1e3:.2f} Mevals/s  packed {eng.packed_bytes/1e9:.2f} GB")
