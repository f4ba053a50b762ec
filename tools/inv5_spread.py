"""Dev tool: per-warp publish / accumulator-ready times of one flow layer of the v5 inverse kernel (CTA 0, all 16 epilogue
warps + the issuer): how far apart do the warps of one phase finish?  usage: python tools/inv5_spread.py [layer_index]"""
import ctypes as C, json, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
from naz_b200 import _lib
which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec, draws, keep, rng = make_case("nsa", 4, 2, [150]*3, 16, 4, seed=1)
N = 148 * 128
x = torch.from_numpy((rng.normal(size=(N, 4)) * 1.5).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, 2)).astype(np.float32)).cuda()
eng = engine_for(spec, draws, engine="tcgen05", options=json.loads(os.environ.get("OPTS", "{}")))
L = _lib.lib()
NEV, NS = 4096, 17
buf = torch.zeros(NS * NEV * 2, dtype=torch.int64, device="cuda")
run = lambda: eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=1)
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer.argtypes = [C.c_void_p]
L.nazb_debug_set_all_warps(1)
L.nazb_debug_set_clock_buffer(buf.data_ptr())
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer(None); L.nazb_debug_set_all_warps(0)
t = buf.cpu().numpy().reshape(NS, NEV, 2)
ev0 = [(int(c), int(v) >> 8, int(v) & 255) for c, v in t[0] if c > 0]
ends = [c for c, st, ev in ev0 if ev == 6]
lo, hi = ends[which], ends[which + 1]
print("layer window", hi - lo)
names = {2: "acc_ok", 3: "ld_ok", 4: "first_go", 5: "spline_done"}
for k in range(8): names[8 + k] = f"pub{k}"
# per (step, event): time per warp
table = {}
for w in range(16):
    for c, v in t[w]:
        c = int(c)
        if lo < c <= hi:
            st, ev = int(v) >> 8, int(v) & 255
            if ev in names: table.setdefault((st, names[ev]), {})[w] = c - lo
iss = [(int(c) - lo, int(v) >> 8, int(v) & 255) for c, v in t[16] if lo < int(c) <= hi]
for (st, ev), d in sorted(table.items(), key=lambda kv: min(kv[1].values())):
    ws = sorted(d.items())
    vals = [v for _, v in ws]
    print(f"step {st:3d} {ev:11s} min {min(vals):6d} max {max(vals):6d} spread {max(vals)-min(vals):5d}  by warp: " + " ".join(f"{w}:{v - min(vals)}" for w, v in ws))
print("issuer:", " ".join(f"{c}:{st}/{ev}" for c, st, ev in iss))
