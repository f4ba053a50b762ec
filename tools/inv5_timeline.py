"""Dev tool: event timeline of one flow layer of the v5 inverse kernel (CTA 0: epilogue slice lane 0 / 1 of rows 0-15 of
quadrant 0 and the issuer).  usage: python tools/inv4_timeline.py [bcast|point] [layer_index]"""
import ctypes as C, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
from naz_b200 import _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "bcast"
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec, draws, keep, rng = make_case("nsa", 4, 2, [150]*3, 16, 4, seed=1)
N = 148 * 128
x = torch.from_numpy((rng.normal(size=(N, 4)) * 1.5).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, 2) if mode == "bcast" else (N, 2)).astype(np.float32)).cuda()
import json, os
eng = engine_for(spec, draws, engine="tcgen05", options=json.loads(os.environ.get("OPTS", "{}")))
L = _lib.lib()
NEV = 4096
buf = torch.zeros(3 * NEV * 2, dtype=torch.int64, device="cuda")
run = lambda: eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=1)
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer.argtypes = [C.c_void_p]
L.nazb_debug_set_clock_buffer(buf.data_ptr())
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer(None)
t = buf.cpu().numpy().reshape(3, NEV, 2)
names = {1: "step_begin", 2: "acc_ok", 3: "tmem_ld_ok", 4: "first_go(pair ok)", 5: "spline_done", 6: "LAYER_END",
         20: "i:step_begin", 21: "i:weights_ok", 22: "i:commit_acc", 23: "i:all_issued"}
for k in range(8): names[8 + k] = f"pub_slice{k}"; names[32 + k] = f"i:slice{k}_ok"; names[48 + k] = f"i:slice{k}_issued"
ev0 = [(int(c), int(v) >> 8, int(v) & 255) for c, v in t[0] if c > 0]
ends = [c for c, st, ev in ev0 if ev == 6]
if len(ends) < which + 2: print("not enough events", len(ev0), len(ends)); sys.exit(1)
lo, hi = ends[which], ends[which + 1]
print(f"{mode}: layer window {hi - lo} cycles; consecutive layer lengths:", [ends[i + 1] - ends[i] for i in range(min(len(ends) - 1, 10))])
allev = []
for slot in range(3):
    for c, v in t[slot]:
        c = int(c)
        if lo < c <= hi: allev.append((c, slot, int(v) >> 8, int(v) & 255))
allev.sort()
prev = lo
lastslot = {0: lo, 1: lo, 2: lo}
print(" t(cyc)  +d_any  +d_own  who   step event")
for c, slot, st, ev in allev:
    who = ["epi0", "epi1", "issr"][slot]
    print(f"{c - lo:7d} {c - prev:7d} {c - lastslot[slot]:7d}  {who}  {st:3d}  {names.get(ev, ev)}")
    prev = c; lastslot[slot] = c
