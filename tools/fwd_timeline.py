"""Dev tool: clock stamps of CTA 0 of the two-tile forward kernel (flow_tc_fwd4_kernel<true>)."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
from naz_b200 import _lib
spec, draws, keep, rng = make_case("nsa", 4, 2, [150] * 3, 16, 2, seed=1)
N = 148 * 256
z = torch.from_numpy(rng.normal(size=(N, 4)).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, 2)).astype(np.float32)).cuda()
eng = engine_for(spec, draws, engine="tcgen05")
L = _lib.lib()
buf = torch.zeros(128 * 16, dtype=torch.int64, device="cuda")
eng.forward(z, ctx); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer.argtypes = [C.c_void_p]
L.nazb_debug_set_clock_buffer(buf.data_ptr())
eng.forward(z, ctx); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer(None)
t = buf.cpu().numpy().reshape(128, 16)
t0 = t[0, 0]
names = ["XF ", "L2 ", "L3 ", "OUT"]
print("gemm | tile0 epi: acc_wait first_slice total | tile1 epi: acc_wait first_slice total | issuer T0: wait_slices issue | T1: wait_slices issue | abs start T0.issue, T1.issue")
for i in range(12, 24):
    r = t[i]
    def epi(o):
        a, b, c, d = r[o:o + 4]
        return f"{b - a:6d} {max(c - b, 0):6d} {d - a:6d}"
    print(f"{i:3d} {names[i % 4]} | {epi(0)} | {epi(4)} | {r[9]-r[8]:6d} {r[10]-r[9]:6d} | {r[12]-r[11]:6d} {r[13]-r[12]:6d} | {r[9]-t0:8d} {r[12]-t0:8d}")
print("one layer (gemm 12 -> 16):", t[16, 9] - t[12, 9], "cycles")
