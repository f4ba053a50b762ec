"""Dev tool: per-step clock stamps of CTA 0 of the tcgen05 kernel (handoff latency breakdown)."""
import ctypes as C, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
from naz_b200 import _lib
direction = sys.argv[1] if len(sys.argv) > 1 else "inverse"
spec, draws, keep, rng = make_case("nsa", 4, 2, [150]*3, 16, 2, seed=1)
N = 148*128
x = torch.from_numpy((rng.normal(size=(N, 4)) * 1.5).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, 2)).astype(np.float32)).cuda()
eng = engine_for(spec, draws, engine="tcgen05")
L = _lib.lib()
prog = (C.c_int * (8*64))()
L.nazb_debug_program.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
nsteps = L.nazb_debug_program(eng._h, 0 if direction == "inverse" else 1, prog, 8*64)
buf = torch.zeros(256*16, dtype=torch.int64, device="cuda")
def run():
    if direction == "inverse": eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=1)
    else: eng.forward(x, ctx)
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer.argtypes = [C.c_void_p]
L.nazb_debug_set_clock_buffer(buf.data_ptr())
run(); torch.cuda.synchronize()
L.nazb_debug_set_clock_buffer(None)
t = buf.cpu().numpy()[:2048].reshape(256, 8) if direction != "inverse" else buf.cpu().numpy().reshape(256, 16)
t0 = t[0, 0]
names = {0: "none", 1: "tanh", 2: "xinv", 3: "xfwd"}
print("step  n ks sp epi ncols wKB | v2 inverse: wait_w issue acc_wait | epi_work sync | total")
names = {0: "none", 1: "tanh", 2: "xinv", 3: "xfwd", 4: "first"}
if direction == "inverse":
    # v3 stamps: epilogue warp 0 [0] step start [1] acc barrier passed [2] TMEM loaded [3] step end;
    #            issuer of chain 0 [8] step start [9] weights landed [10] first A slice landed [11] all MMAs issued
    print("step  n ncrit ks epi ncols wKB | epi: acc_wait  ld   work  total | issuer: w_wait a_wait issue | issue_end - epi_end(prev)")
    for i in range(min(2*nsteps, 127)):
        s = i % nsteps
        n, ks, sp, epi, ncols, wb, dcol, abuf = [prog[s*8+j] for j in range(8)]
        e0, e1, e2, e3 = t[i][:4]
        m8, m9, m10, m11 = t[i][8:12]
        line = f"{i:3d} {n:4d} {ks:2d} {names[epi]:5s} {ncols:3d} {wb/1024:5.1f} |"
        if epi:
            line += f" {e1-e0:6d} {max(e2-e1,0):5d} {e3-max(e2,e1):6d} {e3-e0:6d} |"
            if epi == 1:
                e4, e5, e6 = t[i][4:7]
                line += f" half0: tanh {e4-e2:5d} sts {e5-e4:4d} fence+arrive {e6-e5:4d} |"
        else:
            line += " " * 29 + "|"
        if wb:
            line += f" {m9-m8:6d} {m10-m9:6d} {m11-m10:6d}"
        print(line)
    print("two layers:", t[2*nsteps-1, 3] - t[0, 0], "cycles")
    sys.exit(0)
print("step  n ks sp epi ncols wKB | mma: wait_a  wait_w  issue | epi: wait_acc(after mma issue)  work  signal | step_total")
for i in range(min(2*nsteps, 255)):
    s = i % nsteps
    n, ks, sp, epi, ncols, wb, dcol, abuf = [prog[s*8+j] for j in range(8)]
    m0, m1, m2, m3, e4, e5, e6, e7 = t[i]
    nxt = t[i+1, 0] if i+1 < 256 else m3
    line = f"{i:3d} {n:4d} {ks:2d} {sp:1d} {names[epi]:5s} {ncols:3d} {wb/1024:5.1f} | {m1-m0:7d} {m2-m1:7d} {m3-m2:7d} |"
    if epi:
        line += f" {e5-m3:7d} {e6-e5:7d} {e7-e6:7d} | {e7-m0:7d}  ld_done@{e4-e5 if epi==1 else 0:5d}"
    print(line)
tot = t[2*nsteps-1, 7] - t[0, 0] if nsteps*2 <= 255 else 0
print("two layers:", tot, "cycles")
