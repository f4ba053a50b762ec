"""Dev tool: A/B timing of the tcgen05 inverse (log_prob) kernel variants at a config-3-like dev size.
usage: python tools/inv_ab.py [cfg] [S] [N]   (bench.py is the contract; this is for kernel iteration)"""
import json, os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
cases = {
 "cfg3": ("nsa", 4, 2, [150]*3, 16),
 "cfg4": ("maf", 2, 2, [150]*3, 16),
 "cfg2": ("maf", 6, 4, [150]*3, 16),
 "cfg5a": ("maf", 8, 4, [150]*3, 16),
 "cfg5b": ("maf", 16, 4, [150]*3, 16),
 "cfg5c": ("nsa", 8, 4, [150]*3, 16),
}
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 128 * 4
variants = json.loads(os.environ.get("VARIANTS", "null")) or [
    {"inv_kernel": 3}, {}, {"inv_fold": 0}, {"inv_merge_n": 64}, {"inv_merge_n": 96}, {"inv_merge_n": 136}, {"inv_merge_n": 256},
]
ctx_modes = os.environ.get("CTX", "bcast,point").split(",")
kind, D, C, hidden, L = cases[which]
spec, draws, keep, rng = make_case(kind, D, C, hidden, L, S, seed=1)
x = torch.from_numpy((rng.normal(size=(N, D)) * 1.5).astype(np.float32)).cuda()
ctxs = {"bcast": torch.from_numpy(rng.uniform(size=(1, C)).astype(np.float32)).cuda(),
        "point": torch.from_numpy(rng.uniform(size=(N, C)).astype(np.float32)).cuda()}
base = None
for opt in variants:
    eng = engine_for(spec, draws, engine="tcgen05", options=opt)
    for cm in ctx_modes:
        ctx = ctxs[cm]
        def run():
            return eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=int(os.environ.get("GROUPS", "1")))
        out = run(); torch.cuda.synchronize()
        times = []
        for _ in range(int(os.environ.get("REPS", "3"))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        ev = S * N / (ms * 1e-3)
        chk = float(out["lse_max"].double().mean())
        print(f"{which} {cm:6s} {json.dumps(opt):28s} S={S} N={N}: {ms:9.2f} ms  {ev/1e6:8.2f} Mevals/s  frac={eng.shape.flops_per_eval()*ev/1407.4e12:.3f}  mean(lse_max)={chk:.6f}  wd={eng.get_option('watchdog')}", flush=True)
    del eng
