"""Dev tool: a short tcgen05 inverse (log_prob) run at the config-3 shape, for ncu captures."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import make_case, engine_for
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128
spec, draws, keep, rng = make_case("nsa", 4, 2, [150] * 3, 16, S, seed=1)
x = torch.from_numpy((rng.normal(size=(N, 4)) * 1.5).astype(np.float32)).cuda()
ctx = torch.from_numpy(rng.uniform(size=(1, 2)).astype(np.float32)).cuda()
G = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = engine_for(spec, draws, engine="tcgen05")
for _ in range(2):
    out = eng.inverse(x, ctx, want_lp=False, want_lse=True, n_groups=G)
torch.cuda.synchronize()
print("ok", float(out["lse_max"].float().mean()))
