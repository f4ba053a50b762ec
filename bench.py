#!/usr/bin/env python
"""bench.py — headline benchmark of the draw-batched flow-evaluation path (BASELINE.json).

Workload (config.workload = "cfg3"): Bayesian-flow posterior-predictive log_prob — conditional
quadratic-spline autoregressive flow, D=4 | C=2, hidden [150,150,150], 16 flow layers, K=8 bins,
S = 1000 weight draws x N = 1,000,000 points, output log (1/S) sum_s p(x_n | theta_s)  [N]
(SURVEY.md §8(d) row 3).  One "step" = one pass of that whole job.  Draws are sharded across ranks
(strong scaling: the job is fixed, each of G ranks owns S/G draws) with ONE all-gather of the per-rank
(max, sum-exp) partials per step.

  python bench.py --gpus N --steps K --warmup W            # our arm (libnazb CUDA path)
  python bench.py --impl reference --steps K --warmup W    # reference CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` = whole-job log-prob evals/s with inputs resident in HBM;
`e2e` = the same through the public API with pinned-host inputs (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (kind, D, C, hidden, L, K, S, N)
    "cfg3": ("nsa", 4, 2, [150, 150, 150], 16, 8, 1000, 1_000_000),
    "cfg4": ("maf", 2, 2, [150, 150, 150], 16, 8, 256, 1_000_000),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--draws", type=int, default=None, help="override S (dev only; the judged run uses the default)")
    ap.add_argument("--points", type=int, default=None, help="override N (dev only)")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--groups", type=int, default=0,
                    help="draw groups per rank (0 = auto: ~8 draws per group so a group's packed weights stay L2-resident "
                         "while the CTAs sweep their point tiles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary workloads (configs 2, 4, 5)")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity probe against the fp64 oracle")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_config(name, cfg, S, N):
    """The workload definition: identical keys AND values in both arms (the driver compares the two `config` dicts);
    everything specific to one arm lives under `detail`."""
    kind, D, C, hidden, L, K, _, _ = cfg
    return {"workload": name, "flow": kind, "D": D, "C": C, "hidden": list(hidden), "layers": L, "count_bins": K, "draws": S, "points": N,
            "output": "logsumexp_s(lp) - log S  [N]", "context": "one [C] vector broadcast to all points (calibrate.py:85,126)",
            "l2": "GPU arm: the packed weights of the draws (GBs) are streamed every step and exceed the 126 MB L2; CPU arm: not applicable"}


def flops_per_eval(D, C, hidden, L, M):
    dims = [D + C] + list(hidden) + [M * D]
    return 2 * L * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# reference CPU path (oracle port): torch fp32 on all host cores, the reference's own structure
# --------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, seconds: float, threads=None, device=None):
    """Times oracle/pyro_style.py (the reference's torch path restated; pyro is not installable here) on a
    bounded sample of the workload: on the host cores (device=None: the CPU baseline / reference arm) or, with a CUDA
    device, through stock eager PyTorch on the GPU (the "reference GPU path" of BASELINE.md §3).
    Returns (evals_per_s, cores, sample_description)."""
    import numpy as np
    import torch
    from oracle import flow_oracle as fo
    from oracle import pyro_style as ps
    kind, D, C, hidden, L, K, S, N = cfg
    cores = threads or os.cpu_count()
    torch.set_num_threads(cores)
    rng = np.random.default_rng(2)
    perms = np.stack([rng.permutation(D) for _ in range(L)])
    spec = fo.FlowSpec(kind, D, C, hidden, L, perms, count_bins=K)
    p0 = fo.init_weights(spec, rng, np.float32)
    flow = ps.PyroStyleFlow(kind, None, D, C, hidden, L, K, "quadratic", perms)
    n_draws = 2
    draws = fo.perturb_draws(p0, n_draws, 0.25, rng, np.float32)
    tdraws = [[(torch.from_numpy(W), torch.from_numpy(b)) for (W, b) in layer] for layer in draws]
    ctx = torch.from_numpy(rng.uniform(size=(C,)).astype(np.float32)) if C else None
    on_gpu = device is not None
    if on_gpu:
        flow = flow.to(device)
        flow.base_dist = torch.distributions.Normal(torch.zeros(D, device=device), torch.ones(D, device=device))
        tdraws = [[(W.to(device), b.to(device)) for (W, b) in layer] for layer in tdraws]
        ctx = None if ctx is None else ctx.to(device)

    def run(npts):
        x = torch.from_numpy((rng.normal(size=(npts, D)) * 1.5).astype(np.float32))
        if on_gpu:
            x = x.to(device)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ps.log_prob_draws_reference_loop(flow, tdraws, x, ctx)
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(512)                                   # warm-up
    probe_n = 4096
    t = run(probe_n)
    rate = n_draws * probe_n / t
    npts = int(max(probe_n, min(2_000_000 if on_gpu else 400_000, rate * seconds / n_draws)))
    t = run(npts)
    cpu_reference_rate.last_seconds = t
    where = "eager PyTorch on the B200 (oracle port; upstream ships no Blackwell kernel)" if on_gpu else f"fp32 torch CPU ({cpu_model()})"
    return n_draws * npts / t, cores, f"{n_draws} draws x {npts} points of the {L}-layer {kind} D={D}|C={C} flow, {where}, {t:.1f} s"


def pyro_found():
    try:
        import pyro  # noqa: F401
        return True
    except Exception:
        return False


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    kind, D, C, hidden, L, K, S, N = cfg
    S = args.draws or S
    N = args.points or N
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        rate, cores, sample = cpu_reference_rate(cfg, max(2.0, args.cpu_seconds / 2))
        if i >= args.warmup:
            vals.append(rate)
            secs.append(cpu_reference_rate.last_seconds)
    v = sum(vals) / len(vals)
    out = {
        "impl": "reference", "metric": "log-prob evals/s (weight-draws x points)", "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, cfg, S, N),
        "detail": {"note": "reference CPU path = oracle/pyro_style.py (pyro-ppl is not installable here), bounded sample per step",
                   "pyro_found": pyro_found()},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "cpu_model": cpu_model(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def parity_probe(flow, probe_draws, x, ctx, eng, kind, D, C, hidden, L, K):
    """2 draws x 4096 points of the timed configuration (same weights, same inputs) against the fp64 oracle (checker only).
    viol = fraction outside |cuda - ref| <= 1e-5 + 1e-4 |ref|; worst = largest error / tolerance."""
    import numpy as np
    from oracle import flow_oracle as fo
    spec = fo.FlowSpec(kind, D, C, list(hidden), L, flow.perms().numpy(), count_bins=K)
    out = eng.inverse(x, ctx, None, want_lp=True, s_begin=0, s_count=2)["lp"].cpu().numpy().astype(np.float64)
    d64 = [[(W.astype(np.float64), b.astype(np.float64)) for (W, b) in lay] for lay in probe_draws]
    ref, _ = fo.log_prob_draws(spec, d64, x.cpu().numpy().astype(np.float64), None if ctx is None else ctx.cpu().numpy().astype(np.float64))
    err = np.abs(out - ref)
    tol = 1e-5 + 1e-4 * np.abs(ref)
    return {"viol": float((err > tol).mean()), "worst": float((err / tol).max()), "max_abs_err": float(err.max()),
            "checked": "2 draws x %d points of the timed weights / inputs vs the fp64 oracle (oracle/flow_oracle.py)" % x.shape[0],
            "tolerance": "1e-4 relative + 1e-5 absolute"}


def aux_workloads(dev, gen, peak_tf, engine):
    """SURVEY §8(d) configs 2, 4 and the log_prob leg of config 5 at full size (kernel time, CUDA events); every entry carries its
    own algorithmic F1 and roofline fraction.  Not part of `value`."""
    import torch
    from naz_b200.flows.flow import NormalizingFlow
    specs = {
        # name: (kind, D, C, hidden, L, K, S, N, per-point context, dropout masks)
        "aux_cfg2": ("maf", 6, 4, [150] * 3, 16, 8, 100, 100_000, True, True),
        "aux_cfg4": ("maf", 2, 2, [150] * 3, 16, 8, 256, 1_000_000, True, False),
        "aux_cfg5_logprob_maf_8d": ("maf", 8, 4, [150] * 3, 16, 8, 1000, 10_000, False, False),
        "aux_cfg5_logprob_maf_16d": ("maf", 16, 4, [150] * 3, 16, 8, 1000, 10_000, False, False),
        "aux_cfg5_logprob_nsa_8d": ("nsa", 8, 4, [150] * 3, 16, 8, 1000, 10_000, False, False),
    }
    res = {}
    for name, (kind, D, C, hidden, L, K, S, N, per_point, dropout) in specs.items():
        try:
            torch.manual_seed(5)
            if kind == "nsa":
                fl = NormalizingFlow("nsa", None, D, C, hidden, L, K, engine=engine).to(dev)
            else:
                fl = NormalizingFlow("maf", None, D, C, hidden, L, engine=engine).to(dev)
            x = torch.randn((N, D), device=dev, generator=gen) * 1.5
            ctx = torch.rand((N, C) if per_point else (C,), device=dev, generator=gen)
            if dropout:
                keep = (torch.rand((S, L, len(hidden), max(hidden)), device=dev, generator=gen) > 0.25).float()
                eng = fl.make_engine(fl.current_draw(), keep=keep, p_drop=0.25, device=dev)
            else:
                draws = [[(lin.weight.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((S,) + tuple(lin.weight.shape), device=dev, generator=gen) * 2 - 1)),
                           lin.bias.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((S,) + tuple(lin.bias.shape), device=dev, generator=gen) * 2 - 1)))
                          for lin in arn.layers] for arn in fl.nets]
                eng = fl.make_engine(draws, device=dev)
                del draws
            run = lambda: eng.inverse(x, ctx, None, want_lp=False, want_lse=True, want_sum=True)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            f1 = flops_per_eval(D, C, hidden, L, fl.shape.M)
            ev = float(S) * N / (ms * 1e-3)
            res[name] = {"value": ev, "unit": "evals/s", "ms": ms, "draws": S, "points": N, "flow": f"{kind} {D}|{C} {hidden} x{L}",
                         "context": "per point" if per_point else "one vector", "engine": eng.engine_for("inverse"),
                         "engine_options": eng.options(), "flops_per_eval": f1, "roofline_frac": f1 * ev / 1e12 / peak_tf,
                         "finite": bool(torch.isfinite(o["sum_n"]).all().item())}
            del eng, o, x, ctx, fl
            torch.cuda.empty_cache()
        except Exception as e:
            res[name] = {"unavailable": repr(e)}
    return res


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from naz_b200 import FlowEngine, FlowShape, _lib
    from naz_b200.flows.flow import NormalizingFlow
    from naz_b200.parallel import all_gather_lse, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kind, D, C, hidden, L, K, S, N = CONFIGS[args.config]
    S = args.draws or S
    N = args.points or N
    s_begin, s_end = shard_range(S, rank, world)
    S_loc = s_end - s_begin

    # ---- synthetic model: the public flow object (random-init weights of the named architecture) ----
    torch.manual_seed(2)
    if kind == "nsa":
        flow = NormalizingFlow("nsa", None, D, C, hidden, L, K, engine=args.engine).to(dev)
    else:
        flow = NormalizingFlow("maf", None, D, C, hidden, L, engine=args.engine).to(dev)
    M = flow.shape.M
    # posterior draws theta_s = theta_0 (1 + 0.25 u_s), u ~ U(-1,1) (bflow_jax_maf.py:239-240), generated on device per rank
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    draws = []
    for arn in flow.nets:
        lay = []
        for lin in arn.layers:
            W, b = lin.weight.detach(), lin.bias.detach()
            uW = torch.rand((S_loc,) + tuple(W.shape), device=dev, generator=gen) * 2 - 1
            ub = torch.rand((S_loc,) + tuple(b.shape), device=dev, generator=gen) * 2 - 1
            lay.append((W.unsqueeze(0) * (1 + 0.25 * uW), b.unsqueeze(0) * (1 + 0.25 * ub)))
        draws.append(lay)
    # the first two draws stay on the host for the post-run parity probe against the fp64 oracle
    probe_draws = [[(W[:2].cpu().numpy().copy(), b[:2].cpu().numpy().copy()) for (W, b) in lay] for lay in draws] if rank == 0 else None
    t0 = time.perf_counter()
    eng = flow.make_engine(draws, device=dev)
    torch.cuda.synchronize()
    pack_s = time.perf_counter() - t0
    del draws
    torch.cuda.empty_cache()

    # ---- synthetic points (same on every rank), pinned host copies for the e2e leg ----
    g2 = torch.Generator().manual_seed(7)
    x_host = (torch.randn((N, D), generator=g2) * 1.5).pin_memory()
    c_host = torch.rand((C,), generator=g2).pin_memory() if C else None
    x_dev = x_host.to(dev)
    c_dev = c_host.to(dev) if C else None
    out_host = torch.empty((N,), dtype=torch.float32).pin_memory()

    kernel_ms = []
    n_groups = args.groups if args.groups > 0 else max(1, S_loc // 8)

    def step_device():
        """One pass of the job with inputs resident in HBM."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.inverse(x_dev, c_dev, None, want_lp=False, want_lse=True, n_groups=n_groups)
        e1.record()
        kernel_ms.append((e0, e1))
        return all_gather_lse(out["lse_max"], out["lse_sum"], S)

    def step_e2e():
        """The public API call: pinned-host points in, posterior-predictive log-density out on the host."""
        xd = x_host.to(dev, non_blocking=True)
        cd = c_host.to(dev, non_blocking=True) if C else None
        out = eng.inverse(xd, cd, None, want_lp=False, want_lse=True, n_groups=n_groups)
        res = all_gather_lse(out["lse_max"], out["lse_sum"], S)
        out_host.copy_(res, non_blocking=True)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    kernel_ms.clear()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = step_device()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    total_ms = ev0.elapsed_time(ev1)
    k_ms = sum(a.elapsed_time(b) for a, b in kernel_ms) / max(1, len(kernel_ms))

    # e2e leg: one warm-up, then the same number of steps
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    # auxiliary: the sampling direction (reference `sample`, one conditioner pass per flow layer) on the same draws,
    # cfg-5-sized: every local draw x 10 000 base-noise points shared by the draws (kernel time, CUDA events)
    n_s = 10_000
    z_dev = torch.randn((n_s, D), device=dev, generator=gen)
    eng.forward(z_dev, c_dev)
    torch.cuda.synchronize()
    f0, f1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    xs = eng.forward(z_dev, c_dev)
    f1e.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1e)
    del xs

    # auxiliary: SURVEY §8 f1 — value + gradient of sum_n lp (the NUTS / SVI inner loop), config-4 architecture (maf 2|2),
    # 4 chains x 100 000 points; 1 grad-eval = one (chain, point) pair through the value and the full parameter gradient
    aux_grad = aux_grad_spline = None
    if not args.no_aux:
        # every rank holds the same chains (same seeds) and the same points; with N > 1 GPUs the POINTS are sharded
        # (strong scaling of one gradient step) and one all-reduce of the flat gradient buffer + the values finishes the step
        # (naz_b200.parallel.inverse_grad_point_sharded); the time is the max over ranks of the CUDA-event time of the whole
        # step, all-reduce included
        from naz_b200.parallel import inverse_grad_point_sharded

        def grad_aux(gkind, gD, gC, gK, chains, pts_per_gpu, what):
            torch.manual_seed(3)
            ggen = torch.Generator(device=dev)
            ggen.manual_seed(99)
            gargs = (gD, gC, [150, 150, 150], 16) + ((gK,) if gkind == "nsa" else ())
            gflow = NormalizingFlow(gkind, None, *gargs, engine="simt").to(dev)
            gdraws = [[(lin.weight.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((chains,) + tuple(lin.weight.shape), device=dev, generator=ggen) * 2 - 1)),
                        lin.bias.detach().unsqueeze(0) * (1 + 0.25 * (torch.rand((chains,) + tuple(lin.bias.shape), device=dev, generator=ggen) * 2 - 1)))
                       for lin in arn.layers] for arn in gflow.nets]
            geng = gflow.make_engine(gdraws, device=dev)
            GN = pts_per_gpu * world                            # per-GPU work fixed
            gx = torch.randn((GN, gD), device=dev, generator=ggen) * 1.5
            gc = torch.rand((GN, gC), device=dev, generator=ggen)
            inverse_grad_point_sharded(geng, gx, gc)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            gr = inverse_grad_point_sharded(geng, gx, gc)
            g1.record()
            torch.cuda.synchronize()
            g_t = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(g_t, op=dist.ReduceOp.MAX)
            g_ms = float(g_t.item())
            # algorithmic work of one grad-eval: the value pass (F1, masks as dense zeros) + the two products of the backward pass
            # per linear layer (cotangent . W and the outer product into dW) = 3 F1; roofline = fp32 FMA issue of the CUDA cores
            # (nominal: SMs x 128 lanes x 2 flop x measured SM clock; MEASURED_PEAKS.json holds no fp32 figure)
            g_f1 = 3 * flops_per_eval(gD, gC, [150, 150, 150], 16, 2 if gkind == "maf" else 3 * gK - 1)
            g_peak = world * torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 2 * 1.965e9 / 1e12
            g_ach = g_f1 * chains * GN / (g_ms * 1e-3) / 1e12
            return {"value": chains * GN / (g_ms * 1e-3), "unit": "grad-evals/s", "chains": chains, "points": GN, "ms": g_ms,
                    "flops_per_grad_eval": g_f1,
                    "roofline": {"bound": "fp32 FMA (CUDA cores)", "achieved": g_ach, "peak": g_peak, "unit": "TFLOP/s", "frac": g_ach / g_peak,
                                 "peak_source": "nominal: SMs x 128 FMA lanes x 2 x 1.965 GHz, all GPUs of the run"},
                    "flow": what, "engine": "simt (fp32 CUDA cores)",
                    "parallelism": f"point-sharded x{world}, one all-reduce of the gradient buffer (weak scaling: {pts_per_gpu} points per GPU)",
                    "finite": bool(torch.isfinite(gr["sum_n"]).all().item()),
                    "note": "nazb_inverse_grad: value + d/d(all weights) of sum_n lp per chain (row f1; not part of `value`)"}

        # config-4 architecture (maf 2|2), 4 chains x 100 000 points; 1 grad-eval = one (chain, point) pair through the value and
        # the full parameter gradient;  and the benchmarked spline architecture (config 3: nsa 4|2, K = 8), 2 chains x 50 000 points
        aux_grad = grad_aux("maf", 2, 2, 8, 4, 100_000, "maf 2|2 [150,150,150] x16")
        aux_grad_spline = grad_aux("nsa", 4, 2, 8, 2, 50_000, "nsa 4|2 [150,150,150] x16, K = 8 (quadratic)")

    # parity probe: the weights and inputs that were just timed, 2 draws x 4096 points, against the fp64 oracle
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_probe(flow, probe_draws, x_dev[:4096], c_dev, eng, kind, D, C, hidden, L, K)
    # auxiliary workloads: each with its own F1-based roofline fraction
    aux = {}
    if world == 1 and not args.no_aux:
        peaks0, _ = load_peaks()
        aux = aux_workloads(dev, gen, float(peaks0.get("bf16_tflops_sustained", peaks0.get("bf16_tflops"))), args.engine)
    ref_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            rate_g, _, sample_g = cpu_reference_rate(CONFIGS[args.config], min(args.cpu_seconds, 6.0), device=dev)
            ref_gpu = {"value": rate_g, "unit": "evals/s", "kind": "port", "sample": sample_g}
        except Exception as e:   # the oracle port is test infrastructure: never let it take the bench line down
            ref_gpu = {"unavailable": repr(e)}

    tms = torch.tensor([total_ms, e2e_s * 1e3, k_ms, fwd_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, k_ms, fwd_ms = [float(v) for v in tms.tolist()]
    finite = bool(torch.isfinite(res).all().item())

    if rank == 0:
        evals = float(S) * float(N)
        ms_per_step = total_ms / args.steps
        value = evals / (ms_per_step * 1e-3)
        e2e_value = evals / (e2e_ms * 1e-3 / args.steps)
        peaks, peak_kind = load_peaks()
        f1 = flops_per_eval(D, C, hidden, L, M)
        # dominant kernel: flow_tc_kernel / flow_simt_kernel, one launch per step per rank
        ach_tf = f1 * float(S_loc) * float(N) / (k_ms * 1e-3) / 1e12
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get(args.config)
            except Exception:
                traffic = None
        out = {
            "metric": "log-prob evals/s (weight-draws x points)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 hi/lo-split tcgen05 MMAs, fp32 TMEM accumulate)"
            if eng.engine_for("inverse") == "tcgen05" else "f32",
            "data": "synthetic",
            "config": workload_config(args.config, CONFIGS[args.config], S, N),
            "detail": {"draws_per_gpu": S_loc,
                       "parallelism": f"draw-sharded x{world}, one all-gather of the per-rank log-sum-exp partials [N]",
                       "draw_groups_per_gpu": n_groups, "engine": eng.engine_for("inverse"), "engine_options": eng.options(),
                       "packed_weights_gb": eng.packed_bytes / 1e9,
                       "l2": ("inputs larger than L2: %.1f GB of packed weights streamed per step" % (eng.packed_bytes / 1e9))
                       if eng.packed_bytes > 126e6 else "dev-size run: packed weights fit L2 (not a judged configuration)",
                       "pack_seconds_excluded": pack_s, "result_finite": finite, "pyro_found": pyro_found()},
            "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         "traffic": traffic, "peak_source": f"{peak_kind} bf16 dense, sustained",
                         "note": "achieved = algorithmic F1 (%d flop/eval, masks as dense zeros, never the D-pass count, never reduced by the context fold) x evals per launch / CUDA-event time of the launch (fold kernel + flow_tc_inv5_kernel); 3 fp16 MMAs per algorithmic product (M = 128, A operand in TMEM); bound by the dependency chain of the autoregressive inverse and the epilogue instruction issue (DESIGN.md 5.2)" % f1},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + (C * 4 if C else 0)),
                    "d2h_bytes_per_step": int(N * 4)},
            "gpu_launches": int(launches),
            "clocks": clk,
            "aux_sample_direction": {"value": float(S) * n_s / (fwd_ms * 1e-3), "unit": "samples/s", "draws": S, "points_per_draw": n_s,
                                     "engine": eng.engine_for("forward"), "ms": fwd_ms,
                                     "note": "reference `sample` direction on the same draws (not part of `value`)"},
        }
        if aux_grad is not None:
            out["aux_grad_direction"] = aux_grad
        if aux_grad_spline is not None:
            out["aux_grad_direction_spline"] = aux_grad_spline
        out.update(aux)
        if parity is not None:
            out["parity"] = parity
        if ref_gpu is not None:
            out["reference_gpu"] = ref_gpu
        if not args.no_cpu_baseline and world == 1:
            rate, cores, sample = cpu_reference_rate(CONFIGS[args.config], args.cpu_seconds)
            out["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "cpu_model": cpu_model(), "kind": "port", "sample": sample}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
