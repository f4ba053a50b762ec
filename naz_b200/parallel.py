"""Sharding by weight draw across the GPUs of one box: rank g owns a contiguous slice
of the draws and its packed weights only; points are replicated.  Exactly one collective per call:
an all-gather of the per-rank (max, sum-exp) partials for the per-point posterior predictive, or of the
per-draw sums for importance weights.  Raw [S,N] / [S,N,D] outputs stay sharded.

The gradient path shards the other way: the chains are few and every chain sums over the whole data set,
so rank g takes a contiguous slice of the POINTS, every rank holds all chains, and an all-reduce(SUM) of the flattened
gradient buffer (plus the [S] values) finishes the step (the reference's jax.grad over `jnp.sum(lp)`, bflow_jax_maf.py:233-235)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(S: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first S % world ranks get one extra draw."""
    base, rem = divmod(S, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def combine_lse_partials(pmax: torch.Tensor, psum: torch.Tensor, log_norm: float) -> torch.Tensor:
    """[G,N] partial (max, sum exp) -> log sum.  Pure-torch combine used on CPU (gloo tests); on CUDA the
    product path calls libnazb's nazb_lse_finish (engine.lse_finish)."""
    m = pmax.max(dim=0).values
    safe = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    s = (psum * torch.exp(pmax - safe)).sum(dim=0)
    return safe + torch.log(s) + log_norm


def all_gather_lse(pmax: torch.Tensor, psum: torch.Tensor, S_total: int, log_w_given: bool = False,
                   group=None) -> torch.Tensor:
    """Each rank passes its local [N] (or [g,N]) partials; returns the global posterior predictive [N]."""
    if pmax.dim() == 1:
        pmax, psum = pmax.unsqueeze(0), psum.unsqueeze(0)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if pmax.shape[0] > 1 and world > 1:
        # several draw groups per rank: fold them into ONE (max, sum) pair first so the collective stays [2,1,N]
        if pmax.is_cuda:
            from .engine import lse_finish
            loc = lse_finish(pmax.contiguous(), psum.contiguous(), 0.0)
        else:
            loc = combine_lse_partials(pmax, psum, 0.0)
        pmax, psum = loc.unsqueeze(0), torch.ones_like(loc).unsqueeze(0)
    local = torch.stack([pmax, psum])                      # [2, g, N]
    if world > 1:
        buf = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(buf, local, group=group)
        allp = torch.cat(buf, dim=1)                       # [2, world*g, N]
    else:
        allp = local
    log_norm = 0.0 if log_w_given else -math.log(S_total)
    if allp.is_cuda:
        from .engine import lse_finish
        return lse_finish(allp[0].contiguous(), allp[1].contiguous(), log_norm)
    return combine_lse_partials(allp[0], allp[1], log_norm)


def all_gather_draw_sums(sum_local: torch.Tensor, S_total: int, group=None) -> torch.Tensor:
    """Per-draw sums from every rank, in global draw order (shard_range order). Handles ragged shards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return sum_local
    cap = -(-S_total // world)
    pad = torch.zeros(cap, dtype=sum_local.dtype, device=sum_local.device)
    pad[: sum_local.numel()] = sum_local
    buf = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(buf, pad, group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(S_total, r, world)
        parts.append(buf[r][: e - b])
    return torch.cat(parts)


def point_range(N: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice of the points for the gradient path."""
    return shard_range(N, rank, world)


def all_reduce_value_and_grads(sum_n: torch.Tensor, gW, gb, group=None):
    """Sum the per-rank partial values [S] (float64) and gradients ([L][n_lin] of [S,out,in] / [S,out], fp32) over the
    ranks: one all-reduce of the flat fp32 gradient buffer plus one of the [S] values (kept in float64); returns
    (sum_n, gW, gb) holding the totals, the gradients as views into the reduced buffer."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return sum_n, gW, gb
    leaves = [t for layer in gW for t in layer] + [t for layer in gb for t in layer]
    flat = torch.cat([t.reshape(-1) for t in leaves])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    val = sum_n.clone()
    dist.all_reduce(val, op=dist.ReduceOp.SUM, group=group)
    views, off = [], 0
    for t in leaves:
        views.append(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    it = iter(views)
    gW2 = [[next(it) for _ in layer] for layer in gW]
    gb2 = [[next(it) for _ in layer] for layer in gb]
    return val, gW2, gb2


def inverse_grad_point_sharded(eng, x: torch.Tensor, ctx: Optional[torch.Tensor] = None, bounds=None, group=None):
    """FlowEngine.inverse_grad over this rank's slice of the points, then the all-reduce.  x (and a per-point ctx) are the
    FULL arrays, identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = point_range(x.shape[0], rank, world)
    c = ctx
    if ctx is not None and ctx.dim() == 2 and ctx.shape[0] == x.shape[0]:
        c = ctx[b:e]
    r = eng.inverse_grad(x[b:e], c, bounds)
    val, gW, gb = all_reduce_value_and_grads(r["sum_n"], r["gW"], r["gb"], group)
    return {"sum_n": val, "gW": gW, "gb": gb}
