"""`NormalizingFlow` — drop-in for src/naz/flows/flow.py:24-129 on the libnazb CUDA path, plus the
draw-batched entry points (`log_prob_draws`, `sample_draws`) that replace the reference's per-draw
Python loops (train_flows.py:414-420, calibrate.py:147-150, compute_bic_simpler.py:116-120)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ..engine import FlowEngine, FlowShape
from .transforms import masked_affine_autoregressive, neural_spline_autoregressive, neural_spline_coupling


def _continuous_free_form(*a, **k):
    raise NotImplementedError("'cnf' (FFJORD) is a continuous flow; BASELINE.json's north_star covers discrete flows only")


flow_makers = {"maf": masked_affine_autoregressive, "nsa": neural_spline_autoregressive,
               "nsc": neural_spline_coupling, "cnf": _continuous_free_form}


class _FlowDist:
    """Stand-in for pyro's (Conditional)TransformedDistribution: the reference only touches
    `.transforms` (train_flows.py:38,65) and `.condition(...)` (flow.py:76)."""

    def __init__(self, base_dist, transforms):
        self.base_dist = base_dist
        self.transforms = transforms

    def condition(self, context):
        return self

    def clear_cache(self):
        """train_flows.py:220 calls it after every epoch; nothing is cached between calls here."""


def draws_from_posterior_samples(posterior_samples: Dict[str, torch.Tensor], L, n_lin: int):
    """Reference format `"flow_{i}_{name}" -> tensor[S, ...]` (train_flows.py:71, bflow.py:72) -> pytree.  `L` is the number
    of flow layers or, when Permute / BatchNorm transforms sit between them, the positions of the autoregressive
    transforms in `flow_dist.transforms` (the reference's `i` enumerates every transform)."""
    pos = range(L) if isinstance(L, int) else L
    return [[(posterior_samples[f"flow_{i}_nn.layers.{j}.weight"], posterior_samples[f"flow_{i}_nn.layers.{j}.bias"])
             for j in range(n_lin)] for i in pos]


class Relabelling:
    """`T.Permute` layers (transforms.py:155-156) cost nothing at run time: a permutation between two autoregressive layers is
    a re-labelling of the variables, so it is folded into the conditioner weights when they are packed.  With q the running
    map "actual index k of the reference <-> engine index q[k]" (identity before the first layer), an autoregressive layer
    becomes one on engine variables by moving its input columns (`W0'[:, C + q[k]] = W0[:, C + k]`), its output rows
    (`Wout'[m D + q[k]] = Wout[m D + k]`) and its MADE order (`perm' = q[perm]`); a Permute p then maps q <- q[p].  The
    eval-mode BatchNorm scale / shift behind layer l are re-labelled the same way.  At the ends: log_prob feeds
    `x[..., argsort(q_final)]` (and bounds in that order), sample returns `e[..., q_final]`."""

    def __init__(self, transforms, D: int):
        q = torch.arange(D)
        self.D = D
        self.ar_pos, self.q, self.bn = [], [], []
        for i, t in enumerate(transforms):
            kind = getattr(t, "kind", None)
            if kind in ("maf", "nsa", "nsc"):
                self.ar_pos.append(i); self.q.append(q.clone()); self.bn.append(None)
            elif kind == "permute":
                q = q[t.permutation.cpu()]
            elif kind == "batchnorm":
                if not self.ar_pos or self.bn[-1] is not None:
                    raise NotImplementedError("BatchNorm is supported directly behind an autoregressive layer (transforms.py:157)")
                self.bn[-1] = (t, q.clone())
            else:
                raise NotImplementedError(f"transform {type(t).__name__} is not part of naz's discrete flows")
        self.q_final = q
        self.inv_final = torch.argsort(q)
        self.trivial = all(torch.equal(qi, torch.arange(D)) for qi in self.q + [q])
        self.has_bn = any(b is not None for b in self.bn)

    def fold_layer(self, l: int, lins, C: int):
        """lins: [n_lin] of tensors or (W, b) pairs of flow layer l whose LAST two axes are [out, in] / last axis [out]
        (weights, masks alike; leading draw axes pass through)."""
        if self.trivial:
            return lins
        D, invq = self.D, torch.argsort(self.q[l])
        out = list(lins)

        def cols(W):
            idx = torch.cat([torch.arange(C), C + invq]).to(W.device)
            return W.index_select(-1, idx)

        def rows(T, axis):
            M = T.shape[axis] // D
            idx = torch.cat([m * D + invq for m in range(M)]).to(T.device)
            return T.index_select(axis, idx)

        first, last = out[0], out[-1]
        out[0] = (cols(first[0]), first[1]) if isinstance(first, tuple) else cols(first)
        last = out[-1]
        out[-1] = (rows(last[0], -2), rows(last[1], -1)) if isinstance(last, tuple) else rows(last, -2)
        return out

    def unfold_layer_grads(self, l: int, gW_first, gW_last, gb_last, C: int):
        """Gradients with respect to the FOLDED first / last linear of flow layer l -> the reference's index order."""
        if self.trivial:
            return gW_first, gW_last, gb_last
        D, q = self.D, self.q[l]
        cols = torch.cat([torch.arange(C), C + q]).to(gW_first.device)          # folded column C + q[k] holds reference column C + k
        M = gW_last.shape[-2] // D
        rows = torch.cat([m * D + q for m in range(M)]).to(gW_last.device)
        return gW_first.index_select(-1, cols), gW_last.index_select(-2, rows), gb_last.index_select(-1, rows)

    def fold_perm(self, l: int, perm):
        return self.q[l][perm.cpu().to(torch.int64)]

    def layer_affine(self):
        """(a, b) [L, D] in engine order, identity rows where a layer has no BatchNorm."""
        L = len(self.ar_pos)
        a, b = torch.ones(L, self.D), torch.zeros(L, self.D)
        for l, ent in enumerate(self.bn):
            if ent is not None:
                bn, q = ent
                al, bl = bn.affine()
                inv = torch.argsort(q)
                a[l], b[l] = al.cpu()[inv], bl.cpu()[inv]
        return a, b

    def to_engine(self, x):
        return x if self.trivial else x.index_select(-1, self.inv_final.to(x.device))

    def from_engine(self, e):
        return e if self.trivial else e.index_select(-1, self.q_final.to(e.device))

    def bounds_to_engine(self, bounds):
        if bounds is None or self.trivial:
            return bounds
        lo = torch.as_tensor(bounds["low"] if isinstance(bounds, dict) else bounds[0]).reshape(-1)
        hi = torch.as_tensor(bounds["high"] if isinstance(bounds, dict) else bounds[1]).reshape(-1)
        if lo.numel() == 1:
            return {"low": lo, "high": hi}
        return {"low": lo[self.inv_final.to(lo.device)], "high": hi[self.inv_final.to(hi.device)]}


class _LogProbFn(torch.autograd.Function):
    """Autograd node of `NormalizingFlow.log_prob`.  Minibatches (< `BIG` points): value and gradient come from ONE fp32-engine
    handle (a single re-pack per optimiser step; the backward pass, 6-7x the value pass, dominates a training step anyway).
    Large batches: the value comes from the module's forward engine (tensor cores), so an evaluation that merely forgot
    `torch.no_grad()` keeps its speed.  `backward` is ONE launch of
    nazb_inverse_vjp (the incremental inverse recomputed, then the adjoint recursion; naz_b200/csrc/flow_grad.cu) — so the
    reference's MLE loop `loss = -flow.log_prob(x, condition=y).mean(); loss.backward(); optimizer.step()`
    (train_flows.py:195-213) runs on this path unchanged."""

    BIG = 16384

    @staticmethod
    def forward(ctx, flow, x, cond, *params):
        eng = flow._grad_engine() if x.shape[0] < _LogProbFn.BIG else flow._single_engine()
        lp = eng.inverse(flow.relabel.to_engine(x), cond, flow._bounds_e(), want_lp=True)["lp"][0]
        ctx.flow, ctx.cond = flow, cond
        ctx.versions = tuple(p._version for p in params)
        ctx.params = params
        ctx.save_for_backward(x)
        return lp

    @staticmethod
    def backward(ctx, g):
        flow, (x,) = ctx.flow, ctx.saved_tensors
        if tuple(p._version for p in ctx.params) != ctx.versions:
            raise RuntimeError("NormalizingFlow parameters were modified in place between log_prob and backward")
        eng = flow._grad_engine()
        want_dc = bool(ctx.needs_input_grad[2])
        r = eng.inverse_grad(flow.relabel.to_engine(x), ctx.cond, flow._bounds_e(), want_dx=ctx.needs_input_grad[1],
                             weights=g.detach().reshape(-1), want_dctx=want_dc)
        grads = []
        for l in range(len(flow.nets)):
            gW, gb = [t[0] for t in r["gW"][l]], [t[0] for t in r["gb"][l]]
            gW[0], gW[-1], gb[-1] = flow.relabel.unfold_layer_grads(l, gW[0], gW[-1], gb[-1], flow.shape.C)
            if flow.flow_type == "nsc":
                grads += flow._layers[l].grads_from_made(gW, gb, flow.shape.C)
            else:
                for a, b in zip(gW, gb):
                    grads += [a, b]
        dx = flow.relabel.from_engine(r["dx"][0]).to(x.device).reshape(x.shape) if ctx.needs_input_grad[1] else None
        dc = None
        if want_dc:
            dc = r["dctx"][0]                                   # [N, C]; a broadcast context ([C] or [1, C]) receives the sum over points
            if ctx.cond.dim() == 1 or ctx.cond.shape[0] == 1:
                dc = dc.sum(0).reshape(ctx.cond.shape)
            dc = dc.to(ctx.cond.device)
        return (None, dx, dc) + tuple(grads)


class NormalizingFlow(nn.Module):
    def __init__(self, flow_type, bounds, *flow_maker_args, embedding_net=None, engine: str = "auto", **flow_maker_kwargs):
        super().__init__()
        assert flow_type in list(flow_makers.keys())
        flow_maker = flow_makers[flow_type]
        self.flow_type = flow_type
        self.conditional = True if flow_maker_args[1] > 0 else False
        if embedding_net is not None:
            assert self.conditional
            self.embedding_net = embedding_net
        else:
            self.embedding_net = nn.Identity()
        self.bounds = bounds
        self.theta_dim, self.condition_dim = int(flow_maker_args[0]), int(flow_maker_args[1])
        self.flow, self.transforms, self.nets = flow_maker(*flow_maker_args, **flow_maker_kwargs)
        self.base_dist = torch.distributions.Normal(torch.zeros(self.theta_dim), torch.ones(self.theta_dim))
        self.flow_dist = _FlowDist(self.base_dist, self.transforms)
        self.relabel = Relabelling(self.transforms, self.theta_dim)
        hidden = self.nets[0].hidden_dims
        order = flow_maker_kwargs.get("order", "quadratic")
        count_bins = 8
        if flow_type == "nsa":
            count_bins = flow_maker_kwargs.get("count_bins", flow_maker_args[4] if len(flow_maker_args) > 4 else 8)
        if flow_type == "nsc":
            count_bins = flow_maker_kwargs.get("count_bins", flow_maker_args[4] if len(flow_maker_args) > 4 else 8)
        kind = "maf" if flow_type == "maf" else ("nsa" if order == "quadratic" else "nsa_linear")
        self.shape = FlowShape(kind, self.theta_dim, max(self.condition_dim, 0), list(hidden), len(self.nets), count_bins)
        self.dropout_p = flow_maker_kwargs.get("dropout_p", None)
        # coupling layers reach the engine as single-degree masked conditioners (transforms.py::SplineCoupling): the tensor-core
        # forward (sample) programs take them as they are, the tensor-core inverse programs are built for the MADE degree ladder
        # and decline, so `log_prob` of these flows is served by the fp32 kernel (one conditioner pass per layer there)
        self._engine_kind = engine
        self._layers = [self.transforms[i] for i in self.relabel.ar_pos]
        self._eng1: Optional[FlowEngine] = None
        self._eng1_key = None
        self._engg: Optional[FlowEngine] = None
        self._engg_key = None

    # ------------------------------------------------------------------ packing helpers
    def masks(self):
        if self.flow_type == "nsc":
            dev = next(self.parameters()).device
            return [[m.to(dev) for m in t.made_masks(self.shape.C)] for t in self._layers]
        return [[lin.mask for lin in arn.layers] for arn in self.nets]

    def perms(self):
        if self.flow_type == "nsc":
            return torch.arange(self.theta_dim).repeat(len(self.nets), 1)
        return torch.stack([arn.permutation.cpu() for arn in self.nets])

    def current_draw(self):
        """[L][n_lin](W, b) in the engine's conditioner format (for coupling flows: the single-degree masked form)."""
        if self.flow_type == "nsc":
            return [t.as_made([(lin.weight.detach(), lin.bias.detach()) for lin in t.nn.layers],
                              [g.detach() for g in t.lower_spline.groups(t.order)], self.shape.C) for t in self._layers]
        return [[(lin.weight.detach(), lin.bias.detach()) for lin in arn.layers] for arn in self.nets]

    def _packed_masks(self):
        """Masks / MADE orders as the engine packs them, i.e. with the Permute layers folded in (`Relabelling`)."""
        return [self.relabel.fold_layer(l, ml, self.condition_dim) for l, ml in enumerate(self.masks())]

    def _packed_perms(self):
        return torch.stack([self.relabel.fold_perm(l, perm) for l, perm in enumerate(self.perms())])

    def _fold_draws(self, draws):
        if self.relabel.trivial:
            return draws
        return [self.relabel.fold_layer(l, [(torch.as_tensor(W), torch.as_tensor(b)) for (W, b) in layer], self.condition_dim)
                for l, layer in enumerate(draws)]

    def _new_engine(self, S: int, dev) -> FlowEngine:
        eng = FlowEngine(self.shape, S, device=dev, engine=self._engine_kind)
        if self.relabel.has_bn:
            if self.training:
                raise RuntimeError("BatchNorm flows are evaluated with their moving statistics: call flow.eval() first (the "
                                   "train()-mode batch statistics belong to the training loop, which is out of scope)")
            eng.set_layer_affine(*self.relabel.layer_affine())
        return eng

    def _device(self):
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("naz_b200.NormalizingFlow evaluates on CUDA only (no CPU fallback): call flow.cuda() first")
        return p.device

    def _single_engine(self) -> FlowEngine:
        """Engine holding the module's current parameters (S = 1), re-packed when they change."""
        dev = self._device()
        key = (dev, self.training) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._eng1 is None or self._eng1_key != key:
            if self._eng1 is None or self._eng1.device != dev or self.relabel.has_bn:
                self._eng1 = self._new_engine(1, dev)
            self._eng1.pack(self._fold_draws(self.current_draw()), self._packed_masks(), self._packed_perms())
            self._eng1_key = key
        return self._eng1

    def _grad_engine(self) -> FlowEngine:
        """fp32-engine handle (S = 1) on the current parameters for nazb_inverse_grad / nazb_inverse_vjp."""
        dev = self._device()
        key = (dev,) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._engg is None or self._engg_key != key:
            if self._engg is None or self._engg.device != dev or self.relabel.has_bn:
                self._engg = FlowEngine(self.shape, 1, device=dev, engine="simt")
                if self.relabel.has_bn:                     # eval-mode statistics and affine, constants of the gradient
                    self._engg.set_layer_affine(*self.relabel.layer_affine())
            self._engg.pack(self._fold_draws(self.current_draw()), self._packed_masks(), self._packed_perms())
            self._engg_key = key
        return self._engg

    def _flat_params(self):
        """Parameters in the order `_LogProbFn.backward` returns their gradients."""
        if self.flow_type == "nsc":
            return [p for t in self._layers for p in ([q for lin in t.nn.layers for q in (lin.weight, lin.bias)]
                                                      + t.lower_spline.groups(t.order))]
        return [t for arn in self.nets for lin in arn.layers for t in (lin.weight, lin.bias)]

    def make_engine(self, draws, keep=None, p_drop: float = 0.0, device=None) -> FlowEngine:
        """Engine holding S draws given as the reference pytree `[L][n_lin](W[S,out,in], b[S,out])`
        or as the `"flow_{i}_{name}"` dict (coupling flows: the dict, or a pytree already in the engine's conditioner format)."""
        if isinstance(draws, dict):
            post = draws
            draws = draws_from_posterior_samples(post, self.relabel.ar_pos, len(self.nets[0].layers))
            if self.flow_type == "nsc":
                names = ["unnormalized_widths", "unnormalized_heights", "unnormalized_derivatives", "unnormalized_lambdas"]
                draws = [t.as_made(layer, [post[f"flow_{i}_lower_spline.{n}"] for n in names[:len(t.slot_groups())]], self.shape.C)
                         for i, t, layer in zip(self.relabel.ar_pos, self._layers, draws)]
        draws = self._fold_draws(draws)
        S = 1
        for layer in draws:
            for (W, b) in layer:
                if W.dim() == 3:
                    S = W.shape[0]
        if keep is not None:
            S = keep.shape[0]
        eng = self._new_engine(S, device or self._device())
        eng.pack(draws, self._packed_masks(), self._packed_perms(), keep, p_drop)
        return eng

    def _bounds_e(self):
        return self.relabel.bounds_to_engine(self.bounds)

    def _cond(self, condition):
        if self.conditional:
            assert condition is not None
            return self.embedding_net(condition)
        return None

    # ------------------------------------------------------------------ reference API (flow.py:45-129)
    def log_prob(self, x, *args, condition=None, **kwargs):
        """flow.py:66-79.  Evaluated by libnazb on the module's CURRENT weights.  With autograd enabled and parameters (or x)
        that require grad the result carries a grad_fn whose backward is one nazb_inverse_vjp launch (`_LogProbFn`: masked-affine
        and quadratic-spline flows), so the reference's training loop (train_flows.py:195-213) works unchanged; under
        torch.no_grad() it is a plain tensor.  A dropout flow in train() mode
        would apply a fresh nn.Dropout mask upstream: that stochastic path is `MCDPNormalizingFlow.sample_uncertain` /
        `log_prob_draws(keep=...)` here, so calling log_prob in that state raises instead of silently ignoring dropout."""
        if self.training and self.dropout_p not in (None, 0.0):
            raise RuntimeError("log_prob on a dropout flow in train() mode: call flow.eval() for the deterministic density, or "
                               "log_prob_draws(..., keep=masks, p_drop=p) for explicit MC-dropout masks")
        params = self._flat_params()
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or (isinstance(x, torch.Tensor) and x.requires_grad)
                                        or (isinstance(condition, torch.Tensor) and condition.requires_grad)):
            cond = self._cond(condition)          # may carry the graph of a trainable embedding_net (flow.py:30-36): d lp / d ctx flows back
            if self.relabel.has_bn and self.training:
                raise RuntimeError("BatchNorm flows are evaluated with their moving statistics: call flow.eval() first (the "
                                   "train()-mode batch statistics belong to the training loop, which is out of scope)")
            # BatchNorm (eval mode): its statistics, gamma and beta are constants of this node (they receive no gradient)
            return _LogProbFn.apply(self, x, cond, *params)
        eng = self._single_engine()
        out = eng.inverse(self.relabel.to_engine(x), self._cond(condition), self._bounds_e(), want_lp=True)
        return out["lp"][0]

    def bounded_log_prob(self, x, *args, condition=None, **kwargs):
        if self.bounds is None:
            return self.log_prob(x, *args, condition=condition, **kwargs)
        lp = torch.full_like(x[:, 0], -math.inf)
        low, high = self.bounds["low"].to(x.device), self.bounds["high"].to(x.device)
        valid = ((x > low.expand(x.shape)) & (x < high.expand(x.shape))).all(dim=1)
        if valid.any():
            cond = condition
            if condition is not None and condition.dim() == 2 and condition.shape[0] == x.shape[0]:
                cond = condition[valid]
            lp[valid] = self.log_prob(x[valid, :], *args, condition=cond, **kwargs)
        return lp

    def average_log_prob(self, x, *args, condition=None, **kwargs):
        return torch.mean(self.bounded_log_prob(x, *args, condition=condition, **kwargs))

    def sample(self, *args, condition=None, base_noise=None, **kwargs):
        """`flow.sample([N], condition=c)`; `predict` upstream passes `(cond, [N])` positionally
        (train_flows.py:419) — both spellings are accepted."""
        shape = None
        for a in args:
            if isinstance(a, torch.Tensor) and condition is None and self.conditional:
                condition = a
            elif isinstance(a, (list, tuple, torch.Size)):
                shape = list(a)
        eng = self._single_engine()
        if base_noise is None:
            assert shape is not None, "sample shape required"
            n = int(torch.Size(shape).numel())
            base_noise = torch.randn((n, self.theta_dim), device=eng.device)
        x = self.relabel.from_engine(eng.forward(base_noise.reshape(-1, self.theta_dim), self._cond(condition), self._bounds_e())[0])
        return x.reshape(*(shape if shape is not None else [x.shape[0]]), self.theta_dim)

    # ------------------------------------------------------------------ draw-batched entry points
    def log_prob_draws(self, x, draws, condition=None, reduce: Optional[str] = None, log_w=None, keep=None,
                       p_drop: float = 0.0, engine: Optional[FlowEngine] = None):
        """log p(x_n | theta_s) for all draws.  reduce=None -> [S,N]; "lse" -> posterior predictive
        log (1/S) sum_s p(x_n|theta_s) [N] (or weighted by log_w); "sum" -> sum_n lp[s,n] [S] (float64)."""
        eng = engine or self.make_engine(draws, keep, p_drop)
        cond = self._cond(condition)
        x, bounds = self.relabel.to_engine(torch.as_tensor(x)), self._bounds_e()
        if reduce is None:
            return eng.inverse(x, cond, bounds, want_lp=True)["lp"]
        if reduce == "lse":
            out = eng.inverse(x, cond, bounds, want_lp=False, want_lse=True, log_w=log_w)
            return eng.lse_finish(out["lse_max"], out["lse_sum"], 0.0 if log_w is not None else -math.log(eng.S))
        if reduce == "sum":
            return eng.inverse(x, cond, bounds, want_lp=False, want_sum=True)["sum_n"]
        raise ValueError(reduce)

    def sample_draws(self, draws, n_or_noise, condition=None, keep=None, p_drop: float = 0.0,
                     engine: Optional[FlowEngine] = None, generator=None):
        """x[s, n, :] ~ p(. | theta_s).  `n_or_noise` is a sample count (fresh N(0,I) noise per draw) or
        the base noise itself, [S,N,D] or [N,D] (shared)."""
        eng = engine or self.make_engine(draws, keep, p_drop)
        if isinstance(n_or_noise, int):
            z = torch.randn((eng.S, n_or_noise, self.theta_dim), device=eng.device, generator=generator)
        else:
            z = n_or_noise
        return self.relabel.from_engine(eng.forward(z, self._cond(condition), self._bounds_e()))
