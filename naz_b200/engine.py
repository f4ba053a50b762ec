"""FlowEngine — thin PyTorch-side owner of one libnazb handle.

PyTorch is plumbing only here (device memory, streams, `torch.distributed`); every flow FLOP runs in
the hand-written CUDA behind the C ABI (include/nazb.h).  The engine ingests the reference's own weight
formats: the `torch_to_jax` pytree `[L][n_lin](W[S,out,in], b[S,out])` + masks + permutations
(src/naz/flows/bflow_jax_maf.py:26-46), and optional per-draw dropout keep-masks.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_KINDS = {"maf": _lib.KIND_AFFINE, "nsa": _lib.KIND_RQS, "nsa_linear": _lib.KIND_RLS}
_ENGINES = {"auto": _lib.ENGINE_AUTO, "simt": _lib.ENGINE_SIMT, "tcgen05": _lib.ENGINE_TCGEN05}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def hidden_degrees_from_masks(masks: Sequence[torch.Tensor], perm: torch.Tensor, D: int, C: int) -> Optional[List[List[int]]]:
    """Recover MADE degrees of the hidden units of ONE flow layer from its masks, or None when the
    masks are not of the canonical degree form (then only the Jacobi inverse schedule is valid).
    Degrees follow pyro's convention (bflow_jax_maf.py:57-62): inputs [0]*C ++ (1+rank)."""
    perm = perm.to(torch.int64).cpu()
    rank = torch.empty(D, dtype=torch.int64)
    rank[perm] = torch.arange(D)
    in_deg = torch.cat([torch.zeros(C, dtype=torch.int64), 1 + rank])
    degs: List[torch.Tensor] = []
    prev = in_deg
    for k, m in enumerate(masks[:-1]):
        m = m.detach().cpu() > 0.5
        # unit degree = largest degree it is connected to (hid >= in  <=>  connected)
        d = torch.where(m, prev[None, :].expand_as(m), torch.full_like(m, -1, dtype=torch.int64)).max(dim=1).values
        d = d.clamp(min=0 if C > 0 else 1)
        if not torch.equal(m, d[:, None] >= prev[None, :]):
            return None
        degs.append(d)
        prev = d
    out_deg = (1 + rank).repeat(masks[-1].shape[0] // D)
    if not torch.equal(masks[-1].detach().cpu() > 0.5, out_deg[:, None] > prev[None, :]):
        return None
    for d in degs:
        if (d[1:] < d[:-1]).any() or d.max() >= D:
            return None
    return [d.tolist() for d in degs]


@dataclass
class FlowShape:
    kind: str
    D: int
    C: int
    hidden: List[int]
    L: int
    count_bins: int = 8
    bound: float = 3.0
    clip: Tuple[float, float] = (-5.0, 3.0)

    @property
    def M(self) -> int:
        K = self.count_bins
        return {"maf": 2, "nsa": 3 * K - 1, "nsa_linear": 4 * K - 1}[self.kind]

    def flops_per_eval(self) -> int:
        """Algorithmic F1 = 2 L [(D+C) H1 + sum H_k H_{k+1} + H_last D M]  (SURVEY §8(d))."""
        dims = [self.D + self.C] + list(self.hidden) + [self.M * self.D]
        return 2 * self.L * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


class FlowEngine:
    """One handle = one flow architecture x S weight draws resident on one GPU."""

    def __init__(self, shape: FlowShape, S: int, device=None, engine: str = "auto", inverse_mode: str = "incremental"):
        if not torch.cuda.is_available():
            raise RuntimeError("naz_b200.FlowEngine needs a CUDA (sm_100) device; there is no CPU fallback")
        self.shape = shape
        self.S = int(S)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.lib()
        d = _lib.NazbDesc()
        d.kind = _KINDS[shape.kind]
        d.D, d.C, d.L, d.n_hidden = shape.D, shape.C, shape.L, len(shape.hidden)
        if len(shape.hidden) > _lib.NAZB_MAX_HIDDEN_LAYERS:
            raise ValueError("too many hidden layers")
        for i, h in enumerate(shape.hidden):
            d.hidden[i] = h
        d.count_bins = shape.count_bins
        d.bound = shape.bound
        d.clip_lo, d.clip_hi = shape.clip
        d.S = self.S
        d.engine = _ENGINES[engine]
        d.inverse_mode = _lib.INV_INCREMENTAL if inverse_mode == "incremental" else _lib.INV_JACOBI
        d.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = C.c_void_p()
        rc = self._lib.nazb_create(C.byref(self._h), C.byref(d))
        if rc != 0:
            self._h = None
            raise _lib.NazbError(rc, "nazb_create")
        self.inverse_mode = inverse_mode
        self._keepalive = None

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.nazb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def engine_name(self) -> str:
        return _lib.ENGINE_NAMES[self._lib.nazb_engine_in_use(self._h)]

    def engine_for(self, direction: str) -> str:
        """Engine serving `direction` ("inverse" = log_prob, "forward" = sample) after pack()."""
        return _lib.ENGINE_NAMES[self._lib.nazb_engine_for_direction(self._h, 0 if direction == "inverse" else 1)]

    @property
    def packed_bytes(self) -> int:
        return int(self._lib.nazb_packed_bytes(self._h))

    def _check(self, rc: int, where: str):
        if rc != 0:
            detail = self._lib.nazb_last_cuda_error(self._h).decode() if self._h else ""
            raise _lib.NazbError(rc, where, detail)

    _warned_shapes = set()

    def _warn_simt_directions(self):
        """`engine="auto"` serves a direction whose tensor-core program does not fit tensor memory with the fp32 CUDA-core
        kernel (~10x slower on the bench shapes): say so once per shape instead of degrading silently."""
        if self.engine_name != "tcgen05":
            return
        slow = [d for d in ("inverse", "forward") if self.engine_for(d) == "simt"]
        key = (self.shape.kind, self.shape.D, self.shape.C, tuple(self.shape.hidden), self.shape.count_bins, tuple(slow))
        if slow and key not in FlowEngine._warned_shapes:
            FlowEngine._warned_shapes.add(key)
            import warnings
            warnings.warn(f"naz_b200: the {' and '.join(slow)} direction of this {self.shape.kind} flow (D={self.shape.D}, C={self.shape.C}, "
                          f"hidden={list(self.shape.hidden)}) does not fit the tcgen05 programs (tensor-memory budget) or carries a per-layer "
                          "affine (BatchNorm), and runs on the fp32 SIMT kernel; FlowEngine.engine_for(direction) reports the engine per direction", RuntimeWarning, stacklevel=3)

    # engine options (include/nazb.h: nazb_set_option); nothing in the library reads the environment
    OPTION_NAMES = ("inv_kernel", "inv_merge_n", "inv_fold", "inv_gate", "inv_a_tmem", "inv_align", "inv_trim", "inv_defer", "inv_park",
                    "inv_gaps")

    def set_option(self, name: str, value: int) -> None:
        self._check(self._lib.nazb_set_option(self._h, name.encode(), int(value)), f"nazb_set_option({name})")

    def get_option(self, name: str) -> Optional[int]:
        v = C.c_int32(0)
        rc = self._lib.nazb_get_option(self._h, name.encode(), C.byref(v))
        return int(v.value) if rc == 0 else None

    def options(self) -> dict:
        """Current engine options (for bench / test records); {} on the SIMT engine."""
        out = {}
        for k in self.OPTION_NAMES + ("inv_fold_available", "inv_kernel_in_use", "inv_a_tmem_in_use", "inv_block_width"):
            v = self.get_option(k)
            if v is not None:
                out[k] = v
        return out

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_layer_affine(self, a: Optional[torch.Tensor], b: Optional[torch.Tensor] = None) -> None:
        """Element-wise affine behind every flow layer (eval-mode `T.BatchNorm`, transforms.py:157-158): sampling direction
        x <- a[l] * x + b[l] after layer l.  a, b: [L, D], a > 0; None removes the step.  Call before pack(): such flows are
        served by the SIMT kernel in both directions (include/nazb.h `nazb_set_layer_affine`)."""
        if a is None:
            self._check(self._lib.nazb_set_layer_affine(self._h, None, None, self._stream()), "nazb_set_layer_affine")
            return
        sh = self.shape
        a_h = torch.as_tensor(a).detach().to("cpu", torch.float32).contiguous()
        b_h = torch.as_tensor(b).detach().to("cpu", torch.float32).contiguous()
        if tuple(a_h.shape) != (sh.L, sh.D) or tuple(b_h.shape) != (sh.L, sh.D):
            raise ValueError(f"layer affine must be [L={sh.L}, D={sh.D}]")
        self._check(self._lib.nazb_set_layer_affine(self._h, a_h.data_ptr(), b_h.data_ptr(), self._stream()),
                    "nazb_set_layer_affine")

    # ------------------------------------------------------------------
    def pack_draw_map(self, base, standard_params: torch.Tensor, scale, masks, perms,
                      keep: Optional[torch.Tensor] = None, p_drop: float = 0.0):
        """Pack S draws given as the reference's STANDARD parameters (bflow_jax_maf.py:239-240):
        theta_s = theta_0 * (1 + scale * u_s), applied while packing (no [S, P] `params` array, no per-draw loop).
        base: [L][n_lin] of (W0 [out,in], b0 [out]); standard_params: [S, P] with P in `ravel_pytree` order of the params
        pytree (layer by layer, W then b of each linear, C order) — the layout of posterior["standard_params"]."""
        sh = self.shape
        dims = [sh.D + sh.C] + list(sh.hidden) + [sh.M * sh.D]
        u = _f32c(torch.as_tensor(standard_params), self.device)
        P = sum(dims[j + 1] * dims[j] + dims[j + 1] for j in range(len(dims) - 1)) * sh.L
        if u.dim() != 2 or u.shape[0] != self.S or u.shape[1] != P:
            raise ValueError(f"standard_params must be [S={self.S}, P={P}]")
        if isinstance(scale, torch.Tensor) and scale.numel() > 1:
            # fixed_scale=False of the reference's model (bflow_jax_maf.py:238): `scale` is itself sampled, one value per
            # draw [S] or per draw and parameter [S, P] (multi_scale=True).  The fused pack kernels take one scalar, so this
            # rarely used variant materialises theta = theta_0 * (1 + scale * u) on the device with the reference's
            # un-fused fp32 rounding order (mul, add, mul) and packs the result.
            sc = _f32c(scale, self.device)
            if tuple(sc.shape) not in ((self.S,), (self.S, 1), (self.S, P)):
                raise ValueError(f"scale must be a float, [S={self.S}] or [S={self.S}, P={P}]")
            sc = sc.reshape(self.S, -1)
            flat0 = torch.cat([torch.cat([_f32c(torch.as_tensor(W0), self.device).reshape(-1),
                                          _f32c(torch.as_tensor(b0), self.device).reshape(-1)]) for layer in base for (W0, b0) in layer])
            theta = flat0.unsqueeze(0) * (1.0 + sc * u)
            draws, off = [], 0
            for l in range(sh.L):
                lay = []
                for j in range(len(dims) - 1):
                    out, inn = dims[j + 1], dims[j]
                    W = theta[:, off:off + out * inn].reshape(self.S, out, inn); off += out * inn
                    bb = theta[:, off:off + out]; off += out
                    lay.append((W, bb))
                draws.append(lay)
            return self.pack(draws, masks, perms, keep, p_drop)
        views, off = [], 0
        for l in range(sh.L):
            lay = []
            for j in range(len(dims) - 1):
                out, inn = dims[j + 1], dims[j]
                uw = u[:, off:off + out * inn]; off += out * inn       # strided views into the flat matrix (stride P)
                ub = u[:, off:off + out]; off += out
                lay.append((uw, ub))
            views.append(lay)
        return self.pack(views, masks, perms, keep, p_drop, _base=base, _scale=float(scale), _u_stride=P)

    def pack(self, draws, masks, perms, keep: Optional[torch.Tensor] = None, p_drop: float = 0.0, *, _base=None,
             _scale: float = 0.0, _u_stride: int = 0):
        """draws: [L][n_lin] of (W, b); W is [S,out,in] or [out,in] (shared by all draws), same for b.
        masks: [L][n_lin] of [out,in] 0/1; perms: [L][D] int; keep: [S,L,n_hidden,max(hidden)] 0/1."""
        sh = self.shape
        L, n_lin = sh.L, len(sh.hidden) + 1
        if len(draws) != L or any(len(layer) != n_lin for layer in draws):
            raise ValueError("draws must be [L][n_hidden+1] of (W, b)")
        dev = self.device
        Wt, bt, mt, wst, bst = [], [], [], [], []
        dims = [sh.D + sh.C] + list(sh.hidden) + [sh.M * sh.D]
        for l in range(L):
            for j in range(n_lin):
                W, b = draws[l][j]
                m = _f32c(torch.as_tensor(masks[l][j]), dev)
                out, inn = dims[j + 1], dims[j]
                if _base is not None:
                    # W / b are [S, out*in] / [S, out] views into the flat standard-parameter matrix (draw stride _u_stride)
                    Wt.append(W); bt.append(b); mt.append(m)
                    wst.append(_u_stride); bst.append(_u_stride)
                    continue
                W = _f32c(torch.as_tensor(W), dev)
                b = _f32c(torch.as_tensor(b), dev)
                if W.dim() == 2:
                    W = W.unsqueeze(0)
                if b.dim() == 1:
                    b = b.unsqueeze(0)
                if W.shape[1:] != (out, inn) or b.shape[1:] != (out,) or m.shape != (out, inn):
                    raise ValueError(f"layer {l} linear {j}: expected W[*,{out},{inn}], got {tuple(W.shape)}")
                if W.shape[0] not in (1, self.S) or b.shape[0] not in (1, self.S):
                    raise ValueError("leading dimension of W / b must be S or absent")
                Wt.append(W); bt.append(b); mt.append(m)
                wst.append(0 if W.shape[0] == 1 else out * inn)
                bst.append(0 if b.shape[0] == 1 else out)
        perms_t = torch.as_tensor(perms).to(torch.int64).cpu().contiguous().reshape(L, sh.D)
        hid_deg_arr = None
        if self.inverse_mode == "incremental":
            hk = max(sh.hidden)
            ref = None
            for l in range(L):
                dg = hidden_degrees_from_masks([mt[l * n_lin + j] for j in range(n_lin)], perms_t[l], sh.D, sh.C)
                if dg is None:
                    raise ValueError("masks are not canonical MADE masks; use inverse_mode='jacobi'")
                if ref is None:
                    ref = dg
                elif dg != ref:
                    raise ValueError("hidden degrees differ between flow layers; use inverse_mode='jacobi'")
            hid_deg_arr = (C.c_int32 * (len(sh.hidden) * hk))()
            for j, dj in enumerate(ref):
                for u, v in enumerate(dj):
                    hid_deg_arr[j * hk + u] = int(v)
        keep_t = None
        if keep is not None:
            keep_t = _f32c(torch.as_tensor(keep), dev)
            if tuple(keep_t.shape) != (self.S, L, len(sh.hidden), max(sh.hidden)):
                raise ValueError("keep must be [S, L, n_hidden, max(hidden)]")
        n = L * n_lin
        VP = C.c_void_p * n
        I64 = C.c_int64 * n
        Wp = VP(*[t.data_ptr() for t in Wt])
        bp = VP(*[t.data_ptr() for t in bt])
        mp = VP(*[t.data_ptr() for t in mt])
        perm_arr = (C.c_int64 * (L * sh.D))(*perms_t.flatten().tolist())
        base_t = None
        if _base is not None:
            base_t = [(_f32c(torch.as_tensor(W0), dev), _f32c(torch.as_tensor(b0), dev)) for layer in _base for (W0, b0) in layer]
            for i, (W0, b0) in enumerate(base_t):
                out, inn = dims[i % n_lin + 1], dims[i % n_lin]
                if tuple(W0.shape) != (out, inn) or tuple(b0.shape) != (out,):
                    raise ValueError("base weights must be [out, in] / [out] per linear")
            W0p = VP(*[w.data_ptr() for (w, _) in base_t])
            b0p = VP(*[b_.data_ptr() for (_, b_) in base_t])
            rc = self._lib.nazb_pack_draw_map(self._h, W0p, b0p, Wp, bp, I64(*wst), I64(*bst), float(_scale), mp, perm_arr,
                                              hid_deg_arr, _ptr(keep_t), float(p_drop), self._stream())
            self._check(rc, "nazb_pack_draw_map")
        else:
            rc = self._lib.nazb_pack(self._h, Wp, bp, I64(*wst), I64(*bst), mp, perm_arr, hid_deg_arr,
                                     _ptr(keep_t), float(p_drop), self._stream())
            self._check(rc, "nazb_pack")
        # the pack kernels read the source tensors asynchronously on the current stream
        self._keepalive = (Wt, bt, mt, keep_t, base_t)
        self._warn_simt_directions()
        return self

    # ------------------------------------------------------------------
    def _prep_points(self, x, ctx, bounds):
        sh = self.shape
        x = _f32c(torch.as_tensor(x), self.device)
        c = None
        rows = 1
        if sh.C > 0:
            assert ctx is not None, "condition is required for a conditional flow (flow.py:75)"
            c = _f32c(torch.as_tensor(ctx), self.device)
            if c.dim() == 1:
                c = c.unsqueeze(0)
            rows = c.shape[0]
            if c.shape[-1] != sh.C:
                raise ValueError("condition has wrong width")
        lo = hi = None
        if bounds is not None:
            lo = _f32c(torch.as_tensor(bounds["low"] if isinstance(bounds, dict) else bounds[0]), self.device).reshape(-1)
            hi = _f32c(torch.as_tensor(bounds["high"] if isinstance(bounds, dict) else bounds[1]), self.device).reshape(-1)
            if lo.numel() == 1:
                lo = lo.expand(sh.D).contiguous()
            if hi.numel() == 1:
                hi = hi.expand(sh.D).contiguous()
            if lo.numel() != sh.D or hi.numel() != sh.D:      # the kernels read lo[d], hi[d] for every d < D
                raise ValueError(f"bounds must be scalars or have D = {sh.D} entries (got {lo.numel()} / {hi.numel()})")
        return x, c, rows, lo, hi

    def inverse(self, x, ctx=None, bounds=None, *, want_z=False, want_lp=True, want_lse=False, want_sum=False,
                log_w: Optional[torch.Tensor] = None, s_begin: int = 0, s_count: Optional[int] = None,
                n_groups: Optional[int] = None):
        """Reference `log_prob` direction for draws [s_begin, s_begin+s_count).  Returns a dict with the
        requested outputs: z [S,N,D], lp [S,N], lse (max,sum) partials [G,N], sum_n [S] (float64)."""
        sh = self.shape
        s_count = self.S - s_begin if s_count is None else s_count
        x, c, rows, lo, hi = self._prep_points(x, ctx, bounds)
        if x.dim() != 2 or x.shape[1] != sh.D:
            raise ValueError(f"x must be [N,{sh.D}]")
        N = x.shape[0]
        out = {}
        if N == 0:
            # an empty batch is legal upstream (pyro / jnp return empty arrays); nothing to launch
            if want_z:
                out["z"] = torch.empty((s_count, 0, sh.D), device=self.device, dtype=torch.float32)
            if want_lp:
                out["lp"] = torch.empty((s_count, 0), device=self.device, dtype=torch.float32)
            if want_lse:
                G = n_groups if n_groups else 1
                out["lse_max"] = torch.empty((G, 0), device=self.device, dtype=torch.float32)
                out["lse_sum"] = torch.empty((G, 0), device=self.device, dtype=torch.float32)
            if want_sum:
                out["sum_n"] = torch.zeros((s_count,), device=self.device, dtype=torch.float64)
            return out
        z = torch.empty((s_count, N, sh.D), device=self.device, dtype=torch.float32) if want_z else None
        lp = torch.empty((s_count, N), device=self.device, dtype=torch.float32) if want_lp else None
        lmax = lsum = None
        G = 0
        if want_lse:
            G = n_groups if n_groups else self._default_groups(N, s_count)
            lmax = torch.empty((G, N), device=self.device, dtype=torch.float32)
            lsum = torch.empty((G, N), device=self.device, dtype=torch.float32)
        sum_n = torch.zeros((s_count,), device=self.device, dtype=torch.float64) if want_sum else None
        lw = None if log_w is None else _f32c(torch.as_tensor(log_w), self.device).reshape(-1)
        if lw is not None and lw.numel() != s_count:          # local to [s_begin, s_begin + s_count): the kernel reads log_w[si]
            raise ValueError(f"log_w must have s_count = {s_count} entries (got {lw.numel()})")
        rc = self._lib.nazb_inverse(self._h, s_begin, s_count, x.data_ptr(), _ptr(c), rows, N, _ptr(lo), _ptr(hi),
                                    _ptr(z), _ptr(lp), _ptr(lw), _ptr(lmax), _ptr(lsum), G, _ptr(sum_n), self._stream())
        self._check(rc, "nazb_inverse")
        if want_z:
            out["z"] = z
        if want_lp:
            out["lp"] = lp
        if want_lse:
            out["lse_max"], out["lse_sum"] = lmax, lsum
        if want_sum:
            out["sum_n"] = sum_n
        return out

    def _default_groups(self, N: int, s_count: int) -> int:
        # (a) enough (tile, group) work items to fill the SMs when N is small; (b) when N is large, ~8 draws per group so
        # the packed weights of a group (a few MB per draw) stay L2-resident while every CTA sweeps its point tiles
        # over the same group (measured on cfg 3: DRAM reads 22.6 TB -> one pass over the images, +3.6 % throughput)
        tiles = (N + 63) // 64
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        g = max(1, min(s_count, -(-8 * sms // tiles)))
        if tiles >= 8 * sms:
            g = max(g, s_count // 8)
        return g

    def forward(self, z, ctx=None, bounds=None, *, want_logdet=False, s_begin: int = 0, s_count: Optional[int] = None):
        """Reference `sample` direction.  z: [S,N,D] per-draw base noise or [N,D] shared."""
        sh = self.shape
        s_count = self.S - s_begin if s_count is None else s_count
        z, c, rows, lo, hi = self._prep_points(z, ctx, bounds)
        shared = z.dim() == 2
        if z.shape[-1] != sh.D or (not shared and (z.dim() != 3 or z.shape[0] != s_count)):
            raise ValueError(f"z must be [N,{sh.D}] or [{s_count},N,{sh.D}]")
        N = z.shape[-2]
        x = torch.empty((s_count, N, sh.D), device=self.device, dtype=torch.float32)
        ld = torch.empty((s_count, N), device=self.device, dtype=torch.float32) if want_logdet else None
        if N == 0:   # `flow.sample([0])` is legal upstream: empty result, nothing to launch
            return (x, ld) if want_logdet else x
        rc = self._lib.nazb_forward(self._h, s_begin, s_count, z.data_ptr(), 1 if shared else 0, _ptr(c), rows, N,
                                    _ptr(lo), _ptr(hi), x.data_ptr(), _ptr(ld), self._stream())
        self._check(rc, "nazb_forward")
        return (x, ld) if want_logdet else x

    # ------------------------------------------------------------------
    def inverse_grad(self, x, ctx=None, bounds=None, *, want_dx: bool = False, want_lp: bool = False,
                     s_begin: int = 0, s_count: Optional[int] = None, weights: Optional[torch.Tensor] = None,
                     want_dctx: bool = False):
        """Value and gradient of sum_n log p(x_n | ctx_n; theta_s) per draw (what the reference gets from
        jax.value_and_grad / autograd of bflow_jax_maf.py:233-235).  Returns {"sum_n": [s_count] float64,
        "gW": [L][n_lin] of [S,out,in], "gb": [L][n_lin] of [S,out], "dx": [s_count,N,D], "lp": [s_count,N]};
        rows of gW / gb outside [s_begin, s_begin+s_count) stay zero.  Needs a handle created with engine="simt"
        holding a masked-affine or quadratic neural-spline flow (nazb_inverse_grad returns "unsupported" otherwise).
        `weights` ([N] shared or [s_count, N]): cotangents of lp — the result is then the vector-Jacobian product
        sum_n w[s,n] d lp[s,n] / d theta (nazb_inverse_vjp; "dx" = w d lp / d x) and "sum_n" is not returned.  `want_dctx`:
        also "dctx" [s_count, N, C] = w d lp / d ctx per point (weights default to 1; sum over N for a broadcast context)."""
        sh = self.shape
        if self._keepalive is None:
            raise RuntimeError("inverse_grad: pack() has not been called on this engine")
        s_count = self.S - s_begin if s_count is None else s_count
        x, c, rows, lo, hi = self._prep_points(x, ctx, bounds)
        if x.dim() != 2 or x.shape[1] != sh.D:
            raise ValueError(f"x must be [N,{sh.D}]")
        N = x.shape[0]
        L, n_lin = sh.L, len(sh.hidden) + 1
        dims = [sh.D + sh.C] + list(sh.hidden) + [sh.M * sh.D]
        masks = self._keepalive[2]
        gW = [torch.zeros((self.S, dims[j + 1], dims[j]), device=self.device, dtype=torch.float32)
              for _ in range(L) for j in range(n_lin)]
        gb = [torch.zeros((self.S, dims[j + 1]), device=self.device, dtype=torch.float32)
              for _ in range(L) for j in range(n_lin)]
        n = L * n_lin
        VP = C.c_void_p * n
        I64 = C.c_int64 * n
        sum_n = torch.zeros((s_count,), device=self.device, dtype=torch.float64)
        dx = torch.empty((s_count, N, sh.D), device=self.device, dtype=torch.float32) if want_dx else None
        lp = torch.empty((s_count, N), device=self.device, dtype=torch.float32) if want_lp else None
        tabs = (VP(*[m.data_ptr() for m in masks]), VP(*[t.data_ptr() for t in gW]), VP(*[t.data_ptr() for t in gb]),
                I64(*[dims[i % n_lin + 1] * dims[i % n_lin] for i in range(n)]), I64(*[dims[i % n_lin + 1] for i in range(n)]))
        dctx = None
        if want_dctx:
            if sh.C == 0:
                raise ValueError("want_dctx on a flow without context")
            dctx = torch.empty((s_count, N, sh.C), device=self.device, dtype=torch.float32)
            if weights is None:
                weights = torch.ones((N,), device=self.device, dtype=torch.float32)
        if weights is not None:
            w = _f32c(torch.as_tensor(weights), self.device)
            if tuple(w.shape) not in ((N,), (s_count, N)):
                raise ValueError(f"weights must be [N={N}] or [s_count={s_count}, N]")
            rc = 0 if N == 0 else self._lib.nazb_inverse_vjp(self._h, s_begin, s_count, x.data_ptr(), _ptr(c), rows, N, _ptr(lo), _ptr(hi),
                                                             *tabs, _ptr(dx), _ptr(dctx), _ptr(lp), w.data_ptr(), 0 if w.dim() == 1 else N,
                                                             self._stream())
            self._check(rc, "nazb_inverse_vjp")
            sum_n = None
        else:
            rc = 0 if N == 0 else self._lib.nazb_inverse_grad(self._h, s_begin, s_count, x.data_ptr(), _ptr(c), rows, N, _ptr(lo), _ptr(hi),
                                                              *tabs, _ptr(dx), _ptr(lp), sum_n.data_ptr(), self._stream())
            self._check(rc, "nazb_inverse_grad")
        out = {"sum_n": sum_n,
               "gW": [[gW[l * n_lin + j] for j in range(n_lin)] for l in range(L)],
               "gb": [[gb[l * n_lin + j] for j in range(n_lin)] for l in range(L)]}
        if want_dx:
            out["dx"] = dx
        if want_dctx:
            out["dctx"] = dctx
        if want_lp:
            out["lp"] = lp
        return out

    # ------------------------------------------------------------------
    def lse_finish(self, lse_max: torch.Tensor, lse_sum: torch.Tensor, log_norm: float) -> torch.Tensor:
        G, N = lse_max.shape
        out = torch.empty((N,), device=self.device, dtype=torch.float32)
        if N == 0:
            return out
        rc = self._lib.nazb_lse_finish(lse_max.data_ptr(), lse_sum.data_ptr(), G, N, float(log_norm), out.data_ptr(),
                                       self._stream())
        self._check(rc, "nazb_lse_finish")
        return out


def lse_reduce(lp: torch.Tensor, log_w: Optional[torch.Tensor] = None):
    """Stand-alone cross-draw (max, sum exp) over a materialised lp[S,N] (kernel group 4)."""
    L = _lib.lib()
    lp = lp.contiguous()
    S, N = lp.shape
    m = torch.empty((N,), device=lp.device, dtype=torch.float32)
    s = torch.empty((N,), device=lp.device, dtype=torch.float32)
    lw = None if log_w is None else log_w.to(device=lp.device, dtype=torch.float32).contiguous()
    rc = L.nazb_lse_reduce(lp.data_ptr(), S, N, _ptr(lw), m.data_ptr(), s.data_ptr(),
                           torch.cuda.current_stream(lp.device).cuda_stream)
    if rc != 0:
        raise _lib.NazbError(rc, "nazb_lse_reduce")
    return m, s


def lse_finish(lse_max: torch.Tensor, lse_sum: torch.Tensor, log_norm: float = 0.0) -> torch.Tensor:
    L = _lib.lib()
    if lse_max.dim() == 1:
        lse_max, lse_sum = lse_max.unsqueeze(0), lse_sum.unsqueeze(0)
    lse_max, lse_sum = lse_max.contiguous(), lse_sum.contiguous()
    G, N = lse_max.shape
    out = torch.empty((N,), device=lse_max.device, dtype=torch.float32)
    rc = L.nazb_lse_finish(lse_max.data_ptr(), lse_sum.data_ptr(), G, N, float(log_norm), out.data_ptr(),
                           torch.cuda.current_stream(lse_max.device).cuda_stream)
    if rc != 0:
        raise _lib.NazbError(rc, "nazb_lse_finish")
    return out


def importance(sum_n: torch.Tensor, log_prior: Optional[torch.Tensor] = None, log_q: Optional[torch.Tensor] = None):
    """-> (log_w [S] float64, log_evidence, ess, max_sum) as device tensors (pyro Importance semantics)."""
    L = _lib.lib()
    sum_n = sum_n.to(torch.float64).contiguous()
    S = sum_n.shape[0]
    dev = sum_n.device
    lw = torch.empty((S,), device=dev, dtype=torch.float64)
    out3 = torch.empty((3,), device=dev, dtype=torch.float64)
    lpz = None if log_prior is None else log_prior.to(device=dev, dtype=torch.float32).contiguous()
    lq = None if log_q is None else log_q.to(device=dev, dtype=torch.float32).contiguous()
    rc = L.nazb_importance(sum_n.data_ptr(), _ptr(lpz), _ptr(lq), S, lw.data_ptr(), out3.data_ptr(),
                           torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise _lib.NazbError(rc, "nazb_importance")
    return lw, out3[0], out3[1], out3[2]
