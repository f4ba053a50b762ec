"""Consumers of the [S, N, D] sample tensor: per-draw N-D histograms against fixed bin edges and the
HPD interval across draws of every bin — what `calibrate()` (src/naz/flows/bflow_jax_maf.py:405-460) does on the host
with `np.histogram2d` / `jnp.histogramdd` per draw and `hpd_vectorized` (src/naz/statutils.py:22-46).
CUDA only (libnazb); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def histogramdd_draws(samples: torch.Tensor, edges: Sequence, density: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """samples: CUDA fp32 [S, N, D]; edges: D monotonically increasing 1-D arrays (bin edges per dim, as
    `hist.numpy_bins` / the `bins=` argument of np.histogramdd).  Returns (counts uint32->int64 [S, *nbins],
    density fp32 [S, *nbins] or None) with numpy.histogramdd semantics per draw."""
    if not samples.is_cuda:
        raise RuntimeError("naz_b200.stats needs a CUDA tensor (no CPU fallback)")
    if samples.dim() != 3:
        raise ValueError("samples must be [S, N, D]")
    S, N, D = samples.shape
    if len(edges) != D:
        raise ValueError("one edge array per dimension")
    x = samples.contiguous().float()
    ed = [torch.as_tensor(e, dtype=torch.float64).flatten() for e in edges]
    nb = [int(e.numel()) - 1 for e in ed]
    if min(nb) < 1:
        raise ValueError("each dimension needs at least two edges")
    e_dev = torch.cat(ed).to(x.device)
    total = 1
    for b in nb:
        total *= b
    counts = torch.empty((S, total), dtype=torch.int32, device=x.device)
    dens = torch.empty((S, total), dtype=torch.float32, device=x.device) if density else None
    nb_arr = (C.c_int32 * D)(*nb)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        rc = L.nazb_histogramdd(x.data_ptr(), S, N, D, e_dev.data_ptr(), nb_arr, counts.data_ptr(),
                                dens.data_ptr() if density else None, _stream(x.device))
    if rc != 0:
        raise _lib.NazbError(rc, "nazb_histogramdd")
    shape = (S, *nb)
    return counts.view(shape).to(torch.int64), (dens.view(shape) if density else None)


def hpd_draws(values: torch.Tensor, alpha: float = 0.1) -> torch.Tensor:
    """values: CUDA fp32 [S, ...] (draws first).  Returns [2, ...]: the narrowest interval across draws holding
    floor((1 - alpha) S) + 1 order statistics of every trailing element (statutils.hpd_vectorized)."""
    if not values.is_cuda:
        raise RuntimeError("naz_b200.stats needs a CUDA tensor (no CPU fallback)")
    S = values.shape[0]
    v = values.contiguous().float().view(S, -1)
    M = v.shape[1]
    out = torch.empty((2, M), dtype=torch.float32, device=v.device)
    L = _lib.lib()
    with torch.cuda.device(v.device):
        rc = L.nazb_hpd(v.data_ptr(), S, M, float(alpha), out[0].data_ptr(), out[1].data_ptr(), _stream(v.device))
    if rc != 0:
        if rc == -1 and S - int((1.0 - alpha) * S) <= 0:
            raise ValueError("Too few elements for interval calculation")   # statutils.py:33-34
        raise _lib.NazbError(rc, "nazb_hpd")
    return out.view((2,) + tuple(values.shape[1:]))


def truncnorm_sample(uniform: torch.Tensor, loc, scale, low, high) -> Tuple[torch.Tensor, torch.Tensor]:
    """Truncated-normal guide (priors/TruncatedNormal.py:14-60).  uniform: CUDA fp32 [S, P] in (0, 1); loc / scale / low /
    high: scalars or [P].  Returns (samples [S, P] fp32, log_q [S] float64 = sum_p log q(sample))."""
    if not uniform.is_cuda:
        raise RuntimeError("naz_b200.stats needs a CUDA tensor (no CPU fallback)")
    x = uniform.contiguous().float()
    S, P = x.shape
    dev = x.device
    ps = []
    for v in (loc, scale, low, high):
        t = torch.as_tensor(v, dtype=torch.float32).to(dev).reshape(-1).contiguous()
        if t.numel() not in (1, P):
            raise ValueError("loc / scale / low / high must be scalars or have P elements")
        ps.append(t)
    y = torch.empty_like(x)
    log_q = torch.empty((S,), dtype=torch.float64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.nazb_truncnorm_sample(x.data_ptr(), S, P, ps[0].data_ptr(), ps[0].numel(), ps[1].data_ptr(), ps[1].numel(),
                                     ps[2].data_ptr(), ps[2].numel(), ps[3].data_ptr(), ps[3].numel(), y.data_ptr(),
                                     log_q.data_ptr(), _stream(dev))
    if rc != 0:
        raise _lib.NazbError(rc, "nazb_truncnorm_sample")
    return y, log_q
