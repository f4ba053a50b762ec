"""MC-dropout marginalisation — drop-in for src/naz/flows/mcdpflow.py:27-56.

Upstream loops `niter` times in train() mode, re-drawing an nn.Dropout mask per element per conditioner
call, and copies every iteration to the host.  Here the `niter` masks are drawn up front as explicit
per-unit keep-masks `[niter, L, n_hidden, H]`, folded into `niter` packed weight images,
and all iterations run in ONE kernel launch; the host copy happens once at the end."""
from __future__ import annotations

import numpy as np
import torch

from .flow import NormalizingFlow


class MCDPNormalizingFlow(NormalizingFlow):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.flow_maker = args[0]
        assert kwargs["dropout_p"] not in [None, 0.0]

    def draw_keep_masks(self, niter: int, generator=None) -> torch.Tensor:
        H = max(self.shape.hidden)
        shape = (niter, self.shape.L, len(self.shape.hidden), H)
        u = torch.rand(shape, generator=generator)
        return (u >= self.dropout_p).float()

    def sample_uncertain(self, niter, *args, condition=None, keep=None, base_noise=None, generator=None, **kwargs):
        shape = None
        for a in args:
            if isinstance(a, (list, tuple, torch.Size)):
                shape = list(a)
        n = int(torch.Size(shape).numel()) if shape is not None else base_noise.shape[-2]
        if keep is None:
            keep = self.draw_keep_masks(niter, generator)
        eng = self.make_engine(self.current_draw(), keep=keep, p_drop=float(self.dropout_p))
        z = base_noise if base_noise is not None else torch.randn((niter, n, self.theta_dim), device=eng.device)
        x = self.relabel.from_engine(eng.forward(z, self._cond(condition), self._bounds_e()))
        return x.cpu().detach().numpy()
