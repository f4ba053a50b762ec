"""naz_b200 — B200-native draw-batched normalizing-flow evaluation (the hot path of AnaryaRay1/naz).

Host side: PyTorch (plumbing) mirroring the reference's Python interface for this path.
Device side: hand-written sm_100a CUDA behind the C ABI in include/nazb.h (naz_b200/libnazb.so).
"""
from .engine import FlowEngine, FlowShape, importance, lse_finish, lse_reduce  # noqa: F401

__all__ = ["FlowEngine", "FlowShape", "importance", "lse_finish", "lse_reduce"]
