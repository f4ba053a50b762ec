"""ctypes binding of libnazb.so (include/nazb.h).  No CPU fallback: importing works anywhere, but any
compute call raises if the library is missing or no sm_100 device is present."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnazb.so")

NAZB_MAX_HIDDEN_LAYERS = 8
NAZB_MAX_DIM = 32

KIND_AFFINE, KIND_RQS, KIND_RLS = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
INV_INCREMENTAL, INV_JACOBI = 0, 1
ENGINE_NAMES = {ENGINE_AUTO: "auto", ENGINE_SIMT: "simt", ENGINE_TCGEN05: "tcgen05"}

# every symbol include/nazb.h declares (tests check that the .so exports all of them)
SYMBOLS = [
    "nazb_create", "nazb_destroy", "nazb_engine_in_use", "nazb_engine_for_direction", "nazb_pack", "nazb_inverse", "nazb_forward",
    "nazb_lse_reduce", "nazb_lse_finish", "nazb_importance", "nazb_strerror", "nazb_last_cuda_error",
    "nazb_packed_bytes", "nazb_launch_count", "nazb_histogramdd", "nazb_hpd", "nazb_pack_draw_map", "nazb_truncnorm_sample",
    "nazb_inverse_grad", "nazb_set_option", "nazb_get_option", "nazb_set_layer_affine", "nazb_host_spline_grad", "nazb_inverse_vjp",
]


class NazbDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("D", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("n_hidden", C.c_int32),
        ("hidden", C.c_int32 * NAZB_MAX_HIDDEN_LAYERS), ("count_bins", C.c_int32),
        ("bound", C.c_float), ("clip_lo", C.c_float), ("clip_hi", C.c_float),
        ("S", C.c_int32), ("engine", C.c_int32), ("inverse_mode", C.c_int32), ("device", C.c_int32),
    ]


class NazbError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = f"{where}: {_strerror(status)}"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


_lib = None


def _strerror(status: int) -> str:
    try:
        return lib().nazb_strerror(status).decode()
    except Exception:  # pragma: no cover
        return f"status {status}"


def lib() -> C.CDLL:
    """Load libnazb.so (once).  Raises a loud error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing — build it with `python -m naz_b200.build` (or __graft_entry__.build()). "
            "naz_b200 has no CPU / PyTorch fallback for the flow hot path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.nazb_create.argtypes = [C.POINTER(vp), C.POINTER(NazbDesc)]
    L.nazb_create.restype = C.c_int
    L.nazb_destroy.argtypes = [vp]
    L.nazb_destroy.restype = None
    L.nazb_engine_in_use.argtypes = [vp]
    L.nazb_engine_in_use.restype = C.c_int
    L.nazb_engine_for_direction.argtypes = [vp, C.c_int]
    L.nazb_engine_for_direction.restype = C.c_int
    L.nazb_pack.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(vp),
                            C.POINTER(i64), C.POINTER(i32), vp, f32, vp]
    L.nazb_pack.restype = C.c_int
    L.nazb_pack_draw_map.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), f32,
                                     C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), vp, f32, vp]
    L.nazb_pack_draw_map.restype = C.c_int
    L.nazb_inverse.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.nazb_inverse.restype = C.c_int
    L.nazb_forward.argtypes = [vp, i32, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.nazb_forward.restype = C.c_int
    L.nazb_lse_reduce.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.nazb_lse_reduce.restype = C.c_int
    L.nazb_lse_finish.argtypes = [vp, vp, i32, i32, f32, vp, vp]
    L.nazb_lse_finish.restype = C.c_int
    L.nazb_importance.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.nazb_importance.restype = C.c_int
    L.nazb_strerror.argtypes = [C.c_int]
    L.nazb_strerror.restype = C.c_char_p
    L.nazb_last_cuda_error.argtypes = [vp]
    L.nazb_last_cuda_error.restype = C.c_char_p
    L.nazb_packed_bytes.argtypes = [vp]
    L.nazb_packed_bytes.restype = C.c_int64
    L.nazb_histogramdd.argtypes = [vp, i32, i64, i32, vp, C.POINTER(i32), vp, vp, vp]
    L.nazb_histogramdd.restype = C.c_int
    L.nazb_hpd.argtypes = [vp, i32, i64, C.c_double, vp, vp, vp]
    L.nazb_hpd.restype = C.c_int
    L.nazb_truncnorm_sample.argtypes = [vp, i32, i64, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp]
    L.nazb_truncnorm_sample.restype = C.c_int
    L.nazb_inverse_grad.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.nazb_inverse_grad.restype = C.c_int
    L.nazb_inverse_vjp.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.nazb_inverse_vjp.restype = C.c_int
    L.nazb_set_option.argtypes = [vp, C.c_char_p, i32]
    L.nazb_set_option.restype = C.c_int
    L.nazb_host_spline_grad.argtypes = [f32, i32, f32, i32, vp, vp, vp, vp, vp]
    L.nazb_host_spline_grad.restype = C.c_int
    L.nazb_set_layer_affine.argtypes = [vp, vp, vp, vp]
    L.nazb_set_layer_affine.restype = C.c_int
    L.nazb_get_option.argtypes = [vp, C.c_char_p, C.POINTER(i32)]
    L.nazb_get_option.restype = C.c_int
    L.nazb_launch_count.argtypes = []
    L.nazb_launch_count.restype = C.c_int64
    _lib = L
    return L


def launch_count() -> int:
    return int(lib().nazb_launch_count())
