"""Loads ``libctc_crf_b200.so`` and declares the C ABI of ``include/ctc_crf_b200.h`` for ctypes.

There is no fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

_lib = None


class ctcOptions(C.Structure):
    """gpu_ctc/ctc.h:36-39"""
    _fields_ = [("stream", C.c_void_p), ("blank_label", C.c_int)]


# every symbol include/ctc_crf_b200.h declares (tests check the export list against the header)
SYMBOLS = [
    "Init", "Release", "compute_alpha", "compute_beta_and_grad", "compute_ctc_loss", "get_workspace_size",
    "ctcGetStatusString", "ccb_last_error", "ccb_den_loaded", "ccb_den_info", "ccb_den_alpha_floats", "ccb_den_aux_bytes",
    "ccb_ctc_workspace_bytes", "ccb_den_forward_backward", "ccb_ctc_forward_backward", "ccb_ctc_crf_loss_fwd",
    "ccb_ctc_crf_loss_logits_fwd", "ccb_ctc_loss_fwd", "ccb_ctc_align_workspace_bytes", "ccb_ctc_align",
    "ccb_launch_count", "ccb_debug_timeline", "ccb_plan_create", "ccb_plan_destroy", "ccb_plan_info", "ccb_plan_copy",
]
GLOBALS = ["DEN_NUM_ARCS", "DEN_NUM_STATES"]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ImportError(
            f"{LIB} is missing: build it with `python -m cat_b200.build` (nvcc, sm_100a). "
            "cat_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.Init.argtypes = [C.c_char_p, C.c_int, ip]; L.Init.restype = None
    L.Release.argtypes = [C.c_int, ip]; L.Release.restype = None
    L.compute_alpha.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.compute_alpha.restype = None
    L.compute_beta_and_grad.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.compute_beta_and_grad.restype = None
    L.compute_ctc_loss.argtypes = [vp, vp, ip, ip, ip, C.c_int, C.c_int, fp, vp, ctcOptions]
    L.compute_ctc_loss.restype = C.c_int
    L.get_workspace_size.argtypes = [ip, ip, C.c_int, C.c_int, ctcOptions, C.POINTER(C.c_size_t)]
    L.get_workspace_size.restype = C.c_int
    L.ctcGetStatusString.argtypes = [C.c_int]; L.ctcGetStatusString.restype = C.c_char_p
    L.ccb_last_error.argtypes = []; L.ccb_last_error.restype = C.c_char_p
    L.ccb_den_loaded.argtypes = [C.c_int]; L.ccb_den_loaded.restype = C.c_int
    L.ccb_den_info.argtypes = [C.POINTER(C.c_long)]; L.ccb_den_info.restype = C.c_int
    L.ccb_den_alpha_floats.argtypes = [C.c_int, C.c_int]; L.ccb_den_alpha_floats.restype = C.c_size_t
    L.ccb_den_aux_bytes.argtypes = [C.c_int, C.c_int]; L.ccb_den_aux_bytes.restype = C.c_size_t
    L.ccb_ctc_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]; L.ccb_ctc_workspace_bytes.restype = C.c_size_t
    L.ccb_den_forward_backward.argtypes = [vp, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                           vp, C.c_long, C.c_long, C.c_float, vp, vp, vp]
    L.ccb_den_forward_backward.restype = C.c_int
    L.ccb_ctc_forward_backward.argtypes = [vp, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp,
                                           C.c_int, C.c_int, vp, vp, C.c_long, C.c_long, C.c_float, vp, vp]
    L.ccb_ctc_forward_backward.restype = C.c_int
    L.ccb_ctc_crf_loss_fwd.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int,
                                       C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp]
    L.ccb_ctc_crf_loss_fwd.restype = C.c_int
    L.ccb_ctc_crf_loss_logits_fwd.argtypes = L.ccb_ctc_crf_loss_fwd.argtypes
    L.ccb_ctc_crf_loss_logits_fwd.restype = C.c_int
    L.ccb_ctc_loss_fwd.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int,
                                   C.c_float, vp, vp, vp, vp, vp]
    L.ccb_ctc_loss_fwd.restype = C.c_int
    L.ccb_ctc_align_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ccb_ctc_align_workspace_bytes.restype = C.c_size_t
    L.ccb_ctc_align.argtypes = [vp, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int,
                                vp, vp, vp, vp]
    L.ccb_ctc_align.restype = C.c_int
    L.ccb_launch_count.argtypes = []; L.ccb_launch_count.restype = C.c_long
    L.ccb_debug_timeline.argtypes = [vp, C.c_int, C.c_int]; L.ccb_debug_timeline.restype = None
    L.ccb_plan_create.argtypes = [C.c_char_p, C.c_int, C.c_int]; L.ccb_plan_create.restype = vp
    L.ccb_plan_destroy.argtypes = [vp]; L.ccb_plan_destroy.restype = None
    L.ccb_plan_info.argtypes = [vp, C.POINTER(C.c_long)]; L.ccb_plan_info.restype = C.c_int
    L.ccb_plan_copy.argtypes = [vp, C.c_int, vp, C.c_size_t]; L.ccb_plan_copy.restype = C.c_int
    _lib = L
    return L


def last_error() -> str:
    return lib().ccb_last_error().decode("utf-8", "replace")
