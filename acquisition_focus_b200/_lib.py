"""ctypes binding of libafb200.so (include/afb200.h).

The product path has NO CPU fallback: if the library is missing or a tensor is not on a CUDA
device the call raises.  PyTorch is used for device memory, streams and autograd plumbing only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libafb200.so")

# enums (keep in sync with include/afb200.h)
F32, BF16, F16, I64, I32, I16, U8 = range(7)
BILINEAR, NEAREST = 0, 1
PAD_ZERO, PAD_VALUE, PAD_DEVICE = 0, 1, 2
AFFINE_GRID, AFFINE_PRE, AFFINE_PARAMS = 0, 1, 2
ROT_ANGLE_AXIS, ROT_NORMAL = 0, 1

DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.int64: I64,
          torch.int32: I32, torch.int16: I16, torch.uint8: U8}


class AfbVolume(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int), ("B", C.c_int), ("C", C.c_int), ("D", C.c_int),
                ("H", C.c_int), ("W", C.c_int), ("sB", C.c_int64), ("sC", C.c_int64), ("sD", C.c_int64),
                ("sH", C.c_int64), ("sW", C.c_int64)]


class AfbViews(C.Structure):
    _fields_ = [("kind", C.c_int), ("V", C.c_int), ("theta", C.c_void_p), ("pre", C.c_void_p),
                ("pre_is_f64", C.c_int), ("params", C.c_void_p), ("gpre", C.c_void_p), ("init", C.c_void_p),
                ("R", C.c_int), ("spat", C.c_int), ("offset_clip", C.c_float), ("zoom_clip", C.c_float),
                ("nii_affine", C.c_void_p), ("fov_mm", C.c_double * 3), ("state", C.c_void_p)]


class AfbError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "afb_version": (C.c_int, []),
    "afb_error_string": (C.c_char_p, [C.c_int]),
    "afb_volume_min_workspace_bytes": (C.c_int64, []),
    "afb_volume_min": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_min_mask_bytes": (C.c_int64, [C.c_int64]),
    "afb_volume_min_mask": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_min_grad_fill_mask": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_volume_min_mask_half": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_host_narrow_labels": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "afb_min_grad_fill_mask_half": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_cast_from_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "afb_onehot_expand": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_void_p]),
    "afb_min_count_from_mask": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_probe_read": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_view_state_bytes": (C.c_int64, []),
    "afb_view_prologue": (C.c_int, [C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_slice_pad_grad": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "afb_slice_scatter": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "afb_min_grad_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_slice_fwd": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_slice_fwd3": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbVolume), C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "afb_slice_bwd_workspace_bytes": (C.c_int64, [C.c_int]),
    "afb_slice_bwd": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_slice_onehot_fwd": (C.c_int, [C.POINTER(AfbVolume), C.c_int, C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "afb_slice_onehot_bwd": (C.c_int, [C.POINTER(AfbVolume), C.c_int, C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_min_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_r6_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_r6_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_embed_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "afb_embed_workspace_bytes": (C.c_int64, [C.c_int]),
    "afb_compose_pre_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_upsample2d_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_upsample2d_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_rot3_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_rot3_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_peer_buffer_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "afb_peer_collective": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_label_group_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_label_extent_search": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_void_p, C.c_void_p]),
    "afb_embed_multi_fwd": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                                       C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "afb_embed_multi_bwd": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afb_embed_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


# experiment entries (profiles/ab_*.py only; not part of include/afb200.h)
_EXPERIMENTS = {
    "afbx_slice_scatter_priv_probe": (C.c_int, [C.POINTER(AfbVolume), C.POINTER(AfbViews), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfbError(
                f"{LIB_PATH} not found - build it with `python -m acquisition_focus_b200.build` "
                "(there is no CPU fallback for this path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGNATURES.items()) + list(_EXPERIMENTS.items()):
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().afb_error_string(rc).decode()
        raise AfbError(f"{what} failed: {msg} (code {rc})")


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise AfbError(f"{name} must live on a CUDA device (no CPU fallback); got {t.device}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def volume_desc(vol: torch.Tensor) -> AfbVolume:
    if vol.dtype not in DTYPES:
        raise AfbError(f"unsupported volume dtype {vol.dtype}")
    B, Cc, D, H, W = vol.shape
    s = vol.stride()
    return AfbVolume(vol.data_ptr(), DTYPES[vol.dtype], B, Cc, D, H, W, s[0], s[1], s[2], s[3], s[4])
