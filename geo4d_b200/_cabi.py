"""ctypes binding of libgeo4d_b200.so (the C-ABI product boundary, include/geo4d_b200.h).

There is deliberately NO fallback: if the library is missing, fails to load, or
the device is not a Blackwell part, every op raises.  The CPU oracle lives in
``oracle/`` and is never reachable from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeo4d_b200.so")


class Geo4DError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("K", C.c_int), ("W", C.c_int), ("H", C.c_int), ("N", C.c_int),
        ("a_stride_w", C.c_int64), ("a_stride_h", C.c_int64), ("a_stride_n", C.c_int64),
        ("box_w", C.c_int), ("box_h", C.c_int), ("box_n", C.c_int),
        ("num_taps", C.c_int), ("tap_dx", C.c_int * 9), ("tap_dy", C.c_int * 9),
        ("b", C.c_void_p), ("n_out", C.c_int), ("b_batched", C.c_int),
        ("out", C.c_void_p), ("ldc", C.c_int64), ("out_fp32", C.c_int), ("alpha", C.c_float),
        ("bias", C.c_void_p), ("row_bias", C.c_void_p), ("row_bias_ld", C.c_int64),
        ("rows_per_bias", C.c_int), ("act", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("tile_n", C.c_int32), ("cta_pair", C.c_int32),
        ("split_k", C.c_int32), ("reserved_", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
    ]


class AlignLoopDesc(C.Structure):
    """g4_align_loop_desc of include/geo4d_b200.h (field order and types must match)."""
    _fields_ = [
        ("logd", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("pred", C.c_void_p), ("weight", C.c_void_p), ("invd", C.c_void_p),
        ("edge_ptr", C.c_void_p), ("edge_idx", C.c_void_p), ("scal", C.c_void_p),
        ("poses", C.c_void_p), ("S", C.c_void_p), ("invf", C.c_void_p), ("st", C.c_void_p),
        ("gpose", C.c_void_p), ("gS", C.c_void_p), ("gscal", C.c_void_p), ("gst", C.c_void_p),
        ("part", C.c_void_p), ("bar", C.c_void_p),
        ("im_poses", C.c_void_p), ("im_focal", C.c_void_p), ("pw_poses", C.c_void_p), ("s_depth", C.c_void_p),
        ("t_depth", C.c_void_p), ("ta_poses", C.c_void_p), ("adam_small", C.c_void_p),
        ("traj", C.c_void_p), ("e_img", C.c_void_p), ("valid_traj", C.c_void_p),
        ("N", C.c_int), ("G", C.c_int), ("HW", C.c_int), ("W", C.c_int), ("group_size", C.c_int),
        ("max_edges_per_image", C.c_int),
        ("n_lo", C.c_int), ("n_hi", C.c_int), ("chunks", C.c_int), ("it0", C.c_int), ("it1", C.c_int),
        ("start_b", C.c_int),
        ("temporal_smoothing_weight", C.c_float), ("translation_weight", C.c_float), ("base_scale", C.c_float),
        ("focal_break", C.c_float),
        ("world", C.c_int), ("rank", C.c_int), ("img_lo", C.c_int * 17), ("rec_doubles", C.c_int),
        ("peer_rec", C.c_void_p * 16), ("peer_flag", C.c_void_p * 16),
        ("flag_base", C.c_ulonglong), ("debug_ns", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the library (once).  Raises Geo4DError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Geo4DError(
                f"{LIB_PATH} not found: build it with `python -m geo4d_b200.build` "
                "(there is no CPU/PyTorch fallback for the product path)")
        try:
            _lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise Geo4DError(f"cannot load {LIB_PATH}: {e}") from e
        _lib.geo4d_last_error.restype = C.c_char_p
        _lib.geo4d_abi_version.restype = C.c_int
        # debug switches (see include/geo4d_b200.h): GEO4D_GEMM_PAIR = -1 (cost model) | 0 | 1,
        # GEO4D_GEMM_DIRECT_STORE = 1 forces the per-thread store epilogue
        if os.environ.get("GEO4D_GEMM_PAIR") is not None:
            _lib.geo4d_debug_gemm_pair_mode(int(os.environ["GEO4D_GEMM_PAIR"]))
        if os.environ.get("GEO4D_GEMM_DIRECT_STORE") is not None:
            _lib.geo4d_debug_gemm_direct_store(int(os.environ["GEO4D_GEMM_DIRECT_STORE"]))
    return _lib


def check(rc: int, what: str = "geo4d"):
    if rc != 0:
        msg = lib().geo4d_last_error().decode(errors="replace")
        raise Geo4DError(f"{what} failed (code {rc}): {msg}")


def require_device():
    import torch
    if not torch.cuda.is_available():
        raise Geo4DError("geo4d_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                         "there is no CPU fallback")
    if not lib().geo4d_device_supported():
        raise Geo4DError("geo4d_b200 kernels are compiled for sm_100a only; current device is not "
                         "compute capability 10.x")
