"""
ctypes binding of libpy4cast_b200.so (C ABI declared in include/py4cast_b200.h).

The product path has NO CPU / eager fallback: if the shared library is missing
or a tensor is not on a CUDA device, calls raise.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

P4C_F32, P4C_BF16 = 0, 1
HIDDEN = 64
MAX_SRC = 3
MAX_PREV = 8

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpy4cast_b200.so")
if os.environ.get("P4C_LIB_AB"):  # developer aid: A/B timing of another in-tree build of the same ABI (scripts/ab_build.sh)
    _LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", os.environ["P4C_LIB_AB"])


class Src(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("index", C.c_void_p),
        ("batch_stride", C.c_int64),
        ("ld", C.c_int32),
        ("width", C.c_int32),
        ("seg_rowptr", C.c_void_p),
        ("seg_perm", C.c_void_p),
        ("seg_scale", C.c_void_p),
        ("seg_out", C.c_void_p),
    ]


class Mlp(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("batch", C.c_int32),
        ("nsrc", C.c_int32),
        ("k1", C.c_int32),
        ("out_features", C.c_int32),
        ("layer_norm", C.c_int32),
        ("dtype", C.c_int32),
        ("full_rows", C.c_int32),
        ("src", Src * MAX_SRC),
        ("w1", C.c_void_p),
        ("b1", C.c_void_p),
        ("w2", C.c_void_p),
        ("b2", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("ln_eps", C.c_float),
        ("packed", C.c_void_p),
        ("tile_rows", C.c_int32),
    ]


class MlpBwd(C.Structure):
    _fields_ = [
        ("dy_a", C.c_void_p),
        ("dy_a_index", C.c_void_p),
        ("dy_a_scale", C.c_void_p),
        ("dy_a_batch_stride", C.c_int64),
        ("dy_b", C.c_void_p),
        ("h1", C.c_void_p),
        ("stats", C.c_void_p),
        ("dh1_out", C.c_void_p),
        ("dx", C.c_void_p * MAX_SRC),
        ("dx_ld", C.c_int32 * MAX_SRC),
        ("dx_accumulate", C.c_int32 * MAX_SRC),
        ("dx0_add_dy", C.c_int32),
        ("g_w1", C.c_void_p),
        ("g_b1", C.c_void_p),
        ("g_w2", C.c_void_p),
        ("g_b2", C.c_void_p),
        ("g_gamma", C.c_void_p),
        ("g_beta", C.c_void_p),
        ("wgrad_partials", C.c_void_p),
        ("wgrad_partials_floats", C.c_int64),
        ("partials_mode", C.c_int32),
        ("dx0_batch_sum", C.c_void_p),
        ("dx_group_out", C.c_void_p),
        ("dx_group", C.c_int32),
        ("dx_group_src", C.c_int32),
    ]


PACK_BYTES = 9 * 8192 + 4 * 64 * 4


class Pack(C.Structure):
    _fields_ = [
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("dst", C.c_void_p),
        ("k1", C.c_int32), ("out_features", C.c_int32), ("nsrc", C.c_int32),
        ("width", C.c_int32 * MAX_SRC),
    ]


class WGrad(C.Structure):
    _fields_ = [
        ("partials", C.c_void_p),
        ("g_w1", C.c_void_p), ("g_b1", C.c_void_p), ("g_w2", C.c_void_p), ("g_b2", C.c_void_p),
        ("g_gamma", C.c_void_p), ("g_beta", C.c_void_p),
        ("nblocks", C.c_int32), ("k1", C.c_int32), ("out_features", C.c_int32), ("_pad", C.c_int32),
    ]


class SegPart(C.Structure):
    _fields_ = [
        ("rowptr", C.c_void_p), ("perm", C.c_void_p), ("scale", C.c_void_p),
        ("num_seg", C.c_int32), ("batch", C.c_int32), ("src_batch_stride", C.c_int64),
        ("col0", C.c_int32), ("kblock", C.c_int32),
        ("out", C.c_void_p), ("out_batch_stride", C.c_int64),
        ("accumulate", C.c_int32), ("tile_rows", C.c_int32),
    ]


class ArX(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("nodes", C.c_int32),
        ("f_state", C.c_int32),
        ("f_static", C.c_int32),
        ("f_forcing", C.c_int32),
        ("num_prev", C.c_int32),
        ("prev", C.c_void_p * MAX_PREV),
        ("prev_batch_stride", C.c_int64 * MAX_PREV),
        ("stat", C.c_void_p),
        ("stat_batch_stride", C.c_int64),
        ("forcing", C.c_void_p),
        ("forcing_batch_stride", C.c_int64),
        ("x_ld", C.c_int32),
    ]


class AdamWState(C.Structure):
    _fields_ = [
        ("step", C.c_int64),
        ("lr", C.c_float),
        ("bias1", C.c_float),
        ("bias2", C.c_float),
        ("_pad", C.c_float),
    ]


# every symbol include/py4cast_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_int
SYMBOLS = {
    "p4c_last_error": (C.c_char_p, []),
    "p4c_version": (_int, []),
    "p4c_launch_count": (_i64, []),
    "p4c_reset_launch_count": (None, []),
    "p4c_sizeof": (_int, [_int]),
    "p4c_csr_build": (_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "p4c_index_to_i32": (_int, [_vp, _vp, _i64, _vp]),
    "p4c_segment_sum": (_int, [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _i64, _int, _int, _int, _vp]),
    "p4c_set_engine": (None, [_int]),
    "p4c_set_trace_buffer": (None, [_vp]),
    "p4c_set_cta_reserve": (None, [_int]),
    "p4c_get_cta_reserve": (_int, []),
    "p4c_set_ws_min_tiles": (None, [_i64]),
    "p4c_get_ws_min_tiles": (_i64, []),
    "p4c_ws_launch_count": (_i64, []),
    "p4c_mlp_bwd_caps": (_i32, [C.POINTER(Mlp)]),
    "p4c_set_ws_bwd_min_tiles": (None, [_i64]),
    "p4c_get_ws_bwd_min_tiles": (_i64, []),
    "p4c_ws_bwd_launch_count": (_i64, []),
    "p4c_ws_bwd_bsum_launch_count": (_i64, []),
    "p4c_get_engine": (_int, []),
    "p4c_mlp_forward": (_int, [C.POINTER(Mlp), _vp, _vp, _vp, _vp, _vp, _vp]),
    "p4c_mlp_forward_aggregate": (_int, [C.POINTER(Mlp), _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp]),
    "p4c_mlp_wgrad_partial_floats": (_i64, [C.POINTER(Mlp)]),
    "p4c_mlp_wgrad_blocks": (_i32, [C.POINTER(Mlp)]),
    "p4c_pack_weights": (_int, [_vp, _i32, _vp]),
    "p4c_wgrad_reduce_batched": (_int, [_vp, _i32, _vp]),
    "p4c_seg_linear": (
        _int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i64, _int, _i32, _int, _vp]
    ),
    "p4c_seg_linear_pair": (_int, [_vp, _vp, _i32, _vp, C.POINTER(SegPart), C.POINTER(SegPart), _int, _vp]),
    "p4c_mlp_backward": (_int, [C.POINTER(Mlp), C.POINTER(MlpBwd), _vp]),
    "p4c_rows_matmul": (_int, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _int, _int, _vp]),
    "p4c_batch_sum": (_int, [_vp, _i64, _i32, _i64, _vp, _int, _int, _vp]),
    "p4c_set_ar_vector_paths": (None, [_int]),
    "p4c_ar_concat": (_int, [C.POINTER(ArX), _vp, _int, _vp]),
    "p4c_ar_step_forward": (
        _int,
        [_vp, _int, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i64,
         C.POINTER(ArX), _vp, _int, _vp],
    ),
    "p4c_ar_step_backward": (
        _int,
        [_vp, _i64, _vp, _i64, _vp, _int, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _int,
         C.POINTER(C.c_void_p), _vp],
    ),
    "p4c_loss_denominator": (_int, [_vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp]),
    "p4c_loss_partial_floats": (_i64, [_i64, _i32, _i32, _int]),
    "p4c_loss_forward": (
        _int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _int, _int, _vp, _vp, _vp]
    ),
    "p4c_loss_backward": (
        _int,
        [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _int, _int, _vp, _vp, _vp],
    ),
    "p4c_acc_partial_floats": (_i64, [_i64, _i32, _i32]),
    "p4c_acc_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "p4c_psd_partial_floats": (_i64, [_i64, _i32, _i32, _i32]),
    "p4c_psd_first_row": (_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "p4c_psd_profile": (_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "p4c_feature_stats_partial_bytes": (_i64, [_i32, _i64, _i32]),
    "p4c_feature_stats": (_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "p4c_scale_shift": (_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "p4c_forcing_fill": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp]),
    "p4c_forcing_assemble": (_int, [_vp, _int, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "p4c_adamw_advance": (_int, [_vp, _f32, _f32, _i64, _i64, _f32, _f32, _vp]),
    "p4c_adamw_step": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _vp]),
    "p4c_cast": (_int, [_vp, _int, _vp, _int, _i64, _vp]),
    "p4c_add": (_int, [_vp, _vp, _vp, _i64, _int, _vp]),
}

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the CUDA library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise RuntimeError(
            f"py4cast_b200: CUDA library not found at {_LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C py4cast_b200/csrc`). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    for which, st in enumerate((Src, Mlp, MlpBwd, ArX, AdamWState, Pack, WGrad)):
        if lib.p4c_sizeof(which) != C.sizeof(st):
            raise RuntimeError(
                f"py4cast_b200: ctypes layout of {st.__name__} ({C.sizeof(st)} B) does not match the "
                f"compiled library ({lib.p4c_sizeof(which)} B): rebuild it"
            )
    _lib = lib
    return lib


class P4CError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().p4c_last_error().decode("utf-8", "replace")
        raise P4CError(f"{what} failed (rc={rc}): {msg}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return P4C_F32
    if dt == torch.bfloat16:
        return P4C_BF16
    raise TypeError(f"py4cast_b200 kernels support float32 and bfloat16 storage, got {dt}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise P4CError(
            "py4cast_b200 kernels need CUDA tensors (no CPU fallback); got a tensor on "
            f"{t.device}"
        )
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().p4c_launch_count())


def reset_launch_count() -> None:
    load().p4c_reset_launch_count()
