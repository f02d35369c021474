"""ctypes binding of ``libwsegk.so`` (include/wsegk.h).  There is no fallback: if the library is missing
or a call fails, a RuntimeError is raised — the product never routes through PyTorch ops or the oracle."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwsegk.so")

OP_CAM_NORM, OP_PCM, OP_FULLRES_NORM, OP_SEED_LABELS, OP_PROTO, OP_CONTRAST_LOSS, OP_SELECT, OP_ER_ECR = 1, 2, 3, 4, 5, 6, 7, 8
OP_LAZY_BWD = 9
PCM_PRECISE, PCM_DEFER_FIXUP = 1, 2

_P = c_void_p
_SIGNATURES = {
    "wsegk_version": (c_int, []),
    "wsegk_last_error": (c_char_p, []),
    "wsegk_workspace_bytes": (c_size_t, [c_int] * 7),
    "wsegk_mining_keys": (c_int, [_P, c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint, _P]),
    "wsegk_sgd_step": (c_int, [_P, _P, _P, _P, c_int, c_float, c_float, c_float, c_float, c_int, _P]),
    "wsegk_debug_launch_count": (ctypes.c_ulonglong, []),
    "wsegk_debug_pcm_profile": (c_int, [c_int]),
    "wsegk_debug_pcm_main_ms": (c_int, [ctypes.POINTER(c_float), ctypes.POINTER(c_float)]),
    "wsegk_debug_pcm_fixup_ms": (c_int, [ctypes.POINTER(c_float)]),
    "wsegk_debug_pcm_trace": (c_int, [_P]),
    "wsegk_debug_exchange_stamps": (c_int, [_P]),
    "wsegk_debug_stamp": (c_int, [_P, _P]),
    "wsegk_debug_pcm_trace_bwd": (c_int, [_P]),
    "wsegk_debug_force_radix": (c_int, [c_int]),
    "wsegk_debug_mine_phases": (c_int, [c_int]),
    "wsegk_debug_pcm_schedule": (c_int, [c_int, c_int, c_int, c_int, _P, c_int]),
    "wsegk_pcm_set_mode": (c_int, [c_int]),
    "wsegk_pcm_get_mode": (c_int, []),
    "wsegk_pcm_set_pairs": (c_int, [c_int, c_int]),
    "wsegk_pcm_multi_workspace_bytes": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P]),
    "wsegk_pcm_fwd_infer_multi": (c_int, [c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "wsegk_debug_pcm_layout": (c_int, [c_int] * 5 + [ctypes.POINTER(c_size_t), c_int]),
    "wsegk_cam_norm_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_pcm_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_pcm_fwd_infer": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_pcm_fwd_logits": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_pcm_fwd_ex": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, c_int, _P]),
    "wsegk_pcm_sign_fixup": (c_int, [c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_pcm_bwd_ex": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, c_int, _P]),
    "wsegk_pcm_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_fullres_norm_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_fullres_norm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P,
                                       c_size_t, _P]),
    "wsegk_fullres_norm_up_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P,
                                          c_size_t, _P]),
    "wsegk_fullres_norm_up_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_int, _P, c_size_t, _P]),
    "wsegk_gap_mlsm_up_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_gap_mlsm_up_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "wsegk_minpool_up_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_minpool_up_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_bilinear_resize_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "wsegk_bilinear_resize_bwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "wsegk_select_kth": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "wsegk_gap_mlsm_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_gap_mlsm_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "wsegk_minpool_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_minpool_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "wsegk_er_ecr_fwd": (c_int, [_P] * 11 + [c_int] * 4 + [_P, c_size_t, _P]),
    "wsegk_er_ecr_bwd": (c_int, [_P] * 14 + [c_int] * 4 + [_P]),
    "wsegk_seed_labels_fwd": (c_int, [_P, _P, c_float, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "wsegk_proto_topk_local": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P,
                                       c_size_t, _P]),
    "wsegk_proto_finalize": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, c_int, c_int, _P]),
    "wsegk_proto_exchange_bytes": (c_size_t, [c_int] * 4),
    "wsegk_proto_push": (c_int, [_P, _P, _P, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "wsegk_proto_finalize_symm": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P]),
    "wsegk_contrast_loss_fwd_bwd": (c_int, [_P] * 14 + [c_int] * 5 + [_P, c_size_t, _P]),
    "wsegk_mine_exchange_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "wsegk_contrast_loss_fwd_bwd_global": (c_int, [_P] * 14 + [c_int] * 5 + [_P, c_size_t, _P, _P, c_int, c_int, _P, _P]),
}

_lib = None
MISSING = []


def _missing_symbol(name):
    def raiser(*_a, **_k):
        raise RuntimeError(f"libwsegk.so does not export `{name}` (stale build?); there is no fallback")
    return raiser


def symbols():
    """Entry points declared in include/wsegk.h (tests check the built library exports each one)."""
    return tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m wseg_b200.build` "
                "(wseg_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                MISSING.append(name)            # tests assert this stays empty; calling it raises below
                setattr(handle, name, _missing_symbol(name))
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().wsegk_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (wsegk status {rc}): {last_error()}")


def workspace_bytes(op: int, N: int = 0, C: int = 0, K: int = 0, D: int = 0, h: int = 0, w: int = 0) -> int:
    return int(lib().wsegk_workspace_bytes(op, N, C, K, D, h, w))
