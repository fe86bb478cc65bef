"""ctypes binding of libtransvae_sm100.so (the C ABI declared in include/transvae_sm100.h).

The library is built in-tree by ``deepl-project_b200/build.py``.  There is no CPU fallback: if the shared
object is missing or no sm_100 device is current, the compute entry points raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# TVAE_LIB: load another build of the same ABI (kernel experiments, tools/experiments/)
LIB_PATH = os.environ.get("TVAE_LIB") or os.path.join(_PKG_ROOT, "libtransvae_sm100.so")

MAX_TAPS = 20
MAX_PHASES = 4
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("split", C.c_int32)]


class Tap(C.Structure):
    _fields_ = [("map", C.c_int32), ("c_off", C.c_int32), ("dw", C.c_int32), ("p", C.c_int32), ("dh", C.c_int32),
                ("kblocks", C.c_int32), ("wk_off", C.c_int32)]


class MtGemmDesc(C.Structure):
    _fields_ = [
        ("a0", View), ("a1", View), ("out", View), ("res", View),
        ("w", C.c_void_p), ("n_total", C.c_int32), ("k_total", C.c_int32),
        ("num_phases", C.c_int32),
        ("ntaps", C.c_int32 * MAX_PHASES),
        ("taps", (Tap * MAX_TAPS) * MAX_PHASES),
        ("out_p", C.c_int32 * MAX_PHASES),
        ("out_c_off", C.c_int32 * MAX_PHASES),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("row_scale", C.c_void_p), ("row_shift", C.c_void_p), ("col_sum", C.c_void_p),
        ("rope_tab", C.c_void_p), ("rope_C", C.c_int32), ("rope_H", C.c_int32), ("rope_W", C.c_int32),
        ("q_scale", C.c_float),
        ("out_f32", C.c_void_p), ("out_n", C.c_int32),
        ("act_grad", C.c_int32), ("z", C.c_void_p),
        ("gn_sums", C.c_void_p), ("gn_groups", C.c_int32), ("out_act", C.c_void_p),
        ("gnb_x", C.c_void_p), ("gnb_sums", C.c_void_p), ("gnb_gamma", C.c_void_p), ("gnb_beta", C.c_void_p),
        ("gnb_part", C.c_void_p), ("gnb_groups", C.c_int32), ("gnb_eps", C.c_float), ("gnb_silu", C.c_int32),
    ]


_PROTOS = {
    "tvae_abi_version": (C.c_int, []),
    "tvae_last_error": (C.c_char_p, []),
    "tvae_device_ok": (C.c_int, []),
    "tvae_num_sms": (C.c_int, []),
    "tvae_set_reserved_sms": (C.c_int, [C.c_int32]),
    "tvae_mtgemm": (C.c_int, [C.POINTER(MtGemmDesc), C.c_void_p]),
    "tvae_attn_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_im2col_in": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_groupnorm_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_groupnorm_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "tvae_groupnorm_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "tvae_row_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_void_p]),
    "tvae_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "tvae_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "tvae_reparam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_int32, C.c_void_p]),
    "tvae_loss_l1_kl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "tvae_mtgemm_wgrad": (C.c_int, [C.POINTER(MtGemmDesc), C.c_void_p, C.c_void_p]),
    "tvae_mtgemm_wgrad_bias": (C.c_int, [C.POINTER(MtGemmDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvae_bias_act_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "tvae_bias_act_bwd_4d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_act_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "tvae_groupnorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                     C.c_void_p]),
    "tvae_groupnorm_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                           C.c_int32, C.c_void_p]),
    "tvae_token_norm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_token_norm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_attn_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_attn_bwd_dq_slices": (C.c_int, [C.c_int32]),
    "tvae_attn_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_rope_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                C.c_float, C.c_int32, C.c_void_p]),
    "tvae_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "tvae_latent_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "tvae_dwconv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_dwconv3x3_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "tvae_weight_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_wgrad_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_fold_qkv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "tvae_fold_qkv_bwd": (C.c_int, [C.c_void_p] * 8 + [C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_upconv1_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tvae_grad_sumsq": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "tvae_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_void_p]),
    "tvae_cast_f32_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tvae_multi_tensor_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "tvae_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_PROTOS)


def load() -> C.CDLL:
    """Load the shared object (no device needed) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python deepl-project_b200/build.py` (or __graft_entry__.build()). "
            "The TransVAE B200 path has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tvae_abi_version() != 4:
        raise RuntimeError("libtransvae_sm100.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().tvae_last_error() or b"").decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libtransvae_sm100 {what} failed (rc={rc}): {last_error()}")


def require_device() -> None:
    if not load().tvae_device_ok():
        raise RuntimeError("TransVAE B200 path needs an sm_100 (B200) CUDA device; there is no CPU fallback")
