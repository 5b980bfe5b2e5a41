"""ctypes binding of libcbas_b200.so (the C ABI declared in include/cbas_b200.h).

There is deliberately no CPU fallback: if the shared library is missing or a call fails, the caller gets a
RuntimeError.  `python -m cbas_b200.build` (or `__graft_entry__.build()`) produces the library in-tree.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# CBAS_B200_LIB: load another build of the same library (A/B timing of two kernel versions on one GPU box)
LIB_PATH = os.environ.get("CBAS_B200_LIB") or os.path.join(_HERE, "libcbas_b200.so")

ABI_VERSION = 3  # CBAS_B200_ABI_VERSION in include/cbas_b200.h
OPT_ATTENTION_IMPL, OPT_PRUNE_LAST_LAYER, OPT_RESIZE_KERNEL, OPT_LN_FUSION, OPT_SERPENTINE = 0, 1, 2, 3, 4  # cbas_b200_encoder_set_option

_lib = None
_lock = threading.Lock()

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class EncoderCfg(C.Structure):
    _fields_ = [
        ("hidden", c_int32), ("layers", c_int32), ("heads", c_int32), ("intermediate", c_int32),
        ("prefix_tokens", c_int32), ("mode", c_int32), ("in_h", c_int32), ("in_w", c_int32),
        ("side", c_int32), ("max_frames", c_int32), ("ln_eps", c_float),
        ("resize_taps_y", c_int32), ("resize_taps_x", c_int32), ("patch", c_int32),
    ]


class LayerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "w_qkv", "c1_qkv", "b_qkv", "w_o", "b_o", "w_up", "c1_up", "b_up", "w_down", "b_down")]


class EncoderWeights(C.Structure):
    _fields_ = [
        ("w_patch", c_void_p), ("b_patch", c_void_p), ("prefix", c_void_p),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p), ("lnf_g", c_void_p), ("lnf_b", c_void_p),
        ("layers", C.POINTER(LayerWeights)),
        ("rs_ymin", c_void_p), ("rs_wy", c_void_p), ("rs_xmin", c_void_p), ("rs_wx", c_void_p),
        ("pos_embed", c_void_p),
    ]


class HeadCfg(C.Structure):
    _fields_ = [
        ("in_features", c_int32), ("out_features", c_int32), ("seq_len", c_int32), ("bottleneck", c_int32),
        ("lstm_hidden", c_int32), ("center_window", c_int32), ("ema_alpha", c_float),
        ("use_acceleration", c_int32), ("lstm_layers", c_int32),
    ]


HEAD_WEIGHT_PTRS = (
    "cls_w", "cls_b", "delta_w", "delta_b", "acc_w", "acc_b",
    "cls_ln_g", "cls_ln_b", "delta_ln_g", "delta_ln_b", "acc_ln_g", "acc_ln_b",
    "lin0_w", "lin0_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "att_w", "att_b",
    "w_ih_f", "w_hh_f", "b_ih_f", "b_hh_f", "w_ih_r", "w_hh_r", "b_ih_r", "b_hh_r",
)
HEAD_WEIGHT_PTRS_L1 = ("w_ih_f1", "w_hh_f1", "b_ih_f1", "b_hh_f1", "w_ih_r1", "w_hh_r1", "b_ih_r1", "b_hh_r1")


class HeadWeights(C.Structure):
    _fields_ = ([(n, c_void_p) for n in HEAD_WEIGHT_PTRS] + [("gate", c_float), ("attention_temp", c_float)]
                + [(n, c_void_p) for n in HEAD_WEIGHT_PTRS_L1])


# name -> (restype, argtypes); every symbol include/cbas_b200.h declares
SIGNATURES = {
    "cbas_b200_last_error": (C.c_char_p, []),
    "cbas_b200_abi_version": (C.c_int, []),
    "cbas_b200_launch_count": (C.c_ulonglong, []),
    "cbas_b200_profile_enable": (C.c_int, [C.c_int]),
    "cbas_b200_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "cbas_b200_encoder_create": (C.c_int, [C.POINTER(EncoderCfg), C.POINTER(EncoderWeights), C.POINTER(c_void_p)]),
    "cbas_b200_encoder_destroy": (None, [c_void_p]),
    "cbas_b200_encoder_forward_u8": (C.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "cbas_b200_encoder_forward_u8_plane": (C.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p,
                                                     c_void_p]),
    "cbas_b200_encoder_forward_f32": (C.c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "cbas_b200_encoder_debug_hidden": (C.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p,
                                                 c_void_p]),
    "cbas_b200_gemm_bf16": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p]),
    "cbas_b200_encoder_set_option": (C.c_int, [c_void_p, c_int32, c_int32]),
    "cbas_b200_ln_stats_init": (C.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "cbas_b200_gemm_resid_ln": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int32, c_int32, c_int32, c_void_p]),
    "cbas_b200_gemm_ln_a": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                      c_int32, c_int32, c_float, c_void_p]),
    "cbas_b200_attention_tc_qk_f16": (C.c_int, [c_int32]),
    "cbas_b200_debug_attention_trace": (C.c_int, [c_void_p]),
    "cbas_b200_debug_resize_tiled": (C.c_int, [c_int32]),
    "cbas_b200_debug_gemm_cta_group": (C.c_int, [c_int32]),
    "cbas_b200_layernorm": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "cbas_b200_attention": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p]),
    "cbas_b200_attention_tc_supported": (C.c_int, [c_int32, c_int32, c_int32]),
    "cbas_b200_attention_tc": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_void_p]),
    "cbas_b200_preprocess_green": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int32,
                                             c_void_p]),
    "cbas_b200_preprocess_resize": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int32,
                                              c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32,
                                              c_void_p]),
    "cbas_b200_head_create": (C.c_int, [C.POINTER(HeadCfg), C.POINTER(HeadWeights), C.POINTER(c_void_p)]),
    "cbas_b200_head_destroy": (None, [c_void_p]),
    "cbas_b200_head_infer": (C.c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "cbas_b200_head_forward_windows": (C.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "cbas_b200_actogram_bins": (C.c_int, [c_void_p, c_int64, c_int32, c_int32, c_float, c_int64, c_void_p,
                                          c_void_p]),
    "cbas_b200_actogram_bins_f64": (C.c_int, [c_void_p, c_int64, c_int32, c_int32, C.c_double, c_int64, c_void_p,
                                              c_void_p]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m cbas_b200.build` "
                    "(cbas_b200 has no CPU fallback)")
            handle = C.CDLL(LIB_PATH)
            ab_build = bool(os.environ.get("CBAS_B200_LIB"))  # an older build, for kernel-level A/B timing only
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name, None)
                if fn is None:
                    if ab_build:
                        continue
                    raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild")
                fn.restype = res
                fn.argtypes = args
            if handle.cbas_b200_abi_version() != ABI_VERSION and not ab_build:
                raise RuntimeError("libcbas_b200.so ABI version mismatch; rebuild")
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().cbas_b200_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def launch_count() -> int:
    return int(lib().cbas_b200_launch_count())


PROFILE_TAGS = ("preprocess", "patch_gemm", "layernorm", "qkv_gemm", "attention", "proj_gemm", "up_gemm", "down_gemm",
                "final_ln", "head_split", "head_proj_gemm", "head_features", "head_lin0_gemm", "head_center",
                "head_ih_gemm", "head_lstm", "actogram", "other")


def profile_enable(on: bool) -> None:
    check(lib().cbas_b200_profile_enable(1 if on else 0), "profile_enable")


def profile_read() -> dict:
    """{tag: (total_ms, launches)} for every tag that recorded at least one launch."""
    out = {}
    for i, name in enumerate(PROFILE_TAGS):
        ms, n = C.c_double(), C.c_longlong()
        check(lib().cbas_b200_profile_read(i, C.byref(ms), C.byref(n)), "profile_read")
        if n.value:
            out[name] = (ms.value, n.value)
    return out
