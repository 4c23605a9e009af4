"""ctypes binding of libvipcup.so -- the only way Python reaches the kernels (no torch types cross the ABI)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIP_LIB_PATH") or os.path.join(HERE, "libvipcup.so")   # override: A/B experiments only

VIP_DTYPE_F32, VIP_DTYPE_BF16 = 0, 1
FLAG_HFLIP, FLAG_VFLIP, FLAG_GRAY = 1, 2, 4


class VipError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """vip_epilogue_t of include/vipcup.h"""
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int), ("colscale", C.c_void_p), ("residual", C.c_void_p),
                ("ldr", C.c_int), ("out", C.c_void_p), ("ldc", C.c_int), ("out_dtype", C.c_int),
                ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p),
                ("row_stats", C.c_void_p), ("gap", C.c_void_p), ("gap_rows", C.c_int),
                ("row_gate", C.c_void_p), ("gate_rows", C.c_int), ("residual_lo", C.c_void_p), ("out_lo", C.c_void_p),
                ("row_pivot", C.c_void_p)]


_lib = None

# name -> (restype, argtypes); must list every symbol include/vipcup.h declares (tests check this)
SIGNATURES = {
    "vip_version": (C.c_char_p, []),
    "vip_last_error": (C.c_char_p, []),
    "vip_launch_count": (C.c_int64, []),
    "vip_launch_count_reset": (None, []),
    "vip_preprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vip_preprocess_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "vip_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vip_gemm_bf16_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(Epilogue), C.c_void_p]),
    "vip_conv2d_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(Epilogue), C.c_void_p]),
    "vip_conv2d_slice_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), C.c_void_p]),
    "vip_split_attention2_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vip_avgpool3s2_bf16": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "vip_act_scale_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p]),
    "vip_eca_gate_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "vip_memset_async": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "vip_im2col_bf16": (C.c_int, [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_int, C.c_void_p]),
    "vip_avgpool2_same_bf16": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "vip_global_avgpool_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vip_scale_add_act_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p]),
    "vip_layernorm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_longlong,
                                     C.c_int, C.c_float, C.c_void_p]),
    "vip_row_stats_finalize": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "vip_dwconv3x3_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "vip_maxpool3s2_bf16": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "vip_window_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "vip_dwconv_bf16": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 11 + [C.c_void_p]),
    "vip_layernorm_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "vip_pad_crop": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "vip_head_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_void_p]),
    "vip_scale_cast_fx_bf16": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]),
    "vip_gemm_grouped_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(Epilogue), C.c_void_p]),
    "vip_scale_weights_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vip_mlp_fused_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "vip_cast_f32_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "vip_selftest_div255": (C.c_int, [C.POINTER(C.c_uint64), C.c_void_p]),
    "vip_jpeg_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p]),
    "vip_jpeg_plan": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vip_jpeg_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load libvipcup.so (built by ``__graft_entry__.build()`` / ``python -m vipcup_b200.build``). Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VipError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vip_last_error().decode("utf-8", "replace")
        raise VipError(f"{what or 'libvipcup'} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(lib().vip_launch_count())


def launch_count_reset() -> None:
    lib().vip_launch_count_reset()
