"""Tensor-facing wrappers of the C ABI.  torch supplies device memory, ``data_ptr()`` and the current stream; all
arithmetic happens inside libvipcup.so."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FLAG_GRAY, FLAG_HFLIP, FLAG_VFLIP, VipError  # noqa: F401


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise VipError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise VipError(f"{name} must be contiguous")


def preprocess(src: torch.Tensor, out_hw, crop_yxhw=None, jpeg_q=None, flags=None, out_dtype=torch.float32,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """crop -> bicubic resize -> /255 -> JPEG(q) -> flips -> gray on device tensors.

    src u8 [N,Hs,Ws,3]; crop_yxhw i32 [N,4] | None; jpeg_q i32 [N] (q<0 skips) | None; flags u8 [N] | None.
    Returns [N,Ho,Wo,3] float32 or bfloat16 (reference: dataset/dataset.py:31-37, dataset/augment.py:110-120,142-146).
    """
    _require_cuda(src, "src")
    if src.dtype != torch.uint8 or src.dim() != 4 or src.shape[-1] != 3:
        raise VipError("src must be uint8 [N,Hs,Ws,3]")
    n, hs, ws, _ = src.shape
    ho, wo = int(out_hw[0]), int(out_hw[1])
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise VipError("out_dtype must be float32 or bfloat16")
    if out is None:
        out = torch.empty((n, ho, wo, 3), dtype=out_dtype, device=src.device)
    else:
        _require_cuda(out, "out")
        if tuple(out.shape) != (n, ho, wo, 3) or out.dtype != out_dtype:
            raise VipError("out has the wrong shape or dtype")
    for name, t, dt, shape in (("crop_yxhw", crop_yxhw, torch.int32, (n, 4)), ("jpeg_q", jpeg_q, torch.int32, (n,)),
                               ("flags", flags, torch.uint8, (n,))):
        if t is not None:
            _require_cuda(t, name)
            if t.dtype != dt or tuple(t.shape) != shape:
                raise VipError(f"{name} must be {dt} with shape {shape}")
    with torch.cuda.device(src.device):
        rc = _lib.lib().vip_preprocess(_ptr(src), n, hs, ws, _ptr(crop_yxhw), _ptr(jpeg_q), _ptr(flags), ho, wo,
                                       _ptr(out), _lib.VIP_DTYPE_BF16 if out_dtype == torch.bfloat16 else
                                       _lib.VIP_DTYPE_F32, _stream_ptr())
    _lib.check(rc, "vip_preprocess")
    return out


def preprocess_host(src: torch.Tensor, out_hw, crop_yxhw=None, jpeg_q=None, flags=None, out_dtype=torch.float32,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """Same as :func:`preprocess` for HOST (ideally pinned) tensors: H2D, kernel and D2H are pipelined inside the
    library and the call returns after the result is in ``out``."""
    for name, t in (("src", src), ("crop_yxhw", crop_yxhw), ("jpeg_q", jpeg_q), ("flags", flags), ("out", out)):
        if t is not None and (t.is_cuda or not t.is_contiguous()):
            raise VipError(f"{name} must be a contiguous host tensor")
    n, hs, ws, _ = src.shape
    ho, wo = int(out_hw[0]), int(out_hw[1])
    if out is None:
        out = torch.empty((n, ho, wo, 3), dtype=out_dtype).pin_memory()
    rc = _lib.lib().vip_preprocess_host(_ptr(src), n, hs, ws, _ptr(crop_yxhw), _ptr(jpeg_q), _ptr(flags), ho, wo,
                                        _ptr(out), _lib.VIP_DTYPE_BF16 if out_dtype == torch.bfloat16 else
                                        _lib.VIP_DTYPE_F32)
    _lib.check(rc, "vip_preprocess_host")
    return out


def selftest_div255() -> int:
    import ctypes as C

    bad = C.c_uint64(0)
    rc = _lib.lib().vip_selftest_div255(C.byref(bad), _stream_ptr())
    _lib.check(rc, "vip_selftest_div255")
    return int(bad.value)

