"""JPEG decode on the device: host side of ``vip_jpeg_parse`` / ``vip_jpeg_plan`` / ``vip_jpeg_decode`` (include/vipcup.h).

Stands where the reference calls ``tf.io.read_file`` -> ``tf.image.decode_jpeg(channels=3)`` (dataset/dataset.py:24-28).
The host only reads the files and walks their marker segments (``vip_jpeg_parse``, C, GIL released); the entropy-coded
bytes travel to the GPU as they are (~8x fewer bytes over PCIe than decoded pixels) and Huffman decode, inverse DCT, chroma
upsampling and colour conversion run there, bit-identical to libjpeg-turbo.  Files outside the baseline subset (progressive,
PNG, CMYK ...) are decoded by libjpeg / Pillow on the host -- the reference's own CPU decoder -- into the same output
buffer."""
from __future__ import annotations

import ctypes as C
import io

import numpy as np
import torch

from . import _lib

VIP_JPEG_OK, VIP_JPEG_NOT_JPEG, VIP_JPEG_UNSUPPORTED = 0, 1, 2


class JpegDesc(C.Structure):
    """vip_jpeg_desc of include/vipcup.h"""
    _fields_ = [("status", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32),
                ("hs", C.c_int32 * 3), ("vs", C.c_int32 * 3), ("tq", C.c_int32 * 3), ("td", C.c_int32 * 3),
                ("ta", C.c_int32 * 3), ("restart_interval", C.c_int32), ("scan_offset", C.c_int32),
                ("scan_bytes", C.c_int32), ("file_offset", C.c_int64), ("dst_offset", C.c_int64),
                ("coef_offset", C.c_int64), ("qt", (C.c_uint16 * 64) * 4), ("huff_bits", (C.c_uint8 * 16) * 4),
                ("huff_vals", (C.c_uint8 * 256) * 4)]


DESC_BYTES = C.sizeof(JpegDesc)


def parse(file_bytes) -> JpegDesc:
    """Header walk of one file (host only)."""
    d = JpegDesc()
    _lib.check(_lib.lib().vip_jpeg_parse(bytes(file_bytes), len(file_bytes), C.byref(d)), "vip_jpeg_parse")
    return d


def read_and_parse(path):
    """(file bytes, descriptor) of one file: the unit of work of the dataset's thread pool."""
    with open(path, "rb") as f:
        data = f.read()
    return data, bytes(parse(data))        # the descriptor as bytes: the batch assembly is a join, not a per-image copy


def _host_decode(data):
    from PIL import Image

    with Image.open(io.BytesIO(data)) as im:
        return np.asarray(im.convert("RGB"))


class DecodedBatch:
    """Result of :func:`decode_batch`: one flat device uint8 buffer and the (offset, H, W) of every image in it."""

    def __init__(self, flat, layout, n_device, n_host, err):
        self.flat, self.layout, self.n_device, self.n_host, self.err = flat, layout, n_device, n_host, err

    def __len__(self):
        return len(self.layout)

    def image(self, i):
        off, h, w = self.layout[i]
        return self.flat[off: off + h * w * 3].view(h, w, 3)

    def uniform_shape(self):
        shapes = {(h, w) for _, h, w in self.layout}
        return next(iter(shapes)) if len(shapes) == 1 else None

    def stacked(self):
        """[N,H,W,3] view when every image has the same size (the plan packs the slices densely in order)."""
        hw = self.uniform_shape()
        if hw is None:
            raise ValueError("images of different sizes: use image(i)")
        return self.flat[: len(self.layout) * hw[0] * hw[1] * 3].view(len(self.layout), hw[0], hw[1], 3)

    def check(self):
        """Synchronises and raises if the entropy decoder flagged a corrupt stream."""
        if self.err is not None:
            bad = torch.nonzero(self.err).flatten().tolist()
            if bad:
                raise _lib.VipError(f"corrupt JPEG entropy-coded data in images {bad}")


# byte offsets of the vip_jpeg_desc fields the host touches per batch (vectorised with numpy instead of per-image ctypes)
_OFF_STATUS, _OFF_WIDTH, _OFF_HEIGHT = JpegDesc.status.offset, JpegDesc.width.offset, JpegDesc.height.offset
_OFF_FILE, _OFF_DST = JpegDesc.file_offset.offset, JpegDesc.dst_offset.offset


def decode_batch(files, descs=None, device=None) -> DecodedBatch:
    """files: list of bytes objects (whole files); descs: their parsed descriptors or None.  Returns device pixels.

    Work on the current stream: one H2D copy of the concatenated files, one of the descriptors, two kernels.  The host side
    is a handful of numpy / bytes operations per BATCH (the per-image Python loop it replaced cost 70 us per image and
    starved the GPU in the configs[4] run)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = len(files)
    if n == 0:
        return DecodedBatch(torch.empty((1,), dtype=torch.uint8, device=device), [], 0, 0, None)
    if descs is None:
        descs = [parse(f) for f in files]
    raw = np.frombuffer(b"".join(bytes(d) for d in descs), np.uint8).reshape(n, DESC_BYTES).copy()
    status = raw[:, _OFF_STATUS: _OFF_STATUS + 4].view(np.int32)[:, 0]
    host_pixels = {}
    for i in np.nonzero(status != VIP_JPEG_OK)[0].tolist():
        px = _host_decode(files[i])                    # the reference's CPU decoder for what the kernels do not cover
        host_pixels[i] = px
        raw[i, _OFF_HEIGHT: _OFF_HEIGHT + 4].view(np.int32)[0] = px.shape[0]
        raw[i, _OFF_WIDTH: _OFF_WIDTH + 4].view(np.int32)[0] = px.shape[1]
    on_dev = status == VIP_JPEG_OK
    lens = np.fromiter((len(f) for f in files), np.int64, n)
    padded = np.where(on_dev, (lens + 15) // 16 * 16, 0)       # 16-byte aligned starts (word refills of the bit reader)
    offs = np.cumsum(padded) - padded
    total = int(padded.sum())
    raw[:, _OFF_FILE: _OFF_FILE + 8] = offs.astype(np.int64).view(np.uint8).reshape(n, 8)
    dst_bytes, coef_blocks = C.c_int64(0), C.c_int64(0)
    _lib.check(_lib.lib().vip_jpeg_plan(raw.ctypes.data, n, C.byref(dst_bytes), C.byref(coef_blocks)), "vip_jpeg_plan")
    dst_off = raw[:, _OFF_DST: _OFF_DST + 8].view(np.int64)[:, 0]
    hh = raw[:, _OFF_HEIGHT: _OFF_HEIGHT + 4].view(np.int32)[:, 0]
    ww = raw[:, _OFF_WIDTH: _OFF_WIDTH + 4].view(np.int32)[:, 0]
    layout = list(zip(dst_off.tolist(), hh.tolist(), ww.tolist()))
    flat = torch.empty((max(int(dst_bytes.value), 1),), dtype=torch.uint8, device=device)
    n_dev = n - len(host_pixels)
    err = None
    if n_dev:
        blob = b"".join(f.ljust(int(p), b"\0") for f, p, ok in zip(files, padded.tolist(), on_dev.tolist()) if ok)
        stage = torch.empty((max(total, 16),), dtype=torch.uint8, pin_memory=True)      # caching host allocator
        stage.numpy()[:total] = np.frombuffer(blob, np.uint8)
        dstage = torch.empty((n * DESC_BYTES,), dtype=torch.uint8, pin_memory=True)
        dstage.numpy()[:] = raw.reshape(-1)
        data_d = stage.to(device, non_blocking=True)
        desc_d = dstage.to(device, non_blocking=True)
        coef = torch.empty((max(int(coef_blocks.value), 1) * 64,), dtype=torch.int16, device=device)
        err = torch.empty((n,), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            rc = _lib.lib().vip_jpeg_decode(data_d.data_ptr(), raw.ctypes.data, desc_d.data_ptr(), n, coef.data_ptr(),
                                            flat.data_ptr(), err.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vip_jpeg_decode")
    for i, px in host_pixels.items():
        off, h, w = layout[i]
        flat[off: off + h * w * 3].copy_(torch.from_numpy(np.array(px, copy=True)).reshape(-1), non_blocking=False)
    return DecodedBatch(flat, layout, n_dev, len(host_pixels), err)
