"""Data pipeline with the surface of the reference's ``dataset/dataset.py`` (seeding 12-17, build_decoder 22-48,
build_augmenter 50-59, build_dataset 64-102).

What changed underneath: tf.io.read_file stays on host threads; tf.image.decode_jpeg runs ON THE DEVICE for baseline JPEG
files (vip_jpeg_decode: Huffman decode, integer IDCT, fancy upsampling, colour conversion, bit-identical to libjpeg-turbo;
the host only walks the marker segments) and through Pillow on host threads (the same libjpeg-turbo defaults TF uses:
ISLOW IDCT, fancy upsampling) for everything else (progressive files, PNG, a caller-supplied decode_fn, VIP_JPEG_DEVICE=0);
everything after the decode -- cast, bicubic resize, /255, the augmentations -- is ONE fused CUDA kernel (vip_preprocess)
that writes the [B,H,W,3] batch in device memory, where the backbones consume it.  tf.data's RNG stream cannot be reproduced, so
the TTA decisions of apply_augment (dataset/augment.py:153-182: p=0.8 gate, hflip 0.5, vflip 0.5, gray 0.3) are drawn from
a counter-based Philox-4x32-10 keyed by CFG.seed and indexed by (image position, TTA pass) -- independent of batching and
sharding -- and passed to the kernel as explicit per-image flags."""
from __future__ import annotations

import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import jpeg, ops


def seeding(CFG):
    """dataset/dataset.py:12-17 (the tf seed becomes the numpy Generator that draws the augmentation decisions)."""
    seed = CFG.seed
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    CFG._rng = np.random.default_rng(seed)


def build_decoder(with_labels, img_size, CFG, ext="jpg"):
    """Returns ``decode(path) -> uint8 [Hs,Ws,3]`` (host).  The float conversion / resize / normalisation that the
    reference does here (dataset.py:31-37) happen on the device inside build_dataset."""
    if ext not in ("jpg", "jpeg", "png"):
        raise ValueError("Image extension not supported")  # dataset.py:30

    def decode(path):
        from PIL import Image

        with Image.open(path) as im:
            return np.asarray(im.convert("RGB"))

    decode.device_decodable = ext in ("jpg", "jpeg")   # DeviceDataset may hand the file bytes to vip_jpeg_decode instead

    def decode_with_labels(path, label):
        return decode(path), label

    return decode_with_labels if with_labels else decode


def philox4x32_10(counter, key):
    """Philox-4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"; the generator behind
    tf.random.uniform, dataset/augment.py:11-19) on an array of counters: counter uint32 [n, 4], key (k0, k1) -> uint32
    [n, 4].  Counter-based: every output depends only on (counter, key), not on how many numbers were drawn before."""
    c = np.array(counter, dtype=np.uint64, copy=True).reshape(-1, 4)
    k0, k1 = np.uint64(int(key[0]) & 0xFFFFFFFF), np.uint64(int(key[1]) & 0xFFFFFFFF)
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    w0, w1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    for _ in range(10):
        p0, p1 = m0 * c[:, 0], m1 * c[:, 2]                       # 32 x 32 -> 64-bit products
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0, k1 = (k0 + w0) & mask, (k1 + w1) & mask
    return c.astype(np.uint32)


def uniform_from_bits(u32):
    """uint32 -> float32 in [0, 1) the way TF's random ops do it: 23 random mantissa bits of a float in [1, 2), minus 1."""
    return ((np.asarray(u32, np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)


def draw_augment_flags(indices, pass_idx, seed):
    """apply_augment (dataset/augment.py:153-182) as explicit per-image decisions, uint8 flags [n].

    TF's stateful RNG inside a parallel tf.data map cannot be reproduced; the decisions here come from the same generator
    family used as a pure function: Philox(counter = (image index in the CSV, TTA pass, 0, 0), key = (seed, 0)) gives the
    four uniforms of one image -- gate (`random_float() > 0.8` returns the image unchanged), hflip < 0.5, vflip < 0.5
    (RandomFlip, augment.py:115-120), gray < 0.3 (RandomGray, 142-146).  They do not depend on the batch size, the order of
    evaluation or the number of GPUs the list is sharded over."""
    idx = np.asarray(indices, dtype=np.uint64).reshape(-1)
    ctr = np.stack([idx & np.uint64(0xFFFFFFFF), np.full_like(idx, int(pass_idx)), idx >> np.uint64(32), np.zeros_like(idx)], 1)
    u = uniform_from_bits(philox4x32_10(ctr, (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)))
    gate = ~(u[:, 0] > np.float32(0.80))
    h = (u[:, 1] < np.float32(0.5)) & gate
    v = (u[:, 2] < np.float32(0.5)) & gate
    g = (u[:, 3] < np.float32(0.3)) & gate
    return (h * ops.FLAG_HFLIP + v * ops.FLAG_VFLIP + g * ops.FLAG_GRAY).astype(np.uint8)


def build_augmenter(with_labels=True, img_size=(200, 200), CFG=None):
    """Returns ``augment(indices, pass_idx) -> flags``: the decisions of one TTA pass for the images with those positions
    in the input CSV."""
    def augment(indices, pass_idx=0):
        return draw_augment_flags(indices, pass_idx, getattr(CFG, "seed", 42))

    return augment


class RawJpegs(list):
    """One batch of undecoded files: a list of (file bytes, parsed descriptor); the device decode is memoised so that every
    model of the ensemble consumes the same decoded pixels."""

    def decoded(self, device):
        if getattr(self, "_decoded", None) is None:
            self._decoded = jpeg.decode_batch([f for f, _ in self], [d for _, d in self], device=device)
            if self._decoded.err is not None and getattr(self, "err_sink", None) is not None:
                self.err_sink.append((getattr(self, "first_index", 0), self._decoded.err))   # read after the pass: no sync here
        return self._decoded


class DeviceDataset:
    """Iterable of device batches [B,H,W,3] (bf16 by default).  ``repeat`` / ``steps`` semantics of tf.data are
    replaced by explicit passes: iterating yields ceil(N/B) batches of one pass; call again for the next TTA pass."""

    def __init__(self, paths, batch_size, img_size, decode_fn, augment_fn, augment, out_dtype, device, workers,
                 index_offset=0):
        self.paths, self.batch_size, self.img_size = list(paths), int(batch_size), tuple(int(v) for v in img_size)
        self.index_offset, self.passes_done = int(index_offset), 0   # position of paths[0] in the whole list; TTA pass counter
        self.decode_fn, self.augment_fn, self.augment = decode_fn, augment_fn, augment
        self.out_dtype, self.device = out_dtype, device
        # JPEG decode on the device unless the caller brought a decoder of its own or switched it off
        self.device_decode = (getattr(decode_fn, "device_decodable", False) and os.environ.get("VIP_JPEG_DEVICE", "1") != "0")
        self._decode_errs = []      # (first image index, device int32 [n] flags) per device-decoded batch
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        # the prefetch task waits for the decode tasks: it needs a thread of its own (a single-worker pool would deadlock)
        self.prefetcher = ThreadPoolExecutor(max_workers=1)

    def __len__(self):
        return -(-len(self.paths) // self.batch_size)

    def _decode_batch(self, paths):
        if self.device_decode:
            return RawJpegs(self.pool.map(jpeg.read_and_parse, paths))     # file read + marker walk only
        imgs = list(self.pool.map(self.decode_fn, paths))
        return imgs

    def host_batches(self, depth=2):
        """One pass over the images as (first index, [decoded uint8 HxWx3 arrays] or RawJpegs) per batch; the next ``depth``
        batches are being read / decoded on the thread pool while the caller works on the current one (tf.data prefetch,
        dataset/dataset.py:100)."""
        n = len(self.paths)
        starts = list(range(0, n, self.batch_size))
        pending = [self.prefetcher.submit(self._decode_batch, self.paths[i0: i0 + self.batch_size]) for i0 in starts[:depth]]
        for k, i0 in enumerate(starts):
            imgs = pending.pop(0).result()
            if isinstance(imgs, RawJpegs):
                imgs.err_sink, imgs.first_index = self._decode_errs, i0
            if k + depth < len(starts):
                j0 = starts[k + depth]
                pending.append(self.prefetcher.submit(self._decode_batch, self.paths[j0: j0 + self.batch_size]))
            yield i0, imgs

    def check_decode_errors(self):
        """Raises for files whose entropy-coded data the device decoder flagged as corrupt (tf.image.decode_jpeg raises
        InvalidArgumentError for those, dataset/dataset.py:28); one device -> host read per batch, after the pass."""
        bad = []
        for i0, err in self._decode_errs:
            bad += [self.paths[i0 + j] for j in torch.nonzero(err).flatten().tolist()]
        self._decode_errs.clear()
        if bad:
            raise ValueError(f"corrupt JPEG data in {len(bad)} file(s): " + ", ".join(bad[:5]))

    def flags_for(self, i0, n, pass_idx):
        """Augmentation decisions of images [i0, i0 + n) of this dataset in TTA pass ``pass_idx``."""
        return self.augment_fn(self.index_offset + i0 + np.arange(n), pass_idx)

    def __iter__(self):
        pass_idx, self.passes_done = self.passes_done, self.passes_done + 1
        for i0, imgs in self.host_batches(depth=1):
            flags = self.flags_for(i0, len(imgs), pass_idx) if self.augment else None
            yield self._to_device(imgs, flags)

    def stage(self, imgs):
        """Decoded images of ONE size -> pinned uint8 [n,Hs,Ws,3] -> device (async copy on the current stream); None when
        the sizes differ (the caller then takes the per-size path of ``to_device``).  Undecoded files: H2D of the file
        bytes and the two decode kernels instead."""
        if isinstance(imgs, RawJpegs):
            dec = imgs.decoded(self.device)
            return dec.stacked() if dec.uniform_shape() is not None else None
        shape = imgs[0].shape
        if any(im.shape != shape for im in imgs):
            return None
        st = torch.empty((len(imgs), *shape), dtype=torch.uint8).pin_memory()
        np.stack(imgs, out=st.numpy())
        return st.to(self.device, non_blocking=True)

    def to_device(self, imgs, flags, img_size=None):
        return self._to_device(imgs, flags, img_size)

    def _to_device(self, imgs, flags, img_size=None):
        h, w = img_size or self.img_size
        out = torch.empty((len(imgs), h, w, 3), dtype=self.out_dtype, device=self.device)
        dec = imgs.decoded(self.device) if isinstance(imgs, RawJpegs) else None
        # group by decoded size (test images may have other dimensions, dataset.py:32-34)
        groups = {}
        for i in range(len(imgs)):
            groups.setdefault(tuple(dec.layout[i][1:3]) if dec is not None else imgs[i].shape[:2], []).append(i)
        for (hs, ws), idx in groups.items():
            if dec is not None:
                src = dec.stacked() if len(groups) == 1 else torch.stack([dec.image(i) for i in idx])
            else:
                stage = torch.empty((len(idx), hs, ws, 3), dtype=torch.uint8).pin_memory()
                np.stack([imgs[i] for i in idx], out=stage.numpy())
                src = stage.to(self.device, non_blocking=True)
            fl = None
            if flags is not None:
                fl = torch.from_numpy(np.ascontiguousarray(flags[idx])).to(self.device)
            res = ops.preprocess(src, (h, w), None, None, fl, out_dtype=self.out_dtype)
            if len(groups) == 1:
                return res
            out[torch.as_tensor(idx, device=self.device)] = res
        return out


def build_dataset(paths, labels=None, batch_size=32, cache=True, decode_fn=None, augment_fn=None, dim=(200, 200),
                  augment=True, repeat=True, shuffle=1024, cache_dir="", drop_remainder=False, CFG=None,
                  out_dtype=torch.bfloat16, device=None, workers=None, index_offset=0):
    """Same arguments as dataset/dataset.py:64-102 (labels / cache / shuffle / repeat only matter for training and are
    accepted and ignored on this inference path)."""
    if labels is not None:
        raise NotImplementedError("the training branch of build_dataset is out of scope (inference hot path only)")
    if decode_fn is None:
        decode_fn = build_decoder(False, img_size=CFG.img_size, CFG=CFG)
    if augment_fn is None:
        augment_fn = build_augmenter(False, img_size=CFG.img_size, CFG=CFG)
    device = device or torch.device("cuda", torch.cuda.current_device())
    return DeviceDataset(paths, batch_size, CFG.img_size, decode_fn, augment_fn, augment, out_dtype, device,
                         workers or min(32, os.cpu_count() or 4), index_offset=index_offset)
