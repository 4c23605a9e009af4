"""Data pipeline with the surface of the reference's ``dataset/dataset.py`` (seeding 12-17, build_decoder 22-48,
build_augmenter 50-59, build_dataset 64-102).

What changed underneath: tf.io.read_file / tf.image.decode_jpeg run on host threads through Pillow (the same
libjpeg-turbo defaults TF uses: ISLOW IDCT, fancy upsampling) into pinned uint8 staging buffers; everything after the
decode -- cast, bicubic resize, /255, the augmentations -- is ONE fused CUDA kernel (vip_preprocess) that writes the
[B,H,W,3] batch in device memory, where the backbones consume it.  tf.data's RNG cannot be reproduced, so the TTA
decisions of apply_augment (dataset/augment.py:153-182: p=0.8 gate, hflip 0.5, vflip 0.5, gray 0.3) are drawn from a
numpy Generator seeded with CFG.seed and passed to the kernel as explicit per-image flags."""
from __future__ import annotations

import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops


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

    def decode_with_labels(path, label):
        return decode(path), label

    return decode_with_labels if with_labels else decode


def draw_augment_flags(n, rng):
    """apply_augment (dataset/augment.py:153-182) as explicit decisions: returns uint8 flags [n]."""
    gate = rng.random(n) <= 0.80                      # `if random_float() > augment_prob: return image`
    h = (rng.random(n) < 0.5) & gate                  # RandomFlip prob_hflip=0.5
    v = (rng.random(n) < 0.5) & gate                  # prob_vflip=0.5
    g = (rng.random(n) < 0.3) & gate                  # RandomGray prob=0.3
    return (h * ops.FLAG_HFLIP + v * ops.FLAG_VFLIP + g * ops.FLAG_GRAY).astype(np.uint8)


def build_augmenter(with_labels=True, img_size=(200, 200), CFG=None):
    """Returns ``augment(n) -> flags`` drawing the per-image decisions for one pass over ``n`` images."""
    def augment(n):
        rng = getattr(CFG, "_rng", None) or np.random.default_rng(getattr(CFG, "seed", 42))
        return draw_augment_flags(n, rng)

    return augment


class DeviceDataset:
    """Iterable of device batches [B,H,W,3] (bf16 by default).  ``repeat`` / ``steps`` semantics of tf.data are
    replaced by explicit passes: iterating yields ceil(N/B) batches of one pass; call again for the next TTA pass."""

    def __init__(self, paths, batch_size, img_size, decode_fn, augment_fn, augment, out_dtype, device, workers):
        self.paths, self.batch_size, self.img_size = list(paths), int(batch_size), tuple(int(v) for v in img_size)
        self.decode_fn, self.augment_fn, self.augment = decode_fn, augment_fn, augment
        self.out_dtype, self.device = out_dtype, device
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        # the prefetch task waits for the decode tasks: it needs a thread of its own (a single-worker pool would deadlock)
        self.prefetcher = ThreadPoolExecutor(max_workers=1)

    def __len__(self):
        return -(-len(self.paths) // self.batch_size)

    def _decode_batch(self, paths):
        imgs = list(self.pool.map(self.decode_fn, paths))
        return imgs

    def host_batches(self, depth=2):
        """One pass over the images as (first index, [decoded uint8 HxWx3 arrays]) per batch; the next ``depth`` batches
        are being decoded on the thread pool while the caller works on the current one (tf.data prefetch,
        dataset/dataset.py:100)."""
        n = len(self.paths)
        starts = list(range(0, n, self.batch_size))
        pending = [self.prefetcher.submit(self._decode_batch, self.paths[i0: i0 + self.batch_size]) for i0 in starts[:depth]]
        for k, i0 in enumerate(starts):
            imgs = pending.pop(0).result()
            if k + depth < len(starts):
                j0 = starts[k + depth]
                pending.append(self.prefetcher.submit(self._decode_batch, self.paths[j0: j0 + self.batch_size]))
            yield i0, imgs

    def __iter__(self):
        for _, imgs in self.host_batches(depth=1):
            flags = self.augment_fn(len(imgs)) if self.augment else None
            yield self._to_device(imgs, flags)

    def stage(self, imgs):
        """Decoded images of ONE size -> pinned uint8 [n,Hs,Ws,3] -> device (async copy on the current stream); None when
        the sizes differ (the caller then takes the per-size path of ``to_device``)."""
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
        # group by decoded size (test images may have other dimensions, dataset.py:32-34)
        groups = {}
        for i, im in enumerate(imgs):
            groups.setdefault(im.shape[:2], []).append(i)
        for (hs, ws), idx in groups.items():
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
                  out_dtype=torch.bfloat16, device=None, workers=None):
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
                         workers or min(32, os.cpu_count() or 4))
