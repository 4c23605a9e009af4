"""Regenerates the committed golden fixtures.  Run from the repo root: ``python tests/golden/make_golden.py``.

* ``jpeg_pillow.npz``   -- inputs and outputs of the REAL libjpeg-turbo (through Pillow) for the 4:2:0 round trip:
                           pins oracle.preprocess.jpeg_round_trip_u8 (and through it the CUDA kernel) to the library
                           TensorFlow's tf.image.adjust_jpeg_quality calls.  Third-party dependency of the reference:
                           libjpeg-turbo bundled in TensorFlow (version unpinned by the reference); generated here
                           with Pillow 12.2.0 / libjpeg-turbo API "6.2".
* ``preprocess_small.npz`` -- inputs + oracle outputs of the whole preprocessing path on small ragged images.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preprocess as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(2022)
    # 1) libjpeg-turbo vectors
    imgs, qs, outs = [], [], []
    for k, (h, w) in enumerate([(48, 64), (33, 47), (16, 16), (8, 24), (40, 40), (25, 70)]):
        for q in (65, 80, 93, 100) if k < 4 else (30, 75):
            img = P.synth_image(100 + k, max(h, 16), max(w, 16))[:h, :w] if k % 2 == 0 else \
                rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            imgs.append(img)
            qs.append(q)
            outs.append(P.jpeg_round_trip_pillow(img, q))
    np.savez_compressed(os.path.join(HERE, "jpeg_pillow.npz"), n=len(imgs), q=np.array(qs),
                        **{f"in{i}": a for i, a in enumerate(imgs)}, **{f"out{i}": a for i, a in enumerate(outs)})
    # 2) whole preprocessing path, small
    n, hs, ws, ho, wo = 6, 56, 72, 64, 80
    src = np.stack([P.synth_image(200 + i, hs, ws) for i in range(n)])
    crops = np.array([[0, 0, 56, 72], [3, 5, 40, 50], [10, 2, 46, 70], [0, 20, 56, 52], [7, 7, 33, 41], [1, 1, 54, 70]],
                     dtype=np.int32)
    q = np.array([70, -1, 95, 100, 65, 88], dtype=np.int32)
    flags = np.array([0, 1, 2, 3, 4, 7], dtype=np.uint8)
    out = P.preprocess_batch(src, ho, wo, crops, q, flags)
    np.savez_compressed(os.path.join(HERE, "preprocess_small.npz"), src=src, crops=crops, q=q, flags=flags, out=out,
                        out_hw=np.array([ho, wo]))
    print("wrote fixtures:", os.listdir(HERE))


if __name__ == "__main__":
    main()
