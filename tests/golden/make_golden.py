"""Regenerates the committed golden fixtures.  Run from the repo root: ``python tests/golden/make_golden.py``.

* ``jpeg_pillow.npz``   -- inputs and outputs of the REAL libjpeg-turbo (through Pillow) for the 4:2:0 round trip:
                           pins oracle.preprocess.jpeg_round_trip_u8 (and through it the CUDA kernel) to the library
                           TensorFlow's tf.image.adjust_jpeg_quality calls.  Third-party dependency of the reference:
                           libjpeg-turbo bundled in TensorFlow (version unpinned by the reference); generated here
                           with Pillow 12.2.0 / libjpeg-turbo API "6.2".
* ``preprocess_small.npz`` -- inputs + oracle outputs of the whole preprocessing path on small ragged images.
* ``jpeg_files.npz``    -- JPEG FILES (bytes) and what the real libjpeg-turbo (through Pillow) decodes them to: pins
                           oracle.jpeg_decode and the device decoder (csrc/jpeg_decode.cu).  4:2:0 / 4:2:2 / 4:4:4 / grey,
                           odd sizes, optimised Huffman tables, restart intervals, a progressive file (host path), and
                           real photographs.
* ``photos.npz``        -- the only real images the reference holds: the three 512x512 photographs embedded in
                           models/keras_cv_attention_models/test_images.py:18-22, decoded with Pillow HERE (the reference
                           is importable in this container only) and cropped to 200x200 -- real-image inputs for the
                           preprocessing / backbone parity tests.  Written only when /root/reference is present.
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
    # 3) JPEG files + libjpeg-turbo's decode of them
    import io

    from PIL import Image

    def enc(img, **kw):
        b = io.BytesIO()
        Image.fromarray(img).save(b, "JPEG", **kw)
        return b.getvalue()

    files = []
    for k, (h, w) in enumerate([(200, 200), (64, 80), (33, 47), (17, 16), (8, 8), (1, 1), (50, 35), (100, 260)]):
        img = P.synth_image(300 + k, max(h, 16), max(w, 16))[:h, :w]
        files.append(enc(img, quality=75 + 3 * k))                                   # 4:2:0, standard tables
        files.append(enc(img, quality=92, subsampling=0))                            # 4:4:4
        files.append(enc(img, quality=60, subsampling=1))                            # 4:2:2
        files.append(enc(img, quality=88, optimize=True))                            # per-file Huffman tables
    img = P.synth_image(320, 72, 96)
    files.append(enc(img[:, :, 1], quality=80))                                      # grey
    files.append(enc(img, quality=85, restart_marker_blocks=3))                      # DRI / RSTn
    files.append(enc(img, quality=85, subsampling=0, restart_marker_rows=1))
    files.append(enc(rng.integers(0, 256, (48, 48, 3), dtype=np.uint8), quality=100))   # long codes, large coefficients
    files.append(enc(img, quality=85, progressive=True))                             # outside the device subset
    ref_mod = "/root/reference/models/keras_cv_attention_models/test_images.py"
    photos = None
    if os.path.exists(ref_mod):
        import importlib.util

        spec = importlib.util.spec_from_file_location("ref_test_images", ref_mod)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        full = [np.asarray(Image.fromarray(f()).convert("RGB")) for f in (mod.dog_cat, mod.cat, mod.dog)]
        photos = np.stack([a[156:356, 156:356] for a in full])                       # centre 200x200 crops
        np.savez_compressed(os.path.join(HERE, "photos.npz"), photos=photos)
        for a in photos:
            files.append(enc(a, quality=90))
            files.append(enc(a, quality=78, subsampling=0))
        files.append(enc(full[1], quality=85))                                       # 512x512: several bands per image
    elif os.path.exists(os.path.join(HERE, "jpeg_files.npz")):
        print("reference not present: keeping the photo entries of the existing jpeg_files.npz is not possible; aborting")
        return
    import hashlib

    outs = [np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files]
    # decoded pixels for the small files, SHA-256 of the pixels (shape-prefixed) for the large ones
    np.savez_compressed(os.path.join(HERE, "jpeg_files.npz"), n=len(files),
                        **{f"file{i}": np.frombuffer(f, np.uint8) for i, f in enumerate(files)},
                        **{f"out{i}": a for i, a in enumerate(outs) if a.size <= 100 * 260 * 3},
                        **{f"sha{i}": np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)
                           for i, a in enumerate(outs)})
    print("wrote fixtures:", os.listdir(HERE))


if __name__ == "__main__":
    main()
