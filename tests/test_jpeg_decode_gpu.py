"""GPU parity of the device JPEG decoder (csrc/jpeg_decode.cu through the C ABI) with libjpeg-turbo, bit for bit: the
committed golden files, files encoded here and decoded by Pillow in the same process, and the oracle."""
import io
import os

import numpy as np
import pytest

from test_oracle_jpeg_decode import golden_files, sha

pytestmark = pytest.mark.gpu


def _enc(img, **kw):
    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def _pillow(f):
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))


def test_golden_files_bit_exact(cuda_device):
    import torch

    from vipcup_b200 import jpeg

    items = list(golden_files())
    batch = jpeg.decode_batch([f for _, f, _, _ in items], device=cuda_device)     # one mixed-size, mixed-sampling batch
    torch.cuda.synchronize()
    batch.check()
    assert batch.n_host == 1 and batch.n_device == len(items) - 1                  # the progressive file took the host path
    for k, (i, f, out, digest) in enumerate(items):
        got = batch.image(k).cpu().numpy()
        assert sha(got) == digest, f"file {i}: shape {got.shape}"
        if out is not None:
            assert np.array_equal(got, out)


@pytest.mark.parametrize("kw", [dict(quality=75), dict(quality=95, subsampling=0), dict(quality=50, subsampling=1),
                                dict(quality=90, optimize=True), dict(quality=85, restart_marker_blocks=5),
                                dict(quality=100), dict(quality=5)])
def test_200x200_batch_matches_pillow(cuda_device, kw):
    """The shape main.py sees: a batch of 200x200 files -> one [N,200,200,3] tensor."""
    import torch

    from oracle import preprocess as P
    from vipcup_b200 import jpeg

    files = [_enc(P.synth_image(900 + i), **kw) for i in range(37)]
    batch = jpeg.decode_batch(files, device=cuda_device)
    x = batch.stacked()
    torch.cuda.synchronize()
    batch.check()
    assert batch.n_host == 0 and tuple(x.shape) == (37, 200, 200, 3)
    ref = np.stack([_pillow(f) for f in files])
    assert np.array_equal(x.cpu().numpy(), ref)


def test_ragged_sizes_and_grey(cuda_device):
    import torch

    from oracle import jpeg_decode as J
    from oracle import preprocess as P
    from vipcup_b200 import jpeg

    rng = np.random.default_rng(5)
    files = []
    for k, (h, w) in enumerate([(1, 1), (7, 9), (8, 8), (16, 16), (17, 33), (100, 3), (3, 100), (255, 257), (640, 48), (48, 900)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if k % 2 else P.synth_image(k, max(h, 16), max(w, 16))[:h, :w]
        files += [_enc(img, quality=70 + k), _enc(img, quality=90, subsampling=0), _enc(img, quality=80, subsampling=1),
                  _enc(img[:, :, 0], quality=85)]
    batch = jpeg.decode_batch(files, device=cuda_device)
    torch.cuda.synchronize()
    batch.check()
    assert batch.n_host == 0
    for k, f in enumerate(files):
        got = batch.image(k).cpu().numpy()
        assert np.array_equal(got, _pillow(f)), k
        if got.shape[0] * got.shape[1] <= 64 * 64:
            assert np.array_equal(got, J.decode(f)), k


def test_corrupt_stream_is_flagged_not_fatal(cuda_device):
    import torch

    from oracle import preprocess as P
    from vipcup_b200 import _lib, jpeg

    good = _enc(P.synth_image(1), quality=80)
    d = jpeg.parse(good)
    cut = good[: d.scan_offset + d.scan_bytes // 3] + b"\xff\xd9"          # truncated entropy-coded segment
    batch = jpeg.decode_batch([good, cut, good], device=cuda_device)
    torch.cuda.synchronize()
    assert np.array_equal(batch.image(0).cpu().numpy(), _pillow(good)) and np.array_equal(batch.image(2).cpu().numpy(), _pillow(good))
    got = batch.image(1).cpu().numpy()                                      # the rows decoded before the cut are right
    assert np.array_equal(got[:32], _pillow(good)[:32])
    assert batch.err.cpu().tolist() == [0, 1, 0]                            # bits consumed past the end of the scan
    with pytest.raises(_lib.VipError):
        batch.check()
    assert jpeg.decode_batch([], device=cuda_device).flat.numel() == 1      # empty batch


def test_photos_decode_and_dataset_path(cuda_device, tmp_path):
    """build_dataset with device decode: files on disk -> [B,H,W,3] batches equal to the host-decode path, bit for bit."""
    import torch

    from vipcup_b200.config import Config
    from vipcup_b200.dataset import build_dataset

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "photos.npz"))
    paths = []
    for i, a in enumerate(z["photos"]):
        for q in (70, 95):
            p = str(tmp_path / f"p{i}_{q}.jpg")
            with open(p, "wb") as f:
                f.write(_enc(a, quality=q))
            paths.append(p)
    outs = {}
    for mode in ("1", "0"):
        os.environ["VIP_JPEG_DEVICE"] = mode
        try:
            CFG = Config({"img_size": [224, 224], "seed": 42})
            ds = build_dataset(paths, batch_size=4, CFG=CFG, augment=False, out_dtype=torch.float32, device=cuda_device)
            outs[mode] = torch.cat([b.clone() for b in ds]).cpu().numpy()
        finally:
            os.environ.pop("VIP_JPEG_DEVICE", None)
    assert outs["1"].shape == (6, 224, 224, 3)
    assert np.array_equal(outs["1"].view(np.uint32), outs["0"].view(np.uint32))


def test_dataset_raises_for_corrupt_file(cuda_device, tmp_path):
    """A truncated entropy-coded segment: the device decoder flags it and the dataset raises after the pass, as
    tf.image.decode_jpeg would for the file."""
    import torch

    from oracle import preprocess as P
    from vipcup_b200.config import Config
    from vipcup_b200.dataset import build_dataset
    from vipcup_b200 import jpeg

    good = _enc(P.synth_image(2), quality=80)
    d = jpeg.parse(good)
    paths = []
    for i, data in enumerate([good, good[: d.scan_offset + d.scan_bytes // 2] + b"\xff\xd9", good]):
        p = str(tmp_path / f"f{i}.jpg")
        with open(p, "wb") as f:
            f.write(data)
        paths.append(p)
    ds = build_dataset(paths, batch_size=8, CFG=Config({"img_size": [200, 200], "seed": 1}), augment=False, device=cuda_device)
    for _ in ds:
        pass
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="f1.jpg"):
        ds.check_decode_errors()
