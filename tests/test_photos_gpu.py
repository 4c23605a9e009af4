"""The only real images the reference holds (three photographs, models/keras_cv_attention_models/test_images.py:18-22;
200x200 crops committed as tests/golden/photos.npz) through the whole device path -- JPEG file -> device decode ->
preprocessing -> backbone -- against the oracle pipeline (Pillow decode -> oracle preprocess -> fp32 oracle backbone).
Real photographs have the smooth regions, edges and saturated pixels the blur + noise synthetic images lack."""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _files():
    from PIL import Image

    z = np.load(os.path.join(GOLD, "photos.npz"))
    out = []
    for a in z["photos"]:
        for q in (92, 75):
            b = io.BytesIO()
            Image.fromarray(a).save(b, "JPEG", quality=q)
            out.append(b.getvalue())
    return out


@pytest.mark.parametrize("arch", ["rs50", "gcvit_tiny"])
def test_photos_end_to_end_vs_oracle(cuda_device, arch):
    import torch
    from PIL import Image

    from oracle import preprocess as P
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200 import jpeg, ops

    files = _files()
    hw = 200 if arch == "rs50" else 224
    # oracle side: libjpeg (Pillow) decode -> restated bicubic / 255
    dec = [np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files]
    x_ref = np.stack([P.decode_to_float(d, hw, hw) for d in dec])
    # device side: Huffman / IDCT on the GPU -> fused preprocessing (f32, compared bit for bit with the oracle's)
    batch = jpeg.decode_batch(files, device=cuda_device)
    src = batch.stacked()
    assert np.array_equal(src.cpu().numpy(), np.stack(dec))
    x32 = ops.preprocess(src, (hw, hw), out_dtype=torch.float32)
    assert np.array_equal(x32.cpu().numpy().view(np.uint32), x_ref.view(np.uint32))
    x = ops.preprocess(src, (hw, hw), out_dtype=torch.bfloat16)
    ref_taps, taps = {}, {}
    if arch == "rs50":
        from oracle import resnet_rs as R
        from vipcup_b200.models import ResNetRS

        W = R.random_weights(50, 2, seed=11)
        ref = R.forward(x_ref, W, 50, head_act="softmax", taps=ref_taps)
        model = ResNetRS(50, classes=2, classifier_activation="softmax", device=cuda_device).load_weights(W)
        got = model(x, taps=taps)
        head, stages = ("predictions/kernel", "predictions/bias"), ("stem", "c2", "c3", "c4", "c5")
    else:
        from oracle import gcvit as G
        from vipcup_b200.models import GCViT

        W = G.random_weights("tiny", 2, seed=11)
        ref = G.forward(x_ref, W, "tiny", head_act="softmax", taps=ref_taps)
        model = GCViT("tiny", input_shape=(hw, hw, 3), num_classes=2, head_act="softmax", device=cuda_device).load_weights(W)
        got = model(x, taps=taps)
        head, stages = ("head/kernel", "head/bias"), ("stem", "level0", "level1", "level2", "level3")
    torch.cuda.synchronize()
    check_against_oracle(ref, ref_taps, got, taps, W[head[0]], W[head[1]], stages, logit_tol=1e-2)
