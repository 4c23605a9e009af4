"""GPU parity: the CUDA preprocessing path (through the C ABI) versus the oracle, bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(dev, src, out_hw, crops=None, q=None, flags=None, dtype=None):
    import torch

    from vipcup_b200 import ops

    t = lambda a, dt: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    out = ops.preprocess(t(src, torch.uint8), out_hw, t(crops, torch.int32), t(q, torch.int32), t(flags, torch.uint8),
                         out_dtype=dtype or torch.float32)
    torch.cuda.synchronize()
    return out


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_library_loaded_is_in_tree():
    from vipcup_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH)
    assert b"sm_100a" in _lib.lib().vip_version()


def test_div255_sequence_is_ieee_exact(cuda_device):
    from vipcup_b200 import ops

    assert ops.selftest_div255() == 0


def test_golden_fixture(cuda_device):
    z = np.load(os.path.join(GOLD, "preprocess_small.npz"))
    out = _run(cuda_device, z["src"], tuple(int(v) for v in z["out_hw"]), z["crops"], z["q"], z["flags"])
    assert _bits_equal(out.cpu().numpy(), z["out"])


def test_jpeg_golden_vectors_from_libjpeg_turbo(cuda_device):
    """u8 -> (identity resize, /255, *255.5 trunc is the identity on u8) -> JPEG round trip -> *1/255; compared with
    libjpeg-turbo's own output scaled the same way."""
    z = np.load(os.path.join(GOLD, "jpeg_pillow.npz"))
    for i in range(int(z["n"])):
        img = z[f"in{i}"]
        h, w, _ = img.shape
        out = _run(cuda_device, img[None], (h, w), None, np.array([int(z["q"][i])]), None).cpu().numpy()[0]
        ref = z[f"out{i}"].astype(np.float32) * np.float32(1.0 / 255.0)
        assert _bits_equal(out, ref), f"vector {i}"


@pytest.mark.parametrize("case", [
    dict(hs=200, ws=200, ho=224, wo=224, n=24, crop=True, jpeg=True, flags=True),    # BASELINE config 2 shape
    dict(hs=200, ws=200, ho=200, wo=200, n=8, crop=False, jpeg=False, flags=False),  # main.py 200x200 models
    dict(hs=200, ws=200, ho=200, wo=200, n=8, crop=False, jpeg=True, flags=True),    # 200: 12.5 MCUs, edge padding
    dict(hs=200, ws=200, ho=224, wo=224, n=8, crop=False, jpeg=False, flags=True),   # GCViT 224 path of main.py
    dict(hs=97, ws=131, ho=75, wo=53, n=6, crop=True, jpeg=True, flags=True),        # ragged, downscale, odd sizes
    dict(hs=64, ws=64, ho=256, wo=256, n=3, crop=True, jpeg=True, flags=True),       # maximum JPEG plane size
    dict(hs=31, ws=17, ho=9, wo=7, n=5, crop=False, jpeg=True, flags=True),          # smaller than one MCU row pair
    dict(hs=300, ws=260, ho=200, wo=200, n=4, crop=True, jpeg=False, flags=True),    # other source sizes (dataset.py:33)
    dict(hs=5, ws=4, ho=1, wo=2, n=3, crop=False, jpeg=True, flags=True),            # degenerate: one chroma sample
    dict(hs=33, ws=35, ho=40, wo=24, n=4, crop=True, jpeg=True, flags=True),         # row pitch not a multiple of 4 (byte staging)
])
def test_parity_with_oracle(cuda_device, case):
    from oracle import preprocess as P

    n, hs, ws, ho, wo = case["n"], case["hs"], case["ws"], case["ho"], case["wo"]
    rng = np.random.default_rng(hs * 7 + ws * 3 + ho + n)
    src = np.stack([P.synth_image(i, hs, ws) if i % 3 else rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
                    for i in range(n)])
    crops = q = flags = None
    if case["crop"]:
        h = rng.integers(max(1, hs // 2), hs + 1, n)
        w = rng.integers(max(1, ws // 2), ws + 1, n)
        y0 = (rng.random(n) * (hs - h + 1)).astype(int)
        x0 = (rng.random(n) * (ws - w + 1)).astype(int)
        crops = np.stack([y0, x0, h, w], 1).astype(np.int32)
    if case["jpeg"]:
        q = rng.integers(65, 101, n).astype(np.int32)
        q[::4] = -1 if n > 4 else q[::4]  # mixed batch: some images skip the JPEG stage
        q[-1] = 100
        if n > 2:
            q[1] = 7                       # very low quality: large quantisers
    if case["flags"]:
        flags = rng.integers(0, 8, n).astype(np.uint8)
    got = _run(cuda_device, src, (ho, wo), crops, q, flags).cpu().numpy()
    ref = P.preprocess_batch(src, ho, wo, crops, q, flags)
    bad = np.argwhere(got.view(np.uint32) != ref.view(np.uint32))
    assert bad.size == 0, f"{len(bad)} mismatching elements, first at {bad[0]}: {got[tuple(bad[0])]} vs {ref[tuple(bad[0])]}"


def test_bf16_output_is_rounded_f32(cuda_device):
    import torch

    from oracle import preprocess as P

    src = np.stack([P.synth_image(i, 200, 200) for i in range(4)])
    crops, q, flags = P.synth_decisions(4)
    f32 = _run(cuda_device, src, (224, 224), crops, q, flags)
    b16 = _run(cuda_device, src, (224, 224), crops, q, flags, dtype=torch.bfloat16)
    assert torch.equal(f32.to(torch.bfloat16), b16)


def test_empty_batch_and_errors(cuda_device):
    import torch

    from vipcup_b200 import ops

    out = ops.preprocess(torch.zeros((0, 200, 200, 3), dtype=torch.uint8, device=cuda_device), (224, 224))
    assert out.shape == (0, 224, 224, 3)
    with pytest.raises(ops.VipError):  # JPEG planes larger than shared memory must fail loudly, not fall back
        ops.preprocess(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device=cuda_device), (512, 512),
                       jpeg_q=torch.full((1,), 90, dtype=torch.int32, device=cuda_device))


def test_full_size_properties(cuda_device):
    """BASELINE config 2 at full size (4096 x 200x200 -> 224x224): size-independent checks.
    * every image equals the oracle on a random sample of 32 images,
    * flips commute with the rest of the path: out(flags=f)[n] == flip_f(out(flags=0)[n]),
    * q = -1 for all images equals the no-JPEG launch,
    * determinism: two launches are bit-identical."""
    import torch

    from oracle import preprocess as P

    n = 4096
    base = np.stack([P.synth_image(i, 200, 200) for i in range(64)])
    src = np.tile(base, (n // 64, 1, 1, 1))
    crops, q, flags = P.synth_decisions(n)
    out = _run(cuda_device, src, (224, 224), crops, q, flags)
    out2 = _run(cuda_device, src, (224, 224), crops, q, flags)
    assert torch.equal(out, out2)
    plain = _run(cuda_device, src, (224, 224), crops, q, None)
    h = torch.from_numpy((flags & 1).astype(bool)).to(cuda_device)
    v = torch.from_numpy((flags & 2).astype(bool)).to(cuda_device)
    exp = torch.where(h[:, None, None, None], plain.flip(2), plain)
    exp = torch.where(v[:, None, None, None], exp.flip(1), exp)
    assert torch.equal(out, exp)
    nq = _run(cuda_device, src[:256], (224, 224), crops[:256], np.full(256, -1, np.int32), flags[:256])
    nj = _run(cuda_device, src[:256], (224, 224), crops[:256], None, flags[:256])
    assert torch.equal(nq, nj)
    pick = np.random.default_rng(5).choice(n, 32, replace=False)
    got = out[torch.from_numpy(pick).to(cuda_device)].cpu().numpy()
    ref = P.preprocess_batch(src[pick], 224, 224, crops[pick], q[pick], flags[pick])
    assert _bits_equal(got, ref)


def test_host_buffer_entry_point(cuda_device):
    import torch

    from oracle import preprocess as P
    from vipcup_b200 import ops

    n = 37
    src = np.stack([P.synth_image(i, 200, 200) for i in range(n)])
    crops, q, flags = P.synth_decisions(n)
    out = ops.preprocess_host(torch.from_numpy(src).pin_memory(), (224, 224), torch.from_numpy(crops).pin_memory(),
                              torch.from_numpy(q).pin_memory(), torch.from_numpy(flags).pin_memory())
    ref = P.preprocess_batch(src, 224, 224, crops, q, flags)
    assert _bits_equal(out.numpy(), ref)


@pytest.mark.parametrize("hs,ws,ho,wo,n", [
    (200, 200, 224, 224, 40),   # GCViT path of main.py (streaming kernel, 14 stripes of 16 rows)
    (200, 200, 200, 200, 40),   # ResNet-RS path: identity fast path for unflagged images, tap path for flagged ones
    (180, 220, 200, 200, 5),    # other source sizes (dataset/dataset.py:33-34), down- and up-scale at once
    (512, 512, 200, 200, 3),    # strong downscale: many source rows per stripe
    (64, 64, 256, 256, 3),      # strong upscale: repeated taps, clamped borders
    (8, 16, 8, 8, 7),           # tiny
])
@pytest.mark.parametrize("bf16", [False, True])
def test_streaming_kernel_parity_with_oracle(cuda_device, hs, ws, ho, wo, n, bf16):
    """vip_preprocess without crop / JPEG emulation takes the streaming kernel (csrc/preprocess_stream.cu): bit-exact
    against the oracle for every flag combination (flips as addressing, gray), both output types; bf16 = RN of the f32."""
    import torch

    from oracle import preprocess as P

    rng = np.random.default_rng(hs + 3 * ws + ho)
    src = np.stack([P.synth_image(i, hs, ws) if i % 3 else rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
                    for i in range(n)])
    flags = (np.arange(n) % 8).astype(np.uint8)
    ref = P.preprocess_batch(src, ho, wo, None, None, flags)
    got = _run(cuda_device, src, (ho, wo), None, None, flags, dtype=torch.bfloat16 if bf16 else torch.float32)
    if bf16:
        assert torch.equal(got.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    else:
        bad = np.argwhere(got.cpu().numpy().view(np.uint32) != ref.view(np.uint32))
        assert bad.size == 0, f"{len(bad)} mismatching elements, first at {bad[0]}"
    # the streaming and the fused kernel agree: q = -1 everywhere forces the fused (JPEG-capable) kernel
    fused = _run(cuda_device, src, (ho, wo), None, np.full(n, -1, np.int32), flags, dtype=torch.bfloat16 if bf16 else torch.float32)
    assert torch.equal(got, fused)
    none = _run(cuda_device, src, (ho, wo), None, None, None, dtype=torch.bfloat16 if bf16 else torch.float32)
    assert torch.equal(none[::8], got[::8])     # flags == 0 images: same result with and without a flags array
