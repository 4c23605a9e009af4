"""CPU tests of the preprocessing oracle: pinned against libjpeg-turbo (live through Pillow and through the committed
vectors) and against the known-answer values of SURVEY.md A.2.  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import preprocess as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_jpeg_golden_vectors_from_libjpeg_turbo():
    z = np.load(os.path.join(GOLD, "jpeg_pillow.npz"))
    for i in range(int(z["n"])):
        got = P.jpeg_round_trip_u8(z[f"in{i}"], int(z["q"][i]))
        assert np.array_equal(got, z[f"out{i}"]), f"vector {i} (q={int(z['q'][i])}) differs from libjpeg-turbo"


@pytest.mark.parametrize("hw", [(224, 224), (200, 200), (199, 201), (50, 77), (17, 33), (1, 1), (3, 250)])
@pytest.mark.parametrize("q", [1, 50, 65, 85, 99, 100])
def test_jpeg_integer_restatement_matches_pillow_live(hw, q):
    pytest.importorskip("PIL")
    h, w = hw
    rng = np.random.default_rng(h * 1000 + w + q)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q % 2 else P.synth_image(q, max(h, 8), max(w, 8))[:h, :w]
    assert np.array_equal(P.jpeg_round_trip_u8(img, q), P.jpeg_round_trip_pillow(img, q))


def test_quality_100_tables_are_all_ones():
    l, c = P.quant_tables(100)
    assert (l == 1).all() and (c == 1).all()
    l50, _ = P.quant_tables(50)
    assert l50[0, 0] == 16 and l50[7, 7] == 99


def test_bicubic_known_answers():
    idx, w = P.bicubic_taps(200, 224)
    assert idx[0].tolist() == [0, 0, 0, 1]
    np.testing.assert_allclose(w[0], np.array([0, 0, 1.0248182, -0.02481813], dtype=np.float32), rtol=2e-7, atol=1e-9)
    assert idx[1].tolist() == [0, 0, 1, 2]
    np.testing.assert_allclose(w[1], np.array([0, 0.12485881, 0.93122476, -0.05608368], dtype=np.float32), rtol=2e-7, atol=1e-9)
    assert idx[223].tolist() == [198, 199, 199, 199]
    np.testing.assert_allclose(w[223], np.array([-0.02481813, 1.0248182, 0, 0], dtype=np.float32), rtol=2e-7, atol=1e-9)


def test_identity_resize_is_bit_exact():
    x = np.random.default_rng(0).integers(0, 256, (200, 200, 3), dtype=np.uint8).astype(np.float32)
    assert np.array_equal(P.resize_bicubic(x, 200, 200), x)


def test_resize_is_not_clamped_and_close_to_pillow_bicubic():
    from PIL import Image

    x = np.random.default_rng(1).integers(0, 256, (200, 200), dtype=np.uint8)
    y = P.resize_bicubic(x.astype(np.float32)[..., None], 224, 224)[..., 0]
    assert y.min() < 0 and y.max() > 255
    ref = np.asarray(Image.fromarray(x.astype(np.float32), "F").resize((224, 224), Image.BICUBIC))
    assert np.abs(y - ref).max() < 0.5  # table-quantised weights vs exact weights


def test_float_to_u8_saturate():
    x = np.array([-0.1, 0.0, 0.5, 1.0, 1.2, 254.9 / 255.5], dtype=np.float32)
    assert P.float_to_u8_saturate(x).tolist() == [0, 0, 127, 255, 255, 254]


def test_preprocess_golden_fixture():
    z = np.load(os.path.join(GOLD, "preprocess_small.npz"))
    ho, wo = (int(v) for v in z["out_hw"])
    out = P.preprocess_batch(z["src"], ho, wo, z["crops"], z["q"], z["flags"])
    assert np.array_equal(out.view(np.uint32), z["out"].view(np.uint32))


def test_flags_semantics():
    x = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3) / 18
    assert np.array_equal(P.apply_flags(x, P.FLAG_HFLIP), x[:, ::-1])
    assert np.array_equal(P.apply_flags(x, P.FLAG_VFLIP), x[::-1])
    g = P.apply_flags(x, P.FLAG_GRAY)
    assert np.array_equal(g[..., 0], g[..., 1]) and np.array_equal(g[..., 1], g[..., 2])
