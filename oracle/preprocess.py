"""CPU ORACLE (test infrastructure, NOT a product path) for the preprocessing half of the
vip-cup-2022 inference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The shipped path (``vipcup_b200``) never does and fails loudly
without its CUDA library.

What is restated (reference file:line -> function here):

* ``dataset/dataset.py:31-37``  tf.cast(float32) -> tf.image.resize(method='bicubic') -> ``/ 255.0``
      -> :func:`bicubic_taps`, :func:`resize_bicubic`, :func:`decode_to_float`
* ``dataset/augment.py:110-113``  tf.image.random_jpeg_quality (adjust_jpeg_quality)
      -> :func:`jpeg_round_trip_u8` (pure-integer libjpeg baseline 4:2:0 encode->decode) wrapped by
         :func:`adjust_jpeg_quality`
* ``dataset/augment.py:115-120``  tf.image.flip_left_right / flip_up_down      -> :func:`apply_flags`
* ``dataset/augment.py:142-146``  rgb_to_grayscale -> grayscale_to_rgb         -> :func:`apply_flags`
* ``models/keras_cv_attention_models/imagenet/data.py:56-63`` random-crop -> resize semantic
      (slice, then resize the slice)                                           -> :func:`preprocess_one`

The arithmetic of these ops lives in un-vendored third-party code (TensorFlow, version unpinned by
the reference; libjpeg-turbo bundled inside TF).  TensorFlow is not installable here, so:

PARITY STATUS
  * JPEG round trip: pinned against libjpeg-turbo itself (Pillow 12.2 / libjpeg-turbo "6.2" API in this
    image) -- tests/test_oracle_preprocess.py requires 0 mismatching pixels.
  * bicubic resize / ``/255`` / gray: **parity unpinned** versus real TensorFlow (no TF offline, the
    reference has no tests or golden vectors).  The restatement follows TF's ResizeBicubic CPU kernel
    (``half_pixel_centers=True``, Keys a=-0.5, 1024-entry table, border renormalisation, vertical then
    horizontal, fp32, no FMA contraction, no clamp).  Known-answer values recorded in SURVEY.md A.2 are
    asserted in the tests.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------------------------
# tf.image.resize(method="bicubic", antialias=False) == ResizeBicubic(half_pixel_centers=True)
# ----------------------------------------------------------------------------------------------
_TABLE_SIZE = 1024
_KEYS_A = -0.5


def _coeff_table() -> np.ndarray:
    """Weight LUT, (1025, 2) fp32: [:,0] for |x|<=1 ("near"), [:,1] for x+1 ("far").

    Computed in double from a float abscissa and rounded to fp32 once, like the TF table initialiser.
    """
    a = _KEYS_A
    i = np.arange(_TABLE_SIZE + 1, dtype=np.float64)
    x = (i / _TABLE_SIZE).astype(np.float32).astype(np.float64)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    x1 = (x.astype(np.float32) + np.float32(1.0)).astype(np.float64)
    far = ((a * x1 - 5.0 * a) * x1 + 8.0 * a) * x1 - 4.0 * a
    return np.stack([near, far], axis=1).astype(np.float32)


COEFF_TABLE = _coeff_table()


def bicubic_taps(in_size: int, out_size: int):
    """Per output coordinate: 4 clamped source indices (int32) and 4 fp32 weights.

    scale = float(in)/float(out); src = (o + 0.5) * scale - 0.5 (all fp32); offset = lrintf(delta*1024);
    taps whose unclamped index fell outside [0, in) get weight 0; weights renormalised by 1/sum.
    """
    scale = F32(in_size) / F32(out_size)
    o = np.arange(out_size, dtype=np.float32)
    in_loc_f = (o + F32(0.5)) * scale - F32(0.5)            # fp32 throughout
    in_loc = np.floor(in_loc_f).astype(np.int64)
    delta = in_loc_f - in_loc.astype(np.float32)
    offset = np.rint(delta * F32(_TABLE_SIZE)).astype(np.int64)   # lrintf: round-half-even
    limit = in_size - 1
    raw = np.stack([in_loc - 1, in_loc, in_loc + 1, in_loc + 2], axis=1)
    idx = np.clip(raw, 0, limit)
    w = np.stack(
        [
            COEFF_TABLE[offset, 1],
            COEFF_TABLE[offset, 0],
            COEFF_TABLE[_TABLE_SIZE - offset, 0],
            COEFF_TABLE[_TABLE_SIZE - offset, 1],
        ],
        axis=1,
    ).astype(np.float32)
    w = np.where(idx == raw, w, F32(0.0)).astype(np.float32)
    wsum = ((w[:, 0] + w[:, 1]) + w[:, 2]) + w[:, 3]
    tiny = F32(1000.0) * np.finfo(np.float32).tiny
    inv = np.where(np.abs(wsum) >= tiny, F32(1.0) / wsum, F32(1.0)).astype(np.float32)
    w = (w * inv[:, None]).astype(np.float32)
    return idx.astype(np.int32), w


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """fp32 [h,w,c] -> fp32 [out_h,out_w,c]; vertical taps first, then horizontal, each product and
    each sum rounded to fp32 separately, summed left to right (TF Interpolate1D order)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w, _ = img.shape
    iy, wy = bicubic_taps(h, out_h)
    ix, wx = bicubic_taps(w, out_w)
    # vertical: v[oy, x, c]
    v = img[iy[:, 0]] * wy[:, 0, None, None]
    for k in (1, 2, 3):
        v = v + img[iy[:, k]] * wy[:, k, None, None]
    v = v.astype(np.float32)
    out = v[:, ix[:, 0]] * wx[None, :, 0, None]
    for k in (1, 2, 3):
        out = out + v[:, ix[:, k]] * wx[None, :, k, None]
    return np.ascontiguousarray(out, dtype=np.float32)


def decode_to_float(u8_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """dataset.py:31-37: cast -> resize (always runs, identity at equal size) -> IEEE ``/ 255.0``."""
    x = u8_hwc.astype(np.float32)
    x = resize_bicubic(x, out_h, out_w)
    return np.ascontiguousarray((x / F32(255.0)).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# libjpeg(-turbo) baseline 4:2:0 encode -> decode, integer restatement (entropy coding is lossless
# and skipped).  jcparam.c / jccolor.c / jcsample.c / jfdctint.c / jcdctmgr.c / jidctint.c /
# jdsample.c / jdcolor.c of libjpeg-turbo (the library TF links); not in /root/reference.
# ----------------------------------------------------------------------------------------------
_LUMA_BASE = np.array(
    [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
     14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
     18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)
_CHROMA_BASE = np.array(
    [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64).reshape(8, 8)


def quant_tables(quality: int):
    """jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE). Returns (luma, chroma) int32 [8,8]
    in natural (row = vertical frequency) order."""
    q = int(quality)
    q = 1 if q <= 0 else (100 if q > 100 else q)
    s = 5000 // q if q < 50 else 200 - 2 * q
    out = []
    for base in (_LUMA_BASE, _CHROMA_BASE):
        t = (base * s + 50) // 100
        out.append(np.clip(t, 1, 255).astype(np.int32))
    return out[0], out[1]


_C = dict(c0298=2446, c0390=3196, c0541=4433, c0765=6270, c0899=7373, c1175=9633,
          c1501=12299, c1847=15137, c1961=16069, c2053=16819, c2562=20995, c3072=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, axis, pass1: bool):
    """One pass of jpeg_fdct_islow along ``axis`` (length 8) on an int64 array."""
    d = np.moveaxis(d, axis, -1)
    x = [d[..., i] for i in range(8)]
    t0, t7 = x[0] + x[7], x[0] - x[7]
    t1, t6 = x[1] + x[6], x[1] - x[6]
    t2, t5 = x[2] + x[5], x[2] - x[5]
    t3, t4 = x[3] + x[4], x[3] - x[4]
    t10, t13 = t0 + t3, t0 - t3
    t11, t12 = t1 + t2, t1 - t2
    out = [None] * 8
    if pass1:
        out[0] = (t10 + t11) << 2
        out[4] = (t10 - t11) << 2
        n = 13 - 2
    else:
        out[0] = _descale(t10 + t11, 2)
        out[4] = _descale(t10 - t11, 2)
        n = 13 + 2
    z1 = (t12 + t13) * _C["c0541"]
    out[2] = _descale(z1 + t13 * _C["c0765"], n)
    out[6] = _descale(z1 - t12 * _C["c1847"], n)
    z1 = t4 + t7
    z2 = t5 + t6
    z3 = t4 + t6
    z4 = t5 + t7
    z5 = (z3 + z4) * _C["c1175"]
    t4 = t4 * _C["c0298"]
    t5 = t5 * _C["c2053"]
    t6 = t6 * _C["c3072"]
    t7 = t7 * _C["c1501"]
    z1 = -z1 * _C["c0899"]
    z2 = -z2 * _C["c2562"]
    z3 = -z3 * _C["c1961"] + z5
    z4 = -z4 * _C["c0390"] + z5
    out[7] = _descale(t4 + z1 + z3, n)
    out[5] = _descale(t5 + z2 + z4, n)
    out[3] = _descale(t6 + z2 + z3, n)
    out[1] = _descale(t7 + z1 + z4, n)
    return np.moveaxis(np.stack(out, axis=-1), -1, axis)


def _idct_1d(d, axis, pass1: bool):
    """One pass of jpeg_idct_islow along ``axis`` on an int64 array (input already dequantised)."""
    d = np.moveaxis(d, axis, -1)
    x = [d[..., i] for i in range(8)]
    z2, z3 = x[2], x[6]
    z1 = (z2 + z3) * _C["c0541"]
    t2 = z1 - z3 * _C["c1847"]
    t3 = z1 + z2 * _C["c0765"]
    t0 = (x[0] + x[4]) << 13
    t1 = (x[0] - x[4]) << 13
    t10, t13 = t0 + t3, t0 - t3
    t11, t12 = t1 + t2, t1 - t2
    t0, t1, t2, t3 = x[7], x[5], x[3], x[1]
    z1 = t0 + t3
    z2 = t1 + t2
    z3 = t0 + t2
    z4 = t1 + t3
    z5 = (z3 + z4) * _C["c1175"]
    t0 = t0 * _C["c0298"]
    t1 = t1 * _C["c2053"]
    t2 = t2 * _C["c3072"]
    t3 = t3 * _C["c1501"]
    z1 = -z1 * _C["c0899"]
    z2 = -z2 * _C["c2562"]
    z3 = -z3 * _C["c1961"] + z5
    z4 = -z4 * _C["c0390"] + z5
    t0 = t0 + z1 + z3
    t1 = t1 + z2 + z4
    t2 = t2 + z2 + z3
    t3 = t3 + z1 + z4
    n = (13 - 2) if pass1 else (13 + 2 + 3)
    out = [
        _descale(t10 + t3, n), _descale(t11 + t2, n), _descale(t12 + t1, n), _descale(t13 + t0, n),
        _descale(t13 - t0, n), _descale(t12 - t1, n), _descale(t11 - t2, n), _descale(t10 - t3, n),
    ]
    return np.moveaxis(np.stack(out, axis=-1), -1, axis)


def _plane_round_trip(plane_u8: np.ndarray, table: np.ndarray) -> np.ndarray:
    """u8 [Hp,Wp] (multiples of 8) -> u8 after level shift, FDCT, quantise, dequantise, IDCT, clamp."""
    hp, wp = plane_u8.shape
    b = plane_u8.astype(np.int64).reshape(hp // 8, 8, wp // 8, 8).transpose(0, 2, 1, 3) - 128  # [by,bx,y,x]
    c = _fdct_1d(b, 3, True)      # rows
    c = _fdct_1d(c, 2, False)     # columns -> c[by,bx,v,u], 8x the true DCT
    qt = (table.astype(np.int64) << 3)[None, None]
    mag = (np.abs(c) + (qt >> 1)) // qt           # jcdctmgr.c quantize(): round half away from zero
    coef = np.sign(c) * mag
    deq = coef * table.astype(np.int64)[None, None]
    w = _idct_1d(deq, 2, True)    # columns first (jidctint pass 1)
    p = _idct_1d(w, 3, False)     # then rows
    p = np.clip(p + 128, 0, 255)  # SIMD islow IDCT saturates (packs + 128); see module docstring
    return p.transpose(0, 2, 1, 3).reshape(hp, wp).astype(np.uint8)


def rgb_to_ycc(rgb_u8: np.ndarray):
    """jccolor.c rgb_ycc_convert, 16-bit fixed point. Returns (Y, Cb, Cr) int64 [H,W]."""
    r = rgb_u8[..., 0].astype(np.int64)
    g = rgb_u8[..., 1].astype(np.int64)
    b = rgb_u8[..., 2].astype(np.int64)
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16
    return y, cb, cr


def _pad_edge(a: np.ndarray, rows: int, cols: int) -> np.ndarray:
    return np.pad(a, ((0, rows - a.shape[0]), (0, cols - a.shape[1])), mode="edge")


def _h2v2_downsample(c_full: np.ndarray) -> np.ndarray:
    """jcsample.c h2v2_downsample on an even-sized plane; bias alternates 1,2 along output columns."""
    a = c_full[0::2, 0::2] + c_full[0::2, 1::2] + c_full[1::2, 0::2] + c_full[1::2, 1::2]
    bias = np.where(np.arange(a.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (a + bias) >> 2


def _h2v2_fancy_upsample(c: np.ndarray) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample on the REAL chroma plane [ceil(H/2), ceil(W/2)] with edge replication
    (jdmainct context rows). Returns [2*hc, 2*wc]."""
    c = c.astype(np.int64)
    hc, wc = c.shape
    if wc <= 2:    # jdsample.c jinit_upsampler: fancy upsampling needs downsampled_width > 2, else plain replication
        return np.repeat(np.repeat(c, 2, axis=0), 2, axis=1)
    up = np.concatenate([c[:1], c[:-1]], axis=0)     # row above (replicated at top)
    dn = np.concatenate([c[1:], c[-1:]], axis=0)     # row below (replicated at bottom)
    rows = np.empty((2 * hc, wc), dtype=np.int64)
    rows[0::2] = 3 * c + up
    rows[1::2] = 3 * c + dn
    left = np.concatenate([rows[:, :1], rows[:, :-1]], axis=1)
    right = np.concatenate([rows[:, 1:], rows[:, -1:]], axis=1)
    out = np.empty((2 * hc, 2 * wc), dtype=np.int64)
    out[:, 0::2] = (3 * rows + left + 8) >> 4
    out[:, 1::2] = (3 * rows + right + 7) >> 4
    return out


def jpeg_round_trip_u8(rgb_u8: np.ndarray, quality: int) -> np.ndarray:
    """u8 [H,W,3] -> u8 [H,W,3]: what ``decode_jpeg(encode_jpeg(x, quality=q, chroma_downsampling=True))`` gives
    with libjpeg-turbo defaults (baseline, 4:2:0, ISLOW DCT both ways, fancy upsampling)."""
    h, w, _ = rgb_u8.shape
    lum_t, chr_t = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(rgb_u8)
    # encoder-side padding (jcsample.c expand_right_edge, jcprepct.c expand_bottom_edge)
    h2 = h + (h & 1)
    wy_pad = -(-w // 8) * 8
    wc_full = -(-w // 16) * 16
    y_p = _pad_edge(_pad_edge(y, h2, wy_pad), -(-h // 16) * 16, wy_pad)
    planes = []
    for c in (cb, cr):
        c_full = _pad_edge(c, h2, wc_full)
        cd = _h2v2_downsample(c_full)
        planes.append(_pad_edge(cd, -(-cd.shape[0] // 8) * 8, cd.shape[1]))
    y_r = _plane_round_trip(y_p.astype(np.uint8), lum_t)[:h, :w].astype(np.int64)
    hc, wc = -(-h // 2), -(-w // 2)
    cb_r = _h2v2_fancy_upsample(_plane_round_trip(planes[0].astype(np.uint8), chr_t)[:hc, :wc])[:h, :w] - 128
    cr_r = _h2v2_fancy_upsample(_plane_round_trip(planes[1].astype(np.uint8), chr_t)[:hc, :wc])[:h, :w] - 128
    # jdcolor.c ycc_rgb_convert
    r = y_r + ((91881 * cr_r + 32768) >> 16)
    g = y_r + ((-22554 * cb_r - 46802 * cr_r + 32768) >> 16)
    b = y_r + ((116130 * cb_r + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def float_to_u8_saturate(x: np.ndarray) -> np.ndarray:
    """tf.image.convert_image_dtype(float32 -> uint8, saturate=True): trunc(clip(x * 255.5, 0, 255))."""
    s = (x.astype(np.float32) * F32(255.5)).astype(np.float32)
    return np.clip(s, F32(0.0), F32(255.0)).astype(np.uint8)   # astype truncates


def adjust_jpeg_quality(x_f32: np.ndarray, quality: int, backend: str = "integer") -> np.ndarray:
    """tf.image.adjust_jpeg_quality on a float image in [0,1] (augment.py:112)."""
    u8 = float_to_u8_saturate(x_f32)
    if backend == "integer":
        rt = jpeg_round_trip_u8(u8, quality)
    elif backend == "pillow":
        rt = jpeg_round_trip_pillow(u8, quality)
    else:
        raise ValueError(backend)
    return (rt.astype(np.float32) * F32(1.0 / 255.0)).astype(np.float32)


def jpeg_round_trip_pillow(rgb_u8: np.ndarray, quality: int) -> np.ndarray:
    """The real libjpeg-turbo via Pillow; pins :func:`jpeg_round_trip_u8` (used by tests and as a faster
    CPU-baseline leg)."""
    import io

    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(rgb_u8, "RGB").save(buf, format="JPEG", quality=int(quality), subsampling=2, optimize=False)
    buf.seek(0)
    return np.asarray(Image.open(buf).convert("RGB"))


GRAY_W = (F32(0.2989), F32(0.5870), F32(0.1140))

FLAG_HFLIP, FLAG_VFLIP, FLAG_GRAY = 1, 2, 4


def apply_flags(x: np.ndarray, flags: int) -> np.ndarray:
    """augment.py:115-120 (flips) then 142-146 (gray). Gray = ((R*0.2989 + G*0.5870) + B*0.1140), fp32, each
    product/sum rounded separately (TF's tensordot order is unpinned; this is the oracle's definition)."""
    if flags & FLAG_HFLIP:
        x = x[:, ::-1]
    if flags & FLAG_VFLIP:
        x = x[::-1]
    if flags & FLAG_GRAY:
        g = (x[..., 0] * GRAY_W[0] + x[..., 1] * GRAY_W[1]).astype(np.float32) + x[..., 2] * GRAY_W[2]
        x = np.repeat(g.astype(np.float32)[..., None], 3, axis=-1)
    return np.ascontiguousarray(x, dtype=np.float32)


def preprocess_one(src_u8, out_h, out_w, crop_yxhw=None, jpeg_q=-1, flags=0, jpeg_backend="integer"):
    """One image of the preprocessing path: crop -> bicubic -> /255 -> JPEG(q) (q<0: skip) -> flips -> gray."""
    if crop_yxhw is not None:
        y0, x0, h, w = (int(v) for v in crop_yxhw)
        src_u8 = src_u8[y0:y0 + h, x0:x0 + w]
    x = decode_to_float(src_u8, out_h, out_w)
    if jpeg_q is not None and jpeg_q >= 0:
        x = adjust_jpeg_quality(x, int(jpeg_q), backend=jpeg_backend)
    return apply_flags(x, int(flags))


def preprocess_batch(src_u8, out_h, out_w, crops=None, jpeg_q=None, flags=None, jpeg_backend="integer"):
    n = src_u8.shape[0]
    out = np.empty((n, out_h, out_w, 3), dtype=np.float32)
    for i in range(n):
        out[i] = preprocess_one(
            src_u8[i], out_h, out_w,
            None if crops is None else crops[i],
            -1 if jpeg_q is None else int(jpeg_q[i]),
            0 if flags is None else int(flags[i]),
            jpeg_backend,
        )
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# ----------------------------------------------------------------------------------------------
def synth_image(i: int, h: int = 200, w: int = 200) -> np.ndarray:
    """Deterministic natural-ish u8 image: blurred noise + white noise + gradient."""
    rng = np.random.default_rng(20221000 + i)
    base = rng.random((h // 8 + 2, w // 8 + 2, 3))
    yy = np.linspace(0, base.shape[0] - 1.001, h)
    xx = np.linspace(0, base.shape[1] - 1.001, w)
    y0, x0 = yy.astype(int), xx.astype(int)
    fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
    img = (base[y0][:, x0] * (1 - fy) * (1 - fx) + base[y0 + 1][:, x0] * fy * (1 - fx)
           + base[y0][:, x0 + 1] * (1 - fy) * fx + base[y0 + 1][:, x0 + 1] * fy * fx)
    noise = rng.random((h, w, 3)) * rng.uniform(0.05, 0.2)
    grad = np.linspace(0, rng.uniform(0.0, 0.3), w)[None, :, None]
    img = img * 0.8 + noise + grad
    img = (img - img.min()) / (img.max() - img.min() + 1e-9)
    return np.ascontiguousarray(np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8))


def synth_decisions(n: int, seed: int = 42, src_h: int = 200, src_w: int = 200):
    """Per-image augmentation decisions (CFG.seed=42, main.py:224): square crop side in [160,200],
    uniform offsets, q in [65,100), hflip/vflip ~ Bernoulli(0.5)."""
    rng = np.random.default_rng(seed)
    side = rng.integers(160, min(src_h, src_w) + 1, size=n)
    y0 = (rng.random(n) * (src_h - side + 1)).astype(np.int64)
    x0 = (rng.random(n) * (src_w - side + 1)).astype(np.int64)
    crops = np.stack([y0, x0, side, side], axis=1).astype(np.int32)
    q = rng.integers(65, 100, size=n).astype(np.int32)
    flags = (rng.integers(0, 2, size=n) * FLAG_HFLIP + rng.integers(0, 2, size=n) * FLAG_VFLIP).astype(np.uint8)
    return crops, q, flags
