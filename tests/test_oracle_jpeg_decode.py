"""The oracle's JPEG decoder (oracle/jpeg_decode.py) pinned against the real libjpeg-turbo: the committed golden files
(tests/golden/jpeg_files.npz, decoded by Pillow when the fixture was made) and Pillow live in this process."""
import hashlib
import io
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_files():
    z = np.load(os.path.join(GOLD, "jpeg_files.npz"))
    for i in range(int(z["n"])):
        yield i, z[f"file{i}"].tobytes(), (z[f"out{i}"] if f"out{i}" in z.files else None), z[f"sha{i}"].tobytes()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()


def test_oracle_decoder_matches_libjpeg_turbo_golden_files():
    from oracle import jpeg_decode as J

    checked = skipped = 0
    for i, f, out, digest in golden_files():
        try:
            info = J.parse(f)
        except J.Unsupported:
            skipped += 1                      # the progressive file: host path of the product as well
            continue
        if info["frame"]["h"] * info["frame"]["w"] > 210 * 210:
            continue                          # pure-Python bit loops: small cases only
        got = J.decode(f)
        assert sha(got) == digest, f"file {i}"
        if out is not None:
            assert np.array_equal(got, out), f"file {i}"
        checked += 1
    assert checked >= 40 and skipped == 1


def test_golden_digests_match_the_pillow_in_this_process():
    from PIL import Image

    for i, f, out, digest in golden_files():
        live = np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))
        assert sha(live) == digest, f"file {i}: libjpeg-turbo here decodes differently from the one that made the fixture"


@pytest.mark.parametrize("kw", [dict(quality=75), dict(quality=92, subsampling=0), dict(quality=60, subsampling=1),
                                dict(quality=88, optimize=True), dict(quality=85, restart_marker_blocks=2)])
def test_oracle_decoder_matches_pillow_live(kw):
    from PIL import Image

    from oracle import jpeg_decode as J
    from oracle import preprocess as P

    for k, (h, w) in enumerate([(40, 56), (31, 17), (9, 70), (30, 3), (4, 4), (2, 9)]):   # narrow planes: replicated chroma
        b = io.BytesIO()
        Image.fromarray(P.synth_image(700 + k, max(h, 16), max(w, 16))[:h, :w]).save(b, "JPEG", **kw)
        ref = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))
        assert np.array_equal(J.decode(b.getvalue()), ref), (kw, h, w)


def test_header_parser_of_the_library_agrees_with_the_oracle():
    """vip_jpeg_parse is host code: geometry, tables and the entropy-coded segment it reports equal the oracle's parse."""
    from oracle import jpeg_decode as J
    from vipcup_b200 import jpeg

    for i, f, out, digest in golden_files():
        d = jpeg.parse(f)
        try:
            info = J.parse(f)
        except J.Unsupported:
            assert d.status == jpeg.VIP_JPEG_UNSUPPORTED and (d.width, d.height) == (96, 72)
            continue
        fr = info["frame"]
        assert d.status == jpeg.VIP_JPEG_OK and (d.height, d.width, d.ncomp) == (fr["h"], fr["w"], len(fr["comps"])), i
        assert f[d.scan_offset: d.scan_offset + d.scan_bytes] == info["scan"], i
        assert d.restart_interval == info["ri"]
        for c, comp in enumerate(fr["comps"]):
            if d.ncomp > 1:
                assert (d.hs[c], d.vs[c]) == (comp["hs"], comp["vs"])
            assert np.array_equal(np.array(d.qt[d.tq[c]]).reshape(8, 8), info["qt"][comp["tq"]])
            for cls, sel in ((0, d.td[c]), (1, d.ta[c])):
                bits, vals = info["ht"][(cls, sel)]
                assert list(d.huff_bits[2 * cls + sel]) == bits and list(d.huff_vals[2 * cls + sel])[: len(vals)] == vals
    for junk in (b"", b"\xff", b"\x89PNG\r\n\x1a\n" + b"\0" * 40, b"\xff\xd8\xff\xc0\x00"):
        assert jpeg.parse(junk).status == jpeg.VIP_JPEG_NOT_JPEG
