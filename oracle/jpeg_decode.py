"""ORACLE (test infrastructure, never imported by the product path): CPU restatement of baseline JPEG decoding as
``tf.io.read_file`` -> ``tf.image.decode_jpeg(channels=3)`` performs it (reference: dataset/dataset.py:24-28).

The arithmetic lives in an un-vendored third-party library: the libjpeg-turbo bundled with TensorFlow (version unpinned by
the reference; defaults JDCT_ISLOW + fancy upsampling).  Restated from the published algorithms:
  stream syntax, Huffman decoding (DECODE / RECEIVE / EXTEND), restart intervals      ITU-T T.81 Annex B, F.2.2, E.2.4
  dequantise + integer inverse DCT                                                     jidctint.c  (oracle/preprocess._idct_1d)
  h2v1 / h2v2 fancy upsampling, YCbCr -> RGB                                           jdsample.c, jdcolor.c
PINNED: the same library family is present in this container behind Pillow (libjpeg-turbo), so every function here is
checked bit for bit against ``PIL.Image.open(...).convert("RGB")`` in tests/test_oracle_jpeg_decode.py on 4:2:0 / 4:2:2 /
4:4:4 / grey files, odd sizes, optimised Huffman tables and restart intervals, and against the committed golden files
(tests/golden/jpeg_files.npz, generator tests/golden/make_golden.py).  Pure-Python bit loops: small images only."""
from __future__ import annotations

import numpy as np

from . import preprocess as P

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                   60, 61, 54, 47, 55, 62, 63])


class Unsupported(ValueError):
    pass


def parse(data: bytes) -> dict:
    """Marker walk (T.81 B.1, B.2): tables, frame and scan headers, the entropy-coded segment of a single-scan file."""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG")
    p, qt, ht, frame, ri = 2, {}, {}, None, 0
    while p < len(data):
        assert data[p] == 0xFF
        while data[p] == 0xFF:
            p += 1
        m = data[p]
        p += 1
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        L = int.from_bytes(data[p:p + 2], "big")
        s = data[p + 2:p + L]
        if m == 0xDB:
            i = 0
            while i < len(s):
                pq, tq = s[i] >> 4, s[i] & 15
                n = 128 if pq else 64
                v = np.frombuffer(s[i + 1:i + 1 + n], ">u2" if pq else np.uint8).astype(np.int64)
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = v
                qt[tq] = t.reshape(8, 8)
                i += 1 + n
        elif m == 0xC4:
            i = 0
            while i < len(s):
                tc, th = s[i] >> 4, s[i] & 15
                bits = list(s[i + 1:i + 17])
                n = sum(bits)
                ht[(tc, th)] = (bits, list(s[i + 17:i + 17 + n]))
                i += 17 + n
        elif m in (0xC0, 0xC1):
            if s[0] != 8:
                raise Unsupported("precision")
            frame = dict(h=int.from_bytes(s[1:3], "big"), w=int.from_bytes(s[3:5], "big"),
                         comps=[dict(id=s[6 + 3 * c], hs=s[7 + 3 * c] >> 4, vs=s[7 + 3 * c] & 15, tq=s[8 + 3 * c])
                                for c in range(s[5])])
        elif 0xC2 <= m <= 0xCF and m != 0xC8:
            raise Unsupported("not a sequential Huffman frame")
        elif m == 0xDD:
            ri = int.from_bytes(s[0:2], "big")
        elif m == 0xDA:
            ns = s[0]
            if frame is None or ns != len(frame["comps"]):
                raise Unsupported("multi-scan")
            for c in range(ns):
                frame["comps"][c]["td"], frame["comps"][c]["ta"] = s[2 + 2 * c] >> 4, s[2 + 2 * c] & 15
            start = p + L
            q = start
            while q + 1 < len(data) and not (data[q] == 0xFF and data[q + 1] != 0 and not 0xD0 <= data[q + 1] <= 0xD7):
                q += 1
            return dict(frame=frame, qt=qt, ht=ht, ri=ri, scan=data[start:q])
        p += L
    raise Unsupported("no scan")


class _Bits:
    """Bit reader over the entropy-coded segment with byte stuffing removed (T.81 F.2.2.5 NEXTBIT)."""

    def __init__(self, seg: bytes):
        self.seg, self.pos, self.acc, self.n = seg, 0, 0, 0

    def _byte(self):
        if self.pos >= len(self.seg):
            return 0
        b = self.seg[self.pos]
        if b == 0xFF:
            nxt = self.seg[self.pos + 1] if self.pos + 1 < len(self.seg) else 0xD9
            if nxt == 0:
                self.pos += 2
                return 0xFF
            return 0                       # marker: feed zeros, stay put
        self.pos += 1
        return b

    def get(self, k):
        while self.n < k:
            self.acc = (self.acc << 8) | self._byte()
            self.n += 8
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        self.acc = self.n = 0
        assert self.seg[self.pos] == 0xFF and 0xD0 <= self.seg[self.pos + 1] <= 0xD7, "restart marker expected"
        self.pos += 2


def _huff_table(bits, vals):
    """T.81 Annex C: code -> (length, symbol) as a dict keyed by (length, code)."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[(length, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.get(1)
        if (length, code) in table:
            return table[(length, code)]
    raise ValueError("invalid Huffman code")


def _extend(v, s):
    return v - (1 << s) + 1 if v < (1 << (s - 1)) else v


def decode_coefficients(info):
    """Entropy decode: per component an int64 array [blocks_y, blocks_x, 8, 8] (natural order, not yet dequantised)."""
    fr = info["frame"]
    comps = fr["comps"]
    if len(comps) == 1:
        comps[0]["hs"] = comps[0]["vs"] = 1
    hmax, vmax = max(c["hs"] for c in comps), max(c["vs"] for c in comps)
    mcux, mcuy = -(-fr["w"] // (8 * hmax)), -(-fr["h"] // (8 * vmax))
    coefs = [np.zeros((mcuy * c["vs"], mcux * c["hs"], 64), np.int64) for c in comps]
    tabs = {k: _huff_table(*v) for k, v in info["ht"].items()}
    br, pred, count = _Bits(info["scan"]), [0] * len(comps), 0
    for my in range(mcuy):
        for mx in range(mcux):
            if info["ri"] and count and count % info["ri"] == 0:
                br.restart()
                pred = [0] * len(comps)
            count += 1
            for ci, c in enumerate(comps):
                for v in range(c["vs"]):
                    for h in range(c["hs"]):
                        blk = coefs[ci][my * c["vs"] + v, mx * c["hs"] + h]
                        s = _decode_symbol(br, tabs[(0, c["td"])])
                        pred[ci] += _extend(br.get(s), s) if s else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(br, tabs[(1, c["ta"])])
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[ZIGZAG[k]] = _extend(br.get(s), s)
                                k += 1
                            elif r == 15:
                                k += 16
                            else:
                                break
    return [c.reshape(c.shape[0], c.shape[1], 8, 8) for c in coefs], (hmax, vmax)


def _idct_plane(coef, qt):
    """[by,bx,v,u] coefficients -> u8 plane: dequantise, jidctint columns then rows, +128, clamp."""
    deq = coef * qt[None, None]
    w = P._idct_1d(deq, 2, True)
    p = np.clip(P._idct_1d(w, 3, False) + 128, 0, 255)
    by, bx = coef.shape[:2]
    return p.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def _h2v1_fancy_upsample(c):
    """jdsample.c h2v1_fancy_upsample on the real chroma plane [H, ceil(W/2)] -> [H, 2*wc]."""
    c = c.astype(np.int64)
    if c.shape[1] <= 2:                  # jinit_upsampler: h2v1_upsample (replication) unless downsampled_width > 2
        return np.repeat(c, 2, axis=1)
    left = np.concatenate([c[:, :1], c[:, :-1]], axis=1)
    right = np.concatenate([c[:, 1:], c[:, -1:]], axis=1)
    out = np.empty((c.shape[0], 2 * c.shape[1]), np.int64)
    out[:, 0::2] = (3 * c + left + 1) >> 2
    out[:, 1::2] = (3 * c + right + 2) >> 2
    return out


def decode(data: bytes) -> np.ndarray:
    """Whole file -> uint8 [H,W,3] RGB (grey files replicated to three channels, decode_jpeg(channels=3))."""
    info = parse(data)
    fr = info["frame"]
    h, w, comps = fr["h"], fr["w"], fr["comps"]
    coefs, (hmax, vmax) = decode_coefficients(info)
    planes = [_idct_plane(cf, info["qt"][c["tq"]]) for cf, c in zip(coefs, comps)]
    y = planes[0][:h, :w]
    if len(comps) == 1:
        return np.repeat(y[..., None], 3, axis=2).astype(np.uint8)
    if any((c["hs"], c["vs"]) != (1, 1) for c in comps[1:]) or (hmax, vmax) not in ((1, 1), (2, 1), (2, 2)):
        raise Unsupported("sampling factors")
    hc, wc = -(-h * 1 // vmax), -(-w // hmax)
    ch = []
    for pl in planes[1:]:
        real = pl[:hc, :wc]
        up = real if hmax == 1 else (_h2v1_fancy_upsample(real) if vmax == 1 else P._h2v2_fancy_upsample(real))
        ch.append(up[:h, :w] - 128)
    cb, cr = ch
    r = y + ((91881 * cr + 32768) >> 16)
    g = y + ((-22554 * cb - 46802 * cr + 32768) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)
