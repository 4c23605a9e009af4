"""Keras .h5 checkpoint ingestion (vipcup_b200/h5lite.py; reference: main.py:107,186-194).  No HDF5 library
exists offline, so the reader is checked against (1) files from the independent writer tests/tools/h5write.py in several
tree shapes, (2) structures assembled byte by byte here from the format specification that the writer never produces
(version-2 superblock + OHDR headers with compact links, chunked + shuffled + deflated data, big-endian floats,
variable-length strings through a global heap), (3) the registry path."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "tools"))


def _layers(rng, n=40):
    layers = []
    for i in range(n):
        ws = [(f"block{i}/conv/kernel:0", rng.standard_normal((3, 3, 4, 8)).astype(np.float32)),
              (f"block{i}/conv/bias:0", rng.standard_normal((8,)).astype(np.float32))]
        if i % 5 == 0:
            ws.append((f"block{i}/bn/moving_mean:0", rng.standard_normal((8,))))           # float64
        if i % 6 == 0:
            ws.append((f"block{i}/steps:0", np.arange(3, dtype=np.int64)))
        layers.append((f"block{i}", [] if i % 7 == 3 else ws))
    return layers


@pytest.mark.parametrize("kw", [dict(), dict(wrap_model_weights=True), dict(leaf_cap=3, node_cap=2), dict(leaf_cap=1, node_cap=2),
                                dict(attr_chunk=150)])
def test_round_trip_with_the_independent_writer(tmp_path, kw):
    import h5write

    from vipcup_b200 import h5lite

    layers = _layers(np.random.default_rng(1))
    path = h5write.write_keras_weights(str(tmp_path / "w.h5"), layers, **kw)
    W = h5lite.load_keras_weights(path)
    want = {k[:-2]: a for _, ws in layers for k, a in ws}
    assert list(W) == list(want)                                   # Keras order kept
    for k, a in want.items():
        assert W[k].dtype == a.dtype and np.array_equal(W[k], a), k
    f = h5lite.File(path)
    g = f["model_weights"] if kw.get("wrap_model_weights") else f
    assert g.attrs["backend"] == b"tensorflow" and "block0" in g and g["block0/block0/conv/kernel:0"].shape == (3, 3, 4, 8)
    with pytest.raises(KeyError):
        g["block0/nope"]


def _ohdr(msgs):
    """version-2 object header, 2-byte chunk size, no times / phase change fields; checksum left zero (not verified)"""
    body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
    return b"OHDR" + struct.pack("<BBH", 2, 0x01, len(body)) + body + b"\0\0\0\0"


def test_hand_assembled_new_style_file(tmp_path):
    """Superblock v2, OHDR headers, compact link messages, attribute v3 (variable-length strings in a global heap), a
    chunked 2-D big-endian float dataset with shuffle + deflate, a compact dataset."""
    from vipcup_b200 import h5lite

    buf = bytearray(48)                                            # superblock v2 patched in at the end

    def alloc(b):
        buf.extend(b"\0" * (-len(buf) % 8))
        a = len(buf)
        buf.extend(b)
        return a

    f32be = struct.pack("<B3sI", 0x11, (0x21 | (31 << 8)).to_bytes(3, "little"), 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    # --- chunked dataset [5, 6] in chunks of [4, 4], shuffle (id 2) then deflate (id 1)
    data = np.arange(30, dtype=">f4").reshape(5, 6) * 0.5
    entries = []
    for r0 in (0, 4):
        for c0 in (0, 4):
            chunk = np.zeros((4, 4), ">f4")
            blk = data[r0: r0 + 4, c0: c0 + 4]
            chunk[: blk.shape[0], : blk.shape[1]] = blk
            raw = np.frombuffer(chunk.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()       # shuffle
            comp = zlib.compress(raw)
            entries.append((len(comp), (r0, c0), alloc(comp)))
    node = b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + b"\xff" * 16
    for size, (r0, c0), addr in entries:
        node += struct.pack("<IIQQQ", size, 0, r0, c0, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 8, 8, 0)                                            # final key
    bt = alloc(node)
    filt = struct.pack("<BB", 2, 2) + struct.pack("<HHHI", 2, 0, 1, 4) + struct.pack("<HHHI", 1, 0, 1, 6)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 4, 4, 4)
    dspace = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 5, 6)
    d_chunked = alloc(_ohdr([(0x01, dspace), (0x03, f32be), (0x0B, filt), (0x08, layout)]))
    # --- compact dataset of 3 int16
    i16 = struct.pack("<B3sI", 0x10, (8).to_bytes(3, "little"), 2) + struct.pack("<HH", 0, 16)
    small = np.array([-2, 7, 300], "<i2")
    d_compact = alloc(_ohdr([(0x01, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 3)), (0x03, i16),
                             (0x08, struct.pack("<BBH", 3, 0, small.nbytes) + small.tobytes())]))
    # --- global heap with two strings, attribute of 2 variable-length strings
    strs = [b"alpha/kernel:0", b"beta:0"]
    objs = b""
    for i, s in enumerate(strs):
        objs += struct.pack("<HHIQ", i + 1, 1, 0, len(s)) + s + b"\0" * (-len(s) % 8)
    objs += struct.pack("<HHIQ", 0, 0, 0, 0)
    gcol = alloc(b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(objs)) + objs)
    vlen_dt = struct.pack("<B3sI", 0x19, (1 | (1 << 8)).to_bytes(3, "little"), 16) + struct.pack("<B3sI", 0x30, (0).to_bytes(3, "little"), 1)
    vals = b"".join(struct.pack("<IQI", len(s), gcol, i + 1) for i, s in enumerate(strs))
    name = b"weight_names\0"
    ds1 = struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2)
    attr = struct.pack("<BBHHHB", 3, 0, len(name), len(vlen_dt), len(ds1), 0) + name + vlen_dt + ds1 + vals

    def link(nm, addr):
        return struct.pack("<BBB", 1, 0, len(nm)) + nm + struct.pack("<Q", addr)

    linfo = struct.pack("<BB", 0, 0) + b"\xff" * 16
    grp = alloc(_ohdr([(0x02, linfo), (0x0A, b"\0\0"), (0x06, link(b"w", d_chunked)), (0x06, link(b"small", d_compact)), (0x0C, attr)]))
    root = alloc(_ohdr([(0x02, linfo), (0x06, link(b"layer", grp))]))
    buf[:48] = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(buf), root) + b"\0\0\0\0"
    p = tmp_path / "new.h5"
    p.write_bytes(bytes(buf))
    f = h5lite.File(str(p))
    assert f.keys() == ["layer"] and sorted(f["layer"].keys()) == ["small", "w"]
    got = f["layer/w"].read()
    assert got.dtype == np.dtype(">f4") and np.array_equal(got, data)
    assert np.array_equal(f["layer/small"].read(), small)
    assert [bytes(x) for x in f["layer"].attrs["weight_names"]] == strs


def test_errors_are_loud(tmp_path):
    from vipcup_b200 import h5lite

    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(p))


def test_registry_reads_h5_and_resolves_keras_name_scopes(tmp_path):
    """ckpts/<Arch>-<H>x<W>/ckpt/*.h5 as main.py:186-192 expects them; weight names carry the name scope of the model object
    (here 'resnet-rs-50/'), which resolve_weight_names strips by unique-suffix matching."""
    import h5write

    from vipcup_b200 import registry
    from vipcup_b200.models import ResNetRS

    m = ResNetRS(50, classes=2, device="cpu")
    rng = np.random.default_rng(0)
    W = {k: rng.standard_normal(s).astype(np.float32) for k, s in m.weight_shapes().items()}
    by_layer = {}
    for k, a in W.items():
        by_layer.setdefault(k.split("/")[0], []).append((f"resnet-rs-50/{k}:0", a))
    d = tmp_path / "ResNetRS50-200x200" / "ckpt"
    d.mkdir(parents=True)
    h5write.write_keras_weights(str(d / "fold0.h5"), list(by_layer.items()))
    (tmp_path / "ckpts.json").write_text('[["ResNetRS50-200x200", [200, 200], 0]]')
    entries = registry.scan_checkpoints(str(tmp_path), str(tmp_path / "ckpts.json"))
    assert entries[0][0][0].endswith("fold0.h5")
    got, meta = registry.load_checkpoint(entries[0][0][0])
    assert len(got) == len(W) and meta["head_act"] is None
    res = registry.resolve_weight_names(got, list(W))
    for k, a in W.items():
        assert np.array_equal(res[k], a), k
    with pytest.raises(KeyError):
        registry.resolve_weight_names(got, ["missing/kernel"])
