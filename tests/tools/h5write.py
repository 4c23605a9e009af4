"""Minimal HDF5 WRITER for test fixtures: produces the layout h5py's default settings (libver 'earliest') give a Keras
weights file -- version-0 superblock, version-1 object headers, old-style groups (symbol-table message, version-1 B-tree of
SNOD nodes, local heap), contiguous little-endian datasets, version-1 attribute messages with fixed-length strings --
written independently of vipcup_b200/h5lite.py from the same HDF5 File Format Specification.  ``leaf_cap`` / ``node_cap``
set how many symbols a SNOD and how many children a B-tree node hold (small values force multi-node, two-level trees)."""
import struct

import numpy as np

UNDEF = b"\xff" * 8


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class Writer:
    def __init__(self, leaf_cap=8, node_cap=4):
        self.b = bytearray(96)            # superblock written last
        self.leaf_cap, self.node_cap = leaf_cap, node_cap

    def alloc(self, data):
        self.b += b"\0" * (-len(self.b) % 8)
        addr = len(self.b)
        self.b += data
        return addr

    # ---- messages
    @staticmethod
    def _msg(mtype, data):
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), 0) + data

    @staticmethod
    def _dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)

    @staticmethod
    def _datatype(dt):
        dt = np.dtype(dt)
        if dt.kind == "f":
            e, m = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
            bits = 0x20 | ((dt.itemsize * 8 - 1) << 8) | (1 if dt.byteorder == ">" else 0)
            return (struct.pack("<B3sI", 0x11, bits.to_bytes(3, "little"), dt.itemsize) +
                    struct.pack("<HHBBBBI", 0, dt.itemsize * 8, m, e, 0, m, (1 << (e - 1)) - 1))
        if dt.kind in "iu":
            bits = (8 if dt.kind == "i" else 0) | (1 if dt.byteorder == ">" else 0)
            return struct.pack("<B3sI", 0x10, bits.to_bytes(3, "little"), dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
        if dt.kind == "S":
            return struct.pack("<B3sI", 0x13, (1).to_bytes(3, "little"), dt.itemsize)      # null-padded ASCII
        raise TypeError(dt)

    def _attr(self, name, value):
        arr = np.asarray(value)
        nm = name.encode() + b"\0"
        dtb, dsb = self._datatype(arr.dtype), self._dataspace(arr.shape)
        body = struct.pack("<BxHHH", 1, len(nm), len(dtb), len(dsb)) + _pad8(nm) + _pad8(dtb) + _pad8(dsb) + arr.tobytes()
        assert len(body) < 65000, "attribute too large for one header message"
        return self._msg(0x0C, body)

    def _header(self, msgs):
        body = b"".join(msgs)
        return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body

    # ---- objects
    def dataset(self, arr, attrs=None):
        arr = np.ascontiguousarray(arr)
        data = self.alloc(arr.tobytes()) if arr.size else 0
        layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, arr.nbytes)
        msgs = [self._msg(0x01, self._dataspace(arr.shape)), self._msg(0x03, self._datatype(arr.dtype)), self._msg(0x08, layout)]
        msgs += [self._attr(k, v) for k, v in (attrs or {}).items()]
        return self.alloc(self._header(msgs))

    def group(self, children, attrs=None):
        """children: {name: object header address}"""
        names = sorted(children)
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        heap_seg = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQ", 0, len(heap_data)) + UNDEF + struct.pack("<Q", heap_seg))
        # symbol nodes
        leaves = []
        for i in range(0, max(len(names), 1), self.leaf_cap):
            part = names[i: i + self.leaf_cap]
            ent = b"".join(struct.pack("<QQII16x", offs[n], children[n], 0, 0) for n in part)
            leaves.append((self.alloc(b"SNOD" + struct.pack("<BxH", 1, len(part)) + ent), offs[part[-1]] if part else 0))
        level, nodes = 0, leaves
        while True:
            nxt = []
            for i in range(0, len(nodes), self.node_cap):
                part = nodes[i: i + self.node_cap]
                body = struct.pack("<Q", 0)
                for addr, last in part:
                    body += struct.pack("<QQ", addr, last)
                nxt.append((self.alloc(b"TREE" + struct.pack("<BBH", 0, level, len(part)) + UNDEF + UNDEF + body), part[-1][1]))
            nodes, level = nxt, level + 1
            if len(nodes) == 1:
                break
        msgs = [self._msg(0x11, struct.pack("<QQ", nodes[0][0], heap))]
        msgs += [self._attr(k, v) for k, v in (attrs or {}).items()]
        return self.alloc(self._header(msgs)), nodes[0][0], heap

    def finish(self, root):
        addr, btree, heap = root
        sb = (b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0) +
              struct.pack("<Q", 0) + UNDEF + struct.pack("<Q", len(self.b)) + UNDEF +
              struct.pack("<QQII", 0, addr, 1, 0) + struct.pack("<QQ", btree, heap))
        assert len(sb) == 96
        self.b[:96] = sb
        return bytes(self.b)


def _tree(w, node, attrs):
    """node: nested dict name -> (dict | ndarray); returns the header address (datasets) or the group triple."""
    kids = {}
    for name, v in node.items():
        kids[name] = _tree(w, v, {})[0] if isinstance(v, dict) else w.dataset(v)
    return w.group(kids, attrs)


def write_keras_weights(path, layers, wrap_model_weights=False, leaf_cap=8, node_cap=4, attr_chunk=60000):
    """layers: [(layer name, [(weight name such as 'conv/kernel:0', ndarray), ...])] in Keras order."""
    w = Writer(leaf_cap, node_cap)

    def name_list(prefix, names):
        arr = np.array([n.encode() for n in names], dtype="S") if names else np.zeros((0,), "S1")
        if arr.nbytes <= attr_chunk:
            return {prefix: arr}
        per = max(1, attr_chunk // arr.dtype.itemsize)              # hdf5_format.save_attributes_to_hdf5_group splits
        return {f"{prefix}{i}": arr[k: k + per] for i, k in enumerate(range(0, len(arr), per))}

    kids = {}
    for lname, weights in layers:
        tree = {}
        for wname, arr in weights:
            parts = wname.split("/")
            d = tree
            for p in parts[:-1]:
                d = d.setdefault(p, {})
            d[parts[-1]] = np.asarray(arr)
        kids[lname] = _tree(w, tree, name_list("weight_names", [n for n, _ in weights]))[0]
    attrs = dict(name_list("layer_names", [n for n, _ in layers]), backend=np.array(b"tensorflow"), keras_version=np.array(b"2.9.0"))
    root = w.group(kids, attrs)
    if wrap_model_weights:
        root = w.group({"model_weights": root[0]}, {"keras_version": np.array(b"2.9.0")})
    data = w.finish(root)
    with open(path, "wb") as f:
        f.write(data)
    return path
