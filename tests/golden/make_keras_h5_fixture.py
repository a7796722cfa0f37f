"""Hand-assembles `keras_fcn_skip_tiny.h5` from the HDF5 File Format Specification (version 3.0, sections II-IV),
independently of `page_segmentation_b200/lib/h5.py` (neither its reader nor its writer is imported): the file is laid
out the way libhdf5 1.10/1.12 under h5py 3.x lays out what Keras' `model.save('x.h5')` writes for the reference's
`fcn_skip` graph (network.py:75-84 loads such a file), with the traits h5.py's own writer never produces:

  * superblock v0 with the library's default group K values (leaf 4, internal 16), root symbol-table entry with the
    cached B-tree / heap addresses;
  * version-1 object headers holding NIL, modification-time (0x12), old and new fill-value (0x04 / 0x05) messages and
    CONTINUATION messages (0x10) with the remaining attributes in a second block;
  * old-style groups whose 25 links overflow one symbol-table node (2K = 8 entries): a B-tree node with several SNOD
    children, names in a local heap with a free-list block;
  * `model_config` / `backend` / `keras_version` as VARIABLE-LENGTH UTF-8 strings (global heap collection "GCOL",
    attribute message version 3 with a UTF-8 name encoding), what TF >= 2.5 writes for `str` values;
  * `layer_names` split into `layer_names0` / `layer_names1` (Keras' save_attributes_to_hdf5_group for > 64 KB
    attributes) as fixed-length NULL-PADDED string arrays (numpy `S` dtype), `weight_names` likewise;
  * nested `layer/layer/kernel:0` groups; contiguous (layout v3) datasets, one CHUNKED dataset (v1 chunk B-tree, edge
    chunks) and one COMPACT dataset; a dataspace with the max-dims flag.

The channel counts are a tenth of fcn_skip's so that the fixture stays small.  Expected arrays: `expected_weights()`.
Run `python tests/golden/make_keras_h5_fixture.py` to regenerate the committed bytes.  If a file written by real
Keras ever becomes available it replaces this one.
"""
import json
import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "keras_fcn_skip_tiny.h5")

# Keras' auto-numbered names for model.py:45-92 built first in a fresh process
LAYERS = ["input_1", "lambda", "lambda_1", "conv2d", "conv2d_1", "max_pooling2d", "conv2d_2", "conv2d_3",
          "max_pooling2d_1", "conv2d_4", "conv2d_5", "max_pooling2d_2", "conv2d_6", "conv2d_transpose",
          "conv2d_transpose_1", "concatenate", "conv2d_transpose_2", "concatenate_1", "conv2d_transpose_3",
          "concatenate_2", "conv2d_transpose_4", "concatenate_3", "lambda_2", "logits"]
KERNELS = {                               # (kh, kw, in, out); transposed: (kh, kw, out, in)
    "conv2d": (5, 5, 1, 2), "conv2d_1": (5, 5, 2, 3), "conv2d_2": (5, 5, 3, 4), "conv2d_3": (5, 5, 4, 4),
    "conv2d_4": (5, 5, 4, 6), "conv2d_5": (5, 5, 6, 6), "conv2d_6": (5, 5, 6, 8),
    "conv2d_transpose": (5, 5, 8, 8), "conv2d_transpose_1": (2, 2, 6, 8), "conv2d_transpose_2": (5, 5, 4, 12),
    "conv2d_transpose_3": (2, 2, 3, 10), "conv2d_transpose_4": (2, 2, 2, 7), "logits": (1, 1, 5, 3),
}
BIAS = {"conv2d_transpose": 8, "conv2d_transpose_1": 6, "conv2d_transpose_2": 4, "conv2d_transpose_3": 3,
        "conv2d_transpose_4": 2}
CHUNKED = "conv2d_3"                      # kernel stored chunked (3, 2, 4, 3)-element chunks -> edge chunks on every axis
COMPACT = "logits"                        # bias stored compact


def expected_weights():
    rng = np.random.RandomState(20240519)
    out = {}
    for name in LAYERS:
        if name in KERNELS:
            k = rng.standard_normal(KERNELS[name]).astype("<f4")
            nb = BIAS.get(name, KERNELS[name][3])
            out[name] = (k, rng.standard_normal(nb).astype("<f4"))
    return out


MODEL_CONFIG = json.dumps({"class_name": "Functional", "config": {
    "name": "fcn_skip", "layers": [{"class_name": "InputLayer", "name": "input_1", "config": {"dtype": "float32", "note": "ü"}}],
    "input_layers": [["input_1", 0, 0]], "output_layers": [["logits", 0, 0]]}, "keras_version": "2.6.0", "backend": "tensorflow"},
    ensure_ascii=False)


class File:
    def __init__(self):
        self.b = bytearray(96)            # superblock filled in last
        self.gcol = None

    def alloc(self, data: bytes, align=8) -> int:
        while len(self.b) % align:
            self.b.append(0)
        a = len(self.b)
        self.b += data
        return a


def pad8(x: bytes) -> bytes:
    return x + b"\x00" * (-len(x) % 8)


def msg(mtype, body, flags=0):
    body = pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def nil(nbytes):
    return struct.pack("<HHB3x", 0, nbytes, 0) + b"\x00" * nbytes


def dt_f32():
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def dt_f64():
    return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def dt_fixed_string(n, pad_type=1, cset=0):                    # 1 = null-padded: what h5py maps numpy 'S' to
    return struct.pack("<BBBBI", 0x13, pad_type | (cset << 4), 0, 0, n)


def dt_vlen_utf8():
    base = struct.pack("<BBBBI", 0x13, 0x00 | (1 << 4), 0, 0, 1)      # 1-byte UTF-8 character type
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) + base   # type = string, pad = null-term, cset = UTF-8


def dataspace(shape, with_max=False):
    b = struct.pack("<BBB5x", 1, len(shape), 1 if with_max else 0)
    b += b"".join(struct.pack("<Q", d) for d in shape)
    if with_max:
        b += b"".join(struct.pack("<Q", d) for d in shape)
    return b


class GlobalHeap:
    """One 4096-byte collection; objects are appended, the rest is the free-space object 0."""

    def __init__(self, f: File):
        self.f = f
        self.addr = f.alloc(bytes(4096))
        self.objs = []

    def add(self, data: bytes):
        self.objs.append(data)
        return self.addr, len(self.objs)

    def finish(self):
        out = bytearray(b"GCOL" + struct.pack("<B3xQ", 1, 4096))
        for i, d in enumerate(self.objs, 1):
            out += struct.pack("<HH4xQ", i, 1, len(d)) + pad8(d)
        free = 4096 - len(out)
        assert free >= 16
        out += struct.pack("<HH4xQ", 0, 0, free)
        out += bytes(4096 - len(out))
        self.f.b[self.addr:self.addr + 4096] = out


def attr_v1(name, dt, ds, data):
    nb = name.encode() + b"\x00"
    return msg(0x000C, struct.pack("<BxHHH", 1, len(nb), len(dt), len(ds)) + pad8(nb) + pad8(dt) + pad8(ds) + data)


def attr_v3(name, dt, ds, data, name_cset=1):
    nb = name.encode() + b"\x00"
    return msg(0x000C, struct.pack("<BBHHHB", 3, 0, len(nb), len(dt), len(ds), name_cset) + nb + dt + ds + data)


def attr_vlen_string(heap: GlobalHeap, name, text):
    raw = text.encode("utf-8")
    addr, idx = heap.add(raw)
    return attr_v3(name, dt_vlen_utf8(), dataspace(()), struct.pack("<IQI", len(raw), addr, idx))


def attr_bytes_array(name, items):
    n = max(len(i) for i in items) if items else 1
    data = b"".join(i.ljust(n, b"\x00") for i in items)       # null-PADDED, no terminator when len == n
    return attr_v1(name, dt_fixed_string(n), dataspace((len(items),)), data)


def object_header(f: File, messages, split_at=None):
    """Version-1 object header.  `split_at`: messages from that index on go to a continuation block."""
    mtime = msg(0x0012, struct.pack("<B3xI", 1, 1621411200))
    first = list(messages if split_at is None else messages[:split_at]) + [mtime]
    rest = [] if split_at is None else list(messages[split_at:])
    count = len(first) + len(rest)
    cont_addr = None
    if rest:
        body2 = b"".join(rest) + nil(24)                       # libhdf5 leaves a NIL gap at the end of grown blocks
        count += 1
        cont_addr = f.alloc(body2)
        first.insert(1, msg(0x0010, struct.pack("<QQ", cont_addr, len(body2))))
        count += 1
    first.append(nil(8))
    count += 1
    body = b"".join(first)
    return f.alloc(struct.pack("<BxHII4x", 1, count, 1, len(body)) + body)


def dataset_contiguous(f, arr, with_max=False):
    data = f.alloc(arr.tobytes())
    layout = struct.pack("<BBQQ", 3, 1, data, arr.nbytes)
    fill_old = msg(0x0004, struct.pack("<I", 0))
    fill_new = msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0), flags=1)
    return object_header(f, [msg(0x0001, dataspace(arr.shape, with_max)), msg(0x0003, dt_f32(), flags=1), fill_old,
                             fill_new, msg(0x0008, layout)])


def dataset_compact(f, arr):
    raw = arr.tobytes()
    layout = struct.pack("<BBH", 3, 0, len(raw)) + raw
    return object_header(f, [msg(0x0001, dataspace(arr.shape)), msg(0x0003, dt_f32(), flags=1),
                             msg(0x0005, struct.pack("<BBBB", 2, 1, 2, 0), flags=1), msg(0x0008, layout)])


def dataset_chunked(f, arr, chunk):
    rank = arr.ndim
    entries = []
    grid = [range(0, arr.shape[i], chunk[i]) for i in range(rank)]
    import itertools
    for offs in itertools.product(*grid):
        c = np.zeros(chunk, dtype="<f4")
        sl = tuple(slice(o, min(o + chunk[i], arr.shape[i])) for i, o in enumerate(offs))
        part = arr[sl]
        c[tuple(slice(0, s) for s in part.shape)] = part
        entries.append((offs, f.alloc(c.tobytes()), c.nbytes))

    def key(nbytes, offs):
        return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

    def node(level, items, final_offs):
        """items: (first chunk offsets, child address, chunk bytes)"""
        out = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
        for offs, child, nbytes in items:
            out += key(nbytes, offs) + struct.pack("<Q", child)
        out += key(0, final_offs)
        capacity = 24 + (2 * 32 + 1) * (8 + 8 * (rank + 1)) + 2 * 32 * 8
        return f.alloc(out + bytes(capacity - len(out)))

    end = tuple(arr.shape)
    # two leaf nodes under one level-1 root, so that the reader has to descend
    half = len(entries) // 2
    leaves = [entries[:half], entries[half:]]
    kids = []
    for k, leaf in enumerate(leaves):
        final = leaves[k + 1][0][0] if k + 1 < len(leaves) else end
        kids.append((leaf[0][0], node(0, leaf, final), leaf[0][2]))
    root = node(1, kids, end)
    layout = struct.pack("<BBBQ", 3, 2, rank + 1, root) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", 4)
    return object_header(f, [msg(0x0001, dataspace(arr.shape, with_max=True)), msg(0x0003, dt_f32(), flags=1),
                             msg(0x0005, struct.pack("<BBBB", 2, 3, 2, 0), flags=1), msg(0x0008, layout)])


def group(f: File, children, attrs, split_at=None):
    """Old-style group: local heap + B-tree of symbol-table nodes holding at most 8 entries each (leaf K = 4; nodes
    split in halves as libhdf5 does when one fills up, so they end up 4..8 full).  -> (header, btree, heap)"""
    names = sorted(children, key=lambda s: s.encode())
    heap = bytearray(8)                                       # offset 0: the empty name
    offs = {}
    for n in sorted(children):                                # heap order = insertion order, unrelated to the key order
        offs[n] = len(heap)
        heap += pad8(n.encode() + b"\x00")
    free_off = len(heap)
    size = max(88, len(heap) + 32)
    size += -size % 8
    heap += bytes(size - len(heap))
    struct.pack_into("<QQ", heap, free_off, 1, size - free_off)   # one free block: next = 1 (H5HL_FREE_NULL), its size
    data_addr = f.alloc(bytes(heap))
    heap_addr = f.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, size, free_off, data_addr))
    nodes = [names[i:i + 5] for i in range(0, len(names), 5)] or [[]]
    if len(nodes) > 1 and len(nodes[-1]) < 2:                 # no nearly empty last node
        nodes[-2:] = [nodes[-2] + nodes[-1]]
    snods = []
    for part in nodes:
        s = b"SNOD" + struct.pack("<BxH", 1, len(part))
        for n in part:
            hdr, cache = children[n]
            if cache:                                         # cached symbol-table info for groups (cache type 1)
                s += struct.pack("<QQII", offs[n], hdr, 1, 0) + struct.pack("<QQ", *cache)
            else:
                s += struct.pack("<QQII", offs[n], hdr, 0, 0) + bytes(16)
        s += bytes(8 + 8 * 40 - len(s))
        snods.append(f.alloc(s))
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods) if names else 0, UNDEF, UNDEF) + struct.pack("<Q", 0)
    if names:
        for part, a in zip(nodes, snods):
            tree += struct.pack("<QQ", a, offs[part[-1]])      # child, then the key = heap offset of its largest name
    tree += bytes(24 + 33 * 8 + 32 * 8 - len(tree))
    tree_addr = f.alloc(tree)
    hdr = object_header(f, [msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + attrs,
                        split_at=None if split_at is None else split_at + 1)
    return hdr, tree_addr, heap_addr


def build() -> bytes:
    f = File()
    gh = GlobalHeap(f)
    W = expected_weights()
    layer_groups = {}
    for name in LAYERS:
        if name in W:
            k, b = W[name]
            kd = dataset_chunked(f, k, (3, 2, 4, 3)) if name == CHUNKED else dataset_contiguous(f, k, with_max=(name == "conv2d"))
            bd = dataset_compact(f, b) if name == COMPACT else dataset_contiguous(f, b)
            inner, t, h = group(f, {"kernel:0": (kd, None), "bias:0": (bd, None)}, [])
            wn = [f"{name}/kernel:0".encode(), f"{name}/bias:0".encode()]
            g, t2, h2 = group(f, {name: (inner, (t, h))}, [attr_bytes_array("weight_names", wn)])
        else:
            # weight-less layers: an empty group whose weight_names attribute is a zero-length float64 array (np.array([]))
            g, t2, h2 = group(f, {}, [attr_v1("weight_names", dt_f64(), dataspace((0,)), b"")])
        layer_groups[name] = (g, (t2, h2))
    names = [n.encode() for n in LAYERS]
    mw_attrs = [attr_bytes_array("layer_names0", names[:13]), attr_bytes_array("layer_names1", names[13:]),
                attr_vlen_string(gh, "backend", "tensorflow"), attr_vlen_string(gh, "keras_version", "2.6.0")]
    mw, t, h = group(f, layer_groups, mw_attrs, split_at=1)
    root_attrs = [attr_vlen_string(gh, "keras_version", "2.6.0"), attr_vlen_string(gh, "backend", "tensorflow"),
                  attr_vlen_string(gh, "model_config", MODEL_CONFIG)]
    root, rt, rh = group(f, {"model_weights": (mw, (t, h))}, root_attrs, split_at=2)
    gh.finish()
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(f.b), UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", rt, rh)
    assert len(sb) == 96
    f.b[:96] = sb
    return bytes(f.b)


if __name__ == "__main__":
    data = build()
    with open(OUT, "wb") as fh:
        fh.write(data)
    print(OUT, len(data), "bytes")
