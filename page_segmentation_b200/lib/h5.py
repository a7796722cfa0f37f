"""Minimal HDF5 reader / writer for Keras 2.x model files (no h5py / libhdf5).

The reference loads models with `tf.keras.models.load_model(path.h5)` and falls
back to `model.load_weights(path)` (ocr4all_pixel_classifier/lib/network.py:75-107);
both go through h5py, which is not installed here.  This module reads the subset
of HDF5 that h5py writes with its default `libver='earliest'` settings and that
Keras' `hdf5_format` uses:

  superblock v0/v1, version-1 object headers (with continuation blocks),
  old-style groups (symbol-table message -> v1 B-tree "TREE" + "SNOD" nodes +
  local "HEAP"), contiguous / compact / chunked-uncompressed datasets of
  little-endian floats, attribute messages v1-v3 holding fixed-length or
  variable-length (global heap "GCOL") strings and numeric arrays.

Version-2 object headers ("OHDR") with compact link messages are also parsed,
so files written by `libver='latest'` with small groups load as well; dense
link / attribute storage (fractal heaps) is rejected with a clear error.

Keras layout: root attrs `keras_version`, `backend`, `model_config`; group
`/model_weights` (or the root itself for `save_weights` files) with attr
`layer_names` and per layer a group with attr `weight_names` and the datasets.
Weights are mapped BY ORDER within `layer_names` (layers that own weights),
never by literal layer name, because Keras layer names carry process-global
counters (`conv2d_7`, ...).

`write_keras_h5` authors files of the same old-style flavour (fixtures, export).
Pins (tests/test_h5_pins.py): the reader is checked against a file written by the
real HDF5 library (MATLAB's -v7.3 writer, shipped with scipy's test data) and
against a Keras-layout file assembled byte by byte from the format specification
by tests/golden/make_keras_h5_fixture.py (variable-length UTF-8 attributes,
continuation blocks, multi-node groups, chunked / compact datasets).  h5py and
Keras themselves are absent here, so a real `model.save()` file is still
unseen; the writer is only pinned by this reader.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


# ---------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------
class _Datatype(NamedTuple):
    cls: int
    size: int
    numpy: Optional[np.dtype]
    vlen_string: bool
    str_pad: int


class _Object:
    def __init__(self):
        self.attrs: Dict[str, object] = {}
        self.links: Dict[str, int] = {}          # child name -> object header address
        self.link_order: List[str] = []
        self.dataspace: Optional[Tuple[int, ...]] = None
        self.datatype: Optional[_Datatype] = None
        self.layout = None                        # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", ...)
        self.is_group = False


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        self._objects: Dict[int, _Object] = {}
        self._parse_superblock()

    # -- primitives --------------------------------------------------------
    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off:off + n], "little")

    def _parse_superblock(self):
        base = -1
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise H5Error("not an HDF5 file (signature missing)")
        ver = self.buf[base + 8]
        self.base = base
        if ver in (0, 1):
            self.O = self.buf[base + 13]
            self.L = self.buf[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            p += 4 * self.O                                   # base addr, free space, eof, driver
            # root group symbol table entry
            self.root_addr = self._u(p + self.O, self.O)
        elif ver in (2, 3):
            self.O = self.buf[base + 9]
            self.L = self.buf[base + 10]
            p = base + 12 + 3 * self.O
            self.root_addr = self._u(p, self.O)
        else:
            raise H5Error(f"unsupported superblock version {ver}")

    # -- object headers ------------------------------------------------------
    def obj(self, addr: int) -> _Object:
        if addr in self._objects:
            return self._objects[addr]
        o = _Object()
        self._objects[addr] = o
        a = self.base + addr
        if self.buf[a:a + 4] == b"OHDR":
            self._parse_ohdr_v2(a, o)
        else:
            self._parse_ohdr_v1(a, o)
        return o

    def _parse_ohdr_v1(self, a: int, o: _Object):
        if self.buf[a] != 1:
            raise H5Error(f"unsupported object header version {self.buf[a]} at {a}")
        nmsg = self._u(a + 2, 2)
        size = self._u(a + 8, 4)
        blocks = [(a + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and seen < nmsg:
                mtype = self._u(p, 2)
                msize = self._u(p + 2, 2)
                body = p + 8
                self._message(mtype, body, msize, o, blocks, v2=False)
                seen += 1
                p = body + msize                         # v1 message bodies are 8-byte padded

    def _parse_ohdr_v2(self, a: int, o: _Object):
        flags = self.buf[a + 5]
        p = a + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szlen = 1 << (flags & 3)
        chunk0 = self._u(p, szlen)
        p += szlen
        blocks = [(p, chunk0)]
        track_order = bool(flags & 0x04)
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 4 <= end:
                mtype = self.buf[p]
                msize = self._u(p + 1, 2)
                body = p + 4 + (2 if track_order else 0)
                if mtype == 0 and msize == 0:
                    break
                self._message(mtype, body, msize, o, blocks, v2=True)
                p = body + msize

    def _message(self, mtype: int, body: int, msize: int, o: _Object, blocks, v2: bool):
        if mtype == 0x0011:                                   # symbol table
            o.is_group = True
            btree = self._u(body, self.O)
            heap = self._u(body + self.O, self.O)
            self._walk_group(btree, heap, o)
        elif mtype == 0x0006:                                 # link message (new-style compact group)
            o.is_group = True
            self._link_message(body, o)
        elif mtype == 0x0002:                                 # link info
            o.is_group = True
            p = body + 2 + (8 if self.buf[body + 1] & 1 else 0)
            fheap = self._u(p, self.O)
            if fheap != UNDEF & ((1 << (8 * self.O)) - 1):
                raise H5Error("dense link storage (fractal heap) is not supported")
        elif mtype == 0x0001:
            o.dataspace = self._dataspace(body)
        elif mtype == 0x0003:
            o.datatype = self._datatype(body)
        elif mtype == 0x0008:
            o.layout = self._layout(body)
        elif mtype == 0x000C:
            name, val = self._attribute(body)
            o.attrs[name] = val
        elif mtype == 0x0015:                                 # attribute info
            p = body + 2 + (2 if self.buf[body + 1] & 1 else 0)
            if self._u(p, self.O) != UNDEF & ((1 << (8 * self.O)) - 1):
                raise H5Error("dense attribute storage (fractal heap) is not supported")
        elif mtype == 0x0010:                                 # continuation
            off = self._u(body, self.O)
            length = self._u(body + self.O, self.L)
            a = self.base + off
            if v2:
                if self.buf[a:a + 4] != b"OCHK":
                    raise H5Error("bad continuation chunk")
                blocks.append((a + 4, length - 8))
            else:
                blocks.append((a, length))

    # -- groups ------------------------------------------------------------
    def _heap_string(self, heap_data: int, off: int) -> str:
        a = heap_data + off
        e = self.buf.index(b"\x00", a)
        return self.buf[a:e].decode("utf-8")

    def _walk_group(self, btree: int, heap: int, o: _Object):
        h = self.base + heap
        if self.buf[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap")
        heap_data = self.base + self._u(h + 8 + 2 * self.L, self.O)
        self._walk_btree(btree, heap_data, o)

    def _walk_btree(self, addr: int, heap_data: int, o: _Object):
        a = self.base + addr
        if self.buf[a:a + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        level = self.buf[a + 5]
        used = self._u(a + 6, 2)
        p = a + 8 + 2 * self.O
        for i in range(used):
            p += self.L                                       # key i
            child = self._u(p, self.O)
            p += self.O
            if level > 0:
                self._walk_btree(child, heap_data, o)
            else:
                self._walk_snod(child, heap_data, o)

    def _walk_snod(self, addr: int, heap_data: int, o: _Object):
        a = self.base + addr
        if self.buf[a:a + 4] != b"SNOD":
            raise H5Error("bad symbol table node")
        n = self._u(a + 6, 2)
        p = a + 8
        for _ in range(n):
            name = self._heap_string(heap_data, self._u(p, self.O))
            ohdr = self._u(p + self.O, self.O)
            o.links[name] = ohdr
            o.link_order.append(name)
            p += 2 * self.O + 8 + 16

    def _link_message(self, body: int, o: _Object):
        flags = self.buf[body + 1]
        p = body + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.buf[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nlen_size = 1 << (flags & 3)
        nlen = self._u(p, nlen_size)
        p += nlen_size
        name = self.buf[p:p + nlen].decode("utf-8")
        p += nlen
        if ltype == 0:
            o.links[name] = self._u(p, self.O)
            o.link_order.append(name)

    # -- messages ------------------------------------------------------------
    def _dataspace(self, body: int) -> Tuple[int, ...]:
        ver = self.buf[body]
        rank = self.buf[body + 1]
        p = body + (8 if ver == 1 else 4)
        return tuple(self._u(p + i * self.L, self.L) for i in range(rank))

    def _datatype(self, body: int) -> _Datatype:
        cls = self.buf[body] & 0x0F
        bits0 = self.buf[body + 1]
        size = self._u(body + 4, 4)
        if cls == 0:                                           # fixed point
            signed = bool(bits0 & 0x08)
            order = ">" if bits0 & 1 else "<"
            return _Datatype(cls, size, np.dtype(f"{order}{'i' if signed else 'u'}{size}"), False, 0)
        if cls == 1:                                           # floating point
            order = ">" if bits0 & 1 else "<"
            return _Datatype(cls, size, np.dtype(f"{order}f{size}"), False, 0)
        if cls == 3:                                           # fixed-length string
            return _Datatype(cls, size, np.dtype(f"S{size}"), False, bits0 & 0x0F)
        if cls == 9:                                           # variable length
            is_string = (bits0 & 0x0F) == 1
            if not is_string:
                raise H5Error("variable-length sequences are not supported")
            return _Datatype(cls, size, None, True, 0)
        raise H5Error(f"unsupported datatype class {cls}")

    def _layout(self, body: int):
        ver = self.buf[body]
        if ver == 3:
            lc = self.buf[body + 1]
            if lc == 0:
                sz = self._u(body + 2, 2)
                return ("compact", self.buf[body + 4:body + 4 + sz])
            if lc == 1:
                return ("contiguous", self._u(body + 2, self.O), self._u(body + 2 + self.O, self.L))
            if lc == 2:
                rank = self.buf[body + 2]
                addr = self._u(body + 3, self.O)
                dims = tuple(self._u(body + 3 + self.O + 4 * i, 4) for i in range(rank))
                return ("chunked", addr, dims)
        elif ver in (1, 2):
            rank = self.buf[body + 1]
            lc = self.buf[body + 2]
            p = body + 8
            if lc == 1:
                addr = self._u(p, self.O)
                return ("contiguous", addr, None)
            if lc == 2:
                addr = self._u(p, self.O)
                dims = tuple(self._u(p + self.O + 4 * i, 4) for i in range(rank))
                return ("chunked", addr, dims)
            if lc == 0:
                dims_end = p + 4 * rank
                sz = self._u(dims_end, 4)
                return ("compact", self.buf[dims_end + 4:dims_end + 4 + sz])
        raise H5Error(f"unsupported data layout message version {ver}")

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        a = self.base + addr
        if self.buf[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = self._u(a + 8, self.L)
        p = a + 8 + self.L
        end = a + size
        while p + 8 + self.L <= end:
            idx = self._u(p, 2)
            osz = self._u(p + 8, self.L)
            if idx == 0:
                break
            if idx == index:
                return self.buf[p + 8 + self.L:p + 8 + self.L + osz]
            p += 8 + self.L + (osz + 7) // 8 * 8
        raise H5Error("global heap object not found")

    def _decode(self, dt: _Datatype, shape: Tuple[int, ...], raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if dt.vlen_string:
            out = []
            step = 4 + self.O + 4
            for i in range(n):
                p = i * step
                addr = int.from_bytes(raw[p + 4:p + 4 + self.O], "little")
                idx = int.from_bytes(raw[p + 4 + self.O:p + 8 + self.O], "little")
                out.append(self._global_heap_object(addr, idx).decode("utf-8") if addr else "")
            return out[0] if not shape else np.array(out, dtype=object).reshape(shape)
        arr = np.frombuffer(raw[:n * dt.size], dtype=dt.numpy, count=n)
        if dt.cls == 3:
            vals = [bytes(v).split(b"\x00")[0].decode("utf-8") for v in arr]
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = arr.astype(dt.numpy.newbyteorder("="))
        return arr.reshape(shape) if shape else arr[0]

    def _attribute(self, body: int):
        ver = self.buf[body]
        nsz = self._u(body + 2, 2)
        tsz = self._u(body + 4, 2)
        ssz = self._u(body + 6, 2)
        p = body + 8 + (1 if ver == 3 else 0)
        pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
        name = self.buf[p:p + nsz].split(b"\x00")[0].decode("utf-8")
        p += pad(nsz)
        dt = self._datatype(p)
        p += pad(tsz)
        shape = self._dataspace(p) if ssz >= 4 else ()
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * (dt.size if not dt.vlen_string else 4 + self.O + 4)
        return name, self._decode(dt, shape, self.buf[p:p + nbytes])

    # -- public --------------------------------------------------------------
    def root(self) -> _Object:
        return self.obj(self.root_addr)

    def child(self, o: _Object, name: str) -> _Object:
        if name not in o.links:
            raise KeyError(name)
        return self.obj(o.links[name])

    def resolve(self, o: _Object, path: str) -> _Object:
        for part in path.strip("/").split("/"):
            if part:
                o = self.child(o, part)
        return o

    def read_dataset(self, o: _Object) -> np.ndarray:
        if o.datatype is None or o.dataspace is None or o.layout is None:
            raise H5Error("object is not a dataset")
        dt, shape = o.datatype, o.dataspace
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * dt.size
        kind = o.layout[0]
        if kind == "compact":
            raw = o.layout[1]
        elif kind == "contiguous":
            addr = o.layout[1]
            if addr == UNDEF & ((1 << (8 * self.O)) - 1):
                raw = b"\x00" * nbytes
            else:
                raw = self.buf[self.base + addr:self.base + addr + nbytes]
        else:
            raw = self._read_chunked(o, nbytes)
        return self._decode(dt, shape, raw)

    def _read_chunked(self, o: _Object, nbytes: int) -> bytes:
        _, addr, cdims = o.layout
        shape = o.dataspace
        esz = o.datatype.size
        rank = len(shape)
        cshape = cdims[:rank]
        arr = np.zeros(shape, dtype=o.datatype.numpy)

        def walk(node):
            a = self.base + node
            if self.buf[a:a + 4] != b"TREE":
                raise H5Error("bad chunk B-tree")
            level = self.buf[a + 5]
            used = self._u(a + 6, 2)
            p = a + 8 + 2 * self.O
            keysz = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize = self._u(p, 4)
                fmask = self._u(p + 4, 4)
                offs = tuple(self._u(p + 8 + 8 * i, 8) for i in range(rank))
                child = self._u(p + keysz, self.O)
                p += keysz + self.O
                if level > 0:
                    walk(child)
                else:
                    if fmask != 0:
                        raise H5Error("filtered (compressed) chunks are not supported")
                    chunk = np.frombuffer(self.buf[self.base + child:self.base + child + csize], dtype=o.datatype.numpy)
                    chunk = chunk[:int(np.prod(cshape))].reshape(cshape)
                    sl = tuple(slice(offs[i], min(offs[i] + cshape[i], shape[i])) for i in range(rank))
                    arr[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(addr)
        return arr.tobytes()


class KerasModel(NamedTuple):
    name: Optional[str]                     # model_config.config.name ('fcn_skip' | 'fcn' | 'unet' | ...) or None
    weights: List[Tuple[np.ndarray, np.ndarray]]
    layer_names: List[str]
    keras_version: Optional[str]


def _str_list(v) -> List[str]:
    if v is None:
        return []
    if isinstance(v, str):
        return [v]
    return [x.decode("utf-8") if isinstance(x, bytes) else str(x) for x in np.asarray(v).reshape(-1)]


def _chunked_attr(o: _Object, name: str) -> List[str]:
    """Keras splits attributes larger than 64 KB into name0, name1, ... (hdf5_format)."""
    if name in o.attrs:
        return _str_list(o.attrs[name])
    out, i = [], 0
    while f"{name}{i}" in o.attrs:
        out += _str_list(o.attrs[f"{name}{i}"])
        i += 1
    return out


def load_keras_model(path: str) -> KerasModel:
    """Reads a Keras full-model file (`model.save`) or weights-only file
    (`save_weights`) and returns the (kernel, bias) pairs of the weighted layers in
    `layer_names` order."""
    f = H5File(path)
    root = f.root()
    name = None
    cfg = root.attrs.get("model_config")
    if cfg is not None:
        try:
            name = json.loads(cfg if isinstance(cfg, str) else str(cfg)).get("config", {}).get("name")
        except (ValueError, AttributeError):
            name = None
    wroot = f.child(root, "model_weights") if "model_weights" in root.links else root
    layer_names = _chunked_attr(wroot, "layer_names")
    if not layer_names:
        raise H5Error(f"{path}: no `layer_names` attribute - not a Keras weights file")
    weights: List[Tuple[np.ndarray, np.ndarray]] = []
    kept: List[str] = []
    for ln in layer_names:
        g = f.child(wroot, ln)
        wnames = _chunked_attr(g, "weight_names")
        if not wnames:
            continue
        arrays = [np.asarray(f.read_dataset(f.resolve(g, wn)), dtype=np.float32) for wn in wnames]
        kernels = [a for a in arrays if a.ndim == 4]
        biases = [a for a in arrays if a.ndim == 1]
        if len(kernels) != 1 or len(biases) != 1:
            raise H5Error(f"{path}: layer {ln} has weights {[a.shape for a in arrays]}; expected one 4-D kernel and "
                          "one bias (only conv / transposed-conv layers are in scope)")
        weights.append((kernels[0], biases[0]))
        kept.append(ln)
    kv = root.attrs.get("keras_version")
    return KerasModel(name, weights, kept, kv if isinstance(kv, str) else None)


# ---------------------------------------------------------------------------
# writer (old-style groups, contiguous datasets, fixed-length string attributes)
# ---------------------------------------------------------------------------
class _Writer:
    O = 8
    L = 8
    LEAF_K = 128            # symbol-table nodes hold up to 2K entries: one node per group suffices

    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def align(self, n: int = 8):
        while len(self.buf) % n:
            self.buf.append(0)

    def write(self, b: bytes) -> int:
        self.align(8)
        off = len(self.buf)
        self.buf += b
        return off

    # messages ---------------------------------------------------------------
    @staticmethod
    def _msg(mtype: int, body: bytes) -> bytes:
        body = body + b"\x00" * (-len(body) % 8)
        return struct.pack("<HHBxxx", mtype, len(body), 0) + body

    @staticmethod
    def _dt_float32() -> bytes:
        # class 1 v1, little-endian, mantissa-normalisation 2 (msb set, not stored), sign bit 31
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dt_string(n: int) -> bytes:
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, n)    # null-terminated ASCII, fixed length n

    @staticmethod
    def _dataspace(shape: Sequence[int]) -> bytes:
        b = struct.pack("<BBBxxxxx", 1, len(shape), 0)
        for d in shape:
            b += struct.pack("<Q", d)
        return b

    def _attr(self, name: str, value) -> bytes:
        nb = name.encode("utf-8") + b"\x00"
        if isinstance(value, (str, bytes)):
            vb = value.encode("utf-8") if isinstance(value, str) else value
            n = max(1, len(vb) + 1)
            dt, ds, data = self._dt_string(n), self._dataspace(()), vb.ljust(n, b"\x00")
        else:
            items = [v.encode("utf-8") if isinstance(v, str) else bytes(v) for v in value]
            n = max([len(v) for v in items] + [0]) + 1
            dt, ds = self._dt_string(n), self._dataspace((len(items),))
            data = b"".join(v.ljust(n, b"\x00") for v in items)
        pad = lambda b: b + b"\x00" * (-len(b) % 8)
        body = struct.pack("<BxHHH", 1, len(nb), len(dt), len(ds)) + pad(nb) + pad(dt) + pad(ds) + data
        return self._msg(0x000C, body)

    def _object_header(self, messages: List[bytes]) -> int:
        body = b"".join(messages)
        if len(body) > 0xFFFF_FFF0:
            raise H5Error("object header too large")
        hdr = struct.pack("<BxHII", 1, len(messages), 1, len(body)) + b"\x00" * 4
        return self.write(hdr + body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.write(arr.tobytes())
        layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data_addr, arr.nbytes)
        msgs = [self._msg(0x0001, self._dataspace(arr.shape)), self._msg(0x0003, self._dt_float32()),
                self._msg(0x0008, layout)]
        return self._object_header(msgs)

    def group(self, children: Dict[str, int], attrs: Dict[str, object]) -> int:
        names = sorted(children)                     # symbol-table entries are ordered by name
        if len(names) > 2 * self.LEAF_K:
            raise H5Error("too many links for a single symbol-table node")
        heap_data = bytearray(b"\x00" * 8)           # offset 0 = empty string
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode("utf-8") + b"\x00"
            while len(heap_data) % 8:
                heap_data.append(0)
        heap_data += b"\x00" * 16                    # room for a free block
        data_addr = self.write(bytes(heap_data))
        free_off = len(heap_data) - 16
        struct.pack_into("<QQ", self.buf, data_addr + free_off, 1, 16)      # free block: next = 1 (none), size
        heap_addr = self.write(b"HEAP" + struct.pack("<Bxxx", 0) + struct.pack("<QQQ", len(heap_data), free_off, data_addr))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
        for n in names:
            snod += struct.pack("<QQII", offs[n], children[n], 0, 0) + b"\x00" * 16
        snod += b"\x00" * ((2 * self.LEAF_K - len(names)) * 40)
        snod_addr = self.write(snod)
        tree = b"TREE" + struct.pack("<BBH", 0, 0, 1 if names else 0) + struct.pack("<QQ", UNDEF, UNDEF)
        tree += struct.pack("<Q", 0)
        if names:
            tree += struct.pack("<QQ", snod_addr, offs[names[-1]])
        tree += b"\x00" * ((2 * 16 + 1) * 8 + 2 * 16 * 8 - (len(tree) - 24))
        tree_addr = self.write(tree)
        msgs = [self._msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        msgs += [self._attr(k, v) for k, v in attrs.items()]
        return self._object_header(msgs), tree_addr, heap_addr


def write_keras_h5(path: str, weights: Sequence[Tuple[np.ndarray, np.ndarray]], model_name: Optional[str] = "fcn_skip",
                   layer_names: Optional[Sequence[str]] = None, weights_only: bool = False,
                   extra_layers: Sequence[str] = ()) -> None:
    """Writes (kernel, bias) pairs in the Keras HDF5 layout.  `extra_layers` adds
    weight-less layer groups (Lambda / pooling / concat) like a real Keras file has."""
    w = _Writer()
    w.buf += b"\x00" * 96                                     # superblock placeholder (v0 with 8-byte offsets)
    if layer_names is None:
        layer_names = []
        for i in range(len(weights)):
            # Keras-style counter names; only `logits` is named explicitly (model.py:88)
            layer_names.append("logits" if i == len(weights) - 1 else (f"conv2d_{i}" if i else "conv2d"))
    layer_groups: Dict[str, int] = {}
    for ln, (k, b) in zip(layer_names, weights):
        kd, bd = w.dataset(k), w.dataset(b)
        inner, _, _ = w.group({"kernel:0": kd, "bias:0": bd}, {})
        g, _, _ = w.group({ln: inner}, {"weight_names": [f"{ln}/kernel:0", f"{ln}/bias:0"]})
        layer_groups[ln] = g
    for ln in extra_layers:
        g, _, _ = w.group({}, {"weight_names": []})
        layer_groups[ln] = g
    all_names = list(layer_names) + list(extra_layers)
    common = {"layer_names": all_names, "backend": "tensorflow", "keras_version": "2.5.0"}
    if weights_only:
        root, tree_addr, heap_addr = w.group(layer_groups, common)
    else:
        mw, _, _ = w.group(layer_groups, common)
        cfg = json.dumps({"class_name": "Functional", "config": {"name": model_name, "layers": []}})
        root, tree_addr, heap_addr = w.group({"model_weights": mw},
                                             {"keras_version": "2.5.0", "backend": "tensorflow", "model_config": cfg})
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", _Writer.LEAF_K, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
    w.buf[0:len(sb)] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
