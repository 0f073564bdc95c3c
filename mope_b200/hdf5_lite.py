"""Minimal HDF5 writer / reader for the reference's transition-table files -- no libhdf5, no h5py.

The reference stores its tables as HDF5 (mope/generate_transitions.py:242-323): a flat root group of 2-D float64 datasets
`P<idx>` carrying scalar attributes (`N, s, u, v, gen`, or `N, Nb, u` for bottleneck files) and root attributes `N`,
`breaks`, `frequencies`, `min_bin_size`, `uniform_weight`; mope/transition_data_2d.py:74-85 reads them back with h5py.
Neither h5py nor libhdf5 is available in the build image, so this module writes and parses the subset of the HDF5 file
format (specification version 1.1 objects, the ones `h5py.File(name, "w")` with default settings produces) directly:

    superblock version 0, one root group as a symbol table (local heap + version-1 B-tree with a single leaf + one
    symbol-table node), version-1 object headers (with continuation blocks when reading), dataspace version 1, datatype
    classes 0 (fixed point), 1 (IEEE floats) and 3 (fixed-length strings), contiguous (layout version 3; versions 1 / 2
    when reading) and compact layouts, attribute messages version 1 (1-3 when reading).

Not supported: chunked / filtered datasets, new-style (link-message / fractal-heap) groups, nested groups, variable-length
types.  The reader is checked against an HDF5 file written by an independent implementation (MATLAB's, shipped in SciPy's
test data) in tests/test_hdf5_lite.py; the writer against the reader.  The group uses one symbol-table node sized through
the superblock's "group leaf node K", which keeps the writer to a single B-tree leaf for any number of datasets.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ----------------------------------------------------------------------------------------------------------------------
# datatype / dataspace encoding
# ----------------------------------------------------------------------------------------------------------------------
def _encode_dtype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.float64:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt == np.float32:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)   # null-terminated, ASCII
    raise TypeError("hdf5_lite: unsupported dtype %r" % (dt,))


def _decode_dtype(b: bytes) -> np.dtype:
    cls = b[0] & 0x0F
    bits0, bits1 = b[1], b[2]
    size = struct.unpack_from("<I", b, 4)[0]
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:
        return np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size))
    if cls == 1:
        return np.dtype("%sf%d" % (order, size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise TypeError("hdf5_lite: unsupported datatype class %d" % cls)


def _encode_space(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _decode_space(b: bytes) -> Tuple[int, ...]:
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if b[3] == 2:   # null dataspace
            return (0,)
    else:
        raise ValueError("hdf5_lite: dataspace version %d" % ver)
    return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank))


def _as_array(v) -> np.ndarray:
    if isinstance(v, (bytes, str)):
        raw = v.encode() if isinstance(v, str) else v
        return np.array(raw, dtype="S%d" % (len(raw) + 1))
    a = np.asarray(v)
    if a.dtype == np.bool_:
        a = a.astype(np.int8)
    if a.dtype.kind == "U":
        a = a.astype("S")
        a = a.astype("S%d" % (a.dtype.itemsize + 1))
    if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
        a = a.astype(np.float64)
    return a


def _attr_message(name: str, value) -> bytes:
    a = _as_array(value)
    nm = name.encode() + b"\0"
    dt = _encode_dtype(a.dtype)
    sp = _encode_space(a.shape)
    data = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "iuf" else a).tobytes()
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data
    return _message(0x000C, body)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    if len(body) > 0xFFFF:
        raise ValueError("hdf5_lite: header message of %d bytes exceeds the 64 KB limit" % len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


# ----------------------------------------------------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------------------------------------------------
def write_file(path: str, root_attrs: Dict[str, object], datasets: Dict[str, Tuple[np.ndarray, Dict[str, object]]]) -> None:
    """Write a flat HDF5 file: root attributes + named datasets, each with its own attributes."""
    names = sorted(datasets.keys(), key=lambda s: s.encode())   # symbol-table nodes are ordered by strcmp of the names
    n = len(names)
    leaf_k = max(4, (n + 1) // 2)
    if leaf_k > 0xFFFF:
        raise ValueError("hdf5_lite: too many datasets for one symbol-table node (%d)" % n)
    internal_k = 16

    # local heap data: offset 0 = empty string (the B-tree's first key), then the names
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for nm in names:
        name_off[nm] = len(heap)
        heap += _pad8(nm.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)      # free block: next = 1 (none, the library's end marker), size 16
    heap_data = bytes(heap)

    out = bytearray()

    def alloc(b: bytes, align: int = 8) -> int:
        out.extend(b"\0" * (-len(out) % align))
        addr = len(out)
        out.extend(b)
        return addr

    sb_size = 8 + 8 + 4 + 4 + 4 * 8 + 40          # superblock version 0 with 8-byte offsets / lengths
    out.extend(b"\0" * sb_size)

    # root group: object header, B-tree node, local heap, symbol-table node
    root_msgs_tail = [_attr_message(k, v) for k, v in root_attrs.items()]
    # addresses are needed inside the root header, so lay the fixed-size structures out first
    btree_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    snod_size = 8 + 2 * leaf_k * 40
    btree_addr = alloc(b"\0" * btree_size)
    heap_addr = alloc(b"\0" * 32)
    heap_data_addr = alloc(heap_data)
    snod_addr = alloc(b"\0" * snod_size)
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + root_msgs_tail)
    root_addr = alloc(root_hdr)

    # datasets: object header, then the raw data
    entries = []
    for nm in names:
        arr, attrs = datasets[nm]
        a = _as_array(arr)
        raw = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "iuf" else a).tobytes()
        data_addr = alloc(raw) if len(raw) else UNDEF
        msgs = [
            _message(0x0001, _encode_space(a.shape)),
            _message(0x0003, _encode_dtype(a.dtype), flags=1),
            _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),                      # fill value: late allocation, undefined
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),       # contiguous layout
        ] + [_attr_message(k, v) for k, v in attrs.items()]
        entries.append((name_off[nm], alloc(_object_header(msgs))))

    eof = len(out)
    # local heap header
    out[heap_addr:heap_addr + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)
    # symbol-table node
    snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, n))
    for off, addr in entries:
        snod += struct.pack("<QQII16x", off, addr, 0, 0)
    snod += b"\0" * (snod_size - len(snod))
    out[snod_addr:snod_addr + snod_size] = snod
    # B-tree: one leaf node (level 0) pointing at the symbol-table node; keys are heap offsets of names
    bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if n else 0, UNDEF, UNDEF))
    bt += struct.pack("<Q", 0)
    if n:
        bt += struct.pack("<QQ", snod_addr, name_off[names[-1]])
    bt += b"\0" * (btree_size - len(bt))
    out[btree_addr:btree_addr + btree_size] = bt
    # superblock
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", leaf_k, internal_k, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == sb_size
    out[0:sb_size] = sb
    with open(path, "wb") as f:
        f.write(bytes(out))


# ----------------------------------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        start = 0
        while True:   # the superblock may sit behind a user block of 512, 1024, ... bytes
            if buf[start:start + 8] == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
            if start >= len(buf):
                raise ValueError("hdf5_lite: not an HDF5 file")
        ver = buf[start + 8]
        if ver > 1:
            raise ValueError("hdf5_lite: superblock version %d (new-style files) is not supported" % ver)
        if buf[start + 13] != 8 or buf[start + 14] != 8:
            raise ValueError("hdf5_lite: only 8-byte offsets and lengths are supported")
        p = start + 16
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, p)
        p += 8
        if ver == 1:
            p += 4
        self.base = struct.unpack_from("<Q", buf, p)[0]
        p += 32
        _, self.root_addr, cache, _ = struct.unpack_from("<QQII", buf, p)

    def at(self, addr: int) -> int:
        return self.base + addr

    def messages(self, addr: int):
        """(type, flags, body) of every message of a version-1 object header, following continuation blocks."""
        p = self.at(addr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.b, p)
        if ver != 1:
            raise ValueError("hdf5_lite: object header version %d is not supported" % ver)
        blocks = [(p + 16, hsize)]
        out = []
        while blocks:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.b, q)
                body = self.b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, flags, body))
        return out

    def attribute(self, body: bytes):
        ver = body[0]
        if ver == 1:
            _, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", body, 0)
            p = 8
            name = body[p:p + nsz].split(b"\0")[0].decode(); p += nsz + (-nsz % 8)
            dt = _decode_dtype(body[p:p + dsz]); p += dsz + (-dsz % 8)
            shape = _decode_space(body[p:p + ssz]); p += ssz + (-ssz % 8)
        elif ver in (2, 3):
            _, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", body, 0)
            p = 8 + (1 if ver == 3 else 0)
            name = body[p:p + nsz].split(b"\0")[0].decode(); p += nsz
            dt = _decode_dtype(body[p:p + dsz]); p += dsz
            shape = _decode_space(body[p:p + ssz]); p += ssz
        else:
            raise ValueError("hdf5_lite: attribute message version %d" % ver)
        count = int(np.prod(shape)) if shape else 1
        a = np.frombuffer(body, dtype=dt, count=count, offset=p).reshape(shape)
        return name, self._native(a)

    @staticmethod
    def _native(a: np.ndarray):
        if a.dtype.kind == "S":
            return a.item().split(b"\0")[0].decode() if a.shape == () else np.char.decode(a)
        a = a.astype(a.dtype.newbyteorder("="))
        return a[()] if a.shape == () else a.copy()

    def object(self, addr: int):
        """(array or None, attrs) of the object whose header is at addr."""
        attrs, shape, dt, data = {}, None, None, None
        layout = None
        for mtype, _, body in self.messages(addr):
            if mtype == 0x000C:
                k, v = self.attribute(body)
                attrs[k] = v
            elif mtype == 0x0001:
                shape = _decode_space(body)
            elif mtype == 0x0003:
                dt = _decode_dtype(body)
            elif mtype == 0x0008:
                layout = body
        if layout is not None and dt is not None and shape is not None:
            ver = layout[0]
            count = int(np.prod(shape)) if shape else 1
            if ver == 3:
                cls = layout[1]
                if cls == 1:
                    daddr, _ = struct.unpack_from("<QQ", layout, 2)
                    data = np.frombuffer(self.b, dtype=dt, count=count, offset=self.at(daddr)) if daddr != UNDEF else np.zeros(count, dt)
                elif cls == 0:
                    sz = struct.unpack_from("<H", layout, 2)[0]
                    data = np.frombuffer(layout[4:4 + sz], dtype=dt, count=count)
                else:
                    raise ValueError("hdf5_lite: chunked datasets are not supported")
            elif ver in (1, 2):
                rank, cls = layout[1], layout[2]
                if cls != 1:
                    raise ValueError("hdf5_lite: only contiguous version-%d layouts are supported" % ver)
                daddr = struct.unpack_from("<Q", layout, 8)[0]
                data = np.frombuffer(self.b, dtype=dt, count=count, offset=self.at(daddr))
            else:
                raise ValueError("hdf5_lite: layout message version %d" % ver)
            data = self._native(data.reshape(shape))
        return data, attrs

    def group_entries(self, btree_addr: int, heap_addr: int):
        hp = self.at(heap_addr)
        if self.b[hp:hp + 4] != b"HEAP":
            raise ValueError("hdf5_lite: bad local heap signature")
        data_addr = self.at(struct.unpack_from("<Q", self.b, hp + 24)[0])

        def name_at(off):
            end = self.b.index(b"\0", data_addr + off)
            return self.b[data_addr + off:end].decode()

        out = []

        def walk(addr):
            p = self.at(addr)
            sig = self.b[p:p + 4]
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
                if ntype != 0:
                    raise ValueError("hdf5_lite: not a group B-tree")
                q = p + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", self.b, q + 8 + 16 * i)[0]
                    walk(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", self.b, p + 6)[0]
                for i in range(nsym):
                    off, oaddr = struct.unpack_from("<QQ", self.b, p + 8 + 40 * i)
                    out.append((name_at(off), oaddr))
            else:
                raise ValueError("hdf5_lite: unexpected node signature %r" % sig)
        walk(btree_addr)
        return out

    def read(self):
        root_attrs = {}
        btree = heap = None
        for mtype, _, body in self.messages(self.root_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", body, 0)
            elif mtype == 0x000C:
                k, v = self.attribute(body)
                root_attrs[k] = v
        if btree is None:
            raise ValueError("hdf5_lite: the root group is not a symbol-table group (file written with libver='latest'?)")
        datasets = {}
        for name, addr in self.group_entries(btree, heap):
            data, attrs = self.object(addr)
            if data is not None:
                datasets[name] = (data, attrs)
        return root_attrs, datasets


def read_file(path: str):
    """-> (root_attrs, {name: (array, attrs)}) of a flat, old-style HDF5 file."""
    with open(path, "rb") as f:
        buf = f.read()
    return _Reader(buf).read()
