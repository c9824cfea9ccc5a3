"""Minimal HDF5 writer/reader for the snapshot files of src/save_data.py (no h5py in this image).

The reference stores a snapshot as `h5py.File(...).create_dataset(name, data=array)` calls with
h5py's defaults (src/save_data.py:16-27): a version-0 superblock, an old-style root group
(symbol table = v1 B-tree + local heap + one symbol-table node) and one version-1 object header
per dataset with a simple dataspace, an IEEE little-endian float datatype and a CONTIGUOUS data
layout.  This module writes exactly that subset of the HDF5 file format specification (v1.1 /
"earliest" libver), so the files open with h5py/libhdf5 and `from_file` of the reference reads
them unchanged, and reads the same subset back (including files with a user block and object
header continuation blocks, which real libhdf5 output has: the reader is tested on an
libhdf5-written file that ships with SciPy).

Only what snapshots need: flat root group, up to 2*LEAF_K datasets, float32/float64/int32/int64/
uint32 arrays of any rank (rank 0 = scalar), no chunking, no compression, no attributes."""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 16        # symbol-table node holds 2*LEAF_K entries (h5py's default is 4 -> 8 entries)
INTERNAL_K = 16

_MSG_DATASPACE, _MSG_DATATYPE, _MSG_FILL, _MSG_LAYOUT, _MSG_CONT, _MSG_SYMTAB = 0x1, 0x3, 0x5, 0x8, 0x10, 0x11


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------
# datatype messages (spec IV.A.2.d): class 0 fixed-point, class 1 floating-point
# ------------------------------------------------------------------------------------------------
def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("big-endian arrays are not supported")
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        bits = dt.itemsize * 8
        head = struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize)   # v1|class 1; implied msb; sign bit
        prop = struct.pack("<HHBBBBI", 0, bits, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(f"unsupported dtype {dt}")


def _parse_datatype(b: bytes) -> np.dtype:
    cls, bits0, _, _, size = struct.unpack_from("<BBBBI", b, 0)
    cls &= 0x0F
    if bits0 & 1:
        raise ValueError("big-endian datasets are not supported")
    if cls == 1:
        return np.dtype("<f%d" % size)
    if cls == 0:
        return np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", size))
    raise ValueError(f"unsupported HDF5 datatype class {cls}")


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    # version 1, reserved, #messages, reference count 1, header size; prefix padded to 16 bytes
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _dataset_header(shape, dt, data_addr, nbytes) -> bytes:
    rank = len(shape)
    space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)
    fill = struct.pack("<BBBB", 2, 2, 2, 0)       # v2: allocate late, write fill if set, undefined
    layout = struct.pack("<BBQQ", 3, 1, data_addr if nbytes else UNDEF, nbytes)   # v3, contiguous
    return _object_header([_message(_MSG_DATASPACE, space), _message(_MSG_DATATYPE, _datatype_message(dt), 1),
                           _message(_MSG_FILL, fill), _message(_MSG_LAYOUT, layout)])


# ------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------
class Writer:
    """with Writer(path) as w: w.create_dataset(name, array) ...   Metadata is laid out when the
    file is closed; array bytes are streamed in create_dataset (arrays may be NumPy arrays or
    anything np.asarray accepts; C order is enforced)."""

    DATA_ALIGN = 4096

    def __init__(self, path):
        self.path = path
        self.f = open(path, "wb")
        self.items = []           # (name, shape, dtype, addr, nbytes)
        self.meta_reserve = 96 + 64 + 544 + 32 + (8 + 2 * LEAF_K * 40) + 2 * LEAF_K * (64 + 160)
        self.meta_reserve += -self.meta_reserve % self.DATA_ALIGN
        self.f.seek(self.meta_reserve)
        self.pos = self.meta_reserve

    def create_dataset(self, name, data):
        if any(name == it[0] for it in self.items):
            raise ValueError(f"dataset {name!r} already exists")
        if len(self.items) >= 2 * LEAF_K:
            raise ValueError("too many datasets for one symbol-table node")
        if not name or "/" in name or len(name.encode()) > 55:
            raise ValueError(f"unsupported dataset name {name!r}")
        arr = np.asarray(data)
        if not arr.flags.c_contiguous:      # (np.ascontiguousarray would turn a scalar into shape (1,))
            arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        _datatype_message(arr.dtype)           # validates the dtype before anything is written
        addr = self.pos
        if arr.nbytes:
            self.f.write(memoryview(arr.reshape(-1)).cast("B"))
            self.pos += arr.nbytes
            pad = -self.pos % 8
            self.f.write(b"\0" * pad)
            self.pos += pad
        self.items.append((name, tuple(arr.shape), arr.dtype, addr, arr.nbytes))

    def close(self):
        if self.f is None:
            return
        items = sorted(self.items, key=lambda it: it[0].encode())    # SNOD entries in strcmp order
        # local heap data segment: offset 0 = empty name, then the names, then one free block
        heap_data, name_off = bytearray(8), {}
        for name, *_ in items:
            name_off[name] = len(heap_data)
            heap_data += _pad8(name.encode() + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)             # last free block (next = 1), 16 bytes long
        a_root = 96
        root_hdr_len = 16 + 8 + 16
        a_btree = a_root + root_hdr_len + (-root_hdr_len % 8)
        a_heap = a_btree + 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        a_heap_data = a_heap + 32
        a_snod = a_heap_data + len(heap_data)
        a_obj = a_snod + 8 + 2 * LEAF_K * 40
        headers, obj_addr = [], {}
        for name, shape, dt, addr, nbytes in items:
            h = _dataset_header(shape, dt, addr, nbytes)
            obj_addr[name] = a_obj
            headers.append(h)
            a_obj += len(h) + (-len(h) % 8)
        if a_obj > self.meta_reserve:
            raise RuntimeError("metadata block overflow")
        eof = max(self.pos, self.meta_reserve)
        root_entry = struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + root_entry
        assert len(sb) == 96
        root = _object_header([_message(_MSG_SYMTAB, struct.pack("<QQ", a_btree, a_heap))])
        last_name = name_off[items[-1][0]] if items else 0
        btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, a_snod, last_name)
        btree += b"\0" * (a_heap - a_btree - len(btree))
        heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, a_heap_data)
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(items))
        for name, *_ in items:
            snod += struct.pack("<QQII16x", name_off[name], obj_addr[name], 0, 0)
        snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
        meta = bytearray(self.meta_reserve)
        for addr, blob in [(0, sb), (a_root, root), (a_btree, btree), (a_heap, heap),
                           (a_heap_data, bytes(heap_data)), (a_snod, snod)]:
            meta[addr:addr + len(blob)] = blob
        for (name, *_), h in zip(items, headers):
            meta[obj_addr[name]:obj_addr[name] + len(h)] = h
        self.f.seek(0)
        self.f.write(meta)
        self.f.truncate(eof)
        self.f.close()
        self.f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write(path, datasets):
    """datasets: mapping or iterable of (name, array)."""
    with Writer(path) as w:
        for name, arr in (datasets.items() if hasattr(datasets, "items") else datasets):
            w.create_dataset(name, arr)


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class Reader:
    """Reader(path).keys(), reader[name] -> np.ndarray (a scalar dataset gives a 0-d array).
    Handles version-0/1 superblocks (any user block), old-style groups, v1 object headers with
    continuation blocks, contiguous and compact layouts."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            self.b = f.read(1 << 22)          # metadata lives at the front of our files
            self.f_size = f.seek(0, 2)
        base = 0
        while self.b[base:base + 8] != SIGNATURE:
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.b):
                raise ValueError(f"{path}: not an HDF5 file")
        ver = self.b[base + 8]
        if ver > 1:
            raise ValueError(f"{path}: superblock version {ver} is not supported (write with libver='earliest')")
        so, sl = self.b[base + 13], self.b[base + 14]
        if (so, sl) != (8, 8):
            raise ValueError("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.b, base + 16)
        o = base + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", self.b, o)
        self.base = base if self.base == 0 and base else self.base
        o += 32
        _, root_hdr, cache, _ = struct.unpack_from("<QQII", self.b, o)
        if cache == 1:
            btree, heap = struct.unpack_from("<QQ", self.b, o + 24)
        else:
            btree, heap = self._symtab_of(root_hdr)
        self.entries = {}
        self._walk(btree, self._heap_data(heap))

    # -- low level -------------------------------------------------------------------------------
    def _at(self, addr, n):
        a = self.base + addr
        if a + n <= len(self.b):
            return self.b[a:a + n]
        with open(self.path, "rb") as f:
            f.seek(a)
            return f.read(n)

    def _heap_data(self, addr):
        h = self._at(addr, 32)
        if h[:4] != b"HEAP":
            raise ValueError("bad local heap signature")
        size, _, data = struct.unpack_from("<QQQ", h, 8)
        return self._at(data, size)

    def _walk(self, addr, heap):
        n = self._at(addr, 24)
        if n[:4] == b"SNOD":
            count = struct.unpack_from("<H", n, 6)[0]
            body = self._at(addr + 8, count * 40)
            for i in range(count):
                off, hdr = struct.unpack_from("<QQ", body, i * 40)
                end = heap.index(b"\0", off)
                self.entries[heap[off:end].decode()] = hdr
            return
        if n[:4] != b"TREE" or n[4] != 0:
            raise ValueError("bad group B-tree node")
        used = struct.unpack_from("<H", n, 6)[0]
        body = self._at(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            self._walk(struct.unpack_from("<Q", body, 8 + 16 * i)[0], heap)

    def _messages(self, addr):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self._at(addr, 16), 0)
        if ver != 1:
            raise ValueError(f"object header version {ver} is not supported")
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < nmsg:
            a, sz = blocks.pop(0)
            blk, o = self._at(a, sz), 0
            while o + 8 <= sz and len(out) < nmsg:
                mtype, msz, flags = struct.unpack_from("<HHB", blk, o)
                body = blk[o + 8:o + 8 + msz]
                if mtype == _MSG_CONT:
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                o += 8 + msz
        return out

    def _symtab_of(self, hdr):
        for mtype, body in self._messages(hdr):
            if mtype == _MSG_SYMTAB:
                return struct.unpack_from("<QQ", body, 0)
        raise ValueError("root group has no symbol table message")

    # -- API -------------------------------------------------------------------------------------
    def keys(self):
        return list(self.entries)

    def __contains__(self, name):
        return name in self.entries

    def info(self, name):
        """(shape, dtype, file offset of the raw data or None, nbytes)."""
        shape = dt = None
        layout = None
        for mtype, body in self._messages(self.entries[name]):
            if mtype == _MSG_DATASPACE:
                ver, rank = body[0], body[1]
                o = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, body, o) if rank else ()
            elif mtype == _MSG_DATATYPE:
                dt = _parse_datatype(body)
            elif mtype == _MSG_LAYOUT:
                layout = body
        if shape is None or dt is None or layout is None:
            raise ValueError(f"{name!r} is not a simple dataset")
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize if shape else dt.itemsize
        if layout[0] == 3 and layout[1] == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            return tuple(shape), dt, (None if addr == UNDEF else self.base + addr), min(nbytes, size) if size else nbytes
        if layout[0] == 3 and layout[1] == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            return tuple(shape), dt, ("compact", bytes(layout[4:4 + size])), nbytes
        if layout[0] in (1, 2) and layout[2] == 1:      # old-style message: rank, class, address, dims
            addr = struct.unpack_from("<Q", layout, 8)[0]
            return tuple(shape), dt, (None if addr == UNDEF else self.base + addr), nbytes
        raise ValueError(f"{name!r}: only contiguous and compact layouts are supported")

    def __getitem__(self, name):
        shape, dt, where, nbytes = self.info(name)
        if where is None:
            return np.zeros(shape, dtype=dt)
        if isinstance(where, tuple):
            return np.frombuffer(where[1], dtype=dt, count=nbytes // dt.itemsize).reshape(shape).copy()
        return np.fromfile(self.path, dtype=dt, count=nbytes // dt.itemsize, offset=where).reshape(shape)

    get = __getitem__


def read(path):
    r = Reader(path)
    return {k: r[k] for k in r.keys()}
