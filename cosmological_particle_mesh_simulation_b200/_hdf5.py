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


# ------------------------------------------------------------------------------------------------
# structural validator
# ------------------------------------------------------------------------------------------------
def check(path):
    """Walk a file the way libhdf5 does when it opens every object, applying the consistency rules
    its decoders enforce (format spec v1.1, sections II-IV; H5Fsuper, H5Oprefix/chunk, H5HL, H5B,
    H5Gnode, H5S, H5T, H5Dlayout decode routines), plus one global rule: no two allocated regions
    (superblock, object headers, B-tree nodes, heaps, symbol nodes, raw data) may overlap and all must
    end before the end-of-file address.  Returns {name: (shape, dtype)}; raises ValueError on the first
    violation.  Used by the tests on our own files AND on a libhdf5-written one, so the rules are
    known not to be stricter than the library's."""
    r = Reader(path)
    b, base = r.b, r.base
    regions = []

    def need(cond, msg):
        if not cond:
            raise ValueError(f"{path}: {msg}")

    def claim(addr, n, what):
        need(addr != UNDEF, f"{what}: undefined address")
        need(base + addr + n <= r.f_size, f"{what} [{addr}, {addr + n}) runs past the end of the file")
        need(addr + n <= r.eof, f"{what} [{addr}, {addr + n}) runs past the end-of-file address {r.eof}")
        regions.append((addr, addr + n, what))

    ver = b[base + 8]            # (addresses below are relative to the base address)
    need(ver in (0, 1), "superblock version")
    need(b[base + 13] == 8 and b[base + 14] == 8, "sizes of offsets/lengths")
    need(r.leaf_k > 0 and r.internal_k > 0, "group B-tree K values must be non-zero")
    # libhdf5 refuses a file that is shorter than its stored end-of-file address ("truncated file")
    need(r.eof <= r.f_size, f"file is shorter ({r.f_size}) than its end-of-file address ({r.eof})")
    sb_len = 56 + (4 if ver == 1 else 0) + 40
    regions.append((0, sb_len, "superblock"))
    o = base + 24 + (4 if ver == 1 else 0) + 32
    name_off, root_hdr, cache, _ = struct.unpack_from("<QQII", b, o)
    need(cache in (0, 1), "root entry cache type")

    def object_header(addr, what):
        h = r._at(addr, 16)
        v, _, nmsg, nlink, size = struct.unpack_from("<BBHII", h, 0)
        need(v == 1, f"{what}: object header version {v}")
        need(nlink >= 1, f"{what}: link count")
        need(size % 8 == 0, f"{what}: header size {size} is not a multiple of 8")
        claim(addr, 16 + size, what + " header")
        blocks, out = [(addr + 16, size)], []
        while blocks:
            a, sz = blocks.pop(0)
            blk, p = r._at(a, sz), 0
            while p < sz:
                need(p + 8 <= sz, f"{what}: {sz - p} stray bytes at the end of a header block")
                mtype, msz, flags = struct.unpack_from("<HHB", blk, p)
                need(msz % 8 == 0, f"{what}: message 0x{mtype:x} size {msz} is not a multiple of 8")
                need(p + 8 + msz <= sz, f"{what}: message 0x{mtype:x} runs past its header block")
                body = blk[p + 8:p + 8 + msz]
                if mtype == _MSG_CONT:
                    ca, cl = struct.unpack_from("<QQ", body, 0)
                    claim(ca, cl, what + " continuation block")
                    blocks.append((ca, cl))
                out.append((mtype, body))
                p += 8 + msz
        need(len(out) == nmsg, f"{what}: header says {nmsg} messages, found {len(out)}")
        return out

    def heap(addr):
        h = r._at(addr, 32)
        need(h[:4] == b"HEAP" and h[4] == 0, "local heap signature/version")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        claim(addr, 32, "local heap header")
        claim(data, size, "local heap data")
        seg = r._at(data, size)
        free_blocks, seen = [], set()
        while free != 1:                                   # H5HL_FREE_NULL
            need(free % 8 == 0 and free + 16 <= size and free not in seen, f"bad heap free-list offset {free}")
            seen.add(free)
            nxt, fsz = struct.unpack_from("<QQ", seg, free)
            need(fsz >= 16 and free + fsz <= size, "heap free block size")
            free_blocks.append((free, free + fsz))
            free = nxt
        need(seg[:1] == b"\0", "heap offset 0 must hold the empty name")
        return seg, free_blocks

    def group(btree, heap_addr, what):
        seg, free_blocks = heap(heap_addr)
        names = {}

        def name_at(off):
            need(off < len(seg), f"{what}: name offset {off} outside the heap")
            end = seg.find(b"\0", off)
            need(end >= 0, f"{what}: unterminated name")
            need(not any(lo < end + 1 and off < hi for lo, hi in free_blocks), f"{what}: name overlaps a free heap block")
            return seg[off:end]

        def node(addr, level_expected=None):
            h = r._at(addr, 8)
            if h[:4] == b"SNOD":
                need(h[4] == 1, "symbol node version")
                n = struct.unpack_from("<H", h, 6)[0]
                need(n <= 2 * r.leaf_k, f"symbol node holds {n} > 2*{r.leaf_k} entries")
                claim(addr, 8 + 2 * r.leaf_k * 40, "symbol node")
                body, prev, local = r._at(addr + 8, n * 40), None, []
                for i in range(n):
                    off, hdr, ctype = struct.unpack_from("<QQI", body, i * 40)
                    nm = name_at(off)
                    need(prev is None or prev < nm, f"{what}: symbol node entries not in ascending name order")
                    need(ctype in (0, 1, 2), "symbol entry cache type")
                    prev = nm
                    names[nm.decode()] = hdr
                    local.append(nm)
                return local
            need(h[:4] == b"TREE" and h[4] == 0, "group B-tree node signature/type")
            level, used = h[5], struct.unpack_from("<H", h, 6)[0]
            need(used <= 2 * r.internal_k, "B-tree entries used > 2K")
            need(level_expected is None or level == level_expected, "B-tree level")
            claim(addr, 24 + (2 * r.internal_k + 1) * 8 + 2 * r.internal_k * 8, "B-tree node")
            body = r._at(addr + 24, (2 * used + 1) * 8)
            last = None
            for i in range(used):
                key_lo, child, key_hi = struct.unpack_from("<QQQ", body, 16 * i)
                got = node(child, level - 1 if level > 0 else None)
                if got:
                    need(name_at(key_lo) < got[0] or (i == 0 and name_at(key_lo) == b""), "B-tree left key must be below its child")
                    need(name_at(key_hi) == got[-1], "B-tree right key must equal the child's largest name")
                    need(last is None or last < got[0], "B-tree children out of order")
                    last = got[-1]
            return []

        node(btree)
        return names

    msgs = object_header(root_hdr, "root group")
    stab = [body for t, body in msgs if t == _MSG_SYMTAB]
    need(len(stab) == 1, "root group needs exactly one symbol-table message")
    btree, heap_addr = struct.unpack_from("<QQ", stab[0], 0)
    if cache == 1:
        need(struct.unpack_from("<QQ", b, o + 24) == (btree, heap_addr), "root entry scratch pad disagrees with the symbol-table message")
    entries = group(btree, heap_addr, "root group")

    out = {}
    for name, hdr in entries.items():
        msgs = object_header(hdr, f"dataset {name!r}")
        kinds = [t for t, _ in msgs]
        for t, label in ((_MSG_DATASPACE, "dataspace"), (_MSG_DATATYPE, "datatype"), (_MSG_LAYOUT, "layout")):
            need(kinds.count(t) == 1, f"dataset {name!r}: needs exactly one {label} message")
        space = next(body for t, body in msgs if t == _MSG_DATASPACE)
        need(space[0] in (1, 2) and space[1] <= 32, "dataspace version/rank")
        rank, flags = space[1], space[2]
        need(flags & ~0x3 == 0, "dataspace flags")
        dims_off = 8 if space[0] == 1 else 4
        need(len(space) >= dims_off + 8 * rank * (2 if flags & 1 else 1), "dataspace message too short for its dimensions")
        dt_body = next(body for t, body in msgs if t == _MSG_DATATYPE)
        cls, vers = dt_body[0] & 0x0F, dt_body[0] >> 4
        need(vers in (1, 2, 3), "datatype version")
        dsize = struct.unpack_from("<I", dt_body, 4)[0]
        if cls == 1:
            boff, prec, epos, esize, mpos, msize, bias = struct.unpack_from("<HHBBBBI", dt_body, 8)
            sign = dt_body[2]
            need(boff + prec <= 8 * dsize, "float: precision exceeds the size")
            need(sign < prec and epos + esize <= prec and mpos + msize <= prec, "float: field outside the precision")
            need(mpos + msize <= epos or epos + esize <= mpos, "float: exponent and mantissa overlap")
            need(not (epos <= sign < epos + esize) and not (mpos <= sign < mpos + msize), "float: sign bit inside a field")
            need(bias == (1 << (esize - 1)) - 1, "float: exponent bias")
        elif cls == 0:
            boff, prec = struct.unpack_from("<HH", dt_body, 8)
            need(boff + prec <= 8 * dsize, "integer: precision exceeds the size")
        shape, dt, where, nbytes = r.info(name)
        need(dt.itemsize == dsize, "datatype size")
        lay = next(body for t, body in msgs if t == _MSG_LAYOUT)
        if lay[0] == 3 and lay[1] == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
            npts = int(np.prod(shape, dtype=np.int64)) if shape else 1
            need(size == npts * dsize, f"dataset {name!r}: layout size {size} != {npts} x {dsize}")
            if size:
                claim(addr, size, f"dataset {name!r} raw data")
        elif isinstance(where, int):
            claim(where - base, nbytes, f"dataset {name!r} raw data")
        out[name] = (shape, dt)

    regions.sort()
    for (a0, a1, w0), (b0, b1, w1) in zip(regions, regions[1:]):
        need(a1 <= b0, f"{w0} [{a0}, {a1}) overlaps {w1} [{b0}, {b1})")
    return out
