"""Minimal self-contained HDF5 writer/reader for SpinED's output layout.

There is no libhdf5 / h5py in this image, so the file is produced directly from the HDF5 File
Format Specification (version-0 superblock, version-1 object headers, old-style groups = symbol
table B-tree + local heap, contiguous dataset layout).  Layout written, as the reference does
(/root/reference/src/SpinED.hs:267-294,315-337,354-368,406-410):

    /basis/representatives        u64 [N]
    /hamiltonian/eigenvalues      real [k]
    /hamiltonian/eigenvectors     T [k, N]      (Block (N, k) column-major == row-major (k, N))
    /hamiltonian/residuals        real [k]
    /observables/<name>           complex128 [k]  (compound {r, i}; hdf5-hs's own choice is unpinned)
    /_workspace                   (empty group)

Status: validated by round trip through the reader below and by a second, independently written
spec-level reader/validator (tests/h5_spec_reader.py, tests/test_driver_cpu.py); it has NOT been
opened with libhdf5 here because none is available -- stated in DESIGN.md.

The file is opened "WriteAppend" like the reference (SpinED.hs:287).  Opening reads only the
metadata: datasets stay on disk as read-only memory maps (40/42-spin representatives are 7-26 GB).
Closing a modified file APPENDS -- the raw data of new datasets, then a fresh copy of the (few KB
of) metadata: object headers, local heaps, B-tree and symbol-table nodes -- and finally patches the
end-of-file address and the root symbol-table entry in the superblock.  Raw data already in the
file is never read, moved or rewritten; superseded metadata and deleted datasets become unreferenced
space, as they do with libhdf5 until a repack.  A file that was only read is not touched.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16
HEAP_FREE_NULL = 1


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------
# datatype messages
# ------------------------------------------------------------------------------------------
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "u" or dt.kind == "i":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        if dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        elif dt.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        else:
            raise TypeError(dt)
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + struct.pack(
            "<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
    if dt.kind == "c":
        part = np.dtype("f%d" % (dt.itemsize // 2))
        members = b""
        for i, name in enumerate((b"r", b"i")):
            members += _pad8(name + b"\0") + struct.pack("<IB3xI4x4I", i * part.itemsize, 0, 0, 0, 0, 0, 0)
            members += _dtype_message(part)
        return struct.pack("<BBBBI", 0x10 | 6, 2, 0, 0, dt.itemsize) + members
    raise TypeError(f"unsupported dtype {dt}")


def _parse_dtype(buf: bytes, off: int = 0):
    cls_ver, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls = cls_ver & 0x0F
    if cls == 0:
        return np.dtype(("i" if b0 & 0x08 else "u") + str(size)), off + 12
    if cls == 1:
        return np.dtype("f" + str(size)), off + 20
    if cls == 6:
        n = b0 | (b1 << 8)
        p = off + 8
        parts = []
        for _ in range(n):
            end = buf.index(b"\0", p)
            p += (end - p + 1 + 7) // 8 * 8
            p += 4 + 1 + 3 + 4 + 4 + 16
            d, p = _parse_dtype(buf, p)
            parts.append(d)
        if n == 2 and parts[0] == parts[1] and parts[0].kind == "f":
            return np.dtype("c" + str(size)), p
        raise TypeError("unsupported compound datatype")
    raise TypeError(f"unsupported datatype class {cls}")


# ------------------------------------------------------------------------------------------
# in-memory tree
# ------------------------------------------------------------------------------------------
class Group(OrderedDict):
    """name -> Group | numpy array (new data) | _Stored (data already in the file)"""


class _Stored:
    """A dataset whose raw data already lives in the file: address, shape, dtype; mapped on demand."""

    def __init__(self, path, addr, shape, dtype):
        self.path, self.addr, self.shape, self.dtype = path, int(addr), tuple(int(d) for d in shape), np.dtype(dtype)

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def array(self):
        if self.nbytes == 0:
            return np.zeros(self.shape, dtype=self.dtype)
        return np.memmap(self.path, dtype=self.dtype, mode="r", offset=self.addr, shape=self.shape)


class File:
    def __init__(self, path: str, mode: str = "a"):
        self.path = path
        self.root = Group()
        self.dirty = False
        self.existing = False
        if mode in ("a", "r") and os.path.exists(path) and os.path.getsize(path) > 0:
            self.root = _read_file(path)
            self.existing = True
        elif mode == "r":
            raise FileNotFoundError(path)
        else:
            self.dirty = True  # a new file is written even when it stays empty
        self.mode = mode

    # --- h5py-like helpers -----------------------------------------------------------------
    def _walk_impl(self, path: str, create: bool):
        node = self.root
        parts = [p for p in path.split("/") if p]
        for p in parts[:-1]:
            if p not in node:
                if not create:
                    raise KeyError(path)
                node[p] = Group()
            node = node[p]
            if not isinstance(node, Group):
                raise KeyError(f"{p} is a dataset")
        return node, (parts[-1] if parts else "")

    def exists(self, path: str) -> bool:
        try:
            node, leaf = self._walk(path, False)
        except KeyError:
            return False
        return leaf == "" or leaf in node

    def create_group(self, path: str):
        node, leaf = self._walk(path, True)
        if leaf and leaf not in node:
            node[leaf] = Group()
            self.dirty = True

    def _walk(self, path: str, create: bool):
        before = _count_groups(self.root) if create else 0
        out = self._walk_impl(path, create)
        if create and _count_groups(self.root) != before:
            self.dirty = True
        return out

    def write_dataset(self, path: str, array):
        node, leaf = self._walk(path, True)
        node[leaf] = np.ascontiguousarray(array)
        self.dirty = True

    def read_dataset(self, path: str) -> np.ndarray:
        """The dataset as an array; data already in the file comes back as a read-only memory map."""
        node, leaf = self._walk(path, False)
        if leaf not in node or isinstance(node[leaf], Group):
            raise KeyError(path)
        v = node[leaf]
        return v.array() if isinstance(v, _Stored) else v

    def delete(self, path: str):
        node, leaf = self._walk(path, False)
        del node[leaf]
        self.dirty = True

    def close(self):
        if self.mode != "r" and self.dirty:
            _write_file(self.path, self.root, append=self.existing)
            self.dirty = False
            self.existing = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()


def _count_groups(g: Group) -> int:
    return 1 + sum(_count_groups(c) for c in g.values() if isinstance(c, Group))


# ------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, start: int = 96):
        self.chunks = []  # (address, bytes | ndarray)
        self.pos = start  # a new file: the superblock occupies [0, 96); appending: the old end of file

    def alloc(self, size: int) -> int:
        self.pos = (self.pos + 7) // 8 * 8
        addr = self.pos
        self.pos += size
        return addr

    def put(self, addr: int, data):
        self.chunks.append((addr, data))

    def object_header(self, messages) -> int:
        body = b""
        for mtype, data in messages:
            data = _pad8(data)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        hdr = struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body
        addr = self.alloc(len(hdr))
        self.put(addr, hdr)
        return addr

    def dataset(self, arr) -> int:
        if isinstance(arr, _Stored):  # raw data stays where it is; only the (small) header is written anew
            shape, dtype, nbytes, data_addr = arr.shape, arr.dtype, arr.nbytes, arr.addr
        else:
            arr = np.ascontiguousarray(arr)
            shape = arr.shape if arr.ndim else (1,)
            dtype, nbytes = arr.dtype, arr.nbytes
            data_addr = self.alloc(max(arr.nbytes, 1))
            self.put(data_addr, arr)
        space = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, nbytes)
        return self.object_header([(0x0001, space), (0x0003, _dtype_message(dtype)), (0x0008, layout)])

    def group(self, g: Group):
        """-> (object header address, btree address, heap address)"""
        entries = []  # (name, header address, cache type, scratch)
        for name, child in g.items():
            if isinstance(child, Group):
                oh, bt, hp = self.group(child)
                entries.append((name, oh, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self.dataset(child), 0, b"\0" * 16))
        entries.sort(key=lambda e: e[0].encode())
        # local heap: empty string at 0, then the names, then one free block
        heap = bytearray(b"\0" * 8)
        offsets = {}
        for name, *_ in entries:
            offsets[name] = len(heap)
            heap += _pad8(name.encode() + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", HEAP_FREE_NULL, 16)
        heap_data_addr = self.alloc(len(heap))
        self.put(heap_data_addr, bytes(heap))
        heap_addr = self.alloc(32)
        self.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data_addr))
        # symbol table nodes of at most 2 * LEAF_K entries, one B-tree node above them
        cap = 2 * LEAF_K
        leaves = [entries[i:i + cap] for i in range(0, len(entries), cap)] or [[]]
        if len(leaves) > 2 * INTERNAL_K:
            raise ValueError("too many objects in one group for this writer")
        keys, children = [0], []
        for leaf in leaves:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf))
            for name, oh, ctype, scratch in leaf:
                body += struct.pack("<QQI4x", offsets[name], oh, ctype) + scratch
            body += b"\0" * (40 * (cap - len(leaf)))
            addr = self.alloc(len(body))
            self.put(addr, body)
            children.append(addr)
            keys.append(offsets[leaf[-1][0]] if leaf else 0)
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children), UNDEF, UNDEF)
        for i, c in enumerate(children):
            node += struct.pack("<QQ", keys[i], c)
        node += struct.pack("<Q", keys[len(children)])
        node += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(node))
        bt_addr = self.alloc(len(node))
        self.put(bt_addr, node)
        oh = self.object_header([(0x0011, struct.pack("<QQ", bt_addr, heap_addr))])
        return oh, bt_addr, heap_addr


def _superblock(eof: int, oh: int, bt: int, hp: int) -> bytes:
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQI4xQQ", 0, oh, 1, bt, hp)
    assert len(sb) == 96
    return sb


def _write_file(path: str, root: Group, append: bool = False):
    """append: new raw data and a fresh copy of all metadata go behind the current end of file, then
    the superblock is patched (last, so that an interrupted write leaves the old file readable)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    start = 96
    if append:
        start = (os.path.getsize(path) + 7) // 8 * 8
    w = _Writer(start)
    oh, bt, hp = w.group(root)
    eof = (w.pos + 7) // 8 * 8
    with open(path, "r+b" if append else "wb") as f:
        if not append:
            f.write(b"\0" * 96)
        for addr, data in sorted(w.chunks, key=lambda c: c[0]):
            f.seek(addr)
            if isinstance(data, np.ndarray):
                data.tofile(f)
            else:
                f.write(data)
        f.truncate(eof)
        f.flush()
        os.fsync(f.fileno())
        f.seek(0)
        f.write(_superblock(eof, oh, bt, hp))
        f.flush()
        os.fsync(f.fileno())


# ------------------------------------------------------------------------------------------
# reader (for the resume path: /basis/representatives, SpinED.hs:319-328)
# ------------------------------------------------------------------------------------------
def _read_file(path: str) -> Group:
    with open(path, "rb") as f:
        head = f.read(96)
        if head[:8] != SIGNATURE or head[8] != 0:
            raise ValueError(f"{path}: not an HDF5 file with a version-0 superblock")
        if head[13] != 8 or head[14] != 8:
            raise ValueError("only 8-byte offsets/lengths are supported")
        root_oh = struct.unpack_from("<Q", head, 56 + 8)[0]
        return _read_group(f, root_oh)


def _read_messages(f, addr: int):
    """Messages of a version-1 object header, following continuation messages (0x0010) into their blocks."""
    f.seek(addr)
    version, _, nmsg, _, size = struct.unpack("<BBHII", f.read(12))
    if version != 1:
        raise ValueError("only version-1 object headers are supported")
    f.read(4)
    blocks = [f.read(size)]
    out = []
    while blocks and len(out) < nmsg:
        body, p = blocks.pop(0), 0
        while p + 8 <= len(body) and len(out) < nmsg:
            mtype, msize, _ = struct.unpack_from("<HHB", body, p)
            data = body[p + 8:p + 8 + msize]
            out.append((mtype, data))
            if mtype == 0x0010:  # object header continuation: (offset, length) of another block of messages
                c_off, c_len = struct.unpack_from("<QQ", data)
                f.seek(c_off)
                blocks.append(f.read(c_len))
            p += 8 + msize
    return [(t, d) for t, d in out if t != 0x0010]


def _read_group(f, oh_addr: int) -> Group:
    msgs = dict(_read_messages(f, oh_addr))
    g = Group()
    if 0x0011 not in msgs:
        return g
    bt_addr, heap_addr = struct.unpack_from("<QQ", msgs[0x0011])
    f.seek(heap_addr)
    h = f.read(32)
    assert h[:4] == b"HEAP"
    seg_size, _, seg_addr = struct.unpack_from("<QQQ", h, 8)
    f.seek(seg_addr)
    heap = f.read(seg_size)

    def walk(node_addr):
        f.seek(node_addr)
        hdr = f.read(24)
        assert hdr[:4] == b"TREE"
        _, level, used = struct.unpack_from("<BBH", hdr, 4)
        body = f.read((2 * used + 1) * 8)
        for i in range(used):
            child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
            if level > 0:
                yield from walk(child)
            else:
                f.seek(child)
                s = f.read(8)
                assert s[:4] == b"SNOD"
                n = struct.unpack_from("<H", s, 6)[0]
                ents = f.read(40 * n)
                for j in range(n):
                    name_off, oh, ctype = struct.unpack_from("<QQI", ents, 40 * j)
                    name = heap[name_off:heap.index(b"\0", name_off)].decode()
                    yield name, oh

    for name, oh in list(walk(bt_addr)):
        m = dict(_read_messages(f, oh))
        if 0x0011 in m:
            g[name] = _read_group(f, oh)
            continue
        space, dtm, layout = m[0x0001], m[0x0003], m[0x0008]
        rank = space[1]
        dim_off = 8 if space[0] == 1 else 4
        shape = struct.unpack_from("<%dQ" % rank, space, dim_off)
        dt, _ = _parse_dtype(dtm)
        if layout[0] != 3 or layout[1] != 1:
            raise ValueError(f"dataset {name}: only contiguous version-3 layouts are supported")
        addr, nbytes = struct.unpack_from("<QQ", layout, 2)
        g[name] = _Stored(f.name, addr, shape, dt)
    return g
