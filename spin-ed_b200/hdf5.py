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

Status: validated by round trip through the reader below (tests/test_driver_cpu.py); it has NOT
been opened with libhdf5 here because none is available -- stated in DESIGN.md.

The file is opened "WriteAppend" like the reference (SpinED.hs:287): existing content is read,
datasets are added or overwritten, and the whole file is rewritten on close.
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
    """name -> Group | numpy array"""


class File:
    def __init__(self, path: str, mode: str = "a"):
        self.path = path
        self.root = Group()
        if mode in ("a", "r") and os.path.exists(path) and os.path.getsize(path) > 0:
            self.root = _read_file(path)
        elif mode == "r":
            raise FileNotFoundError(path)
        self.mode = mode

    # --- h5py-like helpers -----------------------------------------------------------------
    def _walk(self, path: str, create: bool):
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

    def write_dataset(self, path: str, array):
        node, leaf = self._walk(path, True)
        node[leaf] = np.ascontiguousarray(array)

    def read_dataset(self, path: str) -> np.ndarray:
        node, leaf = self._walk(path, False)
        if leaf not in node or isinstance(node[leaf], Group):
            raise KeyError(path)
        return node[leaf]

    def delete(self, path: str):
        node, leaf = self._walk(path, False)
        del node[leaf]

    def close(self):
        if self.mode != "r":
            _write_file(self.path, self.root)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()


# ------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.chunks = []  # (address, bytes | ndarray)
        self.pos = 96     # superblock occupies [0, 96)

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

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        shape = arr.shape if arr.ndim else (1,)
        data_addr = self.alloc(max(arr.nbytes, 1))
        self.put(data_addr, arr)
        space = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        return self.object_header([(0x0001, space), (0x0003, _dtype_message(arr.dtype)), (0x0008, layout)])

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


def _write_file(path: str, root: Group):
    w = _Writer()
    oh, bt, hp = w.group(root)
    eof = (w.pos + 7) // 8 * 8
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQI4xQQ", 0, oh, 1, bt, hp)
    assert len(sb) == 96
    tmp = path + ".tmp"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(tmp, "wb") as f:
        f.write(sb)
        for addr, data in sorted(w.chunks, key=lambda c: c[0]):
            f.seek(addr)
            if isinstance(data, np.ndarray):
                data.tofile(f)
            else:
                f.write(data)
        f.truncate(eof)
    os.replace(tmp, path)


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
    f.seek(addr)
    version, _, nmsg, _, size = struct.unpack("<BBHII", f.read(12))
    if version != 1:
        raise ValueError("only version-1 object headers are supported")
    f.read(4)
    body = f.read(size)
    out, p = [], 0
    while p + 8 <= len(body) and len(out) < nmsg:
        mtype, msize, _ = struct.unpack_from("<HHB", body, p)
        out.append((mtype, body[p + 8:p + 8 + msize]))
        p += 8 + msize
    return out


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
        f.seek(addr)
        g[name] = np.fromfile(f, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    return g
