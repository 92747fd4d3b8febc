"""A second, independently written reader for the subset of HDF5 that SpinED's output uses -- test
infrastructure, deliberately sharing no code with spin-ed_b200/hdf5.py.

No libhdf5/h5py exists in this image, so this module plays the part of "another implementation of
the format": it walks a file strictly by the HDF5 File Format Specification (version 0 superblock,
III.A; version 1 B-trees, III.A.1; symbol table nodes, III.B; local heaps, III.D; version 1 object
headers and their messages, IV.A) and REJECTS anything that deviates -- wrong signatures or version
numbers, non-zero reserved bytes, unsorted B-tree keys, entries past a node's capacity, names that
are not NUL-terminated inside the heap, message sizes that are not multiples of eight, dataspace /
datatype / layout triples whose byte counts disagree, addresses beyond the end-of-file address.
What libhdf5 checks when it opens a file is a subset of this.

    tree = read(path)        # {"name": ndarray | {...}}
"""
import struct

import numpy as np

UNDEFINED = (1 << 64) - 1


class FormatError(Exception):
    pass


def _need(cond, what):
    if not cond:
        raise FormatError(what)


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.eof = None

    def u(self, off, size):
        _need(off + size <= len(self.b), f"read of {size} bytes at {off} past the end of the file")
        return int.from_bytes(self.b[off:off + size], "little")

    def zeros(self, off, size, what):
        _need(bytes(self.b[off:off + size]) == b"\0" * size, f"{what}: reserved bytes at {off} are not zero")

    def addr_ok(self, a, what):
        _need(a != UNDEFINED and a < self.eof, f"{what}: address {a} outside the file (eof {self.eof})")

    # ---- III.A superblock, version 0 ----
    def superblock(self):
        b = self.b
        _need(bytes(b[0:8]) == b"\x89HDF\r\n\x1a\n", "format signature")
        _need(b[8] == 0, "superblock version 0 expected")
        _need(b[9] == 0 and b[10] == 0 and b[12] == 0, "free-space / root-group / shared-header versions must be 0")
        self.zeros(11, 1, "superblock")
        _need(b[13] == 8 and b[14] == 8, "8-byte offsets and lengths expected")
        self.zeros(15, 1, "superblock")
        self.leaf_k, self.internal_k = self.u(16, 2), self.u(18, 2)
        _need(self.leaf_k > 0 and self.internal_k > 0, "B-tree K values must be positive")
        _need(self.u(20, 4) == 0, "file consistency flags")
        base, free, eof, drv = (self.u(24 + 8 * i, 8) for i in range(4))
        _need(base == 0, "base address 0 expected")
        _need(free == UNDEFINED and drv == UNDEFINED, "no free-space info / driver info block expected")
        _need(eof == len(b), f"end-of-file address {eof} != file size {len(b)}")
        self.eof = eof
        return self.symbol_entry(56)

    # ---- III.C symbol table entry (40 bytes) ----
    def symbol_entry(self, off):
        name_off, header, cache = self.u(off, 8), self.u(off + 8, 8), self.u(off + 16, 4)
        self.zeros(off + 20, 4, "symbol table entry")
        _need(cache in (0, 1), "cache type 0 (none) or 1 (group) expected")
        scratch = (self.u(off + 24, 8), self.u(off + 32, 8)) if cache == 1 else None
        if cache == 0:
            self.zeros(off + 24, 16, "symbol table entry scratch pad")
        self.addr_ok(header, "object header")
        return name_off, header, scratch

    # ---- III.D local heap ----
    def heap(self, off):
        _need(bytes(self.b[off:off + 4]) == b"HEAP", f"local heap signature at {off}")
        _need(self.b[off + 4] == 0, "local heap version")
        self.zeros(off + 5, 3, "local heap")
        size, free_head, data = self.u(off + 8, 8), self.u(off + 16, 8), self.u(off + 24, 8)
        self.addr_ok(data, "heap data segment")
        _need(data + size <= self.eof, "heap data segment past the end of the file")
        # free list: (next, size) pairs inside the segment, terminated by 1
        seen, f = 0, free_head
        while f != 1 and f != UNDEFINED:
            _need(f % 8 == 0 and f + 16 <= size, f"heap free block at {f} outside the segment")
            nxt, fsz = self.u(data + f, 8), self.u(data + f + 8, 8)
            _need(fsz >= 16 and f + fsz <= size, "heap free block size")
            f = nxt
            seen += 1
            _need(seen < 10000, "heap free list does not terminate")
        return data, size

    def name(self, heap, off):
        data, size = heap
        _need(off < size, "name offset outside the heap")
        end = off
        while end < size and self.b[data + end] != 0:
            end += 1
        _need(end < size, "name is not NUL-terminated inside the heap")
        return bytes(self.b[data + off:data + end]).decode("ascii")

    # ---- III.A.1 version 1 B-tree, node type 0 (group nodes) + III.B symbol table nodes ----
    def btree(self, off, heap, expect_level=None):
        _need(bytes(self.b[off:off + 4]) == b"TREE", f"B-tree signature at {off}")
        _need(self.b[off + 4] == 0, "B-tree node type 0 (group) expected")
        level, used = self.b[off + 5], self.u(off + 6, 2)
        if expect_level is not None:
            _need(level == expect_level, "B-tree level does not decrease by one")
        _need(used <= 2 * self.internal_k, f"B-tree node uses {used} > 2K entries")
        _need(off + 24 + (2 * self.internal_k + 1) * 8 + 2 * self.internal_k * 8 <= self.eof, "B-tree node truncated")
        out = []
        keys = [self.u(off + 24 + 16 * i, 8) for i in range(used + 1)]
        for i in range(used):
            child = self.u(off + 24 + 16 * i + 8, 8)
            self.addr_ok(child, "B-tree child")
            entries = self.btree(child, heap, level - 1) if level > 0 else self.snod(child, heap)
            if entries:
                # key[i] < every name in child i <= key[i+1] (names compared as strings)
                lo = self.name(heap, keys[i]) if keys[i] else ""
                hi = self.name(heap, keys[i + 1])
                _need(all(lo < n[0] <= hi or (lo == "" and n[0] <= hi) for n in entries), "B-tree keys do not bracket the child's names")
            out += entries
        return out

    def snod(self, off, heap):
        _need(bytes(self.b[off:off + 4]) == b"SNOD", f"symbol table node signature at {off}")
        _need(self.b[off + 4] == 1, "symbol table node version 1")
        self.zeros(off + 5, 1, "symbol table node")
        n = self.u(off + 6, 2)
        _need(n <= 2 * self.leaf_k, f"symbol table node holds {n} > 2K entries")
        _need(off + 8 + 40 * 2 * self.leaf_k <= self.eof, "symbol table node truncated")
        out = []
        for i in range(n):
            name_off, header, scratch = self.symbol_entry(off + 8 + 40 * i)
            out.append((self.name(heap, name_off), header, scratch))
        names = [e[0] for e in out]
        _need(names == sorted(names) and len(set(names)) == len(names), "symbol table entries are not sorted / unique")
        return out

    # ---- IV.A.1 version 1 object header ----
    def messages(self, off):
        _need(self.b[off] == 1, f"object header version 1 expected at {off}")
        self.zeros(off + 1, 1, "object header")
        count, refs, size = self.u(off + 2, 2), self.u(off + 4, 4), self.u(off + 8, 4)
        _need(refs >= 1, "object reference count")
        blocks, out = [(off + 16, size)], []
        while blocks:
            p, left = blocks.pop(0)
            _need(p % 8 == 0, "message block is not 8-byte aligned")
            _need(p + left <= self.eof, "object header block past the end of the file")
            while left >= 8 and len(out) < count:
                mtype, msize, flags = self.u(p, 2), self.u(p + 2, 2), self.b[p + 4]
                self.zeros(p + 5, 3, "message header")
                _need(msize % 8 == 0 and msize <= left - 8, f"message size {msize} (type {mtype:#x})")
                body = bytes(self.b[p + 8:p + 8 + msize])
                if mtype == 0x0010:
                    c_off, c_len = struct.unpack_from("<QQ", body)
                    self.addr_ok(c_off, "continuation block")
                    blocks.append((c_off, c_len))
                out.append((mtype, flags, body))
                p += 8 + msize
                left -= 8 + msize
        _need(len(out) == count, f"object header announces {count} messages, {len(out)} found")
        return out

    # ---- IV.A.2.d datatype message ----
    def datatype(self, body, off=0):
        cls, version = body[off] & 0x0F, body[off] >> 4
        _need(version == 1, "datatype message version 1")
        bits = body[off + 1] | (body[off + 2] << 8) | (body[off + 3] << 16)
        size = struct.unpack_from("<I", body, off + 4)[0]
        if cls == 0:  # fixed point
            _need(bits & 0x1 == 0, "little-endian integers expected")
            bit_off, prec = struct.unpack_from("<HH", body, off + 8)
            _need(bit_off == 0 and prec == 8 * size, "integer precision must fill the element")
            return np.dtype(("<i" if bits & 0x8 else "<u") + str(size)), off + 12
        if cls == 1:  # IEEE floating point
            _need(bits & 0x41 == 0, "little-endian floats expected")
            _need((bits >> 4) & 0x3 == 2, "mantissa normalisation: msb implied")
            bit_off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", body, off + 8)
            want = {4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[size]
            _need(((bits >> 8) & 0xFF, eloc, esize, mloc, msize, bias) == want and bit_off == 0 and prec == 8 * size,
                  "not an IEEE-754 single/double description")
            return np.dtype("<f" + str(size)), off + 20
        if cls == 6:  # compound
            members = bits & 0xFFFF
            p, fields, end_prev = off + 8, [], 0
            for _ in range(members):
                e = body.index(b"\0", p)
                nm = body[p:e].decode("ascii")
                p += (e - p + 1 + 7) // 8 * 8
                m_off = struct.unpack_from("<I", body, p)[0]
                _need(body[p + 4] == 0, "member dimensionality 0 expected")
                p += 4 + 1 + 3 + 4 + 4 + 16
                mdt, p = self.datatype(body, p)
                _need(m_off == end_prev, "compound members must be packed")
                end_prev = m_off + mdt.itemsize
                fields.append((nm, mdt))
            _need(end_prev == size, "compound size is the sum of its members")
            if [f[0] for f in fields] == ["r", "i"] and fields[0][1] == fields[1][1] and fields[0][1].kind == "f":
                return np.dtype("<c" + str(size)), p  # {r, i} pair = complex number
            return np.dtype(fields), p
        raise FormatError(f"datatype class {cls} is not part of SpinED's output")

    def dataset(self, msgs, where):
        by = {}
        for t, _, body in msgs:
            _need(t not in by or t == 0x0010, f"{where}: message type {t:#x} appears twice")
            by[t] = body
        _need({0x0001, 0x0003, 0x0008} <= set(by), f"{where}: dataspace, datatype and layout messages are required")
        sp = by[0x0001]
        _need(sp[0] == 1, "dataspace message version 1")
        rank, flags = sp[1], sp[2]
        _need(flags == 0, "no maximum dimensions / permutation expected")
        _need(sp[3:8] == b"\0" * 5, "dataspace reserved bytes")
        shape = struct.unpack_from(f"<{rank}Q", sp, 8)
        dt, _ = self.datatype(by[0x0003])
        lay = by[0x0008]
        _need(lay[0] == 3 and lay[1] == 1, "version 3 contiguous layout expected")
        addr, nbytes = struct.unpack_from("<QQ", lay, 2)
        count = int(np.prod(shape, dtype=np.int64)) if rank else 1
        _need(nbytes == count * dt.itemsize, f"{where}: layout size {nbytes} != {count} x {dt.itemsize}")
        if nbytes:
            self.addr_ok(addr, where)
            _need(addr + nbytes <= self.eof, f"{where}: raw data past the end of the file")
        return np.frombuffer(self.b, dtype=dt, count=count, offset=addr if nbytes else 0).reshape(shape).copy()

    def group(self, header, scratch, where):
        msgs = self.messages(header)
        stab = [body for t, _, body in msgs if t == 0x0011]
        if not stab:
            return self.dataset(msgs, where)
        _need(len(stab) == 1, "one symbol table message per group")
        bt, hp = struct.unpack_from("<QQ", stab[0])
        if scratch is not None:
            _need(scratch == (bt, hp), f"{where}: cached B-tree/heap addresses differ from the symbol table message")
        self.addr_ok(bt, "group B-tree")
        self.addr_ok(hp, "group heap")
        heap = self.heap(hp)
        _need(self.b[heap[0]] == 0, "heap offset 0 must hold the empty string")
        out = {}
        for name, child, sc in self.btree(bt, heap):
            _need(name not in out, f"{where}/{name}: duplicate name")
            out[name] = self.group(child, sc, f"{where}/{name}")
        return out


def read(path):
    """The whole file as nested dicts of arrays; raises FormatError on any deviation from the format."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    _, root, scratch = r.superblock()
    _need(scratch is not None, "the root entry caches its B-tree and heap addresses")
    return r.group(root, scratch, "")
