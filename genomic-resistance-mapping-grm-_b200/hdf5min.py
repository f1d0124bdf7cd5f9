"""Minimal HDF5 writer + reader (no h5py / libhdf5 in the image) for the Kover dataset layout.

Writes the most widely readable flavour of the file format (HDF5 File Format Specification v1.1):
superblock version 0, version-1 object headers, old-style groups (symbol table = v1 B-tree + local
heap + one symbol node), version-3 data layout messages, contiguous or chunked storage with a v1
chunk B-tree and the deflate filter, version-1 attributes.  Only what ``.kover`` files need
(Appendix A of SURVEY.md; producer create.py:311-356 and :214-238, consumer ds.py:26-148):
1-D / 2-D datasets of uint8/16/32/64 and fixed-length byte strings, scalar float64 / string
attributes on the root group and on datasets.

The reader parses the same subset and is used by the tests and by ``from_tsv``-style consumers.
Host-side code by design (BASELINE.json: "HDF5/gzip writing stays on the host"); chunks are
deflated on a thread pool (zlib releases the GIL).
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K = 4           # symbol node holds up to 2K = 8 entries
GROUP_INTERNAL_K = 16
CHUNK_BTREE_K = 32         # default "indexed storage internal node K" of a version-0 superblock
HEAP_FREE_NULL = 1         # libhdf5's H5HL_FREE_NULL: "no free block" marker of a local heap


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- message encoders -------------------------------------------------------------------------
def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "S":
        # class 3 (string), version 1; null-padded, ASCII
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    if dt.kind == "u":
        return struct.pack("<BBBBIHH", 0x10, 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "i":
        return struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    raise TypeError(f"unsupported dtype {dt}")


def _space_msg(shape) -> bytes:
    shape = tuple(int(x) for x in shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _attr_msg(name: str, value) -> bytes:
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        arr = np.array(value if len(value) else b"\0", dtype=f"S{max(1, len(value))}")
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
    nm = name.encode("utf-8") + b"\0"
    dt, sp = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + arr.tobytes()
    return body


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages: list[tuple[int, bytes]]) -> bytes:
    body = b"".join(_message(t, d) for t, d in messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class H5Writer:
    """Sequential writer: datasets are streamed to the file, metadata is written on close()."""

    def __init__(self, path: str, threads: int = 8):
        self.f = open(path, "wb")
        self.f.write(b"\0" * 96)            # superblock + root symbol table entry, filled in on close()
        self.attrs: dict = {}
        self._datasets: list[tuple[str, int]] = []   # (name, object header address)
        self._pool = ThreadPoolExecutor(max_workers=max(1, threads))

    # -- low level ---------------------------------------------------------------------------------
    def _alloc_write(self, data: bytes) -> int:
        pos = self.f.tell()
        pad = -pos % 8
        if pad:
            self.f.write(b"\0" * pad)
            pos += pad
        self.f.write(data)
        return pos

    def _chunk_btree(self, entries, ndims: int, chunk_shape) -> int:
        """entries: sorted list of (offsets tuple, address, nbytes).  Returns the root node address."""
        key_fmt = "<II" + "Q" * (ndims + 1)
        key_size = struct.calcsize(key_fmt)
        node_size = 24 + (2 * CHUNK_BTREE_K + 1) * key_size + 2 * CHUNK_BTREE_K * 8

        def key(nbytes, offs):
            return struct.pack(key_fmt, nbytes, 0, *offs, 0)

        def end_key(last_offs):
            o = list(last_offs)
            o[-1] += chunk_shape[-1]
            return key(0, o)

        level = 0
        # (first offsets, last offsets, address, nbytes) per child
        children = [(o, o, a, n) for o, a, n in entries]
        while True:
            nodes = []
            for i in range(0, max(1, len(children)), 2 * CHUNK_BTREE_K):
                grp = children[i:i + 2 * CHUNK_BTREE_K]
                body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), UNDEF, UNDEF)
                for first, _last, addr, nbytes in grp:
                    body += key(nbytes if level == 0 else 0, first) + struct.pack("<Q", addr)
                body += end_key(grp[-1][1]) if grp else key(0, (0,) * ndims)
                body += b"\0" * (node_size - len(body))
                nodes.append((grp[0][0] if grp else (0,) * ndims, grp[-1][1] if grp else (0,) * ndims, body))
            # link siblings
            addrs = []
            base = self.f.tell() + (-self.f.tell() % 8)
            for j in range(len(nodes)):
                addrs.append(base + j * (node_size + (-node_size % 8)))
            out = []
            for j, (first, last, body) in enumerate(nodes):
                left = addrs[j - 1] if j > 0 else UNDEF
                right = addrs[j + 1] if j + 1 < len(nodes) else UNDEF
                body = body[:8] + struct.pack("<QQ", left, right) + body[24:]
                a = self._alloc_write(body)
                assert a == addrs[j]
                out.append((first, last, a, 0))
            if len(out) == 1:
                return out[0][2]
            children = out
            level += 1

    # -- datasets ----------------------------------------------------------------------------------
    def create_dataset(self, name: str, data, chunks=None, gzip: int = 0, attrs: dict | None = None) -> None:
        arr = np.ascontiguousarray(data)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype.kind == "S" and arr.dtype.itemsize == 0:
            arr = arr.astype("S1")
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        shape = arr.shape
        msgs: list[tuple[int, bytes]] = [(0x0001, _space_msg(shape)), (0x0003, _dtype_msg(arr.dtype))]
        if chunks is None and gzip > 0 and arr.size:
            # h5py picks a chunk shape itself when compression is requested; any shape is valid
            if arr.ndim == 1:
                chunks = (min(shape[0], max(1, (1 << 20) // arr.dtype.itemsize)),)
            else:
                chunks = (1,) * (arr.ndim - 1) + (min(shape[-1], max(1, (1 << 20) // arr.dtype.itemsize)),)
        if chunks is not None and arr.size:
            chunks = tuple(int(min(max(1, c), max(1, s))) for c, s in zip(chunks, shape))
            msgs.append((0x0005, struct.pack("<BBBBI", 2, 3, 0, 1, 0)))       # fill value v2, incremental alloc
            if gzip > 0:
                name_b = b"deflate\0"
                msgs.append((0x000B, struct.pack("<BB2x4x", 1, 1) + struct.pack("<HHHH", 1, len(name_b), 1, 1)
                             + name_b + struct.pack("<I", int(gzip)) + b"\0" * 4))
            entries = self._write_chunks(arr, chunks, gzip)
            root = self._chunk_btree(entries, arr.ndim, chunks)
            layout = struct.pack("<BBB", 3, 2, arr.ndim + 1) + struct.pack("<Q", root)
            layout += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.dtype.itemsize)
            msgs.append((0x0008, layout))
        else:
            msgs.append((0x0005, struct.pack("<BBBBI", 2, 1, 0, 1, 0)))       # fill value v2, early alloc
            addr = self._alloc_write(arr.tobytes()) if arr.size else UNDEF
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
        for k, v in (attrs or {}).items():
            msgs.append((0x000C, _attr_msg(k, v)))
        self._datasets.append((name, self._alloc_write(_object_header(msgs))))

    def _write_chunks(self, arr, chunks, gzip):
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
        offsets = [()]
        for r in grid:
            offsets = [o + (x,) for o in offsets for x in r]

        def make(off):
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, chunks, arr.shape))
            block = arr[sl]
            if block.shape != tuple(chunks):                 # edge chunks are stored full size
                full = np.zeros(chunks, dtype=arr.dtype)
                full[tuple(slice(0, n) for n in block.shape)] = block
                block = full
            raw = np.ascontiguousarray(block).tobytes()
            return zlib.compress(raw, gzip) if gzip > 0 else raw

        entries = []
        step = 64
        for i in range(0, len(offsets), step):
            batch = offsets[i:i + step]
            for off, blob in zip(batch, self._pool.map(make, batch)):
                entries.append((off, self._alloc_write(blob), len(blob)))
        return entries

    # -- close: group structures + superblock --------------------------------------------------------
    def close(self) -> None:
        if self.f is None:
            return
        if len(self._datasets) > 2 * GROUP_LEAF_K:
            raise ValueError("hdf5min supports at most 8 objects in the root group")
        ds = sorted(self._datasets, key=lambda x: x[0].encode())
        heap_data = b"\0" * 8
        name_off = []
        for name, _ in ds:
            name_off.append(len(heap_data))
            heap_data += _pad8(name.encode("utf-8") + b"\0")
        # room for a few more names (kover dataset split adds a "splits" group later), as one free block
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", HEAP_FREE_NULL, 64) + b"\0" * 48
        heap_data_addr = self._alloc_write(heap_data)
        heap_addr = self._alloc_write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(ds))
        for (name, addr), off in zip(ds, name_off):
            snod += struct.pack("<QQII16x", off, addr, 0, 0)
        snod += b"\0" * (8 + 2 * GROUP_LEAF_K * 40 - len(snod))
        snod_addr = self._alloc_write(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if ds else 0, UNDEF, UNDEF)
        tree += struct.pack("<Q", 0)
        if ds:
            tree += struct.pack("<QQ", snod_addr, name_off[-1])
        tree += b"\0" * (24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8 - len(tree))
        tree_addr = self._alloc_write(tree)
        msgs = [(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in self.attrs.items():
            msgs.append((0x000C, _attr_msg(k, v)))
        root_addr = self._alloc_write(_object_header(msgs))
        eof = self.f.tell()
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        assert len(sb) == 96
        self.f.seek(0)
        self.f.write(sb)
        self.f.close()
        self.f = None
        self._pool.shutdown()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- reader ------------------------------------------------------------------------------------------
class H5Dataset:
    def __init__(self, reader, name, shape, dtype, layout, filters, attrs):
        self._r, self.name, self.shape, self.dtype, self._layout, self._filters, self.attrs = \
            reader, name, shape, dtype, layout, filters, attrs
        self.chunks = layout.get("chunks")

    def read(self) -> np.ndarray:
        r = self._r
        n = int(np.prod(self.shape)) if self.shape else 1
        if self._layout["class"] == 1:
            if self._layout["addr"] == UNDEF or n == 0:
                return np.zeros(self.shape, dtype=self.dtype)
            return np.frombuffer(r.buf, dtype=self.dtype, count=n, offset=self._layout["addr"]).reshape(self.shape).copy()
        out = np.zeros(self.shape, dtype=self.dtype)
        chunks = self.chunks
        for offs, addr, nbytes in r._walk_chunk_btree(self._layout["btree"], len(self.shape)):
            raw = r.buf[addr:addr + nbytes]
            if self._filters:
                raw = zlib.decompress(raw)
            block = np.frombuffer(raw, dtype=self.dtype).reshape(chunks)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, self.shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def __getitem__(self, key):
        return self.read()[key]


class H5Reader:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != SIGNATURE or b[8] != 0:
            raise ValueError("not a version-0 superblock HDF5 file")
        assert b[13] == 8 and b[14] == 8
        root_hdr = struct.unpack_from("<Q", b, 64)[0]
        msgs = self._header_messages(root_hdr)
        self.attrs = {}
        self.datasets: dict[str, H5Dataset] = {}
        for t, d in msgs:
            if t == 0x000C:
                k, v = self._attr(d)
                self.attrs[k] = v
            elif t == 0x0011:
                tree, heap = struct.unpack_from("<QQ", d, 0)
                for name, addr in self._group_entries(tree, heap):
                    self.datasets[name] = self._dataset(name, addr)

    def __getitem__(self, name):
        return self.datasets[name]

    def __contains__(self, name):
        return name in self.datasets

    def _header_messages(self, addr):
        b = self.buf
        ver, _, nmsg, _ref, size = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        pos, end, out = addr + 16, addr + 16 + size, []
        while pos < end and len(out) < nmsg:
            t, sz, _flags = struct.unpack_from("<HHB", b, pos)
            out.append((t, b[pos + 8:pos + 8 + sz]))
            pos += 8 + sz
        return out

    def _dtype(self, d):
        cls, b0, b1, _b2, size = struct.unpack_from("<BBBBI", d, 0)
        cls &= 0x0F
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 0:
            return np.dtype(("<i" if b0 & 8 else "<u") + str(size))
        if cls == 1:
            return np.dtype(f"<f{size}")
        raise TypeError(f"unsupported datatype class {cls}")

    def _space(self, d):
        ver, rank, flags = struct.unpack_from("<BBB", d, 0)
        assert ver == 1
        return tuple(struct.unpack_from("<Q", d, 8 + 8 * i)[0] for i in range(rank))

    def _attr(self, d):
        ver, _, nlen, dlen, slen = struct.unpack_from("<BBHHH", d, 0)
        assert ver == 1
        pos = 8
        name = d[pos:pos + nlen].rstrip(b"\0").decode()
        pos += nlen + (-nlen % 8)
        dt = self._dtype(d[pos:pos + dlen]); pos += dlen + (-dlen % 8)
        shape = self._space(d[pos:pos + slen]); pos += slen + (-slen % 8)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(d, dtype=dt, count=n, offset=pos).reshape(shape)
        if shape == ():
            v = arr[()]
            return name, (v.decode() if isinstance(v, bytes) else v.item())
        return name, arr.copy()

    def _group_entries(self, tree, heap):
        b = self.buf
        assert b[heap:heap + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        out = []

        def walk(node):
            assert b[node:node + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            assert ntype == 0
            pos = node + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8)[0]
                pos += 16
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for j in range(nsym):
                        noff, addr = struct.unpack_from("<QQ", b, child + 8 + 40 * j)
                        s = heap_data + noff
                        out.append((b[s:b.index(b"\0", s)].decode(), addr))
        walk(tree)
        return out

    def _dataset(self, name, addr):
        shape = dtype = None
        layout, filters, attrs = {}, [], {}
        for t, d in self._header_messages(addr):
            if t == 0x0001:
                shape = self._space(d)
            elif t == 0x0003:
                dtype = self._dtype(d)
            elif t == 0x0008:
                ver, cls = struct.unpack_from("<BB", d, 0)
                assert ver == 3
                if cls == 1:
                    a, sz = struct.unpack_from("<QQ", d, 2)
                    layout = {"class": 1, "addr": a, "size": sz}
                elif cls == 2:
                    nd = d[2]
                    bt = struct.unpack_from("<Q", d, 3)[0]
                    dims = struct.unpack_from("<" + "I" * nd, d, 11)
                    layout = {"class": 2, "btree": bt, "chunks": tuple(dims[:-1])}
                else:
                    raise TypeError("unsupported layout class")
            elif t == 0x000B:
                nfilt = d[1]
                pos = 8
                for _ in range(nfilt):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", d, pos)
                    pos += 8 + nlen + (-nlen % 8)
                    cd = struct.unpack_from("<" + "I" * ncd, d, pos)
                    pos += 4 * ncd + (4 if ncd % 2 else 0)
                    filters.append((fid, cd))
            elif t == 0x000C:
                k, v = self._attr(d)
                attrs[k] = v
        return H5Dataset(self, name, shape, dtype, layout, filters, attrs)

    def _walk_chunk_btree(self, node, ndims):
        b = self.buf
        key_size = 8 + 8 * (ndims + 1)
        out = []

        def walk(n):
            if n == UNDEF:
                return
            assert b[n:n + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", b, n + 4)
            assert ntype == 1
            pos = n + 24
            for _ in range(used):
                nbytes, _mask = struct.unpack_from("<II", b, pos)
                offs = struct.unpack_from("<" + "Q" * ndims, b, pos + 8)
                child = struct.unpack_from("<Q", b, pos + key_size)[0]
                pos += key_size + 8
                if level > 0:
                    walk(child)
                else:
                    out.append((offs, child, nbytes))
        walk(node)
        return out
