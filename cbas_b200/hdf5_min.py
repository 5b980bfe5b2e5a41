"""Minimal native HDF5 reader/writer for the CBAS embedding files (`<video>_cls.h5`).

The reference writes and reads these files with h5py (cbas.py:413-421,485-507).  h5py / libhdf5 are not
installed where this package is built, so the file format itself is implemented here for exactly the subset the
contract needs, following the HDF5 File Format Specification v1.1 (the "earliest" on-disk format, which is what
h5py emits by default and what every libhdf5 reads):

  * superblock v0, root group = v1 object header + symbol table (v1 group B-tree, SNOD, local heap)
  * one 2-D chunked, resizable dataset: dataspace v1 with max dims, IEEE float datatype (f2/f4/f8),
    fill-value v2, data layout v3 (chunked, v1 chunk B-tree, one or two levels), no filters
  * root-group attributes as scalar variable-length UTF-8 strings (global heap) - what `h5f.attrs[k] = "str"`
    produces, so `h5f.attrs['encoder_model_identifier'] != project_encoder` (startup_page.py:109-110) compares
    str with str - plus fixed-length strings and numeric scalars on the read side

`cbas_b200.store` prefers h5py whenever it is importable and uses this module otherwise.  The reader is checked
against a genuine libhdf5-written file that ships with scipy (tests/test_store.py); files from this writer are
checked by reading them back and by comparing their framing with that file.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class HDF5FormatError(RuntimeError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# =========================================================================================== reader
class _Reader:
    def __init__(self, path: str):
        self.path = path
        self.f = open(path, "rb")
        self.base = self._find_superblock()
        self._parse_superblock()

    def close(self):
        self.f.close()

    def _read(self, addr: int, n: int) -> bytes:
        self.f.seek(self.base + addr)
        b = self.f.read(n)
        if len(b) != n:
            raise HDF5FormatError(f"{self.path}: truncated file (wanted {n} bytes at {addr})")
        return b

    def _find_superblock(self) -> int:
        off = 0
        size = os.path.getsize(self.path)
        while off < size:
            self.f.seek(off)
            if self.f.read(8) == SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise HDF5FormatError(f"{self.path}: not an HDF5 file")

    def _parse_superblock(self):
        b = self._read(0, 96)
        ver = b[8]
        if ver not in (0, 1):
            raise HDF5FormatError(f"{self.path}: superblock version {ver} is not supported by the native reader "
                                  "(file written with libver='latest'?) - install h5py")
        if b[13] != 8 or b[14] != 8:
            raise HDF5FormatError("only 8-byte offsets/lengths are supported")
        off = 24 if ver == 0 else 28
        hdr = self._read(0, off + 32 + 40)
        base_addr, _free, eof, _drv = struct.unpack_from("<QQQQ", hdr, off)
        # addresses in the file are relative to the base address; the superblock sits at it
        if base_addr not in (0, self.base):
            raise HDF5FormatError("unexpected base address")
        self.eof = eof
        ste = hdr[off + 32:off + 72]
        self.root_header = struct.unpack_from("<Q", ste, 8)[0]

    # ---- object headers ------------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        head = self._read(addr, 16)
        if head[0] != 1:
            raise HDF5FormatError(f"object header version {head[0]} not supported (v2 'OHDR' needs h5py)")
        nmsg, _ref, size = struct.unpack_from("<HII", head, 2)
        blocks = [(addr + 16, size)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg:
            a, n = blocks.pop(0)
            data = self._read(a, n)
            p = 0
            while p + 8 <= n and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", data, p)
                body = data[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:  # continuation
                    ca, cn = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cn))
                out.append((mtype, body))
        return out

    # ---- groups --------------------------------------------------------------------------------------
    def links(self, header_addr: int) -> Dict[str, int]:
        for mtype, body in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                return self._group_entries(btree, heap)
        raise HDF5FormatError("object is not an (old-style) group")

    def _heap_data(self, heap_addr: int) -> Tuple[int, int]:
        h = self._read(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise HDF5FormatError("bad local heap signature")
        size, _free, data_addr = struct.unpack_from("<QQQ", h, 8)
        return data_addr, size

    def _heap_name(self, heap: Tuple[int, int], off: int) -> str:
        data_addr, size = heap
        raw = self._read(data_addr + off, min(256, size - off))
        return raw.split(b"\x00", 1)[0].decode("utf-8")

    def _group_entries(self, btree_addr: int, heap_addr: int) -> Dict[str, int]:
        heap = self._heap_data(heap_addr)
        out: Dict[str, int] = {}

        def walk(addr: int):
            h = self._read(addr, 24)
            if h[:4] == b"SNOD":
                n = struct.unpack_from("<H", h, 6)[0]
                ents = self._read(addr + 8, n * 40)
                for i in range(n):
                    name_off, obj = struct.unpack_from("<QQ", ents, i * 40)
                    out[self._heap_name(heap, name_off)] = obj
                return
            if h[:4] != b"TREE" or h[4] != 0:
                raise HDF5FormatError("bad group B-tree node")
            used = struct.unpack_from("<H", h, 6)[0]
            body = self._read(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                walk(struct.unpack_from("<Q", body, 8 + i * 16)[0])

        walk(btree_addr)
        return out

    # ---- datatypes / dataspaces ----------------------------------------------------------------------
    @staticmethod
    def parse_datatype(body: bytes):
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size = struct.unpack_from("<I", body, 4)[0]
        if cls == 1:  # floating point
            if bits & 1:
                raise HDF5FormatError("big-endian floats not supported")
            return ("float", np.dtype(f"<f{size}"))
        if cls == 0:  # fixed point
            signed = bool(bits & 0x08)
            return ("int", np.dtype(("<i" if signed else "<u") + str(size)))
        if cls == 3:  # fixed-length string
            return ("string", size)
        if cls == 9:  # variable length
            if (bits & 0x0F) == 1:
                return ("vlen_string", size)
            raise HDF5FormatError("variable-length sequences are not supported")
        raise HDF5FormatError(f"datatype class {cls} (version {ver}) not supported")

    @staticmethod
    def parse_dataspace(body: bytes):
        ver = body[0]
        rank, flags = body[1], body[2]
        p = 8 if ver == 1 else 4
        dims = list(struct.unpack_from(f"<{rank}Q", body, p)) if rank else []
        p += 8 * rank
        maxdims = list(struct.unpack_from(f"<{rank}Q", body, p)) if (flags & 1 and rank) else list(dims)
        return dims, maxdims

    # ---- attributes ----------------------------------------------------------------------------------
    def attributes(self, header_addr: int) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for mtype, body in self.messages(header_addr):
            if mtype != 0x0C:
                continue
            ver = body[0]
            name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", body, 2)
            p = 8
            if ver == 3:
                p = 9
            pad = _pad8 if ver == 1 else (lambda n: n)
            name = body[p:p + name_sz].split(b"\x00", 1)[0].decode("utf-8")
            p += pad(name_sz)
            dt = self.parse_datatype(body[p:p + dt_sz])
            p += pad(dt_sz)
            dims, _ = self.parse_dataspace(body[p:p + ds_sz])
            p += pad(ds_sz)
            data = body[p:]
            count = int(np.prod(dims)) if dims else 1
            kind = dt[0]
            if kind == "vlen_string":
                vals = []
                for i in range(count):
                    ln, col, idx = struct.unpack_from("<IQI", data, i * 16)
                    vals.append(self._global_heap_object(col, idx)[:ln].decode("utf-8"))
                out[name] = vals[0] if not dims else vals
            elif kind == "string":
                n = dt[1]
                vals = [data[i * n:(i + 1) * n].split(b"\x00", 1)[0] for i in range(count)]
                out[name] = np.bytes_(vals[0]) if not dims else np.array(vals)
            else:
                arr = np.frombuffer(data[:count * dt[1].itemsize], dtype=dt[1])
                out[name] = arr[0] if not dims else arr.reshape(dims)
        return out

    def _global_heap_object(self, col_addr: int, index: int) -> bytes:
        h = self._read(col_addr, 16)
        if h[:4] != b"GCOL":
            raise HDF5FormatError("bad global heap collection")
        size = struct.unpack_from("<Q", h, 8)[0]
        data = self._read(col_addr, size)
        p = 16
        while p + 16 <= size:
            idx, _ref, _r, osz = struct.unpack_from("<HHIQ", data, p)
            if idx == 0:
                break
            if idx == index:
                return data[p + 16:p + 16 + osz]
            p += 16 + _pad8(osz)
        raise HDF5FormatError("global heap object not found")


class Dataset:
    """Read-only 2-D (or N-D) dataset view: `.shape`, `.dtype`, `len()`, slicing on the first axis."""

    def __init__(self, rd: _Reader, header_addr: int):
        self._rd = rd
        self.maxshape = None
        self.chunks = None
        layout = None
        for mtype, body in rd.messages(header_addr):
            if mtype == 0x01:
                dims, maxd = rd.parse_dataspace(body)
                self.shape = tuple(dims)
                self.maxshape = tuple(None if m == UNDEF else m for m in maxd)
            elif mtype == 0x03:
                kind, dt = rd.parse_datatype(body)
                if kind not in ("float", "int"):
                    raise HDF5FormatError("only numeric datasets are supported")
                self.dtype = dt
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                raise HDF5FormatError("filtered (compressed) datasets are not supported by the native reader")
        if layout is None:
            raise HDF5FormatError("dataset has no data layout message")
        if layout[0] in (1, 2):  # pre-1.6.3 layout message (old files): version, rank, class, 5 reserved, address, dims
            rank, self._class = layout[1], layout[2]
            p = 8
            addr = UNDEF
            if self._class != 0:
                addr = struct.unpack_from("<Q", layout, p)[0]
                p += 8
            cd = struct.unpack_from(f"<{rank}I", layout, p)
            p += 4 * rank
            if self._class == 1:
                self._addr, self._size = addr, int(np.prod(self.shape)) * self.dtype.itemsize
            elif self._class == 2:
                self._btree, self.chunks, self._index = addr, tuple(cd[:-1]), None
            else:
                sz = struct.unpack_from("<I", layout, p)[0]
                self._compact = layout[p + 4:p + 4 + sz]
            return
        if layout[0] != 3:
            raise HDF5FormatError(f"data layout message version {layout[0]} is not supported (needs h5py)")
        self._class = layout[1]
        if self._class == 1:  # contiguous
            self._addr, self._size = struct.unpack_from("<QQ", layout, 2)
        elif self._class == 2:  # chunked
            rank1 = layout[2]
            self._btree = struct.unpack_from("<Q", layout, 3)[0]
            cd = struct.unpack_from(f"<{rank1}I", layout, 11)
            self.chunks = tuple(cd[:-1])
            self._index: Optional[Dict[Tuple[int, ...], Tuple[int, int]]] = None
        elif self._class == 0:  # compact
            sz = struct.unpack_from("<H", layout, 2)[0]
            self._compact = layout[4:4 + sz]
        else:
            raise HDF5FormatError("unknown layout class")

    def __len__(self):
        return self.shape[0]

    def _chunk_index(self):
        if self._index is None:
            idx: Dict[Tuple[int, ...], Tuple[int, int]] = {}
            rank1 = len(self.chunks) + 1
            key_sz = 8 + 8 * rank1

            def walk(addr):
                if addr == UNDEF:
                    return
                h = self._rd._read(addr, 24)
                if h[:4] != b"TREE" or h[4] != 1:
                    raise HDF5FormatError("bad chunk B-tree node")
                level, used = h[5], struct.unpack_from("<H", h, 6)[0]
                body = self._rd._read(addr + 24, used * (key_sz + 8) + key_sz)
                for i in range(used):
                    p = i * (key_sz + 8)
                    csize, _mask = struct.unpack_from("<II", body, p)
                    offs = struct.unpack_from(f"<{rank1}Q", body, p + 8)
                    child = struct.unpack_from("<Q", body, p + key_sz)[0]
                    if level == 0:
                        idx[tuple(offs[:-1])] = (child, csize)
                    else:
                        walk(child)

            walk(self._btree)
            self._index = idx
        return self._index

    def read_rows(self, start: int, stop: int) -> np.ndarray:
        start, stop = max(0, start), min(self.shape[0], stop)
        tail = self.shape[1:]
        n = max(0, stop - start)
        out = np.zeros((n,) + tuple(tail), self.dtype)
        if n == 0:
            return out
        row_items = int(np.prod(tail)) if tail else 1
        if self._class == 1:
            if self._addr != UNDEF:
                raw = self._rd._read(self._addr + start * row_items * self.dtype.itemsize,
                                     n * row_items * self.dtype.itemsize)
                out[...] = np.frombuffer(raw, self.dtype).reshape(out.shape)
            return out
        if self._class == 0:
            full = np.frombuffer(self._compact, self.dtype).reshape(self.shape)
            return full[start:stop].copy()
        if any(c != s for c, s in zip(self.chunks[1:], self.shape[1:])):
            raise HDF5FormatError("chunks must span the trailing dimensions")
        cr = self.chunks[0]
        index = self._chunk_index()
        for c0 in range((start // cr) * cr, stop, cr):
            ent = index.get((c0,) + (0,) * (len(self.shape) - 1))
            if ent is None:
                continue  # unallocated chunk reads as the fill value (zero)
            lo, hi = max(start, c0), min(stop, c0 + cr)
            raw = self._rd._read(ent[0] + (lo - c0) * row_items * self.dtype.itemsize,
                                 (hi - lo) * row_items * self.dtype.itemsize)
            out[lo - start:hi - start] = np.frombuffer(raw, self.dtype).reshape((hi - lo,) + tuple(tail))
        return out

    def __getitem__(self, key):
        if isinstance(key, slice):
            start, stop, step = key.indices(self.shape[0])
            if step != 1:
                return self.read_rows(start, stop)[::step]
            return self.read_rows(start, stop)
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            return self.read_rows(k, k + 1)[0]
        if key is Ellipsis or key == ():
            return self.read_rows(0, self.shape[0])
        raise TypeError("native HDF5 datasets support integer and slice indexing on the first axis")


class File:
    """Read-only file: `f["cls"]`, `"cls" in f`, `f.attrs`, context manager - the h5py surface infer_file and the
    project loader use (cbas.py:485-487, startup_page.py:102-110)."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("hdf5_min.File is read-only; use hdf5_min.Writer to create files")
        self._rd = _Reader(path)
        self._links = self._rd.links(self._rd.root_header)
        self.attrs = self._rd.attributes(self._rd.root_header)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def close(self):
        self._rd.close()

    def keys(self):
        return self._links.keys()

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name) -> Dataset:
        if name not in self._links:
            raise KeyError(name)
        return Dataset(self._rd, self._links[name])


# =========================================================================================== writer
def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\x00" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(msgs: List[bytes]) -> bytes:
    data = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(data)) + data


def _float_datatype(dt: np.dtype) -> bytes:
    spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
    sign, eloc, esz, mloc, msz, bias = spec
    return (struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize) +
            struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esz, mloc, msz, bias))


class Writer:
    """Append-only writer of one chunked dataset `[N, width]` plus root string attributes.

    Chunk data is streamed to disk as rows arrive (whole chunks, the last one zero-padded as HDF5 stores it);
    the metadata (object headers, group B-tree, heaps, chunk index) is laid down behind the data by close(),
    which then writes the superblock - a reader never sees a half-indexed file, and the caller publishes the
    finished file with an atomic rename exactly like the reference (cbas.py:410,442)."""

    DATA_START = 2048

    def __init__(self, path: str, name: str, width: int, dtype="f2", chunk_rows: int = 8192,
                 attrs: Optional[Dict[str, str]] = None):
        self.path, self.name, self.width = path, name, int(width)
        self.dtype = np.dtype(dtype).newbyteorder("<")
        if self.dtype.kind != "f":
            raise ValueError("only float datasets are written")
        self.chunk_rows = int(chunk_rows)
        self.attrs = dict(attrs or {})
        self.rows = 0
        self._chunk_bytes = self.chunk_rows * self.width * self.dtype.itemsize
        self._f = open(path, "wb")
        self._f.write(b"\x00" * self.DATA_START)
        self._closed = False

    # rows land at DATA_START + row * rowbytes: chunks are contiguous and in order, so appending is a plain write
    def append(self, block: np.ndarray) -> None:
        block = np.ascontiguousarray(block, dtype=self.dtype)
        if block.ndim != 2 or block.shape[1] != self.width:
            raise ValueError(f"expected [n,{self.width}] rows")
        self._f.seek(self.DATA_START + self.rows * self.width * self.dtype.itemsize)
        self._f.write(block.tobytes())
        self.rows += block.shape[0]

    def flush(self) -> None:
        self._f.flush()

    def abort(self) -> None:
        if not self._closed:
            self._f.close()
            self._closed = True

    def close(self) -> None:
        if self._closed:
            return
        f = self._f
        n_chunks = (self.rows + self.chunk_rows - 1) // self.chunk_rows
        end = self.DATA_START + n_chunks * self._chunk_bytes
        f.seek(self.DATA_START + self.rows * self.width * self.dtype.itemsize)
        f.write(b"\x00" * (end - f.tell()))  # zero-pad the last chunk

        pos = [end]

        def alloc(n: int) -> int:
            a = pos[0]
            pos[0] = _pad8(a + n)
            return a

        blobs: List[Tuple[int, bytes]] = []
        # ---- chunk B-tree (type 1), leaves of <= 64 entries under an optional root
        rank1 = 3
        key_sz = 8 + 8 * rank1
        node_sz = 24 + 65 * key_sz + 64 * 8

        def chunk_key(row0: int, size: int) -> bytes:
            return struct.pack("<II3Q", size, 0, row0, 0, 0)

        def node(level: int, entries: List[Tuple[int, int]], last_row: int) -> int:
            # entries: (first row of the subtree, child address)
            addr = alloc(node_sz)
            body = b""
            for row0, child in entries:
                body += chunk_key(row0, self._chunk_bytes if level == 0 else 0) + struct.pack("<Q", child)
            body += chunk_key(last_row, 0)
            raw = b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF) + body
            blobs.append((addr, raw + b"\x00" * (node_sz - len(raw))))
            return addr

        chunks = [(i * self.chunk_rows, self.DATA_START + i * self._chunk_bytes) for i in range(n_chunks)]
        if n_chunks == 0:
            btree = UNDEF
        elif n_chunks <= 64:
            btree = node(0, chunks, n_chunks * self.chunk_rows)
        else:
            if n_chunks > 64 * 64:
                raise HDF5FormatError("dataset too large for the two-level chunk index of the native writer")
            leaves = []
            for i in range(0, n_chunks, 64):
                part = chunks[i:i + 64]
                leaves.append((part[0][0], node(0, part, part[-1][0] + self.chunk_rows)))
            btree = node(1, leaves, n_chunks * self.chunk_rows)

        # ---- dataset object header
        dataspace = struct.pack("<BBB5x", 1, 2, 1) + struct.pack("<4Q", self.rows, self.width, UNDEF, self.width)
        layout = struct.pack("<BBB", 3, 2, rank1) + struct.pack("<Q", btree) + struct.pack(
            "<3I", self.chunk_rows, self.width, self.dtype.itemsize)
        # v2: incremental allocation (chunked), write the fill value if set, defined with size 0 = library default (0)
        fill = struct.pack("<BBBBI", 2, 3, 2, 1, 0)
        ds_header = _object_header([_msg(0x01, dataspace), _msg(0x03, _float_datatype(self.dtype), 1),
                                    _msg(0x05, fill, 1), _msg(0x08, layout)])
        ds_addr = alloc(len(ds_header))
        blobs.append((ds_addr, ds_header))

        # ---- root group: local heap, symbol-table node, group B-tree
        heap_data = b"\x00" * 8 + self.name.encode("utf-8") + b"\x00"
        name_off = 8
        heap_data += b"\x00" * (_pad8(len(heap_data)) - len(heap_data))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 64) + b"\x00" * 48  # one free block to the end of the segment
        heap_data_addr = alloc(len(heap_data))
        heap_addr = alloc(32)
        blobs.append((heap_data_addr, heap_data))
        blobs.append((heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)))
        snod = b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", name_off, ds_addr, 0, 0)
        snod += b"\x00" * (8 + 2 * 4 * 40 - len(snod))
        snod_addr = alloc(len(snod))
        blobs.append((snod_addr, snod))
        gnode = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off)
        gnode += b"\x00" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(gnode))
        gtree_addr = alloc(len(gnode))
        blobs.append((gtree_addr, gnode))

        # ---- attributes: scalar variable-length UTF-8 strings, values in one global heap collection
        attr_msgs: List[bytes] = []
        if self.attrs:
            objs = b""
            ids = {}
            for i, (k, v) in enumerate(self.attrs.items(), start=1):
                raw = str(v).encode("utf-8")
                ids[k] = (i, len(raw))
                objs += struct.pack("<HHIQ", i, 1, 0, len(raw)) + raw + b"\x00" * (_pad8(len(raw)) - len(raw))
            col_size = max(4096, _pad8(16 + len(objs) + 16))
            free = col_size - 16 - len(objs)
            col = b"GCOL" + struct.pack("<B3xQ", 1, col_size) + objs + struct.pack("<HHIQ", 0, 0, 0, free)
            col += b"\x00" * (col_size - len(col))
            col_addr = alloc(col_size)
            blobs.append((col_addr, col))
            # class 9 (variable length) v1: type = string, null-terminated, UTF-8; 16-byte descriptor; the base type
            # is the 1-byte unsigned integer libhdf5 gives every variable-length string (H5T_NATIVE_UCHAR)
            vlen_dt = (struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) +
                       struct.pack("<BBBBIHH", 0x10, 0x00, 0x00, 0x00, 1, 0, 8))
            scalar_ds = struct.pack("<BBB5x", 1, 0, 0)
            for k, (idx, ln) in ids.items():
                name = k.encode("utf-8") + b"\x00"
                body = struct.pack("<BxHHH", 1, len(name), len(vlen_dt), len(scalar_ds))
                body += name + b"\x00" * (_pad8(len(name)) - len(name))
                body += vlen_dt + b"\x00" * (_pad8(len(vlen_dt)) - len(vlen_dt))
                body += scalar_ds + b"\x00" * (_pad8(len(scalar_ds)) - len(scalar_ds))
                body += struct.pack("<IQI", ln, col_addr, idx)
                attr_msgs.append(_msg(0x0C, body))
        root_header = _object_header([_msg(0x11, struct.pack("<QQ", gtree_addr, heap_addr))] + attr_msgs)
        root_addr = alloc(len(root_header))
        blobs.append((root_addr, root_header))

        eof = pos[0]
        for addr, raw in blobs:
            f.seek(addr)
            f.write(raw)
        f.seek(eof - 1)
        f.write(b"\x00")
        # ---- superblock v0 last: until it is in place the file does not parse as HDF5 at all
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", gtree_addr, heap_addr)
        f.seek(0)
        f.write(sb)
        f.flush()
        os.fsync(f.fileno())
        f.close()
        self._closed = True
