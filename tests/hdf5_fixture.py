"""TEST-ONLY writer for the HDF5 subset rvdb_reader.py reads (superblock 0, version-1 object headers with a
continuation block, symbol-table groups, layout v3: compact / contiguous / chunked with a two-level version-1 chunk
B-tree, filter pipeline v1 with shuffle + LZF) and a small LZF compressor.

Written from the published HDF5 File Format Specification and the liblzf stream format -- NOT by h5py (absent from
this image): a reader that passes against this writer is self-consistent, not proven against real .rvdb files.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def lzf_compress(data: bytes) -> bytes:
    """Greedy LZF encoder (literal runs <= 32 bytes, back references of 3 .. 264 bytes within 8 KB)."""
    n, out, i, lit_start = len(data), bytearray(), 0, 0
    table = {}

    def flush(upto):
        s = lit_start
        while s < upto:
            run = min(32, upto - s)
            out.append(run - 1)
            out.extend(data[s:s + run])
            s += run

    while i + 2 < n:
        key = data[i:i + 3]
        ref = table.get(key)
        table[key] = i
        if ref is not None and 0 < i - ref <= 8192:
            length = 3
            while i + length < n and length < 264 and data[ref + length] == data[i + length]:
                length += 1
            flush(i)
            off, l2 = i - ref - 1, length - 2
            if l2 < 7:
                out.append((l2 << 5) | (off >> 8))
            else:
                out.append((7 << 5) | (off >> 8))
                out.append(l2 - 7)
            out.append(off & 0xFF)
            i += length
            lit_start = i
        else:
            i += 1
    flush(n)
    return bytes(out)


def shuffle(raw: bytes, elem: int) -> bytes:
    a = np.frombuffer(raw, np.uint8)
    n = len(a) // elem
    return a[:n * elem].reshape(n, elem).T.tobytes() + a[n * elem:].tobytes()


class Writer:
    def __init__(self):
        self.buf = bytearray(96)                                  # superblock + root symbol table entry, filled at the end

    def _alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf.extend(data)
        return addr

    @staticmethod
    def _msg(mtype, body: bytes) -> bytes:
        body = body + b"\x00" * (-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _object_header(self, msgs, continuation=None) -> int:
        """msgs in the first block; `continuation` (a list of messages) goes to a separately allocated block."""
        n = len(msgs) + (1 + len(continuation) if continuation else 0)
        if continuation:
            cont = b"".join(continuation)
            caddr = self._alloc(cont)
            msgs = msgs + [self._msg(0x0010, struct.pack("<QQ", caddr, len(cont)))]
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, n, 1, len(body)) + body)

    def group(self, links: dict) -> int:
        """links: {name: object header address}; one SNOD (<= 8 entries)."""
        assert len(links) <= 8
        heap_data, offs = bytearray(b"\x00" * 8), {}
        for name in sorted(links):
            offs[name] = len(heap_data)
            heap_data.extend(name.encode() + b"\x00")
            heap_data.extend(b"\x00" * (-len(heap_data) % 8))
        data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(links))
        for name in sorted(links):
            snod += struct.pack("<QQII16x", offs[name], links[name], 0, 0)
        snod_addr = self._alloc(snod + b"\x00" * (40 * (8 - len(links))))
        last = offs[sorted(links)[-1]] if links else 0
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, last)
        btree = self._alloc(tree)
        return self._object_header([self._msg(0x0011, struct.pack("<QQ", btree, heap))])

    @staticmethod
    def _dtype_msg(dt: np.dtype) -> bytes:
        if dt.kind == "f":
            props = struct.pack("<HHBBBBI", 0, dt.itemsize * 8, 23, 8, 0, 23, 127) if dt.itemsize == 4 else \
                struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            return struct.pack("<BBBBI", 0x11, 0x20, 0x1F if dt.itemsize == 4 else 0x3F, 0, dt.itemsize) + props
        signed = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, signed, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)

    def dataset(self, arr: np.ndarray, chunks=None, lzf=False, do_shuffle=False, compact=False, skip_lzf_on=()) -> int:
        arr = np.ascontiguousarray(arr)
        space = self._msg(0x0001, struct.pack("<BBB5x", 1, arr.ndim, 0) + struct.pack(f"<{arr.ndim}Q", *arr.shape))
        dtype = self._msg(0x0003, self._dtype_msg(arr.dtype))
        extra = []
        if compact:
            raw = arr.tobytes()
            layout = self._msg(0x0008, struct.pack("<BBH", 3, 0, len(raw)) + raw)
        elif chunks is None:
            addr = self._alloc(arr.tobytes())
            layout = self._msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes))
        else:
            entries = []
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
            for idx, start in enumerate(np.array(np.meshgrid(*grid, indexing="ij")).reshape(arr.ndim, -1).T):
                block = np.zeros(chunks, arr.dtype)
                sl = tuple(slice(s, min(s + c, d)) for s, c, d in zip(start, chunks, arr.shape))
                block[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
                raw, mask = block.tobytes(), 0
                if do_shuffle:
                    raw = shuffle(raw, arr.dtype.itemsize)
                if lzf:
                    comp = lzf_compress(raw)
                    if idx in skip_lzf_on or len(comp) >= len(raw):
                        mask |= 1 << (1 if do_shuffle else 0)      # the optional filter was skipped for this chunk
                    else:
                        raw = comp
                entries.append((len(raw), mask, tuple(int(v) for v in start) + (0,), self._alloc(raw)))
            rank1 = arr.ndim + 1

            def node(level, items, children):
                body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
                for (size, mask, offs, _), child in zip(items, children):
                    body += struct.pack("<II", size, mask) + struct.pack(f"<{rank1}Q", *offs) + struct.pack("<Q", child)
                body += struct.pack("<II", 0, 0) + struct.pack(f"<{rank1}Q", *(tuple(arr.shape) + (0,)))   # final key
                return self._alloc(body)

            leaves, firsts = [], []
            for s in range(0, len(entries), 3):                   # tiny fan-out: forces a second tree level
                part = entries[s:s + 3]
                leaves.append(node(0, part, [e[3] for e in part]))
                firsts.append(part[0])
            root = leaves[0] if len(leaves) == 1 else node(1, firsts, leaves)
            layout = self._msg(0x0008, struct.pack("<BBBQ", 3, 2, rank1, root) +
                               struct.pack(f"<{rank1}I", *(tuple(chunks) + (arr.dtype.itemsize,))))
            filt = []
            if do_shuffle:
                filt.append(struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\x00" + struct.pack("<I", arr.dtype.itemsize) + b"\x00" * 4)
            if lzf:
                filt.append(struct.pack("<HHHH", 32000, 8, 1, 3) + b"lzf\x00\x00\x00\x00\x00" +
                            struct.pack("<III", 4, 261, int(np.prod(chunks)) * arr.dtype.itemsize) + b"\x00" * 4)
            if filt:
                extra.append(self._msg(0x000B, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
        return self._object_header([space, dtype, layout], continuation=extra or None)

    def finish(self, root_header: int) -> bytes:
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII16x", 0, root_header, 0, 0)
        assert len(sb) == 96, len(sb)
        self.buf[:96] = sb
        return bytes(self.buf)
