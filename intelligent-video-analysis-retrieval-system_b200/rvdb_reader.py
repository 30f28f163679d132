"""Reader for the reference's ``.rvdb`` container -- no h5py, no lz4.

A ``.rvdb`` file is an HDF5 file written by h5py with its default settings (``h5py.File(path, 'w')``:
unified_index.py:179), holding among others

* ``vectors/embeddings``  float32 [N, D], chunked, ``shuffle`` + ``lzf`` filters  (unified_index.py:943-948, 1557-1562)
* ``metadata/data``       uint8: one LZ4 frame around the JSON metadata list      (unified_index.py:950-956)
* ``faiss_index``         uint8, chunked + lzf: the bytes of ``faiss.write_index``  (unified_index.py:1811-1832)

which the reference reads back in ``_setup_memory_maps`` (unified_index.py:1175-1234; 29 s for 851 284 frames,
logs/system_20250828.log:26).  Neither h5py nor lz4 is installed in this image and there is no network, so this
module restates the published formats for exactly the subset h5py's defaults produce:

* HDF5 File Format Specification: superblock version 0 / 1, version-1 object headers (+ continuation blocks),
  old-style groups (symbol-table message -> version-1 B-tree of group nodes -> ``SNOD`` symbol nodes -> local heap),
  dataspace v1 / v2, fixed-point / floating-point datatypes, data layout v3 (compact, contiguous, chunked with a
  version-1 chunk B-tree), filter pipeline v1 / v2 with shuffle (2), deflate (1) and LZF (32000);
* LZ4 Frame Format v1.6 (magic ``0x184D2204``, FLG / BD, optional content size and dictionary id, linked or
  independent blocks, uncompressed blocks, block / content checksums skipped, skippable frames).

The bulk bytes (LZF chunks, LZ4 blocks, un-shuffle) are decoded by the C entry points in ``csrc/codecs.cu``.

VALIDATION STATUS (also in DESIGN.md): the LZ4 frame path is checked against frames produced by a real LZ4 library
(pyarrow's codec).  The HDF5 path could only be checked against a file assembled by this repository's own test
writer (tests/hdf5_fixture.py), because no HDF5 writer exists in the image and the reference ships no ``.rvdb``:
it is UNVALIDATED against a file written by h5py.  Files written with ``libver='latest'`` (superblock 2/3, link
messages, fractal heaps, version-2 B-trees) are rejected with a clear error rather than mis-read.
"""
from __future__ import annotations

import ctypes as C
import json
import mmap
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from . import _native as nat

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


class RvdbFormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------- LZ4 frame
def lz4_frame_decompress(data) -> bytes:
    """``lz4.frame.decompress`` for one or more concatenated frames."""
    src = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    buf = src.tobytes() if src.size < (1 << 16) else None       # header parsing works on bytes; big inputs stay as views
    view = memoryview(src)
    n, pos, out_parts = len(view), 0, []
    while pos < n:
        if n - pos < 4:
            raise RvdbFormatError("LZ4: trailing bytes after the last frame")
        magic = struct.unpack_from("<I", view, pos)[0]
        pos += 4
        if 0x184D2A50 <= magic <= 0x184D2A5F:                   # skippable frame
            size = struct.unpack_from("<I", view, pos)[0]
            pos += 4 + size
            continue
        if magic != 0x184D2204:
            raise RvdbFormatError(f"LZ4: bad frame magic 0x{magic:08x}")
        flg, bd = view[pos], view[pos + 1]
        pos += 2
        if (flg >> 6) != 1:
            raise RvdbFormatError("LZ4: unsupported frame version")
        block_checksum, has_size, content_checksum, has_dict = (flg >> 4) & 1, (flg >> 3) & 1, (flg >> 2) & 1, flg & 1
        block_max = {4: 1 << 16, 5: 1 << 18, 6: 1 << 20, 7: 1 << 22}.get((bd >> 4) & 7)
        if block_max is None:
            raise RvdbFormatError("LZ4: bad block-size code")
        content_size = None
        if has_size:
            content_size = struct.unpack_from("<Q", view, pos)[0]
            pos += 8
        if has_dict:
            pos += 4
        pos += 1                                                # header checksum
        # pass 1: the block table (sizes are needed before the output can be allocated when no content size is stored)
        blocks, p = [], pos
        while True:
            word = struct.unpack_from("<I", view, p)[0]
            p += 4
            if word == 0:
                break
            size, raw = word & 0x7FFFFFFF, bool(word >> 31)
            if p + size > n:
                raise RvdbFormatError("LZ4: truncated block")
            blocks.append((p, size, raw))
            p += size + (4 if block_checksum else 0)
        end = p + (4 if content_checksum else 0)
        cap = content_size if content_size is not None else len(blocks) * block_max
        out = np.empty(max(cap, 1), np.uint8)
        produced = 0
        for off, size, raw in blocks:
            if raw:
                if produced + size > cap:
                    raise RvdbFormatError("LZ4: output larger than announced")
                out[produced:produced + size] = src[off:off + size]
                got = size
            else:
                got = nat.lib.ivr_lz4_block_decompress(src[off:].ctypes.data, size, out.ctypes.data, produced, cap)
                if got < 0:
                    raise RvdbFormatError("LZ4: malformed block")
            produced += got
        if content_size is not None and produced != content_size:
            raise RvdbFormatError(f"LZ4: frame decoded to {produced} bytes, header says {content_size}")
        out_parts.append(out[:produced].tobytes())
        pos = end
    del buf
    return b"".join(out_parts)


# ------------------------------------------------------------------------------------- HDF5 subset
class _Dataset:
    def __init__(self, f: "Hdf5File", name: str, shape, dtype, layout, filters):
        self.file, self.name, self.shape, self.dtype = f, name, tuple(shape), np.dtype(dtype)
        self.layout, self.filters = layout, filters

    def __repr__(self):
        return f"<rvdb dataset {self.name!r} {self.shape} {self.dtype} {self.layout['class']}>"

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    # -- whole dataset ----------------------------------------------------------------------------
    def read(self) -> np.ndarray:
        cls = self.layout["class"]
        if cls == "compact":
            return np.frombuffer(self.layout["data"], self.dtype, count=int(np.prod(self.shape))).reshape(self.shape).copy()
        if cls == "contiguous":
            if self.layout["address"] == UNDEF:
                return np.zeros(self.shape, self.dtype)
            return np.frombuffer(self.file.buf, self.dtype, count=int(np.prod(self.shape, dtype=np.int64)),
                                 offset=self.layout["address"]).reshape(self.shape).copy()
        out = np.zeros(self.shape, self.dtype)
        for start, block in self.iter_chunks():
            sl = tuple(slice(s, min(s + c, d)) for s, c, d in zip(start, block.shape, self.shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    # -- chunk by chunk (the loader streams row blocks to the device) ---------------------------
    def iter_chunks(self) -> Iterator[Tuple[Tuple[int, ...], np.ndarray]]:
        if self.layout["class"] != "chunked":
            yield (0,) * len(self.shape), self.read()
            return
        cdims = self.layout["chunk"]
        raw_bytes = int(np.prod(cdims, dtype=np.int64)) * self.dtype.itemsize
        scratch = [np.empty(raw_bytes, np.uint8), np.empty(raw_bytes, np.uint8)]
        for offsets, address, size, mask in self.file._chunk_entries(self.layout["btree"], len(cdims) + 1):
            data = np.frombuffer(self.file.buf, np.uint8, count=size, offset=address)
            turn = 0
            for i in range(len(self.filters) - 1, -1, -1):      # filters are undone in reverse order
                fid, cd = self.filters[i]
                if mask & (1 << i):
                    continue                                    # this filter was skipped for this chunk
                dst = scratch[turn]
                turn ^= 1
                if fid == 32000:                                # LZF (h5py): cd_values[2] = uncompressed chunk size
                    got = nat.lib.ivr_lzf_decompress(data.ctypes.data, data.size, dst.ctypes.data, raw_bytes)
                    if got < 0:
                        raise RvdbFormatError(f"{self.name}: malformed LZF chunk at {offsets}")
                    data = dst[:got]
                elif fid == 2:                                  # shuffle
                    nat.check(nat.lib.ivr_unshuffle(data.ctypes.data, data.size, self.dtype.itemsize, dst.ctypes.data))
                    data = dst[:data.size]
                elif fid == 1:                                  # deflate
                    data = np.frombuffer(zlib.decompress(data.tobytes()), np.uint8)
                else:
                    raise RvdbFormatError(f"{self.name}: unsupported HDF5 filter id {fid}")
            if data.size != raw_bytes:
                raise RvdbFormatError(f"{self.name}: chunk at {offsets} decoded to {data.size} bytes, expected {raw_bytes}")
            yield tuple(offsets[:-1]), data.view(self.dtype).reshape(cdims).copy()   # the last offset is the element's (0)

    def iter_row_blocks(self) -> Iterator[Tuple[int, np.ndarray]]:
        """(first_row, rows [m, D]) in ascending row order for a 2-D dataset whose chunks span all columns or not."""
        if len(self.shape) != 2:
            raise RvdbFormatError(f"{self.name}: iter_row_blocks needs a 2-D dataset")
        n, d = self.shape
        if self.layout["class"] != "chunked":
            yield 0, self.read()
            return
        crow, ccol = self.layout["chunk"]
        pending: Dict[int, np.ndarray] = {}
        filled: Dict[int, int] = {}
        per_row_block = (d + ccol - 1) // ccol
        for (r0, c0), block in self.iter_chunks():
            if per_row_block == 1:
                yield r0, block[:min(crow, n - r0), :d]
                continue
            if r0 not in pending:
                pending[r0] = np.empty((min(crow, n - r0), d), self.dtype)
                filled[r0] = 0
            w = min(ccol, d - c0)
            pending[r0][:, c0:c0 + w] = block[:pending[r0].shape[0], :w]
            filled[r0] += 1
            if filled[r0] == per_row_block:
                yield r0, pending.pop(r0)
        if pending:
            raise RvdbFormatError(f"{self.name}: incomplete row blocks {sorted(pending)}")


class Hdf5File:
    """Read-only view of the HDF5 subset described in the module docstring."""

    def __init__(self, path: str):
        self._fh = open(path, "rb")
        self.buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        base = self._find_superblock()
        self._parse_superblock(base)

    def close(self):
        try:
            self.buf.close()
        except BufferError:                                     # a chunk view is still alive somewhere: the GC unmaps it
            pass
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- superblock -----------------------------------------------------------------------------
    def _find_superblock(self) -> int:
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == _SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise RvdbFormatError("not an HDF5 file (no superblock signature)")

    def _parse_superblock(self, base: int):
        b = self.buf
        version = b[base + 8]
        if version not in (0, 1):
            raise RvdbFormatError(
                f"HDF5 superblock version {version}: files written with libver='latest' (superblock 2/3, link messages, "
                "fractal heaps) are not supported; the reference writes with h5py's defaults (version 0)")
        self.size_of_offsets, self.size_of_lengths = b[base + 13], b[base + 14]
        if (self.size_of_offsets, self.size_of_lengths) != (8, 8):
            raise RvdbFormatError("HDF5 files with offsets / lengths that are not 8 bytes wide are not supported")
        p = base + 24 + (4 if version == 1 else 0)
        self.base_address = struct.unpack_from("<Q", b, p)[0]
        if self.base_address != 0 or base != 0:
            raise RvdbFormatError("HDF5 files with a user block / non-zero base address are not supported")
        p += 32                                                 # base, free-space info, end of file, driver info
        # root group symbol table entry
        _name_off, header, cache_type = struct.unpack_from("<QQI", b, p)
        self.root_header = header

    # -- object headers --------------------------------------------------------------------------
    def _messages(self, address: int) -> List[Tuple[int, bytes]]:
        b = self.buf
        if b[address:address + 4] == b"OHDR":
            raise RvdbFormatError("version-2 object headers (libver='latest') are not supported")
        version, _r, n_msgs, _refs, hsize = struct.unpack_from("<BBHII", b, address)
        if version != 1:
            raise RvdbFormatError(f"object header version {version} at {address} is not supported")
        msgs: List[Tuple[int, bytes]] = []
        blocks = [(address + 16, hsize)]                        # 12-byte prefix padded to 16
        while blocks and len(msgs) < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < n_msgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = bytes(b[p + 8:p + 8 + msize])
                p += 8 + msize
                if mtype == 0x0010:                             # continuation: more messages elsewhere
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                msgs.append((mtype, body))
        return msgs

    # -- groups -----------------------------------------------------------------------------------
    def _group_links(self, header: int) -> Dict[str, int]:
        for mtype, body in self._messages(header):
            if mtype == 0x0011:                                 # symbol table: B-tree + local heap
                btree, heap = struct.unpack_from("<QQ", body, 0)
                return self._walk_group_btree(btree, self._heap_data(heap))
            if mtype in (0x0002, 0x0006):
                raise RvdbFormatError("new-style groups (link info / link messages) are not supported")
        raise RvdbFormatError(f"object at {header} is not a group")

    def _heap_data(self, heap: int) -> int:
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise RvdbFormatError("bad local heap signature")
        _size, _free, data = struct.unpack_from("<QQQ", b, heap + 8)
        return data

    def _walk_group_btree(self, node: int, heap_data: int) -> Dict[str, int]:
        b = self.buf
        out: Dict[str, int] = {}
        if node == UNDEF:
            return out
        if b[node:node + 4] != b"TREE":
            raise RvdbFormatError("bad group B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        if ntype != 0:
            raise RvdbFormatError("expected a group B-tree node")
        p = node + 24                                           # signature, type, level, entries, two siblings
        for i in range(used):
            child = struct.unpack_from("<Q", b, p + 8 + i * 16)[0]      # key(8) child(8) key child ... key
            if level > 0:
                out.update(self._walk_group_btree(child, heap_data))
                continue
            if b[child:child + 4] != b"SNOD":
                raise RvdbFormatError("bad symbol node signature")
            n_sym = struct.unpack_from("<H", b, child + 6)[0]
            for s in range(n_sym):
                e = child + 8 + s * 40
                name_off, header = struct.unpack_from("<QQ", b, e)
                start = heap_data + name_off
                end = b.find(b"\x00", start)
                out[b[start:end].decode("utf-8")] = header
        return out

    def _resolve(self, path: str) -> int:
        header = self.root_header
        for part in [p for p in path.split("/") if p]:
            links = self._group_links(header)
            if part not in links:
                raise KeyError(path)
            header = links[part]
        return header

    def __contains__(self, path: str) -> bool:
        try:
            self._resolve(path)
            return True
        except (KeyError, RvdbFormatError):
            return False

    def keys(self, path: str = "/") -> List[str]:
        return sorted(self._group_links(self._resolve(path)))

    def is_group(self, path: str) -> bool:
        return any(t == 0x0011 for t, _ in self._messages(self._resolve(path)))

    # -- datasets ---------------------------------------------------------------------------------
    def dataset(self, path: str) -> _Dataset:
        shape = dtype = layout = None
        filters: List[Tuple[int, Tuple[int, ...]]] = []
        for mtype, body in self._messages(self._resolve(path)):
            if mtype == 0x0001:
                shape = self._dataspace(body)
            elif mtype == 0x0003:
                dtype = self._datatype(body)
            elif mtype == 0x0008:
                layout = self._layout(body)
            elif mtype == 0x000B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            raise RvdbFormatError(f"{path} is not a dataset")
        if layout["class"] == "chunked":
            layout["chunk"] = tuple(layout["chunk_with_elem"][:-1])     # the last entry is the element size
        return _Dataset(self, path, shape, dtype, layout, filters)

    @staticmethod
    def _dataspace(body: bytes):
        version, rank, flags = body[0], body[1], body[2]
        if version == 1:
            p = 8
        elif version == 2:
            p = 4
        else:
            raise RvdbFormatError(f"dataspace version {version} is not supported")
        return struct.unpack_from(f"<{rank}Q", body, p) if rank else ()

    @staticmethod
    def _datatype(body: bytes) -> np.dtype:
        cls, bits0 = body[0] & 0x0F, body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:                                            # fixed point: bit 3 of the first flag byte = signed
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        raise RvdbFormatError(f"datatype class {cls} is not supported (only integers and IEEE floats)")

    @staticmethod
    def _layout(body: bytes) -> dict:
        version, cls = body[0], body[1]
        if version != 3:
            raise RvdbFormatError(f"data layout message version {version} is not supported")
        if cls == 0:
            size = struct.unpack_from("<H", body, 2)[0]
            return {"class": "compact", "data": body[4:4 + size]}
        if cls == 1:
            address, size = struct.unpack_from("<QQ", body, 2)
            return {"class": "contiguous", "address": address, "size": size}
        if cls == 2:
            rank = body[2]
            btree = struct.unpack_from("<Q", body, 3)[0]
            dims = struct.unpack_from(f"<{rank}I", body, 11)
            return {"class": "chunked", "btree": btree, "chunk_with_elem": dims}
        raise RvdbFormatError(f"data layout class {cls} is not supported")

    @staticmethod
    def _filters(body: bytes) -> List[Tuple[int, Tuple[int, ...]]]:
        version, n = body[0], body[1]
        out, p = [], 8 if version == 1 else 2
        for _ in range(n):
            fid = struct.unpack_from("<H", body, p)[0]
            if version == 1 or fid >= 256:
                name_len = struct.unpack_from("<H", body, p + 2)[0]
                _flags, ncd = struct.unpack_from("<HH", body, p + 4)
                p += 8
            else:
                name_len = 0
                _flags, ncd = struct.unpack_from("<HH", body, p + 2)
                p += 6
            if version == 1:
                name_len = (name_len + 7) // 8 * 8
            p += name_len
            cd = struct.unpack_from(f"<{ncd}I", body, p)
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, tuple(cd)))
        return out

    def _chunk_entries(self, node: int, rank_plus_1: int):
        """(chunk offsets, address, stored size, filter mask) for every chunk, by walking the version-1 chunk B-tree."""
        b = self.buf
        if node == UNDEF:
            return
        if b[node:node + 4] != b"TREE":
            raise RvdbFormatError("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        if ntype != 1:
            raise RvdbFormatError("expected a chunk B-tree node")
        key_size = 8 + 8 * rank_plus_1                          # chunk size (4), filter mask (4), rank + 1 offsets
        p = node + 24
        for i in range(used):
            k = p + i * (key_size + 8)
            size, mask = struct.unpack_from("<II", b, k)
            offsets = struct.unpack_from(f"<{rank_plus_1}Q", b, k + 8)
            child = struct.unpack_from("<Q", b, k + key_size)[0]
            if level > 0:
                yield from self._chunk_entries(child, rank_plus_1)
            else:
                yield offsets, child, size, mask


# ------------------------------------------------------------------------------------- the .rvdb layout
def read_rvdb(path: str) -> dict:
    """What ``UnifiedIndex._setup_memory_maps`` reads (unified_index.py:1175-1234): returns
    {"file", "embeddings": dataset | None, "faiss_index": uint8 array | None, "metadata": list}.  The caller closes
    ``file`` when it is done streaming the embeddings."""
    f = Hdf5File(path)
    try:
        faiss_bytes = None
        if "faiss_index" in f:                                  # new format: the raw bytes of faiss.write_index
            faiss_bytes = f.dataset("faiss_index").read().reshape(-1)
        elif "index/faiss" in f:                                # old format: an LZ4 frame around them
            faiss_bytes = np.frombuffer(lz4_frame_decompress(f.dataset("index/faiss").read().reshape(-1)), np.uint8)
        emb = None
        if "vectors/embeddings" in f:
            emb = f.dataset("vectors/embeddings")
        elif "vectors" in f and not f.is_group("vectors"):
            emb = f.dataset("vectors")
        meta: list = []
        for name in ("metadata/data", "metadata"):
            if name in f and not f.is_group(name):
                meta = json.loads(lz4_frame_decompress(f.dataset(name).read().reshape(-1)).decode("utf-8"))
                break
        return {"file": f, "embeddings": emb, "faiss_index": faiss_bytes, "metadata": meta}
    except Exception:
        f.close()
        raise
