"""A ``faiss``-shaped module backed by the sm_100a kernels (boundary level B0).

The reference never computes the flat inner-product search itself: it calls
``faiss.IndexFlatIP`` (unified_index.py:503, 1767-1779; core.py:827, 891, 1208;
system.py:1330).  Installing this module as ``faiss``
(``sys.modules['faiss'] = faiss_compat``) makes the reference's own
``UnifiedIndex.search_vectors`` and ``FAISSRetriever.search`` run on the B200
path unchanged; see INTEGRATION.md.

Only the surface the reference touches is provided.  Everything computes on the
GPU through the C ABI (``_native``); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat

__all__ = ["Index", "IndexFlatIP", "IndexFlatL2", "IndexIVFFlat", "normalize_L2",
           "StandardGpuResources", "index_cpu_to_gpu", "index_gpu_to_cpu", "write_index",
           "read_index", "serialize_index", "deserialize_index", "IO_FLAG_MMAP", "METRIC_INNER_PRODUCT"]

IO_FLAG_MMAP = 1
METRIC_INNER_PRODUCT = 0


def _as_f32_2d(x, d=None, what="x"):
    a = np.ascontiguousarray(x, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError(f"{what} must be a 2-D array, got shape {a.shape}")
    if d is not None and a.shape[1] != d:
        raise ValueError(f"{what} has dimension {a.shape[1]}, index has {d}")   # FAISS asserts d == self.d
    return a


class IndexFlatIP:
    """Exact inner-product index; rows live in HBM as fp16, ids are insertion order.

    ``add`` / ``search`` / ``ntotal`` / ``d`` / ``is_trained`` / ``train`` / ``reset`` follow
    ``faiss.IndexFlatIP`` as used by the reference.
    """

    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int, device: int | None = None):
        self.d = int(d)
        self.is_trained = True
        self.device = nat.default_device() if device is None else int(device)
        h = C.c_void_p()
        nat.check(nat.lib.ivr_index_create(self.d, self.device, C.byref(h)))
        self._h = h
        self.search_path = nat.PATH_AUTO

    # -- lifecycle -------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            nat.lib.ivr_index_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("index is closed")
        return self._h

    @property
    def ntotal(self) -> int:
        return int(nat.lib.ivr_index_ntotal(self._handle()))

    def train(self, x=None):          # flat index: nothing to train (core.py:820-825)
        return None

    def reset(self):
        nat.check(nat.lib.ivr_index_reset(self._handle()))

    def set_window(self, first: int = 0, count: int = -1):
        """Searches score only the stored rows [first, first + count) until cleared (``count < 0``); ids stay
        "stored row + id_offset".  No data moves (``ivr_index_set_window``)."""
        nat.check(nat.lib.ivr_index_set_window(self._handle(), int(first), int(count)))

    def reserve(self, n_rows: int):
        nat.check(nat.lib.ivr_index_reserve(self._handle(), int(n_rows)))

    # -- build -----------------------------------------------------------
    def add(self, x) -> None:
        """Append rows (float32 [n, d] host array, or a CUDA float32 torch tensor)."""
        if _is_cuda_tensor(x):
            return self.add_tensor(x)
        a = _as_f32_2d(x, self.d, "add(x)")
        nat.check(nat.lib.ivr_index_add(self._handle(), a.ctypes.data, a.shape[0]))

    def add_tensor(self, x) -> None:
        import torch
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"add(x): expected [n,{self.d}], got {tuple(x.shape)}")
        x = x.to(dtype=torch.float32).contiguous()
        _check_device(x, self.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        nat.check(nat.lib.ivr_index_add_device(self._handle(), x.data_ptr(), x.shape[0], st))

    def reconstruct_n(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """``faiss.Index.reconstruct_n``: rows [first, first + n) as float32 (the stored fp16 values, widened)."""
        n = self.ntotal - first if n is None else int(n)
        out = np.empty((n, self.d), np.float32)
        if n:
            nat.check(nat.lib.ivr_index_reconstruct(self._handle(), int(first), n, out.ctypes.data))
        return out

    # -- search ----------------------------------------------------------
    def search(self, x, k: int):
        """D, I = index.search(x, k): float32 [nq,k] descending, int64 [nq,k], -1 padded."""
        if _is_cuda_tensor(x):
            return self.search_tensor(x, k)
        q = _as_f32_2d(x, self.d, "search(x)")
        k = int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        nq = q.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        if nq:
            nat.check(nat.lib.ivr_index_search(self._handle(), q.ctypes.data, nq, k,
                                               D.ctypes.data, I.ctypes.data, self.search_path))
        return D, I

    def search_tensor(self, q, k: int, id_offset: int = 0, path: int | None = None):
        """Device-resident search on torch's current stream; returns CUDA tensors (D, I)."""
        import torch
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"search(x): expected [nq,{self.d}], got {tuple(q.shape)}")
        q = q.to(dtype=torch.float32).contiguous()
        _check_device(q, self.device)
        nq, k = q.shape[0], int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        if nq:
            st = torch.cuda.current_stream(q.device).cuda_stream
            nat.check(nat.lib.ivr_index_search_device(
                self._handle(), q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), int(id_offset),
                self.search_path if path is None else path, st))
        return D, I

    def search_keys_tensor(self, q, k: int, id_offset: int = 0, out=None, path: int | None = None):
        """Device-resident search that leaves the hits as packed 64-bit keys (int64 tensor [nq, k] holding
        ``order_preserving(score) << 32 | ~(row + id_offset)``, 0 = padding): the 8-byte-per-hit payload of the
        cross-shard exchange (``ShardedFlatIP``).  ``out`` may be a preallocated contiguous int64 [nq, k] tensor."""
        import torch
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"search(x): expected [nq,{self.d}], got {tuple(q.shape)}")
        q = q.to(dtype=torch.float32).contiguous()
        _check_device(q, self.device)
        nq, k = q.shape[0], int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        if out is None:
            out = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        elif out.shape != (nq, k) or out.dtype != torch.int64 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int64 [nq, k] tensor")
        if nq:
            st = torch.cuda.current_stream(q.device).cuda_stream
            nat.check(nat.lib.ivr_index_search_keys_device(
                self._handle(), q.data_ptr(), nq, k, out.data_ptr(), int(id_offset),
                self.search_path if path is None else path, st))
        return out

    # -- instrumentation ---------------------------------------------------
    def set_timing(self, enable: bool = True):
        nat.check(nat.lib.ivr_index_set_timing(self._handle(), int(bool(enable))))

    def last_timing(self) -> dict:
        ms = (C.c_float * 3)()
        ln = (C.c_int * 3)()
        nat.check(nat.lib.ivr_index_last_timing(self._handle(), ms, ln))
        return {"score_ms": ms[0], "merge_ms": ms[1], "prep_ms": ms[2],
                "score_launches": ln[0], "merge_launches": ln[1], "prep_launches": ln[2],
                "path": {0: "empty", 1: "stream", 2: "mma"}.get(
                    nat.lib.ivr_index_last_path(self._handle()), "?"),
                "kernel": (nat.lib.ivr_index_last_kernel(self._handle()) or b"").decode()}


Index = IndexFlatIP


def IndexFlatL2(d):      # core.py:1204 offers it; the hot path in scope is inner product only
    raise NotImplementedError("ivr_b200 implements IndexFlatIP only (exact inner-product search)")


def IndexIVFFlat(*a, **k):   # unified_index.py:897-926 is dead code; the reference forces FlatIP (core.py:1209-1219)
    raise NotImplementedError("approximate IVF indexes are out of scope; use IndexFlatIP")


def normalize_L2(x) -> None:
    """In-place row L2 normalisation of a float32 [n,d] array (unified_index.py:1776)."""
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or x.ndim != 2:
        raise TypeError("normalize_L2 expects a 2-D float32 numpy array")
    if not x.flags.c_contiguous:
        raise ValueError("normalize_L2 expects a C-contiguous array")
    if x.shape[0]:
        nat.check(nat.lib.ivr_normalize_l2(nat.default_device(), x.ctypes.data, x.shape[0], x.shape[1]))


# -- no-op GPU plumbing the reference probes for (core.py:1057-1066, 1222-1228) --
class StandardGpuResources:
    pass


def index_cpu_to_gpu(res, device, index):
    return index


def index_gpu_to_cpu(index):
    return index


# -- persistence in the FAISS on-disk layout -------------------------------------------------
# The reference stores its index with ``faiss.write_index`` (unified_index.py:1811 -> the 'faiss_index' dataset of
# a .rvdb; core.py:987 -> index.faiss) and restores it with ``faiss.deserialize_index`` / ``read_index``
# (unified_index.py:1182-1188; core.py:1057, 4274-4278; system.py:2387).  For an ``IndexFlatIP`` the published
# FAISS layout (faiss/impl/index_write.cpp: fourcc, write_index_header, WRITEXBVECTOR) is, little-endian:
#
#     offset  size  field
#          0     4  fourcc "IxFI"  (inner product; "IxF2" = L2, "IxFl" = legacy flat)
#          4     4  int32  d
#          8     8  int64  ntotal
#         16    16  int64  dummy, dummy   (FAISS writes 1 << 20 twice)
#         32     1  bool   is_trained
#         33     4  int32  metric_type    (0 = METRIC_INNER_PRODUCT, 1 = METRIC_L2)
#         37     8  uint64 number of float32 values that follow (= ntotal * d)
#         45   4*n  float32 rows, row-major
#
# FAISS itself is absent from this image, so the layout is pinned by a hand-built buffer in the tests, not by a
# file FAISS wrote -- stated in DESIGN.md.
_FLAT_HEADER = 45
_FOURCC_IP, _FOURCC_L2, _FOURCC_LEGACY = b"IxFI", b"IxF2", b"IxFl"
_IO_CHUNK_ROWS = 1 << 16


def _flat_header(d: int, ntotal: int) -> bytes:
    import struct
    return (_FOURCC_IP + struct.pack("<iqqq?i", int(d), int(ntotal), 1 << 20, 1 << 20, True, METRIC_INNER_PRODUCT) +
            struct.pack("<Q", int(ntotal) * int(d)))


def _parse_flat_header(buf) -> tuple:
    import struct
    head = bytes(buf[:_FLAT_HEADER])
    if len(head) < _FLAT_HEADER:
        raise ValueError("not a FAISS index: shorter than a flat-index header")
    fourcc = head[:4]
    if fourcc == _FOURCC_L2:
        raise NotImplementedError("the stored index is an IndexFlatL2; ivr_b200 implements inner product only")
    if fourcc not in (_FOURCC_IP, _FOURCC_LEGACY):
        raise NotImplementedError(f"unsupported FAISS index type {fourcc!r}: only IndexFlatIP ('IxFI') is read "
                                  "(the reference forces every index type to FlatIP: core.py:1209-1219)")
    d, ntotal, _d1, _d2, _trained, metric = struct.unpack("<iqqq?i", head[4:37])
    (count,) = struct.unpack("<Q", head[37:45])
    if metric != METRIC_INNER_PRODUCT:
        raise NotImplementedError(f"metric_type {metric}: only METRIC_INNER_PRODUCT (0) is supported")
    if d <= 0 or ntotal < 0 or count != ntotal * d:
        raise ValueError(f"corrupt flat-index header: d={d} ntotal={ntotal} payload={count} floats")
    return d, ntotal


def _add_payload(index: "IndexFlatIP", payload: np.ndarray, ntotal: int, d: int) -> None:
    rows = payload.reshape(ntotal, d)
    index.reserve(ntotal)
    for s in range(0, ntotal, 1 << 20):              # ivr_index_add streams each block through pinned double buffers
        index.add(rows[s:s + (1 << 20)])


def serialize_index(index: IndexFlatIP) -> np.ndarray:
    """``faiss.serialize_index``: the index as a uint8 array in the FAISS IndexFlatIP layout (see above).  The
    payload is what the index holds: the added rows after their one rounding to fp16."""
    n, d = index.ntotal, index.d
    out = np.empty(_FLAT_HEADER + 4 * n * d, np.uint8)
    out[:_FLAT_HEADER] = np.frombuffer(_flat_header(d, n), np.uint8)
    if n:
        rows = out[_FLAT_HEADER:].view(np.float32).reshape(n, d)
        nat.check(nat.lib.ivr_index_reconstruct(index._handle(), 0, n, rows.ctypes.data))
    return out


def deserialize_index(buf, device: int | None = None) -> IndexFlatIP:
    """``faiss.deserialize_index`` for the IndexFlatIP layout: uint8 array / bytes -> index on the GPU."""
    a = np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else np.asarray(buf)
    if a.dtype != np.uint8 or a.ndim != 1:
        raise TypeError("deserialize_index expects a 1-D uint8 array (or bytes)")
    d, ntotal = _parse_flat_header(a)
    if a.size < _FLAT_HEADER + 4 * ntotal * d:
        raise ValueError("truncated FAISS index: payload shorter than the header announces")
    index = IndexFlatIP(d, device=device)
    if ntotal:
        payload = np.frombuffer(a, np.float32, count=ntotal * d, offset=_FLAT_HEADER) if a.flags.c_contiguous \
            else np.ascontiguousarray(a[_FLAT_HEADER:_FLAT_HEADER + 4 * ntotal * d]).view(np.float32)
        _add_payload(index, payload, ntotal, d)
    return index


def write_index(index: IndexFlatIP, path) -> None:
    """``faiss.write_index``: same bytes as ``serialize_index``, streamed to ``path`` in 64 k-row blocks."""
    n, d = index.ntotal, index.d
    with open(path, "wb") as f:
        f.write(_flat_header(d, n))
        blk = np.empty((min(_IO_CHUNK_ROWS, max(n, 1)), d), np.float32)
        for s in range(0, n, _IO_CHUNK_ROWS):
            m = min(_IO_CHUNK_ROWS, n - s)
            nat.check(nat.lib.ivr_index_reconstruct(index._handle(), s, m, blk.ctypes.data))
            f.write(memoryview(blk[:m]).cast("B"))


def read_index(path, flags: int = 0, device: int | None = None) -> IndexFlatIP:
    """``faiss.read_index`` for an IndexFlatIP file.  The float32 payload is memory-mapped (whatever ``flags`` says:
    ``IO_FLAG_MMAP`` is how core.py:4274 asks for it) and streamed page cache -> pinned double buffer -> HBM by
    ``ivr_index_add``, converted to fp16 rows on the device: no second host copy of the matrix is ever made."""
    with open(path, "rb") as f:
        d, ntotal = _parse_flat_header(f.read(_FLAT_HEADER))
    index = IndexFlatIP(d, device=device)
    if ntotal:
        payload = np.memmap(path, dtype=np.float32, mode="r", offset=_FLAT_HEADER, shape=(ntotal * d,))
        _add_payload(index, payload, ntotal, d)
        del payload
    return index


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _check_device(t, device: int):
    if t.device.index != device:
        raise ValueError(f"tensor lives on cuda:{t.device.index}, index on cuda:{device}")
