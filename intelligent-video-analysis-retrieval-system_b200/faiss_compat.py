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


# -- minimal persistence so reference save/load paths survive --------------
def serialize_index(index: IndexFlatIP) -> np.ndarray:
    """Not provided: index (de)serialisation is the .rvdb container I/O (unified_index.py:1182-1188), which is out of
    scope; rebuild with ``IndexFlatIP.add`` from the stored embeddings instead."""
    raise NotImplementedError(
        "index (de)serialisation is .rvdb I/O, which is out of scope (SURVEY.md section 8f); "
        "rebuild with IndexFlatIP.add from the stored embeddings instead")


def deserialize_index(buf):
    raise NotImplementedError(serialize_index.__doc__)


def write_index(index, path):
    raise NotImplementedError(serialize_index.__doc__)


def read_index(path, flags=0):
    raise NotImplementedError(serialize_index.__doc__)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _check_device(t, device: int):
    if t.device.index != device:
        raise ValueError(f"tensor lives on cuda:{t.device.index}, index on cuda:{device}")
