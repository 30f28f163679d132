"""Oracle O1: NumPy restatement of ``faiss.IndexFlatIP`` (TEST INFRASTRUCTURE ONLY).

The reference never implements the flat inner-product search itself; it calls
third-party FAISS (``import faiss``: unified_index.py:31, core.py:32; version
un-pinned, not vendored, not installed here).  The call sites that define the
contract this file restates:

* ``faiss.IndexFlatIP(dim)`` + ``index.add(chunk.astype('float32'))``
  -- unified_index.py:1767-1779, core.py:1208, core.py:827
* ``faiss.normalize_L2(chunk)``                       -- unified_index.py:1776
* ``D, I = index.search(q.reshape(1, -1), k)``        -- unified_index.py:503
* ``similarities, indices = index.search(q.astype(np.float32), k)`` -- core.py:891
* ``idx == -1`` means "no more results"               -- unified_index.py:508, core.py:902
* ``index.ntotal``                                    -- core.py:268, 843

Contract restated (published FAISS behaviour):
  S = x @ X^T in fp32; per query the k largest, sorted descending; ids are the
  int64 insertion order; when k > ntotal the tail is padded with id -1 and
  score -FLT_MAX (``std::numeric_limits<float>::lowest()``).  FAISS gives no
  guarantee on the order of exactly-equal scores; this oracle breaks ties by
  the LOWER id so its output is canonical.

PARITY UNPINNED against real FAISS (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

NEG_PAD = np.float32(np.finfo(np.float32).min)  # -3.4028235e+38, FAISS CMin<float>::neutral()


def normalize_L2(x: np.ndarray) -> None:
    """In-place row L2 normalisation (``faiss.normalize_L2``; unified_index.py:1776).

    Zero rows are left untouched (FAISS ``fvec_renorm_L2`` skips rows whose
    squared norm is 0).
    """
    if x.dtype != np.float32 or x.ndim != 2:
        raise TypeError("normalize_L2 expects a 2-D float32 array")
    nr = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    nz = nr > 0
    inv = np.ones_like(nr)
    inv[nz] = np.float32(1.0) / np.sqrt(nr[nz], dtype=np.float32)
    x *= inv[:, None]


def normalize_and_validate(features: np.ndarray) -> np.ndarray:
    """Restates ``FAISSRetriever._normalize_and_validate_features`` (core.py:1176-1196)."""
    if not isinstance(features, np.ndarray):
        raise ValueError("Features must be numpy array")
    if features.size == 0:
        raise ValueError("Features array is empty")
    if features.ndim == 1:
        features = features.reshape(1, -1)
    elif features.ndim != 2:
        raise ValueError(f"Features must be 1D or 2D, got {features.ndim}D")
    if not np.isfinite(features).all():
        raise ValueError("Features contain NaN or infinite values")
    norms = np.linalg.norm(features, axis=1, keepdims=True)
    norms[norms == 0] = 1
    return features / norms


def _topk_desc(scores: np.ndarray, ids: np.ndarray, k: int):
    """Exact top-k of each row of ``scores`` (descending, ties -> lower id).

    ``ids`` are the (int64) ids of the columns.  Returns (D[nq,k'], I[nq,k'])
    with k' = min(k, ncols).
    """
    nq, n = scores.shape
    kk = min(k, n)
    if kk == n:
        sel = np.broadcast_to(np.arange(n), (nq, n))
    else:
        sel = np.argpartition(-scores, kk - 1, axis=1)[:, :kk]
    part = np.take_along_axis(scores, sel, axis=1)
    pid = ids[sel]
    # ties at the k-th boundary: argpartition picks arbitrarily -> redo such rows exactly
    if kk < n:
        kth = part.min(axis=1)
        n_ge = (scores >= kth[:, None]).sum(axis=1)
        for r in np.nonzero(n_ge > kk)[0]:
            order = np.lexsort((ids, -scores[r]))[:kk]
            part[r] = scores[r, order]
            pid[r] = ids[order]
    order = np.lexsort((pid, -part), axis=-1)
    return np.take_along_axis(part, order, axis=1), np.take_along_axis(pid, order, axis=1)


class IndexFlatIP:
    """Flat exact inner-product index (oracle)."""

    def __init__(self, d: int):
        self.d = int(d)
        self.is_trained = True
        self._chunks: list[np.ndarray] = []
        self._xb: np.ndarray | None = None
        self.ntotal = 0

    # -- build ---------------------------------------------------------
    def train(self, x):  # no-op for a flat index (core.py:820-825 guards on is_trained)
        return None

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"add expects [n,{self.d}] float32, got {x.shape}")
        self._chunks.append(x.copy())
        self._xb = None
        self.ntotal += x.shape[0]

    def reset(self) -> None:
        self._chunks, self._xb, self.ntotal = [], None, 0

    @property
    def xb(self) -> np.ndarray:
        if self._xb is None:
            self._xb = (np.concatenate(self._chunks, axis=0) if self._chunks
                        else np.zeros((0, self.d), np.float32))
            self._chunks = [self._xb]
        return self._xb

    # -- search --------------------------------------------------------
    def search(self, x: np.ndarray, k: int, *, q_block: int = 1024, db_block: int = 1 << 18):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"search expects [nq,{self.d}] float32, got {x.shape}")
        nq, k = x.shape[0], int(k)
        D = np.full((nq, k), NEG_PAD, np.float32)
        I = np.full((nq, k), -1, np.int64)
        xb = self.xb
        n = xb.shape[0]
        if n == 0 or k == 0 or nq == 0:
            return D, I
        for q0 in range(0, nq, q_block):
            q = x[q0:q0 + q_block]
            bd = bi = None
            for r0 in range(0, n, db_block):
                blk = xb[r0:r0 + db_block]
                s = q @ blk.T                                   # fp32 sgemm
                d_, i_ = _topk_desc(s, np.arange(r0, r0 + blk.shape[0], dtype=np.int64), k)
                if bd is None:
                    bd, bi = d_, i_
                else:                                            # running merge
                    cd = np.concatenate([bd, d_], axis=1)
                    ci = np.concatenate([bi, i_], axis=1)
                    order = np.lexsort((ci, -cd), axis=-1)[:, :k]
                    bd = np.take_along_axis(cd, order, axis=1)
                    bi = np.take_along_axis(ci, order, axis=1)
            kk = bd.shape[1]
            D[q0:q0 + q.shape[0], :kk] = bd
            I[q0:q0 + q.shape[0], :kk] = bi
        return D, I

    def scores_of(self, x: np.ndarray, ids: np.ndarray) -> np.ndarray:
        """fp32 inner products of query i with rows ids[i, :] (for the comparator)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(ids.shape, np.float32)
        xb = self.xb
        for i in range(x.shape[0]):
            v = ids[i] >= 0
            out[i, v] = xb[ids[i, v]] @ x[i]
            out[i, ~v] = NEG_PAD
        return out


def merge_shard_results(Ds, Is, k: int):
    """Oracle for the multi-GPU merge: concat per-shard (D, I) and keep the k best.

    Precedent in the reference: ``_search_with_remote_index`` concatenates the
    per-shard hit lists, sorts by score and truncates (system.py:1721-1746).
    Here the merge is on raw inner products (descending), ties -> lower id.
    """
    cd = np.concatenate(Ds, axis=1)
    ci = np.concatenate(Is, axis=1)
    # padded entries (-1) must sort last even against real -FLT_MAX scores
    key_id = np.where(ci < 0, np.iinfo(np.int64).max, ci)
    order = np.lexsort((key_id, -cd.astype(np.float64)), axis=-1)[:, :k]
    return np.take_along_axis(cd, order, axis=1), np.take_along_axis(ci, order, axis=1)


def pack_keys(D, I):
    """CPU restatement of the 64-bit candidate key of csrc/common.cuh (``make_key``):
    ``order_preserving(float32 score) << 32 | ~uint32(id)``; entries with id < 0 become key 0 (padding).
    Unsigned descending key order == score descending, then lower id."""
    u = np.ascontiguousarray(D, np.float32).view(np.uint32).astype(np.uint64)
    o = np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    ids = np.asarray(I, np.int64)
    low = (~ids.astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    return np.where(ids < 0, np.uint64(0), (o << np.uint64(32)) | low)


def unpack_keys(keys):
    """Inverse of ``pack_keys``: (D float32, I int64); key 0 -> (-FLT_MAX, -1)."""
    keys = np.asarray(keys, np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o)
    D = u.astype(np.uint32).view(np.float32)
    I = ((~keys) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    pad = keys == 0
    return np.where(pad, np.float32(-3.402823466e+38), D).astype(np.float32), np.where(pad, -1, I)


def similarity_relationships(all_features, top: int = 10, threshold: float = 0.7):
    """Restates MetadataManager._build_similarity_relationships (core.py:3493-3531) on
    {folder: (keys, float32 [n, d])}: sklearn-style cosine matrix per folder, for each frame the
    descending order minus its first entry, top `top`, kept where cosine > threshold.
    Returns ({key: [keys]}, {key: {other_key: cosine}}) -- the second map is for tie-aware tests."""
    graph, sims_of = {}, {}
    for keys, x in all_features.values():
        if len(keys) < 2:
            continue
        x = np.asarray(x, np.float32)
        n = np.sqrt(np.einsum("ij,ij->i", x, x))
        n[n == 0] = 1
        xn = x / n[:, None]
        s = xn @ xn.T
        for i, key in enumerate(keys):
            order = np.argsort(s[i])[::-1][1:top + 1]
            graph[key] = [keys[j] for j in order if s[i][j] > threshold]
            sims_of[key] = {keys[j]: float(s[i][j]) for j in range(len(keys))}
    return graph, sims_of
