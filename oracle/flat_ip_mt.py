"""Multi-threaded CPU flat inner-product search: the TIMED CPU baseline (TEST / BENCH INFRASTRUCTURE ONLY).

Same contract as ``oracle/flat_ip.py`` (the ``faiss.IndexFlatIP`` calls of unified_index.py:503, 1767-1779 and
core.py:827, 891), restated the way FAISS's own BLAS path works -- blocks of database rows, one fp32 sgemm per block,
a per-query k-selection of the block, a running merge -- but on torch's CPU kernels so that BOTH stages use every
host thread (``torch.matmul`` -> multi-threaded sgemm, ``torch.topk`` -> rows in parallel).  The NumPy oracle does its
selection with single-threaded ``argpartition``, which made it 4-5x slower than this on 8-16 cores: a CPU baseline
should not lose to its own top-k.  tests/test_oracle_golden.py pins this file against the NumPy oracle.
"""
from __future__ import annotations

import numpy as np
import torch

NEG_PAD = float(np.finfo(np.float32).min)


def normalize_L2(x: np.ndarray) -> None:
    from . import flat_ip
    flat_ip.normalize_L2(x)


class IndexFlatIP:
    """faiss-shaped: ``add`` / ``search`` / ``ntotal`` / ``d`` / ``is_trained`` / ``train`` / ``reset``."""

    def __init__(self, d: int, db_block: int = 1 << 18):
        self.d, self.is_trained, self.db_block = int(d), True, int(db_block)
        self._chunks: list = []
        self._xb = None
        self.ntotal = 0
        self.search_seconds = 0.0        # time spent inside search() (lets a caller split wrapper time from search time)

    def train(self, x=None):
        return None

    def reset(self):
        self._chunks, self._xb, self.ntotal = [], None, 0

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"add expects [n,{self.d}] float32, got {x.shape}")
        self._chunks.append(torch.from_numpy(x))          # no copy: the caller's array is the storage
        self._xb = None
        self.ntotal += x.shape[0]

    def _rows(self):
        if self._xb is None:
            self._xb = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks, 0)
            self._chunks = [self._xb]
        return self._xb

    def search(self, x, k: int):
        import time
        t0 = time.perf_counter()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"search expects [nq,{self.d}] float32, got {x.shape}")
        nq, k = x.shape[0], int(k)
        D = np.full((nq, k), NEG_PAD, np.float32)
        I = np.full((nq, k), -1, np.int64)
        if self.ntotal == 0 or nq == 0 or k == 0:
            return D, I
        xb, q = self._rows(), torch.from_numpy(x)
        bd = bi = None
        for r0 in range(0, xb.shape[0], self.db_block):
            blk = xb[r0:r0 + self.db_block]
            s = q @ blk.T                                            # fp32 sgemm, all threads
            d_, i_ = torch.topk(s, min(k, blk.shape[0]), dim=1)     # per-query selection, rows in parallel
            i_ = i_ + r0
            if bd is None:
                bd, bi = d_, i_
            else:
                cd, ci = torch.cat([bd, d_], 1), torch.cat([bi, i_], 1)
                o = torch.topk(cd, min(k, cd.shape[1]), dim=1)
                bd, bi = o.values, torch.gather(ci, 1, o.indices)
        # canonical order: score descending, ties -> lower id (torch.topk leaves ties in arbitrary order)
        bd, bi = bd.numpy(), bi.numpy()
        order = np.lexsort((bi, -bd), axis=-1)
        kk = bd.shape[1]
        D[:, :kk] = np.take_along_axis(bd, order, axis=1)
        I[:, :kk] = np.take_along_axis(bi, order, axis=1)
        self.search_seconds += time.perf_counter() - t0
        return D, I


def set_threads(n: int | None = None) -> int:
    """Use ``n`` host threads (default: every core this process may run on) for torch AND the BLAS NumPy links --
    explicitly, because launchers such as torch.distributed.run export OMP_NUM_THREADS=1.  Returns the count."""
    import os
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    n = max(1, int(n))
    torch.set_num_threads(n)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def faiss_module(index_cls=None):
    """A ``faiss``-shaped module around this index (for running the reference's wrappers on the host)."""
    import types
    m = types.ModuleType("faiss")
    m.IndexFlatIP = index_cls or IndexFlatIP
    m.Index = m.IndexFlatIP
    m.normalize_L2 = normalize_L2
    m.METRIC_INNER_PRODUCT = 0
    return m
