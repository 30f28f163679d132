"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The
embedding matrix is partitioned into contiguous row blocks at ``add`` time:
rank r owns global ids ``[offsets[r], offsets[r+1])``.  A search replicates the
queries, runs the local exact top-k on every GPU with GLOBAL ids
(``id_offset``), exchanges the ``[nq, k]`` (score, id) lists with ONE
``all_gather`` and folds them with the on-device k-way merge kernel
(``ivr_topk_merge_device``).  Exactness: the global top-k is a subset of the
union of the local top-k lists.

Reference precedent for the semantics: ``_search_with_remote_index``
concatenates the per-shard hit lists, sorts and truncates (system.py:1721-1746;
api.py:1661-1694).  The merge here runs on raw inner products (descending); the
``1 - ip`` mapping of ``search_vectors`` is applied only at that facade.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np


def partition_rows(n_total: int, world_size: int) -> np.ndarray:
    """Contiguous balanced partition: offsets[r] = floor(r * n / world)."""
    return np.array([(r * n_total) // world_size for r in range(world_size + 1)], dtype=np.int64)


def _cuda_merge(device: int):
    from . import _native as nat

    def merge(D_parts, I_parts, k):
        import torch
        n_parts, nq, kk = D_parts.shape
        D = torch.empty((nq, k), dtype=torch.float32, device=D_parts.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=D_parts.device)
        st = torch.cuda.current_stream(D_parts.device).cuda_stream
        nat.check(nat.lib.ivr_topk_merge_device(device, D_parts.data_ptr(), I_parts.data_ptr(),
                                                n_parts, nq, k, D.data_ptr(), I.data_ptr(), st))
        return D, I
    return merge


class ShardedFlatIP:
    """Exact inner-product index row-sharded over the ranks of a process group.

    ``local_index`` / ``merge`` can be injected (the CPU ``gloo`` tests exercise the
    partition / offset / gather logic with a stand-in backend); by default they are
    the CUDA index and the CUDA merge kernel and fail loudly without a GPU.
    """

    def __init__(self, d: int, group=None, device: Optional[int] = None,
                 local_index=None, merge: Optional[Callable] = None):
        import torch.distributed as dist
        self.d = int(d)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if local_index is None:
            from . import _native as nat
            from .faiss_compat import IndexFlatIP
            self.device = nat.default_device() if device is None else device
            local_index = IndexFlatIP(self.d, device=self.device)
            merge = merge or _cuda_merge(self.device)
        else:
            self.device = device
        self.local = local_index
        self._merge = merge
        self.id_offset = 0
        self.ntotal_global = 0

    # ---- build ---------------------------------------------------------
    def add_global(self, x) -> None:
        """Every rank passes the same [n, d] matrix (or a view of it); each keeps its block."""
        n = x.shape[0]
        off = partition_rows(n, self.world)
        if self.ntotal_global != 0:
            raise RuntimeError("add_global supports one contiguous build; use add_local to append")
        self.id_offset = int(off[self.rank])
        self.local.add(x[off[self.rank]:off[self.rank + 1]])
        self.ntotal_global = n

    def add_local(self, x_local, id_offset: int, ntotal_global: int) -> None:
        """This rank's pre-partitioned block (e.g. generated on device), with its global offset."""
        if self.local.ntotal == 0:
            self.id_offset = int(id_offset)
        self.local.add(x_local)
        self.ntotal_global = int(ntotal_global)

    @property
    def ntotal(self) -> int:
        return self.ntotal_global

    # ---- search --------------------------------------------------------
    def search(self, q, k: int):
        """q: [nq, d] tensor (CUDA for the product path), replicated on every rank.
        Returns (D [nq,k] float32 descending, I [nq,k] int64 global ids) on every rank."""
        import torch
        import torch.distributed as dist
        D_loc, I_loc = self.local.search_tensor(q, k, id_offset=self.id_offset)
        if self.world == 1:
            return D_loc, I_loc
        nq = D_loc.shape[0]
        D_all = torch.empty((self.world * nq, k), dtype=D_loc.dtype, device=D_loc.device)
        I_all = torch.empty((self.world * nq, k), dtype=I_loc.dtype, device=I_loc.device)
        dist.all_gather_into_tensor(D_all, D_loc.contiguous(), group=self.group)
        dist.all_gather_into_tensor(I_all, I_loc.contiguous(), group=self.group)
        return self._merge(D_all.view(self.world, nq, k), I_all.view(self.world, nq, k), int(k))
