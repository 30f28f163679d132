"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The
embedding matrix is partitioned into contiguous row blocks at ``add`` time:
rank r owns global ids ``[offsets[r], offsets[r+1])``.  A search replicates the
queries, runs the local exact top-k on every GPU with GLOBAL ids
(``id_offset``) and leaves the hits as PACKED 64-bit keys (score and id in one
word, 8 bytes per hit -- the form the local search already holds before its
final unpack), exchanges the ``[nq, k]`` key lists with ONE
``all_gather_into_tensor`` into a buffer the index owns, and folds them with the
on-device k-way merge kernel (``ivr_topk_merge_keys_device``: one launch, no
scratch, no allocation).  Exactness: the global top-k is a subset of the union
of the local top-k lists.  The key holds a 32-bit id, so a sharded index is
limited to 2^32 - 1 rows in total (checked at ``add``).

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


MAX_SHARDED_ROWS = (1 << 32) - 1          # the exchange key holds a 32-bit global row id


def _cuda_merge(device: int):
    from . import _native as nat

    def merge(keys_parts, k):
        import torch
        n_parts, nq, kk = keys_parts.shape
        D = torch.empty((nq, k), dtype=torch.float32, device=keys_parts.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=keys_parts.device)
        st = torch.cuda.current_stream(keys_parts.device).cuda_stream
        nat.check(nat.lib.ivr_topk_merge_keys_device(device, keys_parts.data_ptr(), n_parts, nq, k,
                                                     D.data_ptr(), I.data_ptr(), st))
        return D, I
    return merge


class ShardedFlatIP:
    """Exact inner-product index row-sharded over the ranks of a process group.

    ``local_index`` / ``merge`` can be injected (the CPU ``gloo`` tests exercise the
    partition / offset / gather logic with a stand-in backend that speaks the same packed-key
    protocol: ``search_keys_tensor(q, k, id_offset, out)`` and ``merge(keys [world, nq, k], k)``);
    by default they are the CUDA index and the CUDA merge kernel and fail loudly without a GPU.
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
        self._keys = None          # [nq, k] local keys and [world, nq, k] gathered keys, reused across searches
        self._gathered = None

    # ---- build ---------------------------------------------------------
    def add_global(self, x) -> None:
        """Every rank passes the same [n, d] matrix (or a view of it); each keeps its block."""
        n = x.shape[0]
        off = partition_rows(n, self.world)
        if self.ntotal_global != 0:
            raise RuntimeError("add_global supports one contiguous build; use add_local to append")
        if n > MAX_SHARDED_ROWS:
            raise ValueError(f"a sharded index holds at most 2^32 - 1 rows in total (got {n})")
        self.id_offset = int(off[self.rank])
        self.local.add(x[off[self.rank]:off[self.rank + 1]])
        self.ntotal_global = n

    def add_local(self, x_local, id_offset: int, ntotal_global: int) -> None:
        """This rank's pre-partitioned block (e.g. generated on device), with its global offset."""
        if int(ntotal_global) > MAX_SHARDED_ROWS:
            raise ValueError(f"a sharded index holds at most 2^32 - 1 rows in total (got {ntotal_global})")
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
        k = int(k)
        if self.world == 1:
            return self.local.search_tensor(q, k, id_offset=self.id_offset)
        nq = q.shape[0]
        if self._keys is None or self._keys.shape != (nq, k) or self._keys.device != q.device:
            self._keys = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            self._gathered = torch.empty((self.world, nq, k), dtype=torch.int64, device=q.device)
        self.local.search_keys_tensor(q, k, id_offset=self.id_offset, out=self._keys)
        dist.all_gather_into_tensor(self._gathered.view(self.world * nq, k), self._keys, group=self.group)
        return self._merge(self._gathered, k)


# ---------------------------------------------------------------------------
# Dedup across GPUs: whole videos per rank, no data-path collective
# ---------------------------------------------------------------------------
def partition_units(lengths, world_size: int) -> np.ndarray:
    """Contiguous assignment of indivisible units (videos) to ranks, balanced by frame count:
    unit u goes to the rank whose ideal frame range contains the unit's midpoint.  Returns unit
    offsets [world_size + 1]; rank r owns units [off[r], off[r+1])."""
    lengths = np.asarray(lengths, dtype=np.int64)
    total = int(lengths.sum())
    off = np.zeros(world_size + 1, dtype=np.int64)
    if total == 0 or len(lengths) == 0:
        off[1:] = len(lengths)
        return off
    mid = np.cumsum(lengths) - lengths / 2.0
    owner = np.minimum((mid * world_size / total).astype(np.int64), world_size - 1)
    for r in range(world_size):
        off[r + 1] = int(np.searchsorted(owner, r, side="right"))
    return off


class ShardedFrameFilter:
    """Near-duplicate pruning of many videos on several GPUs: every rank prunes ITS videos (scene split on
    the consecutive cosine inside each video, then the windowed rule inside each scene -- filter.py:142-315)
    with one fused ``ivr_frame_filter`` call per video; videos are never split, so no halo and no
    collective on the data path.  ``gather`` (optional, control plane) collects the kept indices.

    ``ops`` is the module providing ``calculate_similarities``, ``detect_scene_transitions``,
    ``group_into_scenes`` and ``apply_similarity_filtering_to_scenes`` (default: ``frame_filter``, the
    CUDA path; the world-size-2 gloo test injects the oracle)."""

    def __init__(self, window: int = 8, threshold: float = 0.95, transition_threshold: float = 0.75,
                 min_scene_length: int = 2, group=None, ops=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.transition_threshold, self.min_scene_length = float(transition_threshold), int(min_scene_length)
        self.config = {"enable_similarity_filtering": True, "similarity_threshold": float(threshold),
                       "similarity_window_size": int(window), "use_advanced_similarity_filtering": True,
                       "min_frame_distance": 1}
        if ops is None:
            from . import frame_filter as ops
        self.ops = ops

    def my_units(self, video_bounds):
        off = partition_units([e - s + 1 for s, e in video_bounds], self.world)
        return int(off[self.rank]), int(off[self.rank + 1])

    def filter_videos(self, embeddings, video_bounds) -> list:
        """Kept GLOBAL frame indices of this rank's videos.  ``video_bounds``: inclusive (start, end) per video,
        ascending and non-overlapping; ``embeddings`` may be the whole matrix (only this rank's range is read)."""
        lo, hi = self.my_units(video_bounds)
        if lo >= hi:
            return []
        if hasattr(self.ops, "FrameFilter"):                        # CUDA path: one fused call per video -- the frames cross
            f = self.ops.FrameFilter(window=self.config["similarity_window_size"],     # PCIe once, scene split and greedy
                                     threshold=self.config["similarity_threshold"],    # rule run on the device, and a scene
                                     transition_threshold=self.transition_threshold,   # can never cross a video boundary
                                     min_scene_length=self.min_scene_length)
            kept: list = []
            for vs, ve in video_bounds[lo:hi]:
                kept.extend((f.apply_filters(embeddings[vs:ve + 1]) + vs).tolist())
            return kept
        s0, e1 = video_bounds[lo][0], video_bounds[hi - 1][1]
        x = embeddings[s0:e1 + 1]
        sims = self.ops.calculate_similarities(x)                   # one pass over the rank's frames
        scenes = []
        for vs, ve in video_bounds[lo:hi]:
            inside = sims[vs - s0:ve - s0]                          # cosines between frames of THIS video only
            cuts = self.ops.detect_scene_transitions(inside, self.transition_threshold)
            for a, b in self.ops.group_into_scenes(cuts, ve - vs + 1, self.min_scene_length):
                scenes.append((a + vs - s0, b + vs - s0))
        rows = list(range(s0, e1 + 1))
        return list(self.ops.apply_similarity_filtering_to_scenes(x, rows, scenes, self.config)[1])

    def gather(self, kept_local: list) -> list:
        """All ranks' kept indices in video order (control plane: a Python-object all-gather)."""
        if self.world == 1:
            return list(kept_local)
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, list(kept_local), group=self.group)
        return [i for p in parts for i in p]
