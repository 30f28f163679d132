"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The
embedding matrix is partitioned into contiguous row blocks at ``add`` time:
rank r owns global ids ``[offsets[r], offsets[r+1])``.  A search replicates the
queries, runs the local exact top-k on every GPU with GLOBAL ids
(``id_offset``) and leaves the hits as PACKED 64-bit keys (score and id in one
word, 8 bytes per hit -- the form the local search already holds before its
final unpack), exchanges the ``[nq, k]`` key lists and folds them with the
on-device k-way merge kernel (``ivr_topk_merge_keys_device``: one launch, no
scratch, no allocation).  Exactness: the global top-k is a subset of the union
of the local top-k lists.  The key holds a 32-bit id, so a sharded index is
limited to 2^32 - 1 rows in total (checked at ``add``).

The exchange (``exchange="peer"``, the default on GPUs with peer access) is our
own kernel over NVLink peer memory: every rank PUSHES its keys into a mailbox
in each peer's HBM and raises a flag; the receiver waits with a stream memory
operation, not with a spinning kernel (``csrc/exchange.cu``).  Push, wait and
merge run on a side stream, so ``search_async`` lets the scoring of batch i+1
start while slower peers still finish batch i -- the per-batch straggler wait
of a collective disappears from the throughput.  ``exchange="nccl"`` keeps the
ONE ``all_gather_into_tensor`` path (and is what injected CPU back-ends use).

Reference precedent for the semantics: ``_search_with_remote_index``
concatenates the per-shard hit lists, sorts and truncates (system.py:1721-1746;
api.py:1661-1694).  The merge here runs on raw inner products (descending); the
``1 - ip`` mapping of ``search_vectors`` is applied only at that facade.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np


def partition_rows(n_total: int, world_size: int, weights=None, align: int = 1) -> np.ndarray:
    """Contiguous partition.  Balanced by default: offsets[r] = floor(r * n / world).  With ``weights`` (one
    positive number per rank, e.g. the scoring rates ``ShardedFlatIP.scoring_rate`` measured: GPUs under the same
    power cap differ by a few per cent, and with the pipelined exchange throughput is set by the SLOWEST rank's own
    work) rank r receives a share proportional to its weight; inner boundaries are rounded to ``align`` rows."""
    if weights is None:
        return np.array([(r * n_total) // world_size for r in range(world_size + 1)], dtype=np.int64)
    w = np.asarray(weights, dtype=np.float64)
    if w.shape != (world_size,) or not np.all(np.isfinite(w)) or np.any(w <= 0):
        raise ValueError("weights: one positive finite number per rank")
    cum = np.concatenate([[0.0], np.cumsum(w)]) / w.sum()
    off = np.rint(cum * n_total / align).astype(np.int64) * align
    off[0], off[-1] = 0, n_total
    return np.minimum(np.maximum.accumulate(off), n_total)


MAX_SHARDED_ROWS = (1 << 32) - 1          # the exchange key holds a 32-bit global row id


def rebalance_bounds(bounds, times_ms, nominal, margin: int, gain: float = 0.7, align: int = 1024) -> np.ndarray:
    """One step of the elastic-boundary controller (pure, deterministic: every rank computes the same answer from
    the same all-gathered times).  ``bounds`` [world + 1]: the boundaries the timed searches ran with; ``times_ms``
    [world]: every rank's mean local scoring time.  Rank r's rate is rows / time; its target share is proportional to
    its rate; the shares move ``gain`` of the way there.  Inner boundaries are rounded to ``align`` rows and stay
    within ``margin`` rows of the ``nominal`` partition -- the rows both neighbours hold."""
    bounds = np.asarray(bounds, dtype=np.int64)
    nominal = np.asarray(nominal, dtype=np.int64)
    t = np.asarray(times_ms, dtype=np.float64)
    n = np.diff(bounds).astype(np.float64)
    if len(t) != len(n) or np.any(n <= 0) or not np.all(np.isfinite(t)) or np.any(t <= 0):
        return bounds.copy()
    rate = n / t
    share = n + gain * (n.sum() * rate / rate.sum() - n)
    inner = bounds[0] + np.cumsum(share)[:-1]
    inner = np.rint(inner / align).astype(np.int64) * align
    inner = np.clip(inner, nominal[1:-1] - margin, nominal[1:-1] + margin)
    out = bounds.copy()
    out[1:-1] = inner
    return np.maximum.accumulate(out)


class _EventTimer:
    """Local scoring time of the last searches from CUDA events on the scoring stream (no host sync at record time)."""

    def __init__(self, depth: int):
        self.depth, self.ring = depth, {}

    def start(self, i):
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ring[i] = [e, None]

    def stop(self, i):
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ring[i][1] = e
        self.ring.pop(i - self.depth, None)

    def mean_ms(self, lo: int, hi: int) -> float:
        """Mean over searches [lo, hi); blocks the host until search hi - 1 has been scored."""
        self.ring[hi - 1][1].synchronize()
        return sum(self.ring[i][0].elapsed_time(self.ring[i][1]) for i in range(lo, hi)) / (hi - lo)


class _WallTimer(_EventTimer):
    """The same for host back-ends (CPU tests): the local search is synchronous."""

    def start(self, i):
        import time
        self.ring[i] = [time.perf_counter(), None]

    def stop(self, i):
        import time
        self.ring[i][1] = time.perf_counter()
        self.ring.pop(i - self.depth, None)

    def mean_ms(self, lo: int, hi: int) -> float:
        return 1e3 * sum(self.ring[i][1] - self.ring[i][0] for i in range(lo, hi)) / (hi - lo)


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


class _PeerExchange:
    """The mailbox of ``csrc/exchange.cu`` for one rank (ctypes over ``ivr_exchange_*``)."""

    SLOTS = 2

    def __init__(self, device: int, rank: int, world: int, capacity: int):
        from . import _native as nat
        self.nat, self.device, self.rank, self.world, self.capacity = nat, device, rank, world, int(capacity)
        h = C.c_void_p()
        nat.check(nat.lib.ivr_exchange_create(device, rank, world, self.SLOTS, self.capacity, C.byref(h)))
        self.h = h

    def ipc_handle(self) -> bytes:
        buf = (C.c_uint8 * self.nat.IVR_IPC_HANDLE_BYTES)()
        self.nat.check(self.nat.lib.ivr_exchange_ipc_handle(self.h, buf))
        return bytes(buf)

    def base(self) -> int:
        return int(self.nat.lib.ivr_exchange_base(self.h))

    def connect_ipc(self, handles: bytes) -> None:
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self.nat.check(self.nat.lib.ivr_exchange_connect_ipc(self.h, buf))

    def connect_ptrs(self, bases) -> None:
        arr = (C.c_void_p * self.world)(*[C.c_void_p(int(b)) for b in bases])
        self.nat.check(self.nat.lib.ivr_exchange_connect_ptrs(self.h, arr))

    def push(self, keys_ptr: int, n: int, slot: int, epoch: int, stream: int) -> None:
        self.nat.check(self.nat.lib.ivr_exchange_push(self.h, keys_ptr, n, slot, epoch, stream))

    def wait(self, slot: int, epoch: int, stream: int) -> None:
        self.nat.check(self.nat.lib.ivr_exchange_wait(self.h, slot, epoch, stream))

    def slot_ptr(self, slot: int) -> int:
        return int(self.nat.lib.ivr_exchange_slot(self.h, slot))

    def close(self) -> None:
        if self.h:
            self.nat.lib.ivr_exchange_destroy(self.h)
            self.h = None


class PendingSearch:
    """Result of ``ShardedFlatIP.search_async``: the hits exist once the side stream has merged them.
    ``result()`` orders the CURRENT stream behind that and returns ``(D, I)``; with ``copy=False`` these are the
    index's own per-slot buffers, valid until two further searches have been issued.  ``to_host`` queues the
    device-to-host copies behind the merge (no wait on the scoring stream); ``synchronize`` blocks the host."""

    def __init__(self, D, I, done, stream):
        self.D, self.I, self._done, self._stream = D, I, done, stream

    def result(self, copy: bool = True):
        import torch
        if self._done is not None:
            torch.cuda.current_stream(self.D.device).wait_event(self._done)
        return (self.D.clone(), self.I.clone()) if copy else (self.D, self.I)

    def to_host(self, D_host, I_host) -> "PendingSearch":
        import torch
        if self._stream is None:
            D_host.copy_(self.D, non_blocking=True); I_host.copy_(self.I, non_blocking=True)
            self._done = torch.cuda.Event(); self._done.record()
            return self
        with torch.cuda.stream(self._stream):
            D_host.copy_(self.D, non_blocking=True); I_host.copy_(self.I, non_blocking=True)
            self._done = torch.cuda.Event(); self._done.record(self._stream)
        return self

    def synchronize(self) -> None:
        if self._done is not None:
            self._done.synchronize()


class ShardedFlatIP:
    """Exact inner-product index row-sharded over the ranks of a process group.

    ``local_index`` / ``merge`` can be injected (the CPU ``gloo`` tests exercise the
    partition / offset / gather logic with a stand-in backend that speaks the same packed-key
    protocol: ``search_keys_tensor(q, k, id_offset, out)`` and ``merge(keys [world, nq, k], k)``);
    by default they are the CUDA index and the CUDA merge kernel and fail loudly without a GPU.
    """

    def __init__(self, d: int, group=None, device: Optional[int] = None,
                 local_index=None, merge: Optional[Callable] = None, exchange: Optional[str] = None):
        import os
        import torch.distributed as dist
        self.d = int(d)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        cuda_backend = local_index is None
        if cuda_backend:
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
        # exchange transport: "peer" (mailboxes over NVLink peer memory, pipelined) or "nccl" (one all-gather);
        # None = IVR_EXCHANGE, else peer for the CUDA back-end (falling back to nccl, on every rank together, when the
        # mailboxes cannot be mapped) and nccl for injected back-ends
        want = exchange or (os.environ.get("IVR_EXCHANGE") if cuda_backend else None) or ("peer" if cuda_backend else "nccl")
        if want not in ("peer", "nccl"):
            raise ValueError(f"exchange must be 'peer' or 'nccl', got {want!r}")
        self.exchange = want
        self._auto_exchange = exchange is None and "IVR_EXCHANGE" not in os.environ
        self._mail = None          # _PeerExchange, created at the first search (capacity = nq * k, regrown collectively)
        self._comm = None          # side stream: push -> wait -> merge of consecutive searches, in order
        self._step = 0
        self._slots = None         # per slot: keys, D, I, pushed-event
        # elastic shard boundaries (enabled by add_global / add_local with margin > 0)
        self._nominal = None       # [world + 1] partition the shards were built for
        self._bounds = None        # [world + 1] boundaries the NEXT search runs with (identical on every rank)
        self._margin = 0           # rows of each neighbour this rank also stores on either side
        self._period, self._gain = 8, 0.7
        self._nsearch = 0
        self._timer = None
        self._ctl, self._ctl_ready = None, False   # host-side (gloo) group the controller all-gathers the times on
        self.balance_log = []      # (search number, times_ms, rows per rank) of every controller step

    # ---- build ---------------------------------------------------------
    def add_global(self, x, weights=None, margin: int = 0) -> None:
        """Every rank passes the same [n, d] matrix (or a view of it); each keeps its block (``weights``: see
        ``partition_rows``; the same on every rank).  ``margin`` > 0 also keeps that many rows of either neighbour and
        turns on the elastic boundaries (``enable_elastic``)."""
        n = x.shape[0]
        off = partition_rows(n, self.world, weights)
        if self.ntotal_global != 0:
            raise RuntimeError("add_global supports one contiguous build; use add_local to append")
        if n > MAX_SHARDED_ROWS:
            raise ValueError(f"a sharded index holds at most 2^32 - 1 rows in total (got {n})")
        lo = max(0, int(off[self.rank]) - int(margin))
        hi = min(n, int(off[self.rank + 1]) + int(margin))
        self.id_offset = lo
        self.local.add(x[lo:hi])
        self.ntotal_global = n
        if margin > 0 and self.world > 1:
            self.enable_elastic(off, margin)

    def add_local(self, x_local, id_offset: int, ntotal_global: int) -> None:
        """This rank's pre-partitioned block (e.g. generated on device), with its global offset."""
        if int(ntotal_global) > MAX_SHARDED_ROWS:
            raise ValueError(f"a sharded index holds at most 2^32 - 1 rows in total (got {ntotal_global})")
        if self.local.ntotal == 0:
            self.id_offset = int(id_offset)
        self.local.add(x_local)
        self.ntotal_global = int(ntotal_global)

    def enable_elastic(self, nominal, margin: int, period: int = 8, gain: float = 0.7, ctl_group=None,
                       timer=None) -> None:
        """ELASTIC SHARD BOUNDARIES.  GPUs under the same power cap differ by a few per cent and drift with
        temperature; with the pipelined exchange the job runs at the pace of the slowest rank's own work.  Every rank
        therefore also stores ``margin`` rows of either neighbour (HBM is plentiful: + 2 * margin rows), which lets
        the boundary between two ranks move between two searches WITHOUT moving data: each search scores the rows
        ``[bounds[r], bounds[r+1])`` through ``ivr_index_set_window``.  Every ``period`` searches the ranks
        all-gather their mean local scoring times on a host-side group and apply ``rebalance_bounds`` -- the same
        deterministic step on every rank, so every row is scored by exactly one rank in every search and the hits
        stay exactly those of the static partition.

        ``nominal`` [world + 1]: the partition (the same on every rank); this rank must hold the rows
        ``[nominal[r] - margin, nominal[r+1] + margin)`` clipped to the index, starting at ``id_offset``.
        ``ctl_group``: a gloo group for the controller (created on first use otherwise -- collective)."""
        nominal = np.asarray(nominal, dtype=np.int64)
        if nominal.shape != (self.world + 1,) or nominal[0] != 0 or nominal[-1] != self.ntotal_global:
            raise ValueError("nominal: offsets [world + 1] from 0 to ntotal")
        lo = max(0, int(nominal[self.rank]) - int(margin))
        hi = min(self.ntotal_global, int(nominal[self.rank + 1]) + int(margin))
        if self.id_offset != lo or self.local.ntotal != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo}, {hi}) for margin {margin}: it holds "
                             f"[{self.id_offset}, {self.id_offset + self.local.ntotal})")
        self._nominal, self._bounds, self._margin = nominal, nominal.copy(), int(margin)
        self._period, self._gain = max(4, int(period)), float(gain)
        self._ctl, self._ctl_ready, self._nsearch = ctl_group, ctl_group is not None, 0
        self._timer = timer
        self.balance_log = []

    def _elastic_begin(self, q) -> int:
        """Controller step (every ``period`` searches) + the row window of this search.  Returns the search number."""
        import torch
        import torch.distributed as dist
        i, P = self._nsearch, self._period
        self._nsearch += 1
        if self._timer is None:
            self._timer = _EventTimer(2 * P) if getattr(q, "is_cuda", False) else _WallTimer(2 * P)
        if i >= P and i % P == 0:
            # the two latest searches may still be in flight: use the P - 2 before them (all run with these bounds)
            mine = torch.tensor([self._timer.mean_ms(i - P, i - 2)], dtype=torch.float64)
            if not self._ctl_ready:                      # the controller talks on a HOST-side group: a GPU collective
                try:                                     # would hold SMs the scoring kernels need while it waits
                    if dist.get_backend(self.group) != "gloo":
                        ranks = dist.get_process_group_ranks(self.group) if self.group is not None else None
                        self._ctl = dist.new_group(ranks=ranks, backend="gloo")
                    else:
                        self._ctl = self.group
                except Exception as e:                   # no host-side transport on this box (every rank alike):
                    import warnings                      # keep the boundaries where they are
                    warnings.warn(f"elastic shard boundaries disabled: cannot create the gloo control group ({e})")
                    self._period = 1 << 62
                    self._ctl_ready = True
                    return self._elastic_window(i)
                self._ctl_ready = True
            times = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(times, mine, group=self._ctl)
            times = [float(t.item()) for t in times]
            self._bounds = rebalance_bounds(self._bounds, times, self._nominal, self._margin, self._gain)
            self.balance_log.append((i, times, np.diff(self._bounds).tolist()))
        return self._elastic_window(i)

    def _elastic_window(self, i: int) -> int:
        first = int(self._bounds[self.rank]) - self.id_offset
        self.local.set_window(first, int(self._bounds[self.rank + 1] - self._bounds[self.rank]))
        self._timer.start(i)
        return i

    @property
    def ntotal(self) -> int:
        return self.ntotal_global

    @property
    def rows_per_rank(self):
        """Rows every rank scores in the NEXT search (elastic boundaries), or None for a static partition."""
        return None if self._bounds is None else [int(v) for v in np.diff(self._bounds)]

    def scoring_rate(self, q, k: int, reps: int = 8, warm: int = 3) -> float:
        """Rows per millisecond THIS rank's GPU scores for the batch ``q`` on its current shard (device-timed, local
        search only, no exchange).  All-gather the rates and pass them to ``partition_rows`` / ``add_global`` as
        ``weights`` to size the shards by measured speed."""
        self.local.set_timing(True)
        ms = []
        for i in range(warm + reps):
            self.local.search_keys_tensor(q, int(k), id_offset=self.id_offset)
            t = self.local.last_timing()
            if i >= warm:
                ms.append(t["score_ms"] + t["merge_ms"] + t["prep_ms"])
        self.local.set_timing(False)
        return float(self.local.ntotal) / (sum(ms) / len(ms))

    # ---- search --------------------------------------------------------
    def search(self, q, k: int):
        """q: [nq, d] tensor (CUDA for the product path), replicated on every rank.
        Returns (D [nq,k] float32 descending, I [nq,k] int64 global ids) on every rank."""
        return self.search_async(q, k).result(copy=True)

    def search_async(self, q, k: int) -> PendingSearch:
        """Queue one search and return at once.  The local scoring runs on the current stream; with the peer
        exchange the push / arrival wait / merge run on the index's side stream, so the NEXT ``search_async`` starts
        scoring while this one's exchange is still waiting for slower ranks.  Searches complete in issue order."""
        import torch
        import torch.distributed as dist
        k = int(k)
        if self.world == 1:
            D, I = self.local.search_tensor(q, k, id_offset=self.id_offset)
            return PendingSearch(D, I, None, None)
        nq = q.shape[0]
        if self.exchange == "peer" and nq > 0 and self._ensure_mailbox(nq * k, q.device):
            return self._search_peer(q, nq, k)
        if self._keys is None or self._keys.shape != (nq, k) or self._keys.device != q.device:
            self._keys = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            self._gathered = torch.empty((self.world, nq, k), dtype=torch.int64, device=q.device)
        i = self._elastic_begin(q) if self._bounds is not None else -1
        self.local.search_keys_tensor(q, k, id_offset=self.id_offset, out=self._keys)
        if i >= 0:
            self._timer.stop(i)
        dist.all_gather_into_tensor(self._gathered.view(self.world * nq, k), self._keys, group=self.group)
        D, I = self._merge(self._gathered, k)
        return PendingSearch(D, I, None, None)

    # ---- peer-memory exchange (csrc/exchange.cu) ---------------------------
    def _ensure_mailbox(self, n_entries: int, device) -> bool:
        """Collective: (re)creates the mailboxes when the payload outgrows them and maps the peers' through CUDA
        IPC.  Every rank calls it with the same size (queries are replicated).  Returns False -- on every rank --
        when some rank cannot map a peer and the transport was chosen automatically (then NCCL takes over)."""
        import torch
        import torch.distributed as dist
        if self._mail is not None and n_entries <= self._mail.capacity:
            return True
        from . import _native as nat
        if self._mail is not None:                      # nobody may still be writing into the old mailboxes
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)
            self._mail.close()
            self._mail = None
        # every rank reaches both collectives below whatever happens locally: a rank that cannot create or map a
        # mailbox reports it through the all-reduce instead of leaving the others waiting
        mail, err = None, None
        try:
            mail = _PeerExchange(self.device, self.rank, self.world, n_entries)
            handle = mail.ipc_handle()
        except nat.NativeError as e:
            err, handle = e, bytes(nat.IVR_IPC_HANDLE_BYTES)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(device)
        allh = torch.empty((self.world, nat.IVR_IPC_HANDLE_BYTES), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        if err is None:
            try:
                mail.connect_ipc(allh.cpu().numpy().tobytes())
            except nat.NativeError as e:
                err = e
        flag = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if mail is not None:
                mail.close()
            if not self._auto_exchange:
                raise RuntimeError(f"exchange='peer' is not available on every rank (rank {self.rank}: {err or 'ok'})")
            self.exchange = "nccl"
            return False
        self._mail = mail
        self._step = 0
        self._slots = None
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=device)
        return True

    def _search_peer(self, q, nq: int, k: int) -> PendingSearch:
        import torch
        S = _PeerExchange.SLOTS
        if self._slots is None or self._slots[0]["keys"].shape != (nq, k) or self._slots[0]["keys"].device != q.device:
            if self._slots is not None:
                self._comm.synchronize()
            self._slots = [{"keys": torch.empty((nq, k), dtype=torch.int64, device=q.device),
                            "D": torch.empty((nq, k), dtype=torch.float32, device=q.device),
                            "I": torch.empty((nq, k), dtype=torch.int64, device=q.device),
                            "pushed": None} for _ in range(S)]
        slot, epoch = self._step % S, self._step // S + 1
        self._step += 1
        b = self._slots[slot]
        cur = torch.cuda.current_stream(q.device)
        if b["pushed"] is not None:                     # the push of search (step - S) has read this keys buffer
            cur.wait_event(b["pushed"])
        i = self._elastic_begin(q) if self._bounds is not None else -1
        self.local.search_keys_tensor(q, k, id_offset=self.id_offset, out=b["keys"])
        if i >= 0:
            self._timer.stop(i)
        scored = torch.cuda.Event()
        scored.record(cur)
        comm = self._comm
        comm.wait_event(scored)                          # also orders the merge behind every consumer of D / I[slot]
        cs = comm.cuda_stream                            # that was queued on the current stream before this search
        self._mail.push(b["keys"].data_ptr(), nq * k, slot, epoch, cs)
        b["pushed"] = torch.cuda.Event()
        b["pushed"].record(comm)
        self._mail.wait(slot, epoch, cs)
        from . import _native as nat
        nat.check(nat.lib.ivr_topk_merge_keys_device(self.device, self._mail.slot_ptr(slot), self.world, nq, k,
                                                     b["D"].data_ptr(), b["I"].data_ptr(), cs))
        done = torch.cuda.Event()
        done.record(comm)
        return PendingSearch(b["D"], b["I"], done, comm)

    def close(self) -> None:
        """Collective: waits for every rank's outstanding exchanges, then frees the mailbox."""
        if self._mail is not None:
            import torch
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._mail.close()
            self._mail = None


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
