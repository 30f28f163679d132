"""CPU: the row-sharded search path at world_size 2 over gloo (host logic: partition, global
ids, all_gather, merge).  The local index and the merge are stand-ins built on the oracle --
the CUDA kernels behind them are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip, synth


class _OracleLocal:
    def __init__(self, d):
        self.ix = flat_ip.IndexFlatIP(d)
        self._full, self._first = None, 0

    def set_window(self, first=0, count=-1):
        """ivr_index_set_window on the stand-in: searches see the stored rows [first, first + count) only."""
        if self._full is None:
            self._full = self.ix
        if count < 0:
            self.ix, self._first = self._full, 0
            return
        assert 0 <= first and first + count <= self._full.ntotal
        self.ix, self._first = flat_ip.IndexFlatIP(self._full.d), first
        self.ix.add(self._full.xb[first:first + count])

    @property
    def ntotal(self):
        return self.ix.ntotal

    def add(self, x):
        self.ix.add(np.asarray(x))

    def search_tensor(self, q, k, id_offset=0):
        id_offset += self._first
        D, I = self.ix.search(q.numpy(), k)
        return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + id_offset, -1))

    def search_keys_tensor(self, q, k, id_offset=0, out=None):
        """The packed-key protocol of faiss_compat.IndexFlatIP.search_keys_tensor, restated on the CPU."""
        id_offset += self._first
        D, I = self.ix.search(q.numpy(), k)
        keys = torch.from_numpy(flat_ip.pack_keys(D, np.where(I >= 0, I + id_offset, -1)).view(np.int64))
        if out is None:
            return keys
        out.copy_(keys)
        return out


def _merge(keys_parts, k):
    parts = keys_parts.numpy().view(np.uint64)
    D_all, I_all = zip(*(flat_ip.unpack_keys(p) for p in parts))
    D, I = flat_ip.merge_shard_results(list(D_all), list(I_all), k)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, n, d, k, weights, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ivr_b200.sharded import ShardedFlatIP, partition_rows
        xb = synth.clip_like(n, d, seed=21, n_centres=64)
        xq = synth.clip_like(9, d, seed=22, n_centres=64)
        sh = ShardedFlatIP(d, local_index=_OracleLocal(d), merge=_merge)
        assert sh.exchange == "nccl"                     # injected back-ends use the collective transport
        sh.add_global(xb, weights=weights)
        off = partition_rows(n, world, weights)
        assert sh.id_offset == off[rank] and sh.local.ntotal == off[rank + 1] - off[rank]
        D, I = sh.search(torch.from_numpy(xq), k)
        pend = [sh.search_async(torch.from_numpy(xq), k) for _ in range(3)]      # the async front end, same hits
        for h in pend:
            h.synchronize()
            D2, I2 = h.result(copy=False)
            assert np.array_equal(I2.numpy(), I.numpy()) and np.array_equal(D2.numpy(), D.numpy())
        full = flat_ip.IndexFlatIP(d)
        full.add(xb)
        Dr, Ir = full.search(xq, k)
        ok = np.array_equal(I.numpy(), Ir) and np.allclose(D.numpy(), Dr, atol=1e-6)
        out[rank] = bool(ok) and sh.ntotal == n
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k,weights", [(1001, 10, None), (7, 10, None), (1001, 10, [1.0, 3.0])])
def test_sharded_search_world2(n, k, weights):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, 32, k, weights, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


class _ScriptedTimer:
    """Scoring times for the controller: rank 0's GPU is 'slower' per row than rank 1's."""

    def __init__(self, index, ms_per_row):
        self.index, self.ms_per_row = index, ms_per_row

    def start(self, i):
        pass

    def stop(self, i):
        pass

    def mean_ms(self, lo, hi):
        b = self.index._bounds
        return float(b[self.index.rank + 1] - b[self.index.rank]) * self.ms_per_row


def _elastic_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ivr_b200.sharded import ShardedFlatIP, partition_rows
        n, d, k, margin = 4001, 32, 10, 700
        xb = synth.clip_like(n, d, seed=41, n_centres=64)
        full = flat_ip.IndexFlatIP(d)
        full.add(xb)
        sh = ShardedFlatIP(d, local_index=_OracleLocal(d), merge=_merge)
        sh.add_global(xb, margin=margin)
        nominal = partition_rows(n, world)
        assert sh.id_offset == max(0, nominal[rank] - margin)
        assert sh.local.ntotal == min(n, nominal[rank + 1] + margin) - sh.id_offset
        sh.enable_elastic(nominal, margin, period=4, gain=1.0,
                          timer=_ScriptedTimer(sh, 0.003 if rank == 0 else 0.001))
        ok = True
        for s_ in range(14):
            xq = synth.clip_like(5, d, seed=300 + s_, n_centres=64)
            D, I = sh.search(torch.from_numpy(xq), k)
            Dr, Ir = full.search(xq, k)
            ok = ok and np.array_equal(I.numpy(), Ir) and np.allclose(D.numpy(), Dr, atol=1e-6)
        out[rank] = (bool(ok), [list(map(int, rows)) for _, _, rows in sh.balance_log])
    finally:
        dist.destroy_process_group()


def test_elastic_boundaries_world2_move_rows_to_the_faster_rank_and_stay_exact():
    """Rank 0 is 3x slower per row: the controller (every 4 searches) hands its rows to rank 1 until the margin
    stops it; both ranks apply the same boundaries; every search still returns exactly the full-index hits."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_elastic_worker, args=(2, port, out), nprocs=2, join=True)
    res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == res[1][1] and len(res[0][1]) == 3          # steps at searches 4, 8, 12, identical on both ranks
    first, last = res[0][1][0], res[0][1][-1]
    assert first[0] < 2000 < first[1] and sum(first) == 4001
    assert last == [2000 - 700, 2001 + 700]                        # clipped at the margin: rows both ranks hold


def test_rebalance_bounds_controller():
    from ivr_b200.sharded import partition_rows, rebalance_bounds
    nom = partition_rows(100_000_000, 8)
    speed = np.array([1.0, 1.03, 1.02, 1.0, 1.01, 1.01, 1.0, 0.99])
    b = nom.copy()
    for _ in range(5):
        t = np.diff(b) / speed / 3e5
        b = rebalance_bounds(b, t, nom, margin=500_000)
        assert b[0] == 0 and b[-1] == 100_000_000 and np.all(np.diff(b) > 0)
        assert np.all(np.abs(b - nom) <= 500_000) and np.all(b[1:-1] % 1024 == 0)
    t = np.diff(b) / speed / 3e5
    assert t.max() / t.mean() < 1.001                              # from 1.0175 with equal shards
    # degenerate inputs leave the boundaries alone
    assert rebalance_bounds(nom, [1.0] * 7 + [0.0], nom, 1000).tolist() == nom.tolist()
    assert rebalance_bounds(nom, [float("nan")] * 8, nom, 1000).tolist() == nom.tolist()
    # margin 0 = static partition
    assert rebalance_bounds(nom, 1.0 / speed, nom, 0).tolist() == nom.tolist()


def test_partition_rows():
    from ivr_b200.sharded import partition_rows
    off = partition_rows(100_000_000, 8)
    assert off[0] == 0 and off[-1] == 100_000_000 and np.all(np.diff(off) == 12_500_000)
    off = partition_rows(10, 4)
    assert off.tolist() == [0, 2, 5, 7, 10]


def test_partition_rows_by_measured_speed():
    """Shards sized by each GPU's scoring rate: shares proportional to the weights, boundaries aligned, every
    row owned exactly once; equal weights reproduce the balanced partition."""
    from ivr_b200.sharded import partition_rows
    n = 100_000_000
    rates = [305.1, 310.4, 308.0, 303.9, 309.2, 306.6, 307.7, 304.3]
    off = partition_rows(n, 8, weights=rates, align=1024)
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) > 0)
    assert np.all(off[1:-1] % 1024 == 0)
    share = np.diff(off) / n
    assert np.allclose(share, np.array(rates) / sum(rates), atol=2e-5)
    t = np.diff(off) / np.array(rates)                       # predicted time per rank: flat within the alignment
    assert t.max() / t.min() < 1.0002
    assert partition_rows(n, 8, weights=[2.0] * 8).tolist() == partition_rows(n, 8).tolist()
    assert partition_rows(5, 4, weights=[1, 1, 1, 1], align=1024).tolist()[-1] == 5    # tiny index: still covers all rows
    for bad in ([1.0, 0.0], [1.0, float("nan")], [1.0]):
        with pytest.raises(ValueError):
            partition_rows(10, 2, weights=bad)


# ---------------------------------------------------------------------------
# dedup: whole videos per rank, no data-path collective
# ---------------------------------------------------------------------------
class _OracleDedupOps:
    """filter.py's functions from the oracle, with the product module's call signatures."""
    from oracle import dedup as _od
    calculate_similarities = staticmethod(lambda x: _OracleDedupOps._od.calculate_similarities(list(x)))
    detect_scene_transitions = staticmethod(_od.detect_scene_transitions)
    group_into_scenes = staticmethod(_od.group_into_scenes)

    @staticmethod
    def apply_similarity_filtering_to_scenes(x, rows, scenes, config):
        return _OracleDedupOps._od.apply_similarity_filtering_to_scenes(list(x), rows, scenes, config)


def _videos(seed=5):
    rng = np.random.default_rng(seed)
    lens = [int(v) for v in rng.integers(1, 120, size=9)]
    x = np.concatenate([synth.dedup_frames(n, 32, seed=100 + i)[0] for i, n in enumerate(lens)])
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return x, [(int(s), int(s + n - 1)) for s, n in zip(starts, lens)]


def _dedup_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ivr_b200.sharded import ShardedFrameFilter
        x, bounds = _videos()
        sf = ShardedFrameFilter(window=8, threshold=0.95, ops=_OracleDedupOps)
        out[rank] = sf.gather(sf.filter_videos(x, bounds))
    finally:
        dist.destroy_process_group()


def test_sharded_dedup_world2_equals_per_video_reference_rule():
    from oracle import dedup as od
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dedup_worker, args=(2, port, out), nprocs=2, join=True)
    x, bounds = _videos()
    cfg = {"enable_similarity_filtering": True, "similarity_threshold": 0.95, "similarity_window_size": 8,
           "use_advanced_similarity_filtering": True, "min_frame_distance": 1}
    want = []
    for vs, ve in bounds:                                           # the reference processes one video at a time
        v = list(x[vs:ve + 1])
        scenes = od.group_into_scenes(od.detect_scene_transitions(od.calculate_similarities(v), 0.75), len(v), 2)
        want += [vs + i for i in od.apply_similarity_filtering_to_scenes(v, list(range(len(v))), scenes, cfg)[1]]
    assert out[0] == want and out[1] == want and len(want) > 0


def test_partition_units_never_splits_and_balances():
    from ivr_b200.sharded import partition_units
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        lens = rng.integers(1, 1000, size=50)
        off = partition_units(lens, world)
        assert off[0] == 0 and off[-1] == 50 and np.all(np.diff(off) >= 0)
        loads = [int(lens[off[r]:off[r + 1]].sum()) for r in range(world)]
        assert max(loads) - min(loads) <= 2 * int(lens.max())
    assert partition_units([], 4).tolist() == [0, 0, 0, 0, 0]
    assert partition_units([5], 4).tolist()[-1] == 1 and sum(np.diff(partition_units([5], 4))) == 1
