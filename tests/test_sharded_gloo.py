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

    @property
    def ntotal(self):
        return self.ix.ntotal

    def add(self, x):
        self.ix.add(np.asarray(x))

    def search_tensor(self, q, k, id_offset=0):
        D, I = self.ix.search(q.numpy(), k)
        return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + id_offset, -1))


def _merge(D_all, I_all, k):
    D, I = flat_ip.merge_shard_results(list(D_all.numpy()), list(I_all.numpy()), k)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, n, d, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ivr_b200.sharded import ShardedFlatIP, partition_rows
        xb = synth.clip_like(n, d, seed=21, n_centres=64)
        xq = synth.clip_like(9, d, seed=22, n_centres=64)
        sh = ShardedFlatIP(d, local_index=_OracleLocal(d), merge=_merge)
        sh.add_global(xb)
        off = partition_rows(n, world)
        assert sh.id_offset == off[rank] and sh.local.ntotal == off[rank + 1] - off[rank]
        D, I = sh.search(torch.from_numpy(xq), k)
        full = flat_ip.IndexFlatIP(d)
        full.add(xb)
        Dr, Ir = full.search(xq, k)
        ok = np.array_equal(I.numpy(), Ir) and np.allclose(D.numpy(), Dr, atol=1e-6)
        out[rank] = bool(ok) and sh.ntotal == n
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1001, 10), (7, 10)])
def test_sharded_search_world2(n, k):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, 32, k, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_partition_rows():
    from ivr_b200.sharded import partition_rows
    off = partition_rows(100_000_000, 8)
    assert off[0] == 0 and off[-1] == 100_000_000 and np.all(np.diff(off) == 12_500_000)
    off = partition_rows(10, 4)
    assert off.tolist() == [0, 2, 5, 7, 10]
