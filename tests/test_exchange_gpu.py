"""GPU: the peer-memory hit exchange (csrc/exchange.cu, K7) with several RANKS IN ONE PROCESS on one device.

The mailboxes of the in-process ranks are connected by plain device pointers (``ivr_exchange_connect_ptrs``); the
kernels, flags, stream waits and the merge are exactly the ones the one-process-per-GPU path uses (there the
pointers come from CUDA IPC: ``tools/sharded_check.py`` under ``gpurun --gpus 2``).  Pushes of a step are queued
before its waits so that a single device can never block on its own stream order.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import comparator, flat_ip, synth  # noqa: E402

TOL = 1e-3


def make_ranks(world, capacity):
    import torch
    from ivr_b200.sharded import _PeerExchange
    ranks = [_PeerExchange(0, r, world, capacity) for r in range(world)]
    bases = [m.base() for m in ranks]
    for m in ranks:
        m.connect_ptrs(bases)
    streams = [torch.cuda.Stream() for _ in range(world)]
    return ranks, streams


@pytest.mark.parametrize("world,nq,k", [(2, 64, 100), (3, 7, 5), (8, 256, 100), (2, 1, 1), (4, 33, 50)])
def test_pushed_keys_merge_like_one_index(world, nq, k):
    """Row-sharded search over in-process ranks: local keys -> push -> stream wait -> merge == the oracle over all
    rows, for several steps through both mailbox slots (slot reuse, rising epochs)."""
    import torch
    import ivr_b200
    from ivr_b200 import _native as nat
    from ivr_b200.sharded import partition_rows, _PeerExchange
    d, n = 128, 24000
    xb = synth.clip_like(n, d, seed=5, n_centres=64)
    ref = flat_ip.IndexFlatIP(d)
    ref.add(xb)
    off = partition_rows(n, world)
    shards = []
    for r in range(world):
        idx = ivr_b200.IndexFlatIP(d)
        idx.add(xb[off[r]:off[r + 1]])
        shards.append(idx)
    ranks, streams = make_ranks(world, nq * k)
    keys = [torch.empty((nq, k), dtype=torch.int64, device="cuda") for _ in range(world)]
    D = [torch.empty((nq, k), dtype=torch.float32, device="cuda") for _ in range(world)]
    I = [torch.empty((nq, k), dtype=torch.int64, device="cuda") for _ in range(world)]
    S = _PeerExchange.SLOTS
    for step in range(5):
        xq = synth.clip_like(nq, d, seed=100 + step, n_centres=64)
        q = torch.from_numpy(xq).cuda()
        slot, epoch = step % S, step // S + 1
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                shards[r].search_keys_tensor(q, k, id_offset=int(off[r]), out=keys[r])
                ranks[r].push(keys[r].data_ptr(), nq * k, slot, epoch, streams[r].cuda_stream)
        for r in range(world):
            ranks[r].wait(slot, epoch, streams[r].cuda_stream)
            nat.check(nat.lib.ivr_topk_merge_keys_device(0, ranks[r].slot_ptr(slot), world, nq, k,
                                                         D[r].data_ptr(), I[r].data_ptr(), streams[r].cuda_stream))
        torch.cuda.synchronize()
        Dr, Ir = ref.search(xq, k)
        for r in range(world):
            bad = comparator.compare_topk(D[r].cpu().numpy(), I[r].cpu().numpy(), Dr, Ir,
                                          lambda ids: ref.scores_of(xq, ids), TOL)
            assert not bad, f"step {step} rank {r}: " + "\n".join(bad[:5])
        for r in range(1, world):                      # every rank merged the same mailbox contents
            assert torch.equal(D[r], D[0]) and torch.equal(I[r], I[0])
    for m in ranks:
        m.close()


def test_mailbox_holds_every_senders_keys_bit_for_bit():
    import torch
    world, n = 3, 1000
    ranks, streams = make_ranks(world, n)
    g = torch.Generator().manual_seed(3)
    sent = [torch.randint(1, 2 ** 62, (n,), dtype=torch.int64, generator=g).cuda() for _ in range(world)]
    for r in range(world):
        ranks[r].push(sent[r].data_ptr(), n, 1, 7, streams[r].cuda_stream)
    for r in range(world):
        ranks[r].wait(1, 7, streams[r].cuda_stream)
    torch.cuda.synchronize()

    class Raw:                                         # the slot as a tensor (no copy)
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    for r in range(world):
        got = torch.as_tensor(Raw(ranks[r].slot_ptr(1), world * n), device="cuda").view(world, n)
        for s in range(world):
            assert torch.equal(got[s], sent[s]), f"mailbox of rank {r}, sender {s}"
    for m in ranks:
        m.close()


def test_argument_checks():
    import ctypes as C
    import torch
    from ivr_b200 import _native as nat
    from ivr_b200.sharded import _PeerExchange
    h = C.c_void_p()
    assert nat.lib.ivr_exchange_create(0, 2, 2, 2, 16, C.byref(h)) == nat.IVR_EINVAL          # rank >= world
    assert nat.lib.ivr_exchange_create(0, 0, 65, 2, 16, C.byref(h)) == nat.IVR_EINVAL         # beyond one merge level
    m = _PeerExchange(0, 0, 2, 16)
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    with pytest.raises(nat.NativeError, match="not connected"):
        m.push(buf.data_ptr(), 16, 0, 1, 0)
    m.connect_ptrs([m.base(), m.base()])
    with pytest.raises(nat.NativeError, match="capacity"):
        m.push(buf.data_ptr(), 17, 0, 1, 0)
    with pytest.raises(nat.NativeError):
        m.push(buf.data_ptr(), 16, 2, 1, 0)                                                   # slot out of range
    assert len(m.ipc_handle()) == nat.IVR_IPC_HANDLE_BYTES
    m.close()
    one = _PeerExchange(0, 0, 1, 8)                                                           # world 1: connected to itself
    one.push(buf.data_ptr(), 8, 0, 1, 0)
    one.wait(0, 1, 0)
    torch.cuda.synchronize()
    one.close()


def test_single_rank_sharded_index_front_ends():
    """Without a process group a ShardedFlatIP is one shard: search / search_async / to_host go straight to the local
    index (no exchange), with the same hits."""
    import torch
    import ivr_b200
    d, n, k = 128, 9000, 20
    xb = synth.clip_like(n, d, seed=15, n_centres=32)
    xq = synth.clip_like(6, d, seed=16, n_centres=32)
    sh = ivr_b200.ShardedFlatIP(d, device=0)
    assert sh.world == 1 and sh.rows_per_rank is None
    sh.add_global(xb)
    q = torch.from_numpy(xq).cuda()
    D, I = sh.search(q, k)
    ref = flat_ip.IndexFlatIP(d)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k)
    assert not comparator.compare_topk(D.cpu().numpy(), I.cpu().numpy(), Dr, Ir, lambda ids: ref.scores_of(xq, ids), TOL)
    hD, hI = torch.empty((6, k)).pin_memory(), torch.empty((6, k), dtype=torch.int64).pin_memory()
    h = sh.search_async(q, k).to_host(hD, hI)
    h.synchronize()
    assert torch.equal(hD, D.cpu()) and torch.equal(hI, I.cpu())
    D2, I2 = sh.search_async(q, k).result(copy=False)
    torch.cuda.synchronize()
    assert torch.equal(D2, D) and torch.equal(I2, I)
