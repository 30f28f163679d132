"""Multi-GPU parity check of the row-sharded search (run under torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402
from oracle import comparator, flat_ip, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for n, d, nq, k in ((200_003, 512, 300, 100), (1001, 64, 5, 100), (50_000, 128, 2, 10), (5, 32, 3, 10)):
    xb = synth.clip_like(n, d, seed=7, n_centres=128)
    xq = synth.clip_like(nq, d, seed=8, n_centres=128)
    sh = ivr_b200.ShardedFlatIP(d, device=lr)
    sh.add_global(xb)
    D, I = sh.search(torch.from_numpy(xq).cuda(), k)
    torch.cuda.synchronize()
    if rank == 0:
        ref = flat_ip.IndexFlatIP(d)
        ref.add(xb)
        Dr, Ir = ref.search(xq, k)
        bad = comparator.compare_topk(D.cpu().numpy(), I.cpu().numpy(), Dr, Ir, lambda ids: ref.scores_of(xq, ids), 1e-3)
        print(f"sharded x{world}: n={n} d={d} nq={nq} k={k}: violations={len(bad)} {bad[:2]}", flush=True)
        ok = ok and not bad
    # every rank must hold the same merged result
    chk = torch.stack([I.sum().float(), D.sum()])
    ref_chk = chk.clone()
    dist.broadcast(ref_chk, 0)
    assert torch.equal(chk, ref_chk), "ranks disagree on the merged result"

# ---- the two exchange transports agree bit for bit; pipelined searches complete in order under skewed ranks ----
n, d, nq, k = 400_000, 256, 512, 100
xb = synth.clip_like(n, d, seed=17, n_centres=128)
peer = ivr_b200.ShardedFlatIP(d, device=lr, exchange="peer")
nccl = ivr_b200.ShardedFlatIP(d, device=lr, exchange="nccl")
peer.add_global(xb); nccl.add_global(xb)
qs = [torch.from_numpy(synth.clip_like(nq, d, seed=200 + i, n_centres=128)).cuda() for i in range(12)]
want = [nccl.search(q, k) for q in qs]
torch.cuda.synchronize()
assert peer.exchange == "peer", "the mailboxes could not be mapped (no peer access?)"
pending, got = [], []
for i, q in enumerate(qs):
    if (i + rank) % 3 == 0:
        torch.cuda._sleep(int(3e7))                   # ~15 ms of delay on a different rank every step
    h = peer.search_async(q, k)
    pending.append(h)
    if len(pending) == 2:                             # two searches in flight, like bench.py
        D_, I_ = pending.pop(0).result(copy=True)
        got.append((D_, I_))
while pending:
    got.append(pending.pop(0).result(copy=True))
torch.cuda.synchronize()
same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got, want))
if rank == 0:
    print(f"sharded x{world}: 12 pipelined searches over the peer mailboxes (skewed ranks) == NCCL all-gather path: {same}", flush=True)
ok = ok and same


def timed(fn, reps):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(reps); e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def lockstep(index):
    def run(reps):
        for i in range(reps):
            index.search(qs[i % len(qs)], k)
    return run


def pipelined(reps):
    prev = None
    for i in range(reps):
        h = peer.search_async(qs[i % len(qs)], k)
        if prev is not None:
            prev.result(copy=False)
        prev = h
    prev.result(copy=False)


for fn in (lockstep(nccl), lockstep(peer), pipelined):
    fn(5)
t_nccl, t_peer, t_pipe = timed(lockstep(nccl), 50), timed(lockstep(peer), 50), timed(pipelined, 50)
if rank == 0:
    print(f"sharded x{world}: {n} x {d}, {nq} queries, k={k}: NCCL all-gather {t_nccl:.3f} ms/search, "
          f"peer mailboxes {t_peer:.3f}, peer mailboxes pipelined {t_pipe:.3f}", flush=True)

# ---- elastic shard boundaries: replicated margins + the controller; hits stay those of the static partition ----
el = ivr_b200.ShardedFlatIP(d, device=lr, exchange="peer")
el.add_global(xb, margin=24_000)
el._period = 4
same, pend = True, []
for i in range(30):
    if rank == 0 and i % 2 == 0:
        torch.cuda._sleep(int(2e6))                   # rank 0 busy with something else now and then
    pend.append((i % len(qs), el.search_async(qs[i % len(qs)], k)))
    if len(pend) == 2:
        j, h = pend.pop(0)
        D_, I_ = h.result(copy=False)
        same = same and torch.equal(D_, want[j][0]) and torch.equal(I_, want[j][1])
while pend:
    j, h = pend.pop(0)
    D_, I_ = h.result(copy=False)
    same = same and torch.equal(D_, want[j][0]) and torch.equal(I_, want[j][1])
torch.cuda.synchronize()
b = torch.tensor(el._bounds, device="cuda")
b0 = b.clone()
dist.broadcast(b0, 0)
agree = bool(torch.equal(b, b0))
if rank == 0:
    print(f"sharded x{world}: elastic boundaries, 30 pipelined searches: hits identical to the static partition: {same}; "
          f"{len(el.balance_log)} controller steps, rows per rank {np.diff(el._bounds).tolist()} "
          f"(last times {[round(t, 3) for t in el.balance_log[-1][1]]} ms); every rank holds the same boundaries: {agree}",
          flush=True)
assert agree
ok = ok and same
el.close()
peer.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
