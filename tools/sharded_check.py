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
dist.destroy_process_group()
sys.exit(0 if ok else 1)
