"""BASELINE config E on N GPUs in ONE process group: rows x query-batch sweep of the row-sharded search.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        tools/sweep_multi.py [--rows 10,30,100] [--nq 1,16,256,1024,4096,16384] [--out gpurun_out/sweep_nN]

Same data generator, sharding (whole 500 k-row chunks per rank), search call (ShardedFlatIP.search: local search ->
one all-gather of packed keys -> k-way merge) and timing rule (CUDA events, barrier + sync on both sides, MAX over
ranks) as bench.py -- every cell equals `bench.py --gpus N --rows ROWS --nq NQ --no-extras --no-cpu-baseline` without
paying the index build once per cell.  Parity: 16 queries per cell against an exact fp32 scan of every shard.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ivr_b200  # noqa: E402
from ivr_b200.sharded import ShardedFlatIP, partition_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", default="10,30,100")
ap.add_argument("--nq", default="1,16,256,1024,4096,16384")
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/sweep_multi")
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM, TC, TCS = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0), peaks.get("bf16_tflops_sustained", 1400.0)
d, k = a.dim, a.k
nqs = [int(v) for v in a.nq.split(",")]
cen = bench.centres(d, dev)
Q = bench.gen_queries(max(nqs), d, cen.cpu()).to(dev)
torch.backends.cuda.matmul.allow_tf32 = False
cells = []
for rows_m in [float(v) for v in a.rows.split(",")]:
    n_total = int(rows_m * 1e6)
    chunk = bench.GEN_CHUNK
    c_off = partition_rows(n_total // chunk, world)
    row0, row1 = int(c_off[rank]) * chunk, int(c_off[rank + 1]) * chunk
    index = ShardedFlatIP(d, device=lr) if world > 1 else None
    local = index.local if index else ivr_b200.IndexFlatIP(d, device=lr)
    local.reserve(row1 - row0)
    n_chk = 16
    best_d = torch.full((n_chk, k), -float("inf"), device=dev)
    best_i = torch.full((n_chk, k), -1, dtype=torch.int64, device=dev)
    for c in range(int(c_off[rank]), int(c_off[rank + 1])):
        x = bench.gen_rows(c, chunk, d, cen, dev)
        index.add_local(x, row0, n_total) if index else local.add(x)
        dd, ii = torch.topk(Q[:n_chk] @ x.T, k, dim=1)
        cd, ci = torch.cat([best_d, dd], 1), torch.cat([best_i, ii + c * chunk], 1)
        o = torch.argsort(cd, dim=1, descending=True, stable=True)[:, :k]
        best_d, best_i = torch.gather(cd, 1, o), torch.gather(ci, 1, o)
        del x
    if world > 1:
        gd = [torch.empty_like(best_d) for _ in range(world)]
        gi = [torch.empty_like(best_i) for _ in range(world)]
        dist.all_gather(gd, best_d); dist.all_gather(gi, best_i)
        cd, ci = torch.cat(gd, 1), torch.cat(gi, 1)
        o = torch.argsort(cd, dim=1, descending=True, stable=True)[:, :k]
        best_d, best_i = torch.gather(cd, 1, o), torch.gather(ci, 1, o)
    Dr, Ir = best_d.cpu().numpy(), best_i.cpu().numpy()

    def search(q):
        return index.search(q, k) if index else local.search_tensor(q, k)

    for nq in nqs:
        q = Q[:nq].contiguous()
        D, I = search(q)
        torch.cuda.synchronize()
        m = min(nq, n_chk)
        par = bench.parity_report(D[:m].cpu().numpy(), I[:m].cpu().numpy(), Dr[:m], Ir[:m], k)
        for _ in range(2):
            search(q)
        ts = []
        for _ in range(a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            search(q)
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        local.set_timing(True)
        search(q)
        tm = local.last_timing()
        local.set_timing(False)
        ms = statistics.median(ts)
        n_local = row1 - row0
        c = {"n_gpus": world, "rows": n_total, "rows_per_gpu": n_local, "dim": d, "nq": nq, "k": k, "ms": round(ms, 4),
             "qps": round(nq / ms * 1e3, 1), "kernel": tm["kernel"], "score_ms_rank0": round(tm["score_ms"], 4),
             "tflops_per_gpu": round(2.0 * n_local * d * nq / ms / 1e9, 1), "gbps_per_gpu": round(n_local * d * 2 / ms / 1e6, 1),
             "parity": par["status"], "max_abs_err": par["max_abs_score_error_vs_exact_fp32"]}
        c["frac_hbm"], c["frac_tensor_burst"], c["frac_tensor_sustained"] = (round(c["gbps_per_gpu"] / HBM, 3),
                                                                             round(c["tflops_per_gpu"] / TC, 3),
                                                                             round(c["tflops_per_gpu"] / TCS, 3))
        cells.append(c)
        if rank == 0:
            print(json.dumps(c), flush=True)
    local.close()
    del index, local
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out + ".jsonl", "w") as f:
        for c in cells:
            f.write(json.dumps(c) + "\n")
    with open(a.out + ".md", "w") as f:
        f.write("| GPUs | rows | queries | ms / batch | q/s (all GPUs) | kernel | per-GPU GB/s rows-once (% HBM peak) | per-GPU TFLOP/s "
                "(% burst / % sustained) | parity (16 q) |\n|---|---|---|---|---|---|---|---|---|\n")
        for c in cells:
            f.write(f"| {c['n_gpus']} | {c['rows'] / 1e6:g} M | {c['nq']} | {c['ms']:.3f} | {c['qps']:.0f} | `{c['kernel']}` | "
                    f"{c['gbps_per_gpu']:.0f} ({100 * c['frac_hbm']:.0f} %) | {c['tflops_per_gpu']:.0f} "
                    f"({100 * c['frac_tensor_burst']:.0f} % / {100 * c['frac_tensor_sustained']:.0f} %) | {c['parity']} |\n")
if world > 1:
    dist.destroy_process_group()
