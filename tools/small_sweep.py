"""Small-batch comparison on one GPU: every kernel that can serve nq queries, same index.

    python tools/small_sweep.py ROWS DIM [NQ,NQ,...]
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402

n, d = int(float(sys.argv[1])), int(sys.argv[2])
nqs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8,16,32,64,128").split(",")]
g = torch.Generator(device="cuda").manual_seed(1)
cen = torch.nn.functional.normalize(torch.randn(4096, d, generator=g, device="cuda"), dim=1)


def gen(m):
    z = torch.randint(0, 4096, (m,), generator=g, device="cuda")
    return torch.nn.functional.normalize(cen[z] + (0.5 / d ** 0.5) * torch.randn(m, d, generator=g, device="cuda"), dim=1)


idx = ivr_b200.IndexFlatIP(d)
idx.reserve(n)
for s in range(0, n, 1_000_000):
    idx.add(gen(min(1_000_000, n - s)))
idx.set_timing(True)
Q = gen(max(nqs))
floor_ms = n * idx.d * 2 / 6552.6e9 * 1e3
print(f"rows={n} dim={d}: HBM floor {floor_ms:.3f} ms (rows once at 6552.6 GB/s)")
for nq in nqs:
    q = Q[:nq].contiguous()
    out = []
    ref = None
    for name, path, mode in [("stream", 1, "0"), ("small", 2, "3"), ("qres", 2, "1"), ("xres", 2, "2")]:
        if name == "stream" and nq > 8:
            continue
        if name == "xres" and nq < 64:
            continue
        os.environ["IVR_MMA_MODE"] = mode
        try:
            for _ in range(2):
                D, I = idx.search_tensor(q, 100, path=path)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                D, I = idx.search_tensor(q, 100, path=path)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = idx.last_timing()
            if ref is None:
                ref = I.clone()
            same = float((I == ref).float().mean())
            out.append(f"{name} {statistics.median(ts):.3f} ms (score {t['score_ms']:.3f} merge {t['merge_ms']:.3f}, "
                       f"{100 * floor_ms / statistics.median(ts):.0f}% hbm, ids={same:.4f})")
        except Exception as e:  # unsupported shape for this kernel
            out.append(f"{name} n/a ({str(e)[:40]})")
    print(f"nq={nq:4d}: " + " | ".join(out), flush=True)
