"""Tuning sweep on ONE index: python tools/env_sweep.py ROWS DIM NQ,NQ,... VAR=v1,v2,... [VAR2=...] [--k 100] [--reps 5]

Builds the index once, then times IndexFlatIP.search_tensor (CUDA events, median of --reps after 2 warm-ups) for every
combination of the listed environment knobs (the library reads them per call) and every batch size.
"""
import itertools
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--") and "=" in a)
n, d = int(float(args[0])), int(args[1])
nqs = [int(v) for v in args[2].split(",")]
knobs = [(a.split("=")[0], a.split("=")[1].split(",")) for a in args[3:]]
k, reps = int(opts.get("k", 100)), int(opts.get("reps", 5))
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
hbm_ms, = (n * idx.d * 2 / 6552.6e9 * 1e3,)
print(f"rows={n} dim={d} k={k}: HBM floor {hbm_ms:.3f} ms; tensor floor per query {2.0 * n * d / 1654.2e12 * 1e3:.5f} ms", flush=True)
for nq in nqs:
    q = Q[:nq].contiguous()
    floor = max(hbm_ms, 2.0 * n * d * nq / 1654.2e12 * 1e3)
    for combo in itertools.product(*[v for _, v in knobs]):
        for (name, _), val in zip(knobs, combo):
            os.environ[name] = val
        try:
            for _ in range(2):
                idx.search_tensor(q, k)
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                idx.search_tensor(q, k)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = idx.last_timing()
            ms = statistics.median(ts)
            print(f"nq={nq:5d} " + " ".join(f"{nm}={v}" for (nm, _), v in zip(knobs, combo)) +
                  f": {ms:.3f} ms ({100 * floor / ms:.0f}% of the binding floor {floor:.3f}) kernel={t['kernel']} "
                  f"score={t['score_ms']:.3f} merge={t['merge_ms']:.3f}", flush=True)
        except Exception as e:
            print(f"nq={nq} {combo}: {str(e)[:80]}", flush=True)
