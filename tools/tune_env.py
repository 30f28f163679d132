"""Build one index, then time search under several environment settings (read per call by the library).

    python tools/tune_env.py ROWS DIM NQ "A=1,B=2" "A=3" ...
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402

n, d = int(float(sys.argv[1])), int(sys.argv[2])
nqs = [int(v) for v in sys.argv[3].split(",")]
settings = sys.argv[4:] or [""]
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
touched = set()
for nq in nqs:
    q = Q[:nq].contiguous()
    for st in settings:
        for key in touched:
            os.environ.pop(key, None)
        for kv in filter(None, st.split(",")):
            key, v = kv.split("=")
            os.environ[key] = v
            touched.add(key)
        for _ in range(2):
            idx.search_tensor(q, 100)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx.search_tensor(q, 100)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = idx.last_timing()
        ms = statistics.median(ts)
        print(f"rows={n} nq={nq} [{st or 'default'}]: {ms:.3f} ms  {2.0 * n * d * nq / ms / 1e9:.0f} TFLOP/s  "
              f"kernel={t['kernel']} launches={t['score_launches']}", flush=True)
