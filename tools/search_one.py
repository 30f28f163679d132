"""One search configuration, for ncu captures: python tools/search_one.py N D NQ K PATH [CG] [REPS]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402

n, d, nq, k, path = (int(a) for a in sys.argv[1:6])
cg = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
os.environ["IVR_MMA_CTA_GROUP"] = str(cg)
g = torch.Generator(device="cuda").manual_seed(1)
idx = ivr_b200.IndexFlatIP(d)
idx.reserve(n)
cen = torch.nn.functional.normalize(torch.randn(4096, d, generator=g, device="cuda"), dim=1)
for s in range(0, n, 500_000):
    m = min(500_000, n - s)
    z = torch.randint(0, 4096, (m,), generator=g, device="cuda")
    x = torch.nn.functional.normalize(cen[z] + (0.5 / d ** 0.5) * torch.randn(m, d, generator=g, device="cuda"), dim=1)
    idx.add(x)
z = torch.randint(0, 4096, (nq,), generator=g, device="cuda")
q = torch.nn.functional.normalize(cen[z] + (0.5 / d ** 0.5) * torch.randn(nq, d, generator=g, device="cuda"), dim=1)
idx.search_path = path
idx.set_timing(True)
for _ in range(reps):
    D, I = idx.search_tensor(q, k)
    torch.cuda.synchronize()
    t = idx.last_timing()
print(f"n={n} d={d} nq={nq} k={k} path={t['path']} cg={cg}: score_ms={t['score_ms']:.3f} merge_ms={t['merge_ms']:.3f} "
      f"prep_ms={t['prep_ms']:.3f} TFLOPs={2.0*n*d*nq/t['score_ms']/1e9:.1f} GB/s(rows once)={n*d*2/t['score_ms']/1e6:.0f}")
