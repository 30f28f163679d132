"""Dedup kernel timing at a given size: python tools/dedup_one.py N D W [REPS]  (inputs resident in HBM)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ivr_b200 import _native as nat  # noqa: E402
from oracle import dedup as od, synth  # noqa: E402

n, d, w = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
xs = []
for s in range(0, n, 100_000):
    x, _ = synth.dedup_frames(min(100_000, n - s), d, seed=1000 + s)
    xs.append(torch.from_numpy(x))
x = torch.cat(xs).cuda()
cos = torch.empty(n, dtype=torch.float32, device="cuda")
mask = torch.empty(n, dtype=torch.int32, device="cuda")
keep = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
# pass 1: consecutive cosine -> scenes on the host (as FrameFilter does)
nat.check(nat.lib.ivr_consecutive_cosine_device(0, x.data_ptr(), n, d, cos.data_ptr(), st))
torch.cuda.synchronize()
scenes = od.scenes_from_cosines(cos.cpu().numpy()[1:], n, 0.75, 2)
a = torch.tensor([s for s, _ in scenes], dtype=torch.int64, device="cuda")
b = torch.tensor([e for _, e in scenes], dtype=torch.int64, device="cuda")
nat.check(nat.lib.ivr_dedup_set_timing(1))
ms = (C.c_float * 2)()
best = None
for _ in range(reps):
    nat.check(nat.lib.ivr_dedup_window_device(0, x.data_ptr(), n, d, a.data_ptr(), b.data_ptr(), len(scenes), w,
                                              C.c_float(0.95), keep.data_ptr(), cos.data_ptr(), mask.data_ptr(), st))
    torch.cuda.synchronize()
    nat.check(nat.lib.ivr_dedup_last_timing(ms))
    best = (ms[0], ms[1]) if best is None or ms[0] < best[0] else best
kept = int(keep.sum().item())
# parity on a prefix against the vectorised oracle
m = min(n, 20001)
want = od.window_keep_mask(x[:m].cpu().numpy(), [sc for sc in scenes if sc[1] < m], w, 0.95)
last = max([sc[1] for sc in scenes if sc[1] < m] + [0]) + 1
ok = bool(np.array_equal(keep[:last].cpu().numpy(), want[:last]))
print(f"dedup n={n} d={d} W={w}: banded_ms={best[0]:.3f} resolve_ms={best[1]:.3f} "
      f"GB/s={n*d*4/best[0]/1e6:.0f} frames/s={n/(best[0]+best[1])*1e3:.3e} kept={kept} scenes={len(scenes)} prefix_parity={ok}")
