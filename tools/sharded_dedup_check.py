"""Multi-GPU check + timing of the video-sharded dedup (run under torchrun, one rank per GPU):
whole videos per rank, no data-path collective (second row)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ivr_b200.sharded import ShardedFrameFilter  # noqa: E402
from oracle import dedup as od, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
os.environ["IVR_DEVICE"] = str(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rng = np.random.default_rng(5)
lens = rng.integers(200, 6000, size=120).tolist()                 # 120 videos, ~370 k frames
parts = [synth.dedup_frames(n, 512, seed=1000 + i)[0] for i, n in enumerate(lens)]
x = np.concatenate(parts)
starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
bounds = [(int(s), int(s + n - 1)) for s, n in zip(starts, lens)]
f = ShardedFrameFilter(window=8, threshold=0.95)
kept = f.filter_videos(x, bounds)
dist.barrier()
t0 = time.perf_counter()
kept = f.filter_videos(x, bounds)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
allk = f.gather(kept)
if rank == 0:
    lo, hi = f.my_units(bounds)
    want = []
    cfg = {"enable_similarity_filtering": True, "similarity_threshold": 0.95, "similarity_window_size": 8,
           "use_advanced_similarity_filtering": True, "min_frame_distance": 1}
    for (vs, ve), v in list(zip(bounds, parts))[:6]:               # the CPU rule on the first videos only (it is slow)
        sims = od.consecutive_cosines_fast(v)
        scenes = od.scenes_from_cosines(sims, len(v), 0.75, 2)
        want += [vs + int(i) for i in np.flatnonzero(od.window_keep_mask(v, scenes, 8, 0.95))]
    got = [i for i in allk if i <= bounds[5][1]]
    same = got == want
    print(f"sharded dedup x{world}: {len(x)} frames in {len(lens)} videos, {dt * 1e3:.1f} ms per pass (host frames in, kept lists out), "
          f"{len(x) / dt / 1e6:.2f} M frames/s, kept {len(allk)}; first 6 videos vs the CPU rule: {'identical' if same else 'DIFFER'}", flush=True)
dist.destroy_process_group()
