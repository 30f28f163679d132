"""Host-to-host timings of the SURVEY 8f "next" components on one GPU, with the CPU oracle beside them on a bounded slice.

    python tools/next_kernels_timing.py > gpurun_out/next_kernels.md
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402
from ivr_b200 import frame_filter as ff  # noqa: E402
from oracle import dedup as od, synth, temporal as ot  # noqa: E402


def best(fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


rows = []
rng = np.random.default_rng(0)

# ---- TemporalAnalyzer.find_similar_sequences (core.py:3644-3702): K6 seq_cosine + seq_diag ----
ta = ivr_b200.TemporalAnalyzer()
for nt, nd, d, L in ((64, 20_000, 512, 5), (200, 100_000, 512, 5), (500, 200_000, 768, 8)):
    db = synth.dedup_frames(nd, d, seed=1)[0]
    tg = db[1000:1000 + nt] + np.float32(0.1) * rng.standard_normal((nt, d)).astype(np.float32)
    t, hits = best(lambda: ta.find_similar_sequences(tg, db, sequence_length=L, similarity_threshold=0.8), reps=3)
    flops = 2.0 * nt * nd * d
    m = min(nd, 2000)
    t0 = time.perf_counter()
    ot.find_similar_sequences(tg[:16], db[:m], L, 0.8)
    t_cpu = (time.perf_counter() - t0) * (nt / 16) * (nd / m)
    rows.append(("find_similar_sequences", f"target {nt} x db {nd} x {d}, L={L}", f"{t * 1e3:.2f} ms host-to-host",
                 f"{flops / t / 1e12:.2f} TFLOP/s fp32 SIMT (incl. H2D of {4 * (nt + nd) * d / 1e6:.0f} MB), {len(hits)} hits",
                 f"oracle (line-by-line port, 1 thread) ~{t_cpu:.0f} s extrapolated from a 16 x {m} slice"))

# ---- cluster_similar_frames (filter_research_update.py:113-134): cosine matrix -> eps-neighbourhood bits + host DBSCAN ----
for n, d in ((1024, 384), (4096, 384), (8192, 512)):
    base = rng.standard_normal((max(n // 50, 2), d)).astype(np.float32)
    x = (base[rng.integers(0, len(base), n)] + np.float32(0.12) * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    t, cl = best(lambda: ff.cluster_similar_frames(x, eps=0.05, min_samples=2), reps=3)
    rows.append(("cluster_similar_frames", f"{n} frames x {d}", f"{t * 1e3:.2f} ms host-to-host",
                 f"{2.0 * n * n * d / t / 1e12:.2f} TFLOP/s fp32 SIMT incl. host DBSCAN labelling, {len(cl)} clusters", ""))

# ---- last-kept chain / FIFO-of-10 (filter.py:178-222; filter_research_update.py:316-338): one warp per scene ----
for n, d in ((200_000, 512), (1_000_000, 512)):
    x = synth.dedup_frames(n, d, seed=3)[0] if n <= 200_000 else np.concatenate([synth.dedup_frames(200_000, d, seed=3 + i)[0] for i in range(n // 200_000)])
    sims = np.asarray(ff.calculate_similarities(x), np.float32)
    scenes = od.scenes_from_cosines(sims, n, 0.75, 2)
    cfg = ff.create_config()
    t, out = best(lambda: ff.apply_similarity_filtering_to_scenes(x, list(range(n)), scenes, cfg), reps=2)
    t2, out2 = best(lambda: ff.temporal_window_filter(x[:200_000], 0.95, 10), reps=2)
    rows.append(("filter_similar_frames_in_scene (chain)", f"{n} frames x {d}, {len(scenes)} scenes",
                 f"{t * 1e3:.1f} ms host-to-host (H2D of {n * d * 4 / 1e9:.2f} GB inside)", f"{n / t / 1e6:.2f} M frames/s, kept {len(out[1])}", ""))
    rows.append(("temporal_window_filter (FIFO of 10)", f"200000 frames x {d}, one sequence", f"{t2 * 1e3:.1f} ms host-to-host",
                 f"{200_000 / t2 / 1e6:.2f} M frames/s, kept {len(out2)} (one warp walks the whole sequence: serial by definition)", ""))

# ---- MetadataManager._build_similarity_relationships (core.py:3493-3531): per-folder search with Q = X + fp32 re-rank ----
for folders, per, d in ((20, 500, 512), (50, 2000, 512)):
    meta = {}
    for f in range(folders):
        x = synth.clip_like(per, d, seed=100 + f, n_centres=8)
        meta[f"L{f:02d}"] = [ivr_b200.KeyframeMetadata(folder_name=f"L{f:02d}", image_name=f"{i:05d}", frame_id=i,
                                                       file_path=f"L{f:02d}/{i:05d}.jpg", clip_features=x[i]) for i in range(per)]
    t, g = best(lambda: ivr_b200.build_similarity_relationships(meta), reps=2, warm=1)
    rows.append(("build_similarity_relationships", f"{folders} folders x {per} frames x {d}", f"{t * 1e3:.0f} ms host-to-host",
                 f"{folders * per / t / 1e3:.1f} k frames/s (GPU candidates + fp32 host re-rank), {sum(len(v) for v in g.values())} edges", ""))

print("| component | shape | time | rate | CPU oracle beside it |\n|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
