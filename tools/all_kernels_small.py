"""Small problems through every kernel in one process (a quick all-kernels exerciser; also the input for a
compute-sanitizer pass where that tool is available -- it is closed on the measurement pool).

    python tools/all_kernels_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402
from ivr_b200 import frame_filter as ff  # noqa: E402
from oracle import synth  # noqa: E402

xb = synth.clip_like(3000, 128, seed=1, n_centres=32)
idx = ivr_b200.IndexFlatIP(128)
idx.add(xb)
for mode, path, nq, k in [("0", 1, 1, 10), ("0", 1, 3, 200), ("3", 2, 5, 10), ("3", 2, 40, 100), ("1", 2, 70, 100),
                          ("2", 2, 300, 100), ("1", 2, 20, 300), ("2", 2, 20, 300)]:
    os.environ["IVR_MMA_MODE"] = mode
    idx.search_path = path
    D, I = idx.search(synth.clip_like(nq, 128, seed=2 + nq, n_centres=32), k)
    print("search", mode, path, nq, k, idx.last_timing()["kernel"], int(I[0, 0]), flush=True)
os.environ["IVR_MMA_MODE"] = "0"
xw = synth.clip_like(2000, 768, seed=3, n_centres=16)
iw = ivr_b200.IndexFlatIP(768)
iw.add(xw)
for nq in (4, 300):
    D, I = iw.search(synth.clip_like(nq, 768, seed=4, n_centres=16), 50)
    print("search 768", nq, iw.last_timing()["kernel"], flush=True)
x, _ = synth.dedup_frames_guarded(1500, 512, window=8, thresholds=(0.95, 0.75), seed=5)
print("dedup window", len(ff.FrameFilter(window=8, threshold=0.95).apply_filters(x)), flush=True)
print("dedup fifo", len(ff.temporal_window_filter(x[:400], threshold=0.95, temporal_window=10)), flush=True)
print("dedup chain", len(ff.extract_unique_frames_rule(x[:400], threshold=0.98)), flush=True)
x64, _ = synth.dedup_frames_guarded(700, 64, window=5, thresholds=(0.9,), seed=6)
print("dedup generic", len(ff.FrameFilter(window=5, threshold=0.9).apply_filters(x64)), flush=True)
ta = ivr_b200.TemporalAnalyzer()
print("sequences", len(ta.find_similar_sequences(x64[100:120], x64, sequence_length=5, similarity_threshold=0.6)), flush=True)
print("sanitize run complete")
