"""Golden vectors for TemporalAnalyzer.find_similar_sequences / detect_scene_boundaries, produced by
running the UNMODIFIED reference (core.py:3584-3702, 3812-3832) in the authoring container.

    python tests/golden/make_golden_temporal.py        (needs /root/reference; writes temporal.npz/.json)

TEST INFRASTRUCTURE ONLY.  The committed fixtures are what travels; /root/reference is never read at test time.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shims  # noqa: E402


class _Log:
    def __getattr__(self, _):
        return lambda *a, **k: None


class _Timer:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Perf:
    def timer(self, *a, **k):
        return _Timer()


def frames(n, d, seed):
    """Scenes of geometric length around a per-scene base vector (un-normalised fp32, like raw CLIP output)."""
    rng = np.random.default_rng(seed)
    out, sid = [], []
    s = 0
    while len(out) < n:
        base = rng.standard_normal(d).astype(np.float32)
        for _ in range(int(rng.geometric(1.0 / 12))):
            out.append(base + np.float32(rng.uniform(0.15, 0.5)) * rng.standard_normal(d).astype(np.float32))
            sid.append(s)
        s += 1
    return (np.stack(out[:n]) * np.float32(2.0)).astype(np.float32)


def main():
    with ref_shims.reference_modules(names=("core",)) as mods:
        core = mods["core"]
        ta = core.TemporalAnalyzer.__new__(core.TemporalAnalyzer)
        ta.config, ta.logger, ta.perf_monitor = None, _Log(), _Perf()
        cases, arrays = {}, {}
        for name, n, d, seed, t0, tl, L, thr in [("a", 300, 64, 11, 40, 14, 5, 0.8),
                                                 ("b", 220, 128, 12, 100, 9, 3, 0.7),
                                                 ("c", 150, 64, 13, 10, 6, 6, 0.6)]:
            db = frames(n, d, seed)
            rng = np.random.default_rng(seed + 100)
            target = (db[t0:t0 + tl] + np.float32(0.2) * rng.standard_normal((tl, d)).astype(np.float32)).astype(np.float32)
            hits = ta.find_similar_sequences(target, db, sequence_length=L, similarity_threshold=thr)
            sims = np.array([h[1] for h in hits], np.float64)
            # guard band: no qualifying/non-qualifying window within 1e-4 of the threshold is checked by the test's oracle
            arrays[f"{name}_db"], arrays[f"{name}_target"] = db, target
            cases[name] = {"sequence_length": L, "threshold": thr,
                           "hits": [[int(h[0]), float(h[1])] for h in hits],
                           "scene_boundaries": [[int(a), int(b)] for a, b in
                                                ta.detect_scene_boundaries(db, threshold=0.3, min_scene_length=5)]}
            print(name, "hits", len(hits), "best", sims[:3] if len(sims) else None,
                  "scenes", len(cases[name]["scene_boundaries"]))
        # validate_inputs=False skips the "fewer than 2 * min_scene_length frames -> one scene" shortcut (core.py:3601-3608):
        # every 9-frame slice of db "a" (min_scene_length 4 -> 9 >= 2 * 4 would not take the shortcut anyway; 5 does)
        db = arrays["a_db"]
        cases["short_clip"] = {"min_scene_length": 5, "threshold": 0.3, "length": 9, "slices": []}
        for s0 in range(0, 120):
            sl = db[s0:s0 + 9]
            cases["short_clip"]["slices"].append({
                "start": s0,
                "validated": [[int(a), int(b)] for a, b in ta.detect_scene_boundaries(sl, 0.3, 5, validate_inputs=True)],
                "unvalidated": [[int(a), int(b)] for a, b in ta.detect_scene_boundaries(sl, 0.3, 5, validate_inputs=False)]})
        print("short_clip: slices whose unvalidated answer differs:",
              sum(c["validated"] != c["unvalidated"] for c in cases["short_clip"]["slices"]))
        cases["short"] = {"hits": [[int(h[0]), float(h[1])] for h in
                                   ta.find_similar_sequences(arrays["a_target"][:3], arrays["a_db"], sequence_length=5)]}
        np.savez_compressed(os.path.join(HERE, "temporal.npz"), **arrays)
        with open(os.path.join(HERE, "temporal.json"), "w") as f:
            json.dump(cases, f)


if __name__ == "__main__":
    main()
