"""Golden vectors for filter_research_update.AdvancedKeyframeExtractor.detect_scene_changes (101-111),
.cluster_similar_frames (113-134) and .select_representative_frame (136-155), produced by running the UNMODIFIED reference with real scikit-learn.

    python tests/golden/make_golden_research.py        (needs /root/reference; writes research.npz/.json)

TEST INFRASTRUCTURE ONLY.  ``imagehash`` / ``colorama`` (absent, presentation and image hashing only) and
``transformers`` (network) are stubbed by oracle/ref_shims.py; no reference source is copied.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cluster as oc, ref_shims  # noqa: E402


def frames(n, d, seed):
    """Scenes around a base vector; inside a scene small groups of near-identical frames (sigma 0.02-0.08)
    next to looser ones (0.3-0.6), so DBSCAN at eps = 0.05 finds clusters, border points and noise."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        base = rng.standard_normal(d).astype(np.float32)
        for _ in range(int(rng.integers(2, 7))):
            anchor = base + np.float32(rng.uniform(0.3, 0.6)) * rng.standard_normal(d).astype(np.float32)
            for _ in range(int(rng.integers(1, 6))):
                out.append(anchor + np.float32(rng.uniform(0.02, 0.3)) * rng.standard_normal(d).astype(np.float32))
    return (np.stack(out[:n]) * np.float32(1.7)).astype(np.float32)


def guard_ok(x, eps=0.05, scene_thr=0.7, band=1e-4):
    sim = oc.cosine_matrix(x)
    off = ~np.eye(len(x), dtype=bool)
    cons = np.diag(sim, 1)
    return not (np.abs((1 - sim) - eps) < band)[off].any() and not (np.abs(cons - scene_thr) < band).any()


def main():
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):                          # the module prints a banner at import
        ctx = ref_shims.reference_modules(names=("filter_research_update",))
        mods = ctx.__enter__()
    try:
        fr = mods["filter_research_update"]
        ex = fr.AdvancedKeyframeExtractor.__new__(fr.AdvancedKeyframeExtractor)
        cases, arrays = {}, {}
        for name, n, d, seed in [("s384", 260, 384, 41), ("s64", 180, 64, 42), ("s512", 90, 512, 43)]:
            while True:
                x = frames(n, d, seed)
                if guard_ok(x):
                    break
                seed += 1000
            emb = [x[i] for i in range(n)]
            changes = ex.detect_scene_changes(emb)
            # With the scikit-learn of this image the reference method RAISES ("Negative values in data passed to
            # X") whenever rounding makes some 1 - cos(e_i, e_i) negative: DBSCAN(metric='precomputed') validates
            # non-negativity.  Such slices are recorded as "ValueError"; the others pin the clustering itself.
            clusters, raised = {}, 0
            spans = list(zip(changes[:-1], changes[1:]))
            spans += [(a, min(a + w, n)) for a in range(0, n - 3, 7) for w in (3, 4, 5, 6, 8, 12)]
            for a, b in spans:
                try:
                    clusters[f"{a}:{b}"] = [[int(i) for i in g] for g in ex.cluster_similar_frames(emb[a:b], list(range(b - a)))]
                except ValueError as e:
                    assert "Negative values" in str(e)
                    clusters[f"{a}:{b}"] = "ValueError"
                    raised += 1
            # Phase 3 (select_representative_frame, 136-155) on every group the reference produced
            reps = {}
            for span, groups in clusters.items():
                if groups == "ValueError":
                    continue
                a, b = (int(v) for v in span.split(":"))
                reps[span] = [int(ex.select_representative_frame(g, emb[a:b], None)) for g in groups]
            arrays[name] = x
            cases[name] = {"scene_changes": [int(c) for c in changes], "clusters": clusters, "representatives": reps,
                           "eps": fr.CLUSTER_EPS, "min_samples": fr.MIN_CLUSTER_SIZE, "scene_threshold": fr.SCENE_THRESHOLD}
            ok = [v for v in clusters.values() if v != "ValueError"]
            print(name, "scenes", len(changes) - 1, "slices", len(clusters), "raised", raised,
                  "multi-member groups", sum(1 for v in ok for g in v if len(g) > 1), file=sys.stderr)
        cases["edge"] = {"empty": ex.cluster_similar_frames([], []), "one": ex.cluster_similar_frames([arrays["s64"][0]], [0])}
        np.savez_compressed(os.path.join(HERE, "research.npz"), **arrays)
        with open(os.path.join(HERE, "research.json"), "w") as f:
            json.dump(cases, f)
    finally:
        with contextlib.redirect_stdout(sink):
            ctx.__exit__(None, None, None)


if __name__ == "__main__":
    main()
