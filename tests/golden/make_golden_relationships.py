"""Golden vectors for MetadataManager._build_similarity_relationships (core.py:3493-3531), produced by running
the UNMODIFIED reference method (sklearn cosine_similarity + argsort) in the authoring container.

    python tests/golden/make_golden_relationships.py     (needs /root/reference; writes relationships.npz/.json)

TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shims, synth  # noqa: E402


class _Log:
    def __getattr__(self, _):
        return lambda *a, **k: None


class _Self:
    def __init__(self):
        self.logger = _Log()
        self.similarity_graph = {}


def main():
    with ref_shims.reference_modules(names=("core",)) as mods:
        core = mods["core"]
        sizes = {"L01_V001": 57, "L01_V002": 1, "L02_V001": 23, "L02_V002": 140, "L03_V001": 12}
        feats, all_md = {}, {}
        for fi, (folder, n) in enumerate(sizes.items()):
            x = (synth.clip_like(n, 64, seed=500 + fi, n_centres=4) * np.float32(1.3)).astype(np.float32)   # un-normalised
            feats[folder] = x
            all_md[folder] = [core.KeyframeMetadata(folder_name=folder, image_name=f"{i:04d}", frame_id=i,
                                                    file_path=f"keyframes/{folder}/{i:04d}.jpg",
                                                    clip_features=(x[i] if (i % 11) != 5 else None))
                              for i in range(n)]
        me = _Self()
        core.MetadataManager._build_similarity_relationships(me, np.zeros((1, 64), np.float32), all_md)
        np.savez_compressed(os.path.join(HERE, "relationships.npz"), **feats)
        with open(os.path.join(HERE, "relationships.json"), "w") as f:
            json.dump({"missing_rule": "clip_features is None where frame_id % 11 == 5", "graph": me.similarity_graph}, f)
        print("frames in graph", len(me.similarity_graph), "edges", sum(len(v) for v in me.similarity_graph.values()))


if __name__ == "__main__":
    main()
