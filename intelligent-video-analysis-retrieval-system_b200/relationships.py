"""Per-folder similarity relationships on the B200 path ("next" row 2).

Mirrors ``MetadataManager._build_similarity_relationships`` (core.py:3493-3531): for every folder
(video) the reference computes the full cosine matrix of its keyframes with sklearn and keeps, for
each frame, the top-10 other frames whose cosine exceeds 0.7.  That is the batched search kernel
with Q = X: the folder's rows are L2-normalised, added to a flat inner-product index and searched
against themselves with k = 11 (the fused top-k epilogue never materialises the n x n matrix).

Reference quirk kept: the first entry of the descending order is dropped as "self" whatever it is
(``np.argsort(sim[i])[::-1][1:11]``), so an exact duplicate may take self's place.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import faiss_compat as faiss


def build_similarity_relationships(all_metadata: Dict[str, List], top: int = 10, threshold: float = 0.7,
                                   device: int | None = None) -> Dict[str, List[str]]:
    """{frame key: [keys of its most similar frames in the same folder]} (core.py:3493-3531).

    ``all_metadata``: {folder_name: [KeyframeMetadata-like with .clip_features, .get_unique_key()]}.
    Frames without features are skipped; folders with fewer than 2 featured frames are skipped.
    """
    graph: Dict[str, List[str]] = {}
    index = None
    try:
        for _folder, metadata_list in all_metadata.items():
            feats, keys = [], []
            for m in metadata_list:
                if m.clip_features is not None:
                    feats.append(np.asarray(m.clip_features, dtype=np.float32).reshape(-1))
                    keys.append(m.get_unique_key())
            if len(feats) < 2:
                continue
            x = np.ascontiguousarray(np.stack(feats), dtype=np.float32)
            faiss.normalize_L2(x)                                   # cosine = inner product of unit rows
            if index is None or index.d != x.shape[1]:
                if index is not None:
                    index.close()
                index = faiss.IndexFlatIP(x.shape[1], device=device)
            index.reset()
            index.add(x)
            k = min(top + 1, len(keys))
            D, I = index.search(x, k)
            for i, key in enumerate(keys):
                graph[key] = [keys[j] for s, j in zip(D[i, 1:], I[i, 1:]) if j >= 0 and s > threshold]
    finally:
        if index is not None:
            index.close()
    return graph
